// Calibration for K3 (not part of the library): what does one B200 deliver for a streaming kernel that reads R
// arrays and writes W arrays of `mb` megabytes each (128-bit accesses, 8 in flight per thread, L1::no_allocate)?
// K3 is R = 2 (two student stacks), W = 3 (two gradients + the target) on 58.7 MB arrays.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/rw_micro tools/rw_micro.cu && tools/rw_micro
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

struct Ptrs { const float4* r[3]; float4* w[3]; };

// blocks of 1024 float4 (16 KB, one heat-map) per CTA iteration, like K3's items
template <int R, int W>
__global__ void __launch_bounds__(128) rw_kernel(Ptrs P, long long n_blocks) {
  for (long long blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
    float4 acc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) acc[u] = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = ldg_stream(P.r[r] + blk * 1024 + threadIdx.x + 128 * u);
#pragma unroll
      for (int u = 0; u < 8; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
#pragma unroll
    for (int w = 0; w < W; ++w)
#pragma unroll
      for (int u = 0; u < 8; ++u) stg_stream(P.w[w] + blk * 1024 + threadIdx.x + 128 * u, acc[u]);
    if (W == 0 && acc[0].x == 123.456f) P.w[0][0] = acc[0];
  }
}

int main(int argc, char** argv) {
  const long long n_blocks = argc > 1 ? atoll(argv[1]) : 3584;     // 16 KB blocks per array (c2: B*J = 3584)
  const size_t bytes = (size_t)n_blocks * 16384;
  Ptrs P;
  float* buf[6];
  for (int i = 0; i < 6; ++i) { CK(cudaMalloc(&buf[i], bytes)); CK(cudaMemset(buf[i], i + 1, bytes)); }
  for (int i = 0; i < 3; ++i) { P.r[i] = (const float4*)buf[i]; P.w[i] = (float4*)buf[3 + i]; }
  float* flush; const size_t fb = 256u << 20; CK(cudaMalloc(&flush, fb));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto run = [&](const char* name, int nr, int nw, auto kern) {
    for (int ctas : {4, 5, 8, 12, 16}) {
      float best = 1e9f, sum = 0.f;
      for (int rep = 0; rep < 12; ++rep) {
        CK(cudaMemsetAsync(flush, rep, fb));
        CK(cudaEventRecord(e0));
        kern<<<sms * ctas, 128>>>(P, n_blocks);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep >= 2) { sum += ms; if (ms < best) best = ms; }
      }
      const double mb = (double)bytes * (nr + nw) / 1e6;
      printf("%-16s %2d CTAs/SM: best %6.1f us (%5.0f GB/s)  mean %6.1f us (%5.0f GB/s)\n", name, ctas, best * 1e3, mb / best, sum / 10 * 1e3, mb / (sum / 10));
    }
  };
  run("read 1", 1, 0, rw_kernel<1, 0>);
  run("read 2", 2, 0, rw_kernel<2, 0>);
  run("write 1", 0, 1, rw_kernel<0, 1>);
  run("write 3", 0, 3, rw_kernel<0, 3>);
  run("read 1 write 1", 1, 1, rw_kernel<1, 1>);
  run("read 2 write 3", 2, 3, rw_kernel<2, 3>);
  run("read 2 write 2", 2, 2, rw_kernel<2, 2>);
  return 0;
}
