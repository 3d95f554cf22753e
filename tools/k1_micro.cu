// Micro-benchmark behind K1's staging design (not part of the library): how fast can one B200 pull N 16 KB
// heat-maps into shared memory, for different staging shapes and amounts of per-map work?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/k1_micro tools/k1_micro.cu && tools/k1_micro
// Variants (all: persistent grid of 148 CTAs, maps claimed from a global counter one ahead):
//   stage   W warps/CTA, one 16 KB buffer per warp, bulk copy -> wait -> touch one word
//   scan    + pass A (max / min over the whole staged map, 128-bit LDS)
//   spin    + a dependent ALU chain of `work` instructions per map after the scan (stands for phases L/B/C)
//   ldg     no staging: each warp streams its map with 128-bit LDG (8 in flight per lane), 32 warps/CTA
//   pair    2 warps share a buffer: scan, then the buffer goes to the partner while this warp does its `work`
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

constexpr int kMap = 4096;            // floats per map

__device__ __forceinline__ float scan(const float* s, int lane) {
  const float4* s4 = reinterpret_cast<const float4*>(s);
  float m = -1e30f, n = 1e30f;
  for (int q = lane; q < kMap / 4; q += 256) {
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = s4[q + 32 * u];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      m = fmaxf(fmaxf(m, fmaxf(x[u].x, x[u].y)), fmaxf(x[u].z, x[u].w));
      n = fminf(fminf(n, fminf(x[u].x, x[u].y)), fminf(x[u].z, x[u].w));
    }
  }
  return m - n;
}

__device__ __forceinline__ float spin(float v, int work) {
  // a dependent chain: ~work FFMA at 4 cycles each
  for (int i = 0; i < work; ++i) v = __fmaf_rn(v, 1.0000001f, 1e-7f);
  return v;
}

// mode 0 stage, 1 scan, 2 scan + spin
__global__ void __launch_bounds__(512, 1) k_stage(const float* maps, int N, unsigned long long* counter, float* out, int mode,
                                                  int work, int use_policy, int cap) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* buf = reinterpret_cast<float*>(smem + (size_t)warp * kMap * 4);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)warps * kMap * 4) + warp;
  int* credits = reinterpret_cast<int*>(smem + (size_t)warps * kMap * 4 + 192);     // copies this CTA may still put in flight
  uint64_t pol = 0;
  if (threadIdx.x == 0) *credits = cap;
  if (lane == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); pol = evict_first(); }
  __syncthreads();
  unsigned long long claim = 0;
  auto issue = [&](long long n) {
    if (lane == 0) {
      if (cap > 0) {
        for (;;) {
          if (atomicSub(credits, 1) > 0) break;
          atomicAdd(credits, 1);
          __nanosleep(100);
        }
      }
      mbar_expect(bar, kMap * 4);
      if (use_policy) bulk_g2s(buf, maps + (size_t)n * kMap, kMap * 4, bar, pol);
      else asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(smem_u32(buf)), "l"(maps + (size_t)n * kMap), "r"(kMap * 4), "r"(smem_u32(bar)) : "memory");
    }
  };
  if (lane == 0) claim = atomicAdd(counter, 1ull);
  long long cur = (long long)__shfl_sync(0xffffffffu, claim, 0);
  if (cur < N) issue(cur);
  if (lane == 0) asm volatile("atom.add.relaxed.gpu.global.u64 %0, [%1], 1;" : "=l"(claim) : "l"(counter) : "memory");
  float acc = 0.f;
  for (long long it = 0; cur < N; ++it) {
    mbar_wait(bar, (uint32_t)(it & 1));
    if (cap > 0 && lane == 0) atomicAdd(credits, 1);
    float v = buf[lane];
    if (mode >= 1) v = scan(buf, lane);
    if (mode >= 2) v = spin(v, work);
    acc += v;
    __syncwarp();
    const long long nn = (long long)__shfl_sync(0xffffffffu, claim, 0);
    if (nn < N) {
      issue(nn);
      if (lane == 0) asm volatile("atom.add.relaxed.gpu.global.u64 %0, [%1], 1;" : "=l"(claim) : "l"(counter) : "memory");
    }
    cur = nn;
  }
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

// two warps share one buffer: scan on the buffer, then hand it to the partner and do the `work` chain
__global__ void __launch_bounds__(1024, 1) k_pair(const float* maps, int N, unsigned long long* counter, float* out, int work,
                                                   int nbuf) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = warp >> 1, side = warp & 1;
  float* buf = reinterpret_cast<float*>(smem + (size_t)b * kMap * 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nbuf * kMap * 4);     // full[nbuf], free[nbuf]
  uint64_t* full = bars + b;
  uint64_t* freeb = bars + nbuf + b;
  volatile int* flags = reinterpret_cast<volatile int*>(bars + 2 * nbuf);          // fills[nbuf], gone[nbuf][2]
  volatile int* fills = flags + b;                                                  // fills issued into this buffer
  volatile int* gone = flags + nbuf + 2 * b;                                        // [2]: this side has left
  uint64_t pol = 0;
  if (side == 0 && lane == 0) {
    mbar_init(full, 1); mbar_init(freeb, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    *fills = 0; gone[0] = 0; gone[1] = 0;
  }
  if (lane == 0) pol = evict_first();
  __syncthreads();
  unsigned long long claim = 0;
  if (lane == 0) claim = atomicAdd(counter, 1ull);
  long long cur = (long long)__shfl_sync(0xffffffffu, claim, 0);
  float acc = 0.f;
  // ownership k of the buffer (k = 0, 1, 2, ...) belongs to side k & 1 and ends with an arrive on `freeb` (phase k);
  // the warp's i-th ownership is k = 2 i + side and waits for phase k - 1 -- unless the partner has left
  for (long long i = 0;; ++i) {
    const long long k = 2 * i + side;
    if (cur >= N) {
      if (lane == 0) { gone[side] = 1; __threadfence_block(); }
      break;
    }
    if (k > 0 && lane == 0) {
      uint32_t ok = 0;
      while (!ok && !gone[side ^ 1]) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(smem_u32(freeb)), "r"((uint32_t)((k - 1) & 1)) : "memory");
      }
    }
    __syncwarp();
    int f = 0;
    if (lane == 0) {
      f = *fills; *fills = f + 1;
      mbar_expect(full, kMap * 4);
      bulk_g2s(buf, maps + (size_t)cur * kMap, kMap * 4, full, pol);
      asm volatile("atom.add.relaxed.gpu.global.u64 %0, [%1], 1;" : "=l"(claim) : "l"(counter) : "memory");
    }
    f = __shfl_sync(0xffffffffu, f, 0);
    mbar_wait(full, (uint32_t)(f & 1));
    float v = scan(buf, lane);
    __syncwarp();
    if (lane == 0) { __threadfence_block(); mbar_arrive(freeb); }
    v = spin(v, work);
    acc += v;
    cur = (long long)__shfl_sync(0xffffffffu, claim, 0);
  }
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

__global__ void __launch_bounds__(1024, 1) k_ldg(const float* maps, int N, unsigned long long* counter, float* out, int work) {
  const int lane = threadIdx.x & 31;
  unsigned long long claim = 0;
  if (lane == 0) claim = atomicAdd(counter, 1ull);
  long long cur = (long long)__shfl_sync(0xffffffffu, claim, 0);
  float acc = 0.f;
  while (cur < N) {
    if (lane == 0) asm volatile("atom.add.relaxed.gpu.global.u64 %0, [%1], 1;" : "=l"(claim) : "l"(counter) : "memory");
    const float4* s4 = reinterpret_cast<const float4*>(maps + (size_t)cur * kMap);
    float m = -1e30f;
    for (int q = lane; q < kMap / 4; q += 256) {
      float4 x[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(s4 + q + 32 * u));
#pragma unroll
      for (int u = 0; u < 8; ++u) m = fmaxf(fmaxf(m, fmaxf(x[u].x, x[u].y)), fmaxf(x[u].z, x[u].w));
    }
    acc += spin(m, work);
    cur = (long long)__shfl_sync(0xffffffffu, claim, 0);
  }
  if (acc == 12345.678f) out[threadIdx.x] = acc;
}

int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 28672;
  float *maps, *out, *flush;
  unsigned long long* counter;
  CK(cudaMalloc(&maps, (size_t)N * kMap * 4));
  CK(cudaMalloc(&out, 4096));
  CK(cudaMalloc(&counter, 8));
  const size_t flush_bytes = 256u << 20;
  CK(cudaMalloc(&flush, flush_bytes));
  {   // pseudo-random floats (not zeros: the data must not flatter any lossless path in the memory system)
    float* h = (float*)malloc((size_t)64 << 20);
    uint32_t x = 12345u;
    for (size_t i = 0; i < ((size_t)64 << 20) / 4; ++i) { x = x * 1664525u + 1013904223u; h[i] = (float)(x >> 8) * (1.0f / 16777216.0f) - 0.5f; }
    for (size_t off = 0; off < (size_t)N * kMap * 4; off += (size_t)64 << 20) {
      size_t n = (size_t)N * kMap * 4 - off; if (n > ((size_t)64 << 20)) n = (size_t)64 << 20;
      CK(cudaMemcpy((char*)maps + off, h, n, cudaMemcpyHostToDevice));
    }
    free(h);
  }
  int sms = 0, optin = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0));
  CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  CK(cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const double mb = (double)N * kMap * 4 / 1e6;
  auto time = [&](const char* name, auto launch) {
    float best = 1e9f, sum = 0.f;
    const int reps = 12;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaMemsetAsync(flush, r, flush_bytes));
      CK(cudaMemsetAsync(counter, 0, 8));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r >= 2) { sum += ms; if (ms < best) best = ms; }
    }
    printf("%-44s best %7.1f us (%5.0f GB/s)  mean %7.1f us (%5.0f GB/s)\n", name, best * 1e3, mb / best, sum / reps * 1e3,
           mb / (sum / reps));
    fflush(stdout);
  };
  char name[128];
  for (int pol = 1; pol >= 0; --pol)
    for (int w : {14, 7, 4}) {
      snprintf(name, sizeof name, "stage  %2d warps x 16 KB%s", w, pol ? "" : " (no evict_first)");
      time(name, [&] { k_stage<<<sms, w * 32, (size_t)w * kMap * 4 + 256, 0>>>(maps, N, counter, out, 0, 0, pol, 0); });
    }
  time("scan   14 warps x 16 KB", [&] { k_stage<<<sms, 14 * 32, (size_t)14 * kMap * 4 + 256, 0>>>(maps, N, counter, out, 1, 0, 1, 0); });
  for (int work : {250, 500, 750, 1000, 1500}) {
    snprintf(name, sizeof name, "spin   14 warps x 16 KB, work %4d", work);
    time(name, [&] { k_stage<<<sms, 14 * 32, (size_t)14 * kMap * 4 + 256, 0>>>(maps, N, counter, out, 2, work, 1, 0); });
  }
  for (int cap : {3, 4, 5, 6, 8})
    for (int work : {0, 500, 1000, 1500}) {
      snprintf(name, sizeof name, "cap    14 warps, <= %d copies in flight, work %4d", cap, work);
      time(name, [&] { k_stage<<<sms, 14 * 32, (size_t)14 * kMap * 4 + 256, 0>>>(maps, N, counter, out, work ? 2 : 1, work, 1, cap); });
    }
  for (int nbuf : {8, 10})
    for (int work : {500, 1000, 1500}) {
      snprintf(name, sizeof name, "pair   %2d warps / %2d buffers, work %4d", 2 * nbuf, nbuf, work);
      time(name, [&] { k_pair<<<sms, 2 * nbuf * 32, (size_t)nbuf * kMap * 4 + 1024, 0>>>(maps, N, counter, out, work, nbuf); });
    }
  for (int work : {0, 500, 1000}) {
    snprintf(name, sizeof name, "ldg    32 warps, no staging, work %4d", work);
    time(name, [&] { k_ldg<<<sms, 1024, 0, 0>>>(maps, N, counter, out, work); });
  }
  return 0;
}
