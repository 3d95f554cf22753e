// K4: mean-teacher EMA over every parameter tensor of a model in ONE launch.
//
// Reference: update_ema_variables, utils/parameters.py:4-8 (duplicate utils/udaap/utils_mt.py:34-39):
//   for ema_param, param in zip(ema_model.parameters(), model.parameters()):
//       ema_param.data.mul_(alpha).add_(param.data, alpha=1 - alpha)
// i.e. 2 launches per tensor (908 for a 2-stack hourglass).  Here a chunk table maps CTAs onto
// (tensor, offset) work items; 12 bytes of HBM traffic per parameter, 128-bit accesses where the
// chunk is 16-byte aligned.  Rounding follows ATen: t = ema*alpha (rounded), fma(param, 1-alpha, t).
#include "common.cuh"
#include <stdlib.h>

namespace ubpl {

__global__ void __launch_bounds__(256) ema_multi_kernel(const uint64_t* __restrict__ ema_ptrs,
                                                         const uint64_t* __restrict__ param_ptrs,
                                                         const long long* __restrict__ numels,
                                                         const int32_t* __restrict__ chunk_tensor,
                                                         const long long* __restrict__ chunk_start,
                                                         long long n_chunks, int chunk_elems, float a, float oma,
                                                         const float* __restrict__ alpha_dev) {
  // alpha_dev (optional): {alpha, 1 - alpha} read from device memory, so that a CUDA graph that captured this
  // launch follows the per-epoch alpha = min(1 - 1/(epo+1), ema_decay) of utils/parameters.py:6 without re-capture
  if (alpha_dev) { a = alpha_dev[0]; oma = alpha_dev[1]; }
  for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
    const int t = chunk_tensor[c];
    const long long start = chunk_start[c];
    float* e = reinterpret_cast<float*>(ema_ptrs[t]) + start;
    const float* p = reinterpret_cast<const float*>(param_ptrs[t]) + start;
    long long rem = numels[t] - start;
    const int n = (int)(rem < chunk_elems ? rem : chunk_elems);
    if ((((uintptr_t)e | (uintptr_t)p) & 15) == 0) {
      const int n4 = n >> 2;
      float4* e4 = reinterpret_cast<float4*>(e);
      const float4* p4 = reinterpret_cast<const float4*>(p);
      int i = threadIdx.x;
      for (; i + 3 * (int)blockDim.x < n4; i += 4 * blockDim.x) {      // 8 independent 128-bit loads in flight
        float4 ev[4], pv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { ev[u] = e4[i + u * blockDim.x]; pv[u] = ldg_stream(p4 + i + u * blockDim.x); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ev[u].x = ema1(ev[u].x, pv[u].x, a, oma); ev[u].y = ema1(ev[u].y, pv[u].y, a, oma);
          ev[u].z = ema1(ev[u].z, pv[u].z, a, oma); ev[u].w = ema1(ev[u].w, pv[u].w, a, oma);
          e4[i + u * blockDim.x] = ev[u];
        }
      }
      for (; i < n4; i += blockDim.x) {
        float4 ev = e4[i];
        const float4 pv = __ldg(p4 + i);
        ev.x = ema1(ev.x, pv.x, a, oma); ev.y = ema1(ev.y, pv.y, a, oma);
        ev.z = ema1(ev.z, pv.z, a, oma); ev.w = ema1(ev.w, pv.w, a, oma);
        e4[i] = ev;
      }
      for (int i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) e[i] = ema1(e[i], __ldg(p + i), a, oma);
    } else {
      for (int i = threadIdx.x; i < n; i += blockDim.x) e[i] = ema1(e[i], __ldg(p + i), a, oma);
    }
  }
}

__global__ void __launch_bounds__(256) ema_flat_kernel(float* __restrict__ e, const float* __restrict__ p, long long n,
                                                        float a, float oma) {
  const long long n4 = n >> 2;
  float4* e4 = reinterpret_cast<float4*>(e);
  const float4* p4 = reinterpret_cast<const float4*>(p);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 ev = e4[i];
    const float4 pv = __ldg(p4 + i);
    ev.x = ema1(ev.x, pv.x, a, oma); ev.y = ema1(ev.y, pv.y, a, oma);
    ev.z = ema1(ev.z, pv.z, a, oma); ev.w = ema1(ev.w, pv.w, a, oma);
    e4[i] = ev;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    e[i] = ema1(e[i], __ldg(p + i), a, oma);
}

}  // namespace ubpl

using namespace ubpl;

extern "C" int ubpl_ema_multi_tensor(const uint64_t* ema_ptrs, const uint64_t* param_ptrs, const int64_t* numels,
                                     const int32_t* chunk_tensor, const int64_t* chunk_start, int64_t n_chunks,
                                     int chunk_elems, float alpha, float one_minus_alpha, const float* alpha_dev,
                                     void* stream) {
  UBPL_REQUIRE(ema_ptrs && param_ptrs && numels && chunk_tensor && chunk_start, "ubpl_ema_multi_tensor: NULL pointer");
  UBPL_REQUIRE(n_chunks >= 0 && chunk_elems > 0 && chunk_elems % 4 == 0, "ubpl_ema_multi_tensor: bad chunking");
  if (n_chunks == 0) return UBPL_OK;
  // resident CTAs per SM: 8 by default; UBPL_EMA_CTAS lowers it (the update usually runs beside K1, where fewer,
  // longer streams disturb the staged copies less)
  const char* ev = getenv("UBPL_EMA_CTAS");
  const int per_sm = (ev && atoi(ev) > 0) ? atoi(ev) : 8;
  const long long cap = (long long)sm_count() * per_sm;
  const int grid = (int)(n_chunks < cap ? n_chunks : cap);
  ema_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ema_ptrs, param_ptrs, reinterpret_cast<const long long*>(numels),
                                                           chunk_tensor, reinterpret_cast<const long long*>(chunk_start),
                                                           n_chunks, chunk_elems, alpha, one_minus_alpha, alpha_dev);
  return check_launch("ubpl_ema_multi_tensor");
}

extern "C" int ubpl_ema_flat(float* ema, const float* param, int64_t n, float alpha, float one_minus_alpha,
                             void* stream) {
  UBPL_REQUIRE(ema && param && n >= 0, "ubpl_ema_flat: bad arguments");
  UBPL_REQUIRE(((reinterpret_cast<uintptr_t>(ema) | reinterpret_cast<uintptr_t>(param)) & 15) == 0,
               "ubpl_ema_flat: buffers must be 16-byte aligned");
  if (n == 0) return UBPL_OK;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  ema_flat_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(ema, param, n, alpha, one_minus_alpha);
  return check_launch("ubpl_ema_flat");
}
