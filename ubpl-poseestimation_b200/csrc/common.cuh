// Shared device/host helpers for libubpl_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include "../../include/ubpl_b200.h"

namespace ubpl {

void set_error(const char* fmt, ...);
int check_launch(const char* what);          // cudaGetLastError -> code
int sm_count();
int smem_optin();

#define UBPL_REQUIRE(cond, ...)                  \
  do {                                           \
    if (!(cond)) {                               \
      ubpl::set_error(__VA_ARGS__);              \
      return UBPL_ERR_INVALID;                   \
    }                                            \
  } while (0)

// ---------------------------------------------------------------------------------------------
// mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -- global -> shared staging
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a hardware-defined time)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// L2 policy: the heat-maps are read exactly once -> evict_first.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}

// Asynchronous bulk prefetch of a contiguous global range into L2 (no registers, no smem): src 16-byte
// aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}

// ---------------------------------------------------------------------------------------------
// streaming 128-bit global accesses (read-once / write-once data: keep it out of L1)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_cs(float4* p, const float4& v) {       // evict-first streaming store
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// mean-teacher EMA of one element with ATen's rounding: t = ema*alpha (rounded), then fma(param, 1-alpha, t)
// (utils/parameters.py:8: ema.mul_(alpha).add_(param, alpha=1-alpha))
__device__ __forceinline__ float ema1(float e, float p, float a, float oma) {
  return __fmaf_rn(p, oma, __fmul_rn(e, a));
}

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// torch.max ordering on (value, flat index): larger value wins, NaN beats everything, ties and
// NaN-vs-NaN go to the smaller index (first occurrence in row-major order).
__device__ __forceinline__ bool arg_better(float v, int i, float bv, int bi) {
  const bool vn = (v != v), bn = (bv != bv);
  if (vn || bn) return vn && (!bn || i < bi);
  return (v > bv) || (v == bv && i < bi);
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    if (arg_better(ov, oi, v, i)) {
      v = ov;
      i = oi;
    }
  }
}

// Arg-max over the warp for FINITE values (no NaN among the lanes' v): one CREDUX for the value, one REDUX for the
// smallest index among the lanes that hold it -- torch.max's first-index rule -- and the winning lane's own value
// (so a zero maximum keeps the sign the shuffle tree of warp_argmax would return).  Lanes without a candidate pass
// (-inf, 0x7fffffff).
__device__ __forceinline__ float warp_redux_max(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}
__device__ __forceinline__ void warp_argmax_finite(float& v, int& i) {
  const float m = warp_redux_max(v);
  const int c = (v == m) ? i : 0x7fffffff;
  const int im = __reduce_min_sync(0xffffffffu, c);
  const unsigned who = __ballot_sync(0xffffffffu, c == im);
  v = __shfl_sync(0xffffffffu, v, __ffs(who) - 1);
  i = im;
}

// n / d and n % d for 32-bit n with a host-computed magic pair: q = umulhi(n, mul) >> shr; a power-of-two
// divisor (mul == 0) is a plain shift.
struct FastDiv {
  unsigned mul, shr, d;
  __host__ void init(unsigned div) {
    d = div;
    unsigned l = 0;
    while ((1ull << l) < div) ++l;                       // l = ceil(log2 div)
    if ((1ull << l) == div) { mul = 0; shr = l; return; }
    const unsigned long long m = ((1ull << 32) * ((1ull << l) - div)) / div + 1;
    mul = (unsigned)m; shr = l;
  }
  __device__ __forceinline__ unsigned div(unsigned n) const {
    if (mul == 0) return n >> shr;
    const unsigned t = __umulhi(n, mul);
    return (t + ((n - t) >> 1)) >> (shr - 1);
  }
  __device__ __forceinline__ void divmod(unsigned n, unsigned& q, unsigned& r) const { q = div(n); r = n - q * d; }
};

// ---------------------------------------------------------------------------------------------
// Gaussian target geometry shared by K2 (visibility gate) and K3 (render)
// ---------------------------------------------------------------------------------------------
struct Gauss {
  int kx, ky;       // int(kp): truncation toward zero
  double cx, cy;    // int(kp) * 1.0 / stride
  float vis;
};

// utils/process.py:262-272
__device__ __forceinline__ Gauss gauss_setup(float kxf, float kyf, int img_h, int img_w, float stride, float sigma) {
  Gauss g;
  g.kx = (int)kxf;
  g.ky = (int)kyf;
  const int ulx = (int)((float)g.kx - sigma), uly = (int)((float)g.ky - sigma);
  const int brx = (int)((float)g.kx + sigma + 1.f), bry = (int)((float)g.ky + sigma + 1.f);
  g.vis = (brx >= img_w || bry >= img_h || ulx < 0 || uly < 0) ? 0.f : 1.f;
  g.cx = (double)g.kx * 1.0 / (double)stride;
  g.cy = (double)g.ky * 1.0 / (double)stride;
  return g;
}

// ---------------------------------------------------------------------------------------------
// python-float distance shared by K2 and by the K2 epilogue fused into K1
// ---------------------------------------------------------------------------------------------
int pow_table(const int32_t** keys, const double** vals, const uint32_t** bits, int* n, int* rmax);

struct PowTab {
  const int32_t* key;
  const double* val;
  const uint32_t* bits;   // one bit per radicand: set where libm pow and sqrt disagree
  int n, rmax;
};

// python: ((x1-x2)**2 + (y1-y2)**2) ** 0.5  == libm pow(r, 0.5); see api.cu for the table.
__device__ __forceinline__ double py_dist(double x1, double y1, double x2, double y2, const PowTab& T) {
  const double dx = __dsub_rn(x1, x2), dy = __dsub_rn(y1, y2);
  const double r = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
  double d = sqrt(r);
  if (r <= (double)T.rmax) {
    const int ri = (int)r;
    if ((double)ri == r && ((__ldg(T.bits + (ri >> 5)) >> (ri & 31)) & 1u)) {
      int lo = 0, hi = T.n - 1;
      while (lo <= hi) {
        const int mid = (lo + hi) >> 1;
        const int k = T.key[mid];
        if (k == ri) { d = T.val[mid]; break; }
        if (k < ri) lo = mid + 1; else hi = mid - 1;
      }
    }
  }
  return d;
}

}  // namespace ubpl
