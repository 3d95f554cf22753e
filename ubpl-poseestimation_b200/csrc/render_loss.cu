// K3: Gaussian target render + masked joint-MSE, forward and gradient in one pass over the maps.
//
// Reference semantics (file:line in /root/reference):
//   utils/process.py:253-278,394-397  kps_heatmap / heatmap_gaussian (float64 exp(-D2/2/s/s),
//                                     >1 -> 1, <0.01 -> 0, visibility test in image space)
//   utils/losses.py:8-29     JointMSELoss      utils/losses.py:32-53    JointDistLoss
//   utils/losses.py:169-210  JointPseudoLoss3  utils/losses.py:246-286  JointDistLoss_mt2
//
// Pure streaming kernels (HBM-bound): one CTA per (sample, joint); every student map is read once
// with 128-bit loads, its gradient is written once, the rendered target is written once (or not at
// all).  The Gaussian is separable, exp(-dx^2/2s^2) * exp(-dy^2/2s^2), evaluated in float32
// (<= 5 ulp of the reference's float64 value, tolerance 1e-5); the 0.01 cut is re-evaluated in
// float64 for the rare pixel within 1e-5 of it so the support is identical to the reference's.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

// resident CTAs per SM the streaming kernel is compiled for: 6 (80 registers, spills) or 5 (102 registers)
#ifndef UBPL_K3_OCC_DEFAULT
#define UBPL_K3_OCC_DEFAULT 5
#endif

namespace ubpl {

__device__ __forceinline__ float4 load4(const float* base, int q, int n, bool vec) {
  if (vec) return ldg_stream(reinterpret_cast<const float4*>(base) + q);
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  const int k = q << 2;
  if (k < n) r.x = __ldg(base + k);
  if (k + 1 < n) r.y = __ldg(base + k + 1);
  if (k + 2 < n) r.z = __ldg(base + k + 2);
  if (k + 3 < n) r.w = __ldg(base + k + 3);
  return r;
}
__device__ int g_store_mode = 0;   // experiment knob (UBPL_K3_STORE): 0 = L1::no_allocate, 1 = .cs, 2 = default
__device__ __forceinline__ void store4(float* base, int q, int n, bool vec, const float4& v) {
  if (vec) {
    float4* p = reinterpret_cast<float4*>(base) + q;
    const int m = g_store_mode;
    if (m == 0) stg_stream(p, v); else if (m == 1) stg_cs(p, v); else *p = v;
    return;
  }
  const int k = q << 2;
  if (k < n) base[k] = v.x;
  if (k + 1 < n) base[k + 1] = v.y;
  if (k + 2 < n) base[k + 2] = v.z;
  if (k + 3 < n) base[k + 3] = v.w;
}

// block-wide sum / max over blockDim.x threads (blockDim multiple of 32, <= 1024); result on all threads
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];      // fixed order: deterministic
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

__device__ __forceinline__ float gauss_1d(int k, double c, float sigma) {
  const float d = (float)((double)k - c);
  return expf(-(d * d) / (2.f * sigma * sigma));
}
// the rare pixel within rounding of the 0.01 cut: decide like the reference, in float64
__device__ __noinline__ float gauss_px_exact(int x, int y, double cx, double cy, float sigma) {
  const double dx = (double)x - cx, dy = (double)y - cy;
  const double D2 = dx * dx + dy * dy;
  const double k = exp(-D2 / 2.0 / (double)sigma / (double)sigma);
  return (k < 0.01) ? 0.f : (float)k;
}
__device__ __forceinline__ float gauss_px(const float* ex, const float* ey, int x, int y, const Gauss& g, float sigma) {
  const float v = ex[x] * ey[y];
  if (fabsf(v - 0.01f) < 1e-5f) return gauss_px_exact(x, y, g.cx, g.cy, sigma);
  return (v < 0.01f) ? 0.f : v;
}

// the same with the two separable factors already in registers
__device__ __forceinline__ float gauss_px_v(float exv, float eyv, int x, int y, int kx, int ky, float stride, float sigma) {
  const float v = exv * eyv;
  if (fabsf(v - 0.01f) < 1e-5f)      // centre recomputed as in gauss_setup: only the integer key point stays live
    return gauss_px_exact(x, y, (double)kx * 1.0 / (double)stride, (double)ky * 1.0 / (double)stride, sigma);
  return (v < 0.01f) ? 0.f : v;
}

// One WARP per (sample, joint): no block barriers; every lane keeps 8 independent 128-bit loads in
// flight; the separable Gaussian factors live in a per-warp shared-memory slice.
// CH, CW: the map's shape when the instance is compiled for it (64x64, 128x128: then also S == SS stacks, four warps per
// CTA and every item shared by its CTA), 0 = everything read from the arguments.  Knowing the shape takes the kernel
// from 57.0 to 54.8 us on c2 -- onto the read/write-mix bound of tools/rw_micro.cu (53.2-55.3 us).
template <bool VEC, int SS, int OCC, int CH, int CW>
__global__ void __launch_bounds__(128, OCC) render_mse_kernel(
    const float* __restrict__ kps, const float* __restrict__ gate_in, const float* __restrict__ sample_w,
    const float* __restrict__ pred, long long pB, long long pS, long long pJ, float* __restrict__ grad, long long gB,
    long long gS, long long gJ, float* __restrict__ target, int B, int S_, int J, int H_, int W_, int img_h, int img_w,
    float stride, float sigma, const float* __restrict__ grad_scale, const int32_t* __restrict__ count_in,
    float loss_weight, float* __restrict__ grad_scale_out, float* __restrict__ gate_out,
    float* __restrict__ per_loss, const FastDiv divW4, double* __restrict__ summary, unsigned char* __restrict__ sum_ws,
    int all_coop_) {
  extern __shared__ float sm[];
  const int H = CH ? CH : H_, W = CW ? CW : W_, S = CH ? SS : S_, all_coop = CH ? 1 : all_coop_;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = CH ? 4 : (int)(blockDim.x >> 5);
  // Programmatic dependent launch: when the kernel is launched with programmatic stream serialization its CTAs are
  // scheduled as soon as the SMs of the preceding kernel (K1 or the selector) free up, and wait HERE until that
  // kernel has completed and its results (key points, gates, count) are visible; a no-op for a plain launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float* ex = sm + (size_t)warp * (W + H);   // [W]
  float* ey = ex + W;                        // [H]
  const int HW = H * W, nq = (HW + 3) >> 2;
  float gs = grad_scale ? *grad_scale : 1.f;
  if (count_in) {                                    // weight / n, n = S * #(open gates)  (MT_UBPL.py:266)
    const int cnt = *count_in;
    gs = (cnt > 0) ? loss_weight / (float)cnt : loss_weight;
    if (grad_scale_out && blockIdx.x == 0 && threadIdx.x == 0) *grad_scale_out = gs;
  }
  const float inv_hw = 1.f / (float)HW;
  const long long BJ = (long long)B * J;
  constexpr int U = 8 / SS;
  // this warp's share of the fused reduction lives in shared memory (updated by lane 0 once per item): the
  // streaming loop below is register-bound and must not carry accumulators
  __shared__ double s_red[3][32];
  if (lane == 0) { s_red[0][warp] = 0.0; s_red[1][warp] = 0.0; s_red[2][warp] = 0.0; }
  // Work split.  Phase 0: one warp per (sample, joint), as many full rounds of gridDim*wpb items as there are.
  // Phase 1: the items left over (fewer than one round) are each shared by the wpb warps of a CTA, which split
  // the map's float4s between them -- a warp that had to stream a second whole item on an otherwise idle GPU
  // would be latency-bound and stretch the launch by ~10 us.
  __shared__ float s_part[2][32];
  const long long round_items = (long long)gridDim.x * wpb;
  const long long full = all_coop ? 0 : (BJ / round_items) * round_items;   // all_coop: every item is shared by a CTA
  for (int phase = 0; phase < 2; ++phase) {
  const int nparts = phase ? wpb : 1, part = phase ? warp : 0;
  const long long it_step = phase ? (long long)gridDim.x : round_items;
  const long long it_end = phase ? BJ : full;
  for (long long item = phase ? full + blockIdx.x : (long long)blockIdx.x * wpb + warp; item < it_end; item += it_step) {
    const int b = (int)(item / J), j = (int)(item % J);
    const int q_lo = (int)((long long)nq * part / nparts), q_hi = (int)((long long)nq * (part + 1) / nparts);
    if (VEC && lane == 0 && q_hi > q_lo) {
      // pull this item's student maps towards L2 while the Gaussian factors are set up; the 128-bit
      // loads below then see L2 latency instead of HBM latency
      for (int st = 0; st < S; ++st)
        bulk_prefetch_l2(pred + (long long)b * pB + (long long)st * pS + (long long)j * pJ + ((long long)q_lo << 2),
                         (uint32_t)(q_hi - q_lo) * 16u);
    }
    const Gauss g = gauss_setup(kps[2 * item], kps[2 * item + 1], img_h, img_w, stride, sigma);
    __syncwarp();
    for (int k = lane; k < W + H; k += 32) {
      if (k < W) ex[k] = gauss_1d(k, g.cx, sigma); else ey[k - W] = gauss_1d(k - W, g.cy, sigma);
    }
    __syncwarp();
    const float gate = (gate_in ? gate_in[item] : 1.f) * g.vis;
    const float wb = sample_w ? sample_w[b] : 1.f;
    const float gcoef = gs * 2.f * inv_hw * gate * wb;
    const bool fin = (part == 0) && (lane == 0);         // the lane that writes this item's scalars
    if (fin && gate_out) gate_out[item] = gate;
    // support box: exp(-r^2/2s^2) >= 0.01 needs r <= s*sqrt(2 ln 100) = 3.035 s; 3.05 s + 1 is a safe superset
    const float rad = 3.05f * sigma + 1.f;
    const int xlo = (int)floorf((float)g.cx - rad), xhi = (int)ceilf((float)g.cx + rad);
    const int ylo = (int)floorf((float)g.cy - rad), yhi = (int)ceilf((float)g.cy + rad);
    // SS stacks are streamed together so that the target texels are computed once per float4 and shared
    for (int st0 = 0; st0 < S; st0 += SS) {
      const float* p[SS];
      float* gr[SS];
      float sse[SS];
#pragma unroll
      for (int ss = 0; ss < SS; ++ss) {
        p[ss] = pred + (long long)b * pB + (long long)(st0 + ss) * pS + (long long)j * pJ;
        gr[ss] = grad ? grad + (long long)b * gB + (long long)(st0 + ss) * gS + (long long)j * gJ : nullptr;
        sse[ss] = 0.f;
      }
      float* tg = (target && st0 == 0) ? target + item * HW : nullptr;
      for (int q0 = q_lo + lane; q0 < q_hi; q0 += 32 * U) {
        float4 pv[SS][U];
#pragma unroll
        for (int ss = 0; ss < SS; ++ss) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int q = q0 + 32 * u;
            if (q < q_hi) pv[ss][u] = load4(p[ss], q, HW, VEC);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = q0 + 32 * u;
          if (q >= q_hi) continue;
          const int k = q << 2;
          float4 t;
          if (VEC) {                                    // W % 4 == 0: the four texels share a row
            unsigned yu, xu;
            if (CW) { yu = (unsigned)q / (unsigned)(CW / 4 ? CW / 4 : 1); xu = (unsigned)q - yu * (unsigned)(CW / 4 ? CW / 4 : 1); }
            else divW4.divmod((unsigned)q, yu, xu);
            const int y = (int)yu, x = (int)(xu << 2);
            // outside the Gaussian's support box the target is exactly 0 (the 0.01 cut, process.py:275)
            if (y < ylo || y > yhi || x + 3 < xlo || x > xhi) {
              t = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
              t.x = gauss_px(ex, ey, x, y, g, sigma); t.y = gauss_px(ex, ey, x + 1, y, g, sigma);
              t.z = gauss_px(ex, ey, x + 2, y, g, sigma); t.w = gauss_px(ex, ey, x + 3, y, g, sigma);
            }
          } else {
            float tt[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int kk = min(k + c, HW - 1);
              tt[c] = gauss_px(ex, ey, kk % W, kk / W, g, sigma);
            }
            t = make_float4(tt[0], tt[1], tt[2], tt[3]);
          }
#pragma unroll
          for (int ss = 0; ss < SS; ++ss) {
            float4 v = pv[ss][u];
            if (!VEC) {                                  // padded lanes contribute zero error
              if (k + 1 >= HW) v.y = t.y;
              if (k + 2 >= HW) v.z = t.z;
              if (k + 3 >= HW) v.w = t.w;
            }
            const float4 d = make_float4(v.x - t.x, v.y - t.y, v.z - t.z, v.w - t.w);
            sse[ss] += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
            if (gr[ss]) store4(gr[ss], q, HW, VEC, make_float4(gcoef * d.x, gcoef * d.y, gcoef * d.z, gcoef * d.w));
          }
          if (tg) store4(tg, q, HW, VEC, t);
        }
      }
#pragma unroll
      for (int ss = 0; ss < SS; ++ss) {
        float tot = warp_sum(sse[ss]);
        if (nparts > 1) {                                // add the warps' partial sums in warp order
          if (lane == 0) s_part[ss][warp] = tot;
          __syncthreads();
          tot = 0.f;
          for (int w2 = 0; w2 < nparts; ++w2) tot += s_part[ss][w2];
          __syncthreads();
        }
        const float pl = ((tot * inv_hw) * gate) * wb;
        if (fin && per_loss) per_loss[((long long)b * S + st0 + ss) * J + j] = pl;
        if (summary && fin) { s_red[0][warp] += (double)pl; s_red[1][warp] += (pl > 0.f) ? 1.0 : 0.0; }
      }
    }
    if (summary && fin) s_red[2][warp] += (gate > 0.f) ? 1.0 : 0.0;
  }
  }
  if (summary) {
    // Loss reduction fused into this launch (what loss_finalize_kernel computes with mask = NULL): every CTA
    // leaves its partial sums in the workspace, the CTA that takes the last ticket adds the partials in CTA
    // order (the item -> CTA mapping is static, so the result is reproducible) and returns the ticket to zero.
    __shared__ unsigned s_last;
    double* part = reinterpret_cast<double*>(sum_ws + 16);
    unsigned* ticket = reinterpret_cast<unsigned*>(sum_ws);
    __syncthreads();
    if (threadIdx.x == 0) {
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (int i = 0; i < wpb; ++i) { t0 += s_red[0][i]; t1 += s_red[1][i]; t2 += s_red[2][i]; }
      part[3 * blockIdx.x] = t0; part[3 * blockIdx.x + 1] = t1; part[3 * blockIdx.x + 2] = t2;
      __threadfence();
      s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (unsigned c = threadIdx.x; c < gridDim.x; c += blockDim.x) {
        t0 += __ldcg(part + 3 * c); t1 += __ldcg(part + 3 * c + 1); t2 += __ldcg(part + 3 * c + 2);
      }
      t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
      __syncthreads();
      if (lane == 0) { s_red[0][warp] = t0; s_red[1][warp] = t1; s_red[2][warp] = t2; }
      __syncthreads();
      if (threadIdx.x == 0) {
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        for (int i = 0; i < wpb; ++i) { u0 += s_red[0][i]; u1 += s_red[1][i]; u2 += s_red[2][i]; }
        summary[0] = u0; summary[1] = u1; summary[2] = (double)(BJ * S); summary[3] = u2;
        *ticket = 0u;
      }
    }
  }
}

// The lean form of render_mse_kernel for the shapes the drivers run (128-bit aligned maps, S = 1 or 2, W/4 a power
// of two <= 128, H*W a multiple of 4096): same arithmetic per texel, ~3x fewer instructions per item.  The generic
// kernel above spends 1.4 k warp instructions per item and warp on index arithmetic, guards and the per-warp
// Gaussian set-up (ncu: issue-bound at 43 % with 20 warps per SM, DRAM 51 % busy).  Here a CTA of 128 threads owns
// an item; thread t handles the float4 column t % (W/4) of rows t / (W/4) + R*u (R = 128 / (W/4)), so the column
// factors of the separable Gaussian are four registers per item and a position costs one row factor; ALL loads of an
// item (8 positions x S stacks per thread and chunk) are issued before the Gaussian tables are built, the next item's
// key point / gate are fetched a whole item ahead, and there is ONE block barrier per item (tables and partial sums
// are double-buffered).  Partial sums are added in warp order: reproducible, not bit-identical to the generic kernel
// (different split of the texels over the warps), inside the 1e-5 the tests allow for the loss.
template <int SS, int OCC>
__global__ void __launch_bounds__(128, OCC) render_mse_fast_kernel(
    const float* __restrict__ kps, const float* __restrict__ gate_in, const float* __restrict__ sample_w,
    const float* __restrict__ pred, long long pB, long long pS, long long pJ, float* __restrict__ grad, long long gB,
    long long gS, long long gJ, float* __restrict__ target, int B, int J, int H, int W, int img_h, int img_w,
    float stride, float sigma, const float* __restrict__ grad_scale, const int32_t* __restrict__ count_in,
    float loss_weight, float* __restrict__ grad_scale_out, float* __restrict__ gate_out,
    float* __restrict__ per_loss, double* __restrict__ summary, unsigned char* __restrict__ sum_ws, int w4_shift) {
  extern __shared__ float sm[];                          // two tables of W + H Gaussian factors
  __shared__ float s_part[2][SS][4];
  __shared__ float s_meta[2][2];                         // gate and sample weight of the item whose partials are pending
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  asm volatile("griddepcontrol.wait;" ::: "memory");     // see render_mse_kernel
  const int HW = H * W, nq = HW >> 2, WH = W + H;
  float gs = grad_scale ? *grad_scale : 1.f;
  if (count_in) {
    const int cnt = *count_in;
    gs = (cnt > 0) ? loss_weight / (float)cnt : loss_weight;
    if (grad_scale_out && blockIdx.x == 0 && tid == 0) *grad_scale_out = gs;
  }
  const float inv_hw = 1.f / (float)HW;
  const long long BJ = (long long)B * J;
  const int x4 = tid & ((1 << w4_shift) - 1), yb = tid >> w4_shift, R = 128 >> w4_shift;
  const int x = x4 << 2;
  const float rad = 3.05f * sigma + 1.f;                 // support box, as in render_mse_kernel
  __shared__ double s_acc[3];                            // thread 0: this CTA's share of the fused reduction
  if (tid == 0) { s_acc[0] = 0.0; s_acc[1] = 0.0; s_acc[2] = 0.0; }

  // key point, gate and sample weight of an item are fetched one item ahead by threads 0..3 and handed to the CTA
  // through shared memory (s_next[parity]): no global-load latency at the top of an item, no registers held for it
  __shared__ float s_next[2][4];
  auto fetch_meta = [&](int it, int par) {               // threads 0..3; visible after the next block barrier
    if (tid < 4 && it < (int)BJ) {
      float v;
      if (tid < 2) v = kps[2 * (long long)it + tid];
      else if (tid == 2) v = gate_in ? gate_in[it] : 1.f;
      else v = sample_w ? sample_w[it / J] : 1.f;
      s_next[par][tid] = v;
    }
  };
  auto finalize = [&](int it, int par) {                 // thread 0: scalars of a finished item
    const int b = it / J, j = it - b * J;
    const float gate = s_meta[par][0], wb = s_meta[par][1];
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) {
      const float tot = ((s_part[par][ss][0] + s_part[par][ss][1]) + s_part[par][ss][2]) + s_part[par][ss][3];
      const float pl = ((tot * inv_hw) * gate) * wb;
      if (per_loss) per_loss[((long long)b * SS + ss) * J + j] = pl;
      s_acc[0] += (double)pl; s_acc[1] += (pl > 0.f) ? 1.0 : 0.0;
    }
    s_acc[2] += (gate > 0.f) ? 1.0 : 0.0;
  };
  int item = (int)blockIdx.x, prev = -1;
  fetch_meta(item, 0);
  __syncthreads();
  for (int k = 0; item < (int)BJ; item += (int)gridDim.x, ++k) {
    const int par = k & 1;
    const int b = item / J, j = item - b * J;
    const float kx = s_next[par][0], ky = s_next[par][1], gate_raw = s_next[par][2], wb = s_next[par][3];
    const float* p[SS];
    float* gr[SS];
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) {
      p[ss] = pred + (long long)b * pB + (long long)ss * pS + (long long)j * pJ;
      gr[ss] = grad ? grad + (long long)b * gB + (long long)ss * gS + (long long)j * gJ : nullptr;
    }
    float* tg = target ? target + (long long)item * HW : nullptr;
    float4 pv[SS][8];
#pragma unroll
    for (int ss = 0; ss < SS; ++ss)
#pragma unroll
      for (int u = 0; u < 8; ++u) pv[ss][u] = ldg_stream(reinterpret_cast<const float4*>(p[ss]) + tid + 128 * u);
    fetch_meta(item + (int)gridDim.x, par ^ 1);          // consumed one item later
    const Gauss g = gauss_setup(kx, ky, img_h, img_w, stride, sigma);
    float* ex = sm + par * WH;
    float* ey = ex + W;
    for (int i = tid; i < WH; i += 128) ex[i] = (i < W) ? gauss_1d(i, g.cx, sigma) : gauss_1d(i - W, g.cy, sigma);
    const float gate = gate_raw * g.vis;
    if (tid == 0) {
      s_meta[par][0] = gate; s_meta[par][1] = wb;
      if (gate_out) gate_out[item] = gate;
    }
    __syncthreads();                                     // tables of this item, partial sums of the previous one
    if (tid == 0 && prev >= 0) finalize(prev, par ^ 1);
    const float gcoef = gs * 2.f * inv_hw * gate * wb;
    const int xlo = (int)floorf((float)g.cx - rad), xhi = (int)ceilf((float)g.cx + rad);
    const int ylo = (int)floorf((float)g.cy - rad), yhi = (int)ceilf((float)g.cy + rad);
    const bool col_in = !(x + 3 < xlo || x > xhi);
    const int gkx = g.kx, gky = g.ky;
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
    if (col_in) { e0 = ex[x]; e1 = ex[x + 1]; e2 = ex[x + 2]; e3 = ex[x + 3]; }
    float sse[SS];
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) sse[ss] = 0.f;
    for (int q0 = 0; q0 < nq; q0 += 1024) {
      if (q0) {
#pragma unroll
        for (int ss = 0; ss < SS; ++ss)
#pragma unroll
          for (int u = 0; u < 8; ++u) pv[ss][u] = ldg_stream(reinterpret_cast<const float4*>(p[ss]) + q0 + tid + 128 * u);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int q = q0 + tid + 128 * u;
        const int y = (q0 >> w4_shift) + yb + R * u;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_in && y >= ylo && y <= yhi) {
          const float ev = ey[y];
          t.x = gauss_px_v(e0, ev, x, y, gkx, gky, stride, sigma); t.y = gauss_px_v(e1, ev, x + 1, y, gkx, gky, stride, sigma);
          t.z = gauss_px_v(e2, ev, x + 2, y, gkx, gky, stride, sigma); t.w = gauss_px_v(e3, ev, x + 3, y, gkx, gky, stride, sigma);
        }
#pragma unroll
        for (int ss = 0; ss < SS; ++ss) {
          const float4 v = pv[ss][u];
          const float4 d = make_float4(v.x - t.x, v.y - t.y, v.z - t.z, v.w - t.w);
          sse[ss] += d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w;
          if (gr[ss]) stg_stream(reinterpret_cast<float4*>(gr[ss]) + q, make_float4(gcoef * d.x, gcoef * d.y, gcoef * d.z, gcoef * d.w));
        }
        if (tg) stg_stream(reinterpret_cast<float4*>(tg) + q, t);
      }
    }
#pragma unroll
    for (int ss = 0; ss < SS; ++ss) {
      const float tot = warp_sum(sse[ss]);
      if (lane == 0) s_part[par][ss][warp] = tot;
    }
    prev = item;
  }
  __syncthreads();
  if (tid == 0 && prev >= 0) finalize(prev, ((prev - (int)blockIdx.x) / (int)gridDim.x) & 1);
  if (summary) {
    // the fused loss reduction of render_mse_kernel: per-CTA partials, the CTA with the last ticket adds them in CTA order
    __shared__ unsigned s_last;
    __shared__ double s_fin[3][4];
    double* part = reinterpret_cast<double*>(sum_ws + 16);
    unsigned* ticket = reinterpret_cast<unsigned*>(sum_ws);
    if (tid == 0) {
      part[3 * blockIdx.x] = s_acc[0]; part[3 * blockIdx.x + 1] = s_acc[1]; part[3 * blockIdx.x + 2] = s_acc[2];
      __threadfence();
      s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      double t0 = 0.0, t1 = 0.0, t2 = 0.0;
      for (unsigned c = tid; c < gridDim.x; c += 128) {
        t0 += __ldcg(part + 3 * c); t1 += __ldcg(part + 3 * c + 1); t2 += __ldcg(part + 3 * c + 2);
      }
      t0 = warp_sum(t0); t1 = warp_sum(t1); t2 = warp_sum(t2);
      if (lane == 0) { s_fin[0][warp] = t0; s_fin[1][warp] = t1; s_fin[2][warp] = t2; }
      __syncthreads();
      if (tid == 0) {
        double u0 = 0.0, u1 = 0.0, u2 = 0.0;
        for (int i = 0; i < 4; ++i) { u0 += s_fin[0][i]; u1 += s_fin[1][i]; u2 += s_fin[2][i]; }
        summary[0] = u0; summary[1] = u1; summary[2] = (double)(BJ * SS); summary[3] = u2;
        *ticket = 0u;
      }
    }
  }
}

__global__ void render_targets_kernel(const float* __restrict__ kps, int N, int H, int W, int img_h, int img_w,
                                      float stride, float sigma, float* __restrict__ heatmap,
                                      float* __restrict__ kps_out) {
  extern __shared__ float sm[];
  float* ex = sm;
  float* ey = sm + W;
  const int HW = H * W;
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    const Gauss g = gauss_setup(kps[3 * n], kps[3 * n + 1], img_h, img_w, stride, sigma);
    __syncthreads();
    for (int k = threadIdx.x; k < W + H; k += blockDim.x) {
      if (k < W) ex[k] = gauss_1d(k, g.cx, sigma); else ey[k - W] = gauss_1d(k - W, g.cy, sigma);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < HW; k += blockDim.x) heatmap[(long long)n * HW + k] = gauss_px(ex, ey, k % W, k / W, g, sigma);
    if (threadIdx.x == 0 && kps_out) {
      kps_out[3 * n] = kps[3 * n];
      kps_out[3 * n + 1] = kps[3 * n + 1];
      kps_out[3 * n + 2] = kps[3 * n + 2] * g.vis;
    }
  }
}

// Dense-target loss: CH float4 chunks per thread kept in registers so the gradient (which depends
// on the map's own max through the mask) is written from registers: one HBM read per map.
template <int CH, int MAXT>
__global__ void __launch_bounds__(MAXT) dense_mse_kernel(
    const float* __restrict__ pred, long long pB, long long pS, long long pJ, const float* __restrict__ tgt, int M,
    long long tM, long long tB, long long tS, long long tJ, const float* __restrict__ coef, int mask_mode, float thr,
    float* __restrict__ grad, long long gB, long long gS, long long gJ, int B, int S, int J, int H, int W,
    const float* __restrict__ grad_scale, float* __restrict__ per_loss, float* __restrict__ mask_o,
    float* __restrict__ vmax_p_o, float* __restrict__ vmax_t_o, int vec) {
  __shared__ float red[32];
  const int HW = H * W, nq = (HW + 3) >> 2;
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float inv_hw = 1.f / (float)HW;
  const float invM = (float)M;
  for (long long item = blockIdx.x; item < (long long)B * J; item += gridDim.x) {
    const int b = (int)(item / J), j = (int)(item % J);
    const float cf = coef ? coef[item] : 1.f;
    float4 tb[CH];
    float vt = -INFINITY;
    for (int s = 0; s < S; ++s) {
      if (s == 0 || tS != 0) {
        // float32 mean over the M teacher maps in index order (torch.mean): sum, then one division
        const float* t0 = tgt + (long long)b * tB + (long long)s * tS + (long long)j * tJ;
        float tmx = -INFINITY;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int q = threadIdx.x + u * blockDim.x;
          if (q < nq) {
            float4 a = load4(t0, q, HW, vec);
            for (int m = 1; m < M; ++m) {
              const float4 c = load4(t0 + (long long)m * tM, q, HW, vec);
              a.x = __fadd_rn(a.x, c.x); a.y = __fadd_rn(a.y, c.y); a.z = __fadd_rn(a.z, c.z); a.w = __fadd_rn(a.w, c.w);
            }
            if (M > 1) { a.x = __fdiv_rn(a.x, invM); a.y = __fdiv_rn(a.y, invM); a.z = __fdiv_rn(a.z, invM); a.w = __fdiv_rn(a.w, invM); }
            tb[u] = a;
            const int k = q << 2;
            tmx = fmaxf(tmx, a.x);
            if (vec || k + 1 < HW) tmx = fmaxf(tmx, a.y);
            if (vec || k + 2 < HW) tmx = fmaxf(tmx, a.z);
            if (vec || k + 3 < HW) tmx = fmaxf(tmx, a.w);
          }
        }
        if (mask_mode != 0 || vmax_t_o) vt = block_max(tmx, red);
      }
      const float* p = pred + (long long)b * pB + (long long)s * pS + (long long)j * pJ;
      float4 d[CH];
      float sse = 0.f, pmx = -INFINITY;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = threadIdx.x + u * blockDim.x;
        if (q < nq) d[u] = load4(p, q, HW, vec);
      }
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = threadIdx.x + u * blockDim.x;
        if (q < nq) {
          const int k = q << 2;
          const float4 pv = d[u];
          pmx = fmaxf(pmx, pv.x);
          if (vec || k + 1 < HW) pmx = fmaxf(pmx, pv.y);
          if (vec || k + 2 < HW) pmx = fmaxf(pmx, pv.z);
          if (vec || k + 3 < HW) pmx = fmaxf(pmx, pv.w);
          d[u] = make_float4(pv.x - tb[u].x, pv.y - tb[u].y, pv.z - tb[u].z, pv.w - tb[u].w);
          sse += d[u].x * d[u].x + d[u].y * d[u].y + d[u].z * d[u].z + d[u].w * d[u].w;   // padded lanes: 0-0
        }
      }
      const float tot = block_sum(sse, red);
      float mk = 1.f, vp = 0.f;
      if (mask_mode == 1 || vmax_p_o) vp = block_max(pmx, red);
      if (mask_mode == 1) mk = (vp >= thr && vt >= thr) ? 1.f : 0.f;
      else if (mask_mode == 2) mk = (vt >= thr) ? 1.f : 0.f;
      const long long o = ((long long)b * S + s) * J + j;
      if (threadIdx.x == 0) {
        if (per_loss) per_loss[o] = (tot * inv_hw) * cf;
        if (mask_o) mask_o[o] = mk;
        if (vmax_p_o) vmax_p_o[o] = vp;
        if (vmax_t_o) vmax_t_o[o] = vt;
      }
      if (grad) {
        float* gr = grad + (long long)b * gB + (long long)s * gS + (long long)j * gJ;
        const float gcoef = gs * 2.f * inv_hw * cf * mk;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int q = threadIdx.x + u * blockDim.x;
          if (q < nq) store4(gr, q, HW, vec, make_float4(gcoef * d[u].x, gcoef * d[u].y, gcoef * d[u].z, gcoef * d[u].w));
        }
      }
    }
  }
}

// The same loss for the shapes the drivers use (S <= 2 stacks, maps of <= 4096 texels, 128-bit aligned): every load
// of an item -- the M teacher maps and the S student maps -- is issued before anything is consumed, the five block
// reductions of the generic kernel (teacher max, then error sum and student max per stack: ten barriers per item)
// become ONE multi-value reduction behind ONE barrier (partials double-buffered by item parity), and the gradient
// is written from registers.  Per-thread accumulation order, the warp tree and the warp-order final sums are those
// of dense_mse_kernel, so the results are bit-identical.
template <int S_, bool TPS>     // TPS: the targets have their own stack axis (tS != 0)
__global__ void __launch_bounds__(256) dense_mse_fast_kernel(
    const float* __restrict__ pred, long long pB, long long pS, long long pJ, const float* __restrict__ tgt, int M,
    long long tM, long long tB, long long tS, long long tJ, const float* __restrict__ coef, int mask_mode, float thr,
    float* __restrict__ grad, long long gB, long long gS, long long gJ, int B, int J, int HW,
    const float* __restrict__ grad_scale, float* __restrict__ per_loss, float* __restrict__ mask_o,
    float* __restrict__ vmax_p_o, float* __restrict__ vmax_t_o) {
  constexpr int CH = 4, TS = TPS ? S_ : 1, NV = 2 * S_ + TS;       // values reduced per item: sse[S], pmax[S], tmax[TS]
  __shared__ float red[2][NV][8];
  const int nq = HW >> 2, t = threadIdx.x, lane = t & 31, w = t >> 5, nw = blockDim.x >> 5;
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float inv_hw = 1.f / (float)HW;
  const float fM = (float)M;
  int par = 0;
  for (long long item = blockIdx.x; item < (long long)B * J; item += gridDim.x, par ^= 1) {
    const int b = (int)(item / J), j = (int)(item % J);
    const float cf = coef ? coef[item] : 1.f;
    float4 tb[TS][CH], d[S_][CH];
    // ---- all loads of the item in flight at once --------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < S_; ++s) {
      const float4* p4 = reinterpret_cast<const float4*>(pred + (long long)b * pB + (long long)s * pS + (long long)j * pJ);
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = t + u * blockDim.x;
        d[s][u] = (q < nq) ? ldg_stream(p4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      const float* t0 = tgt + (long long)b * tB + (long long)s * tS + (long long)j * tJ;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = t + u * blockDim.x;
        tb[s][u] = (q < nq) ? ldg_stream(reinterpret_cast<const float4*>(t0) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int m = 1; m < M; ++m) {                      // float32 mean over the teachers in index order (torch.mean)
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int q = t + u * blockDim.x;
          if (q < nq) {
            const float4 c = ldg_stream(reinterpret_cast<const float4*>(t0 + (long long)m * tM) + q);
            tb[s][u].x = __fadd_rn(tb[s][u].x, c.x); tb[s][u].y = __fadd_rn(tb[s][u].y, c.y);
            tb[s][u].z = __fadd_rn(tb[s][u].z, c.z); tb[s][u].w = __fadd_rn(tb[s][u].w, c.w);
          }
        }
      }
    }
    // ---- per-thread partials ------------------------------------------------------------------------------------
    float v[NV];
#pragma unroll
    for (int s = 0; s < TS; ++s) {
      float tmx = -INFINITY;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = t + u * blockDim.x;
        if (q < nq) {
          if (M > 1) {
            tb[s][u].x = __fdiv_rn(tb[s][u].x, fM); tb[s][u].y = __fdiv_rn(tb[s][u].y, fM);
            tb[s][u].z = __fdiv_rn(tb[s][u].z, fM); tb[s][u].w = __fdiv_rn(tb[s][u].w, fM);
          }
          tmx = fmaxf(fmaxf(tmx, fmaxf(tb[s][u].x, tb[s][u].y)), fmaxf(tb[s][u].z, tb[s][u].w));
        }
      }
      v[2 * S_ + s] = tmx;
    }
#pragma unroll
    for (int s = 0; s < S_; ++s) {
      float sse = 0.f, pmx = -INFINITY;
      const int ts = TPS ? s : 0;
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const int q = t + u * blockDim.x;
        if (q < nq) {
          const float4 pv = d[s][u];
          pmx = fmaxf(fmaxf(pmx, fmaxf(pv.x, pv.y)), fmaxf(pv.z, pv.w));
          d[s][u] = make_float4(pv.x - tb[ts][u].x, pv.y - tb[ts][u].y, pv.z - tb[ts][u].z, pv.w - tb[ts][u].w);
          sse += d[s][u].x * d[s][u].x + d[s][u].y * d[s][u].y + d[s][u].z * d[s][u].z + d[s][u].w * d[s][u].w;
        }
      }
      v[s] = sse; v[S_ + s] = pmx;
    }
    // ---- one reduction for all of them: warp trees, then the warps' partials in warp order ------------------------
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = (k < S_) ? warp_sum(v[k]) : warp_max(v[k]);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; ++k) red[par][k][w] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float a = (k < S_) ? 0.f : -INFINITY;
      for (int i = 0; i < nw; ++i) a = (k < S_) ? a + red[par][k][i] : fmaxf(a, red[par][k][i]);
      v[k] = a;
    }
    // ---- outputs ------------------------------------------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < S_; ++s) {
      const float vt = v[2 * S_ + (TPS ? s : 0)], vp = v[S_ + s];
      float mk = 1.f;
      if (mask_mode == 1) mk = (vp >= thr && vt >= thr) ? 1.f : 0.f;
      else if (mask_mode == 2) mk = (vt >= thr) ? 1.f : 0.f;
      const long long o = ((long long)b * S_ + s) * J + j;
      if (t == 0) {
        if (per_loss) per_loss[o] = (v[s] * inv_hw) * cf;
        if (mask_o) mask_o[o] = mk;
        if (vmax_p_o) vmax_p_o[o] = (mask_mode == 1 || vmax_p_o) ? vp : 0.f;
        if (vmax_t_o) vmax_t_o[o] = vt;
      }
      if (grad) {
        float4* gr = reinterpret_cast<float4*>(grad + (long long)b * gB + (long long)s * gS + (long long)j * gJ);
        const float gcoef = gs * 2.f * inv_hw * cf * mk;
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const int q = t + u * blockDim.x;
          if (q < nq) stg_stream(gr + q, make_float4(gcoef * d[s][u].x, gcoef * d[s][u].y, gcoef * d[s][u].z, gcoef * d[s][u].w));
        }
      }
    }
  }
}

__global__ void __launch_bounds__(1024) loss_finalize_kernel(const float* __restrict__ per_loss,
                                                              const float* __restrict__ mask,
                                                              const float* __restrict__ gate, long long BSJ,
                                                              long long BJ, double* out) {
  __shared__ double red[4][32];
  double s = 0.0, np = 0.0, ns = 0.0, ng = 0.0;
  for (long long i = threadIdx.x; i < BSJ; i += blockDim.x) {
    const float l = per_loss[i], m = mask ? mask[i] : 1.f;
    s += (double)l * (double)m;
    np += (l > 0.f) ? 1.0 : 0.0;
    ns += (m > 0.f) ? 1.0 : 0.0;
  }
  if (gate) { for (long long i = threadIdx.x; i < BJ; i += blockDim.x) ng += (gate[i] > 0.f) ? 1.0 : 0.0; }
  else if (threadIdx.x == 0) ng = (double)BJ;
  s = warp_sum(s); np = warp_sum(np); ns = warp_sum(ns); ng = warp_sum(ng);
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { red[0][w] = s; red[1][w] = np; red[2][w] = ns; red[3][w] = ng; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[threadIdx.x][i];
    out[threadIdx.x] = t;
  }
}

// gate_out = gate_in * visibility(kps) (process.py:262-268); count = #(gate_out > 0);
// *grad_scale = loss_weight / (S * count), or loss_weight when count == 0 (MT_UBPL.py:266).
__global__ void __launch_bounds__(1024) gate_prepare_kernel(const float* __restrict__ kps, const float* __restrict__ gate_in,
                                                             long long n, int img_h, int img_w, float stride, float sigma,
                                                             int S, float loss_weight, float* __restrict__ gate_out,
                                                             float* __restrict__ grad_scale, int32_t* __restrict__ count_out) {
  __shared__ int red[32];
  int c = 0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const Gauss g = gauss_setup(kps[2 * i], kps[2 * i + 1], img_h, img_w, stride, sigma);
    const float gt = (gate_in ? gate_in[i] : 1.f) * g.vis;
    if (gate_out) gate_out[i] = gt;
    c += (gt > 0.f) ? 1 : 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    if (count_out) *count_out = S * t;
    if (grad_scale) *grad_scale = (t > 0) ? loss_weight / (float)(S * t) : loss_weight;
  }
}

__global__ void scale_kernel(float* __restrict__ dst, const float* src, long long n, const float* __restrict__ scale) {
  const float s = *scale;
  const long long n4 = n >> 2;
  float4* d4 = reinterpret_cast<float4*>(dst);
  const float4* x4 = reinterpret_cast<const float4*>(src);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    d4[i] = v;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i] * s;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace ubpl

using namespace ubpl;

static int render_mse_impl(const float* kps, const float* gate_in, const float* sample_w, const float* pred,
                           int64_t pB, int64_t pS, int64_t pJ, float* grad, int64_t gB, int64_t gS, int64_t gJ,
                           float* target, int B, int S, int J, int H, int W, int img_h, int img_w, float stride,
                           float sigma, const float* grad_scale, const int32_t* count_in, float loss_weight,
                           float* grad_scale_out, float* gate_out, float* per_loss, double* summary, unsigned char* sum_ws,
                           void* stream) {
  UBPL_REQUIRE(kps && pred, "ubpl_render_mse: NULL pointer");
  UBPL_REQUIRE(!summary || sum_ws, "ubpl_render_mse_sum: summary needs the workspace");
  UBPL_REQUIRE(B >= 0 && S >= 1 && J >= 0 && H > 0 && W > 0 && stride > 0.f && sigma > 0.f, "ubpl_render_mse: bad arguments");
  const long long BJ = (long long)B * J;
  if (BJ == 0) return UBPL_OK;
  const int HW = H * W;
  int vec = (W % 4 == 0) && aligned16(pred) && pB % 4 == 0 && pS % 4 == 0 && pJ % 4 == 0;
  if (grad) vec = vec && aligned16(grad) && gB % 4 == 0 && gS % 4 == 0 && gJ % 4 == 0;
  if (target) vec = vec && aligned16(target) && (HW % 4 == 0);
  static bool knob_set = false;
  if (!knob_set) {
    const int m = getenv("UBPL_K3_STORE") ? atoi(getenv("UBPL_K3_STORE")) : 0;
    cudaMemcpyToSymbol(g_store_mode, &m, sizeof(int));
    knob_set = true;
  }
  FastDiv divW4;
  divW4.init((unsigned)(W >= 4 ? W / 4 : 1));
  const bool occ5 = getenv("UBPL_K3_OCC") ? atoi(getenv("UBPL_K3_OCC")) == 5 : (UBPL_K3_OCC_DEFAULT == 5);
  const int wpb = 4;    // small CTAs: one (b,j) item per warp, finer-grained tail
  const size_t smem = (size_t)wpb * (W + H) * sizeof(float);
  UBPL_REQUIRE(smem <= 48 * 1024, "ubpl_render_mse: heat-map sides too large (%d x %d)", H, W);
  const long long need = (BJ + wpb - 1) / wpb;
  // one resident wave: the kernel loops over whole rounds of grid*wpb items and shares the leftover items
  // between the warps of a CTA, so a second, partial wave of CTAs would only add a tail
  long long cap = (long long)sm_count() * (occ5 ? 5 : 6);
  // every item shared by the 4 warps of a CTA (740 CTAs each streaming one item: 2 reads + 3 writes of 16 KB) instead
  // of one item per warp (2960 warps, 5 streams each): fewer, larger concurrent streams -- 62.0 -> 56.8 us on c2
  // (profiles/README.md round 2; same HBM concurrency effect as K1's copy cap).  UBPL_K3_COOP=0 restores one warp per item.
  const int all_coop = getenv("UBPL_K3_COOP") ? atoi(getenv("UBPL_K3_COOP")) : 1;
  if (getenv("UBPL_K3_CTAS") && atoi(getenv("UBPL_K3_CTAS")) > 0) cap = (long long)sm_count() * atoi(getenv("UBPL_K3_CTAS"));
  if (cap > 4096) cap = 4096;             // UBPL_RENDER_SUM_WS_BYTES holds 4096 CTA partials
  // fewer items than one round of warps: a CTA per item (the kernel's cooperative phase) keeps 4 warps streaming
  // every item instead of one, which matters when the maps are large and the items few (fly: 128x128, B*J = 1024)
  (void)need;
  const int grid = (int)(BJ < cap ? BJ : cap);
  // launched with programmatic stream serialization (UBPL_K3_PDL=0 switches it off): see griddepcontrol.wait in the kernel
  cudaLaunchConfig_t lc;
  memset(&lc, 0, sizeof(lc));
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3((unsigned)(wpb * 32)); lc.dynamicSmemBytes = smem; lc.stream = (cudaStream_t)stream;
  cudaLaunchAttribute lattr[1];
  lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  lattr[0].val.programmaticStreamSerializationAllowed = 1;
  const bool pdl = getenv("UBPL_K3_PDL") ? atoi(getenv("UBPL_K3_PDL")) != 0 : true;
  lc.attrs = lattr; lc.numAttrs = pdl ? 1 : 0;
  // the lean kernel for the drivers' shapes (see render_mse_fast_kernel) is opt-in (UBPL_K3_FAST=1): measured on c2 it
  // is no faster than the generic kernel (57.2 vs 56.5 us back to back) although it executes ~3x fewer instructions
  // -- K3 is bound by the memory system's read/write mix, not by issue slots (profiles/README.md round 2)
  const int w4 = W / 4;
  const bool fast_ok = vec && (S == 1 || S == 2) && w4 >= 1 && w4 <= 128 && (w4 & (w4 - 1)) == 0 && (HW % 4096 == 0) &&
                       (getenv("UBPL_K3_FAST") && atoi(getenv("UBPL_K3_FAST")) != 0) && !getenv("UBPL_K3_STORE") &&
                       all_coop;
  if (fast_ok) {
    int sh = 0;
    while ((1 << sh) < w4) ++sh;
    // resident CTAs per SM the lean kernel is compiled for (UBPL_K3_FAST_OCC = 4, 5 or 6; one resident wave)
    int focc = getenv("UBPL_K3_FAST_OCC") ? atoi(getenv("UBPL_K3_FAST_OCC")) : 5;
    if (focc < 4) focc = 4;
    if (focc > 6) focc = 6;
    long long fcap = (long long)sm_count() * focc;
    if (getenv("UBPL_K3_CTAS") && atoi(getenv("UBPL_K3_CTAS")) > 0) fcap = (long long)sm_count() * atoi(getenv("UBPL_K3_CTAS"));
    if (fcap > 4096) fcap = 4096;
    lc.gridDim = dim3((unsigned)(BJ < fcap ? BJ : fcap));
    lc.dynamicSmemBytes = 2 * (size_t)(W + H) * sizeof(float);
#define UBPL_LAUNCH_FAST(SSV, OCCV)                                                                                    \
    cudaLaunchKernelEx(&lc, render_mse_fast_kernel<SSV, OCCV>, kps, gate_in, sample_w, pred, (long long)pB, (long long)pS, \
                       (long long)pJ, grad, (long long)gB, (long long)gS, (long long)gJ, target, B, J, H, W, img_h, img_w,  \
                       stride, sigma, grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, summary,   \
                       sum_ws, sh)
    if (S == 2) { if (focc == 4) UBPL_LAUNCH_FAST(2, 4); else if (focc == 5) UBPL_LAUNCH_FAST(2, 5); else UBPL_LAUNCH_FAST(2, 6); }
    else { if (focc == 4) UBPL_LAUNCH_FAST(1, 4); else if (focc == 5) UBPL_LAUNCH_FAST(1, 5); else UBPL_LAUNCH_FAST(1, 6); }
#undef UBPL_LAUNCH_FAST
    return check_launch("ubpl_render_mse");
  }
  // instances compiled for 64x64 / 128x128 maps (see the kernel; UBPL_K3_SHAPES=0: shapes read at run time)
  const int shape = (vec && occ5 && all_coop && (S == 1 || S == 2) && wpb == 4 &&
                     !(getenv("UBPL_K3_SHAPES") && atoi(getenv("UBPL_K3_SHAPES")) == 0))
                        ? ((H == 64 && W == 64) ? 64 : (H == 128 && W == 128) ? 128 : 0) : 0;
#define UBPL_LAUNCH_SHAPED(SSV, C)                                                                                     \
  cudaLaunchKernelEx(&lc, render_mse_kernel<true, SSV, 5, C, C>,                                                       \
      kps, gate_in, sample_w, pred, (long long)pB, (long long)pS, (long long)pJ, grad, (long long)gB, (long long)gS, (long long)gJ, target, B, S, J, H, W, img_h, img_w, stride, sigma, \
      grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, divW4, summary, sum_ws, all_coop)
  if (shape) {
    if (shape == 64) { if (S == 2) UBPL_LAUNCH_SHAPED(2, 64); else UBPL_LAUNCH_SHAPED(1, 64); }
    else { if (S == 2) UBPL_LAUNCH_SHAPED(2, 128); else UBPL_LAUNCH_SHAPED(1, 128); }
    return check_launch("ubpl_render_mse");
  }
#undef UBPL_LAUNCH_SHAPED
#define UBPL_LAUNCH_RENDER(V, SSV)                                                                                     \
  if (occ5) cudaLaunchKernelEx(&lc, render_mse_kernel<V, SSV, 5, 0, 0>,                                                \
      kps, gate_in, sample_w, pred, (long long)pB, (long long)pS, (long long)pJ, grad, (long long)gB, (long long)gS, (long long)gJ, target, B, S, J, H, W, img_h, img_w, stride, sigma, \
      grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, divW4, summary, sum_ws, all_coop);       \
  else cudaLaunchKernelEx(&lc, render_mse_kernel<V, SSV, 6, 0, 0>,                                                     \
      kps, gate_in, sample_w, pred, (long long)pB, (long long)pS, (long long)pJ, grad, (long long)gB, (long long)gS, (long long)gJ, target, B, S, J, H, W, img_h, img_w, stride, sigma, \
      grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, divW4, summary, sum_ws, all_coop)
  if (vec) {
    if (S % 2 == 0) UBPL_LAUNCH_RENDER(true, 2); else UBPL_LAUNCH_RENDER(true, 1);
  } else {
    if (S % 2 == 0) UBPL_LAUNCH_RENDER(false, 2); else UBPL_LAUNCH_RENDER(false, 1);
  }
#undef UBPL_LAUNCH_RENDER
  return check_launch("ubpl_render_mse");
}

extern "C" int ubpl_render_mse(const float* kps, const float* gate_in, const float* sample_w, const float* pred,
                               int64_t pB, int64_t pS, int64_t pJ, float* grad, int64_t gB, int64_t gS, int64_t gJ,
                               float* target, int B, int S, int J, int H, int W, int img_h, int img_w, float stride,
                               float sigma, const float* grad_scale, const int32_t* count_in, float loss_weight,
                               float* grad_scale_out, float* gate_out, float* per_loss, void* stream) {
  return render_mse_impl(kps, gate_in, sample_w, pred, pB, pS, pJ, grad, gB, gS, gJ, target, B, S, J, H, W, img_h, img_w,
                         stride, sigma, grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, nullptr,
                         nullptr, stream);
}

extern "C" int ubpl_render_mse_sum(const float* kps, const float* gate_in, const float* sample_w, const float* pred,
                                   int64_t pB, int64_t pS, int64_t pJ, float* grad, int64_t gB, int64_t gS, int64_t gJ,
                                   float* target, int B, int S, int J, int H, int W, int img_h, int img_w, float stride,
                                   float sigma, const float* grad_scale, const int32_t* count_in, float loss_weight,
                                   float* grad_scale_out, float* gate_out, float* per_loss, double* summary,
                                   void* sum_ws, void* stream) {
  UBPL_REQUIRE(summary && sum_ws && (reinterpret_cast<uintptr_t>(sum_ws) & 15) == 0, "ubpl_render_mse_sum: summary and a 16-byte aligned workspace are required");
  if ((long long)B * J == 0) {
    cudaError_t e = cudaMemsetAsync(summary, 0, 4 * sizeof(double), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_render_mse_sum: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
    return UBPL_OK;
  }
  return render_mse_impl(kps, gate_in, sample_w, pred, pB, pS, pJ, grad, gB, gS, gJ, target, B, S, J, H, W, img_h, img_w,
                         stride, sigma, grad_scale, count_in, loss_weight, grad_scale_out, gate_out, per_loss, summary,
                         reinterpret_cast<unsigned char*>(sum_ws), stream);
}

extern "C" int ubpl_render_targets(const float* kps, int N, int H, int W, int img_h, int img_w, float stride,
                                   float sigma, float* heatmap, float* kps_out, void* stream) {
  UBPL_REQUIRE(kps && heatmap && N >= 0 && H > 0 && W > 0 && stride > 0.f && sigma > 0.f, "ubpl_render_targets: bad arguments");
  if (N == 0) return UBPL_OK;
  const int grid = N < sm_count() * 8 ? N : sm_count() * 8;
  render_targets_kernel<<<grid, 256, (size_t)(W + H) * sizeof(float), (cudaStream_t)stream>>>(kps, N, H, W, img_h, img_w,
                                                                                              stride, sigma, heatmap, kps_out);
  return check_launch("ubpl_render_targets");
}

extern "C" int ubpl_dense_mse(const float* pred, int64_t pB, int64_t pS, int64_t pJ, const float* tgt, int M,
                              int64_t tM, int64_t tB, int64_t tS, int64_t tJ, const float* coef, int mask_mode,
                              float thr, float* grad, int64_t gB, int64_t gS, int64_t gJ, int B, int S, int J, int H,
                              int W, const float* grad_scale, float* per_loss, float* mask, float* vmax_p,
                              float* vmax_t, void* stream) {
  UBPL_REQUIRE(pred && tgt, "ubpl_dense_mse: NULL pointer");
  UBPL_REQUIRE(B >= 0 && S >= 1 && J >= 0 && H > 0 && W > 0 && M >= 1, "ubpl_dense_mse: bad dims");
  UBPL_REQUIRE(mask_mode >= 0 && mask_mode <= 2, "ubpl_dense_mse: mask_mode must be 0, 1 or 2");
  const long long BJ = (long long)B * J;
  if (BJ == 0) return UBPL_OK;
  const long long HW = (long long)H * W;
  UBPL_REQUIRE(HW <= 65536, "ubpl_dense_mse: heat-maps above 65536 texels are not supported");
  int vec = (HW % 4 == 0) && aligned16(pred) && pB % 4 == 0 && pS % 4 == 0 && pJ % 4 == 0 && aligned16(tgt) &&
            tM % 4 == 0 && tB % 4 == 0 && tS % 4 == 0 && tJ % 4 == 0;
  if (grad) vec = vec && aligned16(grad) && gB % 4 == 0 && gS % 4 == 0 && gJ % 4 == 0;
  const int nq = (int)((HW + 3) / 4);
  const int grid = (int)(BJ < (long long)sm_count() * 16 ? BJ : (long long)sm_count() * 16);
  cudaStream_t st = (cudaStream_t)stream;
  const bool fast_ok = vec && S <= 2 && nq <= 4 * 256 && (getenv("UBPL_DENSE_FAST") ? atoi(getenv("UBPL_DENSE_FAST")) != 0 : true);
  if (fast_ok) {
    int threads = ((nq + 3) / 4 + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    const int g2 = (int)(BJ < (long long)sm_count() * 8 ? BJ : (long long)sm_count() * 8);
#define UBPL_DENSE_FAST(SV, TP)                                                                                          \
    dense_mse_fast_kernel<SV, TP><<<g2, threads, 0, st>>>(pred, pB, pS, pJ, tgt, M, tM, tB, tS, tJ, coef, mask_mode, thr, grad, \
                                                         gB, gS, gJ, B, J, (int)HW, grad_scale, per_loss, mask, vmax_p, vmax_t)
    if (S == 1) { if (tS != 0) UBPL_DENSE_FAST(1, true); else UBPL_DENSE_FAST(1, false); }
    else { if (tS != 0) UBPL_DENSE_FAST(2, true); else UBPL_DENSE_FAST(2, false); }
#undef UBPL_DENSE_FAST
    return check_launch("ubpl_dense_mse");
  }
  if (nq <= 4 * 256) {
    int threads = ((nq + 3) / 4 + 31) / 32 * 32;
    if (threads < 32) threads = 32;
    dense_mse_kernel<4, 256><<<grid, threads, 0, st>>>(pred, pB, pS, pJ, tgt, M, tM, tB, tS, tJ, coef, mask_mode, thr, grad,
                                                       gB, gS, gJ, B, S, J, H, W, grad_scale, per_loss, mask, vmax_p, vmax_t, vec);
  } else if (nq <= 4 * 1024) {
    int threads = ((nq + 3) / 4 + 31) / 32 * 32;
    dense_mse_kernel<4, 1024><<<grid, threads, 0, st>>>(pred, pB, pS, pJ, tgt, M, tM, tB, tS, tJ, coef, mask_mode, thr, grad,
                                                  gB, gS, gJ, B, S, J, H, W, grad_scale, per_loss, mask, vmax_p, vmax_t, vec);
  } else {
    int threads = ((nq + 15) / 16 + 31) / 32 * 32;
    dense_mse_kernel<16, 1024><<<grid, threads, 0, st>>>(pred, pB, pS, pJ, tgt, M, tM, tB, tS, tJ, coef, mask_mode, thr, grad,
                                                   gB, gS, gJ, B, S, J, H, W, grad_scale, per_loss, mask, vmax_p, vmax_t, vec);
  }
  return check_launch("ubpl_dense_mse");
}

extern "C" int ubpl_loss_finalize(const float* per_loss, const float* mask, const float* gate, int B, int S, int J,
                                  double* out, void* stream) {
  UBPL_REQUIRE(per_loss && out && B >= 0 && S >= 1 && J >= 0, "ubpl_loss_finalize: bad arguments");
  loss_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(per_loss, mask, gate, (long long)B * S * J, (long long)B * J, out);
  return check_launch("ubpl_loss_finalize");
}

extern "C" int ubpl_gate_prepare(const float* kps, const float* gate_in, int64_t n, int img_h, int img_w, float stride,
                                 float sigma, int S, float loss_weight, float* gate_out, float* grad_scale,
                                 int32_t* count_out, void* stream) {
  UBPL_REQUIRE(kps && n >= 0 && S >= 1 && stride > 0.f && sigma > 0.f, "ubpl_gate_prepare: bad arguments");
  gate_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(kps, gate_in, n, img_h, img_w, stride, sigma, S, loss_weight,
                                                            gate_out, grad_scale, count_out);
  return check_launch("ubpl_gate_prepare");
}

extern "C" int ubpl_scale(float* dst, const float* src, int64_t n, const float* scale, void* stream) {
  UBPL_REQUIRE(dst && src && scale && n >= 0, "ubpl_scale: bad arguments");
  UBPL_REQUIRE(aligned16(dst) && aligned16(src), "ubpl_scale: buffers must be 16-byte aligned");
  if (n == 0) return UBPL_OK;
  long long blocks = (n / 4 + 255) / 256;
  if (blocks > (long long)sm_count() * 16) blocks = (long long)sm_count() * 16;
  if (blocks < 1) blocks = 1;
  scale_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, n, scale);
  return check_launch("ubpl_scale");
}
