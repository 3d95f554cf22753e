// K2: per-joint uncertainty (cross-view / cross-model dispersion) and pseudo-label selection.
//
// Reference semantics (file:line in /root/reference):
//   utils/evaluation.py:40-58    uncertainty_fromDistance
//   utils/process.py:53-68       coord_distance (python floats: ((dx)**2+(dy)**2)**0.5), coord_avgDistance
//   utils/business.py:109-161    assess_pseudo_unc2 (two teachers: intDist, extDist, ensemble weights)
//   utils/business.py:43-46,173-217  _calReliabilityThr / filter_pseudo2 (global quantile, strict >)
//   utils/business.py:237-261,375-376  pseudo_filter_mixUnc / _calUncValue (fixed threshold)
//
// The data here is tiny (K*M*B*J coordinates); everything is float64 in the reference's python
// expression order so the masks come out bit-identical.  One thread per (sample, joint).
#include "common.cuh"

namespace ubpl {

__global__ void view_dispersion_kernel(const float* __restrict__ preds, const float* __restrict__ mean_in, int K,
                                       long long BJ, float* out_mean,
                                       double* out_dist, float* out_unc32, uint8_t* out_legal, uint32_t* max_bits,
                                       int sentinel_illegal, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float u32 = 0.f;
  if (i < BJ) {
    float sx = preds[2 * i], sy = preds[2 * i + 1];
    bool legal = (sx >= 0.f) && (sy >= 0.f);
    for (int k = 1; k < K; ++k) {
      const float x = preds[2 * ((long long)k * BJ + i)], y = preds[2 * ((long long)k * BJ + i) + 1];
      sx = __fadd_rn(sx, x);
      sy = __fadd_rn(sy, y);
      legal = legal && (x >= 0.f) && (y >= 0.f);
    }
    float mx = __fdiv_rn(sx, (float)K), my = __fdiv_rn(sy, (float)K);         // torch.mean (float32)
    if (mean_in) { mx = mean_in[2 * i]; my = mean_in[2 * i + 1]; }            // caller-supplied preds_mean
    double acc = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double x = (double)preds[2 * ((long long)k * BJ + i)], y = (double)preds[2 * ((long long)k * BJ + i) + 1];
      acc = __dadd_rn(acc, py_dist(x, y, (double)mx, (double)my, T));        // sum(dists)
    }
    const double avg = __ddiv_rn(acc, (double)K);                             // / len(dists)
    u32 = (float)avg;                                                         // torch.tensor(dist_avg)
    if (out_mean) { out_mean[2 * i] = mx; out_mean[2 * i + 1] = my; }
    if (out_dist) out_dist[i] = (sentinel_illegal && !legal) ? 999.0 : avg;   // business.py:123 sentinel
    if (out_unc32) out_unc32[i] = u32;
    if (out_legal) out_legal[i] = legal ? 1 : 0;
  }
  if (max_bits) {
    const float m = warp_max(u32);                                            // distances are >= 0
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(max_bits, __float_as_uint(m));
  }
}

__global__ void unc_normalize_kernel(const float* __restrict__ unc32, const uint32_t* __restrict__ max_bits,
                                     long long n, float* out_unc, float* out_uncW) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float mx = __uint_as_float(*max_bits);
  const float u = __fdiv_rn(unc32[i], mx);            // unc / unc.max()  (0/0 -> NaN like torch)
  if (out_unc) out_unc[i] = u;
  if (out_uncW) out_uncW[i] = expf(-u);
}

__global__ void assess_dual_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                   const float* __restrict__ pmean, const float* __restrict__ a1,
                                   const float* __restrict__ a2, int K, long long BJ, double* legal_o,
                                   double* int1_o, double* int2_o, double* ext_o, double* w1_o, double* w2_o,
                                   double* coord_o, float* coord32_o, int32_t* zero_div, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BJ) return;
  const double x1 = p1[2 * i], y1 = p1[2 * i + 1], x2 = p2[2 * i], y2 = p2[2 * i + 1];
  const bool ori_legal = (x1 >= 0 && y1 >= 0) && (x2 >= 0 && y2 >= 0);
  bool g1 = true, g2 = true;
  for (int k = 0; k < K; ++k) {
    const long long o = 2 * ((long long)k * BJ + i);
    g1 = g1 && (a1[o] >= 0.f) && (a1[o + 1] >= 0.f);
    g2 = g2 && (a2[o] >= 0.f) && (a2[o + 1] >= 0.f);
  }
  double legal = ori_legal ? 1.0 : 0.0, w1 = 0.5, w2 = 0.5, d1 = 999.0, d2 = 999.0, ext = 999.0;
  // third pred-set: bus.preds_mean(p1, p2) = torch.mean(stack([p1, p2], -1), -1) in float32 (business.py:297-300)
  double cx = pmean ? (double)pmean[2 * i] : (double)__fdiv_rn(__fadd_rn(p1[2 * i], p2[2 * i]), 2.f);
  double cy = pmean ? (double)pmean[2 * i + 1] : (double)__fdiv_rn(__fadd_rn(p1[2 * i + 1], p2[2 * i + 1]), 2.f);
  if (ori_legal && g1 && g2) {
    // coord_avgDistance: itertools.combinations order, sequential float64 sum, / count
    double s1 = 0.0, s2 = 0.0;
    int cnt = 0;
    for (int u = 0; u < K; ++u)
#pragma unroll 4
      for (int v = u + 1; v < K; ++v) {
        const long long ou = 2 * ((long long)u * BJ + i), ov = 2 * ((long long)v * BJ + i);
        s1 = __dadd_rn(s1, py_dist(a1[ou], a1[ou + 1], a1[ov], a1[ov + 1], T));
        s2 = __dadd_rn(s2, py_dist(a2[ou], a2[ou + 1], a2[ov], a2[ov + 1], T));
        ++cnt;
      }
    d1 = __ddiv_rn(s1, (double)cnt);      // K < 2: 0/0 -> NaN (the reference raises ZeroDivisionError)
    d2 = __ddiv_rn(s2, (double)cnt);
    const double den = __dadd_rn(d1, d2);
    if (den == 0.0) {
      atomicAdd(zero_div, 1);             // business.py:135 divides by zero here
    } else {
      w1 = __ddiv_rn(d1, den);
      w2 = __ddiv_rn(d2, den);
    }
    cx = __dadd_rn(__dmul_rn(w1, x1), __dmul_rn(w2, x2));
    cy = __dadd_rn(__dmul_rn(w1, y1), __dmul_rn(w2, y2));
    legal = 1.0;
    double se = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const long long o = 2 * ((long long)k * BJ + i);
      se = __dadd_rn(se, py_dist(a1[o], a1[o + 1], a2[o], a2[o + 1], T));
    }
    ext = __ddiv_rn(se, (double)K);
  }
  if (legal_o) legal_o[i] = legal;
  if (int1_o) int1_o[i] = d1;
  if (int2_o) int2_o[i] = d2;
  if (ext_o) ext_o[i] = ext;
  if (w1_o) w1_o[i] = w1;
  if (w2_o) w2_o[i] = w2;
  if (coord_o) { coord_o[2 * i] = cx; coord_o[2 * i + 1] = cy; }
  if (coord32_o) { coord32_o[2 * i] = (float)cx; coord32_o[2 * i + 1] = (float)cy; }
}

// ---------------------------------------------------------------------------------------------
// selection
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t key_of(double v) {
  const uint64_t b = (uint64_t)__double_as_longlong(v);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double value_of(uint64_t k) {
  const uint64_t b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__global__ void extrema_init_kernel(double* ext) {
  ext[0] = 0.0;     // dist_max over dist < 999 (business.py:176)
  ext[1] = 999.0;   // dist_min
}
__global__ void extrema_kernel(const double* __restrict__ dist, long long n, double* ext) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double e = dist[i];
  // distances are >= 0, so the raw bit pattern orders like the value
  if (e > 0.0 && e < 999.0) atomicMax(reinterpret_cast<unsigned long long*>(ext), (unsigned long long)__double_as_longlong(e));
  if (e >= 0.0 && e < 999.0) atomicMin(reinterpret_cast<unsigned long long*>(ext + 1), (unsigned long long)__double_as_longlong(e));
}

__global__ void reliability_kernel(const double* __restrict__ dist, const double* __restrict__ legal, long long n,
                                   const double* __restrict__ ext, double reliableDistMin, double* rel,
                                   uint64_t* keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double dmax = ext[0], dmin = ext[1];
  if (dmax == 0.0) dmax = 999.0;                      // business.py:181
  if (dmin > reliableDistMin) dmin = reliableDistMin; // business.py:182
  const double e = dist[i];
  const double e2 = (e != 999.0) ? e : dmax;
  const double unc = (legal[i] > 0.0) ? __ddiv_rn(__dsub_rn(e2, dmin), __dsub_rn(dmax, dmin)) : 1.0;
  const double r = __dsub_rn(1.0, unc);
  rel[i] = r;
  if (keys) keys[i] = key_of(r);
}

__global__ void key_hist_kernel(const uint64_t* __restrict__ keys, long long n, const uint64_t* __restrict__ prefix,
                                int shift, uint32_t* hist, int bits = 16) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = keys[i];
  if (shift + bits < 64) {
    const uint64_t pf = *prefix;
    if ((k >> (shift + bits)) != (pf >> (shift + bits))) return;
  }
  atomicAdd(hist + (uint32_t)((k >> shift) & ((1ull << bits) - 1ull)), 1u);
}

// One CTA of 32 warps walks the nb = 2^bits bins (1024 <= nb <= 65536) from the TOP (largest keys first)
// and finds the bin that holds rank *k_rem: coalesced reads (a warp sums 32 consecutive bins per step),
// then three short parallel scans (warp totals -> 32-bin groups of the owning warp -> bins of the owning
// group).  zero_after != 0 clears the histogram for the next pass once every thread has read it.
__global__ void __launch_bounds__(1024) select_descend_kernel(uint32_t* __restrict__ hist, int shift, uint64_t* prefix,
                                                               long long* k_rem, int zero_after, int bits = 16) {
  __shared__ unsigned part[2048];            // sums of 32-bin groups, descending order
  __shared__ unsigned long long wtot[32];
  __shared__ int s_w;
  __shared__ unsigned long long s_acc;
  const int t = threadIdx.x, w = t >> 5, lane = t & 31;
  const int nb = 1 << bits, span = nb >> 5, gpw = span >> 5;      // bins per warp, 32-bin groups per warp
  unsigned long long wsum = 0;
  for (int g = 0; g < gpw; ++g) {
    const int d = w * span + g * 32 + lane;            // descending bin index
    unsigned c = hist[nb - 1 - d];
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) part[w * gpw + g] = c;
    wsum += c;
  }
  if (lane == 0) wtot[w] = wsum;
  __syncthreads();
  const unsigned long long k = (unsigned long long)*k_rem;
  if (t == 0) {
    unsigned long long acc = 0;
    int W = 31;
    for (int u = 0; u < 32; ++u) {
      if (acc + wtot[u] > k) { W = u; break; }
      acc += wtot[u];
    }
    if (W == 31 && !(acc + wtot[31] > k)) {             // rank beyond the population: clamp to the lowest bins
      acc = 0;
      for (int u = 0; u < 31; ++u) acc += wtot[u];
    }
    s_w = W; s_acc = acc;
  }
  __syncthreads();
  if (w == 0) {
    const int W = s_w;
    const unsigned long long acc = s_acc;
    const int per = (gpw + 31) >> 5;                     // groups per lane (1, or 2 when gpw = 64)
    unsigned p0 = 0, p1 = 0;
    if (lane * per < gpw) p0 = part[W * gpw + lane * per];
    if (per == 2 && lane * per + 1 < gpw) p1 = part[W * gpw + lane * per + 1];
    const unsigned long long pair = (unsigned long long)p0 + p1;
    unsigned long long incl = pair;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned long long excl = incl - pair;
    const bool hit = (acc + excl + pair > k);
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    const int L = m ? (__ffs(m) - 1) : min(31, (gpw - 1) / per);
    const unsigned long long base = acc + __shfl_sync(0xffffffffu, excl, L);
    const unsigned q0 = __shfl_sync(0xffffffffu, p0, L);
    int G = L * per;
    unsigned long long acc2 = base;
    if (per == 2 && !(base + q0 > k)) { G = L * per + 1; acc2 = base + q0; }
    // the 32 bins of group G
    const int d = W * span + G * 32 + lane;
    const unsigned c = hist[nb - 1 - d];
    unsigned long long ci = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long v = __shfl_up_sync(0xffffffffu, ci, o);
      if (lane >= o) ci += v;
    }
    const unsigned long long ce = ci - c;
    const bool hit2 = (acc2 + ce + c > k);
    const unsigned m2 = __ballot_sync(0xffffffffu, hit2);
    const int B = m2 ? (__ffs(m2) - 1) : 31;
    if (lane == B) {
      const uint64_t bin = (uint64_t)(nb - 1 - d);
      const uint64_t mask = ~(((1ull << bits) - 1ull) << shift);
      *prefix = ((shift + bits < 64 ? *prefix : 0ull) & mask) | (bin << shift);
      *k_rem = (long long)(k - (acc2 + ce));
    }
  }
  if (zero_after) {
    __syncthreads();
    for (int i = t; i < nb; i += 1024) hist[i] = 0u;
  }
}

// ---- selection on the DISTANCE keys (multi-GPU path) ------------------------------------------------
// reliability = 1 - (e2 - dmin)/(dmax - dmin) is a monotone non-increasing function of e2 (every IEEE step
// is monotone), illegal and sentinel items have reliability exactly 0 = f(dmax), so the k-th LARGEST
// reliability is f(k-th SMALLEST e') with e' = dist for legal non-sentinel items and +inf otherwise.  The
// radix select can therefore run on keys that do not depend on the global extrema, and the extrema
// all-reduce rides in the same NCCL group as the first histogram.  key' = ~key_of(e') turns "k-th smallest"
// into the "k-th largest" the descend kernel finds.
__global__ void dist_keys_extrema_kernel(const double* __restrict__ dist, const double* __restrict__ legal, long long n,
                                         uint64_t* __restrict__ keys, double* ext) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double e = dist[i];
  if (e > 0.0 && e < 999.0) atomicMax(reinterpret_cast<unsigned long long*>(ext), (unsigned long long)__double_as_longlong(e));
  if (e >= 0.0 && e < 999.0) atomicMin(reinterpret_cast<unsigned long long*>(ext + 1), (unsigned long long)__double_as_longlong(e));
  const double ep = (legal[i] > 0.0 && e != 999.0) ? e : __longlong_as_double(0x7ff0000000000000ll);
  keys[i] = ~key_of(ep);
}

__global__ void apply_dist_kernel(const double* __restrict__ dist, const double* __restrict__ legal, long long n, int J,
                                  const uint64_t* __restrict__ prefix, const double* __restrict__ ext,
                                  double reliableDistMin, double reliableThr, double* rel, uint8_t* enable,
                                  float* gate32, int32_t* counts, double* thr_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  double dmax = ext[0], dmin = ext[1];
  if (dmax == 0.0) dmax = 999.0;                      // business.py:181
  if (dmin > reliableDistMin) dmin = reliableDistMin; // business.py:182
  const double ek = value_of(~(*prefix));              // k-th smallest e'
  const double den = __dsub_rn(dmax, dmin);
  const double rk = (ek > 1.0e300) ? 0.0 : __dsub_rn(1.0, __ddiv_rn(__dsub_rn(ek, dmin), den));
  const double thr = (rk > reliableThr) ? rk : reliableThr;
  if (i == 0 && thr_out) *thr_out = thr;
  if (i >= n) return;
  const double e = dist[i];
  const double e2 = (e != 999.0) ? e : dmax;
  const double unc = (legal[i] > 0.0) ? __ddiv_rn(__dsub_rn(e2, dmin), den) : 1.0;
  const double r = __dsub_rn(1.0, unc);
  rel[i] = r;
  const bool en = r > thr;
  if (enable) enable[i] = en ? 1 : 0;
  if (gate32) gate32[i] = en ? 1.f : 0.f;
  if (en && counts) {
    atomicAdd(counts + (int)(i % J), 1);
    atomicAdd(counts + J, 1);
  }
}

__global__ void select_apply_kernel(const double* __restrict__ rel, long long n, int J,
                                    const uint64_t* __restrict__ prefix, double reliableThr, uint8_t* enable,
                                    float* gate32, int32_t* counts, double* thr_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const double kth = value_of(*prefix);
  const double thr = (kth > reliableThr) ? kth : reliableThr;     // max(args.reliableThr, scores[k])
  if (i == 0 && thr_out) *thr_out = thr;
  if (i >= n) return;
  const bool en = rel[i] > thr;
  if (enable) enable[i] = en ? 1 : 0;
  if (gate32) gate32[i] = en ? 1.f : 0.f;
  if (en && counts) {
    atomicAdd(counts + (int)(i % J), 1);
    atomicAdd(counts + J, 1);
  }
}

__global__ void select_fixed_kernel(const double* __restrict__ dist, const double* __restrict__ legal, long long n,
                                    int J, double distThrMax, uint8_t* enable, float* gate32, int32_t* counts,
                                    double* unc_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double thr = __dsub_rn(1.0, exp(-__ddiv_rn(__dmul_rn(distThrMax, 3.0), 5.0)));   // _calUncValue(distThrMax*3)
  const double unc = __dsub_rn(1.0, exp(-__ddiv_rn(dist[i], 5.0)));
  const bool en = (legal == nullptr || legal[i] > 0.0) && (unc <= thr);
  if (enable) enable[i] = en ? 1 : 0;
  if (gate32) gate32[i] = en ? 1.f : 0.f;
  if (unc_out) unc_out[i] = unc;
  if (en && counts) {
    atomicAdd(counts + (int)(i % J), 1);
    atomicAdd(counts + J, 1);
  }
}

// Fused K2 for the mean-teacher fixed-threshold path, one launch, one thread per (sample, joint):
// dispersion (evaluation.py:44-54) -> unc = 1-exp(-d/5) <= 1-exp(-3*distThrMax/5) (business.py:237-261,
// 375-376) -> gate = enable * visibility (process.py:262-268) -> count += S per open gate (losses.py:29).
__global__ void __launch_bounds__(128) k2_view_fixed_kernel(const float* __restrict__ preds, int K, long long BJ, int J,
                                                             double distThrMax, int img_h, int img_w, float stride,
                                                             float sigma, int S, float* out_mean, double* out_dist,
                                                             uint8_t* out_legal, uint8_t* enable, float* gate_out,
                                                             int32_t* count_out, int32_t* counts, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int open_gate = 0, sel = 0;
  if (i < BJ) {
    const double thr = __dsub_rn(1.0, exp(-__ddiv_rn(__dmul_rn(distThrMax, 3.0), 5.0)));
    float sx = preds[2 * i], sy = preds[2 * i + 1];
    bool legal = (sx >= 0.f) && (sy >= 0.f);
    for (int k = 1; k < K; ++k) {
      const float x = preds[2 * ((long long)k * BJ + i)], y = preds[2 * ((long long)k * BJ + i) + 1];
      sx = __fadd_rn(sx, x);
      sy = __fadd_rn(sy, y);
      legal = legal && (x >= 0.f) && (y >= 0.f);
    }
    const float mx = __fdiv_rn(sx, (float)K), my = __fdiv_rn(sy, (float)K);
    double acc = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const double x = (double)preds[2 * ((long long)k * BJ + i)], y = (double)preds[2 * ((long long)k * BJ + i) + 1];
      acc = __dadd_rn(acc, py_dist(x, y, (double)mx, (double)my, T));
    }
    const double dist = legal ? __ddiv_rn(acc, (double)K) : 999.0;
    const double unc = __dsub_rn(1.0, exp(-__ddiv_rn(dist, 5.0)));
    const bool en = legal && (unc <= thr);
    const Gauss g = gauss_setup(mx, my, img_h, img_w, stride, sigma);
    const float gt = (en ? 1.f : 0.f) * g.vis;
    if (out_mean) { out_mean[2 * i] = mx; out_mean[2 * i + 1] = my; }
    if (out_dist) out_dist[i] = dist;
    if (out_legal) out_legal[i] = legal ? 1 : 0;
    if (enable) enable[i] = en ? 1 : 0;
    gate_out[i] = gt;
    if (en) atomicAdd(counts + (int)(i % J), 1);
    sel = en ? 1 : 0;
    open_gate = (gt > 0.f) ? 1 : 0;
  }
  open_gate = __reduce_add_sync(0xffffffffu, open_gate);
  sel = __reduce_add_sync(0xffffffffu, sel);
  if ((threadIdx.x & 31) == 0) {
    if (open_gate) atomicAdd(count_out, S * open_gate);
    if (sel) atomicAdd(counts + J, sel);
  }
}

// error / PCK flag of predictions against ground truth (business.py:37-40, evaluation.py:78-89):
// error = dist(pred, gt); norm_b = dist(gt[b, ref0], gt[b, ref1]); acc = error / norm < pck_thr.
__global__ void coord_error_kernel(const double* __restrict__ pred, const float* __restrict__ gt, int gt_stride,
                                   long long n_sets, int B, int J, int ref0, int ref1, double pck_thr,
                                   double* err, int32_t* acc, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long BJ = (long long)B * J;
  if (i >= n_sets * BJ) return;
  const long long bj = i % BJ;
  const int b = (int)(bj / J);
  const float* g = gt + bj * gt_stride;
  const float* g0 = gt + ((long long)b * J + ref0) * gt_stride;
  const float* g1 = gt + ((long long)b * J + ref1) * gt_stride;
  const double e = py_dist(pred[2 * i], pred[2 * i + 1], (double)g[0], (double)g[1], T);
  const double norm = py_dist((double)g0[0], (double)g0[1], (double)g1[0], (double)g1[1], T);
  err[i] = e;
  acc[i] = (__ddiv_rn(e, norm) < pck_thr) ? 1 : 0;
}

// Single-GPU quantile selection in ONE launch (one CTA): extrema -> reliability -> exact k-th order
// statistic (8 passes of 8 bits over the monotone keys, 256-bin shared-memory histogram) -> masks.
// Same arithmetic as the multi-kernel path (ubpl_dist_extrema ... ubpl_select_apply), which remains the
// multi-GPU path because the histograms must be all-reduced between the passes.
__global__ void __launch_bounds__(1024) select_quantile_local_kernel(const double* __restrict__ dist,
                                                                      const double* __restrict__ legal, long long n,
                                                                      int J, long long k_rank, double reliableThr,
                                                                      double reliableDistMin, double* rel,
                                                                      uint64_t* keys, uint8_t* enable, float* gate32,
                                                                      int32_t* counts, double* thr_out, double* ext_out) {
  __shared__ unsigned long long s_ext[2];
  __shared__ unsigned int s_hist[256];
  __shared__ unsigned long long s_prefix;
  __shared__ long long s_krem;
  const int t = threadIdx.x;
  if (t == 0) { s_ext[0] = 0ull; s_ext[1] = (unsigned long long)__double_as_longlong(999.0); s_prefix = 0ull; s_krem = k_rank; }
  for (int j = t; j <= J; j += blockDim.x) counts[j] = 0;
  __syncthreads();
  for (long long i = t; i < n; i += blockDim.x) {
    const double e = dist[i];
    if (e > 0.0 && e < 999.0) atomicMax(&s_ext[0], (unsigned long long)__double_as_longlong(e));
    if (e >= 0.0 && e < 999.0) atomicMin(&s_ext[1], (unsigned long long)__double_as_longlong(e));
  }
  __syncthreads();
  double dmax = __longlong_as_double((long long)s_ext[0]), dmin = __longlong_as_double((long long)s_ext[1]);
  if (t == 0 && ext_out) { ext_out[0] = dmax; ext_out[1] = dmin; }
  if (dmax == 0.0) dmax = 999.0;
  if (dmin > reliableDistMin) dmin = reliableDistMin;
  for (long long i = t; i < n; i += blockDim.x) {
    const double e = dist[i];
    const double e2 = (e != 999.0) ? e : dmax;
    const double unc = (legal[i] > 0.0) ? __ddiv_rn(__dsub_rn(e2, dmin), __dsub_rn(dmax, dmin)) : 1.0;
    const double r = __dsub_rn(1.0, unc);
    rel[i] = r;
    keys[i] = key_of(r);
  }
  __syncthreads();
  for (int shift = 56; shift >= 0; shift -= 8) {
    if (t < 256) s_hist[t] = 0u;
    __syncthreads();
    const unsigned long long pf = s_prefix;
    for (long long i = t; i < n; i += blockDim.x) {
      const uint64_t k = keys[i];
      if (shift == 56 || (k >> (shift + 8)) == (pf >> (shift + 8))) atomicAdd(&s_hist[(unsigned)((k >> shift) & 0xffu)], 1u);
    }
    __syncthreads();
    if (t == 0) {
      long long acc = 0, k = s_krem;
      int bin = 0;
      for (int b = 255; b >= 0; --b) {
        const long long c = s_hist[b];
        bin = b;
        if (acc + c > k) break;
        acc += c;
      }
      s_prefix = pf | ((unsigned long long)bin << shift);
      s_krem = k - acc;
    }
    __syncthreads();
  }
  const double kth = value_of(s_prefix);
  const double thr = (kth > reliableThr) ? kth : reliableThr;
  if (t == 0 && thr_out) *thr_out = thr;
  for (long long i = t; i < n; i += blockDim.x) {
    const bool en = rel[i] > thr;
    if (enable) enable[i] = en ? 1 : 0;
    if (gate32) gate32[i] = en ? 1.f : 0.f;
    if (en) { atomicAdd(counts + (int)(i % J), 1); atomicAdd(counts + J, 1); }
  }
}

__global__ void pair_distance_kernel(const double* __restrict__ c1, const double* __restrict__ c2, long long n,
                                     double* __restrict__ out, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = py_dist(c1[2 * i], c1[2 * i + 1], c2[2 * i], c2[2 * i + 1], T);
}

// ---------------------------------------------------------------------------------------------
// a13: the mixed-distance uncertainty of two teachers (utils/business.py:220-234, 302-346, 378-406)
// ---------------------------------------------------------------------------------------------
// One thread per key point.  Per teacher m: error = dist(pred_m, gt), PCK flag against the per-sample norm
// dist(gt[ref0], gt[ref1]) (evaluation.py:78-89), score = clamp01(scores_m[0][j]) (the reference reads batch
// row 0, business.py:310-311), coord_aug = python-float mean of the A augmented views, intDist = mean pairwise
// distance of the views in itertools.combinations order (process.py:57-68).  Shared: extDist = dist(pred_1,
// pred_2), aExtDist = dist(coord_aug_1, coord_aug_2).  Distances with integer radicands reproduce CPython's
// libm pow bit for bit (PowTab); coord_aug means are not integers, so aExtDist (and error when gt is
// fractional) is the IEEE sqrt, at most one ulp from CPython -- the drop-in recomputes those two on the host.
__global__ void mix_dists_kernel(const float* __restrict__ gt, int gt_stride, int ref0, int ref1, double pck_thr,
                                 const float* __restrict__ p1, const float* __restrict__ p2,
                                 const float* __restrict__ s1, const float* __restrict__ s2,
                                 const float* __restrict__ a1, const float* __restrict__ a2, int B, int J, int A,
                                 double* err1, double* err2, int32_t* acc1, int32_t* acc2, double* score1,
                                 double* score2, double* caug1, double* caug2, double* int1, double* int2,
                                 double* ext, double* aext, PowTab T) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)B * J;
  if (i >= n) return;
  const int b = (int)(i / J), j = (int)(i % J);
  double norm = 1.0;
  if (gt) {
    const float* g0 = gt + ((long long)b * J + ref0) * gt_stride;
    const float* g1 = gt + ((long long)b * J + ref1) * gt_stride;
    norm = py_dist((double)g0[0], (double)g0[1], (double)g1[0], (double)g1[1], T);
  }
  double cx[2], cy[2];
  for (int m = 0; m < 2; ++m) {
    const float* p = m ? p2 : p1;
    const float* a = (m ? a2 : a1) + i * A * 2;
    const float* sc = m ? s2 : s1;
    if (gt) {
      const float* g = gt + i * gt_stride;
      const double e = py_dist((double)p[2 * i], (double)p[2 * i + 1], (double)g[0], (double)g[1], T);
      (m ? err2 : err1)[i] = e;
      (m ? acc2 : acc1)[i] = (__ddiv_rn(e, norm) < pck_thr) ? 1 : 0;
    }
    if (sc) { const double v = (double)sc[j]; (m ? score2 : score1)[i] = fmax(0.0, fmin(1.0, v)); }
    double sx = 0.0, sy = 0.0;
    for (int k = 0; k < A; ++k) { sx = __dadd_rn(sx, (double)a[2 * k]); sy = __dadd_rn(sy, (double)a[2 * k + 1]); }
    cx[m] = __ddiv_rn(sx, (double)A); cy[m] = __ddiv_rn(sy, (double)A);
    (m ? caug2 : caug1)[2 * i] = cx[m]; (m ? caug2 : caug1)[2 * i + 1] = cy[m];
    double s = 0.0;
    int cnt = 0;
    for (int u = 0; u < A; ++u)
      for (int v = u + 1; v < A; ++v) {
        s = __dadd_rn(s, py_dist((double)a[2 * u], (double)a[2 * u + 1], (double)a[2 * v], (double)a[2 * v + 1], T));
        ++cnt;
      }
    (m ? int2 : int1)[i] = __ddiv_rn(s, (double)cnt);        // A < 2: 0/0 = NaN (the reference raises ZeroDivisionError)
  }
  ext[i] = py_dist((double)p1[2 * i], (double)p1[2 * i + 1], (double)p2[2 * i], (double)p2[2 * i + 1], T);
  aext[i] = py_dist(cx[0], cy[0], cx[1], cy[1], T);
}

// The stateful half: push (intDist, extDist, aExtDist) into this key point's 3-deep history, take the 0.5/0.3/0.2
// moving averages (business.py:395-405), mixDist (:327), the three <= distThrMax tests on the averages and
// unc = 1-exp(-mixDist/5) or 999 (:343); then the fixed rule of pseudo_filter_mixUnc (:237-261): enable =
// unc <= 1-exp(-3*distThrMax/5), with the score gate of pseudo_filter_mixUnc2 (:264-268) when score != NULL
// (unc = 999 where score < *score_thr).  hist [3][n][3] float64 (newest last), hist_len [n] int32.
__global__ void mix_unc_kernel(const double* __restrict__ intd, const double* __restrict__ extd,
                               const double* __restrict__ aextd, long long n, int J, double distThrMax, double* hist,
                               int32_t* hist_len, const double* __restrict__ score, const double* __restrict__ score_thr,
                               double* lma_out, double* mix_out, double* unc_out, uint8_t* enable, float* gate32,
                               int32_t* counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double cur[3] = {intd[i], extd[i], aextd[i]};
  const int len = hist_len[i];
  double l[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    double* h = hist + ((long long)q * n + i) * 3;          // h[2] newest
    const double h1 = h[2], h2 = h[1];                      // previous newest, the one before
    h[0] = h2; h[1] = h1; h[2] = cur[q];
    if (len == 0) l[q] = cur[q];
    else if (len == 1) l[q] = __dadd_rn(__dmul_rn(cur[q], 0.5 + 0.3), __dmul_rn(h1, 0.2));
    else l[q] = __dadd_rn(__dadd_rn(__dmul_rn(cur[q], 0.5), __dmul_rn(h1, 0.3)), __dmul_rn(h2, 0.2));
    if (lma_out) lma_out[(long long)q * n + i] = l[q];
  }
  hist_len[i] = len < 2 ? len + 1 : 2;
  const double mix = __dadd_rn(l[0], (l[1] > 0.0) ? __ddiv_rn(__dadd_rn(l[1], l[2]), 2.0) : l[2]);
  const bool ok = (l[0] <= distThrMax) && (l[1] <= distThrMax) && (l[2] <= distThrMax);
  double unc = ok ? __dsub_rn(1.0, exp(-__ddiv_rn(mix, 5.0))) : 999.0;
  if (score && score[i] < *score_thr) unc = 999.0;
  if (mix_out) mix_out[i] = mix;
  if (unc_out) unc_out[i] = unc;
  const double thr = __dsub_rn(1.0, exp(-__ddiv_rn(__dmul_rn(distThrMax, 3.0), 5.0)));
  const bool en = unc <= thr;
  if (enable) enable[i] = en ? 1 : 0;
  if (gate32) gate32[i] = en ? 1.f : 0.f;
  if (en && counts) { atomicAdd(counts + (int)(i % J), 1); atomicAdd(counts + J, 1); }
}

static inline int blocks_for(long long n, int t) { return (int)((n + t - 1) / t); }

}  // namespace ubpl

using namespace ubpl;

#define GET_POWTAB(T)                                                            \
  PowTab T;                                                                      \
  {                                                                              \
    int rc_ = pow_table(&T.key, &T.val, &T.bits, &T.n, &T.rmax);                           \
    if (rc_ != UBPL_OK) return rc_;                                               \
  }

extern "C" int ubpl_view_dispersion(const float* preds, const float* mean_in, int K, int B, int J, float* out_mean,
                                    double* out_dist,
                                    float* out_unc32, uint8_t* out_legal, uint32_t* max_bits, int sentinel_illegal,
                                    void* stream) {
  UBPL_REQUIRE(preds != nullptr && K >= 1 && B >= 0 && J >= 0, "ubpl_view_dispersion: bad arguments");
  const long long BJ = (long long)B * J;
  if (BJ == 0) return UBPL_OK;
  GET_POWTAB(T);
  view_dispersion_kernel<<<blocks_for(BJ, 128), 128, 0, (cudaStream_t)stream>>>(preds, mean_in, K, BJ, out_mean, out_dist,
                                                                                 out_unc32, out_legal, max_bits, sentinel_illegal, T);
  return check_launch("ubpl_view_dispersion");
}

extern "C" int ubpl_unc_normalize(const float* unc32, const uint32_t* max_bits, int64_t n, float* out_unc,
                                  float* out_uncW, void* stream) {
  UBPL_REQUIRE(unc32 && max_bits && n >= 0, "ubpl_unc_normalize: bad arguments");
  if (n == 0) return UBPL_OK;
  unc_normalize_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(unc32, max_bits, n, out_unc, out_uncW);
  return check_launch("ubpl_unc_normalize");
}

extern "C" int ubpl_assess_dual(const float* p1, const float* p2, const float* pmean, const float* aug1,
                                const float* aug2, int K, int B, int J, double* legal, double* intDist1,
                                double* intDist2, double* extDist, double* w1, double* w2, double* coord,
                                float* coord32, int32_t* zero_div, void* stream) {
  UBPL_REQUIRE(p1 && p2 && aug1 && aug2 && zero_div, "ubpl_assess_dual: NULL pointer");
  UBPL_REQUIRE(K >= 1 && B >= 0 && J >= 0, "ubpl_assess_dual: bad dims");
  const long long BJ = (long long)B * J;
  if (BJ == 0) return UBPL_OK;
  GET_POWTAB(T);
  assess_dual_kernel<<<blocks_for(BJ, 128), 128, 0, (cudaStream_t)stream>>>(p1, p2, pmean, aug1, aug2, K, BJ, legal,
                                                                             intDist1, intDist2, extDist, w1, w2,
                                                                             coord, coord32, zero_div, T);
  return check_launch("ubpl_assess_dual");
}

extern "C" int ubpl_dist_extrema(const double* dist, int64_t n, double* ext, void* stream) {
  UBPL_REQUIRE(ext && (dist || n == 0) && n >= 0, "ubpl_dist_extrema: bad arguments");
  extrema_init_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ext);
  if (n > 0) extrema_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dist, n, ext);
  return check_launch("ubpl_dist_extrema");
}

extern "C" int ubpl_reliability(const double* dist, const double* legal, int64_t n, const double* ext,
                                double reliableDistMin, double* reliability, uint64_t* keys, void* stream) {
  UBPL_REQUIRE(dist && legal && ext && reliability && n >= 0, "ubpl_reliability: bad arguments");
  if (n == 0) return UBPL_OK;
  reliability_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dist, legal, n, ext, reliableDistMin,
                                                                            reliability, keys);
  return check_launch("ubpl_reliability");
}

extern "C" int ubpl_key_histogram(const uint64_t* keys, int64_t n, const uint64_t* prefix, int shift,
                                  uint32_t* hist, int clear_first, void* stream) {
  UBPL_REQUIRE(hist && (keys || n == 0) && n >= 0, "ubpl_key_histogram: bad arguments");
  UBPL_REQUIRE(shift == 48 || shift == 32 || shift == 16 || shift == 0, "ubpl_key_histogram: shift must be 48/32/16/0");
  UBPL_REQUIRE(shift == 48 || prefix != nullptr, "ubpl_key_histogram: prefix is NULL");
  if (clear_first) {
    cudaError_t e = cudaMemsetAsync(hist, 0, 65536 * sizeof(uint32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_key_histogram: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  }
  if (n > 0) key_hist_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(keys, n, prefix, shift, hist);
  return check_launch("ubpl_key_histogram");
}

extern "C" int ubpl_select_descend(uint32_t* hist, int shift, uint64_t* prefix, int64_t* k_rem, int zero_after,
                                   void* stream) {
  UBPL_REQUIRE(hist && prefix && k_rem, "ubpl_select_descend: NULL pointer");
  UBPL_REQUIRE(shift == 48 || shift == 32 || shift == 16 || shift == 0, "ubpl_select_descend: bad shift");
  select_descend_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(hist, shift, prefix, reinterpret_cast<long long*>(k_rem), zero_after);
  return check_launch("ubpl_select_descend");
}

extern "C" int ubpl_select_apply(const double* reliability, int64_t n, int J, const uint64_t* prefix,
                                 double reliableThr, uint8_t* enable, float* gate32, int32_t* counts,
                                 double* thr_out, void* stream) {
  UBPL_REQUIRE(reliability && prefix && (enable || gate32) && n >= 0 && J >= 1, "ubpl_select_apply: bad arguments");
  if (counts) {
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(J + 1) * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_select_apply: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  }
  const long long nn = n > 0 ? n : 1;
  select_apply_kernel<<<blocks_for(nn, 256), 256, 0, (cudaStream_t)stream>>>(reliability, n, J, prefix, reliableThr,
                                                                             enable, gate32, counts, thr_out);
  return check_launch("ubpl_select_apply");
}

extern "C" int ubpl_select_fixed(const double* dist, const double* legal, int64_t n, int J, double distThrMax,
                                 uint8_t* enable, float* gate32, int32_t* counts, double* unc_out, void* stream) {
  UBPL_REQUIRE(dist && (enable || gate32) && n >= 0 && J >= 1, "ubpl_select_fixed: bad arguments");
  if (counts) {
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(J + 1) * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_select_fixed: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  }
  if (n == 0) return UBPL_OK;
  select_fixed_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dist, legal, n, J, distThrMax, enable,
                                                                            gate32, counts, unc_out);
  return check_launch("ubpl_select_fixed");
}

extern "C" int ubpl_k2_view_fixed(const float* preds, int K, int B, int J, double distThrMax, int img_h, int img_w,
                                  float stride, float sigma, int S, float* out_mean, double* out_dist,
                                  uint8_t* out_legal, uint8_t* enable, float* gate_out, int32_t* count_out,
                                  int32_t* counts, void* stream) {
  UBPL_REQUIRE(preds && gate_out && counts && count_out && K >= 1 && B >= 0 && J >= 1 && S >= 1 && stride > 0.f && sigma > 0.f,
               "ubpl_k2_view_fixed: bad arguments");
  UBPL_REQUIRE(counts + J + 1 == count_out, "ubpl_k2_view_fixed: count_out must directly follow counts[J+1]");
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(J + 2) * sizeof(int32_t), (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("ubpl_k2_view_fixed: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  const long long BJ = (long long)B * J;
  if (BJ == 0) return UBPL_OK;
  GET_POWTAB(T);
  k2_view_fixed_kernel<<<blocks_for(BJ, 64), 64, 0, (cudaStream_t)stream>>>(preds, K, BJ, J, distThrMax, img_h, img_w,
                                                                             stride, sigma, S, out_mean, out_dist, out_legal,
                                                                             enable, gate_out, count_out, counts, T);
  return check_launch("ubpl_k2_view_fixed");
}

extern "C" int ubpl_coord_error(const double* pred, const float* gt, int gt_stride, int64_t n_sets, int B, int J,
                                int ref0, int ref1, double pck_thr, double* err, int32_t* acc, void* stream) {
  UBPL_REQUIRE(pred && gt && err && acc && gt_stride >= 2 && n_sets >= 0 && B >= 0 && J >= 1, "ubpl_coord_error: bad arguments");
  UBPL_REQUIRE(ref0 >= 0 && ref0 < J && ref1 >= 0 && ref1 < J, "ubpl_coord_error: pck_ref out of range");
  const long long n = n_sets * B * J;
  if (n == 0) return UBPL_OK;
  GET_POWTAB(T);
  coord_error_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(pred, gt, gt_stride, n_sets, B, J, ref0, ref1,
                                                                          pck_thr, err, acc, T);
  return check_launch("ubpl_coord_error");
}

extern "C" int ubpl_select_quantile_local(const double* dist, const double* legal, int64_t n, int J, int64_t k_rank,
                                          double reliableThr, double reliableDistMin, double* reliability,
                                          uint64_t* keys, uint8_t* enable, float* gate32, int32_t* counts,
                                          double* thr_out, double* ext_out, void* stream) {
  UBPL_REQUIRE(dist && legal && reliability && keys && counts && (enable || gate32) && n >= 1 && J >= 1,
               "ubpl_select_quantile_local: bad arguments");
  UBPL_REQUIRE(k_rank >= 0 && k_rank < n, "ubpl_select_quantile_local: rank out of range");
  select_quantile_local_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(dist, legal, n, J, k_rank, reliableThr, reliableDistMin,
                                                                      reliability, keys, enable, gate32, counts, thr_out, ext_out);
  return check_launch("ubpl_select_quantile_local");
}

extern "C" int ubpl_pair_distance(const double* c1, const double* c2, int64_t n, double* out, void* stream) {
  UBPL_REQUIRE(c1 && c2 && out && n >= 0, "ubpl_pair_distance: bad arguments");
  if (n == 0) return UBPL_OK;
  GET_POWTAB(T);
  pair_distance_kernel<<<blocks_for(n, 256), 256, 0, (cudaStream_t)stream>>>(c1, c2, n, out, T);
  return check_launch("ubpl_pair_distance");
}

extern "C" int ubpl_mix_dists(const float* gt, int gt_stride, int ref0, int ref1, double pck_thr, const float* p1,
                              const float* p2, const float* s1, const float* s2, const float* a1, const float* a2,
                              int B, int J, int A, double* err1, double* err2, int32_t* acc1, int32_t* acc2,
                              double* score1, double* score2, double* caug1, double* caug2, double* int1, double* int2,
                              double* ext, double* aext, void* stream) {
  UBPL_REQUIRE(B >= 0 && J >= 1 && A >= 1, "ubpl_mix_dists: bad dims");
  if (B == 0) return UBPL_OK;
  UBPL_REQUIRE(p1 && p2 && a1 && a2 && caug1 && caug2 && int1 && int2 && ext && aext, "ubpl_mix_dists: NULL pointer");
  UBPL_REQUIRE(!gt || (gt_stride >= 2 && err1 && err2 && acc1 && acc2 && ref0 >= 0 && ref0 < J && ref1 >= 0 && ref1 < J),
               "ubpl_mix_dists: gt needs err/acc outputs and valid pck_ref");
  UBPL_REQUIRE((!s1 || score1) && (!s2 || score2), "ubpl_mix_dists: scores need score outputs");
  const long long n = (long long)B * J;
  if (n == 0) return UBPL_OK;
  GET_POWTAB(T);
  mix_dists_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(gt, gt_stride, ref0, ref1, pck_thr, p1, p2, s1, s2, a1,
                                                                         a2, B, J, A, err1, err2, acc1, acc2, score1, score2,
                                                                         caug1, caug2, int1, int2, ext, aext, T);
  return check_launch("ubpl_mix_dists");
}

extern "C" int ubpl_mix_unc(const double* intDist, const double* extDist, const double* aExtDist, int64_t n, int J,
                            double distThrMax, double* hist, int32_t* hist_len, const double* score,
                            const double* score_thr, double* lma_out, double* mix_out, double* unc_out, uint8_t* enable,
                            float* gate32, int32_t* counts, void* stream) {
  UBPL_REQUIRE(intDist && extDist && aExtDist && hist && hist_len && n >= 0 && J >= 1, "ubpl_mix_unc: bad arguments");
  UBPL_REQUIRE(!score || score_thr, "ubpl_mix_unc: score needs score_thr");
  if (counts) {
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)(J + 1) * sizeof(int32_t), (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_mix_unc: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  }
  if (n == 0) return UBPL_OK;
  mix_unc_kernel<<<blocks_for(n, 128), 128, 0, (cudaStream_t)stream>>>(intDist, extDist, aExtDist, n, J, distThrMax, hist, hist_len,
                                                                       score, score_thr, lma_out, mix_out, unc_out, enable,
                                                                       gate32, counts);
  return check_launch("ubpl_mix_unc");
}

#include "nccl_select.inc"
#include "p2p_select.inc"
