// K1: back-warp + flip (+ optional left/right joint swap) + arg-max decode, fused; and the materialising
// warp (affine_back2).
//
// Reference semantics (file:line in /root/reference):
//   utils/augment.py:37-47      affine_back2  = F.affine_grid + F.grid_sample(bilinear, zeros,
//                               align_corners=True) + per-sample W mirror
//   utils/udaap/transforms.py:20-57   flip_back: mirror + exchange of the left/right joint channels (optional
//                               here: swap_perm[J]; NULL = the reference's live path, which has no swap)
//   utils/udaap/evaluation.py:13-30   get_preds (first arg-max, 1-based, zero where max <= 0)
//   utils/udaap/transforms.py:151-168 transform(invert=1) -> trunc + 1 (image space)
//   utils/process.py:362-373    quarter-offset refinement (kps_fromHeatmap2)
//
// Float op order of the warp is ATen's CPU order, recovered bit-exactly (oracle/ubpl_oracle.py):
//   base grid  lin(k) = k < n/2 ? fma(step,k,-1) : fma(-step,n-1-k,1),  step = 2/(n-1)
//   gx = fma(y, t01, x*t00) + t02;  ix = (gx+1)*((W-1)/2);  w = ix-floor(ix); e = 1-w; ...
//   out = fma(v_se, se, fma(v_sw, sw, fma(v_ne, ne, v_nw*nw)))
// All of it is written with explicit _rn intrinsics so nvcc cannot contract or reassociate.
//
// Decode strategy (one warp per heat-map, map staged in shared memory by a 1-D bulk async copy):
// a bilinear sample is a convex combination of its four corner texels (zero outside), so an
// output pixel can only reach the maximum if one of its corners is >= the maximum.  The warp
//   A) scans the staged map for its max / min / arg-max (conflict-free 128-bit LDS),
//   L) evaluates exactly the <= 30 output pixels around the pre-image of the arg-max texel;
//      their best value L is a lower bound of the warped maximum,
//   B) rescans for "candidate" texels v >= T = L - |L| * 2^-19 (the slack covers every
//      rounding in the interpolation) and takes their bounding box,
//   C) evaluates exactly every output pixel whose 2x2 footprint can touch that box and reduces
//      (value, canonical index) with torch.max's first-index tie rule.
// Pixels outside C have all four corners < T, hence a computed value < L: they cannot win or tie.
// Maps where this does not apply (max <= 0 after warp and not solvable geometrically, NaN/Inf, singular
// theta, huge candidate box: structure-less maps) are decoded exhaustively -- by ALL warps of the CTA
// together, straight out of the posting warp's staging buffer (CoopJob below), so the result is exact in
// every case and no map is left to a second launch.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace ubpl {

unsigned long long* work_counter(cudaStream_t stream);   // api.cu: a zeroed device counter for this launch

// K2 fused into the K1 epilogue (mean-teacher path, M = 1): every map that finishes bumps the arrival
// counter of its (sample, joint); the warp that brings it to K (all views decoded) computes the dispersion
// (utils/evaluation.py:44-54) and, in mode 2, the fixed-threshold rule, the visibility gate and the counts
// (business.py:237-261,375-376, process.py:262-268, losses.py:29) -- the arithmetic of
// view_dispersion_kernel / k2_view_fixed_kernel, without a second launch.
struct K2Fuse {
  int mode;                   // 0 off; one teacher: 1 dispersion only (mean, dist, legal), 2 + fixed rule, gate and
                              // counts; two teachers (assess_pseudo_unc2): 3 ensemble coord / extDist / legal, 4 + fixed rule
  int K;                      // maps per item (= V: K views of one teacher, or 2 x K/2 views of two teachers)
  int32_t* zero_div;          // modes 3/4: number of items whose two intDists are both 0 (business.py:135 divides by zero)
  int32_t* status;            // set to 1 when a hand-off word never arrived (bounded spin ran out): the results of
                              // this launch are void and the host raises (ops.warp_decode_k2 -> "status")
  unsigned* arrive;           // [B*J] arrival counters, zero before the launch
  unsigned long long* slots;  // [K][B*J] hand-off words ~pack(x, y); 0 = not written yet (zero before the launch)
  double distThrMax;
  double thr;                 // 1 - exp(-3*distThrMax/5), evaluated on the host with the libm CPython uses
  int img_h, img_w, S;
  float stride, sigma;
  float* mean;                // [B*J, 2]
  double* dist;               // [B*J]   (999 for items with an illegal view)
  uint8_t* legal;             // [B*J]
  uint8_t* enable;            // [B*J]   mode 2
  float* gate;                // [B*J]   mode 2
  int32_t* counts;            // [J+2]   mode 2: per-joint selected, total selected, S * #(open gates)
  PowTab T;
};

// K4 inside K1 (optional, ubpl_warp_decode_k2_ema): the mean-teacher EMA of ubpl_ema_multi_tensor, chunk by chunk,
// done by the warps that have run out of maps.  K1's launch ends with ~2 map times (~20 us on c2) in which ever fewer
// warps still decode and HBM is no longer saturated by the staged copies; the EMA depends on nothing in the chain, so
// its 12 bytes per parameter fill that tail instead of costing a launch of their own (a separate EMA kernel cannot run
// beside K1: a CTA that holds 225 KB of shared memory has the SM to itself, so "beside K1" was "in front of K1").
struct TailEma {
  const uint64_t* ema_ptrs;
  const uint64_t* param_ptrs;
  const long long* numels;
  const int32_t* chunk_tensor;
  const long long* chunk_start;
  long long n_chunks;            // 0: no EMA in this launch
  int chunk_elems;
  float a, oma;
  const float* alpha_dev;        // optional {alpha, 1 - alpha} in device memory (CUDA-graph replays follow the epoch)
  unsigned long long* next;      // piece counter (kEmaPiece elements each), zero before the launch
};

struct WDParams {
  FastDiv divJ, divB, divW;
  float stepx, stepy, sfx, sfy;
  const float* maps;
  long long sV, sB, sJ;
  int V, B, J, H, W;
  const float* theta;
  const uint8_t* flip;
  const int32_t* swap_perm;   // [J] or NULL: output joint j of a FLIPPED view is decoded from source channel swap_perm[j]
  const double* dec;
  int do_warp, refine, use_bulk;
  int32_t* out_idx;
  float* out_max;
  float* out_xy;
  float* out_hm_xy;
  unsigned long long* stats;
  unsigned long long* work;   // global claim counter (zeroed before the launch)
  const unsigned char* pf_ptr;   // optional: a global range (the student maps K3 reads next) that warps which
  unsigned long long pf_bytes;   // ran out of maps prefetch into L2, chunk by chunk, while the last maps finish
  unsigned long long* pf_next;   // its chunk counter (zeroed before the launch)
  int pf_every;                  // an idle warp prefetches one 32 KB chunk every pf_every polls (~0.25 us each)
  int inflight_cap;              // at most this many bulk copies in flight per CTA (0 = one per warp, no cap): with
                                 // thousands of concurrent 16 KB streams over a large footprint HBM falls into a
                                 // low-efficiency regime (profiles/README.md, round 2), so the copies are metered
  int dbg;                       // UBPL_K1_DBG (timing experiments only, results void): 1 skip phases L/B/C and the
                                 // exhaustive decode, 2 skip the epilogue, 4 skip only the exhaustive decode, 8 skip pass A;
                                 // 16 (results valid) leaves a %globaltimer timeline in stats[8..19], 32 adds per-warp
                                 // records from stats[32] on (stats must then hold 32 + 2*16*gridDim words), see tl_now
  K2Fuse k2;                  // optional K2 epilogue run by the warp that decodes the last view of a (sample, joint)
  TailEma ema;                // optional K4 done by the warps that have run out of maps
};

// FAST (template): the kernel instance for the call the fused chain makes -- warp on, maps staged by bulk copies, no
// quarter-offset refinement, no joint swap, no heat-map-space coordinates, no timing masks.  With these decided at
// compile time the decode loop is 5-8 % faster (c2 98.8 -> 93.5 us, c4 195 -> 184 us, c5 1713 -> 1574 us): the code is
// a third shorter and -- without the refinement in the epilogue -- the map's transform is dead after the decode.
// Every other call takes the generic instance, where the same fields are read at run time.
#define UBPL_F(expr, constant) (FAST ? (constant) : (expr))

// CH, CW (template): the map's height / width when the instance is compiled for one shape (64x64, 128x128), 0 = read
// from the parameters.  With the shape known the divisions by W are shifts, pass A is four (sixteen) straight
// batches, pass B's sweep over a residue class is one load: c2's K1 93.4 -> 85.5 us, c3's 49.0 -> 40.8 us, c4's
// 184 -> 176 us, and ptxas needs 100 registers instead of 124.
template <int CW>
__device__ __forceinline__ void divmod_w(const FastDiv& d, unsigned n, unsigned& q, unsigned& r) {
  if (CW) { q = n / (unsigned)CW; r = n - q * (unsigned)CW; } else d.divmod(n, q, r);
}

struct Xform {
  float t00, t01, t02, t10, t11, t12;
  float stepx, stepy, sfx, sfy;
  int H, W;
  bool flip;
};

struct ArgMax {
  float v;
  int i;
};

// Timeline probe of UBPL_K1_DBG=16 (tools/k1_ab.py prints it): stats[8]/[9] first / last CTA start, [10]/[11] first /
// last "first map landed" over the warps, [12]/[13] first / last "warp ran out of maps", [14] last warp exit,
// [15] ns summed over the warps waiting for staged copies, [16] ns waiting for a copy ticket, [17] ns in the
// exhaustive decode (own job or alone), [18] ns helping other warps' jobs, [19] ns between a warp's first landed map
// and its running out of maps.  The min slots are initialised to a large value by the host.
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ float lin_coord(int k, int n, float step) {
  if (n <= 1) return 0.f;  // ATen linspace_from_neg_one: a single step sits at 0
  return (k < (n >> 1)) ? __fmaf_rn(step, (float)k, -1.f) : __fmaf_rn(-step, (float)(n - 1 - k), 1.f);
}

// Bilinear sample at the normalised grid coordinates (gx, gy): ATen's op order, zero padding.
__device__ __forceinline__ float sample_at(const float* __restrict__ s, const Xform& X, float gx, float gy) {
  const float ix = __fmul_rn(__fadd_rn(gx, 1.f), X.sfx);
  const float iy = __fmul_rn(__fadd_rn(gy, 1.f), X.sfy);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float w = __fsub_rn(ix, x0f), e = __fsub_rn(1.f, w);
  const float n = __fsub_rn(iy, y0f), so = __fsub_rn(1.f, n);
  const float nw = __fmul_rn(so, e), ne = __fmul_rn(so, w), sw = __fmul_rn(n, e), se = __fmul_rn(n, w);
  const int x0 = (int)fminf(fmaxf(x0f, -2.f), (float)(X.W + 1));
  const int y0 = (int)fminf(fmaxf(y0f, -2.f), (float)(X.H + 1));
  const bool xa = (unsigned)x0 < (unsigned)X.W, xb = (unsigned)(x0 + 1) < (unsigned)X.W;
  const bool ya = (unsigned)y0 < (unsigned)X.H, yb = (unsigned)(y0 + 1) < (unsigned)X.H;
  const float* r0 = s + y0 * X.W + x0;
  const float v_nw = (xa & ya) ? r0[0] : 0.f;
  const float v_ne = (xb & ya) ? r0[1] : 0.f;
  const float v_sw = (xa & yb) ? r0[X.W] : 0.f;
  const float v_se = (xb & yb) ? r0[X.W + 1] : 0.f;
  float acc = __fmul_rn(v_nw, nw);
  acc = __fmaf_rn(v_ne, ne, acc);
  acc = __fmaf_rn(v_sw, sw, acc);
  acc = __fmaf_rn(v_se, se, acc);
  return acc;
}
// Bilinear sample at the normalised base-grid coordinates (xl, yl) of one output pixel from the source map
// s[H*W]: ATen's op order, zero padding.
__device__ __forceinline__ float eval_at(const float* __restrict__ s, const Xform& X, float xl, float yl) {
  const float gx = __fadd_rn(__fmaf_rn(yl, X.t01, __fmul_rn(xl, X.t00)), X.t02);
  const float gy = __fadd_rn(__fmaf_rn(yl, X.t11, __fmul_rn(xl, X.t10)), X.t12);
  return sample_at(s, X, gx, gy);
}
// The same with the column's products ax = xl * t00, ay = xl * t10 formed once per column (separately rounded
// products, so the bits are those of eval_at).
__device__ __forceinline__ float eval_col(const float* __restrict__ s, const Xform& X, float ax, float ay, float yl) {
  const float gx = __fadd_rn(__fmaf_rn(yl, X.t01, ax), X.t02);
  const float gy = __fadd_rn(__fmaf_rn(yl, X.t11, ay), X.t12);
  return sample_at(s, X, gx, gy);
}

// Exact bilinear sample of output pixel (row i, column jw in the WARPED frame, i.e. before the
// mirror) from the staged source map s[H*W].
__device__ __forceinline__ float eval_px(const float* __restrict__ s, const Xform& X, int i, int jw) {
  return eval_at(s, X, lin_coord(jw, X.W, X.stepx), lin_coord(i, X.H, X.stepy));
}

// The same with the base grid read from the CTA's tables (lx[W], ly[H] = lin_coord of every column / row).
__device__ __forceinline__ float eval_tab(const float* __restrict__ s, const Xform& X, const float* lx, const float* ly,
                                          int i, int jw) {
  return eval_at(s, X, lx[jw], ly[i]);
}

// The unnormalised source coordinates (ix, iy) of output pixel (row i, warped column jw), bit-identical
// to the ones eval_at interpolates at.
__device__ __forceinline__ void grid_px(const Xform& X, const float* lx, const float* ly, int i, int jw, float& ix,
                                        float& iy) {
  const float xl = lx[jw], yl = ly[i];
  const float gx = __fadd_rn(__fmaf_rn(yl, X.t01, __fmul_rn(xl, X.t00)), X.t02);
  const float gy = __fadd_rn(__fmaf_rn(yl, X.t11, __fmul_rn(xl, X.t10)), X.t12);
  ix = __fmul_rn(__fadd_rn(gx, 1.f), X.sfx);
  iy = __fmul_rn(__fadd_rn(gy, 1.f), X.sfy);
}

__device__ __forceinline__ void load_xform(Xform& X, const float* theta, const uint8_t* flip, long long vb, int H,
                                           int W) {
  const float* t = theta + vb * 6;
  X.t00 = t[0]; X.t01 = t[1]; X.t02 = t[2]; X.t10 = t[3]; X.t11 = t[4]; X.t12 = t[5];
  X.H = H; X.W = W;
  X.flip = flip ? (flip[vb] != 0) : false;
}
__device__ __forceinline__ void grid_consts(Xform& X, int H, int W) {
  X.stepx = (W > 1) ? __fdiv_rn(2.f, (float)(W - 1)) : 0.f;
  X.stepy = (H > 1) ? __fdiv_rn(2.f, (float)(H - 1)) : 0.f;
  X.sfx = (float)((double)(W - 1) / 2.0);
  X.sfy = (float)((double)(H - 1) / 2.0);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// Exhaustive decode of rows [row0, row1) of the warped map staged in SHARED memory at byte address `sa` (the
// tables lx / ly at lxa / lya).  Each lane walks whole columns (the column terms of the affine grid are hoisted),
// so the tie rule is carried by the index comparison.  FINITE = true: the map holds no NaN / Inf, the running
// arg-max is a branch-free select; false: torch.max's NaN rules (arg_better).
template <bool FINITE>
__device__ __forceinline__ ArgMax decode_rows(uint32_t sa, uint32_t lxa, uint32_t lya, float t00, float t01, float t02,
                                              float t10, float t11, float t12, float sfx, float sfy, int H, int W,
                                              int flip, int lane, int row0, int row1) {
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  const float fW1 = (float)(W + 1), fH1 = (float)(H + 1);
  for (int jo = lane; jo < W; jo += 32) {
    const int jw = flip ? (W - 1 - jo) : jo;
    const float xl = lds_f32(lxa + 4u * (uint32_t)jw);
    const float ax = __fmul_rn(xl, t00), ay = __fmul_rn(xl, t10);
#pragma unroll 4
    for (int i = row0; i < row1; ++i) {
      const float yl = lds_f32(lya + 4u * (uint32_t)i);
      const float gx = __fadd_rn(__fmaf_rn(yl, t01, ax), t02);
      const float gy = __fadd_rn(__fmaf_rn(yl, t11, ay), t12);
      const float ix = __fmul_rn(__fadd_rn(gx, 1.f), sfx);
      const float iy = __fmul_rn(__fadd_rn(gy, 1.f), sfy);
      const float x0f = floorf(ix), y0f = floorf(iy);
      const float w = __fsub_rn(ix, x0f), e = __fsub_rn(1.f, w);
      const float n = __fsub_rn(iy, y0f), so = __fsub_rn(1.f, n);
      const int x0 = (int)fminf(fmaxf(x0f, -2.f), fW1);
      const int y0 = (int)fminf(fmaxf(y0f, -2.f), fH1);
      const bool xa = (unsigned)x0 < (unsigned)W, xb = (unsigned)(x0 + 1) < (unsigned)W;
      const bool ya = (unsigned)y0 < (unsigned)H, yb = (unsigned)(y0 + 1) < (unsigned)H;
      const uint32_t r0 = sa + 4u * (uint32_t)(y0 * W + x0);
      const float v_nw = (xa & ya) ? lds_f32(r0) : 0.f;
      const float v_ne = (xb & ya) ? lds_f32(r0 + 4u) : 0.f;
      const float v_sw = (xa & yb) ? lds_f32(r0 + 4u * (uint32_t)W) : 0.f;
      const float v_se = (xb & yb) ? lds_f32(r0 + 4u * (uint32_t)W + 4u) : 0.f;
      float acc = __fmul_rn(v_nw, __fmul_rn(so, e));
      acc = __fmaf_rn(v_ne, __fmul_rn(so, w), acc);
      acc = __fmaf_rn(v_sw, __fmul_rn(n, e), acc);
      acc = __fmaf_rn(v_se, __fmul_rn(n, w), acc);
      const int k = i * W + jo;
      if (FINITE) {
        const bool better = (acc > bv) | ((acc == bv) & (k < bi));
        bv = better ? acc : bv;
        bi = better ? k : bi;
      } else if (arg_better(acc, k, bv, bi)) {
        bv = acc; bi = k;
      }
    }
  }
  if (FINITE) warp_argmax_finite(bv, bi); else warp_argmax(bv, bi);
  ArgMax r;
  r.v = bv; r.i = bi;
  return r;
}

__device__ __noinline__ ArgMax decode_exhaustive(uint32_t sa, uint32_t lxa, uint32_t lya, float t00, float t01, float t02,
                                                 float t10, float t11, float t12, float sfx, float sfy, int H, int W,
                                                 int flags, int lane, int row0, int row1) {
  // flags: bit 0 mirror, bit 1 the map may hold NaN / Inf
  if (flags & 2) return decode_rows<false>(sa, lxa, lya, t00, t01, t02, t10, t11, t12, sfx, sfy, H, W, flags & 1, lane, row0, row1);
  return decode_rows<true>(sa, lxa, lya, t00, t01, t02, t10, t11, t12, sfx, sfy, H, W, flags & 1, lane, row0, row1);
}

// ---------------------------------------------------------------------------------------------------
// Cooperative exhaustive decode inside the CTA.  A warp whose map needs all H*W output pixels posts it as a
// job: the map stays in the poster's staging buffer (shared memory, visible to the whole CTA) and is cut
// into kBands bands of rows.  Every warp of the CTA looks at the job word once per map (and keeps looking
// once it has run out of maps); whoever sees an open job claims bands until none is left.  The poster works
// on its own job too, waits for the last band, merges the kBands partial arg-maxes and carries on -- ~2 us
// instead of the ~25 us a single warp needs, so such a map no longer produces a tail and needs no second
// kernel launch.  One job per CTA at a time (a second poster helps the first job while it waits).
// ---------------------------------------------------------------------------------------------------
constexpr int kBands = 16;
constexpr int kChunk = 8;         // maps per chunk of the work distribution
constexpr int kChunkRing = 8;     // chunk descriptors kept per CTA
struct CoopJob {
  int owner;        // 0 = free, w + 1 = warp w holds the job slot
  int band_next;    // next band to claim; >= kBands: no open job
  int bands_done;   // bands whose partial result is written
  int warps_done;   // warps of the CTA that have run out of maps
  int local_next;   // the CTA's map tickets (see claim_map): ticket s -> chunk s / kChunk, offset s % kChunk
  int chunk_base[kChunkRing];    // first map index of the chunks the CTA holds, and for which chunk number (+1) each
  int chunk_ready[kChunkRing];   // ring slot is valid
  unsigned issued;  // copy tickets handed out / copies that have landed: a warp with ticket t issues its copy once
  unsigned landed;  // t - landed < WDParams::inflight_cap (FIFO, no retry races; both unused when the cap is 0)
  int flags;        // bit 0: mirrored view, bit 1: the map may hold NaN / Inf (NaN-aware compare)
  unsigned src;     // shared-memory byte address of the poster's staged map
  float t[6];
  float pv[kBands];
  int pi[kBands];
};

__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// Claims and decodes bands of the open job, if any (warp-wide call).
template <int CH, int CW>
__device__ __forceinline__ void coop_help(CoopJob* cj, const WDParams& p, const float* lx, const float* ly, int lane) {
  const int H = CH ? CH : p.H, W = CW ? CW : p.W;
  for (;;) {
    int b = kBands;
    if (lane == 0 && ld_volatile_s32(&cj->band_next) < kBands) b = atomicAdd(&cj->band_next, 1);
    b = __shfl_sync(0xffffffffu, b, 0);
    if (b >= kBands) return;
    __threadfence_block();                                   // the job's parameters were written before it opened
    const unsigned src = *reinterpret_cast<const volatile unsigned*>(&cj->src);
    const volatile float* t = cj->t;
    const int flags = ld_volatile_s32(&cj->flags);
    const int r0 = (H * b) / kBands, r1 = (H * (b + 1)) / kBands;
    const ArgMax r = decode_exhaustive(src, smem_u32(lx), smem_u32(ly), t[0], t[1], t[2], t[3], t[4], t[5], p.sfx, p.sfy,
                                       H, W, flags, lane, r0, r1);
    if (lane == 0) {
      *reinterpret_cast<volatile float*>(&cj->pv[b]) = r.v;
      *reinterpret_cast<volatile int*>(&cj->pi[b]) = r.i;
      __threadfence_block();
      atomicAdd(&cj->bands_done, 1);
    }
  }
}

// Posts the map staged at `s` as the CTA's job, works on it, and returns its exact arg-max (warp-wide call).
template <int CH, int CW>
__device__ __forceinline__ ArgMax coop_exhaustive(CoopJob* cj, const WDParams& p, const float* s, const Xform& X,
                                                  bool nan_aware, const float* lx, const float* ly, int warp, int lane) {
  {
    // Take the CTA's job slot.  If another warp holds it, structure-less maps are frequent here (several at once in
    // one CTA): cooperation then only serialises them, so this warp decodes its map on its own -- every warp busy
    // with a whole map is the best the SM can do when most maps need all their pixels.
    int got = 0;
    if (lane == 0) got = (atomicCAS(&cj->owner, 0, warp + 1) == 0) ? 1 : 0;
    got = __shfl_sync(0xffffffffu, got, 0);
    if (!got)
      return decode_exhaustive(smem_u32(s), smem_u32(lx), smem_u32(ly), X.t00, X.t01, X.t02, X.t10, X.t11, X.t12, p.sfx, p.sfy,
                               CH ? CH : p.H, CW ? CW : p.W, (X.flip ? 1 : 0) | (nan_aware ? 2 : 0), lane, 0, CH ? CH : p.H);
  }
  if (lane == 0) {
    *reinterpret_cast<volatile unsigned*>(&cj->src) = smem_u32(s);
    volatile float* t = cj->t;
    t[0] = X.t00; t[1] = X.t01; t[2] = X.t02; t[3] = X.t10; t[4] = X.t11; t[5] = X.t12;
    *reinterpret_cast<volatile int*>(&cj->flags) = (X.flip ? 1 : 0) | (nan_aware ? 2 : 0);
    *reinterpret_cast<volatile int*>(&cj->bands_done) = 0;
    __threadfence_block();
    atomicExch(&cj->band_next, 0);                           // opens the job
  }
  __syncwarp();
  coop_help<CH, CW>(cj, p, lx, ly, lane);
  if (lane == 0) {
    while (ld_volatile_s32(&cj->bands_done) < kBands) {
    }
    __threadfence_block();
  }
  __syncwarp();
  ArgMax r;
  r.v = -INFINITY; r.i = 0x7fffffff;
  if (lane < kBands) {
    r.v = *reinterpret_cast<const volatile float*>(&cj->pv[lane]);
    r.i = ld_volatile_s32(&cj->pi[lane]);
  }
  warp_argmax(r.v, r.i);
  __syncwarp();
  if (lane == 0) {
    __threadfence_block();
    atomicExch(&cj->owner, 0);
  }
  return r;
}

// One PIECE (kEmaPiece elements of one chunk) of the EMA (see TailEma) by one warp, the arithmetic of ema_multi_kernel
// (csrc/ema.cu).  A piece is one memory round trip: all of a lane's 8 + 8 128-bit loads are in flight together (a whole
// 8192-element chunk per warp would be 16 dependent round trips, ~20 us -- longer than the tail it is meant to fill).
// 8 + 8 loads fit since the specialised instances need ~100 registers (in the 128-register generic code they made the
// decode loop spill and cost K1 12 %; there the pieces were 512 elements).  The caller's table should have
// chunk_elems = kEmaPiece: one claim per work item, no empty claims on small tensors (c2 step 161.9 -> 157 us).
constexpr int kEmaPiece = 1024;
__device__ __forceinline__ void ema_piece(const TailEma& E, unsigned c, int pieces, float a, float oma, int lane) {
  const unsigned chunk = c / (unsigned)pieces;             // 32-bit: a 64-bit division is a call
  const int lo = (int)(c - chunk * (unsigned)pieces) * kEmaPiece;
  const int t = E.chunk_tensor[chunk];
  const long long start = E.chunk_start[chunk];
  const long long rem = E.numels[t] - start;
  const int n = (int)(rem < E.chunk_elems ? rem : E.chunk_elems);
  const int hi = min(n, lo + kEmaPiece);
  if (lo >= hi) return;
  float* e = reinterpret_cast<float*>(E.ema_ptrs[t]) + start + lo;
  const float* q = reinterpret_cast<const float*>(E.param_ptrs[t]) + start + lo;
  const int m = hi - lo;
  if ((((uintptr_t)e | (uintptr_t)q) & 15) == 0) {
    const int n4 = m >> 2;                                 // <= 256: at most 8 float4 per lane
    float4* e4 = reinterpret_cast<float4*>(e);
    const float4* q4 = reinterpret_cast<const float4*>(q);
    float4 ev[8], pv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = lane + 32 * u;
      if (i < n4) { ev[u] = e4[i]; pv[u] = ldg_stream(q4 + i); }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = lane + 32 * u;
      if (i < n4) {
        ev[u].x = ema1(ev[u].x, pv[u].x, a, oma); ev[u].y = ema1(ev[u].y, pv[u].y, a, oma);
        ev[u].z = ema1(ev[u].z, pv[u].z, a, oma); ev[u].w = ema1(ev[u].w, pv[u].w, a, oma);
        e4[i] = ev[u];
      }
    }
    for (int k = (n4 << 2) + lane; k < m; k += 32) e[k] = ema1(e[k], __ldg(q + k), a, oma);
  } else {
    for (int k = lane; k < m; k += 32) e[k] = ema1(e[k], __ldg(q + k), a, oma);
  }
}

// 3-input float min that PROPAGATES NaN (SASS FMNMX3.NAN): the running minimum turns NaN if any texel
// is NaN and -inf if any is -inf, which is how pass A detects non-finite maps for free.
__device__ __forceinline__ float min3_nan(float a, float b, float c) {
  float r;
  asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// Pass A over the staged map at float4 granularity: per-lane max (value, first float4 index) and a
// NaN-propagating running min.
__device__ __forceinline__ void scan_max(const float* s, int HW, int lane, float& bv, int& bq, float& mn) {
  bv = -INFINITY; bq = 0; mn = INFINITY;
  const int nq = HW >> 2;
  const float4* s4 = reinterpret_cast<const float4*>(s);
  int q = lane;
  for (; q + 224 < nq; q += 256) {
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = s4[q + 32 * u];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float m4 = fmaxf(max3(x[u].x, x[u].y, x[u].z), x[u].w);
      mn = min3_nan(x[u].z, x[u].w, min3_nan(x[u].x, x[u].y, mn));
      if (m4 > bv) { bv = m4; bq = q + 32 * u; }
    }
  }
  for (; q < nq; q += 32) {
    const float4 x = s4[q];
    const float m4 = fmaxf(max3(x.x, x.y, x.z), x.w);
    mn = min3_nan(x.z, x.w, min3_nan(x.x, x.y, mn));
    if (m4 > bv) { bv = m4; bq = q; }
  }
}

// Source channel of output joint j of (view, sample) vb: the joint itself, or its left/right partner when
// the view is flipped and a swap table is given.
template <bool FAST>
__device__ __forceinline__ unsigned src_joint(const WDParams& p, unsigned vb, unsigned j) {
  if (UBPL_F(p.swap_perm, (const int32_t*)nullptr) && p.flip && p.flip[vb]) return (unsigned)p.swap_perm[j];
  return j;
}

template <bool FAST>
__device__ __forceinline__ const float* map_src(const WDParams& p, long long n) {
  unsigned vb, j, v, b;
  p.divJ.divmod((unsigned)n, vb, j);
  p.divB.divmod(vb, v, b);
  j = src_joint<FAST>(p, vb, j);
  return p.maps + (long long)v * p.sV + (long long)b * p.sB + (long long)j * p.sJ;
}

// Dynamic work distribution in two levels.  Maps are handed out in chunks of kChunk consecutive indices: CTA b owns
// chunks b and gridDim + b from the start, every further chunk is drawn from the global counter p.work (chunk number
// 2 * gridDim + counter).  Inside the CTA the warps draw tickets from a shared-memory counter; the warp that draws the
// first ticket of chunk q fetches chunk q + 2 from the global counter, two chunks (>= one map per warp) ahead of its
// use.  One global atomic per kChunk maps instead of one per map: the single hot address was a measurable stall (the
// per-map claims of 2072 warps queued at one L2 atomic unit, ~8 % of the warps' time).  Chunk numbers grow along a
// CTA's ticket sequence, so the first ticket that lands beyond the last map ends the warp.  Lane 0 only.
__device__ __forceinline__ long long claim_map(const WDParams& p, CoopJob* cj) {
  const int s = atomicAdd(&cj->local_next, 1);
  const int q = s / kChunk, o = s - q * kChunk;
  if (o == 0) {
    const unsigned long long g = atomicAdd(p.work, 1ull);
    const unsigned long long base = (2ull * gridDim.x + g) * kChunk;
    *reinterpret_cast<volatile int*>(&cj->chunk_base[(q + 2) % kChunkRing]) = base > 0x7fffffffull ? 0x7fffffff : (int)base;
    __threadfence_block();
    *reinterpret_cast<volatile int*>(&cj->chunk_ready[(q + 2) % kChunkRing]) = q + 3;
  }
  while (ld_volatile_s32(&cj->chunk_ready[q % kChunkRing]) != q + 1) {
  }
  __threadfence_block();
  return (long long)ld_volatile_s32(&cj->chunk_base[q % kChunkRing]) + o;
}

// Called by lane 0.  With a cap, first takes a ticket and waits until fewer than `cap` of the CTA's copies are in
// flight (the waiter of a copy counts it as landed); the other lanes of the warp wait at the next warp-wide
// operation meanwhile.
template <bool FAST>
__device__ __forceinline__ void issue_map(const WDParams& p, long long n, float* dst, uint64_t* bar, uint64_t pol,
                                          uint32_t bytes, CoopJob* cj) {
  const float* src = map_src<FAST>(p, n);
  if (p.inflight_cap > 0) {
    const unsigned t = atomicAdd(&cj->issued, 1u);
    while ((int)(t - *reinterpret_cast<volatile unsigned*>(&cj->landed)) >= p.inflight_cap) __nanosleep(64);
  }
  mbar_arrive_expect_tx(bar, bytes);
  bulk_g2s(dst, src, bytes, bar, pol);
}

// K2 of one (sample, joint), run by the warp that decoded its last view (see K2Fuse).  Same float op order
// as view_dispersion_kernel / k2_view_fixed_kernel: float32 sequential sums for the mean, float64 python
// distances summed in view order.
__device__ __forceinline__ void k2_item(const WDParams& p, long long item, int j, int lane) {
  const K2Fuse& f = p.k2;
  const int K = f.K;
  const long long BJ = (long long)p.B * p.J;
  float x = 0.f, y = 0.f;
  if (lane < K) {
    // every view handed its coordinates over as ONE 64-bit word that is never zero once written, so the word
    // is its own flag: no fence on the writer's side, a (rarely taken) spin here
    unsigned long long* sp = f.slots + ((long long)lane * BJ + item);
    unsigned long long v;
    int spins = 0;
    do {                                                 // bounded: a lost word must not turn into a hung GPU
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(sp) : "memory");
    } while (v == 0ull && ++spins < (1 << 24));
    if (v == 0ull && f.status) atomicExch(f.status, 1);  // ... but it must not pass silently either
    v = ~v;
    x = __uint_as_float((unsigned)(v & 0xffffffffull));
    y = __uint_as_float((unsigned)(v >> 32));
    *sp = 0ull;                                          // ready for the next launch on this workspace
  }
  const bool legal = __all_sync(0xffffffffu, (lane >= K) || ((x >= 0.f) && (y >= 0.f)));
  float sx = __shfl_sync(0xffffffffu, x, 0), sy = __shfl_sync(0xffffffffu, y, 0);
  for (int k = 1; k < K; ++k) {
    sx = __fadd_rn(sx, __shfl_sync(0xffffffffu, x, k));
    sy = __fadd_rn(sy, __shfl_sync(0xffffffffu, y, k));
  }
  const float mx = __fdiv_rn(sx, (float)K), my = __fdiv_rn(sy, (float)K);        // torch.mean (float32)
  double dk = 0.0;
  if (lane < K) dk = py_dist((double)x, (double)y, (double)mx, (double)my, f.T);
  double acc = 0.0;
  for (int k = 0; k < K; ++k) acc = __dadd_rn(acc, __shfl_sync(0xffffffffu, dk, k));   // sum(dists), view order
  if (lane != 0) return;
  const double dist = legal ? __ddiv_rn(acc, (double)K) : 999.0;                  // business.py:123 sentinel
  if (f.mean) { f.mean[2 * item] = mx; f.mean[2 * item + 1] = my; }
  if (f.dist) f.dist[item] = dist;
  if (f.legal) f.legal[item] = legal ? 1 : 0;
  if (f.mode == 2) {
    const double thr = f.thr;
    const double unc = __dsub_rn(1.0, exp(-__ddiv_rn(dist, 5.0)));
    const bool en = legal && (unc <= thr);
    const Gauss g = gauss_setup(mx, my, f.img_h, f.img_w, f.stride, f.sigma);
    const float gt = (en ? 1.f : 0.f) * g.vis;
    if (f.enable) f.enable[item] = en ? 1 : 0;
    f.gate[item] = gt;
    if (en) { atomicAdd(f.counts + j, 1); atomicAdd(f.counts + p.J, 1); }
    if (gt > 0.f) atomicAdd(f.counts + p.J + 1, f.S);
  }
}

// K2 of one (sample, joint) with TWO teachers (utils/business.py:109-161 in the array form of assess_dual_kernel,
// same float op order): lanes [0, Kt) hold teacher 1's views, lanes [Kt, 2Kt) teacher 2's.  p_m = float32 view
// mean of teacher m; when every coordinate is legal: intDist_m = mean pairwise distance of teacher m's views in
// itertools.combinations order, w_m = intDist_m / (intDist_1 + intDist_2), coord = w1*p1 + w2*p2 (float64),
// extDist = mean_k dist(view_k of teacher 1, view_k of teacher 2); otherwise the 999 sentinels and the plain
// float32 mean of p1 and p2.  Mode 4 adds the fixed rule on extDist, the visibility gate and the counts.
__device__ __forceinline__ void k2_item_dual(const WDParams& p, long long item, int j, int lane) {
  const K2Fuse& f = p.k2;
  const int Kt = f.K >> 1;
  const long long BJ = (long long)p.B * p.J;
  float x = 0.f, y = 0.f;
  if (lane < f.K) {
    unsigned long long* sp = f.slots + ((long long)lane * BJ + item);
    unsigned long long v;
    int spins = 0;
    do {
      asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(sp) : "memory");
    } while (v == 0ull && ++spins < (1 << 24));
    if (v == 0ull && f.status) atomicExch(f.status, 1);  // ... but it must not pass silently either
    v = ~v;
    x = __uint_as_float((unsigned)(v & 0xffffffffull));
    y = __uint_as_float((unsigned)(v >> 32));
    *sp = 0ull;
  }
  const unsigned legal_mask = __ballot_sync(0xffffffffu, (lane < f.K) && (x >= 0.f) && (y >= 0.f));
  const unsigned m1 = (Kt >= 32) ? 0xffffffffu : ((1u << Kt) - 1u);
  const bool g1 = (legal_mask & m1) == m1, g2 = ((legal_mask >> Kt) & m1) == m1;
  float s1x = __shfl_sync(0xffffffffu, x, 0), s1y = __shfl_sync(0xffffffffu, y, 0);
  float s2x = __shfl_sync(0xffffffffu, x, Kt), s2y = __shfl_sync(0xffffffffu, y, Kt);
  for (int k = 1; k < Kt; ++k) {
    s1x = __fadd_rn(s1x, __shfl_sync(0xffffffffu, x, k));
    s1y = __fadd_rn(s1y, __shfl_sync(0xffffffffu, y, k));
    s2x = __fadd_rn(s2x, __shfl_sync(0xffffffffu, x, Kt + k));
    s2y = __fadd_rn(s2y, __shfl_sync(0xffffffffu, y, Kt + k));
  }
  const float p1x = __fdiv_rn(s1x, (float)Kt), p1y = __fdiv_rn(s1y, (float)Kt);
  const float p2x = __fdiv_rn(s2x, (float)Kt), p2y = __fdiv_rn(s2y, (float)Kt);
  const bool ori_legal = (p1x >= 0.f && p1y >= 0.f) && (p2x >= 0.f && p2y >= 0.f);
  const bool full = ori_legal && g1 && g2;                                    // warp-uniform
  double sd[2] = {0.0, 0.0}, se = 0.0;
  const int P = Kt * (Kt - 1) / 2;
  if (full) {
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      for (int base = 0; base < P; base += 32) {
        // pair number base + lane in combinations order -> (u, v)
        int q = base + lane, u = 0;
        while (u < Kt - 1 && q >= Kt - 1 - u) { q -= Kt - 1 - u; ++u; }
        const int vv = u + 1 + q;
        const bool have = (base + lane) < P;
        const int su = m * Kt + (have ? u : 0), sv = m * Kt + (have ? vv : 0);
        const float xu = __shfl_sync(0xffffffffu, x, su), yu = __shfl_sync(0xffffffffu, y, su);
        const float xv = __shfl_sync(0xffffffffu, x, sv), yv = __shfl_sync(0xffffffffu, y, sv);
        double d = 0.0;
        if (have) d = py_dist((double)xu, (double)yu, (double)xv, (double)yv, f.T);
        const int cnt = min(32, P - base);
        for (int t = 0; t < cnt; ++t) sd[m] = __dadd_rn(sd[m], __shfl_sync(0xffffffffu, d, t));
      }
    }
    // extDist: view k of teacher 1 against view k of teacher 2
    const float ox = __shfl_sync(0xffffffffu, x, (lane < Kt) ? lane + Kt : lane);
    const float oy = __shfl_sync(0xffffffffu, y, (lane < Kt) ? lane + Kt : lane);
    double de = 0.0;
    if (lane < Kt) de = py_dist((double)x, (double)y, (double)ox, (double)oy, f.T);
    for (int k = 0; k < Kt; ++k) se = __dadd_rn(se, __shfl_sync(0xffffffffu, de, k));
  }
  if (lane != 0) return;
  double legal = ori_legal ? 1.0 : 0.0, ext = 999.0;
  double cx = (double)__fdiv_rn(__fadd_rn(p1x, p2x), 2.f), cy = (double)__fdiv_rn(__fadd_rn(p1y, p2y), 2.f);
  if (full) {
    const double d1 = __ddiv_rn(sd[0], (double)P), d2 = __ddiv_rn(sd[1], (double)P);   // K < 2: 0/0 -> NaN
    const double den = __dadd_rn(d1, d2);
    double w1 = 0.5, w2 = 0.5;
    if (den == 0.0) {
      if (f.zero_div) atomicAdd(f.zero_div, 1);
    } else {
      w1 = __ddiv_rn(d1, den);
      w2 = __ddiv_rn(d2, den);
    }
    cx = __dadd_rn(__dmul_rn(w1, (double)p1x), __dmul_rn(w2, (double)p2x));
    cy = __dadd_rn(__dmul_rn(w1, (double)p1y), __dmul_rn(w2, (double)p2y));
    legal = 1.0;
    ext = __ddiv_rn(se, (double)Kt);
  }
  const float c32x = (float)cx, c32y = (float)cy;
  if (f.mean) { f.mean[2 * item] = c32x; f.mean[2 * item + 1] = c32y; }
  if (f.dist) f.dist[item] = ext;
  if (f.legal) f.legal[item] = legal > 0.0 ? 1 : 0;
  if (f.mode == 4) {
    const double thr = f.thr;
    const double unc = __dsub_rn(1.0, exp(-__ddiv_rn(ext, 5.0)));
    const bool en = (legal > 0.0) && (unc <= thr);
    const Gauss g = gauss_setup(c32x, c32y, f.img_h, f.img_w, f.stride, f.sigma);
    const float gt = (en ? 1.f : 0.f) * g.vis;
    if (f.enable) f.enable[item] = en ? 1 : 0;
    f.gate[item] = gt;
    if (en) { atomicAdd(f.counts + j, 1); atomicAdd(f.counts + p.J, 1); }
    if (gt > 0.f) atomicAdd(f.counts + p.J + 1, f.S);
  }
}

// Epilogue of one map (warp-wide call): arg-max -> heat-map coordinates (mask, optional quarter-offset
// refinement) -> image-space coordinates -> outputs -> (optional) arrival at the item's K2.
template <bool FAST, int CH, int CW>
__device__ __forceinline__ void finish_map(const WDParams& p, long long n, int v, int b, int j, const float* s, const Xform& X,
                                           const float* lx, const float* ly, float rv, int ri, double dc0, double dc1,
                                           double dc2, double dc3, int lane, long long& pend_item, unsigned& pend_old) {
  const int H = CH ? CH : p.H, W = CW ? CW : p.W;
  unsigned ayu, axu;
  divmod_w<CW>(p.divW, (unsigned)ri & 0x7fffffffu, ayu, axu);
  const int ax = (int)axu, ay = (int)ayu;             // 0-based arg-max, canonical frame
  float hx = 0.f, hy = 0.f;
  const bool keep = rv > 0.f;                          // maxval.gt(0): NaN -> masked
  if (keep) { hx = (float)(ax + 1); hy = (float)(ay + 1); }
  const int refine = UBPL_F(p.refine, 0);
  const bool do_ref = (refine == 2) || (refine == 1 && j < 2);
  if (do_ref) {
    // process.py:366-371: 1 < px < res[0] and 1 < py < res[1] on the 1-based coordinates
    if (keep && ax >= 1 && ax <= W - 2 && ay >= 1 && ay <= H - 2) {
      float nb = 0.f;
      if (lane < 4) {
        const int di = (lane == 2) ? 1 : (lane == 3 ? -1 : 0);
        const int dj = (lane == 0) ? 1 : (lane == 1 ? -1 : 0);
        const int i = ay + di, jo = ax + dj;
        nb = UBPL_F(p.do_warp, 1) ? eval_tab(s, X, lx, ly, i, X.flip ? (W - 1 - jo) : jo) : s[i * W + jo];
      }
      const float xp = __shfl_sync(0xffffffffu, nb, 0), xm = __shfl_sync(0xffffffffu, nb, 1);
      const float yp = __shfl_sync(0xffffffffu, nb, 2), ym = __shfl_sync(0xffffffffu, nb, 3);
      const float dx = xp - xm, dy = yp - ym;
      hx += (dx > 0.f) ? 0.25f : ((dx < 0.f) ? -0.25f : 0.f);
      hy += (dy > 0.f) ? 0.25f : ((dy < 0.f) ? -0.25f : 0.f);
    }
  }
  if (refine != 0) { hx += 0.5f; hy += 0.5f; }        // process.py:372 (+0.5 for every joint)
  float ox = hx, oy = hy;
  if (lane == 0) {
    if (p.out_idx) p.out_idx[n] = ri;
    if (p.out_max) p.out_max[n] = rv;
    if (UBPL_F(p.out_hm_xy, (float*)nullptr)) { p.out_hm_xy[2 * n] = hx; p.out_hm_xy[2 * n + 1] = hy; }
    if (p.out_xy) {
      if (p.dec) {
        // np.dot row: (a00*(x-1) + 0*(y-1)) + a02, astype(int) truncation, +1
        const double tx = __dadd_rn(__dmul_rn(dc0, (double)hx - 1.0), dc1);
        const double ty = __dadd_rn(__dmul_rn(dc2, (double)hy - 1.0), dc3);
        ox = (float)(trunc(tx) + 1.0);
        oy = (float)(trunc(ty) + 1.0);
      }
      p.out_xy[2 * n] = ox; p.out_xy[2 * n + 1] = oy;
    }
  }
  if (p.k2.mode) {
    // Hand this view's coordinates to its (sample, joint) and count the view in.  No fence: the coordinates
    // travel as one self-flagging 64-bit word (see k2_item), the counter is a relaxed atomic, and its result is
    // not consumed here -- a release fence plus a blocking atomic per map cost ~13 us per launch on the warps'
    // critical path.  The caller resolves the ticket one map later (k2_resolve).
    pend_item = (long long)b * p.J + j;
    if (lane == 0) {
      const long long BJ = (long long)p.B * p.J;
      const unsigned long long word = ~((unsigned long long)__float_as_uint(ox) | ((unsigned long long)__float_as_uint(oy) << 32));
      asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p.k2.slots + ((long long)v * BJ + pend_item)), "l"(word) : "memory");
      asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], 1;" : "=r"(pend_old) : "l"(p.k2.arrive + pend_item) : "memory");
    }
  }
}

// Second half of the K2 hand-off: the warp that saw K-1 earlier arrivals owns the item and computes its K2.
__device__ __forceinline__ void k2_resolve(const WDParams& p, long long& pend_item, unsigned pend_old, int lane) {
  if (pend_item < 0) return;
  const unsigned old = __shfl_sync(0xffffffffu, pend_old, 0);
  if (old == (unsigned)(p.k2.K - 1)) {
    __syncwarp();
    if (lane == 0) p.k2.arrive[pend_item] = 0u;          // ready for the next launch on this workspace
    unsigned ib, ij;
    p.divJ.divmod((unsigned)pend_item, ib, ij);
    if (p.k2.mode >= 3) k2_item_dual(p, pend_item, (int)ij, lane);
    else k2_item(p, pend_item, (int)ij, lane);
  }
  pend_item = -1;
}

// What pass A leaves for the later phases.
struct PassA {
  float bv; int bi;                  // warp-uniform source maximum and its first flat index
  float lane_max;                    // this lane's best float4 maximum
  float c0, f0, C00, C01, C10, C11;  // approximate pixel-space affine offset and inverse (boxes only)
};

// Phases L, B and C (see the header comment) on the staged map.  On return: `exhaustive` asks for the
// exhaustive decode, otherwise (rv, ri) is the exact result.
template <int CH, int CW>
__device__ __forceinline__ void decode_pruned(const WDParams& p, const float* __restrict__ s, const Xform& X,
                                              const float* lx, const float* ly, const PassA& A, int lane, float& rv,
                                              int& ri, bool& exhaustive, unsigned long long& n_eval) {
  const int H = CH ? CH : p.H, W = CW ? CW : p.W, HW = H * W;
  const float c0 = A.c0, f0 = A.f0, C00 = A.C00, C01 = A.C01, C10 = A.C10, C11 = A.C11;
  const float bv = A.bv;
  const int bi = A.bi;
  float L = -INFINITY; int Li = 0x7fffffff;
  const float kSlack = 1.9073486328125e-06f;   // 2^-19
  float T = 0.f;
  bool prune = false, solved = false;
  {
    // ---- phase L: lower bound from the pixels around the pre-image of the arg-max texel
    unsigned biy, bix;
    divmod_w<CW>(p.divW, (unsigned)bi, biy, bix);
    const float sx = (float)bix - c0, sy = (float)biy - f0;
    const float oj = C00 * sx + C01 * sy, oi = C10 * sx + C11 * sy;
    if (oj > -4.f && oj < (float)W + 4.f && oi > -4.f && oi < (float)H + 4.f && lane < 30) {
      const int r = (lane * 43) >> 8;                        // lane / 6 for lane < 30
      const int jw = (int)floorf(oj) - 2 + (lane - 6 * r);
      const int i = (int)floorf(oi) - 2 + r;
      if (jw >= 0 && jw < W && i >= 0 && i < H) {
        L = eval_tab(s, X, lx, ly, i, jw);
        Li = i * W + (X.flip ? (W - 1 - jw) : jw);
      }
    }
    n_eval += 30;
    warp_argmax_finite(L, Li);
    // Candidate threshold.  For a pixel whose four (zero-extended) corners are all < T the computed sample
    // is < L: the fma chain rounds by at most 4 ulp of sum(w|v|), and negative corners lower the exact
    // value by more than the rounding they add, so 2^-19 relative slack covers it.
    T = L - fabsf(L) * kSlack;
    prune = (L > 0.f) && (T > 0.f);          // zero padding cannot be a candidate when T > 0
  }
  if (!prune) {
    // The warped maximum is not known to be positive (an all-negative map, or a map whose maximum is
    // exactly 0).  Along a row the computed ix and iy are monotone in the column (every rounding step
    // is monotone), so the row ends classify the whole frame:
    //   inside : every pixel samples with all four corners in bounds -> the convex bound holds for any
    //            sign (no zero padding involved) and the pruned search stays valid;
    //   Z      : pixels with ix <= -1 | ix >= W | iy <= -1 | iy >= H read nothing but padding (value
    //            exactly 0); every other pixel of an all-negative map is < 0, so the maximum is 0 at the
    //            first Z pixel in canonical order.
    bool inside = true;
    int zrow = 0x7fffffff;
    for (int i = lane; i < H; i += 32) {
#pragma unroll
      for (int endc = 0; endc < 2; ++endc) {
        float ix, iy;
        grid_px(X, lx, ly, i, endc ? W - 1 : 0, ix, iy);
        inside = inside && (ix >= 0.f) && (ix <= (float)(W - 1)) && (iy >= 0.f) && (iy <= (float)(H - 1));
        const bool z = (ix <= -1.f) || (ix >= (float)W) || (iy <= -1.f) || (iy >= (float)H);
        if (z) zrow = min(zrow, i);
      }
    }
    inside = __all_sync(0xffffffffu, inside);
    zrow = __reduce_min_sync(0xffffffffu, zrow);
    if (inside && L > -INFINITY) {
      prune = true;
    } else if (bv < -1e-20f && zrow < H) {
      int zcol = 0x7fffffff;
      for (int jo = lane; jo < W; jo += 32) {
        float ix, iy;
        grid_px(X, lx, ly, zrow, X.flip ? (W - 1 - jo) : jo, ix, iy);
        if ((ix <= -1.f) || (ix >= (float)W) || (iy <= -1.f) || (iy >= (float)H)) zcol = min(zcol, jo);
      }
      zcol = __reduce_min_sync(0xffffffffu, zcol);
      rv = eval_tab(s, X, lx, ly, zrow, X.flip ? (W - 1 - zcol) : zcol);   // +-0, exactly what the warp produces there
      ri = zrow * W + zcol;
      solved = true;
      n_eval += 2 * H + W;
    } else {
      exhaustive = true;
    }
  }
  if (prune && !solved && !exhaustive) {
    // ---- pass B: bounding box of the candidate texels (v >= T) -----------------------
    int txmin = W, txmax = -1, tymin = H, tymax = -1;
    const int nq = HW >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(s);
    auto hit = [&](int k) {
      unsigned ty, tx;
      divmod_w<CW>(p.divW, (unsigned)k, ty, tx);
      txmin = min(txmin, (int)tx); txmax = max(txmax, (int)tx); tymin = min(tymin, (int)ty); tymax = max(tymax, (int)ty);
    };
    auto visit = [&](const float4& x, int q) {
      if (fmaxf(max3(x.x, x.y, x.z), x.w) >= T) {
        if (x.x >= T) hit((q << 2) + 0);
        if (x.y >= T) hit((q << 2) + 1);
        if (x.z >= T) hit((q << 2) + 2);
        if (x.w >= T) hit((q << 2) + 3);
      }
    };
    // lane l scanned the float4s q = l (mod 32) in pass A and knows their maximum (lane_max):
    // only the residue classes whose maximum reaches T can hold candidates.  All 32 lanes
    // re-read one such class together (32 float4 per step).
    unsigned hot = __ballot_sync(0xffffffffu, A.lane_max >= T);
    if (__popc(hot) <= 12) {
      while (hot) {
        const int h = __ffs(hot) - 1;
        hot &= hot - 1;
        for (int q = h + 32 * lane; q < nq; q += 1024) visit(s4[q], q);
      }
    } else {
#pragma unroll 4
      for (int q = lane; q < nq; q += 32) visit(s4[q], q);
    }
    for (int k = (nq << 2) + lane; k < HW; k += 32)
      if (s[k] >= T) hit(k);
    txmin = __reduce_min_sync(0xffffffffu, txmin); txmax = __reduce_max_sync(0xffffffffu, txmax);
    tymin = __reduce_min_sync(0xffffffffu, tymin); tymax = __reduce_max_sync(0xffffffffu, tymax);
    // pre-image of [txmin-1, txmax+1] x [tymin-1, tymax+1] -> bounding box in the warped frame
    const float X0 = (float)(txmin - 1) - c0, X1 = (float)(txmax + 1) - c0;
    const float Y0 = (float)(tymin - 1) - f0, Y1 = (float)(tymax + 1) - f0;
    const float ja = C00 * X0, jb = C00 * X1, jc = C01 * Y0, jd = C01 * Y1;
    const float ia = C10 * X0, ib = C10 * X1, ic = C11 * Y0, id = C11 * Y1;
    const float jlo = fminf(ja, jb) + fminf(jc, jd), jhi = fmaxf(ja, jb) + fmaxf(jc, jd);
    const float ilo = fminf(ia, ib) + fminf(ic, id), ihi = fmaxf(ia, ib) + fmaxf(ic, id);
    const float m = 0.03f;
    const int jmin = max(0, (int)ceilf(fmaxf(jlo - m, -1.f))), jmax = min(W - 1, (int)floorf(fminf(jhi + m, (float)W)));
    const int imin = max(0, (int)ceilf(fmaxf(ilo - m, -1.f))), imax = min(H - 1, (int)floorf(fminf(ihi + m, (float)H)));
    const int bw = jmax - jmin + 1, bh = imax - imin + 1;
    const int area = (bw > 0 && bh > 0) ? bw * bh : 0;
    if (area > 768 || area * 4 > HW) {
      exhaustive = true;
    } else {
      // ---- phase C: exact evaluation of every pixel that can touch a candidate ---------
      rv = L; ri = Li;
      int ci = lane / max(bw, 1), cj = lane - ci * bw;          // (row, col) of this lane's first pixel in the box
      const int di = 32 / max(bw, 1), dj = 32 - di * bw;
      for (int t = lane; t < area; t += 32, ci += di, cj += dj) {
        if (cj >= bw) { cj -= bw; ++ci; }
        const int i = imin + ci, jw = jmin + cj;
        const float v = eval_tab(s, X, lx, ly, i, jw);
        const int k = i * W + (X.flip ? (W - 1 - jw) : jw);
        if ((v > rv) || (v == rv && k < ri)) { rv = v; ri = k; }
      }
      n_eval += area;
      warp_argmax_finite(rv, ri);
    }
  }
}

// One warp per heat-map, maps claimed from a global counter, each staged in the warp's shared-memory buffer by
// a 1-D bulk async copy (TMA engine); the next copy starts when the map is finished.
// TL: the %globaltimer probe of UBPL_K1_DBG=16 is compiled in (its accumulators cost registers).
// EMA: the idle phase can do the mean-teacher EMA (TailEma).  A separate instance because the mere presence of that
// code costs the decode loop 2-4 % (measured A/B on c2 / c3 / c5 with the EMA never executed: 98.2 -> 100.3 us,
// 52.0 -> 53.7 us, 1714 -> 1779 us; same registers, no spills -- code placement): launches without an EMA keep the
// instance that does not have it.
template <bool TL, bool EMA, bool FAST, int CH, int CW>
__global__ void __launch_bounds__(512, 1) warp_decode_kernel(const WDParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int do_warp = UBPL_F(p.do_warp, 1), use_bulk = UBPL_F(p.use_bulk, 1), dbg = UBPL_F(p.dbg, 0);
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = CH ? CH : p.H, W = CW ? CW : p.W, HW = H * W;
  const uint32_t map_bytes = (uint32_t)HW * 4u;
  const uint32_t buf_stride = (map_bytes + 127u) & ~127u;                                        // 128 B aligned
  float* buf0 = reinterpret_cast<float*>(smem_raw + (size_t)warp * buf_stride);
  unsigned char* tail = smem_raw + (size_t)warps * buf_stride;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail) + warp;                                      // 16 x 8 bytes
  CoopJob* cj = reinterpret_cast<CoopJob*>(tail + 128);                                          // 256 bytes
  float* lx = reinterpret_cast<float*>(tail + 384);                                              // [W] base grid, columns
  float* ly = lx + W;                                                                            // [H] base grid, rows

  const long long N = (long long)p.V * p.B * p.J;
  uint64_t pol = 0;
  const bool tl = TL && p.stats;
  unsigned long long tl_first = 0, tl_wait = 0, tl_ticket = 0, tl_exh = 0, tl_help = 0;
  if (tl && threadIdx.x == 0) { const unsigned long long t = tl_now(); atomicMin(p.stats + 8, t); atomicMax(p.stats + 9, t); }
  if (threadIdx.x == 0) {
    cj->owner = 0; cj->band_next = kBands; cj->bands_done = 0; cj->warps_done = 0; cj->issued = 0u; cj->landed = 0u;
    cj->local_next = 0;
    for (int k = 0; k < kChunkRing; ++k) { cj->chunk_base[k] = 0; cj->chunk_ready[k] = 0; }
    const unsigned long long c0 = (unsigned long long)blockIdx.x * kChunk, c1 = ((unsigned long long)gridDim.x + blockIdx.x) * kChunk;
    cj->chunk_base[0] = c0 > 0x7fffffffull ? 0x7fffffff : (int)c0; cj->chunk_ready[0] = 1;
    cj->chunk_base[1] = c1 > 0x7fffffffull ? 0x7fffffff : (int)c1; cj->chunk_ready[1] = 2;
  }
  for (int k = threadIdx.x; k < W; k += blockDim.x) lx[k] = lin_coord(k, W, p.stepx);
  for (int k = threadIdx.x; k < H; k += blockDim.x) ly[k] = lin_coord(k, H, p.stepy);
  // The per-map transform (theta, flip) of the NEXT map is loaded as soon as that map is known -- before the epilogue
  // of the current one -- so that the ~0.3 us of L2 latency is off the path between two maps.
  struct Next {
    long long n;
    float t00, t01, t02, t10, t11, t12;
    bool flip;
  } nx;
  nx.n = N; nx.t00 = nx.t01 = nx.t02 = nx.t10 = nx.t11 = nx.t12 = 0.f; nx.flip = false;
  auto fetch_next = [&]() {                              // draws the warp's next map and starts its transform loads
    long long nn = 0;
    if (lane == 0) nn = claim_map(p, cj);
    nx.n = (long long)__shfl_sync(0xffffffffu, (unsigned long long)nn, 0);
    if (nx.n < N && do_warp) {
      const unsigned vb = p.divJ.div((unsigned)nx.n);
      // volatile asm: the loads are ISSUED here (a plain load may be sunk to its first use, after the epilogue)
      const float* t = p.theta + (long long)vb * 6;
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t00) : "l"(t));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t01) : "l"(t + 1));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t02) : "l"(t + 2));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t10) : "l"(t + 3));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t11) : "l"(t + 4));
      asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(nx.t12) : "l"(t + 5));
      unsigned fl = 0;
      if (p.flip) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(fl) : "l"(p.flip + vb));
      nx.flip = fl != 0;
    }
  };
  if (use_bulk && lane == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
    pol = l2_evict_first_policy();
  }
  __syncthreads();
  fetch_next();
  if (use_bulk && lane == 0 && nx.n < N) {
    const unsigned long long t0 = tl ? tl_now() : 0ull;
    issue_map<FAST>(p, nx.n, buf0, bar, pol, map_bytes, cj);
    if (tl) tl_ticket += tl_now() - t0;
  }

  unsigned long long n_slow = 0, n_eval = 0, n_maps = 0;
  long long pend_item = -1;                              // K2 ticket of the previous map (see finish_map)
  unsigned pend_old = 0;
  long long it = 0;                                      // staged copies waited for so far (mbarrier parity)
  for (;; ++it) {
    const long long n = nx.n;
    if (n >= N) break;
    const float* s = buf0;
    // an exhaustive job posted by another warp of this CTA: help before this map (whose copy is in flight).  The
    // decision is taken by lane 0 and broadcast: coop_help is a warp-wide call, every lane must take the same branch
    // (the lanes are not necessarily converged here -- lane 0 may come late out of issue_map's wait).
    if (do_warp) {
      int open = 0;
      if (lane == 0) open = ld_volatile_s32(&cj->band_next) < kBands ? 1 : 0;
      if (__shfl_sync(0xffffffffu, open, 0)) {
        const unsigned long long t0 = tl ? tl_now() : 0ull;
        coop_help<CH, CW>(cj, p, lx, ly, lane);
        if (tl) tl_help += tl_now() - t0;
      }
    }
    unsigned vbu, ju, vu, bu;
    p.divJ.divmod((unsigned)n, vbu, ju);
    p.divB.divmod(vbu, vu, bu);
    const int j = (int)ju, b = (int)bu;
    // ---- per-map transform set-up, issued BEFORE waiting for the staged map so that the global loads of
    // theta / flip / dec and the inverse-affine arithmetic overlap the copy latency
    Xform X;
    X.H = H; X.W = W; X.flip = false;
    X.t00 = X.t01 = X.t02 = X.t10 = X.t11 = X.t12 = 0.f;
    X.stepx = p.stepx; X.stepy = p.stepy; X.sfx = p.sfx; X.sfy = p.sfy;
    PassA A;
    A.c0 = A.f0 = A.C00 = A.C01 = A.C10 = A.C11 = 0.f;
    bool bad_xform = false;
    double dc0 = 0.0, dc1 = 0.0, dc2 = 0.0, dc3 = 0.0;
    if (p.dec) { const double* c = p.dec + (size_t)b * 4; dc0 = c[0]; dc1 = c[1]; dc2 = c[2]; dc3 = c[3]; }
    if (do_warp) {
      X.t00 = nx.t00; X.t01 = nx.t01; X.t02 = nx.t02; X.t10 = nx.t10; X.t11 = nx.t11; X.t12 = nx.t12; X.flip = nx.flip;
      // pixel-space affine  ix = a*jw + bb*i + c0 ; iy = d*jw + e*i + f0  (approximate, for boxes only)
      const float a = X.t00 * X.stepx * X.sfx, bb = X.t01 * X.stepy * X.sfx;
      const float d = X.t10 * X.stepx * X.sfy, e = X.t11 * X.stepy * X.sfy;
      A.c0 = (X.t02 + 1.f - X.t00 - X.t01) * X.sfx; A.f0 = (X.t12 + 1.f - X.t10 - X.t11) * X.sfy;
      const float det = a * e - bb * d;
      const float nrm = fabsf(a) + fabsf(bb) + fabsf(d) + fabsf(e);
      bad_xform = !(fabsf(det) > 1e-5f * nrm * nrm) || !(nrm < 1e4f) || W <= 1 || H <= 1;
      const float idet = 1.f / det;
      A.C00 = e * idet; A.C01 = -bb * idet; A.C10 = -d * idet; A.C11 = a * idet;
    }
    if (use_bulk) {
      const unsigned long long t0 = tl ? tl_now() : 0ull;
      mbar_wait(bar, (uint32_t)(it & 1));
      if (tl) {
        const unsigned long long t1 = tl_now();
        tl_wait += t1 - t0;
        if (it == 0) tl_first = t1;
      }
      if (p.inflight_cap > 0 && lane == 0) atomicAdd(&cj->landed, 1u);
    } else {
      const float* gsrc = map_src<FAST>(p, n);
      for (int k = lane; k < HW; k += 32) buf0[k] = __ldg(gsrc + k);
      __syncwarp();
    }
    ++n_maps;

    // ---- pass A: raw max / location / min / finiteness -----------------------------------------
    float bv, mn; int bq;
    if (dbg & 8) { bv = s[lane]; bq = lane >> 2; mn = bv; }
    else scan_max(s, HW, lane, bv, bq, mn);
    A.lane_max = bv;                       // per-lane float4 maximum (pass B reuses it)
    int bi = 0x7fffffff;                  // flat index of the lane's first maximum
    if (bv > -INFINITY) {
      const float4 x = reinterpret_cast<const float4*>(s)[bq];
      bi = (bq << 2) + ((x.x == bv) ? 0 : (x.y == bv) ? 1 : (x.z == bv) ? 2 : 3);
    }
    for (int k = ((HW >> 2) << 2) + lane; k < HW; k += 32) {   // tail when H*W % 4 != 0
      const float x = s[k];
      if (x > bv) { bv = x; bi = k; }
      mn = min3_nan(x, x, mn);
    }
    // NaN -> mn is NaN; -inf -> mn == -inf; +inf -> bv == +inf
    const bool nonfinite = __any_sync(0xffffffffu, !(mn >= -FLT_MAX) || !(bv <= FLT_MAX));
    float rv = bv; int ri = bi;           // result (value, canonical flat index)

    if (!do_warp) {
      if (nonfinite) {                    // torch.max: the first NaN wins; +-Inf compare normally
        rv = -INFINITY; ri = 0x7fffffff;
        for (int k = lane; k < HW; k += 32) { const float x = s[k]; if (arg_better(x, k, rv, ri)) { rv = x; ri = k; } }
        warp_argmax(rv, ri);
      } else {
        warp_argmax_finite(rv, ri);
      }
    } else {
      bool exhaustive = nonfinite || bad_xform;
      if (!exhaustive) {
        warp_argmax_finite(bv, bi);       // warp-uniform source max / location
        A.bv = bv; A.bi = bi;
        if (dbg & 1) { rv = bv; ri = bi; }
        else decode_pruned<CH, CW>(p, s, X, lx, ly, A, lane, rv, ri, exhaustive, n_eval);
      }
      if (exhaustive && (dbg & 5)) { exhaustive = false; rv = bv; ri = bi & 0xfff; }
      if (exhaustive) {
        // NaN-aware compare for non-finite maps and for degenerate / huge transforms (their grid can overflow to
        // Inf - Inf = NaN weights); everything else yields finite samples
        const unsigned long long t0 = tl ? tl_now() : 0ull;
        const ArgMax r = coop_exhaustive<CH, CW>(cj, p, s, X, nonfinite || bad_xform, lx, ly, warp, lane);
        if (tl) tl_exh += tl_now() - t0;
        rv = r.v; ri = r.i;
        ++n_slow;
        n_eval += HW;
      }
    }

    fetch_next();                                        // next map + its transform loads, consumed after the epilogue
    if (do_warp) {                                     // second look at the CTA's job word, half a map after the first
      int open = 0;
      if (lane == 0) open = ld_volatile_s32(&cj->band_next) < kBands ? 1 : 0;
      if (__shfl_sync(0xffffffffu, open, 0)) {
        const unsigned long long t0 = tl ? tl_now() : 0ull;
        coop_help<CH, CW>(cj, p, lx, ly, lane);
        if (tl) tl_help += tl_now() - t0;
      }
    }
    if (!(dbg & 2)) {
      k2_resolve(p, pend_item, pend_old, lane);          // the previous map's ticket has long arrived by now
      finish_map<FAST, CH, CW>(p, n, (int)vu, b, j, s, X, lx, ly, rv, ri, dc0, dc1, dc2, dc3, lane, pend_item, pend_old);
    } else if (rv == 123.456f && p.out_max) p.out_max[n] = rv;
    __syncwarp();
    if (use_bulk && lane == 0 && nx.n < N) {                                                    // buffer handed on
      const unsigned long long t0 = tl ? tl_now() : 0ull;
      issue_map<FAST>(p, nx.n, buf0, bar, pol, map_bytes, cj);
      if (tl) tl_ticket += tl_now() - t0;
    }
  }
  k2_resolve(p, pend_item, pend_old, lane);
  if (tl && lane == 0 && n_maps) {
    const unsigned long long t = tl_now();
    atomicMin(p.stats + 10, tl_first); atomicMax(p.stats + 11, tl_first);
    atomicMin(p.stats + 12, t); atomicMax(p.stats + 13, t);
    atomicAdd(p.stats + 15, tl_wait); atomicAdd(p.stats + 16, tl_ticket);
    atomicAdd(p.stats + 17, tl_exh); atomicAdd(p.stats + 18, tl_help); atomicAdd(p.stats + 19, t - tl_first);
    if (dbg & 32) {                                    // per-warp record: stats[32 + 2*g] = time out of maps,
      const unsigned g = blockIdx.x * 16u + (unsigned)warp;   // [33 + 2*g] = maps | exhaustive << 16 | exhaustive ns << 32
      p.stats[32 + 2 * g] = t;
      p.stats[33 + 2 * g] = n_maps | (n_slow << 16) | (tl_exh << 32);
    }
  }
  if (p.stats && lane == 0 && n_maps) {
    atomicAdd(p.stats + 0, n_slow);
    atomicAdd(p.stats + 1, n_eval);
    atomicAdd(p.stats + 2, n_maps);
  }
  // Out of maps: stay while other warps of the CTA may still post an exhaustive job.  Meanwhile -- HBM is going idle as
  // the last maps finish -- do the EMA's pieces (TailEma: real work that would otherwise be a launch of its own; the
  // warp leaves only when no piece is left), then pull the range the next kernel reads (pf_ptr: the student maps of
  // K3) into L2 for as long as other warps are still decoding.
  if (do_warp) {
    __syncwarp();
    if (lane == 0) atomicAdd(&cj->warps_done, 1);
    constexpr unsigned kPfChunk = 32768u;
    bool pf_live = p.pf_ptr != nullptr;
    bool ema_live = EMA && p.ema.n_chunks > 0;
    const int ema_pieces = (p.ema.chunk_elems + kEmaPiece - 1) / kEmaPiece;
    const unsigned ema_total = (unsigned)p.ema.n_chunks * (unsigned)ema_pieces;
    float ea = p.ema.a, eoma = p.ema.oma;
    if (ema_live && p.ema.alpha_dev) { ea = p.ema.alpha_dev[0]; eoma = p.ema.alpha_dev[1]; }
    int tick = 0;                                        // one chunk per pf_every polls: the idle warps ask for about
    for (;;) {                                           // the bandwidth they used while they were decoding
      int st = 0;                                        // lane 0 decides for the warp: 2 leave, 1 help, 0 idle
      if (lane == 0) st = (!ema_live && ld_volatile_s32(&cj->warps_done) >= warps) ? 2 : (ld_volatile_s32(&cj->band_next) < kBands ? 1 : 0);
      st = __shfl_sync(0xffffffffu, st, 0);
      if (st == 2) break;
      if (st == 1) {
        coop_help<CH, CW>(cj, p, lx, ly, lane);
        continue;
      }
      if (ema_live) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(p.ema.next, 1ull);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= (unsigned long long)ema_total) { ema_live = false; continue; }
        ema_piece(p.ema, (unsigned)c, ema_pieces, ea, eoma, lane);
        continue;
      }
      if (pf_live && (tick++ % p.pf_every) == 0) {
        unsigned long long c = 0;
        if (lane == 0) c = atomicAdd(p.pf_next, 1ull);
        c = __shfl_sync(0xffffffffu, c, 0);
        const unsigned long long off = c * kPfChunk;
        if (off >= p.pf_bytes) { pf_live = false; continue; }
        const unsigned long long left = p.pf_bytes - off;
        if (lane == 0) bulk_prefetch_l2(p.pf_ptr + off, (uint32_t)(left < kPfChunk ? left : kPfChunk));
      }
      __nanosleep(200);
    }
  }
  if (tl && lane == 0) atomicMax(p.stats + 14, tl_now());
}

// -----------------------------------------------------------------------------------------------
// affine_back2 materialised: one CTA per map, source staged in shared memory, coalesced stores.
// swap_perm (optional): output channel c of a FLIPPED sample is warped from source channel swap_perm[c]
// (utils/udaap/transforms.py:20-57 flip_back); NULL = no exchange (utils/augment.py:37-47).
// -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) warp_materialize_kernel(const float* __restrict__ in, long long sN, long long sC,
                                                               float* __restrict__ out, long long oN, long long oC,
                                                               int N, int C, int H, int W,
                                                               const float* __restrict__ theta,
                                                               const uint8_t* __restrict__ flip,
                                                               const int32_t* __restrict__ swap_perm, int vec) {
  extern __shared__ __align__(16) float sm[];
  const int HW = H * W;
  for (long long m = blockIdx.x; m < (long long)N * C; m += gridDim.x) {
    const long long n = m / C;
    const int c = (int)(m % C);
    const int cs = (swap_perm && flip && flip[n]) ? swap_perm[c] : c;
    const float* src = in + n * sN + (long long)cs * sC;
    float* dst = out + n * oN + (long long)c * oC;
    __syncthreads();
    if (vec) {                                               // 128-bit streaming loads of the source map
      const float4* s4 = reinterpret_cast<const float4*>(src);
      for (int q = threadIdx.x; q < (HW >> 2); q += blockDim.x) reinterpret_cast<float4*>(sm)[q] = ldg_stream(s4 + q);
    } else {
      for (int k = threadIdx.x; k < HW; k += blockDim.x) sm[k] = __ldg(src + k);
    }
    __syncthreads();
    Xform X;
    load_xform(X, theta, flip, n, H, W);
    grid_consts(X, H, W);
    if (vec) {
      // four consecutive output pixels of a row per thread, one 128-bit store.  When the CTA covers whole rows per sweep
      // (blockDim % (W/4) == 0) a thread keeps its four columns for the whole map: their base-grid coordinates and the
      // products with t00 / t10 are formed once (12 registers) and a pixel costs its row term plus the sample
      const int w4 = W >> 2;
      if ((int)blockDim.x % w4 == 0) {
        const int rows = (int)blockDim.x / w4, i0 = (int)threadIdx.x / w4, jo = ((int)threadIdx.x - i0 * w4) << 2;
        float ax[4], ay[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float xl = lin_coord(X.flip ? (W - 1 - u - jo) : jo + u, W, X.stepx);
          ax[u] = __fmul_rn(xl, X.t00); ay[u] = __fmul_rn(xl, X.t10);
        }
        for (int i = i0; i < H; i += rows) {
          const float yl = lin_coord(i, H, X.stepy);
          float4 o;
          o.x = eval_col(sm, X, ax[0], ay[0], yl); o.y = eval_col(sm, X, ax[1], ay[1], yl);
          o.z = eval_col(sm, X, ax[2], ay[2], yl); o.w = eval_col(sm, X, ax[3], ay[3], yl);
          stg_stream(reinterpret_cast<float4*>(dst) + i * w4 + (jo >> 2), o);
        }
      } else {
        for (int q = threadIdx.x; q < (HW >> 2); q += blockDim.x) {
          const int i = q / w4, jo = (q - i * w4) << 2;
          const float yl = lin_coord(i, H, X.stepy);
          float4 o;
          o.x = eval_at(sm, X, lin_coord(X.flip ? (W - 1 - jo) : jo, W, X.stepx), yl);
          o.y = eval_at(sm, X, lin_coord(X.flip ? (W - 2 - jo) : jo + 1, W, X.stepx), yl);
          o.z = eval_at(sm, X, lin_coord(X.flip ? (W - 3 - jo) : jo + 2, W, X.stepx), yl);
          o.w = eval_at(sm, X, lin_coord(X.flip ? (W - 4 - jo) : jo + 3, W, X.stepx), yl);
          stg_stream(reinterpret_cast<float4*>(dst) + q, o);
        }
      }
    } else {
      for (int k = threadIdx.x; k < HW; k += blockDim.x) {
        const int i = k / W, jo = k - i * W;
        dst[k] = eval_px(sm, X, i, X.flip ? (W - 1 - jo) : jo);
      }
    }
  }
}

// fliplr_back_tensor (utils/augment.py:247-252): exact mirror of the W axis, row per warp.
__global__ void mirror_w_kernel(const float* __restrict__ in, float* __restrict__ out, long long rows, int W) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* src = in + r * W;
  float* dst = out + r * W;
  for (int x = threadIdx.x & 31; x < W; x += 32) dst[x] = src[W - 1 - x];
}

}  // namespace ubpl

using namespace ubpl;

static int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// The call the fused chain makes (see UBPL_F): it gets the specialised kernel instance.  UBPL_K1_FAST=0 forces the generic one.
static bool k1_fast(const WDParams& p) {
  return p.do_warp && p.use_bulk && p.refine == 0 && !p.swap_perm && !p.out_hm_xy && p.dbg == 0 && env_int("UBPL_K1_FAST", 1) != 0;
}

// Fills the geometry / tuning fields of p and launches the kernel.  p.work, p.k2, p.pf_* and the outputs are
// set by the caller.
static int launch_k1(WDParams& p, cudaStream_t stream, bool* ema_carried = nullptr) {
  const int H = p.H, W = p.W;
  const long long N = (long long)p.V * p.B * p.J;
  const long long HW = (long long)H * W;
  UBPL_REQUIRE(HW <= (1 << 24), "ubpl_warp_decode: heat-map too large (%lld texels)", HW);
  UBPL_REQUIRE(N < (1ll << 31), "ubpl_warp_decode: too many maps in one call (%lld)", N);
  const size_t map_bytes = (size_t)HW * 4;
  p.divJ.init((unsigned)p.J); p.divB.init((unsigned)p.B); p.divW.init((unsigned)W);
  p.stepx = (W > 1) ? 2.f / (float)(W - 1) : 0.f;
  p.stepy = (H > 1) ? 2.f / (float)(H - 1) : 0.f;
  p.sfx = (float)((double)(W - 1) / 2.0);
  p.sfy = (float)((double)(H - 1) / 2.0);
  p.dbg = env_int("UBPL_K1_DBG", 0);
  // at most 8 staged copies in flight per SM (1184 over the GPU) by default: measured on B200, more concurrent 16 KB
  // streams than that can tip HBM into a regime that delivers 4.3-4.6 TB/s instead of 6+ (tools/k1_micro.cu,
  // profiles/README.md round 2); UBPL_K1_INFLIGHT overrides (0 = no cap)
  p.inflight_cap = env_int("UBPL_K1_INFLIGHT", 8);
  // bulk async copy needs 16-byte aligned sources and sizes; otherwise the warp copies the map itself
  p.use_bulk = ((reinterpret_cast<uintptr_t>(p.maps) & 15) == 0) && (map_bytes % 16 == 0) && (p.sV % 4 == 0) &&
               (p.sB % 4 == 0) && (p.sJ % 4 == 0);
  // shared memory: one staging buffer per warp, then 16 mbarriers (128 B), the CTA's CoopJob (256 B) and the
  // base-grid tables lx[W], ly[H]
  static_assert(sizeof(CoopJob) <= 256, "CoopJob must fit its shared-memory slot");
  const size_t buf_stride = (map_bytes + 127) & ~(size_t)127;
  const size_t tail = 384 + (((size_t)(W + H) * 4 + 127) & ~(size_t)127);
  UBPL_REQUIRE(buf_stride + tail <= (size_t)smem_optin(), "ubpl_warp_decode: a %dx%d map does not fit in shared memory", H, W);
  int warps = (int)(((size_t)smem_optin() - tail) / buf_stride);
  // tuning knob (read per call so that one process can compare settings): UBPL_K1_WARPS caps the warps per CTA
  const int env_warps = env_int("UBPL_K1_WARPS", 0);
  if (env_warps > 0 && env_warps < warps) warps = env_warps;
  if (warps > 16) warps = 16;
  if (warps < 1) warps = 1;
  const size_t smem = (size_t)warps * buf_stride + tail;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaSuccess;
    const void* inst[] = {(const void*)warp_decode_kernel<false, false, false, 0, 0>, (const void*)warp_decode_kernel<true, true, false, 0, 0>,
                          (const void*)warp_decode_kernel<false, false, true, 0, 0>, (const void*)warp_decode_kernel<false, true, true, 0, 0>,
                          (const void*)warp_decode_kernel<false, false, true, 64, 64>, (const void*)warp_decode_kernel<false, true, true, 64, 64>,
                          (const void*)warp_decode_kernel<false, false, true, 128, 128>, (const void*)warp_decode_kernel<false, true, true, 128, 128>};
    for (const void* f : inst)
      if (e == cudaSuccess) e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin());
    if (e != cudaSuccess) { set_error("ubpl_warp_decode: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
    attr_set = true;
  }
  long long need = (N + warps - 1) / warps;
  int grid = (int)(need < sm_count() ? need : sm_count());
  // the EMA rides along only in the specialised instance (and in the probe build); otherwise the caller issues it
  const bool probe = (p.dbg & 16) && p.stats;
  if (p.ema.n_chunks > 0 && !probe && !k1_fast(p)) p.ema.n_chunks = 0;
  if (ema_carried) *ema_carried = p.ema.n_chunks > 0;
  // instance: probe build / generic / the fused chain's call, the latter compiled for 64x64 and 128x128 maps as well
  // (UBPL_K1_SHAPES=0: shapes read at run time)
  const bool fast = k1_fast(p), ema = p.ema.n_chunks > 0;
  const int shape = (fast && env_int("UBPL_K1_SHAPES", 1)) ? ((H == 64 && W == 64) ? 64 : (H == 128 && W == 128) ? 128 : 0) : 0;
#define UBPL_K1_LAUNCH(...) warp_decode_kernel<__VA_ARGS__><<<grid, warps * 32, smem, stream>>>(p)
  if (probe) UBPL_K1_LAUNCH(true, true, false, 0, 0);
  else if (!fast) UBPL_K1_LAUNCH(false, false, false, 0, 0);
  else if (shape == 64) { if (ema) UBPL_K1_LAUNCH(false, true, true, 64, 64); else UBPL_K1_LAUNCH(false, false, true, 64, 64); }
  else if (shape == 128) { if (ema) UBPL_K1_LAUNCH(false, true, true, 128, 128); else UBPL_K1_LAUNCH(false, false, true, 128, 128); }
  else { if (ema) UBPL_K1_LAUNCH(false, true, true, 0, 0); else UBPL_K1_LAUNCH(false, false, true, 0, 0); }
#undef UBPL_K1_LAUNCH
  return check_launch("ubpl_warp_decode");
}

extern "C" int ubpl_warp_decode(const float* maps, int64_t sV, int64_t sB, int64_t sJ, int V, int B, int J, int H,
                                int W, const float* theta, const uint8_t* flip, const int32_t* swap_perm,
                                const double* dec, int do_warp, int refine, int32_t* out_idx, float* out_max,
                                float* out_xy, float* out_hm_xy, int64_t* stats, int32_t* ws, void* stream) {
  UBPL_REQUIRE(maps != nullptr, "ubpl_warp_decode: maps is NULL");
  UBPL_REQUIRE(V >= 0 && B >= 0 && J >= 0 && H > 0 && W > 0, "ubpl_warp_decode: bad dims V=%d B=%d J=%d H=%d W=%d", V, B, J, H, W);
  UBPL_REQUIRE(!do_warp || theta != nullptr, "ubpl_warp_decode: theta is NULL with do_warp=1");
  UBPL_REQUIRE(refine >= 0 && refine <= 2, "ubpl_warp_decode: refine must be 0, 1 or 2");
  UBPL_REQUIRE(!ws || (reinterpret_cast<uintptr_t>(ws) & 7) == 0, "ubpl_warp_decode: ws must be 8-byte aligned");
  const long long N = (long long)V * B * J;
  if (N == 0) return UBPL_OK;
  WDParams p;
  memset(&p, 0, sizeof(p));
  p.maps = maps; p.sV = sV; p.sB = sB; p.sJ = sJ; p.V = V; p.B = B; p.J = J; p.H = H; p.W = W;
  p.theta = theta; p.flip = flip; p.swap_perm = (do_warp && flip) ? swap_perm : nullptr; p.dec = dec;
  p.do_warp = do_warp; p.refine = refine;
  p.out_idx = out_idx; p.out_max = out_max; p.out_xy = out_xy; p.out_hm_xy = out_hm_xy;
  p.stats = reinterpret_cast<unsigned long long*>(stats);
  if (ws) {
    // the caller's own claim counter: private to this launch (and to a CUDA graph that captured it)
    cudaError_t e = cudaMemsetAsync(ws, 0, 8, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("ubpl_warp_decode: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
    p.work = reinterpret_cast<unsigned long long*>(ws);
  } else {
    p.work = work_counter((cudaStream_t)stream);
    if (!p.work) return UBPL_ERR_CUDA;
  }
  return launch_k1(p, (cudaStream_t)stream);
}

// Workspace of ubpl_warp_decode_k2(_ema), int32 words: [0,1] claim counter, [32,33] prefetch chunk counter, [34] number
// of items with a zero intDist sum, [35] status, [64,65] EMA work-item counter (own 128-byte lines for the three
// counters), [128, 128+J+2) counts,
// then the B*J arrival counters and the V*B*J 64-bit hand-off words.  All of it is cleared by ONE memset node per
// launch.
static inline long long k2_ws_arrive_off(int J) { return 128 + ((J + 2 + 1) & ~1); }
static inline long long k2_ws_slots_off(int B, int J) { return (k2_ws_arrive_off(J) + (long long)B * J + 1) & ~1ll; }
static inline long long k2_ws_zero_words(int V, int B, int J) { return k2_ws_slots_off(B, J) + 2ll * V * B * J; }

extern "C" int64_t ubpl_warp_decode_k2_ws_bytes(int V, int B, int J) {
  if (V < 0 || B < 0 || J < 0) return 0;
  return 4 * (k2_ws_zero_words(V, B, J) + 4);
}

static int warp_decode_k2_impl(const float* maps, int64_t sV, int64_t sB, int64_t sJ, int V, int B, int J, int H,
                               int W, const float* theta, const uint8_t* flip, const int32_t* swap_perm,
                               const double* dec, int refine, int32_t* out_idx, float* out_max, float* out_xy,
                               int k2_mode, double distThrMax, int img_h, int img_w, float stride, float sigma,
                               int S, float* mean, double* dist, uint8_t* legal, uint8_t* enable, float* gate,
                               int64_t* stats, int32_t* ws, int64_t ws_bytes, const void* prefetch,
                               int64_t prefetch_bytes, const TailEma* ema, void* stream) {
  UBPL_REQUIRE(V >= 1 && V <= 32 && B >= 0 && J >= 1 && H > 0 && W > 0, "ubpl_warp_decode_k2: bad dims V=%d B=%d J=%d H=%d W=%d (1 <= V <= 32)", V, B, J, H, W);
  UBPL_REQUIRE(ws && (B == 0 || (maps && theta && out_xy)), "ubpl_warp_decode_k2: NULL pointer");
  UBPL_REQUIRE(refine >= 0 && refine <= 2, "ubpl_warp_decode_k2: refine must be 0, 1 or 2");
  UBPL_REQUIRE(k2_mode >= 1 && k2_mode <= 4, "ubpl_warp_decode_k2: k2_mode must be 1, 2 (one teacher) or 3, 4 (two teachers)");
  UBPL_REQUIRE(k2_mode < 3 || (V % 2 == 0), "ubpl_warp_decode_k2: two teachers need an even number of maps per key point (V = 2K)");
  UBPL_REQUIRE((k2_mode & 1) || ((gate || B == 0) && S >= 1 && stride > 0.f && sigma > 0.f), "ubpl_warp_decode_k2: modes 2 and 4 need gate, S, stride, sigma");
  UBPL_REQUIRE(ws_bytes >= ubpl_warp_decode_k2_ws_bytes(V, B, J), "ubpl_warp_decode_k2: workspace too small");
  UBPL_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 7) == 0, "ubpl_warp_decode_k2: workspace must be 8-byte aligned");
  UBPL_REQUIRE(!prefetch || ((reinterpret_cast<uintptr_t>(prefetch) & 15) == 0 && prefetch_bytes >= 0),
               "ubpl_warp_decode_k2: the prefetch range must be 16-byte aligned");
  const long long zero_words = k2_ws_zero_words(V, B, J);
  cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)zero_words * 4, (cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("ubpl_warp_decode_k2: memset: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
  const long long N = (long long)V * B * J;
  if (N == 0) {                                        // nothing to decode: the EMA still has to happen
    if (ema && ema->n_chunks > 0)
      return ubpl_ema_multi_tensor(ema->ema_ptrs, ema->param_ptrs, reinterpret_cast<const int64_t*>(ema->numels), ema->chunk_tensor,
                                   reinterpret_cast<const int64_t*>(ema->chunk_start), ema->n_chunks, ema->chunk_elems, ema->a,
                                   ema->oma, ema->alpha_dev, stream);
    return UBPL_OK;
  }
  WDParams p;
  memset(&p, 0, sizeof(p));
  // The EMA rides in this launch when the launch is short enough for that to pay: the kernel instance that can do the
  // EMA decodes 2-4 % slower (see warp_decode_kernel), a launch of its own costs the EMA ~10 us more than K1's tail
  // does -- break-even around 1.9 GB of maps.  Above 1.5 GB (UBPL_K1_EMA_MAX_MB) the EMA follows as its own launch.
  bool ema_after = false;
  if (ema && ema->n_chunks > 0) {
    const long long max_mb = env_int("UBPL_K1_EMA_MAX_MB", 1536);
    if (N * (long long)H * W * 4 <= (max_mb << 20)) {
      p.ema = *ema;
      p.ema.next = reinterpret_cast<unsigned long long*>(ws + 64);    // its own 128-byte line, cleared by the memset above
    } else {
      ema_after = true;
    }
  }
  p.maps = maps; p.sV = sV; p.sB = sB; p.sJ = sJ; p.V = V; p.B = B; p.J = J; p.H = H; p.W = W;
  p.theta = theta; p.flip = flip; p.swap_perm = flip ? swap_perm : nullptr; p.dec = dec; p.do_warp = 1; p.refine = refine;
  p.out_idx = out_idx; p.out_max = out_max; p.out_xy = out_xy; p.out_hm_xy = nullptr;
  p.stats = reinterpret_cast<unsigned long long*>(stats);
  p.work = reinterpret_cast<unsigned long long*>(ws);
  if (prefetch && prefetch_bytes >= 16) {
    p.pf_ptr = reinterpret_cast<const unsigned char*>(prefetch);
    p.pf_bytes = (unsigned long long)prefetch_bytes & ~15ull;
    p.pf_next = reinterpret_cast<unsigned long long*>(ws + 32);
    p.pf_every = env_int("UBPL_K1_PF_EVERY", 8);
    if (p.pf_every < 1) p.pf_every = 1;
  }
  K2Fuse& f = p.k2;
  f.mode = k2_mode; f.K = V;
  f.counts = ws + 128;
  f.zero_div = ws + 34;
  f.status = ws + 35;
  f.arrive = reinterpret_cast<unsigned*>(ws + k2_ws_arrive_off(J));
  f.slots = reinterpret_cast<unsigned long long*>(ws + k2_ws_slots_off(B, J));
  f.distThrMax = distThrMax; f.img_h = img_h; f.img_w = img_w; f.S = S; f.stride = stride; f.sigma = sigma;
  f.thr = 1.0 - exp(-(distThrMax * 3.0) / 5.0);       // business.py:375-376 with CPython's libm exp
  f.mean = mean; f.dist = dist; f.legal = legal; f.enable = enable; f.gate = gate;
  int rc = pow_table(&f.T.key, &f.T.val, &f.T.bits, &f.T.n, &f.T.rmax);
  if (rc != UBPL_OK) return rc;
  bool carried = false;
  rc = launch_k1(p, (cudaStream_t)stream, &carried);
  if (ema && ema->n_chunks > 0 && !carried) ema_after = true;
  if (rc == UBPL_OK && ema_after)
    rc = ubpl_ema_multi_tensor(ema->ema_ptrs, ema->param_ptrs, reinterpret_cast<const int64_t*>(ema->numels), ema->chunk_tensor,
                               reinterpret_cast<const int64_t*>(ema->chunk_start), ema->n_chunks, ema->chunk_elems, ema->a,
                               ema->oma, ema->alpha_dev, stream);
  return rc;
}

extern "C" int ubpl_warp_decode_k2(const float* maps, int64_t sV, int64_t sB, int64_t sJ, int V, int B, int J, int H,
                                   int W, const float* theta, const uint8_t* flip, const int32_t* swap_perm,
                                   const double* dec, int refine, int32_t* out_idx, float* out_max, float* out_xy,
                                   int k2_mode, double distThrMax, int img_h, int img_w, float stride, float sigma,
                                   int S, float* mean, double* dist, uint8_t* legal, uint8_t* enable, float* gate,
                                   int64_t* stats, int32_t* ws, int64_t ws_bytes, const void* prefetch,
                                   int64_t prefetch_bytes, void* stream) {
  return warp_decode_k2_impl(maps, sV, sB, sJ, V, B, J, H, W, theta, flip, swap_perm, dec, refine, out_idx, out_max, out_xy,
                             k2_mode, distThrMax, img_h, img_w, stride, sigma, S, mean, dist, legal, enable, gate, stats, ws,
                             ws_bytes, prefetch, prefetch_bytes, nullptr, stream);
}

extern "C" int ubpl_warp_decode_k2_ema(const float* maps, int64_t sV, int64_t sB, int64_t sJ, int V, int B, int J, int H,
                                       int W, const float* theta, const uint8_t* flip, const int32_t* swap_perm,
                                       const double* dec, int refine, int32_t* out_idx, float* out_max, float* out_xy,
                                       int k2_mode, double distThrMax, int img_h, int img_w, float stride, float sigma,
                                       int S, float* mean, double* dist, uint8_t* legal, uint8_t* enable, float* gate,
                                       int64_t* stats, int32_t* ws, int64_t ws_bytes, const void* prefetch,
                                       int64_t prefetch_bytes, const uint64_t* ema_ptrs, const uint64_t* param_ptrs,
                                       const int64_t* numels, const int32_t* chunk_tensor, const int64_t* chunk_start,
                                       int64_t n_chunks, int chunk_elems, float alpha, float one_minus_alpha,
                                       const float* alpha_dev, void* stream) {
  UBPL_REQUIRE(n_chunks >= 0 && (n_chunks == 0 || (ema_ptrs && param_ptrs && numels && chunk_tensor && chunk_start)),
               "ubpl_warp_decode_k2_ema: NULL pointer in the EMA tables");
  UBPL_REQUIRE(n_chunks == 0 || (chunk_elems > 0 && chunk_elems % 4 == 0), "ubpl_warp_decode_k2_ema: bad chunking");
  UBPL_REQUIRE(n_chunks * (int64_t)((chunk_elems + kEmaPiece - 1) / kEmaPiece) < (1ll << 31),
               "ubpl_warp_decode_k2_ema: too many EMA work items (%lld chunks of %d elements)", (long long)n_chunks, chunk_elems);
  TailEma E;
  memset(&E, 0, sizeof(E));
  E.ema_ptrs = ema_ptrs; E.param_ptrs = param_ptrs; E.numels = reinterpret_cast<const long long*>(numels);
  E.chunk_tensor = chunk_tensor; E.chunk_start = reinterpret_cast<const long long*>(chunk_start);
  E.n_chunks = n_chunks; E.chunk_elems = chunk_elems; E.a = alpha; E.oma = one_minus_alpha; E.alpha_dev = alpha_dev;
  return warp_decode_k2_impl(maps, sV, sB, sJ, V, B, J, H, W, theta, flip, swap_perm, dec, refine, out_idx, out_max, out_xy,
                             k2_mode, distThrMax, img_h, img_w, stride, sigma, S, mean, dist, legal, enable, gate, stats, ws,
                             ws_bytes, prefetch, prefetch_bytes, &E, stream);
}

extern "C" int ubpl_warp_materialize(const float* in, int64_t sN, int64_t sC, float* out, int64_t oN, int64_t oC,
                                     int N, int C, int H, int W, const float* theta, const uint8_t* flip,
                                     const int32_t* swap_perm, void* stream) {
  UBPL_REQUIRE(in && out && theta, "ubpl_warp_materialize: NULL pointer");
  UBPL_REQUIRE(N >= 0 && C >= 0 && H > 0 && W > 0, "ubpl_warp_materialize: bad dims");
  if ((long long)N * C == 0) return UBPL_OK;
  const size_t smem = (size_t)H * W * 4;
  UBPL_REQUIRE((int)smem <= smem_optin(), "ubpl_warp_materialize: map does not fit in shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(warp_materialize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin());
    if (e != cudaSuccess) { set_error("ubpl_warp_materialize: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return UBPL_ERR_CUDA; }
    attr_set = true;
  }
  long long maps = (long long)N * C;
  int grid = (int)(maps < (long long)sm_count() * 8 ? maps : (long long)sm_count() * 8);
  const int vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                  sN % 4 == 0 && sC % 4 == 0 && oN % 4 == 0 && oC % 4 == 0;
  warp_materialize_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(in, sN, sC, out, oN, oC, N, C, H, W, theta, flip,
                                                                      flip ? swap_perm : nullptr, vec);
  return check_launch("ubpl_warp_materialize");
}

extern "C" int ubpl_mirror_w(const float* in, float* out, int64_t rows, int W, void* stream) {
  UBPL_REQUIRE(in && out && rows >= 0 && W > 0, "ubpl_mirror_w: bad arguments");
  if (rows == 0) return UBPL_OK;
  const int wpb = 8;
  mirror_w_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(in, out, rows, W);
  return check_launch("ubpl_mirror_w");
}
