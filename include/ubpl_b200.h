/*
 * ubpl_b200.h -- C ABI of libubpl_b200.so: the B200 (sm_100a) pseudo-label hot path of
 * Qi2019KB/UBPL-PoseEstimation.
 *
 * The reference is pure Python (no FFI of its own); every entry point below
 * replaces the body of one reference function, cited as file:line relative to /root/reference.
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in _host; the caller owns every
 *     buffer, kernels never allocate; outputs are pre-allocated by the caller;
 *   - heat-maps: the inner [H, W] plane is contiguous; outer dims are addressed with ELEMENT
 *     strides (the reference slices `outs_ema[m, a, :, -1]`, so outer strides are arbitrary);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success or a negative code; ubpl_last_error() gives the text.
 *     Nothing falls back to the CPU: without a usable sm_100 device the calls fail.
 */
#ifndef UBPL_B200_H
#define UBPL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UBPL_OK 0
#define UBPL_ERR_INVALID (-1)
#define UBPL_ERR_CUDA (-2)
#define UBPL_ERR_UNSUPPORTED (-3)

const char* ubpl_last_error(void);
int ubpl_version(void);
/* sm count, compute capability and opt-in shared memory per block of the current device. */
int ubpl_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes);

/* ---- K1: back-warp + flip + arg-max decode, fused (each heat-map is read from HBM once) --------
 * Replaces AugmentUtils.affine_back2 (utils/augment.py:37-47) followed by get_preds /
 * final_preds (utils/udaap/evaluation.py:13-30,215-238), transform_preds
 * (utils/udaap/transforms.py:151-168) and the scores of ProcessUtils.kps_fromHeatmap(_mul)
 * (utils/process.py:321-336); refine != 0 adds the quarter-offset of kps_fromHeatmap2
 * (utils/process.py:362-373).
 *
 * maps    [V, B, J, H, W] float32 with element strides (sV, sB, sJ)
 * theta   [V, B, 2, 3] float32 contiguous, NULL when do_warp == 0 (plain decode of raw maps)
 * flip    [V, B] uint8 contiguous (NULL = no flips)
 * dec     [B, 4] float64 (a00, a02, a11, a12) of the inverse decode transform (NULL = skip the
 *         image-space transform; out_xy then equals out_hm_xy)
 * refine  0 = off (final_preds parity), 1 = joints 0 and 1 only (kps_fromHeatmap2 bug parity,
 *         utils/process.py:363), 2 = all joints
 * outputs, all contiguous [V, B, J]: out_idx int32 (flat arg-max index in the canonical frame,
 *         first maximum in row-major order), out_max float32 (unmasked maximum = `scores`),
 *         out_xy float32[2] (image space, integer valued), out_hm_xy float32[2] (1-based heat-map
 *         coordinates after the max<=0 mask and the optional refinement); any may be NULL.
 * swap_perm optional int32[J] (NULL = none, the reference's live path): output joint j of a FLIPPED
 *         view is decoded from source channel swap_perm[j] -- the left/right exchange of flip_back
 *         (utils/udaap/transforms.py:20-57) folded into the un-flip; ignored when flip is NULL.
 * stats   optional int64[4] device counters, incremented: [0] maps decoded by exhaustive
 *         evaluation, [1] output pixels evaluated, [2] maps, [3] unused.
 * ws      optional int32[>= 2] device scratch, 8-byte aligned: the launch's private work-claim counter
 *         (cleared by the call).  NULL takes a slot of a process-wide ring instead, which a captured CUDA
 *         graph would keep using while eager calls cycle through it -- pass ws when capturing.
 * Maps that cannot be pruned (NaN/Inf, singular theta, structure-less maps) are decoded exhaustively by
 * all warps of the CTA together inside the same launch; there is no second kernel.
 */
int ubpl_warp_decode(const float* maps, int64_t sV, int64_t sB, int64_t sJ,
                     int V, int B, int J, int H, int W,
                     const float* theta, const uint8_t* flip, const int32_t* swap_perm, const double* dec,
                     int do_warp, int refine,
                     int32_t* out_idx, float* out_max, float* out_xy, float* out_hm_xy,
                     int64_t* stats, int32_t* ws, void* stream);

/* K1 with the per-joint part of K2 fused into its epilogue (mean-teacher path: one teacher, V = K views).
 * Decodes like ubpl_warp_decode(do_warp = 1); in addition every decoded map bumps the arrival counter of
 * its (sample, joint) and the warp that completes the K views computes, with the arithmetic of
 * ubpl_view_dispersion / ubpl_k2_view_fixed (utils/evaluation.py:44-54, utils/business.py:237-261,375-376,
 * utils/process.py:262-268, utils/losses.py:29):
 *   k2_mode 1: mean [B,J,2] f32, dist [B,J] f64 (999 where a view is illegal), legal [B,J] u8
 *   k2_mode 2: + enable [B,J] u8, gate [B,J] f32 (= enable * visibility) and the counts below
 *   k2_mode 3: two teachers (V = 2K maps per key point, teacher-major): BusinessUtils.assess_pseudo_unc2
 *              (utils/business.py:109-161) on the float32 view means p1, p2: mean = float32 ensemble coordinate
 *              w1*p1 + w2*p2 with w_m = intDist_m / (intDist_1 + intDist_2), dist = extDist (999 = illegal),
 *              legal; ws[33] counts the key points whose two intDists are both 0 (the reference divides by zero)
 *   k2_mode 4: + the fixed rule on extDist, gate and counts as in mode 2
 * so the chain needs no K2 launch (modes 2, 4) or only the quantile selector (modes 1, 3).
 * ws   int32 workspace of ubpl_warp_decode_k2_ws_bytes(V, B, J) bytes, 8-byte aligned; the call clears it
 *      with one memset node.  After the launch ws[128 .. 128+J) = selected items per joint, ws[128+J] = total
 *      selected, ws[128+J+1] = S * #(gate > 0) (mode 2) -- the `count_in` of ubpl_render_mse; ws[34] = key points
 *      whose two intDists are both 0 (modes 3/4); ws[35] = status: non-zero when a hand-off word of the K2
 *      epilogue never arrived (the launch's K2 outputs are then void; the caller must check it and raise).
 * prefetch / prefetch_bytes (optional, 16-byte aligned): a global range the NEXT kernel will read (the student
 *      maps of K3).  Warps that have run out of maps pull it into L2 in 32 KB chunks while the last maps finish.
 * swap_perm as in ubpl_warp_decode.  Requires 1 <= V <= 32.  mean/dist/legal/enable may be NULL. */
int64_t ubpl_warp_decode_k2_ws_bytes(int V, int B, int J);
int ubpl_warp_decode_k2(const float* maps, int64_t sV, int64_t sB, int64_t sJ,
                        int V, int B, int J, int H, int W,
                        const float* theta, const uint8_t* flip, const int32_t* swap_perm, const double* dec,
                        int refine, int32_t* out_idx, float* out_max, float* out_xy,
                        int k2_mode, double distThrMax, int img_h, int img_w, float stride, float sigma, int S,
                        float* mean, double* dist, uint8_t* legal, uint8_t* enable, float* gate,
                        int64_t* stats, int32_t* ws, int64_t ws_bytes, const void* prefetch, int64_t prefetch_bytes,
                        void* stream);

/* K1 + K2 epilogue as above, with K4 inside: the warps that have run out of maps do the mean-teacher EMA of
 * ubpl_ema_multi_tensor (same tables, same arithmetic, update_ema_variables utils/parameters.py:4-8) chunk by chunk
 * while the last maps are decoded -- K1's launch ends with ~2 map times in which HBM is no longer saturated by the
 * staged copies, and the EMA depends on nothing in the chain.  A separate EMA kernel cannot run beside K1 (K1's CTAs
 * hold all of an SM's shared memory), so this is the only way to overlap the two.  n_chunks = 0: no EMA.  A warp
 * claims 1024 elements of a chunk at a time (one memory round trip), so a table with chunk_elems = 1024 -- one claim per
 * work item, no empty claims on small tensors -- is the efficient one here (the Python side keeps both).  For a
 * launch of more than 1.5 GB of maps (UBPL_K1_EMA_MAX_MB) the call issues the EMA as a launch of its own right behind
 * K1 instead: the kernel instance that carries the EMA code decodes 2-4 % slower, which only pays on short launches. */
int ubpl_warp_decode_k2_ema(const float* maps, int64_t sV, int64_t sB, int64_t sJ,
                        int V, int B, int J, int H, int W,
                        const float* theta, const uint8_t* flip, const int32_t* swap_perm, const double* dec,
                        int refine, int32_t* out_idx, float* out_max, float* out_xy,
                        int k2_mode, double distThrMax, int img_h, int img_w, float stride, float sigma, int S,
                        float* mean, double* dist, uint8_t* legal, uint8_t* enable, float* gate,
                        int64_t* stats, int32_t* ws, int64_t ws_bytes, const void* prefetch, int64_t prefetch_bytes,
                        const uint64_t* ema_ptrs, const uint64_t* param_ptrs, const int64_t* numels,
                            const int32_t* chunk_tensor, const int64_t* chunk_start, int64_t n_chunks, int chunk_elems,
                            float alpha, float one_minus_alpha, const float* alpha_dev, void* stream);

/* Materialises the back-warped (and un-flipped) maps: the tensor AugmentUtils.affine_back2
 * returns (utils/augment.py:37-47).  in [N, C, H, W] strides (sN, sC); out likewise (oN, oC);
 * theta [N,2,3]; flip [N] uint8 or NULL; swap_perm int32[C] or NULL: output channel c of a flipped sample
 * comes from source channel swap_perm[c] (flip_back's left/right exchange, utils/udaap/transforms.py:20-57).
 * Bit-identical to ATen's CPU grid_sample op order. */
int ubpl_warp_materialize(const float* in, int64_t sN, int64_t sC, float* out, int64_t oN, int64_t oC,
                          int N, int C, int H, int W, const float* theta, const uint8_t* flip,
                          const int32_t* swap_perm, void* stream);

/* AugmentUtils.fliplr_back_tensor (utils/augment.py:247-252): out[r, x] = in[r, W-1-x] for `rows`
 * contiguous rows of W floats. */
int ubpl_mirror_w(const float* in, float* out, int64_t rows, int W, void* stream);

/* ---- K2: per-joint uncertainty ----------------------------------------------------------------
 * View dispersion of one teacher: EvaluationUtils.uncertainty_fromDistance
 * (utils/evaluation.py:40-58) numerator.  preds [K, B, J, 2] float32 contiguous; mean_in [B,J,2]
 * float32 = the caller's preds_mean, or NULL to use the float32 view mean.
 * out_mean [B,J,2] float32 (torch.mean over views), out_dist [B,J] float64 (mean distance to the
 * mean), out_unc32 [B,J] float32 (the same, as the float32 tensor the reference builds),
 * out_legal [B,J] uint8 (all views x>=0 and y>=0), max_bits: uint32 device scalar that receives
 * atomicMax of the float32 bits of out_unc32 (caller zeroes it).  sentinel_illegal != 0 stores
 * 999 in out_dist for items with an illegal view (the sentinel of utils/business.py:123). */
int ubpl_view_dispersion(const float* preds, const float* mean_in, int K, int B, int J,
                         float* out_mean, double* out_dist, float* out_unc32, uint8_t* out_legal,
                         uint32_t* max_bits, int sentinel_illegal, void* stream);
/* unc = unc32 / max, uncW = exp(-unc)  (utils/evaluation.py:56-57); n = B*J. */
int ubpl_unc_normalize(const float* unc32, const uint32_t* max_bits, int64_t n,
                       float* out_unc, float* out_uncW, void* stream);

/* Two-teacher assessment: BusinessUtils.assess_pseudo_unc2 (utils/business.py:109-161) in array
 * form.  p1, p2, pmean [B,J,2] float32; aug1, aug2 [K,B,J,2] float32.  Outputs float64 [B,J]:
 * legal, intDist1, intDist2, extDist, w1, w2; coord [B,J,2] (+ coord32, its float32 copy);
 * pmean NULL = bus.preds_mean(p1, p2) (business.py:297-300); zero_div: int32 device counter of
 * items where intDist1+intDist2 == 0 (the reference raises ZeroDivisionError there,
 * business.py:135; the kernel keeps w = 0.5/0.5 and counts). */
int ubpl_assess_dual(const float* p1, const float* p2, const float* pmean,
                     const float* aug1, const float* aug2, int K, int B, int J,
                     double* legal, double* intDist1, double* intDist2, double* extDist,
                     double* w1, double* w2, double* coord, float* coord32, int32_t* zero_div,
                     void* stream);

/* Prediction error and PCK flag against ground truth (BusinessUtils._check_predsQuality,
 * utils/business.py:37-40; EvaluationUtils.acc_pck_pseudo(_norm), utils/evaluation.py:78-89) for
 * n_sets prediction sets pred [n_sets,B,J,2] float64 against gt [B,J,gt_stride] float32 (x, y first):
 * err = dist(pred, gt), acc = err / dist(gt[b,ref0], gt[b,ref1]) < pck_thr. */
int ubpl_coord_error(const double* pred, const float* gt, int gt_stride, int64_t n_sets, int B, int J,
                     int ref0, int ref1, double pck_thr, double* err, int32_t* acc, void* stream);

/* ProcessUtils.coord_distance (utils/process.py:53-54) for n coordinate pairs, float64 [n,2] each. */
int ubpl_pair_distance(const double* c1, const double* c2, int64_t n, double* out, void* stream);

/* ---- K2: pseudo-label selection ---------------------------------------------------------------
 * BusinessUtils.filter_pseudo2 (utils/business.py:173-217) / _calReliabilityThr (:43-46).
 * Step 1: local extrema of dist (float64[n]; 999 = sentinel): ext[0] = max over dist<999 (0 if
 * none), ext[1] = min over all.  On several GPUs the caller all-reduces ext (MAX / MIN). */
int ubpl_dist_extrema(const double* dist, int64_t n, double* ext, void* stream);
/* Step 2: reliability[n] (float64) from dist, legal and the GLOBAL extrema (ext as above, still
 * raw: the dist_max==0 -> 999 and reliableDistMin clamps of business.py:181-182 are applied
 * inside), plus the monotone uint64 sort key of every reliability. */
int ubpl_reliability(const double* dist, const double* legal, int64_t n, const double* ext,
                     double reliableDistMin, double* reliability, uint64_t* keys, void* stream);
/* Step 3: exact k-th order statistic by radix select, 16 bits per pass.  hist[65536] uint32 of
 * key bits [shift, shift+16) over the keys whose bits above shift+16 equal `prefix`'s (pass 0:
 * shift = 48, all keys).  The caller all-reduces hist over ranks (SUM) between the passes. */
int ubpl_key_histogram(const uint64_t* keys, int64_t n, const uint64_t* prefix, int shift,
                       uint32_t* hist, int clear_first, void* stream);
/* Given the (global) histogram of a pass, descend: finds the bin holding rank *k_rem (0-based,
 * counted from the LARGEST key), updates *prefix |= bin << shift and *k_rem; zero_after != 0 clears the
 * histogram afterwards (then the next ubpl_key_histogram can pass clear_first = 0). */
int ubpl_select_descend(uint32_t* hist, int shift, uint64_t* prefix, int64_t* k_rem, int zero_after, void* stream);
/* Step 4: thr = max(reliableThr, value(prefix)); enable[n] = reliability > thr (uint8; gate32 is
 * the same as float32 0/1, either may be NULL);
 * counts[J+1] int32 per-joint and total selected (item i is joint i % J); thr_out float64. */
int ubpl_select_apply(const double* reliability, int64_t n, int J, const uint64_t* prefix,
                      double reliableThr, uint8_t* enable, float* gate32, int32_t* counts,
                      double* thr_out, void* stream);
/* The whole of steps 1-4 in one launch for the single-GPU case (no collective needed): k_rank =
 * int((n-1)*reliablePCT) counted from the largest reliability; keys is uint64[n] scratch; ext_out
 * (float64[2], optional) receives the raw extrema. */
int ubpl_select_quantile_local(const double* dist, const double* legal, int64_t n, int J, int64_t k_rank,
                               double reliableThr, double reliableDistMin, double* reliability,
                               uint64_t* keys, uint8_t* enable, float* gate32, int32_t* counts,
                               double* thr_out, double* ext_out, void* stream);
/* Multi-GPU form of steps 1-4: the extrema and the four key histograms are all-reduced with NCCL over
 * the communicator created by ubpl_nccl_init (one process per GPU); kernels and collectives are enqueued
 * on `stream` by this single call (capturable in a CUDA graph).  n = this rank's items, k_rank =
 * int((n_total-1)*reliablePCT) over ALL ranks' items; ws = device scratch of UBPL_SELECT_WS_BYTES. */
#define UBPL_SELECT_WS_BYTES (2 * 8 + 8 + 8 + 65536 * 4)
int ubpl_select_quantile_dist(const double* dist, const double* legal, int64_t n, int J, int64_t k_rank,
                              double reliableThr, double reliableDistMin, double* reliability,
                              uint64_t* keys, uint8_t* enable, float* gate32, int32_t* counts,
                              double* thr_out, void* ws, void* stream);
/* ---- a13: mixed-distance uncertainty of two teachers (BusinessUtils.pseudo_cal_unc / pseudo_filter_mixUnc(2),
 * utils/business.py:220-294, 302-346, 378-406) ----------------------------------------------------------------
 * ubpl_mix_dists, one thread per key point, per teacher m: err_m = dist(pred_m, gt) and the PCK flag acc_m
 * (gt [B,J,gt_stride] float32 or NULL to skip), score_m = clamp01(scores_m[0][j]) (the reference reads batch row
 * 0; scores [B,J] float32 or NULL), caug_m [B,J,2] = python-float mean of the A augmented views a_m [B,J,A,2],
 * int_m = mean pairwise distance of the views; shared ext = dist(pred_1, pred_2), aext = dist(caug_1, caug_2).
 * All outputs float64 [B,J].  Integer radicands reproduce CPython's pow bit for bit; aext (and err for a
 * fractional gt) is the IEEE sqrt, <= 1 ulp from CPython. */
int ubpl_mix_dists(const float* gt, int gt_stride, int ref0, int ref1, double pck_thr,
                   const float* p1, const float* p2, const float* s1, const float* s2,
                   const float* a1, const float* a2, int B, int J, int A,
                   double* err1, double* err2, int32_t* acc1, int32_t* acc2, double* score1, double* score2,
                   double* caug1, double* caug2, double* int1, double* int2, double* ext, double* aext,
                   void* stream);
/* ubpl_mix_unc: the stateful half for ONE teacher.  hist float64 [3][n][3] + hist_len int32 [n] are the device
 * form of args.mdsN_lma_cache (zero-initialised, updated in place): the call pushes (intDist, extDist,
 * aExtDist), forms the 0.5/0.3/0.2 moving averages (lma_out [3][n], optional), mixDist (mix_out), unc =
 * 1-exp(-mixDist/5) or 999 when an average exceeds distThrMax (unc_out), and the fixed rule enable = unc <=
 * 1-exp(-3*distThrMax/5) with per-joint counts [J+1]; score/score_thr (optional) apply the median-score gate of
 * pseudo_filter_mixUnc2 first. */
int ubpl_mix_unc(const double* intDist, const double* extDist, const double* aExtDist, int64_t n, int J,
                 double distThrMax, double* hist, int32_t* hist_len, const double* score, const double* score_thr,
                 double* lma_out, double* mix_out, double* unc_out, uint8_t* enable, float* gate32,
                 int32_t* counts, void* stream);

/* The selector as ONE kernel per GPU (single CTA), single- and multi-GPU:
 *   use_p2p = 0: this GPU's n items are the whole population (k_rank < n; keys = uint64[n] scratch);
 *   use_p2p = 1: the all-reduce of the uncertainty histogram BASELINE.json names, done over NVLink peer memory
 *                inside the one kernel: every rank keeps its own distance keys and, per radix pass (12 bits, then
 *                11 at a time), stores the digit histogram of its keys (<= 16 KB) into its slot of every peer's
 *                exchange buffer (ubpl_p2p_alloc / ubpl_p2p_open), publishes a flag (st.release.sys), waits for
 *                the peers' flags and adds the R histograms found in its own buffer; once the bin that holds the
 *                global rank has <= 64 keys the ranks exchange those keys instead of another histogram.  O(n) work
 *                per rank for any number of ranks, typically three NVLink round trips; no NCCL call, no host
 *                work, capturable in a CUDA graph; k_rank = int((n_total-1)*reliablePCT) over all ranks.
 *                Every rank must make the same sequence of calls (it is a collective).  n above 16384 items per
 *                rank needs the keys scratch.
 * legal: float64[n] or uint8[n] (exactly one non-NULL).  Outputs as ubpl_select_quantile_local.
 * kps != NULL (float32 [n,2], image space) folds ubpl_gate_prepare into the launch: gate32 = enable *
 * visibility, *count_out = S * #(gate32 > 0), *grad_scale = loss_weight / count. */
int ubpl_select_quantile_fused(const double* dist, const double* legal_f64, const uint8_t* legal_u8, int64_t n,
                               int J, int64_t k_rank, double reliableThr, double reliableDistMin,
                               double* reliability, uint64_t* keys, uint8_t* enable, float* gate32,
                               int32_t* counts, double* thr_out, double* ext_out,
                               const float* kps, int img_h, int img_w, float stride, float sigma, int S,
                               float loss_weight, float* grad_scale, int32_t* count_out,
                               int use_p2p, void* stream);

/* Test vehicle of the multi-GPU selector's exchange protocol: R ranks emulated by the R blocks of ONE cooperative
 * launch on one GPU (for boxes with fewer GPUs than ranks).  Rank r owns items [r*n_stride, r*n_stride + n_per_rank[r])
 * of dist / legal_u8 / keys / reliability / enable / gate32, row r of counts [R, J+1] and thr_out[r]; n_per_rank is a
 * HOST array; xbuf holds R zeroed exchange buffers of xbuf_stride >= ubpl_p2p_buffer_bytes(R, cap) bytes each. */
int ubpl_select_quantile_emul(const double* dist, const uint8_t* legal_u8, int R, const int64_t* n_per_rank,
                              int64_t n_stride, int J, int64_t k_rank, double reliableThr, double reliableDistMin,
                              double* reliability, uint64_t* keys, uint8_t* enable, float* gate32, int32_t* counts,
                              double* thr_out, void* xbuf, int64_t xbuf_stride, int64_t cap, void* stream);
/* Developer hook: 64 device uint64 that receive globaltimer stamps of the selector's phases ([63] = count);
 * NULL switches it off. */
int ubpl_select_debug_stamps(void* dev_u64x64);
/* Peer-memory exchange buffer of the call above, one per process (= per GPU).  ubpl_p2p_alloc allocates it for
 * `nranks` ranks of at most `max_items` items each and returns its 64-byte CUDA-IPC handle (host buffer); the
 * handles of all ranks, concatenated in rank order (any transport, e.g. torch.distributed all_gather), go to
 * ubpl_p2p_open, which maps the peers' buffers.  ubpl_p2p_status: 0, or 1 after a launch in which a peer did
 * not arrive within the time-out (UBPL_P2P_TIMEOUT_MS, read at ubpl_p2p_alloc, default 60 s); the outputs of
 * that launch are NaN-poisoned (threshold NaN, every mask 0) instead of hanging -- callers must check the status
 * at a point where a sync is acceptable (dist.check_p2p) and raise. */
int64_t ubpl_p2p_buffer_bytes(int nranks, int64_t max_items);
int ubpl_p2p_alloc(int nranks, int64_t max_items, void* handle_out64_host);
int ubpl_p2p_open(const void* handles_host, int nranks, int rank);
int ubpl_p2p_close(void);
int ubpl_p2p_ranks(void);
int ubpl_p2p_status(void);
/* NCCL plumbing for ubpl_select_quantile_dist (libnccl.so.2 is resolved with dlopen): rank 0 obtains a 128-byte
 * unique id (host buffer), shares it by any means (torch.distributed broadcast), every rank calls init. */
int ubpl_nccl_unique_id(void* out128_host);
int ubpl_nccl_init(const void* id128_host, int nranks, int rank);
int ubpl_nccl_destroy(void);
int ubpl_nccl_ranks(void);   /* 0 when no communicator exists */

/* Fixed rule: enable = legal && 1-exp(-dist/5) <= 1-exp(-3*distThrMax/5)
 * (BusinessUtils.pseudo_filter_mixUnc / _calUncValue, utils/business.py:237-261,375-376). */
int ubpl_select_fixed(const double* dist, const double* legal, int64_t n, int J, double distThrMax,
                      uint8_t* enable, float* gate32, int32_t* counts, double* unc_out, void* stream);

/* Fused K2 of the mean-teacher fixed-threshold path, one launch: ubpl_view_dispersion (with the 999
 * sentinel) + ubpl_select_fixed + ubpl_gate_prepare.  preds [K,B,J,2]; outputs as in those three
 * (out_mean is the pseudo key point the targets are rendered at); out_mean, out_dist, out_legal and
 * enable may be NULL.  counts is int32[J+2]: per-joint and total selected, then count_out =
 * &counts[J+1] = S * #(gate_out > 0), which ubpl_render_mse turns into the gradient scale. */
int ubpl_k2_view_fixed(const float* preds, int K, int B, int J, double distThrMax, int img_h, int img_w,
                       float stride, float sigma, int S, float* out_mean, double* out_dist,
                       uint8_t* out_legal, uint8_t* enable, float* gate_out, int32_t* count_out,
                       int32_t* counts, void* stream);

/* ---- K3: Gaussian target render + masked joint-MSE, forward and gradient in one pass ----------
 * ProcessUtils.kps_heatmap (utils/process.py:253-278,394-397) fused into JointMSELoss
 * (utils/losses.py:8-29).  kps [B,J,2] float32 image-space coordinates; gate_in [B,J] float32
 * (key-point weight / enable), may be NULL (=1); sample_w [B] float32 or NULL.
 * pred [B,S,J,H,W] strides (pB,pS,pJ); grad same shape, strides (gB,gS,gJ), may be NULL;
 * target [B,J,H,W] contiguous, may be NULL (not materialised).
 * img_h/img_w: input resolution (256); stride = inpRes/outRes; sigma = kernelSize*sigma (3).
 * gate_out [B,J] float32 = gate_in * visibility (process.py:267-268).
 * per_loss [B,S,J] float32 = mean_HW (p-t)^2 * gate_out * sample_w.
 * grad = gs * 2/(HW) * gate_out * sample_w * (p - t) where gs is read from device memory so the
 * caller can fold weight/n (MT_UBPL.py:266) without a host sync: gs = *grad_scale (NULL = 1), or,
 * when count_in != NULL, gs = loss_weight / *count_in (loss_weight if the count is 0), which is
 * also stored to *grad_scale_out. */
int ubpl_render_mse(const float* kps, const float* gate_in, const float* sample_w,
                    const float* pred, int64_t pB, int64_t pS, int64_t pJ,
                    float* grad, int64_t gB, int64_t gS, int64_t gJ,
                    float* target, int B, int S, int J, int H, int W,
                    int img_h, int img_w, float stride, float sigma,
                    const float* grad_scale, const int32_t* count_in, float loss_weight,
                    float* grad_scale_out, float* gate_out, float* per_loss, void* stream);
/* ubpl_render_mse with the loss reduction fused into the same launch: every CTA leaves its partial sums in
 * the workspace and the CTA that finishes last adds them in CTA order and writes summary float64[4] =
 * (sum(per_loss), #(per_loss > 0), B*S*J, #(gate_out > 0)), i.e. what ubpl_loss_finalize(per_loss, NULL,
 * gate_out) returns.  sum_ws: UBPL_RENDER_SUM_WS_BYTES of device memory, 16-byte aligned, whose first word is
 * zero before the first launch; the kernel returns it to zero (launches sharing a workspace must not overlap). */
#define UBPL_RENDER_SUM_WS_BYTES (16 + 24 * 4096)
int ubpl_render_mse_sum(const float* kps, const float* gate_in, const float* sample_w,
                        const float* pred, int64_t pB, int64_t pS, int64_t pJ,
                        float* grad, int64_t gB, int64_t gS, int64_t gJ,
                        float* target, int B, int S, int J, int H, int W,
                        int img_h, int img_w, float stride, float sigma,
                        const float* grad_scale, const int32_t* count_in, float loss_weight,
                        float* grad_scale_out, float* gate_out, float* per_loss,
                        double* summary, void* sum_ws, void* stream);
/* kps_heatmap alone (utils/process.py:253-278): kps [N,3] float32 (x,y,w) -> heatmap [N,H,W],
 * kps_out [N,3] with w *= visibility. */
int ubpl_render_targets(const float* kps, int N, int H, int W, int img_h, int img_w, float stride,
                        float sigma, float* heatmap, float* kps_out, void* stream);

/* Dense-target masked joint-MSE forward + gradient: JointMSELoss / JointDistLoss
 * (utils/losses.py:8-53), JointPseudoLoss3 (:169-210), JointDistLoss_mt2 (:246-286).
 * pred [B,S,J,H,W] strides (pB,pS,pJ).  tgt: M maps per (b,s,j), averaged in float32 in index
 * order then divided by M (torch.mean): strides (tM,tB,tS,tJ); tS = 0 shares one target between
 * the stacks.  coef [B,J] float32 = gate*sample weight (NULL = 1).
 * mask_mode 0: none; 1: (max p >= thr) && (max t >= thr)  (losses.py:187-193);
 *           2: (max t >= thr)  (losses.py:270-271).
 * Outputs [B,S,J] float32: per_loss = mean_HW (p-t)^2 * coef (before the mask, as the
 * reference counts `loss > 0` on it), mask, vmax_p, vmax_t (any may be NULL).
 * grad (may be NULL) = grad_scale * 2/(HW) * coef * mask * (p - t). */
int ubpl_dense_mse(const float* pred, int64_t pB, int64_t pS, int64_t pJ,
                   const float* tgt, int M, int64_t tM, int64_t tB, int64_t tS, int64_t tJ,
                   const float* coef, int mask_mode, float thr,
                   float* grad, int64_t gB, int64_t gS, int64_t gJ,
                   int B, int S, int J, int H, int W, const float* grad_scale,
                   float* per_loss, float* mask, float* vmax_p, float* vmax_t, void* stream);
/* Reduces the [B,S,J] planes: out[0] = sum(per_loss*mask) (mask NULL = 1), out[1] = #(per_loss>0),
 * out[2] = #(mask>0), out[3] = #(gate>0) over gate[B,J] (NULL -> B*J); all as float64[4]. */
int ubpl_loss_finalize(const float* per_loss, const float* mask, const float* gate,
                       int B, int S, int J, double* out, void* stream);
/* gate_out[n] = gate_in * visibility(kps) (utils/process.py:262-268; gate_in NULL = 1), count_out =
 * S * #(gate_out > 0) (the `n` JointMSELoss returns, utils/losses.py:29) and *grad_scale =
 * loss_weight / count (loss_weight if count == 0), the factor projects/MT_UBPL.py:266 applies to
 * the criterion's sum -- computed on device so ubpl_render_mse can write the final gradient. */
int ubpl_gate_prepare(const float* kps, const float* gate_in, int64_t n, int img_h, int img_w, float stride,
                      float sigma, int S, float loss_weight, float* gate_out, float* grad_scale,
                      int32_t* count_out, void* stream);
/* dst[0..n) = src[0..n) * *scale (device scalar; dst may alias src) -- the backward of the autograd
 * wrappers; both 16-byte aligned. */
int ubpl_scale(float* dst, const float* src, int64_t n, const float* scale, void* stream);

/* ---- N1: canonical key points -> the frame of every augmented view --------------------------
 * ProcessUtils.kps_fliplr (utils/process.py:239-242) + AugmentUtils.affine_kps (utils/augment.py:151-156) /
 * transform (utils/udaap/transforms.py:151-158) for all views, samples and joints in one launch.
 * kps [B,J,3] float32 (x, y, weight); mats [V,B,2,3] float64 = rows 0,1 of get_transform(centre, scale, res, rot)
 * of every (view, sample) (host-built: its sin/cos are numpy's); flips [V,B] uint8 or NULL; out [V,B,J,3].
 * The result feeds ubpl_render_targets / ubpl_render_mse to render the targets in each student view's frame. */
int ubpl_view_kps(const float* kps, const double* mats, const uint8_t* flips, float img_w, int V, int B, int J,
                  float* out, void* stream);

/* ---- N2 / N3: PCK evaluation and the feature-decorrelation loss ---------------------------
 * ubpl_acc_pck: EvaluationUtils.acc_pck (utils/evaluation.py:92-139) on device.  preds [bs,k,p_stride>=2],
 * gts [bs,k,g_stride>=2] float32; errs/accs float32 [k+1] (per joint, then the mean over joints; accs[k] = -1
 * for a joint without a visible gt, which the mean skips); dists / dists_ref [k,bs] optional. */
int ubpl_acc_pck(const float* preds, int p_stride, const float* gts, int g_stride, int bs, int k,
                 int ref0, int ref1, float pck_thr, float* errs, float* accs, float* dists, float* dists_ref,
                 void* stream);
/* ubpl_features_cov: ProcessUtils.features_cov (utils/process.py:19-31) forward + gradient in one pass.
 * f1, f2 [rows, L] float32 contiguous (rows = bs*n*c, L = h*w); cov [rows] = off-diagonal covariance of the two
 * rows; *value = mean |cov|; g1, g2 [rows, L] (both or neither) = d value / d f1, d f2. */
int ubpl_features_cov(const float* f1, const float* f2, int64_t rows, int L, float* cov, float* value,
                      float* g1, float* g2, void* stream);

/* ---- K4: mean-teacher EMA, all parameter tensors in one launch ---------------------------------
 * update_ema_variables (utils/parameters.py:4-8): ema = ema*alpha + (1-alpha)*param, float32,
 * evaluated as fma(param, 1-alpha, ema*alpha) like ATen.  ema_ptrs/param_ptrs: device arrays of
 * n_tensors device addresses; chunk_tensor/chunk_start: device arrays describing n_chunks work
 * items (tensor id, first element); chunk_elems elements per chunk; numels[n_tensors].
 * alpha_dev (optional, device float[2] = {alpha, 1 - alpha}): when not NULL the kernel reads the two factors
 * from it instead of the by-value arguments, so a CUDA graph that captured the launch follows the per-epoch
 * alpha = min(1 - 1/(epo+1), ema_decay) of the reference (the host rewrites the two floats between replays). */
int ubpl_ema_multi_tensor(const uint64_t* ema_ptrs, const uint64_t* param_ptrs, const int64_t* numels,
                          const int32_t* chunk_tensor, const int64_t* chunk_start, int64_t n_chunks,
                          int chunk_elems, float alpha, float one_minus_alpha, const float* alpha_dev,
                          void* stream);
/* Contiguous special case (flattened parameter buffer). */
int ubpl_ema_flat(float* ema, const float* param, int64_t n, float alpha, float one_minus_alpha,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UBPL_B200_H */
