// N1, N2 and N3 of SURVEY.md section 8(f): key points into the frame of every augmented view (for in-frame
// target rendering), the PCK evaluation of validate() and the multi-view feature decorrelation loss.
//
// Reference semantics (file:line in /root/reference):
//   utils/udaap/transforms.py:119-158, utils/augment.py:151-156, utils/process.py:239-242  (N1)
//   utils/evaluation.py:92-139  acc_pck / _acc_calDists / _acc_counting (float32 torch tensors)
//   utils/process.py:19-31      features_cov / torch_cov
#include "common.cuh"

namespace ubpl {

// ---- N2 ------------------------------------------------------------------------------------------------
// One thread per joint walks the batch in order (float32, like the reference's tensors): dist = ||pred - gt||
// where gt_x > 1 and gt_y > 1, else -1; errs[k] = sum over ALL b (the -1 entries included, evaluation.py:101)
// / bs; accs[k] = #(valid, dist/norm_b < thr) / #valid or -1.  Thread 0 then forms the means over the joints.
__global__ void __launch_bounds__(1024) acc_pck_kernel(const float* __restrict__ preds, int p_stride,
                                                        const float* __restrict__ gts, int g_stride, int bs, int k,
                                                        int ref0, int ref1, float thr, float* errs, float* accs,
                                                        float* dists_out, float* dref_out) {
  extern __shared__ float sh[];           // errs[k], accs[k]
  for (int kk = threadIdx.x; kk < k; kk += blockDim.x) {
    float sum = 0.f;
    int valid = 0, hit = 0;
    for (int i = 0; i < bs; ++i) {
      const float* g = gts + ((long long)i * k + kk) * g_stride;
      const float* p = preds + ((long long)i * k + kk) * p_stride;
      float d = -1.f, dr = -1.f;
      if (g[0] > 1.f && g[1] > 1.f) {
        const float* a = gts + ((long long)i * k + ref0) * g_stride;
        const float* b = gts + ((long long)i * k + ref1) * g_stride;
        const float nx = __fsub_rn(a[0], b[0]), ny = __fsub_rn(a[1], b[1]);
        const float norm = __fsqrt_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)));
        const float dx = __fsub_rn(p[0], g[0]), dy = __fsub_rn(p[1], g[1]);
        d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        dr = __fdiv_rn(d, norm);
        ++valid;
        hit += (dr < thr) ? 1 : 0;
      }
      sum = __fadd_rn(sum, d);
      if (dists_out) dists_out[(long long)kk * bs + i] = d;
      if (dref_out) dref_out[(long long)kk * bs + i] = dr;
    }
    sh[kk] = __fdiv_rn(sum, (float)bs);
    sh[k + kk] = valid > 0 ? (float)(1.0 * (double)hit / (double)valid) : -1.f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float es = 0.f, as = 0.f;
    int an = 0;
    for (int kk = 0; kk < k; ++kk) {
      errs[kk] = sh[kk];
      accs[kk] = sh[k + kk];
      es = __fadd_rn(es, sh[kk]);
      if (sh[k + kk] >= 0.f) { as = __fadd_rn(as, sh[k + kk]); ++an; }
    }
    errs[k] = __fdiv_rn(es, (float)k);
    accs[k] = an ? __fdiv_rn(as, (float)an) : 0.f;
  }
}

// ---- N3 ------------------------------------------------------------------------------------------------
// One warp per (b, n, c) row of L = h*w features of the two inputs: means, the off-diagonal covariance
// cov = sum (x1-m1)(x2-m2) / (L-1), and -- when asked -- the gradient rows
//   g1 = coef * sign(cov) * (x2 - m2),  g2 = coef * sign(cov) * (x1 - m1),  coef = 1 / ((L-1) * rows)
// of value = mean_rows |cov| (the mean-subtraction terms cancel: centred rows sum to zero).  Rows of 1024
// floats stay in registers, so each input is read from HBM once and each gradient written once; longer
// rows are re-read (from L1/L2).
template <int PER>   // float4 per lane held in registers (0: stream from memory every pass)
__global__ void __launch_bounds__(256) features_cov_kernel(const float* __restrict__ f1, const float* __restrict__ f2,
                                                           long long rows, int L, float* __restrict__ cov_out,
                                                           float* __restrict__ g1, float* __restrict__ g2, float coef) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float invL = 1.f / (float)L;
  for (long long r = warp; r < rows; r += nwarps) {
    const float* a = f1 + r * L;
    const float* b = f2 + r * L;
    float s1 = 0.f, s2 = 0.f, c = 0.f;
    if (PER > 0) {
      const float4* a4 = reinterpret_cast<const float4*>(a);
      const float4* b4 = reinterpret_cast<const float4*>(b);
      float4 x[PER > 0 ? PER : 1], y[PER > 0 ? PER : 1];
#pragma unroll
      for (int u = 0; u < PER; ++u) { x[u] = ldg_stream(a4 + lane + 32 * u); y[u] = ldg_stream(b4 + lane + 32 * u); }
#pragma unroll
      for (int u = 0; u < PER; ++u) { s1 += (x[u].x + x[u].y) + (x[u].z + x[u].w); s2 += (y[u].x + y[u].y) + (y[u].z + y[u].w); }
      const float m1 = warp_sum(s1) * invL, m2 = warp_sum(s2) * invL;
#pragma unroll
      for (int u = 0; u < PER; ++u) {
        x[u].x -= m1; x[u].y -= m1; x[u].z -= m1; x[u].w -= m1;
        y[u].x -= m2; y[u].y -= m2; y[u].z -= m2; y[u].w -= m2;
        c += (x[u].x * y[u].x + x[u].y * y[u].y) + (x[u].z * y[u].z + x[u].w * y[u].w);
      }
      const float cov = warp_sum(c) / (float)(L - 1);
      if (lane == 0) cov_out[r] = cov;
      if (g1) {
        const float k = coef * ((cov > 0.f) ? 1.f : ((cov < 0.f) ? -1.f : 0.f));
        float4* o1 = reinterpret_cast<float4*>(g1 + r * L);
        float4* o2 = reinterpret_cast<float4*>(g2 + r * L);
#pragma unroll
        for (int u = 0; u < PER; ++u) {
          stg_stream(o1 + lane + 32 * u, make_float4(k * y[u].x, k * y[u].y, k * y[u].z, k * y[u].w));
          stg_stream(o2 + lane + 32 * u, make_float4(k * x[u].x, k * x[u].y, k * x[u].z, k * x[u].w));
        }
      }
    } else {
      for (int i = lane; i < L; i += 32) { s1 += __ldg(a + i); s2 += __ldg(b + i); }
      const float m1 = warp_sum(s1) * invL, m2 = warp_sum(s2) * invL;
      for (int i = lane; i < L; i += 32) c += (__ldg(a + i) - m1) * (__ldg(b + i) - m2);
      const float cov = warp_sum(c) / (float)(L - 1);
      if (lane == 0) cov_out[r] = cov;
      if (g1) {
        const float k = coef * ((cov > 0.f) ? 1.f : ((cov < 0.f) ? -1.f : 0.f));
        for (int i = lane; i < L; i += 32) {
          g1[r * L + i] = k * (__ldg(b + i) - m2);
          g2[r * L + i] = k * (__ldg(a + i) - m1);
        }
      }
    }
  }
}

// ---- N1 ------------------------------------------------------------------------------------------------
// Canonical key points -> the frame of every augmented view (what the Dataset does per sample and joint on the
// host): x <- img_w - x for a flipped view (process.py:239-242, float32), then for the visible key points
// (y > 0, augment.py:154) transform() of transforms.py:151-158: v = (x-1, y-1, 1) with the subtraction in
// float32, np.dot(t, v) -- measured on numpy: fma(t00, vx, t01*vy) + t02 in float64 -- truncation, + 1.
// mats [V*B, 2, 3] float64 are rows 0 and 1 of get_transform for the view (built on the host from the view's
// centre / scale / angle, whose sin/cos must come from the host's numpy to stay bit-identical).
__global__ void view_kps_kernel(const float* __restrict__ kps, const double* __restrict__ mats,
                                const uint8_t* __restrict__ flips, float img_w, int V, int B, int J,
                                float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = (long long)V * B * J;
  if (i >= n) return;
  const int j = (int)(i % J);
  const long long vb = i / J;
  const int b = (int)(vb % B);
  const float* k = kps + ((long long)b * J + j) * 3;
  float x = k[0], y = k[1];
  if (flips && flips[vb]) x = __fsub_rn(img_w, x);
  if (y > 0.f) {
    const double* t = mats + vb * 6;
    const double vx = (double)__fsub_rn(x, 1.f), vy = (double)__fsub_rn(y, 1.f);
    const double rx = __dadd_rn(__fma_rn(t[0], vx, __dmul_rn(t[1], vy)), t[2]);
    const double ry = __dadd_rn(__fma_rn(t[3], vx, __dmul_rn(t[4], vy)), t[5]);
    x = (float)(trunc(rx) + 1.0);
    y = (float)(trunc(ry) + 1.0);
  }
  out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = k[2];
}

// mean of |x| over n values, one CTA, fixed order (reproducible); out[0] = mean
__global__ void __launch_bounds__(1024) abs_mean_kernel(const float* __restrict__ x, long long n, float* out) {
  __shared__ double red[32];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) s += (double)fabsf(x[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    out[0] = (float)(t / (double)n);
  }
}

}  // namespace ubpl

using namespace ubpl;

extern "C" int ubpl_acc_pck(const float* preds, int p_stride, const float* gts, int g_stride, int bs, int k, int ref0,
                            int ref1, float pck_thr, float* errs, float* accs, float* dists, float* dists_ref,
                            void* stream) {
  UBPL_REQUIRE(preds && gts && errs && accs, "ubpl_acc_pck: NULL pointer");
  UBPL_REQUIRE(bs >= 1 && k >= 1 && k <= 4096 && p_stride >= 2 && g_stride >= 2, "ubpl_acc_pck: bad dims (1 <= k <= 4096)");
  UBPL_REQUIRE(ref0 >= 0 && ref0 < k && ref1 >= 0 && ref1 < k, "ubpl_acc_pck: pck_ref out of range");
  const int threads = k < 1024 ? ((k + 31) / 32) * 32 : 1024;
  acc_pck_kernel<<<1, threads, (size_t)2 * k * sizeof(float), (cudaStream_t)stream>>>(preds, p_stride, gts, g_stride, bs, k, ref0,
                                                                                     ref1, pck_thr, errs, accs, dists, dists_ref);
  return check_launch("ubpl_acc_pck");
}

extern "C" int ubpl_view_kps(const float* kps, const double* mats, const uint8_t* flips, float img_w, int V, int B, int J,
                             float* out, void* stream) {
  UBPL_REQUIRE(V >= 0 && B >= 0 && J >= 0, "ubpl_view_kps: bad dims");
  const long long n = (long long)V * B * J;
  if (n == 0) return UBPL_OK;
  UBPL_REQUIRE(kps && mats && out, "ubpl_view_kps: NULL pointer");
  view_kps_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(kps, mats, flips, img_w, V, B, J, out);
  return check_launch("ubpl_view_kps");
}

extern "C" int ubpl_features_cov(const float* f1, const float* f2, int64_t rows, int L, float* cov, float* value,
                                 float* g1, float* g2, void* stream) {
  UBPL_REQUIRE(f1 && f2 && cov && value && rows >= 1 && L >= 2, "ubpl_features_cov: bad arguments");
  UBPL_REQUIRE((g1 == nullptr) == (g2 == nullptr), "ubpl_features_cov: give both gradients or none");
  const float coef = 1.f / ((float)(L - 1) * (float)rows);
  const bool vec = (L % 128 == 0) && ((reinterpret_cast<uintptr_t>(f1) | reinterpret_cast<uintptr_t>(f2) |
                                       reinterpret_cast<uintptr_t>(g1) | reinterpret_cast<uintptr_t>(g2)) & 15) == 0;
  const int wpb = 8;
  long long blocks = (rows + wpb - 1) / wpb;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (vec && L == 1024) features_cov_kernel<8><<<(int)blocks, wpb * 32, 0, st>>>(f1, f2, rows, L, cov, g1, g2, coef);
  else if (vec && L == 512) features_cov_kernel<4><<<(int)blocks, wpb * 32, 0, st>>>(f1, f2, rows, L, cov, g1, g2, coef);
  else if (vec && L == 256) features_cov_kernel<2><<<(int)blocks, wpb * 32, 0, st>>>(f1, f2, rows, L, cov, g1, g2, coef);
  else features_cov_kernel<0><<<(int)blocks, wpb * 32, 0, st>>>(f1, f2, rows, L, cov, g1, g2, coef);
  int rc = check_launch("ubpl_features_cov");
  if (rc != UBPL_OK) return rc;
  abs_mean_kernel<<<1, 1024, 0, st>>>(cov, rows, value);
  return check_launch("ubpl_features_cov(mean)");
}
