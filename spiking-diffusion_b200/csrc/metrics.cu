// (f4) Quality metrics on the GPU: the algebra of the reference's evaluation block that needs no pretrained weights.
//
//   sd_metric_mse / sd_metric_ssim   reconstruction test, R/main.py:303-323: F.mse_loss and
//                                    metric.pytorch_ssim.SSIM(window_size=11) (R/metric/pytorch_ssim/__init__.py:17-65)
//   sd_metric_feature_stats          mu = mean(act), sigma = np.cov(act, rowvar=False)   (R/metric/Fid_score.py:100-113)
//   sd_metric_frechet                ||mu1-mu2||^2 + tr(s1) + tr(s2) - 2 tr(sqrtm(s1 s2)) with the reference's own
//                                    "sqrtm": U diag(sqrt(S)) Vh of the SVD of s1 s2   (R/metric/Fid_score.py:14-17,116-173)
//   sd_metric_poly_mmd2              kernel-inception-distance estimator on given features: unbiased MMD^2 with the
//                                    polynomial kernel (x.y / d + 1)^3 (torchmetrics KernelInceptionDistance, used at
//                                    R/main.py:465-490; torchmetrics is a third-party dependency absent from the reference
//                                    tree and unpinned there -- its published estimator is restated)
//   sd_metric_inception_score        exp(mean_i KL(p_i || mean p)) per split, mean / std over splits
//                                    (R/metric/IS_score.py:58-72, scipy.stats.entropy semantics)
//
// The Inception feature extractor itself needs pretrained weights that are not available offline; these entry points
// take its outputs (features / class probabilities).  All reductions are two-stage with a fixed order (deterministic);
// statistics and the Frechet distance are fp64 like the numpy reference.
#include <math.h>
#include "common.cuh"

namespace sd {

constexpr int kRedThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double r = red[0];
  __syncthreads();
  return r;
}

// out[0] = scale * sum of `n` partials, one block, fixed order
__global__ void __launch_bounds__(kRedThreads) final_sum_kernel(const double* __restrict__ part, int64_t n, double scale,
                                                                double* __restrict__ out64, float* __restrict__ out32) {
  __shared__ double red[kRedThreads];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += kRedThreads) acc += part[i];
  const double r = block_sum(acc, red) * scale;
  if (threadIdx.x == 0) {
    if (out64) *out64 = r;
    if (out32) *out32 = (float)r;
  }
}

// ---- MSE --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRedThreads) mse_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  int64_t n, double* __restrict__ part) {
  __shared__ double red[kRedThreads];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = __fsub_rn(a[i], b[i]);
    acc += (double)__fmul_rn(d, d);          // each squared difference in fp32 as torch does, the sum in fp64
  }
  const double r = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// ---- SSIM -------------------------------------------------------------------------------------------------
struct SsimWindow { float w[15 * 15]; };

// One thread = one pixel of one (image, channel) plane; zero padding (F.conv2d(padding = ws // 2)).
__global__ void __launch_bounds__(kRedThreads) ssim_partial_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                                   int64_t planes, int H, int W, int ws, SsimWindow win,
                                                                   double* __restrict__ part, float* __restrict__ per_plane_sum) {
  __shared__ double red[kRedThreads];
  const int64_t total = planes * H * W;
  const int half = ws / 2;
  const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const int y = (int)((i / W) % H);
    const int64_t pl = i / ((int64_t)W * H);
    const float* p1 = img1 + pl * H * W;
    const float* p2 = img2 + pl * H * W;
    float mu1 = 0.f, mu2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int ky = 0; ky < ws; ++ky) {
      const int yy = y + ky - half;
      if (yy < 0 || yy >= H) continue;
      for (int kx = 0; kx < ws; ++kx) {
        const int xx = x + kx - half;
        if (xx < 0 || xx >= W) continue;
        const float wv = win.w[ky * ws + kx];
        const float a = p1[yy * W + xx], b = p2[yy * W + xx];
        mu1 = fmaf(wv, a, mu1);
        mu2 = fmaf(wv, b, mu2);
        s11 = fmaf(wv, __fmul_rn(a, a), s11);
        s22 = fmaf(wv, __fmul_rn(b, b), s22);
        s12 = fmaf(wv, __fmul_rn(a, b), s12);
      }
    }
    const float mu1_sq = __fmul_rn(mu1, mu1), mu2_sq = __fmul_rn(mu2, mu2), mu1_mu2 = __fmul_rn(mu1, mu2);
    const float sg1 = __fsub_rn(s11, mu1_sq), sg2 = __fsub_rn(s22, mu2_sq), sg12 = __fsub_rn(s12, mu1_mu2);
    const float num = __fmul_rn(__fadd_rn(__fmul_rn(2.f, mu1_mu2), C1), __fadd_rn(__fmul_rn(2.f, sg12), C2));
    const float den = __fmul_rn(__fadd_rn(__fadd_rn(mu1_sq, mu2_sq), C1), __fadd_rn(__fadd_rn(sg1, sg2), C2));
    const float v = __fdiv_rn(num, den);
    acc += (double)v;
    if (per_plane_sum) atomicAdd(per_plane_sum + pl, v);   // optional per-(image, channel) sums (size_average=False)
  }
  const double r = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// ---- feature statistics ---------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(kRedThreads) col_mean_kernel(const TIn* __restrict__ act, int64_t N, int d, double* __restrict__ mu) {
  __shared__ double red[kRedThreads];
  const int j = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < N; i += kRedThreads) acc += (double)act[i * d + j];
  const double r = block_sum(acc, red);
  if (threadIdx.x == 0) mu[j] = r / (double)N;
}

// sigma[a][b] = sum_i (x[i][a] - mu[a]) (x[i][b] - mu[b]) / (N - 1); 32 x 32 output tile per block, fp64
template <typename TIn>
__global__ void __launch_bounds__(256) cov_kernel(const TIn* __restrict__ act, const double* __restrict__ mu, int64_t N, int d,
                                                  double* __restrict__ sigma) {
  __shared__ double sa[32][33], sb[32][33];
  const int ta = blockIdx.y * 32, tb = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;     // 32 x 8 threads, each 4 outputs (rows ty, ty+8, ...)
  double acc[4] = {0, 0, 0, 0};
  for (int64_t i0 = 0; i0 < N; i0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      const int64_t i = i0 + r;
      sa[r][tx] = (i < N && ta + tx < d) ? (double)act[i * d + ta + tx] - mu[ta + tx] : 0.0;
      sb[r][tx] = (i < N && tb + tx < d) ? (double)act[i * d + tb + tx] - mu[tb + tx] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const double bv = sb[r][tx];
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[k] = fma(sa[r][ty + 8 * k], bv, acc[k]);
    }
    __syncthreads();
  }
  for (int k = 0; k < 4; ++k) {
    const int a = ta + ty + 8 * k, b = tb + tx;
    if (a < d && b < d) sigma[(int64_t)a * d + b] = acc[k] / (double)(N - 1);
  }
}

// C = A * B, all d x d row-major fp64 (32 x 32 tiles)
__global__ void __launch_bounds__(256) dgemm_nn_kernel(const double* __restrict__ A, const double* __restrict__ B, int d,
                                                       double* __restrict__ C) {
  __shared__ double sa[32][33], sb[32][33];
  const int ta = blockIdx.y * 32, tb = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      sa[r][tx] = (ta + r < d && k0 + tx < d) ? A[(int64_t)(ta + r) * d + k0 + tx] : 0.0;
      sb[r][tx] = (k0 + r < d && tb + tx < d) ? B[(int64_t)(k0 + r) * d + tb + tx] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double bv = sb[k][tx];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma(sa[ty + 8 * q][k], bv, acc[q]);
    }
    __syncthreads();
  }
  for (int q = 0; q < 4; ++q) {
    const int a = ta + ty + 8 * q, b = tb + tx;
    if (a < d && b < d) C[(int64_t)a * d + b] = acc[q];
  }
}

// ---- one-sided Jacobi SVD (Hestenes) -----------------------------------------------------------------------
// Bt and Vt hold the COLUMNS of B = A V and of V as contiguous rows ([dp][d], dp = d rounded up to even; a padding row
// is all zero).  One launch = one round of the round-robin tournament: dp/2 disjoint column pairs, one block each.
__global__ void __launch_bounds__(kRedThreads) jacobi_round_kernel(double* __restrict__ Bt, double* __restrict__ Vt, int d, int dp,
                                                                   int round, double tol, int* __restrict__ rotated) {
  __shared__ double red[kRedThreads];
  __shared__ double cs[2];
  // circle method: player dp-1 is fixed, the others rotate
  const int k = blockIdx.x, m = dp - 1;
  int p, q;
  if (k == 0) { p = m; q = round % m; }
  else { p = (round + k) % m; q = (round - k + m) % m; }
  if (p > q) { const int t = p; p = q; q = t; }
  double* bp = Bt + (int64_t)p * d;
  double* bq = Bt + (int64_t)q * d;
  double a = 0.0, b = 0.0, g = 0.0;
  for (int i = threadIdx.x; i < d; i += kRedThreads) {
    const double x = bp[i], y = bq[i];
    a = fma(x, x, a); b = fma(y, y, b); g = fma(x, y, g);
  }
  a = block_sum(a, red); b = block_sum(b, red); g = block_sum(g, red);
  if (threadIdx.x == 0) {
    double c = 1.0, s = 0.0;
    if (fabs(g) > tol * sqrt(a * b) && a > 0.0 && b > 0.0) {
      const double zeta = (b - a) / (2.0 * g);
      const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
      c = 1.0 / sqrt(1.0 + t * t);
      s = c * t;
      *rotated = 1;
    }
    cs[0] = c; cs[1] = s;
  }
  __syncthreads();
  const double c = cs[0], s = cs[1];
  if (s == 0.0) return;
  double* vp = Vt + (int64_t)p * d;
  double* vq = Vt + (int64_t)q * d;
  for (int i = threadIdx.x; i < d; i += kRedThreads) {
    const double x = bp[i], y = bq[i];
    bp[i] = c * x - s * y; bq[i] = s * x + c * y;
    const double u = vp[i], w = vq[i];
    vp[i] = c * u - s * w; vq[i] = s * u + c * w;
  }
}

__global__ void set_identity_kernel(double* __restrict__ Vt, int d, int dp) {
  const int64_t total = (int64_t)dp * d;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    Vt[i] = (i / d == i % d) ? 1.0 : 0.0;
}

// part[i] = sqrt(s_i) * (u_i . v_i) = (b_i . v_i) / sqrt(|b_i|)   for column i of B = A V (0 for a null column)
__global__ void __launch_bounds__(kRedThreads) sqrtm_trace_partial_kernel(const double* __restrict__ Bt, const double* __restrict__ Vt,
                                                                          int d, double* __restrict__ part) {
  __shared__ double red[kRedThreads];
  const double* b = Bt + (int64_t)blockIdx.x * d;
  const double* v = Vt + (int64_t)blockIdx.x * d;
  double nn = 0.0, bv = 0.0;
  for (int i = threadIdx.x; i < d; i += kRedThreads) { nn = fma(b[i], b[i], nn); bv = fma(b[i], v[i], bv); }
  nn = block_sum(nn, red); bv = block_sum(bv, red);
  if (threadIdx.x == 0) {
    const double sv = sqrt(nn);
    part[blockIdx.x] = sv > 0.0 ? bv / sqrt(sv) : 0.0;
  }
}

// part[0] = |mu1 - mu2|^2, part[1] = tr(s1), part[2] = tr(s2)
__global__ void __launch_bounds__(kRedThreads) frechet_terms_kernel(const double* __restrict__ mu1, const double* __restrict__ mu2,
                                                                    const double* __restrict__ s1, const double* __restrict__ s2, int d,
                                                                    double* __restrict__ part) {
  __shared__ double red[kRedThreads];
  double a = 0.0, t1 = 0.0, t2 = 0.0;
  for (int i = threadIdx.x; i < d; i += kRedThreads) {
    const double df = mu1[i] - mu2[i];
    a = fma(df, df, a);
    t1 += s1[(int64_t)i * d + i];
    t2 += s2[(int64_t)i * d + i];
  }
  a = block_sum(a, red); t1 = block_sum(t1, red); t2 = block_sum(t2, red);
  if (threadIdx.x == 0) { part[0] = a; part[1] = t1; part[2] = t2; }
}

__global__ void frechet_combine_kernel(const double* __restrict__ terms, const double* __restrict__ tr_covmean, double* __restrict__ out) {
  *out = terms[0] + terms[1] + terms[2] - 2.0 * tr_covmean[0];
}

// ---- polynomial-kernel MMD^2 ---------------------------------------------------------------------------------
// sums of k(x_i, y_j) = (gamma x_i . y_j + coef)^degree over a 32 x 32 tile of pairs; `skip_diag`: leave out i == j
__global__ void __launch_bounds__(256) poly_kernel_sum_kernel(const float* __restrict__ X, const float* __restrict__ Y, int m, int d,
                                                              int degree, double gamma, double coef, int skip_diag,
                                                              double* __restrict__ part) {
  __shared__ float sx[32][33], sy[32][33];
  __shared__ double red[kRedThreads];
  const int ti = blockIdx.y * 32, tj = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  double acc[4] = {0, 0, 0, 0};
  for (int k0 = 0; k0 < d; k0 += 32) {
    for (int r = ty; r < 32; r += 8) {
      sx[r][tx] = (ti + r < m && k0 + tx < d) ? X[(int64_t)(ti + r) * d + k0 + tx] : 0.f;
      sy[r][tx] = (tj + r < m && k0 + tx < d) ? Y[(int64_t)(tj + r) * d + k0 + tx] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < 32; ++k) {
      const double yv = (double)sy[tx][k];
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[q] = fma((double)sx[ty + 8 * q][k], yv, acc[q]);
    }
    __syncthreads();
  }
  double s = 0.0;
  for (int q = 0; q < 4; ++q) {
    const int i = ti + ty + 8 * q, j = tj + tx;
    if (i < m && j < m && !(skip_diag && i == j)) {
      const double base = acc[q] * gamma + coef;
      double pw = 1.0;
      for (int e = 0; e < degree; ++e) pw *= base;
      s += pw;
    }
  }
  const double r = block_sum(s, red);
  if (threadIdx.x == 0) part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
}

__global__ void mmd_combine_kernel(const double* __restrict__ sums, int m, double* __restrict__ out) {
  const double mm = (double)m;
  *out = (sums[0] + sums[1]) / (mm * (mm - 1.0)) - 2.0 * sums[2] / (mm * mm);
}

// ---- inception score -------------------------------------------------------------------------------------
// py[k] = mean over the split's rows of the ROW-NORMALISED probabilities is NOT what scipy does: entropy(pk, qk)
// normalises pk and qk separately; qk = py = mean of the raw rows (R/metric/IS_score.py:62-66).
__global__ void __launch_bounds__(kRedThreads) is_col_mean_kernel(const double* __restrict__ preds, int64_t row0, int64_t rows, int K,
                                                                  double* __restrict__ py) {
  __shared__ double red[kRedThreads];
  const int k = blockIdx.x;
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < rows; i += kRedThreads) acc += preds[(row0 + i) * K + k];
  const double r = block_sum(acc, red);
  if (threadIdx.x == 0) py[k] = r / (double)rows;
}
// kl[i] = sum_k p_k log(p_k / q_k) with p = row / sum(row), q = py / sum(py)
__global__ void __launch_bounds__(kRedThreads) is_row_kl_kernel(const double* __restrict__ preds, int64_t row0, int K,
                                                                const double* __restrict__ py, double* __restrict__ kl) {
  __shared__ double red[kRedThreads];
  const double* row = preds + (row0 + blockIdx.x) * K;
  double sp = 0.0, sq = 0.0;
  for (int k = threadIdx.x; k < K; k += kRedThreads) { sp += row[k]; sq += py[k]; }
  sp = block_sum(sp, red); sq = block_sum(sq, red);
  double acc = 0.0;
  for (int k = threadIdx.x; k < K; k += kRedThreads) {
    const double p = row[k] / sp, q = py[k] / sq;
    if (p > 0.0) acc += p * log(p / q);     // scipy.special.rel_entr: 0 where p == 0 (inf where q == 0 < p)
  }
  const double r = block_sum(acc, red);
  if (threadIdx.x == 0) kl[blockIdx.x] = r;
}
__global__ void is_split_score_kernel(const double* __restrict__ kl_sum, int64_t rows, double* __restrict__ score) {
  *score = exp(kl_sum[0] / (double)rows);
}
__global__ void is_mean_std_kernel(const double* __restrict__ scores, int splits, double* __restrict__ mean_out,
                                   double* __restrict__ std_out) {
  double m = 0.0;
  for (int i = 0; i < splits; ++i) m += scores[i];
  m /= splits;
  double v = 0.0;
  for (int i = 0; i < splits; ++i) v += (scores[i] - m) * (scores[i] - m);
  *mean_out = m;
  *std_out = sqrt(v / splits);     // np.std: population standard deviation
}

static int grid_cap(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  const int64_t cap = (int64_t)(sm_count() > 0 ? sm_count() : 148) * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace sd

using namespace sd;

extern "C" {

int64_t sd_metric_workspace_bytes(int64_t n_elements, int d, int m) {
  // partial sums of the element-wise reductions + (Frechet) product, B^T, V^T, column partials + (MMD) tile partials
  const int64_t dp = d + (d & 1);
  int64_t elems = 4096 + 16 + (n_elements > 0 ? n_elements : 0);   // n_elements: K + N / splits for the inception score
  if (d > 0) elems += 3 * dp * (int64_t)d + dp + 16;
  if (m > 0) { const int64_t t = (m + 31) / 32; elems += 3 * t * t + 16; }
  return elems * (int64_t)sizeof(double);
}

int sd_metric_mse(const float* a, const float* b, int64_t n, float* out_dev, void* workspace, void* stream) {
  SD_REQUIRE(a && b && out_dev && workspace && n >= 1, "metric_mse: bad arguments");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  double* part = (double*)workspace;
  const int blocks = grid_cap(n, kRedThreads) > 4096 ? 4096 : grid_cap(n, kRedThreads);
  mse_partial_kernel<<<blocks, kRedThreads, 0, st>>>(a, b, n, part);
  SD_LAUNCH_CHECK();
  final_sum_kernel<<<1, kRedThreads, 0, st>>>(part, blocks, 1.0 / (double)n, nullptr, out_dev);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_metric_ssim(const float* img1, const float* img2, int N, int C, int H, int W, int window_size,
                   const float* window_host, float* out_dev, float* per_plane_sum_or_null, void* workspace, void* stream) {
  SD_REQUIRE(img1 && img2 && out_dev && workspace && window_host, "metric_ssim: null pointer argument");
  SD_REQUIRE(N >= 1 && C >= 1 && H >= 1 && W >= 1, "metric_ssim: bad shape");
  SD_REQUIRE(window_size >= 1 && window_size <= 15 && (window_size & 1), "metric_ssim: window_size must be odd and <= 15");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  SsimWindow win;
  memset(&win, 0, sizeof(win));
  for (int i = 0; i < window_size * window_size; ++i) win.w[i] = window_host[i];
  const int64_t planes = (int64_t)N * C, total = planes * H * W;
  if (per_plane_sum_or_null) SD_CUDA(cudaMemsetAsync(per_plane_sum_or_null, 0, (size_t)planes * sizeof(float), st));
  double* part = (double*)workspace;
  int blocks = grid_cap(total, kRedThreads);
  if (blocks > 4096) blocks = 4096;
  ssim_partial_kernel<<<blocks, kRedThreads, 0, st>>>(img1, img2, planes, H, W, window_size, win, part, per_plane_sum_or_null);
  SD_LAUNCH_CHECK();
  final_sum_kernel<<<1, kRedThreads, 0, st>>>(part, blocks, 1.0 / (double)total, nullptr, out_dev);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_metric_feature_stats(const void* act, int act_is_f64, int64_t N, int d, double* mu_out, double* sigma_out, void* stream) {
  SD_REQUIRE(act && mu_out && sigma_out && N >= 2 && d >= 1, "metric_feature_stats: bad arguments (N >= 2)");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  dim3 grid((d + 31) / 32, (d + 31) / 32);
  if (act_is_f64) {
    col_mean_kernel<double><<<d, kRedThreads, 0, st>>>((const double*)act, N, d, mu_out);
    SD_LAUNCH_CHECK();
    cov_kernel<double><<<grid, 256, 0, st>>>((const double*)act, mu_out, N, d, sigma_out);
  } else {
    col_mean_kernel<float><<<d, kRedThreads, 0, st>>>((const float*)act, N, d, mu_out);
    SD_LAUNCH_CHECK();
    cov_kernel<float><<<grid, 256, 0, st>>>((const float*)act, mu_out, N, d, sigma_out);
  }
  SD_LAUNCH_CHECK();
  return SD_OK;
}

// Synchronous in the stream (the Jacobi sweeps are repeated until a device-side flag says no rotation was needed).
int sd_metric_frechet(const double* mu1, const double* sigma1, const double* mu2, const double* sigma2, int d, double* out_dev,
                      int* sweeps_out, void* workspace, void* stream) {
  SD_REQUIRE(mu1 && sigma1 && mu2 && sigma2 && out_dev && workspace && d >= 1, "metric_frechet: bad arguments");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  const int dp = d + (d & 1);
  double* ws = (double*)workspace;
  double* part = ws;                         // 4096 + 16
  double* Bt = ws + 4096 + 16;               // [dp][d]: rows = columns of A = sigma1 sigma2, i.e. A^T = sigma2 sigma1
  double* Vt = Bt + (int64_t)dp * d;         // [dp][d]
  double* colpart = Vt + (int64_t)dp * d;    // [dp]
  int* flag = (int*)(colpart + dp);
  dim3 grid((d + 31) / 32, (d + 31) / 32);
  SD_CUDA(cudaMemsetAsync(Bt, 0, (size_t)dp * d * sizeof(double), st));
  dgemm_nn_kernel<<<grid, 256, 0, st>>>(sigma2, sigma1, d, Bt);
  SD_LAUNCH_CHECK();
  set_identity_kernel<<<grid_cap((int64_t)dp * d, 256), 256, 0, st>>>(Vt, d, dp);
  SD_LAUNCH_CHECK();
  int sweeps = 0;
  if (dp >= 2) {
    for (; sweeps < 60; ++sweeps) {
      SD_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
      for (int r = 0; r < dp - 1; ++r) {
        jacobi_round_kernel<<<dp / 2, kRedThreads, 0, st>>>(Bt, Vt, d, dp, r, 1e-15, flag);
      }
      SD_LAUNCH_CHECK();
      int h = 0;
      SD_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
      SD_CUDA(cudaStreamSynchronize(st));
      if (!h) break;
    }
  }
  if (sweeps_out) *sweeps_out = sweeps;
  sqrtm_trace_partial_kernel<<<d, kRedThreads, 0, st>>>(Bt, Vt, d, colpart);
  SD_LAUNCH_CHECK();
  final_sum_kernel<<<1, kRedThreads, 0, st>>>(colpart, d, 1.0, part + 8, nullptr);
  SD_LAUNCH_CHECK();
  frechet_terms_kernel<<<1, kRedThreads, 0, st>>>(mu1, mu2, sigma1, sigma2, d, part);
  SD_LAUNCH_CHECK();
  frechet_combine_kernel<<<1, 1, 0, st>>>(part, part + 8, out_dev);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_metric_poly_mmd2(const float* fx, const float* fy, int m, int d, int degree, double gamma, double coef, double* out_dev,
                        void* workspace, void* stream) {
  SD_REQUIRE(fx && fy && out_dev && workspace && m >= 2 && d >= 1 && degree >= 1 && degree <= 8, "metric_poly_mmd2: bad arguments");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  const int t = (m + 31) / 32;
  double* ws = (double*)workspace;
  double* sums = ws;                 // [3]
  double* tiles = ws + 16;           // 3 x t x t
  dim3 grid(t, t);
  const float* xs[3] = {fx, fy, fx};
  const float* ys[3] = {fx, fy, fy};
  for (int k = 0; k < 3; ++k) {
    poly_kernel_sum_kernel<<<grid, 256, 0, st>>>(xs[k], ys[k], m, d, degree, gamma, coef, k < 2, tiles + (int64_t)k * t * t);
    SD_LAUNCH_CHECK();
    final_sum_kernel<<<1, kRedThreads, 0, st>>>(tiles + (int64_t)k * t * t, (int64_t)t * t, 1.0, sums + k, nullptr);
    SD_LAUNCH_CHECK();
  }
  mmd_combine_kernel<<<1, 1, 0, st>>>(sums, m, out_dev);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_metric_inception_score(const double* preds, int64_t N, int K, int splits, double* mean_out_dev, double* std_out_dev,
                              void* workspace, void* stream) {
  SD_REQUIRE(preds && mean_out_dev && std_out_dev && workspace && N >= 1 && K >= 1 && splits >= 1 && splits <= 64 && N / splits >= 1,
             "metric_inception_score: bad arguments");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  double* ws = (double*)workspace;
  double* scores = ws;            // [64]
  double* klsum = ws + 64;        // [1]
  double* py = ws + 72;           // [K]
  double* kl = py + K;            // [rows]
  const int64_t rows = N / splits;
  for (int s = 0; s < splits; ++s) {
    const int64_t row0 = s * rows;
    is_col_mean_kernel<<<K, kRedThreads, 0, st>>>(preds, row0, rows, K, py);
    SD_LAUNCH_CHECK();
    is_row_kl_kernel<<<(unsigned)rows, kRedThreads, 0, st>>>(preds, row0, K, py, kl);
    SD_LAUNCH_CHECK();
    final_sum_kernel<<<1, kRedThreads, 0, st>>>(kl, rows, 1.0, klsum, nullptr);
    SD_LAUNCH_CHECK();
    is_split_score_kernel<<<1, 1, 0, st>>>(klsum, rows, scores + s);
    SD_LAUNCH_CHECK();
  }
  is_mean_std_kernel<<<1, 1, 0, st>>>(scores, splits, mean_out_dev, std_out_dev);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
