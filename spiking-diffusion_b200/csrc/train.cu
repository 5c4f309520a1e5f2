// Training-path kernels on CUDA cores (fp32, exact products):
//   sd_conv_wgrad        weight / bias gradient of layer.Conv2d / layer.ConvTranspose2d in 'm' mode
//   sd_bn_train_forward  train-mode BatchNorm2d (batch statistics over T*N*H*W)      SJ/activation_based/layer.py:458-465
//   sd_bn_backward       its backward
// The input gradient of a convolution is itself a (transposed) convolution and reuses conv_simt.cu.
// Parity target is torch autograd in fp32; they are not on the sampling hot path.
#include "common.cuh"

namespace sd {

// Weight gradient as a tiled GEMM on CUDA cores.  With U the tensor that is gathered through the kernel window and V the
// one on the anchor grid,
//   conv      : U = x  (a = ci), V = gy (b = co), anchor = output pixel:  gw[co, ci, tap] = sum_p U[p + tap] V[p]
//   transposed: U = gy (a = co), V = x  (b = ci), anchor = input pixel :  gw[ci, co, tap] = sum_p U[p + tap] V[p]
// where "p + tap" is the position anchor * stride - pad + (ky, kx) on U's grid.  One block = 64 rows (tap, a) x 64
// columns b, reduction over a slice of the n_outer * Ha * Wa anchor pixels (grid.z splits), 4 x 4 outputs per thread;
// partial sums go to a [splits][rows][cols] workspace and a second kernel adds them in a fixed order (deterministic) and
// scatters to the reference's parameter layout.
constexpr int kWgBR = 64, kWgBC = 64, kWgBP = 16, kWgPad = 68;

__global__ void __launch_bounds__(256) conv_wgrad_tiled_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                               float* __restrict__ part, sd_conv_desc d, int64_t n_outer,
                                                               int64_t p_per_split) {
  __shared__ __align__(16) float Us[2][kWgBP][kWgPad];
  __shared__ __align__(16) float Vs[2][kWgBP][kWgPad];
  const int taps = d.kh * d.kw;
  const int A = d.transposed ? d.C_out : d.C_in;    // channels of U
  const int Bc = d.transposed ? d.C_in : d.C_out;   // channels of V
  const float* U = d.transposed ? gy : x;
  const float* V = d.transposed ? x : gy;
  const int Ha = d.transposed ? d.H_in : d.H_out, Wa = d.transposed ? d.W_in : d.W_out;   // anchor grid (V)
  const int Hu = d.transposed ? d.H_out : d.H_in, Wu = d.transposed ? d.W_out : d.W_in;   // gathered grid (U)
  const int R = taps * A;
  const int64_t P = n_outer * Ha * Wa;
  const int r0 = blockIdx.x * kWgBR, c0 = blockIdx.y * kWgBC;
  const int64_t p_begin = (int64_t)blockIdx.z * p_per_split;
  const int64_t p_end = p_begin + p_per_split < P ? p_begin + p_per_split : P;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int pl = tid & 15, el = tid >> 4;   // load mapping: pixel fastest (coalesced along x), 16 rows / cols per pass
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // (tap, a) of the four U rows this thread loads
  int r_a[4], r_ky[4], r_kx[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + el + 16 * j;
    const int tap = r < R ? r / A : 0;
    r_a[j] = r < R ? r - tap * A : -1;
    r_ky[j] = tap / d.kw;
    r_kx[j] = tap - r_ky[j] * d.kw;
  }
  // the next slice of 16 anchor pixels is fetched into registers while the current one is multiplied
  float u_reg[4], v_reg[4];
  auto fetch = [&](int64_t pb) {
    const int64_t pidx = pb + pl;
    const bool p_ok = pidx < p_end;
    int64_t n = 0;
    int ay = 0, ax = 0;
    if (p_ok) {
      n = pidx / ((int64_t)Ha * Wa);
      const int pp = (int)(pidx - n * Ha * Wa);
      ay = pp / Wa;
      ax = pp - ay * Wa;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = 0.f, v = 0.f;
      if (p_ok && r_a[j] >= 0) {
        const int uy = ay * d.stride - d.pad + r_ky[j], ux = ax * d.stride - d.pad + r_kx[j];
        if (uy >= 0 && uy < Hu && ux >= 0 && ux < Wu) u = U[((n * A + r_a[j]) * Hu + uy) * Wu + ux];
      }
      const int c = c0 + el + 16 * j;
      if (p_ok && c < Bc) v = V[((n * Bc + c) * Ha + ay) * Wa + ax];
      u_reg[j] = u;
      v_reg[j] = v;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      Us[buf][pl][el + 16 * j] = u_reg[j];
      Vs[buf][pl][el + 16 * j] = v_reg[j];
    }
  };
  if (p_begin < p_end) {
    fetch(p_begin);
    stage(0);
  }
  __syncthreads();
  int buf = 0;
  for (int64_t pb = p_begin; pb < p_end; pb += kWgBP, buf ^= 1) {
    const bool more = pb + kWgBP < p_end;
    if (more) fetch(pb + kWgBP);
#pragma unroll
    for (int pp = 0; pp < kWgBP; ++pp) {
      const float4 a4 = *reinterpret_cast<const float4*>(&Us[buf][pp][tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Vs[buf][pp][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) stage(buf ^ 1);
    __syncthreads();
  }
  float* out = part + (int64_t)blockIdx.z * R * Bc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + tx * 4 + i;
    if (r >= R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + ty * 4 + j;
      if (c < Bc) out[(int64_t)r * Bc + c] = acc[i][j];
    }
  }
}

// gw (reference layout) = sum over splits of part[split][(tap, a)][b]
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ gw,
                                                                sd_conv_desc d, int splits) {
  const int taps = d.kh * d.kw;
  const int A = d.transposed ? d.C_out : d.C_in, Bc = d.transposed ? d.C_in : d.C_out;
  const int64_t total = (int64_t)taps * A * Bc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(int64_t)z * total + i];
    const int b = (int)(i % Bc);
    const int64_t r = i / Bc;
    const int a = (int)(r % A), tap = (int)(r / A);
    // conv: a = ci, b = co -> gw[co][ci][tap];  transposed: a = co, b = ci -> gw[ci][co][tap]
    gw[((int64_t)b * A + a) * taps + tap] = s;
  }
}

// per-channel sum of x[n, c, :] over n and the plane: bias gradient and BN reductions share it
template <int NQ>
__device__ __forceinline__ void block_reduce(float (&v)[NQ], float* smem /* [8][NQ] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], off);
    if (lane == 0) smem[warp * NQ + q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += smem[w * NQ + q];
    v[q] = s;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) channel_sum_kernel(const float* __restrict__ g, float* __restrict__ out,
                                                          int64_t n_outer, int C, int64_t HW) {
  __shared__ float sm[8];
  const int c = blockIdx.x;
  float v[1] = {0.f};
  for (int64_t i = threadIdx.x; i < n_outer * HW; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    v[0] += g[(n * C + c) * HW + p];
  }
  block_reduce<1>(v, sm);
  if (threadIdx.x == 0) out[c] = v[0];
}

// y = (x - mean) * invstd * gamma + beta with batch statistics; mean / biased var written for the backward pass and
// for the running-statistics update done by the caller (momentum, unbiased correction: F.batch_norm semantics).
__global__ void __launch_bounds__(256) bn_train_forward_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ y,
                                                               float* __restrict__ mean_out, float* __restrict__ var_out,
                                                               int64_t n_outer, int C, int64_t HW, float eps) {
  __shared__ float sm[16];
  const int c = blockIdx.x;
  const int64_t cnt = n_outer * HW;
  float v[1] = {0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    v[0] += x[(n * C + c) * HW + p];
  }
  block_reduce<1>(v, sm);
  const float mean = v[0] / (float)cnt;
  float q[1] = {0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const float dlt = x[(n * C + c) * HW + p] - mean;
    q[0] = fmaf(dlt, dlt, q[0]);
  }
  block_reduce<1>(q, sm);
  const float var = q[0] / (float)cnt;
  const float invstd = rsqrtf(var + eps);
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const int64_t o = (n * C + c) * HW + p;
    y[o] = fmaf((x[o] - mean) * invstd, g, b);
  }
  if (threadIdx.x == 0) { mean_out[c] = mean; var_out[c] = var; }
}

// gx = gamma * invstd * (gy - mean(gy) - xhat * mean(gy * xhat)),  ggamma = sum gy * xhat,  gbeta = sum gy
__global__ void __launch_bounds__(256) bn_backward_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                          const float* __restrict__ mean, const float* __restrict__ var,
                                                          const float* __restrict__ gamma, float* __restrict__ gx,
                                                          float* __restrict__ ggamma, float* __restrict__ gbeta,
                                                          int64_t n_outer, int C, int64_t HW, float eps) {
  __shared__ float sm[16];
  const int c = blockIdx.x;
  const int64_t cnt = n_outer * HW;
  const float mu = mean[c], invstd = rsqrtf(var[c] + eps);
  float v[2] = {0.f, 0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const int64_t o = (n * C + c) * HW + p;
    const float g = gy[o];
    v[0] += g;
    v[1] = fmaf(g, (x[o] - mu) * invstd, v[1]);
  }
  block_reduce<2>(v, sm);
  const float gm = gamma ? gamma[c] : 1.f;
  const float m0 = v[0] / (float)cnt, m1 = v[1] / (float)cnt;
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const int64_t o = (n * C + c) * HW + p;
    const float xhat = (x[o] - mu) * invstd;
    gx[o] = gm * invstd * (gy[o] - m0 - xhat * m1);
  }
  if (threadIdx.x == 0) {
    if (ggamma) ggamma[c] = v[1];
    if (gbeta) gbeta[c] = v[0];
  }
}

// ---- pieces of train-mode BatchNorm for statistics shared across GPUs (SyncBN) ------------------------------------
// local mean and centred sum of squares M2 per channel; the caller combines the ranks' (count, mean, M2) triples with the
// parallel-variance formula, which keeps the two-pass accuracy of the single-GPU kernel
__global__ void __launch_bounds__(256) bn_local_stats_kernel(const float* __restrict__ x, float* __restrict__ mean_out,
                                                             float* __restrict__ m2_out, int64_t n_outer, int C,
                                                             int64_t HW) {
  __shared__ float sm[16];
  const int c = blockIdx.x;
  const int64_t cnt = n_outer * HW;
  float v[1] = {0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    v[0] += x[(n * C + c) * HW + p];
  }
  block_reduce<1>(v, sm);
  const float mean = v[0] / (float)cnt;
  float q[1] = {0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const float dlt = x[(n * C + c) * HW + p] - mean;
    q[0] = fmaf(dlt, dlt, q[0]);
  }
  block_reduce<1>(q, sm);
  if (threadIdx.x == 0) { mean_out[c] = mean; m2_out[c] = q[0]; }
}

// local sums of gy and gy * xhat per channel (xhat from the GLOBAL mean / variance)
__global__ void __launch_bounds__(256) bn_backward_reduce_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                 const float* __restrict__ mean, const float* __restrict__ var,
                                                                 float* __restrict__ sum_gy, float* __restrict__ sum_gy_xhat,
                                                                 int64_t n_outer, int C, int64_t HW, float eps) {
  __shared__ float sm[16];
  const int c = blockIdx.x;
  const int64_t cnt = n_outer * HW;
  const float mu = mean[c], invstd = rsqrtf(var[c] + eps);
  float v[2] = {0.f, 0.f};
  for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int64_t n = i / HW, p = i - n * HW;
    const int64_t o = (n * C + c) * HW + p;
    const float g = gy[o];
    v[0] += g;
    v[1] = fmaf(g, (x[o] - mu) * invstd, v[1]);
  }
  block_reduce<2>(v, sm);
  if (threadIdx.x == 0) { sum_gy[c] = v[0]; sum_gy_xhat[c] = v[1]; }
}

// gx = gamma * invstd * (gy - mean_gy - xhat * mean_gy_xhat) with the GLOBAL means of gy and gy * xhat
__global__ void __launch_bounds__(256) bn_backward_apply_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                const float* __restrict__ mean, const float* __restrict__ var,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ mean_gy,
                                                                const float* __restrict__ mean_gy_xhat, float* __restrict__ gx,
                                                                int64_t n_outer, int C, int64_t HW, float eps) {
  const int64_t total = n_outer * C * HW;
  for (int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((o / HW) % C);
    const float invstd = rsqrtf(var[c] + eps);
    const float xhat = (x[o] - mean[c]) * invstd;
    const float gm = gamma ? gamma[c] : 1.f;
    gx[o] = gm * invstd * (gy[o] - mean_gy[c] - xhat * mean_gy_xhat[c]);
  }
}

}  // namespace sd


using namespace sd;

extern "C" {

// reduction splits: enough blocks for two waves of the SMs, slices of at least 256 anchor pixels
static int wgrad_splits(const sd_conv_desc* d) {
  const int taps = d->kh * d->kw;
  const int A = d->transposed ? d->C_out : d->C_in, Bc = d->transposed ? d->C_in : d->C_out;
  const int64_t P = (int64_t)d->T * d->B * (d->transposed ? d->H_in * d->W_in : d->H_out * d->W_out);
  const int64_t tiles = (int64_t)((taps * A + kWgBR - 1) / kWgBR) * ((Bc + kWgBC - 1) / kWgBC);
  int64_t s = (2 * 148 + tiles - 1) / tiles;
  const int64_t smax = P / 256 > 0 ? P / 256 : 1;
  if (s > smax) s = smax;
  if (s > 64) s = 64;
  return (int)(s < 1 ? 1 : s);
}

int64_t sd_conv_wgrad_workspace_bytes(const sd_conv_desc* d) {
  if (!d || validate_conv_desc(d) != SD_OK) return 0;
  return (int64_t)wgrad_splits(d) * d->kh * d->kw * d->C_in * d->C_out * (int64_t)sizeof(float);
}

int sd_conv_wgrad(const sd_conv_desc* d, const float* x, const float* grad_out, float* grad_w, float* grad_bias,
                  void* workspace, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  SD_REQUIRE(x && grad_out && grad_w && workspace, "null pointer argument (workspace: sd_conv_wgrad_workspace_bytes)");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  const int64_t n_outer = (int64_t)d->T * d->B;
  const int taps = d->kh * d->kw;
  const int A = d->transposed ? d->C_out : d->C_in, Bc = d->transposed ? d->C_in : d->C_out;
  const int64_t P = n_outer * (d->transposed ? d->H_in * d->W_in : d->H_out * d->W_out);
  const int splits = wgrad_splits(d);
  int64_t per = (P + splits - 1) / splits;
  per = (per + kWgBP - 1) / kWgBP * kWgBP;
  dim3 grid((unsigned)((taps * A + kWgBR - 1) / kWgBR), (unsigned)((Bc + kWgBC - 1) / kWgBC), (unsigned)splits);
  conv_wgrad_tiled_kernel<<<grid, 256, 0, st>>>(x, grad_out, (float*)workspace, *d, n_outer, per);
  SD_LAUNCH_CHECK();
  const int64_t total = (int64_t)taps * A * Bc;
  conv_wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096), 256, 0, st>>>(
      (const float*)workspace, grad_w, *d, splits);
  SD_LAUNCH_CHECK();
  if (grad_bias) {
    channel_sum_kernel<<<(unsigned)d->C_out, 256, 0, st>>>(grad_out, grad_bias, n_outer, d->C_out,
                                                           (int64_t)d->H_out * d->W_out);
    SD_LAUNCH_CHECK();
  }
  return SD_OK;
}

int sd_bn_train_forward(const float* x, const float* gamma, const float* beta, float* y, float* mean_out,
                        float* var_out, int64_t n_outer, int C, int64_t HW, float eps, void* stream) {
  SD_REQUIRE(n_outer >= 1 && C >= 1 && HW >= 1, "bn_train_forward: bad shape");
  SD_REQUIRE(x && y && mean_out && var_out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  bn_train_forward_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(x, gamma, beta, y, mean_out, var_out, n_outer, C, HW,
                                                                     eps);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_bn_backward(const float* x, const float* grad_out, const float* mean, const float* var, const float* gamma,
                   float* grad_x, float* grad_gamma, float* grad_beta, int64_t n_outer, int C, int64_t HW, float eps,
                   void* stream) {
  SD_REQUIRE(n_outer >= 1 && C >= 1 && HW >= 1, "bn_backward: bad shape");
  SD_REQUIRE(x && grad_out && mean && var && grad_x, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  bn_backward_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(x, grad_out, mean, var, gamma, grad_x, grad_gamma, grad_beta,
                                                                n_outer, C, HW, eps);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_bn_local_stats(const float* x, float* mean_out, float* m2_out, int64_t n_outer, int C, int64_t HW, void* stream) {
  SD_REQUIRE(n_outer >= 1 && C >= 1 && HW >= 1, "bn_local_stats: bad shape");
  SD_REQUIRE(x && mean_out && m2_out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  bn_local_stats_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(x, mean_out, m2_out, n_outer, C, HW);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_bn_backward_reduce(const float* x, const float* grad_out, const float* mean, const float* var, float* sum_gy,
                          float* sum_gy_xhat, int64_t n_outer, int C, int64_t HW, float eps, void* stream) {
  SD_REQUIRE(n_outer >= 1 && C >= 1 && HW >= 1, "bn_backward_reduce: bad shape");
  SD_REQUIRE(x && grad_out && mean && var && sum_gy && sum_gy_xhat, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  bn_backward_reduce_kernel<<<(unsigned)C, 256, 0, as_stream(stream)>>>(x, grad_out, mean, var, sum_gy, sum_gy_xhat, n_outer,
                                                                       C, HW, eps);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_bn_backward_apply(const float* x, const float* grad_out, const float* mean, const float* var, const float* gamma,
                         const float* mean_gy, const float* mean_gy_xhat, float* grad_x, int64_t n_outer, int C, int64_t HW,
                         float eps, void* stream) {
  SD_REQUIRE(n_outer >= 1 && C >= 1 && HW >= 1, "bn_backward_apply: bad shape");
  SD_REQUIRE(x && grad_out && mean && var && mean_gy && mean_gy_xhat && grad_x, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  const int64_t total = n_outer * C * HW;
  int64_t bl = (total + 255) / 256;
  if (bl > 148 * 16) bl = 148 * 16;
  bn_backward_apply_kernel<<<(unsigned)bl, 256, 0, as_stream(stream)>>>(x, grad_out, mean, var, gamma, mean_gy, mean_gy_xhat,
                                                                        grad_x, n_outer, C, HW, eps);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
