// Training-path kernels, first correct versions on CUDA cores:
//   sd_conv_wgrad        weight / bias gradient of layer.Conv2d / layer.ConvTranspose2d in 'm' mode
//   sd_bn_train_forward  train-mode BatchNorm2d (batch statistics over T*N*H*W)      SJ/activation_based/layer.py:458-465
//   sd_bn_backward       its backward
// The input gradient of a convolution is itself a (transposed) convolution and reuses conv_simt.cu.
// These kernels are correctness-first (parity with torch autograd); they are not on the sampling hot path.
#include "common.cuh"

namespace sd {

constexpr int kMaxTaps = 25;

// One block per (co, ci): threads sweep (n, oy, ox), accumulate one partial sum per tap, block-reduce.
//   conv      : gw[co,ci,ky,kx] = sum gy[n,co,oy,ox] * x[n,ci,oy*s-p+ky, ox*s-p+kx]
//   transposed: gw[ci,co,ky,kx] = sum x[n,ci,iy,ix] * gy[n,co,iy*s-p+ky, ix*s-p+kx]
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                         float* __restrict__ gw, sd_conv_desc d, int64_t n_outer) {
  const int co = blockIdx.x / d.C_in, ci = blockIdx.x % d.C_in;
  const int taps = d.kh * d.kw;
  float acc[kMaxTaps];
#pragma unroll
  for (int k = 0; k < kMaxTaps; ++k) acc[k] = 0.f;
  // the "anchor" grid is the conv output (conv) or the conv-transpose input (transposed)
  const int Ha = d.transposed ? d.H_in : d.H_out, Wa = d.transposed ? d.W_in : d.W_out;
  const int Hb = d.transposed ? d.H_out : d.H_in, Wb = d.transposed ? d.W_out : d.W_in;
  const int64_t total = n_outer * Ha * Wa;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int ax = (int)(i % Wa);
    int64_t r = i / Wa;
    const int ay = (int)(r % Ha);
    const int64_t n = r / Ha;
    // value on the anchor grid and plane on the other grid
    const float va = d.transposed ? x[((n * d.C_in + ci) * Ha + ay) * Wa + ax] : gy[((n * d.C_out + co) * Ha + ay) * Wa + ax];
    const float* pb = d.transposed ? gy + (n * d.C_out + co) * (int64_t)Hb * Wb : x + (n * d.C_in + ci) * (int64_t)Hb * Wb;
#pragma unroll
    for (int k = 0; k < kMaxTaps; ++k) {
      if (k < taps) {
        const int ky = k / d.kw, kx = k - ky * d.kw;
        const int by = ay * d.stride - d.pad + ky, bx = ax * d.stride - d.pad + kx;
        if (by >= 0 && by < Hb && bx >= 0 && bx < Wb) acc[k] = fmaf(va, pb[by * Wb + bx], acc[k]);
      }
    }
  }
  __shared__ float red[8][kMaxTaps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMaxTaps; ++k) {
    if (k < taps) {
      float v = acc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) red[warp][k] = v;
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < taps) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
    const int64_t idx = d.transposed ? (((int64_t)ci * d.C_out + co) * taps + threadIdx.x)
                                     : (((int64_t)co * d.C_in + ci) * taps + threadIdx.x);
    gw[idx] = v;
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

}  // namespace sd

using namespace sd;

extern "C" {

int sd_conv_wgrad(const sd_conv_desc* d, const float* x, const float* grad_out, float* grad_w, float* grad_bias,
                  void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  SD_REQUIRE(d->kh * d->kw <= kMaxTaps, "conv_wgrad: kernel larger than 5x5 is not supported");
  SD_REQUIRE(x && grad_out && grad_w, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  const int64_t n_outer = (int64_t)d->T * d->B;
  conv_wgrad_kernel<<<(unsigned)(d->C_out * d->C_in), 256, 0, st>>>(x, grad_out, grad_w, *d, n_outer);
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

}  // extern "C"
