// Vector-quantiser kernels: spike feature mix, nearest-codeword search, codebook gather.
//
// sd_vq_lookup follows R/snn_model/vae_model.py:87-95 in the reference's operation order,
//   d[m,k] = (sum_d z[m,d]^2 + sum_d e[k,d]^2) - 2 * (z[m,:] . e[k,:]),   idx[m] = argmin_k d[m,k] (first min),
// all in fp32.  One warp owns one token: the 16-dim z row lives in registers (broadcast), the codebook tile
// is staged once per CTA in shared memory together with |e_k|^2, lanes stride over the codes and a
// warp-shuffle (min, index) reduction with lowest-index tie-break picks the winner.  The problem is tiny
// (M*K*D MACs, 0.2 MFLOP per image) and latency-bound; bytes moved are M*D*4 (z) + K*D*4 (codebook, L2
// resident) + M*8 (idx).
#include "common.cuh"

namespace sd {

// z[m, d] = (1 - alpha) * sum_t coef[t] * s[t] + alpha * (sum_t s[t]) / T      (vae_model.py:42)
__global__ void vq_feature_kernel(const __half* __restrict__ spk, const float* __restrict__ alpha_dev,
                                  MemoutCoef coef, float* __restrict__ z, int T, int B, int D, int H, int W) {
  StfGeom g(B, H, W);
  const int D8 = c8(D);
  const float alpha = *alpha_dev;
  const float one_minus = __fsub_rn(1.0f, alpha);
  const int64_t total = (int64_t)B * H * W * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(i % D);
    int64_t r = i / D;
    int x = (int)(r % W); r /= W;
    int y = (int)(r % H);
    int b = (int)(r / H);
    const int64_t row = g.row(b, y, x);
    float mem = 0.f, cnt = 0.f;
    for (int t = 0; t < T; ++t) {
      float s = __half2float(spk[g.at(t, D8, d, row)]);
      mem = __fadd_rn(mem, __fmul_rn(s, coef.c[t]));
      cnt = __fadd_rn(cnt, s);
    }
    float a = __fmul_rn(one_minus, mem);
    float bterm = __fdiv_rn(__fmul_rn(alpha, cnt), (float)T);
    z[i] = __fadd_rn(a, bterm);
  }
}

constexpr int kVqWarps = 8;
constexpr int kVqTileK = 256;  // codes staged per shared-memory tile

template <int D>
__global__ void __launch_bounds__(kVqWarps * 32) vq_lookup_kernel(const float* __restrict__ z,
                                                                  const float* __restrict__ cb,
                                                                  int64_t* __restrict__ idx,
                                                                  float* __restrict__ margin, int64_t M, int K) {
  __shared__ float s_cb[kVqTileK * (D + 1)];  // +1 padding: lanes read rows k, k+32.. -> distinct banks
  __shared__ float s_ee[kVqTileK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t tokens_per_cta = kVqWarps * 4;  // each warp walks 4 tokens per codebook tile pass
  auto stage = [&](int k0, int kt) {
    __syncthreads();
    for (int i = threadIdx.x; i < kt * D; i += blockDim.x) {
      int k = i / D, d = i - k * D;
      s_cb[k * (D + 1) + d] = cb[(int64_t)(k0 + k) * D + d];
    }
    __syncthreads();
    for (int k = threadIdx.x; k < kt; k += blockDim.x) {
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) acc = __fadd_rn(acc, __fmul_rn(s_cb[k * (D + 1) + d], s_cb[k * (D + 1) + d]));
      s_ee[k] = acc;
    }
    __syncthreads();
  };
  const bool single = K <= kVqTileK;  // whole codebook fits one tile: stage it once per CTA
  if (single) stage(0, K);
  for (int64_t base = (int64_t)blockIdx.x * tokens_per_cta; base < M; base += (int64_t)gridDim.x * tokens_per_cta) {
    float best[4], second[4];
    int besti[4];
    float zz[4];
    float zr[4][D];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      best[q] = INFINITY; second[q] = INFINITY; besti[q] = 0;
      int64_t m = base + warp * 4 + q;
      float acc = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        zr[q][d] = (m < M) ? z[m * D + d] : 0.f;
        acc = __fadd_rn(acc, __fmul_rn(zr[q][d], zr[q][d]));
      }
      zz[q] = acc;
    }
    for (int k0 = 0; k0 < K; k0 += kVqTileK) {
      const int kt = min(kVqTileK, K - k0);
      if (!single) stage(k0, kt);
      for (int k = lane; k < kt; k += 32) {
        const float* e = &s_cb[k * (D + 1)];
        const float ee = s_ee[k];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float dot = 0.f;
#pragma unroll
          for (int d = 0; d < D; ++d) dot = fmaf(zr[q][d], e[d], dot);
          float dist = __fsub_rn(__fadd_rn(zz[q], ee), __fmul_rn(2.0f, dot));
          if (dist < best[q]) { second[q] = best[q]; best[q] = dist; besti[q] = k0 + k; }
          else if (dist < second[q]) { second[q] = dist; }
        }
      }
    }
    // warp argmin with lowest-index tie-break; also carries the runner-up for the margin output
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float b = best[q], s2 = second[q];
      int bi = besti[q];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, b, off);
        float os = __shfl_xor_sync(0xffffffffu, s2, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        bool take = (ob < b) || (ob == b && oi < bi);
        float loser = take ? b : ob;
        if (take) { b = ob; bi = oi; }
        s2 = fminf(fminf(s2, os), loser);
      }
      int64_t m = base + warp * 4 + q;
      if (lane == 0 && m < M) {
        idx[m] = bi;
        if (margin) margin[m] = s2 - b;
      }
    }
  }
}

// Any embedding dimension (the reference's --embedding_dim is free; 8 / 16 / 32 use the register-blocked kernel above):
// one warp per token, lanes over the codes, the same arithmetic order (sequential |z|^2 and |e|^2, FMA dot product,
// (|z|^2 + |e|^2) - 2 z.e, first index on ties).
__global__ void __launch_bounds__(256) vq_lookup_generic_kernel(const float* __restrict__ z, const float* __restrict__ cb,
                                                                int64_t* __restrict__ idx, float* __restrict__ margin,
                                                                int64_t M, int D, int K) {
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t m = warp_global; m < M; m += n_warps) {
    const float* zr = z + m * D;
    float zz = 0.f;
    for (int d = 0; d < D; ++d) zz = __fadd_rn(zz, __fmul_rn(zr[d], zr[d]));
    float b = INFINITY, s2 = INFINITY;
    int bi = 0;
    for (int k = lane; k < K; k += 32) {
      const float* e = cb + (int64_t)k * D;
      float ee = 0.f, dot = 0.f;
      for (int d = 0; d < D; ++d) {
        ee = __fadd_rn(ee, __fmul_rn(e[d], e[d]));
        dot = fmaf(zr[d], e[d], dot);
      }
      const float dist = __fsub_rn(__fadd_rn(zz, ee), __fmul_rn(2.0f, dot));
      if (dist < b) { s2 = b; b = dist; bi = k; }
      else if (dist < s2) { s2 = dist; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, b, off);
      float os = __shfl_xor_sync(0xffffffffu, s2, off);
      int oi = __shfl_xor_sync(0xffffffffu, bi, off);
      bool take = (ob < b) || (ob == b && oi < bi);
      float loser = take ? b : ob;
      if (take) { b = ob; bi = oi; }
      s2 = fminf(fminf(s2, os), loser);
    }
    if (lane == 0) {
      idx[m] = bi;
      if (margin) margin[m] = s2 - b;
    }
  }
}

__global__ void vq_gather_kernel(const int64_t* __restrict__ idx, const float* __restrict__ cb,
                                 float* __restrict__ out, int B, int D, int H, int W, int K) {
  const int64_t total = (int64_t)B * D * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int d = (int)(r % D);
    int b = (int)(r / D);
    int64_t k = idx[((int64_t)b * H + y) * W + x];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);  // nn.Embedding would raise on out-of-range ids; clamp defensively
    out[i] = cb[k * D + d];
  }
}

static inline unsigned grid_cap(int64_t blocks) {
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

}  // namespace sd

using namespace sd;

extern "C" {

int sd_vq_feature(const void* spikes_stf, const float* alpha_dev, const float* coef_host, float* z, int T, int B,
                  int D, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && T <= SD_MAX_T, "vq_feature: T=%d out of range", T);
  SD_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1, "vq_feature: bad shape");
  SD_REQUIRE(spikes_stf && alpha_dev && coef_host && z, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  MemoutCoef hc;
  for (int t = 0; t < SD_MAX_T; ++t) hc.c[t] = t < T ? coef_host[t] : 0.f;
  int64_t n = (int64_t)B * H * W * D;
  vq_feature_kernel<<<grid_cap((n + 255) / 256), 256, 0, as_stream(stream)>>>((const __half*)spikes_stf, alpha_dev, hc,
                                                                            z, T, B, D, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_vq_lookup(const float* z, const float* codebook, int64_t* idx, float* margin, int64_t M, int D, int K,
                 void* stream) {
  SD_REQUIRE(M >= 0 && K >= 1 && D >= 1, "vq_lookup: bad M=%lld D=%d K=%d", (long long)M, D, K);
  if (M == 0) return SD_OK;
  SD_REQUIRE(z && codebook && idx, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  unsigned grid = grid_cap((M + kVqWarps * 4 - 1) / (kVqWarps * 4));
  switch (D) {
    case 8: vq_lookup_kernel<8><<<grid, kVqWarps * 32, 0, st>>>(z, codebook, idx, margin, M, K); break;
    case 16: vq_lookup_kernel<16><<<grid, kVqWarps * 32, 0, st>>>(z, codebook, idx, margin, M, K); break;
    case 32: vq_lookup_kernel<32><<<grid, kVqWarps * 32, 0, st>>>(z, codebook, idx, margin, M, K); break;
    default:
      vq_lookup_generic_kernel<<<grid_cap((M * 32 + 255) / 256), 256, 0, st>>>(z, codebook, idx, margin, M, D, K);
      break;
  }
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_vq_gather(const int64_t* idx, const float* codebook, float* out_nchw, int B, int D, int H, int W, int K,
                 void* stream) {
  SD_REQUIRE(B >= 1 && D >= 1 && H >= 1 && W >= 1 && K >= 1, "vq_gather: bad shape");
  SD_REQUIRE(idx && codebook && out_nchw, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  int64_t n = (int64_t)B * D * H * W;
  vq_gather_kernel<<<grid_cap((n + 255) / 256), 256, 0, as_stream(stream)>>>(idx, codebook, out_nchw, B, D, H, W, K);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
