// CUDA-core fused conv -> BN -> LIF (and the three non-spiking tails), fp32 accumulation.
//
// This is the path for the layers that are NOT tensor-core shaped:
//   * real-valued inputs with a tiny contraction (enc.conv1 K=9, den.conv1 K=18, vq.poisson K=16) whose input
//     is identical at every timestep (R/main.py:133, R/snn_model/vq_diffusion.py:198, vae_model.py:56), so
//     the convolution is evaluated ONCE and the LIF runs on a constant current;
//   * the stride-2 / transposed layers of the VQ-VAE (0.06 % of the sampling FLOPs).
// It is also the exact-order fp32 cross-check for the tcgen05 kernel in tests.
//
// Mapping: one thread = one output neuron (b, oy, ox, co) for ALL T timesteps; co is the fastest index in a
// warp so packed weights [tap][ci][co] are read coalesced and the input element is a warp-wide broadcast.
// Spike inputs are read 8 channels (16 B) at a time from the STF planes.
#include <stdlib.h>
#include "common.cuh"

namespace sd {

struct SimtParams {
  sd_conv_desc d;
  const void* in;
  const void* in2;
  const float* w;       // [kh*kw][C_in][C_out]
  const float* scale;
  const float* shift;
  float* v;
  void* out;
  __half* out_sum;
  MemoutCoef coef;
  float in_scalar;      // SD_IN_TOKENS: value of input channel 1
};

template <int TMAX>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtParams p) {
  const sd_conv_desc& d = p.d;
  const int T = d.T;
  const int Cout = d.C_out, Cin = d.C_in;
  const int64_t total = (int64_t)d.B * d.H_out * d.W_out * Cout;
  const StfGeom gin(d.B, d.H_in, d.W_in);
  const StfGeom gout(d.B, d.H_out, d.W_out);
  const int Cin8_0 = c8(d.C_in0), Cin8_1 = c8(Cin - d.C_in0);
  const int Cout8 = c8(Cout);
  const bool const_in = d.in_kind == SD_IN_REAL_CONST;
  const int TA = const_in ? 1 : (d.in_kind == SD_IN_STF ? d.in_T : T);  // timesteps actually convolved

  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int co = (int)(i % Cout);
    int64_t r = i / Cout;
    const int ox = (int)(r % d.W_out); r /= d.W_out;
    const int oy = (int)(r % d.H_out);
    const int b = (int)(r / d.H_out);

    float acc[TMAX];
#pragma unroll
    for (int t = 0; t < TMAX; ++t) acc[t] = 0.f;

    for (int ky = 0; ky < d.kh; ++ky) {
      int iy;
      if (d.transposed) {
        int ny = oy + d.pad - ky;
        if (ny < 0 || ny % d.stride) continue;
        iy = ny / d.stride;
      } else {
        iy = oy * d.stride - d.pad + ky;
      }
      if (iy < 0 || iy >= d.H_in) continue;
      for (int kx = 0; kx < d.kw; ++kx) {
        int ix;
        if (d.transposed) {
          int nx = ox + d.pad - kx;
          if (nx < 0 || nx % d.stride) continue;
          ix = nx / d.stride;
        } else {
          ix = ox * d.stride - d.pad + kx;
        }
        if (ix < 0 || ix >= d.W_in) continue;
        const float* wt = p.w + (int64_t)(ky * d.kw + kx) * Cin * Cout + co;
        if (d.in_kind == SD_IN_STF) {
          const int64_t row = gin.row(b, iy, ix);
          for (int cc = 0; cc < Cin; cc += 8) {
            const bool seg1 = cc >= d.C_in0;
            const __half* base = seg1 ? (const __half*)p.in2 : (const __half*)p.in;
            const int C8s = seg1 ? Cin8_1 : Cin8_0;
            const int cl = seg1 ? cc - d.C_in0 : cc;
            float wv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) wv[j] = (cc + j < Cin) ? wt[(int64_t)(cc + j) * Cout] : 0.f;
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
              if (t < TA) {
                const uint4 raw = *reinterpret_cast<const uint4*>(base + gin.at(t, C8s, cl, row));
                const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 f = __half22float2(h2[j]);
                  acc[t] = fmaf(f.x, wv[2 * j], acc[t]);
                  acc[t] = fmaf(f.y, wv[2 * j + 1], acc[t]);
                }
              }
            }
          }
        } else {
          const float* xin = (const float*)p.in;
          const int64_t plane = (int64_t)d.H_in * d.W_in;
          for (int ci = 0; ci < Cin; ++ci) {
            const float wv = wt[(int64_t)ci * Cout];
#pragma unroll
            for (int t = 0; t < TMAX; ++t) {
              if (t < TA) {
                const float xv = xin[(((int64_t)t * d.B + b) * Cin + ci) * plane + (int64_t)iy * d.W_in + ix];
                acc[t] = fmaf(xv, wv, acc[t]);
              }
            }
          }
        }
      }
    }

    const float sc = p.scale[co], sh = p.shift[co];
    if (d.out_kind == SD_OUT_LIF) {
      const int64_t orow = gout.row(b, oy, ox);
      const int64_t vidx = ((int64_t)(co >> 3) * gout.R_alloc + orow) * 8 + (co & 7);
      float v = p.v ? p.v[vidx] : (d.hard_reset ? d.v_reset : 0.f);
      float cnt = 0.f;
      __half* outp = (__half*)p.out;
#pragma unroll
      for (int t = 0; t < TMAX; ++t) {
        if (t < T) {
          const float x = fmaf(acc[TA == 1 ? 0 : t], sc, sh);
          float h;
          if (d.hard_reset) h = __fadd_rn(v, __fdiv_rn(__fsub_rn(x, __fsub_rn(v, d.v_reset)), d.tau));
          else              h = __fadd_rn(v, __fdiv_rn(__fsub_rn(x, v), d.tau));
          const bool s = h >= d.v_threshold;
          v = d.hard_reset ? (s ? d.v_reset : h) : (s ? __fsub_rn(h, d.v_threshold) : h);
          cnt += s ? 1.f : 0.f;
          outp[gout.at(t, Cout8, co, orow)] = __float2half_rn(s ? 1.f : 0.f);
        }
      }
      if (p.v) p.v[vidx] = v;
      if (p.out_sum) p.out_sum[gout.at(0, Cout8, co, orow)] = __float2half_rn(cnt);
    } else if (d.out_kind == SD_OUT_REAL_SEQ) {
      float* outp = (float*)p.out;
      const int64_t plane = (int64_t)d.H_out * d.W_out;
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < T)
          outp[(((int64_t)t * d.B + b) * Cout + co) * plane + (int64_t)oy * d.W_out + ox] =
              fmaf(acc[TA == 1 ? 0 : t], sc, sh);
    } else if (d.out_kind == SD_OUT_MEMOUT_TANH) {
      float m = 0.f;
#pragma unroll
      for (int t = 0; t < TMAX; ++t)
        if (t < T) m = __fadd_rn(m, __fmul_rn(fmaf(acc[TA == 1 ? 0 : t], sc, sh), p.coef.c[t]));
      ((float*)p.out)[(((int64_t)b * Cout + co) * d.H_out + oy) * d.W_out + ox] = tanhf(m);
    } else {  // SD_OUT_MEAN_T
      float m = 0.f;
      if (TA == 1 && d.in_kind == SD_IN_STF) {
        // input already summed over T: sum_t (conv_t + bias) = conv(sum_t s_t) + T * bias
        m = fmaf(acc[0], sc, __fmul_rn(sh, (float)T));
      } else {
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
          if (t < T) m = __fadd_rn(m, fmaf(acc[TA == 1 ? 0 : t], sc, sh));
      }
      ((float*)p.out)[(((int64_t)b * d.H_out + oy) * d.W_out + ox) * Cout + co] = __fdiv_rn(m, (float)T);
    }
  }
}


// Real-valued input and output, fp32 [T*B, C, H, W] (un-fused layer.Conv2d / ConvTranspose2d, the training-mode forward
// and, as the adjoint convolution, the input gradient): implicit GEMM on CUDA cores.  128 output pixels x 64 output
// channels per block, 8 x 4 outputs per thread; K runs tap-major with 8 input channels per slice, so the window
// geometry is evaluated once per tap and the inner loop has no integer division; the next slice is fetched into
// registers while the current one is multiplied out of shared memory (two buffers, one barrier per slice).  The
// accumulation order per output (tap-major, channel inner, one FMA chain) is the generic kernel's.
__global__ void __launch_bounds__(256) conv_real_tiled_kernel(const SimtParams p) {
  constexpr int BM = 128, BN = 64, BK = 8;
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const sd_conv_desc& d = p.d;
  const int Cin = d.C_in, Cout = d.C_out, taps = d.kh * d.kw;
  const int64_t plane_in = (int64_t)d.H_in * d.W_in, plane_out = (int64_t)d.H_out * d.W_out;
  const int64_t M = (int64_t)d.T * d.B * plane_out;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // load mapping: A pixel ml, channels kl + 2 j (j < 4); B column nl, channels kb + 4 j (j < 2)
  const int ml = tid & 127, kl = tid >> 7;
  const int nl = tid & 63, kb = tid >> 6;
  const int64_t m_ld = m0 + ml;
  const bool m_ok = m_ld < M;
  int oy = 0, ox = 0;
  int64_t img = 0;
  if (m_ok) {
    img = m_ld / plane_out;
    const int pp = (int)(m_ld - img * plane_out);
    oy = pp / d.W_out;
    ox = pp - oy * d.W_out;
  }
  const float* xin = (const float*)p.in + img * Cin * plane_in;
  const bool n_ok = n0 + nl < Cout;
  const int slices_per_tap = (Cin + BK - 1) / BK;
  const int n_slices = taps * slices_per_tap;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // window geometry of this thread's load pixel for the current tap
  int tap = 0, ci0 = 0;
  const float* a_ptr = nullptr;   // x at (iy, ix) of channel 0, or null if the tap falls outside
  auto set_tap = [&](int tp) {
    const int ky = tp / d.kw, kx = tp - ky * d.kw;
    int iy, ix;
    bool ok = m_ok;
    if (d.transposed) {
      const int ny = oy + d.pad - ky, nx = ox + d.pad - kx;
      ok = ok && ny >= 0 && nx >= 0 && (ny % d.stride) == 0 && (nx % d.stride) == 0;
      iy = ny / d.stride; ix = nx / d.stride;
    } else {
      iy = oy * d.stride - d.pad + ky; ix = ox * d.stride - d.pad + kx;
    }
    ok = ok && iy >= 0 && iy < d.H_in && ix >= 0 && ix < d.W_in;
    a_ptr = ok ? xin + (int64_t)iy * d.W_in + ix : nullptr;
  };
  float a_reg[4], b_reg[2];
  auto fetch = [&]() {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int ci = ci0 + kl + 2 * j;
      a_reg[j] = (a_ptr != nullptr && ci < Cin) ? a_ptr[(int64_t)ci * plane_in] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci0 + kb + 4 * j;
      b_reg[j] = (n_ok && ci < Cin) ? p.w[((int64_t)tap * Cin + ci) * Cout + n0 + nl] : 0.f;
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int j = 0; j < 4; ++j) As[buf][kl + 2 * j][ml] = a_reg[j];
#pragma unroll
    for (int j = 0; j < 2; ++j) Bs[buf][kb + 4 * j][nl] = b_reg[j];
  };
  auto advance = [&]() {
    ci0 += BK;
    if (ci0 >= Cin) { ci0 = 0; ++tap; if (tap < taps) set_tap(tap); }
  };
  set_tap(0);
  fetch();
  stage(0);
  advance();
  __syncthreads();
  for (int sidx = 0; sidx < n_slices; ++sidx) {
    const int buf = sidx & 1;
    const bool more = sidx + 1 < n_slices;
    if (more) fetch();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) { stage(buf ^ 1); advance(); }
    __syncthreads();
  }
  float* outp = (float*)p.out;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
    if (m >= M) continue;
    const int64_t im = m / plane_out;
    const int64_t pp = m - im * plane_out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + ty * 4 + j;
      if (co < Cout) outp[(im * Cout + co) * plane_out + pp] = fmaf(acc[i][j], p.scale[co], p.shift[co]);
    }
  }
}

// Real-valued input that is constant over T (enc.conv1, den.conv1, vq.poisson): the convolution is evaluated once
// and the LIF runs on a constant current.  One thread = one output pixel x 8 output channels, so every timestep's
// spikes leave as ONE 16-byte store into the STF plane, consecutive lanes -> consecutive rows (coalesced); the
// 8-channel weight slice is a warp-uniform 2 x float4 load per (tap, ci) from the packed [tap][ci][co] array.
// FAST (hard reset to 0, tau a power of two) and OUT8 (u8 STF8 output) are template parameters and the t loop is rolled:
// the kernel runs once per diffusion step on a cold instruction cache, so its size is part of its latency.
template <bool FAST, bool OUT8>
__global__ void __launch_bounds__(256) conv_real_const_lif_kernel(const SimtParams p) {
  const sd_conv_desc& d = p.d;
  const int T = d.T, Cout = d.C_out, Cin = d.C_in;
  // the whole weight block [kh*kw][C_in][C_out] (a few KB for the real-input layers: K-dim 9 / 18 / 16) is staged in
  // shared memory once per block; threads of a warp share the output-channel chunk, so the reads are broadcasts
  extern __shared__ float sw[];
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < d.kh * d.kw * Cin * Cout; i += blockDim.x) sw[i] = p.w[i];   // weights: not produced by a kernel of the chain
  __syncthreads();
  pdl_wait();                // the tokens / input of this step, and the buffers written below
  const bool tokens = d.in_kind == SD_IN_TOKENS;
  const int64_t* __restrict__ tok = (const int64_t*)p.in;
  const int Cout8 = Cout >> 3;
  const int64_t npix = (int64_t)d.B * d.H_out * d.W_out;
  const int64_t total = npix * Cout8;
  const StfGeom gout(d.B, d.H_out, d.W_out);
  const float* __restrict__ xin = (const float*)p.in;
  const int64_t plane = (int64_t)d.H_in * d.W_in;
  const float inv_tau = 1.0f / d.tau;
  constexpr bool fast = FAST;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int chunk = (int)(i / npix);
    int pix = (int)(i - (int64_t)chunk * npix);
    const int ox = pix % d.W_out; pix /= d.W_out;
    const int oy = pix % d.H_out;
    const int b = pix / d.H_out;
    const int co = chunk << 3;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int ky = 0; ky < d.kh; ++ky) {
      const int iy = oy * d.stride - d.pad + ky;
      if (iy < 0 || iy >= d.H_in) continue;
      for (int kx = 0; kx < d.kw; ++kx) {
        const int ix = ox * d.stride - d.pad + kx;
        if (ix < 0 || ix >= d.W_in) continue;
        const float* wt = sw + (ky * d.kw + kx) * Cin * Cout + co;
        const float* xp = xin + ((int64_t)b * Cin) * plane + (int64_t)iy * d.W_in + ix;
        for (int ci = 0; ci < Cin; ++ci) {
          // SD_IN_TOKENS: channel 0 = the token id as a float, channel 1 = the diffusion time (vq_diffusion.py:195-197)
          const float xv = tokens ? (ci == 0 ? (float)tok[(int64_t)b * plane + (int64_t)iy * d.W_in + ix] : p.in_scalar)
                                  : xp[(int64_t)ci * plane];
          const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * Cout);
          const float4 w1 = *reinterpret_cast<const float4*>(wt + ci * Cout + 4);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]);
          acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]);
          acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      }
    }
    const int64_t orow = gout.row(b, oy, ox);
    const int64_t vidx = ((int64_t)chunk * gout.R_alloc + orow) * 8;
    float x[8], v[8], cnt[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      x[j] = fmaf(acc[j], p.scale[co + j], p.shift[co + j]);
      v[j] = p.v ? p.v[vidx + j] : (d.hard_reset ? d.v_reset : 0.f);
      cnt[j] = 0.f;
    }
    __half* outp = (__half*)p.out;
    constexpr bool out8 = OUT8;
    // A rolled loop on purpose: the kernel runs once per diffusion step on a cold instruction cache, and fully unrolled
    // the T = 16 body is ~40 KB of straight-line code whose fetch alone took ~35 us of the 46 us launch (time grew with
    // TMAX, not with the work).  Nothing in the body is indexed by t at compile time.
#pragma unroll 1
    for (int t = 0; t < T; ++t) {
      {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float h;
          bool s;
          if constexpr (fast) {
            h = fmaf(__fsub_rn(x[j], v[j]), inv_tau, v[j]);
            s = h >= d.v_threshold;
            v[j] = s ? 0.f : h;
          } else {
            const float dv = d.hard_reset ? __fsub_rn(x[j], __fsub_rn(v[j], d.v_reset)) : __fsub_rn(x[j], v[j]);
            h = __fadd_rn(v[j], __fdiv_rn(dv, d.tau));
            s = h >= d.v_threshold;
            v[j] = d.hard_reset ? (s ? d.v_reset : h) : (s ? __fsub_rn(h, d.v_threshold) : h);
          }
          cnt[j] += s ? 1.f : 0.f;
          if constexpr (out8) {       // STF8: one byte per spike, 8 channels = half of a 16-byte row
            const uint32_t bit = (s ? 1u : 0u) << (8 * (j & 3));
            if (j & 3) pk[j >> 2] |= bit; else pk[j >> 2] = bit;
          } else {
            const uint32_t bits = s ? 0x3C00u : 0u;
            if (j & 1) pk[j >> 1] |= bits << 16; else pk[j >> 1] = bits;
          }
        }
        if constexpr (out8) {
          const int64_t plane = (int64_t)(Cout8 >> 1) * gout.R_alloc * 16;
          uint8_t* o = (uint8_t*)p.out + (int64_t)(t * 2) * plane + ((int64_t)(chunk >> 1) * gout.R_alloc + orow) * 16 +
                       (chunk & 1) * 8;
          *reinterpret_cast<uint2*>(o) = make_uint2(pk[0], pk[1]);
          *reinterpret_cast<uint2*>(o + plane) = make_uint2(pk[0] << 7, pk[1] << 7);      // the 128*s plane
        } else {
          *reinterpret_cast<uint4*>(outp + gout.at(t, Cout8, co, orow)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    if (p.v) {
#pragma unroll
      for (int j = 0; j < 8; ++j) p.v[vidx + j] = v[j];
    }
    if (p.out_sum) {
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const __half2 h2 = __floats2half2_rn(cnt[j], cnt[j + 1]);
        pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      *reinterpret_cast<uint4*>(p.out_sum + gout.at(0, Cout8, co, orow)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}


// Spike-input (STF) layers of the VQ-VAE that are not 3x3/stride-1 (enc.conv2 s2, enc.conv3 1x1, dec.convT1/2 s2,
// dec.convT3 + memout + tanh).  One thread = one output pixel x 8 output channels x all T timesteps: the 8x8 weight
// block of an (input chunk, tap) is a warp-uniform load held in registers and reused by the T timesteps (64*T FMAs per
// 16 weight loads + T 16-byte spike loads).  For stride-2 transposed convolutions pixels are enumerated
// phase-major (oy%2, ox%2 outermost) so that a warp shares one set of valid taps and does not diverge.
template <int TMAX>
__global__ void __launch_bounds__(128) conv_spike8_kernel(const SimtParams p) {
  const sd_conv_desc& d = p.d;
  const int T = d.T, Cout = d.C_out, Cin = d.C_in;
  const int G8 = (Cout + 7) >> 3;
  const int64_t npix = (int64_t)d.B * d.H_out * d.W_out;
  const int64_t total = npix * G8;
  const StfGeom gin(d.B, d.H_in, d.W_in);
  const StfGeom gout(d.B, d.H_out, d.W_out);
  const int Cin8_0 = c8(d.C_in0), Cin8_1 = c8(Cin - d.C_in0);
  const int Cout8 = c8(Cout);
  const bool phase_major = d.transposed && d.stride == 2 && (d.H_out % 2 == 0) && (d.W_out % 2 == 0);
  const int Hh = d.H_out >> 1, Wh = d.W_out >> 1;
  const float inv_tau = 1.0f / d.tau;
  int tau_exp;
  const bool fast = frexpf(d.tau, &tau_exp) == 0.5f && d.hard_reset && d.v_reset == 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int grp = (int)(i / npix);
    int64_t pix = i - (int64_t)grp * npix;
    int ox, oy, b;
    if (phase_major) {
      const int64_t per = (int64_t)d.B * Hh * Wh;
      const int ph = (int)(pix / per);
      int64_t q = pix - ph * per;
      const int qx = (int)(q % Wh); q /= Wh;
      const int qy = (int)(q % Hh);
      b = (int)(q / Hh);
      oy = qy * 2 + (ph >> 1);
      ox = qx * 2 + (ph & 1);
    } else {
      ox = (int)(pix % d.W_out); pix /= d.W_out;
      oy = (int)(pix % d.H_out);
      b = (int)(pix / d.H_out);
    }
    const int co = grp << 3;
    float acc[TMAX][8];
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[t][j] = 0.f;

    for (int ky = 0; ky < d.kh; ++ky) {
      int iy;
      if (d.transposed) {
        const int ny = oy + d.pad - ky;
        if (ny < 0 || ny % d.stride) continue;
        iy = ny / d.stride;
      } else {
        iy = oy * d.stride - d.pad + ky;
      }
      if (iy < 0 || iy >= d.H_in) continue;
      for (int kx = 0; kx < d.kw; ++kx) {
        int ix;
        if (d.transposed) {
          const int nx = ox + d.pad - kx;
          if (nx < 0 || nx % d.stride) continue;
          ix = nx / d.stride;
        } else {
          ix = ox * d.stride - d.pad + kx;
        }
        if (ix < 0 || ix >= d.W_in) continue;
        const float* wt = p.w + (int64_t)(ky * d.kw + kx) * Cin * Cout + co;
        const int64_t row = gin.row(b, iy, ix);
        for (int cc = 0; cc < Cin; cc += 8) {
          const bool seg1 = cc >= d.C_in0;
          const __half* base = seg1 ? (const __half*)p.in2 : (const __half*)p.in;
          const int C8s = seg1 ? Cin8_1 : Cin8_0;
          const int cl = seg1 ? cc - d.C_in0 : cc;
          float w[8][8];
#pragma unroll
          for (int ci = 0; ci < 8; ++ci) {
            if (cc + ci < Cin) {
              if ((Cout & 7) == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(wt + (int64_t)(cc + ci) * Cout));
                const float4 c = __ldg(reinterpret_cast<const float4*>(wt + (int64_t)(cc + ci) * Cout + 4));
                w[ci][0] = a.x; w[ci][1] = a.y; w[ci][2] = a.z; w[ci][3] = a.w;
                w[ci][4] = c.x; w[ci][5] = c.y; w[ci][6] = c.z; w[ci][7] = c.w;
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) w[ci][j] = (co + j < Cout) ? __ldg(wt + (int64_t)(cc + ci) * Cout + j) : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) w[ci][j] = 0.f;
            }
          }
#pragma unroll
          for (int t = 0; t < TMAX; ++t) {
            if (t < T) {
              const uint4 raw = *reinterpret_cast<const uint4*>(base + gin.at(t, C8s, cl, row));
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(h2[q]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  acc[t][j] = fmaf(f.x, w[2 * q][j], acc[t][j]);
                  acc[t][j] = fmaf(f.y, w[2 * q + 1][j], acc[t][j]);
                }
              }
            }
          }
        }
      }
    }

    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sc[j] = (co + j < Cout) ? p.scale[co + j] : 0.f;
      sh[j] = (co + j < Cout) ? p.shift[co + j] : 0.f;
    }
    if (d.out_kind == SD_OUT_LIF) {
      const int64_t orow = gout.row(b, oy, ox);
      const int64_t vidx = ((int64_t)grp * gout.R_alloc + orow) * 8;
      float v[8], cnt[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        v[j] = p.v ? p.v[vidx + j] : (d.hard_reset ? d.v_reset : 0.f);
        cnt[j] = 0.f;
      }
      __half* outp = (__half*)p.out;
#pragma unroll
      for (int t = 0; t < TMAX; ++t) {
        if (t < T) {
          uint32_t pk[4];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x = fmaf(acc[t][j], sc[j], sh[j]);
            float h;
            bool s;
            if (fast) {
              h = fmaf(__fsub_rn(x, v[j]), inv_tau, v[j]);
              s = h >= d.v_threshold;
              v[j] = s ? 0.f : h;
            } else {
              const float dv = d.hard_reset ? __fsub_rn(x, __fsub_rn(v[j], d.v_reset)) : __fsub_rn(x, v[j]);
              h = __fadd_rn(v[j], __fdiv_rn(dv, d.tau));
              s = h >= d.v_threshold;
              v[j] = d.hard_reset ? (s ? d.v_reset : h) : (s ? __fsub_rn(h, d.v_threshold) : h);
            }
            s = s && (co + j < Cout);
            cnt[j] += s ? 1.f : 0.f;
            const uint32_t bits = s ? 0x3C00u : 0u;
            if (j & 1) pk[j >> 1] |= bits << 16; else pk[j >> 1] = bits;
          }
          *reinterpret_cast<uint4*>(outp + gout.at(t, Cout8, co, orow)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      if (p.v) {
#pragma unroll
        for (int j = 0; j < 8; ++j) p.v[vidx + j] = v[j];
      }
      if (p.out_sum) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const __half2 h2 = __floats2half2_rn(cnt[j], cnt[j + 1]);
          pk[j >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        *reinterpret_cast<uint4*>(p.out_sum + gout.at(0, Cout8, co, orow)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    } else {  // SD_OUT_MEMOUT_TANH: sum_t coef[t] * (conv_t * scale + shift) -> tanh, fp32 [B, C_out, H_out, W_out]
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (co + j < Cout) {
          float m = 0.f;
#pragma unroll
          for (int t = 0; t < TMAX; ++t)
            if (t < T) m = __fadd_rn(m, __fmul_rn(fmaf(acc[t][j], sc[j], sh[j]), p.coef.c[t]));
          ((float*)p.out)[(((int64_t)b * Cout + co + j) * d.H_out + oy) * d.W_out + ox] = tanhf(m);
        }
      }
    }
  }
}


// Spike-input layer with a handful of output channels and the memout + tanh tail (dec.convT3: 32 -> image channels,
// R/snn_model/vae_model.py:147,185).  One thread = one output pixel x all (<= NCO) output channels x all T timesteps;
// the whole weight tensor [tap][ci][co] sits in shared memory and is read as warp-uniform broadcasts.  The FMA order
// per output (tap, input channel) is the one of conv_spike8_kernel, so both give the same bits.
template <int TMAX, int NCO>
__global__ void __launch_bounds__(128) conv_spike_fewout_memout_kernel(const SimtParams p) {
  extern __shared__ float w_s[];   // [taps][Cin][NCO]
  const sd_conv_desc& d = p.d;
  const int T = d.T, Cout = d.C_out, Cin = d.C_in, taps = d.kh * d.kw;
  for (int i = threadIdx.x; i < taps * Cin * NCO; i += blockDim.x) {
    const int j = i % NCO, r = i / NCO;
    w_s[i] = j < Cout ? p.w[(int64_t)r * Cout + j] : 0.f;
  }
  __syncthreads();
  const StfGeom gin(d.B, d.H_in, d.W_in);
  const int Cin8_0 = c8(d.C_in0), Cin8_1 = c8(Cin - d.C_in0);
  const int64_t npix = (int64_t)d.B * d.H_out * d.W_out;
  for (int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; pix < npix; pix += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(pix % d.W_out);
    const int64_t r0 = pix / d.W_out;
    const int oy = (int)(r0 % d.H_out);
    const int b = (int)(r0 / d.H_out);
    float acc[TMAX][NCO];
#pragma unroll
    for (int t = 0; t < TMAX; ++t)
#pragma unroll
      for (int j = 0; j < NCO; ++j) acc[t][j] = 0.f;
    for (int ky = 0; ky < d.kh; ++ky) {
      int iy;
      if (d.transposed) {
        const int ny = oy + d.pad - ky;
        if (ny < 0 || ny % d.stride) continue;
        iy = ny / d.stride;
      } else {
        iy = oy * d.stride - d.pad + ky;
      }
      if (iy < 0 || iy >= d.H_in) continue;
      for (int kx = 0; kx < d.kw; ++kx) {
        int ix;
        if (d.transposed) {
          const int nx = ox + d.pad - kx;
          if (nx < 0 || nx % d.stride) continue;
          ix = nx / d.stride;
        } else {
          ix = ox * d.stride - d.pad + kx;
        }
        if (ix < 0 || ix >= d.W_in) continue;
        const float* wt = w_s + (ky * d.kw + kx) * Cin * NCO;
        const int64_t row = gin.row(b, iy, ix);
        for (int cc = 0; cc < Cin; cc += 8) {
          const bool seg1 = cc >= d.C_in0;
          const __half* base = seg1 ? (const __half*)p.in2 : (const __half*)p.in;
          const int C8s = seg1 ? Cin8_1 : Cin8_0;
          const int cl = seg1 ? cc - d.C_in0 : cc;
          float w[8][NCO];
#pragma unroll
          for (int ci = 0; ci < 8; ++ci)
#pragma unroll
            for (int j = 0; j < NCO; ++j) w[ci][j] = (cc + ci < Cin) ? wt[(cc + ci) * NCO + j] : 0.f;
#pragma unroll
          for (int t = 0; t < TMAX; ++t) {
            if (t < T) {
              const uint4 raw = *reinterpret_cast<const uint4*>(base + gin.at(t, C8s, cl, row));
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 f = __half22float2(h2[q]);
#pragma unroll
                for (int j = 0; j < NCO; ++j) {
                  acc[t][j] = fmaf(f.x, w[2 * q][j], acc[t][j]);
                  acc[t][j] = fmaf(f.y, w[2 * q + 1][j], acc[t][j]);
                }
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NCO; ++j) {
      if (j < Cout) {
        const float sc = p.scale[j], sh = p.shift[j];
        float m = 0.f;
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
          if (t < T) m = __fadd_rn(m, __fmul_rn(fmaf(acc[t][j], sc, sh), p.coef.c[t]));
        ((float*)p.out)[(((int64_t)b * Cout + j) * d.H_out + oy) * d.W_out + ox] = tanhf(m);
      }
    }
  }
}

// w (reference layout) -> [tap][ci][co]
__global__ void pack_simt_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int kh,
                                 int kw, int transposed) {
  const int64_t total = (int64_t)kh * kw * Cin * Cout;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int co = (int)(i % Cout);
    int64_t r = i / Cout;
    int ci = (int)(r % Cin);
    int tap = (int)(r / Cin);
    int ky = tap / kw, kx = tap % kw;
    int64_t src = transposed ? ((((int64_t)ci * Cout + co) * kh + ky) * kw + kx)
                             : ((((int64_t)co * Cin + ci) * kh + ky) * kw + kx);
    out[i] = w[src];
  }
}

int validate_conv_desc(const sd_conv_desc* d) {
  SD_REQUIRE(d != nullptr, "null descriptor");
  SD_REQUIRE(d->T >= 1 && d->T <= SD_MAX_T, "conv: T=%d out of range [1,%d]", d->T, SD_MAX_T);
  SD_REQUIRE(d->B >= 1 && d->C_in >= 1 && d->C_out >= 1 && d->H_in >= 1 && d->W_in >= 1 && d->H_out >= 1 &&
                 d->W_out >= 1, "conv: non-positive dimension");
  SD_REQUIRE(d->kh >= 1 && d->kw >= 1 && d->stride >= 1 && d->pad >= 0, "conv: bad kernel/stride/padding");
  SD_REQUIRE(d->in_kind >= SD_IN_REAL_CONST && d->in_kind <= SD_IN_TOKENS, "conv: bad in_kind %d", d->in_kind);
  if (d->in_kind == SD_IN_TOKENS) SD_REQUIRE(d->C_in == 2 && !d->transposed, "conv: SD_IN_TOKENS is the denoiser's 2-channel input");
  SD_REQUIRE(d->out_kind >= SD_OUT_LIF && d->out_kind <= SD_OUT_CURRENT_SEQ, "conv: bad out_kind %d", d->out_kind);
  if (d->out_kind == SD_OUT_CURRENT_SEQ) SD_REQUIRE(d->in_kind == SD_IN_STF8 && d->C_out % 16 == 0, "conv: SD_OUT_CURRENT_SEQ is the tensor-core int8 path's output (STF8 input, C_out %% 16 == 0)");
  if (d->in_kind == SD_IN_STF8) SD_REQUIRE(d->in_T == d->T && d->C_in0 == d->C_in && d->C_in % 16 == 0,
                                           "conv: STF8 input needs in_T == T, one segment, C_in %% 16 == 0");
  if (d->out_kind == SD_OUT_LIF8) SD_REQUIRE(d->C_out % 16 == 0, "conv: STF8 output needs C_out %% 16 == 0");
  if (d->in_kind == SD_IN_STF) {
    SD_REQUIRE(d->in_T == d->T || d->in_T == 1, "conv: in_T must be T or 1");
    SD_REQUIRE(d->C_in0 >= 1 && d->C_in0 <= d->C_in, "conv: C_in0 out of range");
    SD_REQUIRE(d->C_in0 == d->C_in || d->C_in0 % 8 == 0, "conv: concat boundary must be a multiple of 8");
  }
  if (d->out_kind == SD_OUT_LIF || d->out_kind == SD_OUT_LIF8)
    SD_REQUIRE(d->tau > 1.0f, "LIFNode requires tau > 1, got %f", (double)d->tau);
  // output size consistent with torch's formulas (nn.Conv2d / nn.ConvTranspose2d docs)
  if (!d->transposed) {
    SD_REQUIRE(d->H_out == (d->H_in + 2 * d->pad - d->kh) / d->stride + 1 &&
                   d->W_out == (d->W_in + 2 * d->pad - d->kw) / d->stride + 1, "conv: output size mismatch");
  } else {
    int h0 = (d->H_in - 1) * d->stride - 2 * d->pad + d->kh, w0 = (d->W_in - 1) * d->stride - 2 * d->pad + d->kw;
    SD_REQUIRE(d->H_out >= h0 && d->H_out < h0 + d->stride && d->W_out >= w0 && d->W_out < w0 + d->stride,
               "conv_transpose: output size mismatch (output_padding must be < stride)");
  }
  return SD_OK;
}

}  // namespace sd

using namespace sd;

extern "C" {

int64_t sd_conv_weight_bytes_simt(const sd_conv_desc* d) {
  if (!d) return 0;
  return (int64_t)d->kh * d->kw * d->C_in * d->C_out * (int64_t)sizeof(float);
}

int sd_conv_pack_weights_simt(const sd_conv_desc* d, const float* w, void* packed, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  SD_REQUIRE(w && packed, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  int64_t n = (int64_t)d->kh * d->kw * d->C_in * d->C_out;
  int64_t blocks = (n + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  pack_simt_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(w, (float*)packed, d->C_out, d->C_in, d->kh,
                                                                    d->kw, d->transposed);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_conv_lif_simt(const sd_conv_desc* d, const sd_conv_args* a, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  SD_REQUIRE(a && a->in && a->weights && a->scale && a->shift && a->out, "null pointer argument");
  if (d->in_kind == SD_IN_STF && d->C_in0 < d->C_in) SD_REQUIRE(a->in2 != nullptr, "conv: in2 missing for concat");
  if (d->out_kind == SD_OUT_MEMOUT_TANH) SD_REQUIRE(a->memout_coef_host != nullptr, "conv: memout_coef_host missing");
  SD_DEVICE_OR_RETURN();
  SimtParams p;
  p.d = *d;
  p.in = a->in; p.in2 = a->in2; p.w = (const float*)a->weights; p.scale = a->scale; p.shift = a->shift;
  p.v = a->v; p.out = a->out; p.out_sum = (__half*)a->out_sum;
  p.in_scalar = a->in_scalar;
  for (int t = 0; t < SD_MAX_T; ++t)
    p.coef.c[t] = (a->memout_coef_host && t < d->T) ? a->memout_coef_host[t] : 0.f;
  int64_t n = (int64_t)d->B * d->H_out * d->W_out * d->C_out;
  int64_t blocks = (n + 255) / 256;
  int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = as_stream(stream);
  if (d->in_kind == SD_IN_STF8 ||
      (d->out_kind == SD_OUT_LIF8 && !((d->in_kind == SD_IN_REAL_CONST || d->in_kind == SD_IN_TOKENS) && !d->transposed))) {
    set_error("conv_simt: the u8 spike format (STF8) is produced by the constant-input layer and consumed by "
              "sd_conv_lif_tc (nsplit = 3) only");
    return SD_ERR_UNSUPPORTED;
  }
  const size_t w_smem = (size_t)d->kh * d->kw * d->C_in * d->C_out * sizeof(float);
  if (d->in_kind == SD_IN_TOKENS && !((d->out_kind == SD_OUT_LIF || d->out_kind == SD_OUT_LIF8) && d->C_out % 8 == 0 &&
                                      w_smem <= 48 * 1024)) {
    set_error("conv_simt: SD_IN_TOKENS needs a LIF output with C_out %% 8 == 0 and a weight block of at most 48 KB");
    return SD_ERR_UNSUPPORTED;
  }
  if ((d->in_kind == SD_IN_REAL_CONST || d->in_kind == SD_IN_TOKENS) && (d->out_kind == SD_OUT_LIF || d->out_kind == SD_OUT_LIF8) &&
      !d->transposed && d->C_out % 8 == 0 && w_smem <= 48 * 1024) {
    int64_t n8 = (int64_t)d->B * d->H_out * d->W_out * (d->C_out / 8);
    int64_t bl = (n8 + 255) / 256;
    if (bl > cap) bl = cap;
    int tau_exp;
    const bool fast = frexpf(d->tau, &tau_exp) == 0.5f && d->hard_reset && d->v_reset == 0.f;
    const bool out8 = d->out_kind == SD_OUT_LIF8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)bl);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = w_smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = sd_pdl_enabled(d->concurrent) ? 1 : 0;
    cudaError_t lrc;
    if (fast && out8) lrc = cudaLaunchKernelEx(&cfg, conv_real_const_lif_kernel<true, true>, p);
    else if (fast) lrc = cudaLaunchKernelEx(&cfg, conv_real_const_lif_kernel<true, false>, p);
    else if (out8) lrc = cudaLaunchKernelEx(&cfg, conv_real_const_lif_kernel<false, true>, p);
    else lrc = cudaLaunchKernelEx(&cfg, conv_real_const_lif_kernel<false, false>, p);
    SD_CUDA(lrc);
    SD_LAUNCH_CHECK();
    return SD_OK;
  }
  static const bool force_generic = getenv("SD_SIMT_GENERIC") != nullptr;   // debugging aid: exact-order generic kernel
  if (!force_generic && d->in_kind == SD_IN_REAL_SEQ && d->out_kind == SD_OUT_REAL_SEQ) {
    const int64_t M = (int64_t)d->T * d->B * d->H_out * d->W_out;
    dim3 grid((unsigned)((M + 127) / 128), (unsigned)((d->C_out + 63) / 64));
    conv_real_tiled_kernel<<<grid, 256, 0, st>>>(p);
    SD_LAUNCH_CHECK();
    return SD_OK;
  }
  if (!force_generic && d->in_kind == SD_IN_STF && d->in_T == d->T && d->T <= 16 && d->out_kind == SD_OUT_MEMOUT_TANH &&
      d->C_out <= 4 && (int64_t)d->kh * d->kw * d->C_in * 4 * sizeof(float) <= 48 * 1024) {
    const int64_t npix = (int64_t)d->B * d->H_out * d->W_out;
    int64_t bl = (npix + 127) / 128;
    if (bl > cap * 2) bl = cap * 2;
    const int nco = d->C_out <= 1 ? 1 : (d->C_out <= 2 ? 2 : 4);
    const size_t smem = (size_t)d->kh * d->kw * d->C_in * nco * sizeof(float);
#define SD_FEWOUT(TM)                                                                                          \
  do {                                                                                                         \
    if (nco == 1) conv_spike_fewout_memout_kernel<TM, 1><<<(unsigned)bl, 128, smem, st>>>(p);                  \
    else if (nco == 2) conv_spike_fewout_memout_kernel<TM, 2><<<(unsigned)bl, 128, smem, st>>>(p);             \
    else conv_spike_fewout_memout_kernel<TM, 4><<<(unsigned)bl, 128, smem, st>>>(p);                           \
  } while (0)
    if (d->T <= 4) SD_FEWOUT(4); else if (d->T <= 8) SD_FEWOUT(8); else SD_FEWOUT(16);
#undef SD_FEWOUT
    SD_LAUNCH_CHECK();
    return SD_OK;
  }
  if (!force_generic && d->in_kind == SD_IN_STF && d->in_T == d->T && d->T <= 8 &&
      (d->out_kind == SD_OUT_MEMOUT_TANH || (d->out_kind == SD_OUT_LIF && d->C_out % 8 == 0))) {
    int64_t n8 = (int64_t)d->B * d->H_out * d->W_out * ((d->C_out + 7) / 8);
    int64_t bl = (n8 + 127) / 128;
    if (bl > cap * 2) bl = cap * 2;
    if (d->T <= 4) conv_spike8_kernel<4><<<(unsigned)bl, 128, 0, st>>>(p);
    else conv_spike8_kernel<8><<<(unsigned)bl, 128, 0, st>>>(p);
    SD_LAUNCH_CHECK();
    return SD_OK;
  }
  // TMAX bounds both the accumulator array and the unrolled epilogue; pick the smallest that fits T.
  if (d->T <= 4) conv_simt_kernel<4><<<(unsigned)blocks, 256, 0, st>>>(p);
  else if (d->T <= 8) conv_simt_kernel<8><<<(unsigned)blocks, 256, 0, st>>>(p);
  else if (d->T <= 16) conv_simt_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(p);
  else conv_simt_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(p);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
