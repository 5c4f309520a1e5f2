// Stand-alone multi-step LIF neuron, MembraneOutputLayer, layout conversions and small element-wise ops.
//
// sd_lif_forward is the drop-in for LIFNode.multi_step_forward (eval branch,
// SJ/activation_based/neuron.py:799-809 and the three sibling kernels): fp32 current in, fp32 {0,1} spikes
// out, membrane state read and written in place.  It is HBM-bound: 8 B per neuron-timestep (x in, spike
// out) + 8 B per neuron (v in/out).  One thread owns 4 neurons (128-bit loads/stores) and walks T in
// registers; loads for up to 4 timesteps are issued before use so that each thread keeps 4 independent
// 16-byte requests in flight.  Arithmetic uses explicit round-to-nearest intrinsics in the reference's
// operation order so that no FMA contraction changes a result bit.
#include <stdarg.h>
#include <mutex>
#include "common.cuh"

namespace sd {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Device facts are cached PER DEVICE (a process may drive several GPUs: the current device at call time decides).
static std::mutex g_dev_mu;
constexpr int kMaxDevices = 64;
struct DevInfo { int state = -1; int sm_count = 0, max_thr = 0, cc_major = 0, cc_minor = 0; };   // state: -1 unknown, 0 ok
static DevInfo g_dev[kMaxDevices];
static int g_no_device = -1;   // -1 unknown, 0 some device exists, else error code

int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev >= 0 && dev < kMaxDevices ? dev : 0;
}

int check_device() {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_no_device < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
      cudaGetLastError();
      g_no_device = SD_ERR_NO_DEVICE;
      set_error("no CUDA device available (libsd_b200 has no CPU fallback): %s", cudaGetErrorString(e));
      return g_no_device;
    }
    g_no_device = 0;
  }
  if (g_no_device != 0) {
    set_error("no sm_100 CUDA device available (libsd_b200 has no CPU fallback)");
    return g_no_device;
  }
  DevInfo& d = g_dev[current_device_index()];
  if (d.state < 0) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, current_device_index()) != cudaSuccess || p.major != 10) {
      d.state = SD_ERR_NO_DEVICE;
    } else {
      d.sm_count = p.multiProcessorCount;
      d.max_thr = p.maxThreadsPerMultiProcessor;
      d.cc_major = p.major;
      d.cc_minor = p.minor;
      d.state = SD_OK;
    }
  }
  if (d.state != SD_OK) set_error("device is not compute capability 10.x (libsd_b200 is built for sm_100a only)");
  return d.state;
}
int sm_count() { return g_dev[current_device_index()].sm_count; }
int max_threads_per_sm() { return g_dev[current_device_index()].max_thr; }

// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w));
}

struct LifParams {
  float tau, v_th, v_reset, decay_keep;  // decay_keep = (float)(1 - 1/tau) computed in double
};

template <bool HARD, bool DECAY>
__device__ __forceinline__ float lif_step(float x, float& v, const LifParams& p, float& h_out) {
  float h;
  if (HARD) {
    if (DECAY) {  // v + (x - (v - v_reset)) / tau                    neuron.py:805
      h = __fadd_rn(v, __fdiv_rn(__fsub_rn(x, __fsub_rn(v, p.v_reset)), p.tau));
    } else {      // v - (v - v_reset) / tau + x                      neuron.py:831
      h = __fadd_rn(__fsub_rn(v, __fdiv_rn(__fsub_rn(v, p.v_reset), p.tau)), x);
    }
  } else {
    if (DECAY) {  // v + (x - v) / tau                                neuron.py:858
      h = __fadd_rn(v, __fdiv_rn(__fsub_rn(x, v), p.tau));
    } else {      // v * (1 - 1/tau) + x                              neuron.py:884
      h = __fadd_rn(__fmul_rn(v, p.decay_keep), x);
    }
  }
  h_out = h;
  float s = (h >= p.v_th) ? 1.0f : 0.0f;
  if (HARD) {     // v_reset * spike + (1 - spike) * v                neuron.py:807
    v = __fadd_rn(__fmul_rn(p.v_reset, s), __fmul_rn(__fsub_rn(1.0f, s), h));
  } else {        // v - spike * v_threshold                          neuron.py:860
    v = __fsub_rn(h, __fmul_rn(s, p.v_th));
  }
  return s;
}

template <bool HARD, bool DECAY, bool WRITE_H>
__global__ void __launch_bounds__(256) lif_vec4_kernel(const float4* __restrict__ x, float4* __restrict__ v,
                                                       float4* __restrict__ spk, float4* __restrict__ hseq, int T,
                                                       int64_t N4, LifParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 vv = v[i];
    int t = 0;
    for (; t + 4 <= T; t += 4) {
      float4 xs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xs[u] = ld_stream(x + (int64_t)(t + u) * N4 + i);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 s, h;
        s.x = lif_step<HARD, DECAY>(xs[u].x, vv.x, p, h.x);
        s.y = lif_step<HARD, DECAY>(xs[u].y, vv.y, p, h.y);
        s.z = lif_step<HARD, DECAY>(xs[u].z, vv.z, p, h.z);
        s.w = lif_step<HARD, DECAY>(xs[u].w, vv.w, p, h.w);
        st_stream(spk + (int64_t)(t + u) * N4 + i, s);
        if (WRITE_H) st_stream(hseq + (int64_t)(t + u) * N4 + i, h);
      }
    }
    for (; t < T; ++t) {
      float4 xx = ld_stream(x + (int64_t)t * N4 + i);
      float4 s, h;
      s.x = lif_step<HARD, DECAY>(xx.x, vv.x, p, h.x);
      s.y = lif_step<HARD, DECAY>(xx.y, vv.y, p, h.y);
      s.z = lif_step<HARD, DECAY>(xx.z, vv.z, p, h.z);
      s.w = lif_step<HARD, DECAY>(xx.w, vv.w, p, h.w);
      st_stream(spk + (int64_t)t * N4 + i, s);
      if (WRITE_H) st_stream(hseq + (int64_t)t * N4 + i, h);
    }
    v[i] = vv;
  }
}

// Scalar variant for N % 4 != 0 or unaligned pointers (ragged shapes).
template <bool HARD, bool DECAY, bool WRITE_H>
__global__ void __launch_bounds__(256) lif_scalar_kernel(const float* __restrict__ x, float* __restrict__ v,
                                                         float* __restrict__ spk, float* __restrict__ hseq, int T,
                                                         int64_t N, LifParams p) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float vv = v[i];
    for (int t = 0; t < T; ++t) {
      float h;
      float s = lif_step<HARD, DECAY>(x[(int64_t)t * N + i], vv, p, h);
      spk[(int64_t)t * N + i] = s;
      if (WRITE_H) hseq[(int64_t)t * N + i] = h;
    }
    v[i] = vv;
  }
}

template <bool HARD, bool DECAY, bool WRITE_H>
static int launch_lif(const float* x, float* v, float* spk, float* h, int T, int64_t N, LifParams p,
                      cudaStream_t st) {
  const bool vec = (N % 4 == 0) && (((uintptr_t)x | (uintptr_t)v | (uintptr_t)spk | (uintptr_t)h) % 16 == 0);
  const int64_t work = vec ? N / 4 : N;
  int64_t blocks = (work + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;  // 8 resident CTAs of 256 threads per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (vec) {
    lif_vec4_kernel<HARD, DECAY, WRITE_H><<<(unsigned)blocks, 256, 0, st>>>(
        (const float4*)x, (float4*)v, (float4*)spk, (float4*)h, T, N / 4, p);
  } else {
    lif_scalar_kernel<HARD, DECAY, WRITE_H><<<(unsigned)blocks, 256, 0, st>>>(x, v, spk, h, T, N, p);
  }
  SD_LAUNCH_CHECK();
  return SD_OK;
}

// ---------------------------------------------------------------------------------------------------
// Surrogate-gradient BPTT of the LIF recurrence (training branch: SJ/activation_based/neuron.py:210-258 with the
// ATan surrogate, surrogate.py:663-678).  Same reverse-time recurrence as the reference's generated kernel
// (SURVEY.md Appendix A): one thread per neuron walks t = T-1 .. 0 carrying dL/dh.  HBM-bound: 12 B per
// neuron-timestep (grad_spike in, h in, grad_x out).
struct LifBwdParams {
  float inv_tau, v_th, v_reset, alpha;
  int hard_reset, decay_input, detach_reset;
};

__global__ void __launch_bounds__(256) lif_backward_kernel(const float* __restrict__ grad_spike,
                                                           const float* __restrict__ grad_v_last,
                                                           const float* __restrict__ h_seq, float* __restrict__ grad_x,
                                                           float* __restrict__ grad_v_init, int T, int64_t N,
                                                           LifBwdParams p) {
  const float keep = 1.0f - p.inv_tau;          // d h[t+1] / d v[t]
  const float half_pi_alpha = 1.57079632679489661923f * p.alpha;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float grad_v = grad_v_last ? grad_v_last[i] : 0.f;   // dL/dv[T-1] from outside (the stored state)
    for (int t = T - 1; t >= 0; --t) {
      const float h = h_seq[(int64_t)t * N + i];
      const float over = h - p.v_th;
      const float s = over >= 0.f ? 1.f : 0.f;
      const float ax = half_pi_alpha * over;
      const float g = p.alpha / 2.0f / (1.0f + ax * ax);          // d s / d h   (surrogate.py:663-665)
      float dv_dh;                                                 // d v[t] / d h[t]
      if (p.hard_reset) dv_dh = p.detach_reset ? (1.f - s) : (1.f - s) + (p.v_reset - h) * g;
      else              dv_dh = p.detach_reset ? 1.f : 1.f - p.v_th * g;
      const float grad_h = grad_v * dv_dh + grad_spike[(int64_t)t * N + i] * g;
      grad_x[(int64_t)t * N + i] = p.decay_input ? grad_h * p.inv_tau : grad_h;
      grad_v = grad_h * keep;                                      // flows to v[t-1]
    }
    if (grad_v_init) grad_v_init[i] = grad_v;
  }
}

// ---------------------------------------------------------------------------------------------------
__global__ void memout_kernel(const float* __restrict__ x, float* __restrict__ out, int T, int64_t N,
                              MemoutCoef coef, int apply_tanh) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc = __fadd_rn(acc, __fmul_rn(x[(int64_t)t * N + i], coef.c[t]));
    out[i] = apply_tanh ? tanhf(acc) : acc;
  }
}

__global__ void to_uint8_kernel(const float* __restrict__ p, uint8_t* __restrict__ out, int64_t N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float v = __fadd_rn(p[i], 0.5f);
    v = fminf(fmaxf(v, 0.f), 1.f);
    out[i] = (uint8_t)(__fmul_rn(v, 255.f));  // truncation, as numpy's astype(uint8)
  }
}

__global__ void stf_from_nchw_kernel(const float* __restrict__ x, __half* __restrict__ stf, int T, int B, int C,
                                     int H, int W) {
  StfGeom g(B, H, W);
  const int C8 = c8(C);
  const int64_t total = (int64_t)T * B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xw = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C); r /= C;
    int b = (int)(r % B);
    int t = (int)(r / B);
    stf[g.at(t, C8, c, g.row(b, y, xw))] = __float2half_rn(x[i]);
  }
}

__global__ void stf_to_nchw_kernel(const __half* __restrict__ stf, float* __restrict__ x, int T, int B, int C, int H,
                                   int W) {
  StfGeom g(B, H, W);
  const int C8 = c8(C);
  const int64_t total = (int64_t)T * B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xw = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C); r /= C;
    int b = (int)(r % B);
    int t = (int)(r / B);
    x[i] = __half2float(stf[g.at(t, C8, c, g.row(b, y, xw))]);
  }
}

// STF8 (u8 spikes for the kind::i8 layers): [T][2][C/16][R_alloc][16]; plane 0 of a timestep holds s, plane 1 holds 128*s
__global__ void stf8_from_nchw_kernel(const float* __restrict__ x, uint8_t* __restrict__ stf, int T, int B, int C, int H,
                                      int W) {
  StfGeom g(B, H, W);
  const int C16 = C / 16;
  const int64_t total = (int64_t)T * B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xw = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C); r /= C;
    int b = (int)(r % B);
    int t = (int)(r / B);
    const uint8_t s = x[i] != 0.f ? 1 : 0;
    const int64_t row = g.row(b, y, xw);
    stf[(((int64_t)(t * 2) * C16 + (c >> 4)) * g.R_alloc + row) * 16 + (c & 15)] = s;
    stf[(((int64_t)(t * 2 + 1) * C16 + (c >> 4)) * g.R_alloc + row) * 16 + (c & 15)] = (uint8_t)(s << 7);
  }
}

// checks both planes: returns s from the s plane, and NaN where the 128*s plane disagrees (a format error shows up in tests)
__global__ void stf8_to_nchw_kernel(const uint8_t* __restrict__ stf, float* __restrict__ x, int T, int B, int C, int H,
                                    int W) {
  StfGeom g(B, H, W);
  const int C16 = C / 16;
  const int64_t total = (int64_t)T * B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xw = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C); r /= C;
    int b = (int)(r % B);
    int t = (int)(r / B);
    const int64_t row = g.row(b, y, xw);
    const uint8_t s = stf[(((int64_t)(t * 2) * C16 + (c >> 4)) * g.R_alloc + row) * 16 + (c & 15)];
    const uint8_t s128 = stf[(((int64_t)(t * 2 + 1) * C16 + (c >> 4)) * g.R_alloc + row) * 16 + (c & 15)];
    x[i] = (s <= 1 && s128 == (uint8_t)(s << 7)) ? (float)s : __int_as_float(0x7fc00000);
  }
}

// Zero-insertion 2x upsampling of a spike tensor: out[t, c, b, 2y, 2x] = in[t, c, b, y, x], every other pixel 0.
// A stride-2 ConvTranspose2d(k=3, p=1, output_padding=1) is exactly a stride-1 3x3 convolution (pad 1, flipped taps)
// of this tensor, which puts the decoder on the tcgen05 kernel.  One thread = one 16-byte (8-channel) element.
__global__ void __launch_bounds__(256) stf_upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                             int TC8, int B, int H, int W) {
  const StfGeom gi(B, H, W), go(B, 2 * H, 2 * W);
  const int64_t rows_out = (int64_t)B * go.P;
  const int64_t total = (int64_t)TC8 * rows_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / rows_out, r = i - plane * rows_out;
    const int b = (int)(r / go.P);
    const int pp = (int)(r - (int64_t)b * go.P);
    const int oy = pp / go.W, ox = pp - oy * go.W;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (((oy | ox) & 1) == 0) v = in[plane * gi.R_alloc + gi.row(b, oy >> 1, ox >> 1)];
    out[plane * go.R_alloc + go.G + r] = v;
  }
}

// out[t, c, b, y, x] = in[t, c, b, 2y, 2x]: the outputs of a stride-2 3x3 convolution (pad 1) are the even positions of
// the stride-1 convolution, which runs on the tcgen05 kernel.
__global__ void __launch_bounds__(256) stf_subsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out,
                                                              int TC8, int B, int H, int W, int Ho, int Wo) {
  const StfGeom gi(B, H, W), go(B, Ho, Wo);
  const int64_t rows_out = (int64_t)B * go.P;
  const int64_t total = (int64_t)TC8 * rows_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t plane = i / rows_out, r = i - plane * rows_out;
    const int b = (int)(r / go.P);
    const int pp = (int)(r - (int64_t)b * go.P);
    const int oy = pp / go.W, ox = pp - oy * go.W;
    out[plane * go.R_alloc + go.G + r] = in[plane * gi.R_alloc + gi.row(b, 2 * oy, 2 * ox)];
  }
}

// y[n, c, i] = x[n, c, i] * scale[c] + shift[c]   (un-fused eval-mode BatchNorm2d, SJ/activation_based/layer.py:458-465)
__global__ void channel_affine_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                      const float* __restrict__ shift, float* __restrict__ out, int64_t total, int C,
                                      int64_t HW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    out[i] = fmaf(x[i], scale[c], shift[c]);
  }
}

// membrane state: reference layout fp32 [B, C, H, W] <-> planar fused layout [C/8][R_alloc][8]
__global__ void state_convert_kernel(const float* __restrict__ src, float* __restrict__ dst, int B, int C, int H, int W,
                                     int to_planar) {
  StfGeom g(B, H, W);
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int xw = (int)(i % W);
    int64_t r = i / W;
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C);
    int b = (int)(r / C);
    const int64_t pl = ((int64_t)(c >> 3) * g.R_alloc + g.row(b, y, xw)) * 8 + (c & 7);
    if (to_planar) dst[pl] = src[i]; else dst[i] = src[pl];
  }
}

static inline unsigned grid_for(int64_t n, int threads = 256) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (unsigned)b;
}

}  // namespace sd

using namespace sd;

extern "C" {

const char* sd_last_error(void) { return sd::g_err; }
int sd_version(void) { return 1; }

int sd_device_info(int* sms, int* max_thr, int* cc_major, int* cc_minor) {
  SD_DEVICE_OR_RETURN();
  const DevInfo& d = g_dev[current_device_index()];
  if (sms) *sms = d.sm_count;
  if (max_thr) *max_thr = d.max_thr;
  if (cc_major) *cc_major = d.cc_major;
  if (cc_minor) *cc_minor = d.cc_minor;
  return SD_OK;
}

int64_t sd_stf_guard(int W) { return stf_guard(W); }
int64_t sd_stf_rows(int B, int H, int W) { return stf_rows(B, H, W); }
int64_t sd_stf_bytes(int T, int B, int C, int H, int W) {
  return (int64_t)T * c8(C) * stf_rows(B, H, W) * 8 * (int64_t)sizeof(__half);
}

int sd_lif_forward(const float* x_seq, float* v, float* spike_seq, float* h_seq, int T, int64_t N, float tau,
                   float v_threshold, float v_reset, int hard_reset, int decay_input, void* stream) {
  // assert isinstance(tau, float) and tau > 1.   SJ/activation_based/neuron.py:707
  SD_REQUIRE(tau > 1.0f, "LIFNode requires tau > 1, got %f", (double)tau);
  SD_REQUIRE(T >= 0 && N >= 0, "negative size T=%d N=%lld", T, (long long)N);
  if (T == 0 || N == 0) return SD_OK;
  SD_REQUIRE(x_seq && v && spike_seq, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  LifParams p{tau, v_threshold, v_reset, (float)(1.0 - 1.0 / (double)tau)};
  cudaStream_t st = as_stream(stream);
#define SD_LIF_DISPATCH(HARD, DECAY)                                                         \
  return h_seq ? launch_lif<HARD, DECAY, true>(x_seq, v, spike_seq, h_seq, T, N, p, st)     \
               : launch_lif<HARD, DECAY, false>(x_seq, v, spike_seq, nullptr, T, N, p, st)
  if (hard_reset) {
    if (decay_input) { SD_LIF_DISPATCH(true, true); } else { SD_LIF_DISPATCH(true, false); }
  } else {
    if (decay_input) { SD_LIF_DISPATCH(false, true); } else { SD_LIF_DISPATCH(false, false); }
  }
#undef SD_LIF_DISPATCH
}

int sd_lif_backward(const float* grad_spike_seq, const float* grad_v_last, const float* h_seq, float* grad_x_seq,
                    float* grad_v_init, int T, int64_t N, float tau, float v_threshold, float v_reset, int hard_reset,
                    int decay_input, int detach_reset, float alpha, void* stream) {
  SD_REQUIRE(tau > 1.0f, "LIFNode requires tau > 1, got %f", (double)tau);
  SD_REQUIRE(T >= 0 && N >= 0, "negative size");
  if (T == 0 || N == 0) return SD_OK;
  SD_REQUIRE(grad_spike_seq && h_seq && grad_x_seq, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  LifBwdParams p{1.0f / tau, v_threshold, v_reset, alpha, hard_reset, decay_input, detach_reset};
  lif_backward_kernel<<<grid_for(N), 256, 0, as_stream(stream)>>>(grad_spike_seq, grad_v_last, h_seq, grad_x_seq,
                                                                  grad_v_init, T, N, p);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_memout(const float* x, float* out, const float* coef_host, int T, int64_t N, int apply_tanh, void* stream) {
  SD_REQUIRE(T >= 1 && T <= SD_MAX_T && N >= 0, "memout: bad T=%d or N=%lld", T, (long long)N);
  SD_REQUIRE(coef_host != nullptr, "memout: coef_host is null");
  if (N == 0) return SD_OK;
  SD_REQUIRE(x && out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  MemoutCoef hc;
  for (int t = 0; t < SD_MAX_T; ++t) hc.c[t] = t < T ? coef_host[t] : 0.f;
  memout_kernel<<<grid_for(N), 256, 0, as_stream(stream)>>>(x, out, T, N, hc, apply_tanh);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_to_uint8(const float* pred, uint8_t* out, int64_t N, void* stream) {
  SD_REQUIRE(N >= 0, "negative size");
  if (N == 0) return SD_OK;
  SD_REQUIRE(pred && out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  to_uint8_kernel<<<grid_for(N), 256, 0, as_stream(stream)>>>(pred, out, N);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_channel_affine(const float* x, const float* scale, const float* shift, float* out, int64_t n_outer, int C,
                      int64_t HW, void* stream) {
  SD_REQUIRE(n_outer >= 0 && C >= 1 && HW >= 0, "channel_affine: bad shape");
  const int64_t total = n_outer * C * HW;
  if (total == 0) return SD_OK;
  SD_REQUIRE(x && scale && shift && out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  channel_affine_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(x, scale, shift, out, total, C, HW);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_state_convert(const float* src, float* dst, int B, int C, int H, int W, int to_planar, void* stream) {
  SD_REQUIRE(B >= 1 && C >= 1 && H >= 1 && W >= 1, "state_convert: bad shape");
  SD_REQUIRE(src && dst, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  state_convert_kernel<<<grid_for((int64_t)B * C * H * W), 256, 0, as_stream(stream)>>>(src, dst, B, C, H, W, to_planar);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf_from_nchw(const float* x, void* stf, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 1 && H >= 1 && W >= 1, "stf_from_nchw: bad shape");
  SD_REQUIRE(x && stf, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  SD_CUDA(cudaMemsetAsync(stf, 0, (size_t)sd_stf_bytes(T, B, C, H, W), st));
  int64_t n = (int64_t)T * B * C * H * W;
  stf_from_nchw_kernel<<<grid_for(n), 256, 0, st>>>(x, (__half*)stf, T, B, C, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf8_from_nchw(const float* x, void* stf, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 16 && C % 16 == 0 && H >= 1 && W >= 1, "stf8_from_nchw: bad shape (C must be a multiple of 16)");
  SD_REQUIRE(x && stf, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  SD_CUDA(cudaMemsetAsync(stf, 0, (size_t)sd_stf_bytes(T, B, C, H, W), st));   // same byte count as the fp16 format
  int64_t n = (int64_t)T * B * C * H * W;
  stf8_from_nchw_kernel<<<grid_for(n), 256, 0, st>>>(x, (uint8_t*)stf, T, B, C, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf8_to_nchw(const void* stf, float* x, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 16 && C % 16 == 0 && H >= 1 && W >= 1, "stf8_to_nchw: bad shape (C must be a multiple of 16)");
  SD_REQUIRE(x && stf, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  int64_t n = (int64_t)T * B * C * H * W;
  stf8_to_nchw_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>((const uint8_t*)stf, x, T, B, C, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf_upsample2x(const void* in, void* out, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 1 && H >= 1 && W >= 1, "stf_upsample2x: bad shape");
  SD_REQUIRE(in && out && in != out, "stf_upsample2x: null or aliased pointer argument");
  SD_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "stf_upsample2x: buffers must be 16-byte aligned");
  SD_DEVICE_OR_RETURN();
  const int TC8 = T * c8(C);
  const int64_t n = (int64_t)TC8 * B * 4 * H * W;
  stf_upsample2x_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>((const uint4*)in, (uint4*)out, TC8, B, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf_subsample2x(const void* in, void* out, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 1 && H >= 1 && W >= 1, "stf_subsample2x: bad shape");
  SD_REQUIRE(in && out && in != out, "stf_subsample2x: null or aliased pointer argument");
  SD_REQUIRE((((uintptr_t)in | (uintptr_t)out) & 15) == 0, "stf_subsample2x: buffers must be 16-byte aligned");
  SD_DEVICE_OR_RETURN();
  const int TC8 = T * c8(C), Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const int64_t n = (int64_t)TC8 * B * Ho * Wo;
  stf_subsample2x_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>((const uint4*)in, (uint4*)out, TC8, B, H, W, Ho, Wo);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_stf_to_nchw(const void* stf, float* x, int T, int B, int C, int H, int W, void* stream) {
  SD_REQUIRE(T >= 1 && B >= 1 && C >= 1 && H >= 1 && W >= 1, "stf_to_nchw: bad shape");
  SD_REQUIRE(x && stf, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  int64_t n = (int64_t)T * B * C * H * W;
  stf_to_nchw_kernel<<<grid_for(n), 256, 0, as_stream(stream)>>>((const __half*)stf, x, T, B, C, H, W);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
