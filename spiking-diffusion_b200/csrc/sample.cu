// Absorbing-state diffusion sampling step with a torch-compatible counter-based Philox stream.
//
// torch draws `rand_like` (R/snn_model/vq_diffusion.py:118) and the `exponential_` inside
// Categorical.sample -> multinomial (vq_diffusion.py:136-138) from Philox4x32-10 through
// distribution_elementwise_grid_stride_kernel (TORCH/include/ATen/native/cuda/DistributionTemplates.h:64-87):
// thread `idx` initialises curand_init(seed, subsequence = idx, offset) and per grid-stride round produces one
// curand_uniform4, whose component ii goes to element li = idx + tpg*(4*round + ii), tpg = 256*grid,
// grid = min(sm_count * (max_threads_per_sm/256), ceil(numel/256)).  Because Philox is counter-based the value
// of ANY element can be computed directly:  counter = (offset/4 + round, 0, idx, 0), key = seed.
// That is what these kernels do, so the fused step kernel needs no intermediate [tokens, K] tensor of
// exponentials, and a batch shard can draw the values of its global element indices.
//
// The fused step kernel is HBM-bound on the logits: 4*K B read per token (+ 9 B of token state), one warp per
// token, logits read with 128-bit loads.
#include "common.cuh"

namespace sd {

struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __host__ __device__ static inline uint4 round10(uint4 c, uint2 k) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
#ifdef __CUDA_ARCH__
      uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
      uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
#else
      uint64_t p0 = (uint64_t)kM0 * c.x, p1 = (uint64_t)kM1 * c.z;
      uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
      c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
      k.x += kW0;
      k.y += kW1;
    }
    return c;
  }
};

// curand's _curand_uniform: (0, 1]
__device__ __forceinline__ float u32_to_uniform(uint32_t x) {
  return __fadd_rn(__fmul_rn((float)x, 2.3283064365386963e-10f), 2.3283064365386963e-10f / 2.0f);
}

struct PhiloxCall {
  uint64_t seed;
  uint64_t offset4;  // offset / 4
  uint64_t tpg;      // threads per grid of the torch launch being reproduced
};

// raw 32-bit draw of global element li
__device__ __forceinline__ uint32_t philox_element(const PhiloxCall& c, uint64_t li) {
  const uint64_t idx = li % c.tpg;
  const uint64_t q = li / c.tpg;
  const uint64_t round = q >> 2;
  const uint32_t ii = (uint32_t)(q & 3);
  const uint64_t ctr = c.offset4 + round;
  uint4 r = Philox::round10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)idx, (uint32_t)(idx >> 32)),
                            make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32)));
  return ii == 0 ? r.x : (ii == 1 ? r.y : (ii == 2 ? r.z : r.w));
}

// Same draw when the caller already knows li = q * tpg + idx (no 64-bit division per element).
__device__ __forceinline__ uint32_t philox_element_qr(const PhiloxCall& c, uint64_t q, uint32_t idx) {
  const uint64_t ctr = c.offset4 + (q >> 2);
  const uint32_t ii = (uint32_t)(q & 3);
  uint4 r = Philox::round10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u),
                            make_uint2((uint32_t)c.seed, (uint32_t)(c.seed >> 32)));
  return ii == 0 ? r.x : (ii == 1 ? r.y : (ii == 2 ? r.z : r.w));
}

// torch.rand: value = rand*1 + 0, then the (0,1] -> [0,1) bound flip   (DistributionTemplates.h:493-503)
__device__ __forceinline__ float torch_uniform(uint32_t raw) {
  float v = u32_to_uniform(raw);
  return v == 1.0f ? 0.0f : v;
}
// Tensor.exponential_(1): -log(u) with the u >= 1 - eps/2 guard        (TransformationHelper.h:129-146).
// at::log<float> on the device is the fast __logf approximation (TORCH/include/ATen/NumericUtils.h:150-157), so
// the same intrinsic is used here: measured bit-identical to torch on the B200 (tests/test_gpu_sampling.py).
__device__ __forceinline__ float torch_exponential(uint32_t raw) {
  float v = u32_to_uniform(raw);
  const float eps = 1.1920928955078125e-07f;
  float lg = (v >= 1.0f - eps / 2.0f) ? (-eps / 2.0f) : __logf(v);
  return __fmul_rn(-1.0f, lg);  // (-1 / lambda) * log with lambda = 1
}

static void torch_policy(int64_t numel, uint64_t* tpg, uint64_t* inc) {
  // calc_execution_policy, DistributionTemplates.h:50-62 (block 256, unroll 4, 4 offsets per curand call)
  uint64_t n = (uint64_t)numel;
  uint64_t grid = (n + 255) / 256;
  uint64_t cap = (uint64_t)sm_count() * (uint64_t)(max_threads_per_sm() / 256);
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  *tpg = grid * 256;
  *inc = ((n - 1) / (256 * grid * 4) + 1) * 4;
}

template <bool EXPO>
__global__ void philox_fill_kernel(float* __restrict__ out, int64_t numel, int64_t index_base, PhiloxCall c) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t raw = philox_element(c, (uint64_t)(index_base + i));
    out[i] = EXPO ? torch_exponential(raw) : torch_uniform(raw);
  }
}

// One warp per token; each lane keeps ceil(K/32) logits in registers (K <= 1024).
// Templated on the register-array size so that K = 128 compiles to a compact loop (the 32-wide unroll is ~100 KB of
// code and thrashes the instruction cache when only 4 of its 32 iterations execute).  VEC: K is a multiple of 128 and
// lane l holds elements 128 j + 4 l + e (128-bit loads); otherwise any K, lane l holds elements 32 j + l, and the
// missing tail elements are -inf (probability 0, never drawn).
constexpr int kMaxPerLaneAll = 32;

template <int kMaxPerLane, bool VEC>
__global__ void __launch_bounds__(256) sample_step_kernel(const float* __restrict__ logits, int64_t* __restrict__ x_t,
                                                          uint8_t* __restrict__ unmasked, int64_t* __restrict__ x0_hat,
                                                          int64_t n_tokens, int K, float inv_t, float inv_temp,
                                                          PhiloxCall cu, PhiloxCall ce, int64_t token_base,
                                                          const uint64_t* __restrict__ rng_dev) {
  pdl_launch_dependents();   // the next step's conv1 may stage its weights while this grid samples
  pdl_wait();
  if (rng_dev != nullptr) {
    // graph-replayable form: (seed, base offset) live in device memory; cu/ce carry offsets relative to the base
    const uint64_t seed = rng_dev[0], base4 = rng_dev[1] >> 2;
    cu.seed = ce.seed = seed;
    cu.offset4 += base4;
    ce.offset4 += base4;
    token_base += (int64_t)rng_dev[2];   // where this call's tokens start in the global stream (chunked large batches)
  }
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int per_lane = VEC ? K >> 5 : (K + 31) >> 5;
  for (int64_t tok = warp_global; tok < n_tokens; tok += n_warps) {
    const int64_t gtok = token_base + tok;
    // ---- where to unmask (vq_diffusion.py:118-124) ----
    const float u = torch_uniform(philox_element(cu, (uint64_t)gtok));
    const bool was_unmasked = unmasked[tok] != 0;
    const bool change = (u < inv_t) && !was_unmasked;
    // ---- Categorical(logits / temp): normalise by logsumexp, softmax -> probs (categorical.py:78) ----
    // element k of lane: k = lane*4 + 128*j + e  (128-bit loads, consecutive lanes -> consecutive 16 B)
    float l[kMaxPerLane];
    float mx = -INFINITY;
    if (VEC) {
      const float4* row = reinterpret_cast<const float4*>(logits + tok * (int64_t)K);
#pragma unroll
      for (int j = 0; j < kMaxPerLane / 4; ++j) {
        if (j * 4 < per_lane) {
          float4 q = row[j * 32 + lane];
          // torch's CUDA `tensor / python_float` multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu)
          l[4 * j + 0] = __fmul_rn(q.x, inv_temp); l[4 * j + 1] = __fmul_rn(q.y, inv_temp);
          l[4 * j + 2] = __fmul_rn(q.z, inv_temp); l[4 * j + 3] = __fmul_rn(q.w, inv_temp);
          mx = fmaxf(fmaxf(fmaxf(l[4 * j], l[4 * j + 1]), fmaxf(l[4 * j + 2], l[4 * j + 3])), mx);
        }
      }
    } else {
      const float* row = logits + tok * (int64_t)K;
#pragma unroll
      for (int j = 0; j < kMaxPerLane; ++j) {
        if (j < per_lane) {
          const int k = j * 32 + lane;
          l[j] = k < K ? __fmul_rn(row[k], inv_temp) : -INFINITY;
          mx = fmaxf(mx, l[j]);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j)
      if (j < per_lane) sum = __fadd_rn(sum, expf(__fsub_rn(l[j], mx)));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, off));
    const float lse = __fadd_rn(logf(sum), mx);
    // normalised logits n = l - lse; probs = softmax(n) = exp(n - max n) / sum exp(n - max n)
    float nmx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j)
      if (j < per_lane) { l[j] = __fsub_rn(l[j], lse); nmx = fmaxf(nmx, l[j]); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) nmx = fmaxf(nmx, __shfl_xor_sync(0xffffffffu, nmx, off));
    float s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j)
      if (j < per_lane) { l[j] = expf(__fsub_rn(l[j], nmx)); s2 = __fadd_rn(s2, l[j]); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s2 = __fadd_rn(s2, __shfl_xor_sync(0xffffffffu, s2, off));
    // ---- multinomial(1) = argmax(probs / Exp(1)), first index on ties ----
    // element index li = gtok*K + k = q*tpg + r: one 64-bit division per token, then r walks with at most one wrap
    // per step (K <= tpg because tpg >= min(numel, 256) and K >= 128 divides numel)
    const uint64_t li0 = (uint64_t)gtok * (uint64_t)K;
    const uint64_t q0 = li0 / ce.tpg;
    const uint32_t r0 = (uint32_t)(li0 - q0 * ce.tpg);
    const uint32_t tpg32 = (uint32_t)ce.tpg;
    float best = -INFINITY;
    int besti = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
      if (j < per_lane) {
        const int k = VEC ? (j >> 2) * 128 + lane * 4 + (j & 3) : j * 32 + lane;
        if (!VEC && k >= K) continue;
        const float p = __fdiv_rn(l[j], s2);
        uint32_t r = r0 + (uint32_t)k;
        uint64_t qq = q0;
        while (r >= tpg32) { r -= tpg32; ++qq; }
        const float q = torch_exponential(philox_element_qr(ce, qq, r));
        const float val = __fdiv_rn(p, q);
        if (val > best || (val == best && k < besti)) { best = val; besti = k; }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, off);
      int oi = __shfl_xor_sync(0xffffffffu, besti, off);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    if (lane == 0) {
      if (x0_hat) x0_hat[tok] = besti;
      if (change) { x_t[tok] = besti; unmasked[tok] = 1; }
    }
  }
}

__global__ void denoiser_input_kernel(const int64_t* __restrict__ x_t, float* __restrict__ out, int B, int HW, float t) {
  const int64_t total = (int64_t)B * HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t b = i / HW, p = i % HW;
    out[(b * 2 + 0) * HW + p] = (float)x_t[i];
    out[(b * 2 + 1) * HW + p] = t;
  }
}

static inline unsigned grid_cap(int64_t blocks) {
  int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

static int philox_fill(bool expo, float* out, int64_t numel, uint64_t seed, uint64_t offset, int64_t index_base,
                       int64_t numel_global, uint64_t* inc_out, void* stream) {
  SD_REQUIRE(numel >= 0 && index_base >= 0 && numel_global >= index_base + numel, "philox: bad range");
  SD_REQUIRE(offset % 4 == 0, "philox: generator offset must be a multiple of 4");
  SD_DEVICE_OR_RETURN();
  if (numel_global == 0) { if (inc_out) *inc_out = 0; return SD_OK; }
  PhiloxCall c;
  uint64_t inc;
  torch_policy(numel_global, &c.tpg, &inc);
  c.seed = seed; c.offset4 = offset / 4;
  if (inc_out) *inc_out = inc;
  if (numel == 0) return SD_OK;
  SD_REQUIRE(out != nullptr, "null pointer argument");
  if (expo) philox_fill_kernel<true><<<grid_cap((numel + 255) / 256), 256, 0, as_stream(stream)>>>(out, numel, index_base, c);
  else philox_fill_kernel<false><<<grid_cap((numel + 255) / 256), 256, 0, as_stream(stream)>>>(out, numel, index_base, c);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // namespace sd

using namespace sd;

extern "C" {

int sd_philox_offset_increment(int64_t numel_global, uint64_t* inc_out) {
  SD_REQUIRE(numel_global >= 0 && inc_out, "bad argument");
  SD_DEVICE_OR_RETURN();
  if (numel_global == 0) { *inc_out = 0; return SD_OK; }
  uint64_t tpg;
  torch_policy(numel_global, &tpg, inc_out);
  return SD_OK;
}

int sd_philox_uniform(float* out, int64_t numel, uint64_t seed, uint64_t offset, int64_t index_base,
                      int64_t numel_global, uint64_t* inc_out, void* stream) {
  return philox_fill(false, out, numel, seed, offset, index_base, numel_global, inc_out, stream);
}

int sd_philox_exponential(float* out, int64_t numel, uint64_t seed, uint64_t offset, int64_t index_base,
                          int64_t numel_global, uint64_t* inc_out, void* stream) {
  return philox_fill(true, out, numel, seed, offset, index_base, numel_global, inc_out, stream);
}

static int sample_step_impl(const float* logits, int64_t* x_t, uint8_t* unmasked, int64_t* x0_hat, int64_t n_tokens,
                            int K, int t, float temp, uint64_t seed, uint64_t offset_uniform,
                            uint64_t offset_exponential, int64_t token_base, int64_t n_tokens_global,
                            const uint64_t* rng_dev, void* stream) {
  SD_REQUIRE(n_tokens >= 0 && token_base >= 0 && n_tokens_global >= token_base + n_tokens, "sample_step: bad token range");
  SD_REQUIRE(K >= 1 && K <= 32 * kMaxPerLaneAll, "sample_step: K=%d must be in [1, %d]", K, 32 * kMaxPerLaneAll);
  SD_REQUIRE(t >= 1, "sample_step: t must be >= 1");
  SD_REQUIRE(temp > 0.f, "sample_step: temperature must be positive");
  SD_REQUIRE(offset_uniform % 4 == 0 && offset_exponential % 4 == 0, "sample_step: offsets must be multiples of 4");
  if (n_tokens == 0) return SD_OK;
  SD_REQUIRE(logits && x_t && unmasked, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  PhiloxCall cu, ce;
  uint64_t inc;
  torch_policy(n_tokens_global, &cu.tpg, &inc);
  torch_policy(n_tokens_global * (int64_t)K, &ce.tpg, &inc);
  cu.seed = ce.seed = seed;
  cu.offset4 = offset_uniform / 4;
  ce.offset4 = offset_exponential / 4;
  const float inv_t = 1.0f / (float)t;  // `1 / t_mask.float()`  (vq_diffusion.py:118)
  int64_t blocks = (n_tokens * 32 + 255) / 256;
#define SD_SAMPLE_LAUNCH(PL, VEC)                                                                                  \
  sample_step_kernel<PL, VEC><<<grid_cap(blocks), 256, 0, as_stream(stream)>>>(logits, x_t, unmasked, x0_hat, n_tokens, K, \
                                                                               inv_t, 1.0f / temp, cu, ce, token_base, rng_dev)
  if (K % 128 == 0 && (((uintptr_t)logits) & 15) == 0) {
    if (K <= 128) SD_SAMPLE_LAUNCH(4, true);
    else if (K <= 256) SD_SAMPLE_LAUNCH(8, true);
    else if (K <= 512) SD_SAMPLE_LAUNCH(16, true);
    else SD_SAMPLE_LAUNCH(32, true);
  } else {   // any codebook size (the reference's --codebook_size is free, R/main.py:58)
    if (K <= 128) SD_SAMPLE_LAUNCH(4, false);
    else if (K <= 256) SD_SAMPLE_LAUNCH(8, false);
    else if (K <= 512) SD_SAMPLE_LAUNCH(16, false);
    else SD_SAMPLE_LAUNCH(32, false);
  }
#undef SD_SAMPLE_LAUNCH
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_sample_step(const float* logits, int64_t* x_t, uint8_t* unmasked, int64_t* x0_hat, int64_t n_tokens, int K,
                   int t, float temp, uint64_t seed, uint64_t offset_uniform, uint64_t offset_exponential,
                   int64_t token_base, int64_t n_tokens_global, void* stream) {
  return sample_step_impl(logits, x_t, unmasked, x0_hat, n_tokens, K, t, temp, seed, offset_uniform,
                          offset_exponential, token_base, n_tokens_global, nullptr, stream);
}

int sd_sample_step_dev(const float* logits, int64_t* x_t, uint8_t* unmasked, int64_t* x0_hat, int64_t n_tokens, int K,
                       int t, float temp, const uint64_t* rng_dev, uint64_t rel_offset_uniform,
                       uint64_t rel_offset_exponential, int64_t token_base, int64_t n_tokens_global, void* stream) {
  SD_REQUIRE(rng_dev != nullptr, "sample_step_dev: rng_dev is null");
  return sample_step_impl(logits, x_t, unmasked, x0_hat, n_tokens, K, t, temp, 0, rel_offset_uniform,
                          rel_offset_exponential, token_base, n_tokens_global, rng_dev, stream);
}

int sd_denoiser_input(const int64_t* x_t, float* out, int B, int H, int W, int t, void* stream) {
  SD_REQUIRE(B >= 1 && H >= 1 && W >= 1, "denoiser_input: bad shape");
  SD_REQUIRE(x_t && out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  int64_t n = (int64_t)B * H * W;
  denoiser_input_kernel<<<grid_cap((n + 255) / 256), 256, 0, as_stream(stream)>>>(x_t, out, B, H * W, (float)t);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

}  // extern "C"
