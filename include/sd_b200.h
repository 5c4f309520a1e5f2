/*
 * sd_b200.h -- C-ABI of the B200-native Spiking-Diffusion hot path (libsd_b200.so).
 *
 * Every entry point replaces one piece of the reference's PyTorch/SpikingJelly path; the reference
 * interface each one stands in for is cited as file:line with the prefixes of SURVEY.md section 0:
 *   R/  = Spiking-Diffusion-release/            SJ/ = inside R/spikingjelly.zip
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types in any signature.
 *   - Return value: 0 = OK, non-zero = error (SD_ERR_*); sd_last_error() gives a thread-local message.
 *     No exception crosses the boundary.
 *   - Every compute call takes an explicit stream (a cudaStream_t passed as void*) and only ENQUEUES
 *     work on it; nothing synchronises.  All compute calls are CUDA-graph capturable.
 *   - The caller owns every buffer.  The library never allocates or frees device memory; sizes of
 *     internal-format buffers are obtained from the sd_*_bytes / sd_stf_* queries below.
 *   - Pointers are device pointers to contiguous memory on the current device, 16-byte aligned.
 *   - Thread-safety: "one stream, one caller"; the only global state is a per-device attribute cache
 *     guarded by a mutex.
 *   - There is no CPU fallback: on a machine without an sm_100 device every compute call returns
 *     SD_ERR_NO_DEVICE.
 *
 * Internal activation format ("spike tile format", STF), used between fused layers
 *   A spike tensor that is logically [T, B, C, H, W] is stored as fp16
 *       [T][C/8][R_alloc][8]          (8 channels = 16 bytes innermost)
 *   where rows enumerate the pixels densely, P = H*W rows per image,
 *       row(b, y, x) = G + b*P + y*W + x,
 *   G = sd_stf_guard(W) zero guard rows before and after, R_alloc = sd_stf_rows(B, H, W) (rounded up to a
 *   multiple of 128 rows + 2 guards).  Guard rows and the tail are ALWAYS ZERO: buffers are zero-filled
 *   once by the caller and kernels only ever write valid rows.  With this layout a 3x3/stride-1/pad-1
 *   convolution is nine row-shifted GEMMs over the same shared-memory tile (shift = dy*W + dx rows); the
 *   rows whose shifted source would wrap across an image border (y == 0 for dy = -1, x == W-1 for dx = +1,
 *   ...) are excluded per tap with the tcgen05.mma disable-output-lane mask, so no padding is stored and
 *   every row of an M tile is useful work.
 */
#ifndef SD_B200_H_
#define SD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SD_OK 0
#define SD_ERR_INVALID 1      /* bad argument (mirrors the reference's ValueError / assert sites) */
#define SD_ERR_CUDA 2         /* a CUDA runtime call failed; see sd_last_error() */
#define SD_ERR_NO_DEVICE 3    /* no CUDA device of compute capability 10.x */
#define SD_ERR_UNSUPPORTED 4  /* shape outside what the kernel was built for */
#define SD_MAX_T 32            /* largest supported number of SNN timesteps */

/* ---- library --------------------------------------------------------------------------------- */
const char* sd_last_error(void);
int sd_version(void);                      /* ABI version, currently 1 */
int sd_device_info(int* sm_count, int* max_threads_per_sm, int* cc_major, int* cc_minor);

/* ---- STF geometry ---------------------------------------------------------------------------- */
int64_t sd_stf_guard(int W);                       /* leading/trailing guard rows */
int64_t sd_stf_rows(int B, int H, int W);          /* R_alloc */
int64_t sd_stf_bytes(int T, int B, int C, int H, int W); /* bytes of one STF tensor (C rounded up to 8) */

/* fp32 [T,B,C,H,W] (reference layout, SJ/activation_based/layer.py:164-173) <-> STF */
int sd_stf_from_nchw(const float* x, void* stf, int T, int B, int C, int H, int W, void* stream);
int sd_stf_to_nchw(const void* stf, float* x, int T, int B, int C, int H, int W, void* stream);
/* STF8: the u8 spike format of the kind::i8 layers, [T][2][C/16][R_alloc][16] (rows as in STF; per timestep the planes
 * of s in {0,1} and of 128*s in {0,128}; same byte count as STF, sd_stf_bytes).  Conversions from / to the reference's
 * fp32 [T, B, C, H, W] spike tensors, for API boundaries and tests (precedent for packed spikes:
 * SJ/activation_based/tensor_cache.py:13-90).  C must be a multiple of 16. */
int sd_stf8_from_nchw(const float* x, void* stf8, int T, int B, int C, int H, int W, void* stream);
int sd_stf8_to_nchw(const void* stf8, float* x, int T, int B, int C, int H, int W, void* stream);

/* Zero-insertion 2x upsampling of an STF tensor: out (T, B, C, 2H, 2W) with out[.., 2y, 2x] = in[.., y, x] and zeros
 * elsewhere.  ConvTranspose2d(k=3, s=2, p=1, output_padding=1) (R/snn_model/vae_model.py:139-146) equals a stride-1
 * 3x3 convolution with flipped taps of this tensor, which is how the decoder runs on sd_conv_lif_tc. */
int sd_stf_upsample2x(const void* in, void* out, int T, int B, int C, int H, int W, void* stream);
/* out (T, B, C, ceil(H/2), ceil(W/2)) with out[.., y, x] = in[.., 2y, 2x]: the outputs of Conv2d(k=3, s=2, p=1)
 * (R/snn_model/vae_model.py:107-114) are the even positions of the stride-1 convolution computed by sd_conv_lif_tc
 * (the LIF recurrence is per neuron, so dropping the other positions afterwards is exact). */
int sd_stf_subsample2x(const void* in, void* out, int T, int B, int C, int H, int W, void* stream);

/* Un-fused eval-mode BatchNorm2d (SJ/activation_based/layer.py:458-465 -> F.batch_norm with running stats):
 * out[n, c, i] = x[n, c, i] * scale[c] + shift[c], x fp32 [n_outer, C, HW]. */
int sd_channel_affine(const float* x, const float* scale, const float* shift, float* out, int64_t n_outer, int C,
                      int64_t HW, void* stream);

/* LIF membrane state between the reference layout fp32 [B, C, H, W] (LIFNode.v, neuron.py:260-263) and the planar
 * layout fp32 [C/8][R_alloc][8] used by the fused layers.  to_planar != 0: src = [B,C,H,W], dst = planar. */
int sd_state_convert(const float* src, float* dst, int B, int C, int H, int W, int to_planar, void* stream);

/* ---- (a1) multi-step LIF neuron ---------------------------------------------------------------
 * Replaces LIFNode.multi_step_forward, eval branch:
 *   SJ/activation_based/neuron.py:971-1010 -> jit_eval_multi_step_forward_{hard,soft}_reset_{,no_}decay_input
 *   (:799-809, :827-835, :854-862, :880-888), state protocol neuron.py:260-263.
 * x_seq, spike_seq: fp32 [T, N]; v: fp32 [N], read and written in place (persists across calls).
 * hard_reset != 0: v = v_reset on spike; else soft reset (v -= v_th), v_reset ignored.
 * h_seq_or_null: optional fp32 [T, N] pre-fire potential (used by tests / training).
 */
int sd_lif_forward(const float* x_seq, float* v, float* spike_seq, float* h_seq_or_null,
                   int T, int64_t N, float tau, float v_threshold, float v_reset,
                   int hard_reset, int decay_input, void* stream);

/* (a2 / f.1) surrogate-gradient BPTT of the same recurrence, ATan surrogate with slope `alpha`
 * (SJ/activation_based/neuron.py:210-258, surrogate.py:663-678; the reference's generated kernel is reproduced in
 * SURVEY.md Appendix A).  h_seq is the pre-fire potential saved by sd_lif_forward.  grad_v_last (optional, [N]) is
 * dL/dv after the last step; grad_v_init (optional, [N]) receives dL/dv before the first step. */
int sd_lif_backward(const float* grad_spike_seq, const float* grad_v_last_or_null, const float* h_seq,
                    float* grad_x_seq, float* grad_v_init_or_null, int T, int64_t N, float tau, float v_threshold,
                    float v_reset, int hard_reset, int decay_input, int detach_reset, float alpha, void* stream);

/* ---- (a6) MembraneOutputLayer -----------------------------------------------------------------
 * Replaces R/snn_model/snn_layers.py:28-41 (coef[t] = 0.8^(T-1-t), any T).
 * x: fp32 [T, N] -> out fp32 [N]; apply_tanh != 0 fuses the torch.tanh of R/snn_model/vae_model.py:186.
 * coef_host: HOST pointer to the module's `coef` buffer (T floats).  The buffer is part of the reference's
 * checkpoints (`memout.coef`), so the values are taken from the caller rather than recomputed.
 */
int sd_memout(const float* x, float* out, const float* coef_host, int T, int64_t N, int apply_tanh, void* stream);

/* ---- (a7) vector quantiser --------------------------------------------------------------------
 * sd_vq_feature replaces R/snn_model/vae_model.py:42-46:
 *   z = (1-alpha)*memout(x) + alpha*sum_t x / T, written NHWC-flat as fp32 [B*H*W, D].
 *   spikes: STF [T][D/8][R_alloc][8].
 * sd_vq_lookup replaces get_code_indices (vae_model.py:87-95): argmin_k (|z|^2 + |e_k|^2) - 2 z.e_k,
 *   first index on ties.  z: fp32 [M, D], codebook: fp32 [K, D], idx: int64 [M].  Any D >= 1 (8 / 16 / 32 use a
 *   register-blocked kernel).
 *   margin_or_null: optional fp32 [M], second-best minus best distance.
 * sd_vq_gather replaces quantize + permute(0,3,1,2) (vae_model.py:97-99, :54): idx int64 [B*H*W] ->
 *   fp32 [B, D, H, W].
 */
int sd_vq_feature(const void* spikes_stf, const float* alpha_dev, const float* coef_host, float* z, int T, int B,
                  int D, int H, int W, void* stream);
int sd_vq_lookup(const float* z, const float* codebook, int64_t* idx, float* margin_or_null,
                 int64_t M, int D, int K, void* stream);
int sd_vq_gather(const int64_t* idx, const float* codebook, float* out_nchw, int B, int D, int H, int W, int K,
                 void* stream);

/* ---- (a4,a5,a8,a9,a11) fused conv -> BN -> LIF layers -------------------------------------------
 * One descriptor serves both implementations:
 *   sd_conv_lif_simt : CUDA-core fp32 direct convolution (any kernel/stride/padding, transposed or not,
 *                      real-valued or spike input).  Used for the real-input layers (K-dim 9/18/16)
 *                      and the stride-2 / transposed layers of the VQ-VAE.
 *   sd_conv_lif_tc   : tcgen05 implicit GEMM (3x3, stride 1, pad 1, spike input, C_in % 16 == 0),
 *                      accumulators for all T timesteps resident in TMEM, BN + LIF in the epilogue.
 * Replaces layer.Conv2d / layer.ConvTranspose2d -> layer.BatchNorm2d -> neuron.LIFNode in 'm' mode:
 *   SJ/activation_based/layer.py:164-173, :316-325, :458-465; neuron.py:799-809; used at
 *   R/snn_model/vae_model.py:34-38,109-124,139-155 and R/snn_model/vq_diffusion.py:161-187.
 */
enum {
  SD_IN_REAL_CONST = 0, /* fp32 [B, C_in, H_in, W_in], identical at every timestep (R/main.py:133 repeat) */
  SD_IN_REAL_SEQ = 1,   /* fp32 [T, B, C_in, H_in, W_in] */
  SD_IN_STF = 2,        /* fp16 STF spikes (or T-summed spike counts when in_T == 1) */
  SD_IN_STF8 = 3,       /* u8 STF8 spikes: [T][2][C/16][R_alloc][16], per timestep the planes of s in {0,1} and of
                           128*s in {0,128}; operand of the kind::i8 layers (sd_conv_lif_tc with nsplit = 3) */
  SD_IN_TOKENS = 4      /* the denoiser's input built on the fly: args.in = int64 token ids [B, H_in, W_in], channel 0 =
                           (float)token, channel 1 = args.in_scalar (the diffusion time t), identical at every timestep:
                           cat(x, t * ones) -> repeat(T) of R/snn_model/vq_diffusion.py:195-198.  C_in must be 2. */
};
enum {
  SD_OUT_LIF = 0,       /* BN affine -> LIF over T -> spikes (STF) [+ optional T-sum STF] */
  SD_OUT_REAL_SEQ = 1,  /* affine only -> fp32 [T, B, C_out, H_out, W_out] (un-fused layer.Conv2d) */
  SD_OUT_MEMOUT_TANH = 2, /* affine -> sum_t 0.8^(T-1-t) y_t -> tanh -> fp32 [B, C_out, H_out, W_out]
                             (decoder tail, R/snn_model/vae_model.py:152-153,186) */
  SD_OUT_MEAN_T = 3,    /* affine -> sum_t y_t / T -> fp32 [B, H_out, W_out, C_out] (channels last)
                             (denoiser read-out, R/snn_model/vq_diffusion.py:205-206) */
  SD_OUT_LIF8 = 4,      /* like SD_OUT_LIF with the spikes written as STF8 (u8); the optional T-sum stays fp16 STF */
  SD_OUT_CURRENT_SEQ = 5 /* sd_conv_lif_tc with nsplit = 3 only: affine only (the un-fused layer.Conv2d of the training
                           branch on tensor cores) -> fp32 in the planar STF row geometry [T][C_out/8][R_alloc][8]; guard
                           and pad rows are not written */
};

typedef struct sd_conv_desc {
  int T;                 /* timesteps */
  int B;                 /* images */
  int C_in, H_in, W_in;
  int C_out, H_out, W_out;
  int kh, kw, stride, pad;
  int transposed;        /* 0: Conv2d, 1: ConvTranspose2d (output_padding implied by H_out/W_out) */
  int in_kind;           /* SD_IN_* */
  int out_kind;          /* SD_OUT_* */
  int in_T;              /* SD_IN_STF only: T, or 1 if the input is a T-summed count tensor */
  int C_in0;             /* SD_IN_STF only: channels taken from `in` (the rest, C_in - C_in0, from `in2`;
                            torch.cat((x5, x1), dim=2), R/snn_model/vq_diffusion.py:205) */
  float tau, v_threshold, v_reset;
  int hard_reset;        /* LIFNode(v_reset=None) <=> 0 */
  int nsplit;            /* tc only: 1 or 2 = fp16 terms per fp32 weight (kind::f16); 3 = three int8 digits of a 22-bit
                            fixed-point weight (kind::i8, exact int32 accumulation), see sd_conv_pack_weights_tc */
  int concurrent;        /* tc only, tuning hint: how many launches of this size the caller keeps in flight on
                            different streams (0 or 1 = this launch has the GPU to itself).  A lone small batch is
                            cut into narrower N tiles to occupy more SMs; concurrent sub-batches are not. */
} sd_conv_desc;

typedef struct sd_conv_args {
  const void* in;        /* per in_kind */
  const void* in2;       /* second concat segment (STF) or NULL */
  const void* weights;   /* simt: fp32 from sd_conv_pack_weights_simt; tc: fp16 from sd_conv_pack_weights_tc */
  const float* scale;    /* [C_out] folded BN scale  gamma / sqrt(var + eps)                (1 if no BN) */
  const float* shift;    /* [C_out] folded BN shift  (bias - mean) * scale + beta           (bias if no BN) */
  float* v;              /* LIF state fp32 [C_out8/8][R_alloc_out][8]-planar, or NULL = start from v_reset
                            and discard (the sampler resets every step, R/snn_model/vq_diffusion.py:129) */
  void* out;             /* SD_OUT_LIF: STF spikes;  otherwise fp32 per out_kind */
  void* out_sum;         /* SD_OUT_LIF: optional STF [1][C_out/8][R][8] holding sum_t spikes, or NULL */
  const float* memout_coef_host; /* SD_OUT_MEMOUT_TANH: HOST pointer, T floats (the `memout.coef` buffer) */
  void* workspace;       /* tc only: scratch of sd_conv_workspace_bytes(desc) bytes (0 for T <= 4), or NULL; used to
                            carry the membrane potential between the T/4 passes when `v` is NULL */
  float in_scalar;       /* SD_IN_TOKENS: the value of input channel 1 (diffusion time t as a float) */
} sd_conv_args;

int64_t sd_conv_weight_bytes_simt(const sd_conv_desc* d);
int64_t sd_conv_weight_bytes_tc(const sd_conv_desc* d);
int64_t sd_conv_workspace_bytes(const sd_conv_desc* d);
/* Key of the packed tc weight layout (N tile, K block, CTA pairing).  The layout depends on the batch size through the
 * tile configuration: descriptors that differ only in B / concurrent may share one packed buffer iff their keys are
 * equal.  Returns -1 for unsupported descriptors. */
int64_t sd_conv_weight_layout_tc(const sd_conv_desc* d);
/* w: the reference's own parameter layout, fp32 [C_out, C_in, kh, kw] (Conv2d) or [C_in, C_out, kh, kw]
 * (ConvTranspose2d).  tc packing splits each weight into nsplit fp16 terms after an exact power-of-two
 * per-output-channel scaling; chan_scale_out [C_out] receives the inverse scaling to be multiplied
 * into `scale` by the caller. */
int sd_conv_pack_weights_simt(const sd_conv_desc* d, const float* w, void* packed, void* stream);
int sd_conv_pack_weights_tc(const sd_conv_desc* d, const float* w, void* packed, float* chan_scale_out,
                            void* stream);
int sd_conv_lif_simt(const sd_conv_desc* d, const sd_conv_args* a, void* stream);
int sd_conv_lif_tc(const sd_conv_desc* d, const sd_conv_args* a, void* stream);
/* 1 if sd_conv_lif_tc supports the descriptor on this build. */
int sd_conv_tc_supported(const sd_conv_desc* d);
/* Diagnostics (tools/trace_tc.py): while `buf` (device memory, >= grid * 64 int64) is set, sd_conv_lif_tc launches
 * write per-CTA cycle stamps of the MMA and epilogue warps to it.  Pass NULL to switch tracing off.  Only the
 * -DSD_TRACE build of the library carries the stamps (the shipped kernel has no diagnostics code); the shipped library
 * returns SD_ERR_UNSUPPORTED for a non-NULL buffer.  No reference counterpart. */
int sd_debug_tc_trace(void* buf);
/* Diagnostics (tools/bench_layers.py): the SD_TC_* tile-shape knobs are read from the environment once per process;
 * this re-reads them.  No reference counterpart. */
int sd_debug_tc_reload_knobs(void);

/* ---- (f.1) training-path kernels (fp32 on CUDA cores: tiled implicit GEMMs) -------------------------------------
 * The reference trains through torch autograd over F.conv2d / F.conv_transpose2d / F.batch_norm
 * (SJ/activation_based/layer.py:164-173,316-325,458-465).  The input gradient of a convolution is the adjoint
 * convolution and is computed with sd_conv_lif_simt (SD_IN_REAL_SEQ -> SD_OUT_REAL_SEQ); these entry points add the
 * weight/bias gradient and train-mode BatchNorm.
 * sd_conv_wgrad: d describes the FORWARD op; x fp32 [T,B,C_in,H_in,W_in], grad_out fp32 [T,B,C_out,H_out,W_out];
 *   grad_w in the reference parameter layout ([C_out,C_in,kh,kw] or [C_in,C_out,kh,kw]), grad_bias [C_out] or NULL.
 * sd_bn_train_forward: x, y fp32 [n_outer, C, HW]; batch mean / biased variance per channel are written for the
 *   backward pass and the caller's running-statistics update.
 * workspace: sd_conv_wgrad_workspace_bytes(d) bytes of scratch (partial sums of the split reduction, added in a fixed
 *   order: the result is deterministic). */
int64_t sd_conv_wgrad_workspace_bytes(const sd_conv_desc* d);
int sd_conv_wgrad(const sd_conv_desc* d, const float* x, const float* grad_out, float* grad_w, float* grad_bias_or_null,
                  void* workspace, void* stream);
int sd_bn_train_forward(const float* x, const float* gamma_or_null, const float* beta_or_null, float* y,
                        float* mean_out, float* var_out, int64_t n_outer, int C, int64_t HW, float eps, void* stream);
int sd_bn_backward(const float* x, const float* grad_out, const float* mean, const float* var,
                   const float* gamma_or_null, float* grad_x, float* grad_gamma_or_null, float* grad_beta_or_null,
                   int64_t n_outer, int C, int64_t HW, float eps, void* stream);
/* The same BatchNorm with statistics shared across GPUs (the optional training exchange step of SURVEY.md 8(e); the
 * reference would use torch.nn.SyncBatchNorm).  The caller all-gathers (count, mean, M2) and all-reduces the two backward
 * sums over NCCL between these calls (activation_based/layer.py:_SyncBNFn); the normalisation itself is
 * sd_channel_affine with scale = gamma * invstd, shift = beta - mean * scale. */
int sd_bn_local_stats(const float* x, float* mean_out, float* m2_out, int64_t n_outer, int C, int64_t HW, void* stream);
int sd_bn_backward_reduce(const float* x, const float* grad_out, const float* mean, const float* var, float* sum_gy,
                          float* sum_gy_xhat, int64_t n_outer, int C, int64_t HW, float eps, void* stream);
int sd_bn_backward_apply(const float* x, const float* grad_out, const float* mean, const float* var, const float* gamma,
                         const float* mean_gy, const float* mean_gy_xhat, float* grad_x, int64_t n_outer, int C, int64_t HW,
                         float eps, void* stream);

/* ---- (a12) absorbing-diffusion sampling step ----------------------------------------------------
 * Torch-compatible Philox4x32-10 streams (TORCH/include/ATen/native/cuda/DistributionTemplates.h:50-87):
 * element li of a call with (seed, offset) is component (li / tpg) % 4 of
 * philox(key = seed, counter = (offset/4 + li / (4*tpg), subsequence = li % tpg)), tpg = 256 * grid,
 * grid = min(sm_count * (max_threads_per_sm / 256), ceil(numel / 256)).
 * sd_philox_uniform     == torch.rand / rand_like on CUDA (fp32, [0,1) after the bound flip, :493-503)
 * sd_philox_exponential == Tensor.exponential_(1) on CUDA (TORCH/include/ATen/core/TransformationHelper.h:129-146)
 * *offset_increment_out (host) receives what torch adds to the generator offset for that call.
 */
int sd_philox_uniform(float* out, int64_t numel, uint64_t seed, uint64_t offset, int64_t index_base,
                      int64_t numel_global, uint64_t* offset_increment_out, void* stream);
int sd_philox_exponential(float* out, int64_t numel, uint64_t seed, uint64_t offset, int64_t index_base,
                          int64_t numel_global, uint64_t* offset_increment_out, void* stream);
int sd_philox_offset_increment(int64_t numel_global, uint64_t* inc_out);

/* One reverse-diffusion step, replacing the loop body of AbsorbingDiffusion.sample
 * (R/snn_model/vq_diffusion.py:111-140) after the denoiser call:
 *   changes = rand < 1/t ; changes &= ~unmasked ; unmasked |= changes           (:118-124)
 *   probs = Categorical(logits = logits / temp).probs                           (:134-137)
 *   x0_hat = argmax(probs / Exp(1))                                             (:138, multinomial n=1)
 *   x_t[changes] = x0_hat[changes]                                              (:140)
 * logits: fp32 [n_tokens, K] (channels last), 1 <= K <= 1024 (128-bit loads when K is a multiple of 128 and the
 * pointer is 16-byte aligned); x_t: int64 [n_tokens]; unmasked: uint8 [n_tokens].
 * The shard [token_base, token_base + n_tokens) of a global batch of n_tokens_global tokens draws the
 * Philox values of its GLOBAL element indices, so a batch sharded over GPUs reproduces the single-GPU
 * stream.  offset_uniform / offset_exponential are the generator offsets of the two
 * draws of this step.  x0_hat_or_null: optional int64 [n_tokens] (the raw categorical draw).
 */
int sd_sample_step(const float* logits, int64_t* x_t, uint8_t* unmasked, int64_t* x0_hat_or_null,
                   int64_t n_tokens, int K, int t, float temp, uint64_t seed, uint64_t offset_uniform,
                   uint64_t offset_exponential, int64_t token_base, int64_t n_tokens_global, void* stream);

/* Same step with (seed, base generator offset, extra token base) read from DEVICE memory (rng_dev[0] = seed,
 * rng_dev[1] = offset, a multiple of 4, rng_dev[2] = number of tokens added to token_base) and the two per-step
 * offsets given relative to that base: the launch parameters are then independent of the RNG state and of the
 * position of the batch in the global stream, so ONE CUDA graph of the whole sampling loop can be replayed with a new
 * stream, or for the next chunk of a large batch, by rewriting 24 bytes of device memory. */
int sd_sample_step_dev(const float* logits, int64_t* x_t, uint8_t* unmasked, int64_t* x0_hat_or_null,
                       int64_t n_tokens, int K, int t, float temp, const uint64_t* rng_dev,
                       uint64_t rel_offset_uniform, uint64_t rel_offset_exponential, int64_t token_base,
                       int64_t n_tokens_global, void* stream);

/* Builds the denoiser's first-layer input cat(x_t as float, t) (R/snn_model/vq_diffusion.py:195-196):
 * x_t int64 [B*H*W] -> fp32 [B, 2, H, W]. */
int sd_denoiser_input(const int64_t* x_t, float* out, int B, int H, int W, int t, void* stream);

/* clip(pred + 0.5, 0, 1) * 255 -> uint8 (R/main.py:401). */
int sd_to_uint8(const float* pred, uint8_t* out, int64_t N, void* stream);

/* ---- (f.4) quality metrics on given features (the step after the path) -----------
 * The algebra of the reference's evaluation block that needs no pretrained network.  workspace: device memory of
 * sd_metric_workspace_bytes(n_elements, d, m) bytes (pass the sizes of the call: d for the Frechet distance, m for the
 * MMD, K + N / splits as n_elements for the inception score, 0 otherwise).  Scalars are written to DEVICE memory.
 *   sd_metric_mse            F.mse_loss(a, b)                                            R/main.py:319
 *   sd_metric_ssim           metric.pytorch_ssim.SSIM(window_size)(img1, img2): gaussian window (sigma 1.5), zero padding,
 *                            C1 = 0.01^2, C2 = 0.03^2, mean of the SSIM map; window_host = the window_size^2 fp32 window
 *                            (create_window); per_plane_sum_or_null [N*C] gets the per-plane sums (size_average=False)
 *                                                                R/metric/pytorch_ssim/__init__.py:7-37, R/main.py:320-321
 *   sd_metric_feature_stats  mu = mean(act, 0), sigma = np.cov(act, rowvar=False) in fp64; act fp32 or fp64 [N, d]
 *                                                                                        R/metric/Fid_score.py:100-113
 *   sd_metric_frechet        |mu1-mu2|^2 + tr(s1) + tr(s2) - 2 tr(U sqrt(S) Vh), U S Vh = svd(s1 s2): the reference's own
 *                            sqrtm (an SVD, not scipy's), by a one-sided Jacobi SVD in fp64.  Blocks the stream (sweeps are
 *                            repeated until none rotates).                         R/metric/Fid_score.py:14-17,116-173
 *   sd_metric_poly_mmd2      unbiased MMD^2 with k(x, y) = (gamma x.y + coef)^degree on fp32 features [m, d] of both sets:
 *                            (sum_{i!=j} k_xx + sum_{i!=j} k_yy) / (m (m-1)) - 2 sum k_xy / m^2 -- the estimator of
 *                            torchmetrics.image.kid (poly_mmd), which R/main.py:465-490 calls; torchmetrics is not part of
 *                            the reference tree and no version is pinned there
 *   sd_metric_inception_score  per split exp(mean_i KL(p_i || mean_i p_i)) with scipy.stats.entropy's normalisation, then
 *                            mean and population std over the splits; preds fp64 [N, K]        R/metric/IS_score.py:58-72
 */
int64_t sd_metric_workspace_bytes(int64_t n_elements, int d, int m);
int sd_metric_mse(const float* a, const float* b, int64_t n, float* out_dev, void* workspace, void* stream);
int sd_metric_ssim(const float* img1, const float* img2, int N, int C, int H, int W, int window_size,
                   const float* window_host, float* out_dev, float* per_plane_sum_or_null, void* workspace, void* stream);
int sd_metric_feature_stats(const void* act, int act_is_f64, int64_t N, int d, double* mu_out, double* sigma_out, void* stream);
int sd_metric_frechet(const double* mu1, const double* sigma1, const double* mu2, const double* sigma2, int d,
                      double* out_dev, int* sweeps_out, void* workspace, void* stream);
int sd_metric_poly_mmd2(const float* fx, const float* fy, int m, int d, int degree, double gamma, double coef,
                        double* out_dev, void* workspace, void* stream);
int sd_metric_inception_score(const double* preds, int64_t N, int K, int splits, double* mean_out_dev, double* std_out_dev,
                              void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SD_B200_H_ */
