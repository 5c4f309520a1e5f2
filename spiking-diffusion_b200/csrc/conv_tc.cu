// Fused 3x3 conv -> BatchNorm -> multi-step LIF as ONE tcgen05 implicit-GEMM kernel (sm_100a).
//
// Replaces layer.Conv2d -> layer.BatchNorm2d -> neuron.LIFNode for the spike->spike layers of the denoiser
// (R/snn_model/vq_diffusion.py:165-187 via SJ/activation_based/layer.py:164-173,458-465 and neuron.py:799-809),
// which carry >99.9 % of the sampling FLOPs.  The stride-2 layers of the VQ-VAE use it too: the
// transposed convolutions of the decoder as 3x3 convolutions of zero-inserted spikes (sd_stf_upsample2x) and enc.conv2 at
// stride 1 followed by sd_stf_subsample2x (engine._UpsampledConvT / _Stride1Conv).
//
// GEMM view.  Activations live in the STF layout (include/sd_b200.h): per timestep and per 8-channel chunk a
// plane of 16-byte rows, rows = the pixels of all images flattened densely (H*W rows per image, no padding).  For
// that layout the 3x3/stride-1/pad-1 convolution is nine GEMMs whose A operands are the SAME rows shifted by
// dy*W + dx.  A shifted row that would cross an image border (the zero padding of the convolution) is removed by
// masking the affected OUTPUT rows of that tap's MMAs with the tcgen05.mma disable-output-lane mask (y == 0 for
// dy = -1, y == H-1 for dy = +1, x == 0 for dx = -1, x == W-1 for dx = +1), so every row of an M tile is useful
// work.  In shared memory the planes are exactly the tcgen05 "no-swizzle, K-major" canonical layout (core matrix =
// 8 rows x 16 B contiguous, SBO = 128 B between 8-row groups, LBO = plane stride between 8-channel chunks), so a tap
// is just a different 16-byte-aligned start address in the A descriptor: the input tile is loaded ONCE per K-block
// and reused by all 9 taps (9x less L2->SMEM traffic than im2col).
//
//   M tile  = 128 consecutive rows of the flat pixel sequence (tiles may straddle images)
//   N tile  = 128 output channels (64 / 32 when a small batch would leave most SMs without a tile)
//   K block = 32 input channels (64 for the linear read-out), x 9 taps, x nsplit fp16 weight terms
//   D       = one [128 x N] fp32 accumulator per timestep, up to 4 timesteps resident in TMEM (512 columns, two
//             stages when N <= 64).  T <= 4: one pass, the membrane potential stays in registers.  T = 8 / 16: 2 / 4
//             passes of 4 timesteps over the same tile, potential and spike counts carried in an L2-resident plane.
//
// Exactness.  Spikes (and T-summed spike counts) are exact in fp16.  Each fp32 weight is scaled by an exact
// per-output-channel power of two and split into nsplit fp16 terms (hi, lo = 22 significant bits for
// nsplit = 2); every product is exact and accumulation is fp32 in the tensor core.  Both terms accumulate into
// the same TMEM accumulator (the A tile is shared, only the B descriptor changes).
//
// CTA pairs.  By default two CTAs of a cluster (two SMs of a TPC) compute one M = 256 tile with
// tcgen05.mma.cta_group::2: each SM stages its own 128 A rows and HALF of the B block, so the shared-memory operand
// fetch per SM drops from 8 KB to 6 KB per 64-cycle MMA (measured +20-25 % per layer, profiles/r01_experiments.md).
// The leader CTA issues the MMAs; the peer's relay warp forwards "stage landed" to the leader's barriers with remote
// mbarrier arrives, tcgen05.commit multicasts stage releases and "accumulator ready" to both CTAs, and the peer's
// epilogue releases the accumulator with a remote arrive.
//
// Warp roles (384 threads, 1 CTA / SM, persistent over tiles):
//   warps 0-7  epilogue: tcgen05.ld -> BN affine -> LIF over all T in registers -> fp16 spikes / T-sum / v
//   warp  8    A producer  (cp.async.bulk global -> shared, one copy per (t, chunk) plane, mbarrier tx)
//   warp  9    B producer  (cp.async.bulk of one pre-packed [split][chunk][n][8] weight block per tap)
//   warp 10    MMA issuer  (one lane issues tcgen05.mma, tcgen05.commit releases stages)
//   warp 11    TMEM allocator; in the peer CTA of a pair also the relay
#include <stdlib.h>
#include <mutex>
#include "common.cuh"
#include <type_traits>

namespace sd {

#ifndef SD_TC_EPI_WARPS
#define SD_TC_EPI_WARPS 8
#endif
constexpr int kEpiWarps = SD_TC_EPI_WARPS;          // epilogue warps: 8 or 12 (each TMEM lane quarter is served by kEpiWarps / 4)
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kTcThreads = kEpiThreads + 128;       // + A producer, B producer, MMA issuer, TMEM allocator / relay
constexpr int kWarpA = kEpiWarps, kWarpB = kEpiWarps + 1, kWarpMma = kEpiWarps + 2, kWarpAux = kEpiWarps + 3;
[[maybe_unused]] constexpr int kRegsOther = 128;        // warps 8-11 after setmaxnreg.dec (SD_TC_SETMAXNREG experiment)
[[maybe_unused]] constexpr int kRegsEpilogue = 192;    // warps 0-7 after setmaxnreg.inc: 256 * 216 + 128 * 72 = 64512 <= 65536
constexpr int kMaxAStages = 4;
constexpr int kMaxBStages = 8;
#ifndef SD_TC_B_PIECES
#define SD_TC_B_PIECES 1
#endif
constexpr int kBPieces = SD_TC_B_PIECES;   // concurrent bulk copies per weight stage
constexpr uint32_t kSmemBudget = 220 * 1024;
// shared-memory header: barriers + TMEM slot in the first 1 KB, then two buffers of (scale[256], shift[256]) for the
// epilogue (the N tile's folded-BN affine, staged while the MMAs of the tile run)
constexpr uint32_t kHdrBytes = 1024 + 2 * 2 * 256 * 4;
constexpr int kMaxColsPerThread = 16 * ((8 + kEpiWarps / 4 - 1) / (kEpiWarps / 4));   // N = 128: 64 (8 warps) / 48 (12 warps)
constexpr uint32_t kVSmemBytes = kEpiThreads * kMaxColsPerThread * 4;   // potentials of one tile between its passes: [columns][threads]

struct TcConfig {
  int i8;          // 1: kind::i8 path (nsplit == 3): u8 spikes x three s8 weight digits, two int32 accumulators
                   // (hi = 128*d0 + d1, lo = d2) per timestep, exact integer accumulation
  int accs;        // TMEM accumulators per timestep (1, or 2 for i8)
  int rowch;       // channels per 16-byte operand row: 8 (fp16) or 16 (u8 / s8)
  int T_acc;       // timesteps per pass (min(T, 4) for fp16 LIF, 2 for i8, 1 for the T-summed linear read-out)
  int pair;        // 1: 2-CTA clusters, tcgen05.mma.cta_group::2 (M = 256 per pair), half of every B block per CTA
  int n_tchunks;   // passes per tile: T / T_acc.  T = 8 / 16 run as 2 / 4 passes of 4 timesteps at N = 128 with the
                   // membrane potential carried between passes in an L2-resident fp32 plane, instead of one pass at
                   // N = 64 / 32 (whose A-operand fetch per MMA is amortised over too few columns)
  int tpar;        // 1: T-parallel small-batch mode.  Every (tile, pass) is an independent work unit whose epilogue writes
                   // the BN-affined input currents (fp32) to the workspace; a second, bandwidth-bound kernel runs the LIF
                   // recurrence over all T.  Used when a multi-pass layer would otherwise occupy a fraction of the SMs.
  int N_TILE;
  int KBLK;        // input channels per K block
  int acc_stages;
  int a_stages, b_stages;
  int halo, rows_ld;
  int ndx;         // 1: one copy of the rows, taps shift by dy*Wp+dx rows (16-byte aligned starts);
                   // 3: three copies pre-shifted by dx = -1,0,+1 so that every tap start is 128-byte aligned
                   //    (needs Wp % 8 == 0).  Measured: no gain, the cost of an MMA does not depend on the alignment
  uint32_t a_stage_bytes, b_stage_bytes, smem_bytes;
  int n_tiles, m_tiles, num_kblocks, c0_blocks;
  uint32_t v_smem_off;   // != 0: multi-pass tiles keep the membrane potential in shared memory between their passes
                         // (offset of the [64][256] fp32 block); 0: single pass, T-parallel mode, or no room
};

struct TcParams {
  const __half* in0;
  const __half* in1;
  const __half* wpack;
  const float* scale;
  const float* shift;
  __half* out_spk;
  uint8_t* out_spk8;        // i8 path: spikes as STF8 ([T][2][C/16][R_alloc][16] u8: planes of s and of 128*s)
  __half* out_sum;
  float* out_real;
  float* cur;               // tpar mode: currents [T][C_out/8][R_alloc][8] fp32 (workspace)
  float* v;                 // state plane (caller's LIF state or workspace), or null
  int v_load_initial;       // first pass starts from *v (else from v_reset)
  int v_store_final;        // last pass writes v back
  int64_t R_alloc, G, R_valid;
  int C8_0, C8_1, Cout, Cout8;
  int T, H, W, Wp, P;
  int nsplit, out_kind, hard_reset;
  float tau, v_th, v_reset;
  uint32_t idesc;
  long long* trace;         // diagnostics (sd_debug_tc_trace): per-CTA cycle stamps, or null
  int dbg;                  // diagnostics (env SD_TC_DBG): 1 = skip the spike stores, 2 = skip the TMEM loads
  TcConfig c;
};
[[maybe_unused]] constexpr int kTraceStride = 128;  // int64 slots per CTA: [0] entry, [1] set-up done, [2] exit, then 8 per tile pass;
                                                     // from 64: warp 0's column groups of pass 1 (4 stamps each: start, state loaded, LIF done, end)

// ---------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with a disable-output-lane mask: bit i of the 128-bit mask set => row i of D is neither written nor
// accumulated by this MMA.
__device__ __forceinline__ void tc_mma_f16_masked(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate, uint32_t m0, uint32_t m1, uint32_t m2,
                                                  uint32_t m3) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}
// kind::i8: u8 / s8 operands (K = 32 per instruction), int32 accumulators
__device__ __forceinline__ void tc_mma_i8_masked(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate, uint32_t m0, uint32_t m1, uint32_t m2,
                                                 uint32_t m3) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m0), "r"(m1), "r"(m2), "r"(m3)
      : "memory");
}
// ---- cta_group::2 (two SMs of a cluster pair share one M = 256 MMA) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32  remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64  _, [remAddr32];\n\t"
      "}" ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16_masked_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                       uint32_t accumulate, const uint32_t (&m)[8]) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]),
      "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7])
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_masked_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                      uint32_t accumulate, const uint32_t (&m)[8]) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
      "}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(m[0]), "r"(m[1]), "r"(m[2]), "r"(m[3]),
      "r"(m[4]), "r"(m[5]), "r"(m[6]), "r"(m[7])
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also names the destination registers as in/out operands, so no use of them can be scheduled above it
__device__ __forceinline__ void tc_ld_wait_on(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// Shared-memory matrix descriptor, no swizzle, K-major (cute::UMMA::SmemDescriptor field layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 0 (none)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 rx;\n\t"
      ".reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t"
      "}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// The centre tap is visited first: the first MMA of a tile overwrites the accumulator (accumulate = 0) and must
// therefore have every output row enabled; all other taps mask the rows at the image border they would cross.
__device__ __forceinline__ int tap_order(int i) { return i == 0 ? 4 : (i <= 4 ? i - 1 : i); }

struct PipeState {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) { stage = 0; phase ^= 1; }
  }
};

// ---------------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------------
// NSPLIT = fp16 terms per weight, KSTEPS = KBLK / 16 (tcgen05.mma K = 16): compile-time so that the MMA issue loop
// unrolls into descriptor-low-word additions with immediates.
// PAIR: the CTA is one half of a 2-CTA cluster; the pair computes an M = 256 tile with tcgen05.mma.cta_group::2 issued by
// the leader (rank 0).  Each SM stages its own 128 A rows and HALF of the B tile (N/2 rows), which cuts the
// shared-memory operand fetch per MMA from A 4 KB + B 4 KB to A 4 KB + B 2 KB per SM.
// NSPLIT == 3 selects the kind::i8 path: operands are bytes (16 channels per 16-byte row, K = 32 per MMA, KSTEPS =
// KBLK / 32); per timestep the A stage holds the planes of s and of 128*s, the B stage three digit blocks
// [d0 | d1 | d2] with w_fix = 256 * (128 * d0 + d1) + d2, and TMEM two int32 accumulators: hi += A128 x d0 + A1 x d1,
// lo += A1 x d2.  Three K = 32 MMAs cover 32 input channels where the fp16 path needs four K = 16 MMAs.
template <int NSPLIT, int KSTEPS, bool PAIR>
__global__ void __launch_bounds__(kTcThreads, 1) conv3x3_tc_kernel(const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TcConfig& c = p.c;
  constexpr bool I8 = NSPLIT == 3;
  constexpr int ACCS = I8 ? 2 : 1;            // TMEM accumulators per timestep
  // barrier block (first 256 B), then A stages, then B stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  const uint32_t bar_base = smem_u32(bars);
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (kMaxAStages + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + kMaxBStages + s); };
  auto acc_full = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + s); };
  auto acc_empty = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + 2 + s); };
  // 16-byte aligned (tcgen05.alloc traps on a misaligned destination); both CTAs of a pair pass the same offset
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 8 * (2 * kMaxAStages + 2 * kMaxBStages + 4));
  // pair only: "the peer's stage has landed" barriers in the leader, arrived remotely by the peer's relay warp
  auto a_full_peer = [&](int s) { return bar_base + 8u * (2 * kMaxAStages + 2 * kMaxBStages + 6 + s); };
  auto b_full_peer = [&](int s) { return bar_base + 8u * (3 * kMaxAStages + 2 * kMaxBStages + 6 + s); };
  const uint32_t a_base = smem_u32(smem + kHdrBytes);
  float* const s_affine = reinterpret_cast<float*>(smem + 1024);     // [2 buffers][scale 256 | shift 256]
  float* const s_v = c.v_smem_off ? reinterpret_cast<float*>(smem + c.v_smem_off) : nullptr;
  const uint32_t b_base = a_base + c.a_stages * c.a_stage_bytes;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  // cycle-stamp diagnostics exist only in the -DSD_TRACE build (tools/trace_tc.py); the shipped kernel has none of it
#ifdef SD_TRACE
  long long* const trace = p.trace ? p.trace + (int64_t)blockIdx.x * kTraceStride : nullptr;
  const int dbg = p.dbg;
#else
  constexpr long long* trace = nullptr;
  constexpr int dbg = 0;
#endif
  if (trace && threadIdx.x == 0) trace[0] = clock64();
  // programmatic dependent launch: let the next kernel of the chain start its prologue now; this kernel's own prologue,
  // its weight stream and its first MMAs' B operands do not depend on the previous kernel -- only the A producer and the
  // epilogue warps wait for it (pdl_wait below)
  if (threadIdx.x == 0) pdl_launch_dependents();

  if (threadIdx.x == 0) {
    for (int s = 0; s < c.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); mbar_init(a_full_peer(s), 1); }
    for (int s = 0; s < c.b_stages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); mbar_init(b_full_peer(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), PAIR ? 2 * kEpiWarps : kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpAux) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "r"(512u));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (trace && threadIdx.x == 0) trace[1] = clock64();

  // work units: (M tile, N tile) per CTA, or (pair of M tiles, N tile) per cluster
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  // In T-parallel mode a unit is one (tile, pass); otherwise a unit is a tile and its passes run back to back.
  const int unit_passes = c.tpar ? 1 : c.n_tchunks;
  const int total_tiles = (PAIR ? (c.m_tiles + 1) / 2 : c.m_tiles) * c.n_tiles * (c.tpar ? c.n_tchunks : 1);
  auto tile_of = [&](int unit) { return c.tpar ? unit / c.n_tchunks : unit; };
  auto pass0_of = [&](int unit) { return c.tpar ? unit % c.n_tchunks : 0; };
  auto m_tile_of = [&](int unit) {
    const int tl = tile_of(unit);
    return PAIR ? 2 * (tl / c.n_tiles) + (int)cta_rank : tl / c.n_tiles;
  };
  const int chunks = I8 ? c.KBLK >> 4 : c.KBLK >> 3;   // 16-byte-row chunks (8 fp16 / 16 u8 channels) per K block
  const int planes_t = I8 ? 2 * chunks : chunks * c.ndx;   // planes per timestep in an A stage
  const uint32_t plane_bytes = (uint32_t)c.rows_ld * 16u;

  // Register budget by role (setmaxnreg works on aligned groups of four warps): the producer / MMA / relay warps of the
  // third warpgroup need few registers and give theirs up; the two epilogue warpgroups take them, so the LIF epilogue
  // (16 columns x 2 timesteps in flight, affine, potentials) is no longer scheduled against the 168-register ceiling.
#ifdef SD_TC_SETMAXNREG
  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegsOther));
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegsEpilogue));
#endif

  if (warp == kWarpA) {
    // ===== A producer =====
    pdl_wait();                                      // the spikes are the previous kernel's output
    PipeState st;
    const int ncopy = c.T_acc * planes_t;
    for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
      // an odd number of M tiles leaves the last pair with a phantom second tile: it re-reads the last real tile (its
      // epilogue writes nothing because its rows are >= R_valid)
      const int64_t row0 = (int64_t)min(m_tile_of(tile), c.m_tiles - 1) * kTileRows;
      for (int kbi = 0; kbi < c.num_kblocks * unit_passes; ++kbi) {
        const int kb = kbi % c.num_kblocks, tch = pass0_of(tile) + kbi / c.num_kblocks;
        if (lane == 0) {
          mbar_wait(a_empty(st.stage), st.phase ^ 1);
          mbar_expect_tx(a_full(st.stage), c.a_stage_bytes);
        }
        __syncwarp();
        const bool seg1 = kb >= c.c0_blocks;
        const __half* src = seg1 ? p.in1 : p.in0;
        const int C8 = seg1 ? p.C8_1 : p.C8_0;
        const int chunk0 = (seg1 ? kb - c.c0_blocks : kb) * chunks;
        for (int i = lane; i < ncopy; i += 32) {
          if constexpr (I8) {
            // stage order [t][scale: s, 128*s][chunk]; global STF8 order [t][scale][C/16][row][16]
            const int tl = i / planes_t, rem = i - tl * planes_t;
            const int sc = rem / chunks, ch = rem - sc * chunks;
            const int t = tch * c.T_acc + tl;
            const uint8_t* g = reinterpret_cast<const uint8_t*>(src) +
                               ((((int64_t)(t * 2 + sc) * C8 + chunk0 + ch) * p.R_alloc) + p.G + row0 - c.halo) * 16;
            bulk_g2s(a_base + st.stage * c.a_stage_bytes + (uint32_t)i * plane_bytes, g, plane_bytes, a_full(st.stage));
          } else {
            const int pl = i / c.ndx, dxi = i - pl * c.ndx;          // plane (t, chunk) and its dx copy
            const int tl = pl / chunks, ch = pl - tl * chunks;
            const int t = tch * c.T_acc + tl;
            const int dx = c.ndx == 3 ? dxi - 1 : 0;
            const __half* g = src + ((((int64_t)t * C8 + chunk0 + ch) * p.R_alloc) + p.G + row0 - c.halo + dx) * 8;
            bulk_g2s(a_base + st.stage * c.a_stage_bytes + (uint32_t)i * plane_bytes, g, plane_bytes, a_full(st.stage));
          }
        }
        st.advance(c.a_stages);
      }
    }
  } else if (warp == kWarpB) {
    // ===== B producer (whole warp converged; kBPieces lanes issue one bulk copy each per stage) =====
    PipeState st;
    const int64_t stage_halfs = c.b_stage_bytes / 2;
    const int halves = PAIR ? 2 : 1;                 // pair: each CTA stages its own half (N/2 rows) of every B block
    for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
      const int n_tile = tile_of(tile) % c.n_tiles;
      const __half* wsrc = p.wpack + (int64_t)n_tile * c.num_kblocks * 9 * stage_halfs * halves + (int64_t)cta_rank * stage_halfs;
      for (int itt = 0; itt < c.num_kblocks * 9 * unit_passes; ++itt) {
        const int it = itt % (c.num_kblocks * 9);      // every T pass streams the same weights again
        const int kb = it / 9, tap = tap_order(it - kb * 9);
        mbar_wait(b_empty(st.stage), st.phase ^ 1);
        if (lane == 0) mbar_expect_tx(b_full(st.stage), c.b_stage_bytes);
        __syncwarp();
        // One bulk copy per stage (kBPieces = 1).  Measured (tools/probe_bulk.cu, profiles/r02_experiments.md): a
        // cp.async.bulk costs the issuing thread ~600 cycles whatever its size (12 KB: ~20 B/clk, 48 KB: ~78 B/clk per SM),
        // so this loop delivers one tap per ~600-730 cycles -- enough for the 768 cycles the 12 MMAs of a tap take at
        // 2 timesteps per pass, not for one timestep per pass (384); cutting a stage into 8 concurrent copies is slower
        // (conv5 MMA phase 58.6 k -> 62 k cycles).  It is an issue-rate limit, not a bandwidth limit.
        if (lane < kBPieces) {
          const uint32_t piece = c.b_stage_bytes / kBPieces;        // stage sizes are multiples of 1 KB
          bulk_g2s(b_base + st.stage * c.b_stage_bytes + lane * piece,
                   reinterpret_cast<const uint8_t*>(wsrc + (int64_t)(kb * 9 + tap) * stage_halfs * halves) + lane * piece,
                   piece, b_full(st.stage));
        }
        __syncwarp();
        st.advance(c.b_stages);
      }
    }
  } else if (PAIR && warp == kWarpAux && !leader) {
    // ===== relay (peer CTA): tell the leader's MMA warp when this CTA's stages have landed, in consumption order =====
    PipeState sa, sb;
    for (int tile = unit0; tile < total_tiles; tile += unit_stride) {
      for (int kbi = 0; kbi < c.num_kblocks * unit_passes; ++kbi) {
        mbar_wait(a_full(sa.stage), sa.phase);
        if (lane == 0) mbar_arrive_remote(a_full_peer(sa.stage), 0);
        __syncwarp();
        for (int ti = 0; ti < 9; ++ti) {
          mbar_wait(b_full(sb.stage), sb.phase);
          if (lane == 0) mbar_arrive_remote(b_full_peer(sb.stage), 0);
          __syncwarp();
          sb.advance(c.b_stages);
        }
        sa.advance(c.a_stages);
      }
    }
  } else if (warp == kWarpMma && leader) {
    // ===== MMA issuer (whole warp converged; one elected lane issues tcgen05.mma / tcgen05.commit) =====
    PipeState sa, sb, sc;
    // descriptor words: lo = start>>4 | (LBO>>4)<<16, hi = SBO>>4 | version(1)<<14; all stepping is done on lo
    const uint32_t desc_hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo_const = (((uint32_t)c.ndx * plane_bytes) >> 4) << 16;   // LBO: next 8-channel chunk
    const uint32_t b_lbo = (uint32_t)(PAIR ? c.N_TILE / 2 : c.N_TILE) * 16u;   // rows of B staged in THIS CTA
    const uint32_t b_lo_const = (b_lbo >> 4) << 16;
    const uint32_t a_step_t = ((uint32_t)planes_t * plane_bytes) >> 4;
    const uint32_t a_off_128 = ((uint32_t)chunks * plane_bytes) >> 4;          // i8: from the s planes to the 128*s planes
    const uint32_t a_step_k = ((uint32_t)(2 * c.ndx) * plane_bytes) >> 4;
    const uint32_t b_step_sp = ((uint32_t)chunks * b_lbo) >> 4;
    const uint32_t b_step_k = (2u * b_lbo) >> 4;
    constexpr int MW = PAIR ? 8 : 4;   // 32-lane mask words: 128 output rows per CTA
    // one iteration per (tile, T pass)
    int trace_it = 0;
    for (int tile = unit0, tch = 0; tile < total_tiles;
         (++tch == unit_passes) ? (tch = 0, tile += unit_stride) : 0) {
      long long stall = 0;
      // rows of this tile on the top / bottom / left / right border of their image: the taps that would read across
      // that border have these output rows disabled.  Computed before the accumulator wait, so it overlaps the
      // epilogue of the previous tile.
      uint32_t m_up[MW], m_dn[MW], m_lf[MW], m_rt[MW];
      {
        const int64_t row0 = (int64_t)m_tile_of(tile) * kTileRows;   // leader: first row of the pair
        const int pp0 = (int)(row0 % p.P);
#pragma unroll
        for (int w = 0; w < MW; ++w) {
          const int pp = (pp0 + w * 32 + lane) % p.P;
          const int y = pp / p.W, x = pp - y * p.W;
          m_up[w] = __ballot_sync(0xffffffffu, y == 0);
          m_dn[w] = __ballot_sync(0xffffffffu, y == p.H - 1);
          m_lf[w] = __ballot_sync(0xffffffffu, x == 0);
          m_rt[w] = __ballot_sync(0xffffffffu, x == p.W - 1);
        }
      }
      mbar_wait(acc_empty(sc.stage), sc.phase ^ 1);
      tc_fence_after();
      if (trace && lane == 0 && trace_it < 7) trace[3 + trace_it * 8 + 0] = clock64();
      const uint32_t d_base = tmem_base + (uint32_t)(sc.stage * c.T_acc * ACCS * c.N_TILE);
      for (int kb = 0; kb < c.num_kblocks; ++kb) {
        const long long w0 = trace ? clock64() : 0;
        mbar_wait(a_full(sa.stage), sa.phase);
        if (PAIR) mbar_wait(a_full_peer(sa.stage), sa.phase);
        if (trace) stall += clock64() - w0;
        const uint32_t a_stage = a_base + sa.stage * c.a_stage_bytes;
        for (int ti = 0; ti < 9; ++ti) {
          const int tap = tap_order(ti);
          const long long w1 = trace ? clock64() : 0;
          mbar_wait(b_full(sb.stage), sb.phase);
          if (PAIR) mbar_wait(b_full_peer(sb.stage), sb.phase);
          if (trace) {
            stall += clock64() - w1;
            if (kb == 0 && ti == 0 && lane == 0 && trace_it < 7) trace[3 + trace_it * 8 + 1] = clock64();
          }
          tc_fence_after();
          if (elect_one()) {
            const int dy = tap / 3 - 1, kx = tap % 3;
            const uint32_t a_off = c.ndx == 3 ? (uint32_t)kx * plane_bytes + (uint32_t)(c.halo + dy * p.Wp) * 16u
                                              : (uint32_t)(c.halo + dy * p.Wp + kx - 1) * 16u;
            uint32_t a_lo = a_lo_const | ((a_stage + a_off) >> 4);
            const uint32_t b_lo0 = b_lo_const | ((b_base + sb.stage * c.b_stage_bytes) >> 4);
            uint32_t d = d_base;
            const uint32_t first = (kb | ti) != 0 ? 1u : 0u;
            uint32_t km[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int w = 0; w < MW; ++w)
              km[w] = (dy < 0 ? m_up[w] : (dy > 0 ? m_dn[w] : 0u)) | (kx == 0 ? m_lf[w] : (kx == 2 ? m_rt[w] : 0u));
            for (int t = 0; t < c.T_acc; ++t) {
              if constexpr (I8) {
                const uint32_t d_lo = d + (uint32_t)c.N_TILE;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) {
                  const uint64_t a1 = ((uint64_t)desc_hi << 32) | (a_lo + ks * a_step_k);
                  const uint64_t a128 = ((uint64_t)desc_hi << 32) | (a_lo + a_off_128 + ks * a_step_k);
                  const uint64_t bd0 = ((uint64_t)desc_hi << 32) | (b_lo0 + ks * b_step_k);
                  const uint64_t bd1 = ((uint64_t)desc_hi << 32) | (b_lo0 + b_step_sp + ks * b_step_k);
                  const uint64_t bd2 = ((uint64_t)desc_hi << 32) | (b_lo0 + 2 * b_step_sp + ks * b_step_k);
                  const uint32_t acc0 = ks != 0 ? 1u : first;
                  if (PAIR) {
                    tc_mma_i8_masked_pair(d, a128, bd0, p.idesc, acc0, km);        // hi += (128 s) x d0
                    tc_mma_i8_masked_pair(d, a1, bd1, p.idesc, 1u, km);            // hi += s x d1
                    tc_mma_i8_masked_pair(d_lo, a1, bd2, p.idesc, acc0, km);       // lo += s x d2
                  } else {
                    tc_mma_i8_masked(d, a128, bd0, p.idesc, acc0, km[0], km[1], km[2], km[3]);
                    tc_mma_i8_masked(d, a1, bd1, p.idesc, 1u, km[0], km[1], km[2], km[3]);
                    tc_mma_i8_masked(d_lo, a1, bd2, p.idesc, acc0, km[0], km[1], km[2], km[3]);
                  }
                }
              } else {
#pragma unroll
                for (int sp = 0; sp < NSPLIT; ++sp) {
#pragma unroll
                  for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint64_t adesc = ((uint64_t)desc_hi << 32) | (a_lo + ks * a_step_k);
                    const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (b_lo0 + sp * b_step_sp + ks * b_step_k);
                    if (PAIR) tc_mma_f16_masked_pair(d, adesc, bdesc, p.idesc, (sp | ks) != 0 ? 1u : first, km);
                    else if (dbg & 4) tc_mma_f16(d, adesc, bdesc, p.idesc, (sp | ks) != 0 ? 1u : first);   // timing experiment
                    else tc_mma_f16_masked(d, adesc, bdesc, p.idesc, (sp | ks) != 0 ? 1u : first, km[0], km[1], km[2], km[3]);
                  }
                }
              }
              a_lo += a_step_t;
              d += (uint32_t)(ACCS * c.N_TILE);
            }
            if (PAIR) { tc_commit_pair(b_empty(sb.stage)); if (ti == 8) tc_commit_pair(a_empty(sa.stage)); }
            else { tc_commit(b_empty(sb.stage)); if (ti == 8) tc_commit(a_empty(sa.stage)); }
          }
          __syncwarp();
          sb.advance(c.b_stages);
        }
        sa.advance(c.a_stages);
      }
      if (elect_one()) { if (PAIR) tc_commit_pair(acc_full(sc.stage)); else tc_commit(acc_full(sc.stage)); }
      __syncwarp();
      if (trace && lane == 0 && trace_it < 7) {
        trace[3 + trace_it * 8 + 2] = clock64();
        trace[3 + trace_it * 8 + 3] = stall;
      }
      ++trace_it;
      sc.advance(c.acc_stages);
    }
  } else if (warp < kEpiWarps) {
    // ===== epilogue =====
    PipeState sc;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    // the kEpiWarps / 4 warps of a lane quarter share the tile's 16-column groups
    constexpr int kParts = kEpiWarps / 4;
    const int n_groups = c.N_TILE >> 4;
    const int col_lo = 16 * ((n_groups * (warp >> 2)) / kParts);
    const int col_hi = 16 * ((n_groups * ((warp >> 2) + 1)) / kParts);
    const float inv_T = 1.0f / (float)p.T;
    // x / tau == x * (1/tau) bit for bit when tau is a power of two (the reference uses tau = 2)
    int tau_exp;
    const bool tau_pow2 = frexpf(p.tau, &tau_exp) == 0.5f;
    const float inv_tau = 1.0f / p.tau;
    const bool fast_lif = tau_pow2 && p.hard_reset && p.v_reset == 0.f;
    int trace_it = 0;
    pdl_wait();                                      // before the first read of state / counts and the first store
    for (int tile = unit0, pass = 0; tile < total_tiles;
         (++pass == unit_passes) ? (pass = 0, tile += unit_stride) : 0) {
      const int tch = pass0_of(tile) + pass;
      const bool first_pass = tch == 0, last_pass = tch == c.n_tchunks - 1;
      const int n0 = (tile_of(tile) % c.n_tiles) * c.N_TILE;
      const int64_t r = (int64_t)m_tile_of(tile) * kTileRows + q * 32 + lane;  // row (without guard)
      const int pp = (int)(r % p.P);
      const int py = pp / p.W, px = pp - py * p.W;
      const bool valid = r < p.R_valid;
      // Stage the N tile's folded-BN affine in shared memory while the MMAs of this tile run (the first pass of a tile
      // does it; the buffers alternate between tiles, and the named barrier below orders fill -> use).
      const int tile_par = ((tile - unit0) / unit_stride) & 1;
      float* const sS = s_affine + tile_par * 512;
      float* const sH = sS + 256;
      if (pass == 0) {
        const int e = threadIdx.x;                    // 256 epilogue threads: one column each (N_TILE <= 256)
        if (e < c.N_TILE) {
          const bool in = n0 + e < p.Cout;
          sS[e] = in ? __ldg(p.scale + n0 + e) : 0.f;
          sH[e] = in ? __ldg(p.shift + n0 + e) : 0.f;
        }
      }
      // lean i8 LIF epilogue (below): everything that is fixed for the pass, computed while the MMAs of the pass still run
      const uint32_t t_base0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc.stage * c.T_acc * ACCS * c.N_TILE);
      const bool lean = I8 && p.out_kind == SD_OUT_LIF && !c.tpar && fast_lif && n0 + col_lo < p.Cout;
      const bool v_via_smem = s_v != nullptr;
      const bool want_sum = p.out_sum != nullptr, want_spk = p.out_spk8 != nullptr;
      const int T_acc = c.T_acc;
      const uint32_t n_tile = (uint32_t)c.N_TILE;
      const float vth = p.v_th;
      const int64_t row_off = p.G + r;
      int64_t grp_bytes = p.R_alloc * 16;                       // next 16-channel chunk of an STF8 plane
      int64_t plane = (int64_t)(p.Cout8 >> 1) * grp_bytes;      // from the s plane to the 128*s plane of a timestep
      asm volatile("" : "+l"(grp_bytes), "+l"(plane));          // keep them in registers (no re-derivation per store)
      const int64_t chunk0 = (int64_t)((n0 + col_lo) >> 3) * p.R_alloc + row_off;   // 8-channel chunk index of group 0
      uint8_t* o_grp = p.out_spk8 + (int64_t)(tch * T_acc * 2) * plane + (chunk0 - ((int64_t)((n0 + col_lo) >> 4)) * p.R_alloc) * 16;
      float* v_grp = p.v != nullptr ? p.v + chunk0 * 8 : nullptr;            // state plane [C/8][R_alloc][8] fp32
      __half* sum_grp = want_sum ? p.out_sum + chunk0 * 8 : nullptr;         // T-sums [C/8][R_alloc][8] fp16
      const int64_t chunk_stride = p.R_alloc * 8;                            // elements to the next 8-channel chunk
      float* sv = s_v + threadIdx.x;                                         // [column][thread]
      const float* aS = sS + col_lo;
      const float* aH = sH + col_lo;
      uint32_t t_grp = t_base0 + (uint32_t)col_lo;
      const int last_col = min(col_hi, p.Cout - n0);                         // groups with n < Cout
      const bool load_v_global = p.v != nullptr && valid && (!first_pass || p.v_load_initial) && !(v_via_smem && !first_pass);
      const bool load_sum = !first_pass && want_sum && valid;
      const bool store_v_global = p.v != nullptr && valid && (v_via_smem ? (last_pass && p.v_store_final) : (!last_pass || p.v_store_final));
      asm volatile("" : "+l"(o_grp), "+l"(v_grp), "+l"(sum_grp), "+r"(t_grp));   // ... and kept above the wait
      asm volatile("" ::"r"(n0), "r"(pp), "r"(py), "r"(px));   // keep the tile's index arithmetic (divisions) above the wait
      mbar_wait(acc_full(sc.stage), sc.phase);
      tc_fence_after();
      if (pass == 0) asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory");
      if (trace && threadIdx.x == 0 && trace_it < 7) trace[3 + trace_it * 8 + 4] = clock64();
      const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sc.stage * c.T_acc * ACCS * c.N_TILE);
      // accumulator of local timestep tl, 16 columns from cc: one TMEM load (fp32), or for i8 two (int32 hi and lo)
      // that conv_value() merges into the fp32 convolution value in place: (float)(256 * hi + lo), a single rounding of
      // the exact integer sum (the per-channel scaling guarantees that 256 * hi + lo fits an int32)
      auto ld_acc = [&](int tl, int cc, uint32_t (&a)[16], uint32_t (&lo)[16]) {
        tc_ld16(t_base + (uint32_t)(tl * ACCS * c.N_TILE + cc), a);
        if constexpr (I8) tc_ld16(t_base + (uint32_t)((tl * ACCS + 1) * c.N_TILE + cc), lo);
      };
      auto conv_value = [&](uint32_t (&a)[16], uint32_t (&lo)[16]) {
        tc_ld_wait_on(a);
        if constexpr (I8) {
          tc_ld_wait_on(lo);
#pragma unroll
          for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__int2float_rn((int)a[j] * 256 + (int)lo[j]));
        }
      };
      // ---- kind::i8, LIF with hard reset to 0 and tau a power of two (every layer of the denoiser): lean path ----
      // Same arithmetic as the general code below, but everything that does not change inside a pass is computed once
      // (addresses advance by pointer increments; no parameter re-reads, no per-group 64-bit index arithmetic): the general
      // code spends ~140 instructions per 16-column group and ~45 per timestep on that, a quarter of the epilogue.
      bool epilogue_done = false;
      if constexpr (I8) {
        if (lean) {
          epilogue_done = true;
          for (int cc = col_lo; cc < last_col; cc += 16) {
            if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4)] = clock64();
            float sc_[16], sh_[16], v[16];
#pragma unroll
            for (int j = 0; j < 16; j += 4) {       // warp-uniform shared-memory reads (broadcast)
              const float4 a = *reinterpret_cast<const float4*>(aS + j);
              const float4 b = *reinterpret_cast<const float4*>(aH + j);
              sc_[j] = a.x; sc_[j + 1] = a.y; sc_[j + 2] = a.z; sc_[j + 3] = a.w;
              sh_[j] = b.x; sh_[j + 1] = b.y; sh_[j + 2] = b.z; sh_[j + 3] = b.w;
            }
            if (v_via_smem && !first_pass) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = sv[j * kEpiThreads];
            } else if (load_v_global) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const float4 a = *reinterpret_cast<const float4*>(v_grp + h * chunk_stride);
                const float4 b = *reinterpret_cast<const float4*>(v_grp + h * chunk_stride + 4);
                v[8 * h] = a.x; v[8 * h + 1] = a.y; v[8 * h + 2] = a.z; v[8 * h + 3] = a.w;
                v[8 * h + 4] = b.x; v[8 * h + 5] = b.y; v[8 * h + 6] = b.z; v[8 * h + 7] = b.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            uint4 sum_raw[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};   // fp16 counts of the earlier passes
            if (load_sum) {
              sum_raw[0] = *reinterpret_cast<const uint4*>(sum_grp);
              sum_raw[1] = *reinterpret_cast<const uint4*>(sum_grp + chunk_stride);
            }
            uint32_t cnt8[4] = {0u, 0u, 0u, 0u};     // this pass's counts, one byte per column
            if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 1] = clock64();
            if (trace && threadIdx.x == 0 && trace_it < 7 && cc == col_lo) trace[3 + trace_it * 8 + 6] = clock64();
            uint8_t* o = o_grp;
            uint32_t ta = t_grp;
            for (int tl = 0; tl < T_acc; ++tl) {
              uint32_t hi[16], lo[16];
              tc_ld16(ta, hi);
              tc_ld16(ta + n_tile, lo);
              ta += 2 * n_tile;
              tc_ld_wait_on(hi);
              tc_ld_wait_on(lo);
              uint32_t packed[4] = {0u, 0u, 0u, 0u};
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                // (float)(256 * hi + lo): one rounding of the exact integer sum; then BN affine, charge, fire, reset
                const float x = fmaf(__int2float_rn((int)hi[j] * 256 + (int)lo[j]), sc_[j], sh_[j]);
                const float h = fmaf(__fsub_rn(x, v[j]), inv_tau, v[j]);
                const uint32_t pat = 1u << (8 * (j & 3));
                asm("{\n\t"
                    ".reg .pred q;\n\t"
                    "setp.ge.f32 q, %2, %3;\n\t"
                    "selp.f32 %0, 0f00000000, %2, q;\n\t"
                    "@q or.b32 %1, %1, %4;\n\t"
                    "}"
                    : "=f"(v[j]), "+r"(packed[j >> 2])
                    : "f"(h), "f"(vth), "r"(pat));
              }
#pragma unroll
              for (int k = 0; k < 4; ++k) cnt8[k] += packed[k];     // four byte counters per word (at most T <= 16)
              if (valid && want_spk) {
                *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint4*>(o + plane) = make_uint4(packed[0] << 7, packed[1] << 7, packed[2] << 7, packed[3] << 7);
              }
              o += 2 * plane;
            }
            if (trace && threadIdx.x == 0 && trace_it < 7 && cc == col_lo) trace[3 + trace_it * 8 + 7] = clock64();
            if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 2] = clock64();
            if (want_sum && valid) {   // byte counters of this pass -> fp16, added to the earlier passes' counts (exact)
              __half2* cnt2 = reinterpret_cast<__half2*>(sum_raw);
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const uint32_t w = cnt8[k >> 1] >> (16 * (k & 1));
                cnt2[k] = __hadd2(cnt2[k], __halves2half2(__ushort2half_rn((unsigned short)(w & 0xFFu)),
                                                          __ushort2half_rn((unsigned short)((w >> 8) & 0xFFu))));
              }
              *reinterpret_cast<uint4*>(sum_grp) = sum_raw[0];
              *reinterpret_cast<uint4*>(sum_grp + chunk_stride) = sum_raw[1];
            }
            if (v_via_smem && !last_pass) {
#pragma unroll
              for (int j = 0; j < 16; ++j) sv[j * kEpiThreads] = v[j];
            } else if (store_v_global) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                *reinterpret_cast<float4*>(v_grp + h * chunk_stride) = make_float4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
                *reinterpret_cast<float4*>(v_grp + h * chunk_stride + 4) = make_float4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
              }
            }
            if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 3] = clock64();
            o_grp += grp_bytes;
            if (v_grp != nullptr) v_grp += 2 * chunk_stride;
            if (want_sum) sum_grp += 2 * chunk_stride;
            sv += 16 * kEpiThreads;
            aS += 16; aH += 16;
            t_grp += 16;
          }
        }
      }
      for (int cc = col_lo; cc < col_hi && !epilogue_done; cc += 16) {
        const int n = n0 + cc;  // first output channel of this 16-column group (warp-uniform)
        if (n >= p.Cout) continue;  // zero-padded tail of the last N tile
        if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4)] = clock64();
        float sc_[16], sh_[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {       // warp-uniform shared-memory reads (broadcast)
          const float4 a = *reinterpret_cast<const float4*>(sS + cc + j);
          const float4 b = *reinterpret_cast<const float4*>(sH + cc + j);
          sc_[j] = a.x; sc_[j + 1] = a.y; sc_[j + 2] = a.z; sc_[j + 3] = a.w;
          sh_[j] = b.x; sh_[j + 1] = b.y; sh_[j + 2] = b.z; sh_[j + 3] = b.w;
        }
        if (c.tpar) {
          // T-parallel mode: this pass only delivers the input currents x[t] = conv * scale + shift of its timesteps
          for (int tl = 0; tl < c.T_acc; ++tl) {
            uint32_t acc[16], acc_lo[16];
            ld_acc(tl, cc, acc, acc_lo);
            conv_value(acc, acc_lo);
            if (valid) {
              const int t = tch * c.T_acc + tl;
              float* o = p.cur + (((int64_t)t * p.Cout8 + (n >> 3)) * p.R_alloc + p.G + r) * 8;
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float4 lo, hi;
                lo.x = fmaf(__uint_as_float(acc[8 * h + 0]), sc_[8 * h + 0], sh_[8 * h + 0]);
                lo.y = fmaf(__uint_as_float(acc[8 * h + 1]), sc_[8 * h + 1], sh_[8 * h + 1]);
                lo.z = fmaf(__uint_as_float(acc[8 * h + 2]), sc_[8 * h + 2], sh_[8 * h + 2]);
                lo.w = fmaf(__uint_as_float(acc[8 * h + 3]), sc_[8 * h + 3], sh_[8 * h + 3]);
                hi.x = fmaf(__uint_as_float(acc[8 * h + 4]), sc_[8 * h + 4], sh_[8 * h + 4]);
                hi.y = fmaf(__uint_as_float(acc[8 * h + 5]), sc_[8 * h + 5], sh_[8 * h + 5]);
                hi.z = fmaf(__uint_as_float(acc[8 * h + 6]), sc_[8 * h + 6], sh_[8 * h + 6]);
                hi.w = fmaf(__uint_as_float(acc[8 * h + 7]), sc_[8 * h + 7], sh_[8 * h + 7]);
                *reinterpret_cast<float4*>(o + (int64_t)h * p.R_alloc * 8) = lo;
                *reinterpret_cast<float4*>(o + (int64_t)h * p.R_alloc * 8 + 4) = hi;
              }
            }
          }
        } else if (p.out_kind == SD_OUT_LIF) {
          float v[16];
          __half2 cnt2[8];   // spike counts so far (exact in fp16: at most T <= 16)
          const bool want_sum = p.out_sum != nullptr;
          const int64_t vrow = ((int64_t)(n >> 3) * p.R_alloc + p.G + r) * 8;  // chunk n/8; next chunk + R_alloc*8
          // between the passes of a tile the potential stays in shared memory ([column][thread]: conflict-free); the
          // state plane in global memory is only read for a caller-provided initial state and written for the caller
          const bool v_via_smem = s_v != nullptr;
          float* const sv = s_v + (cc - col_lo) * kEpiThreads + threadIdx.x;     // this thread's 16 columns, stride kEpiThreads
          if (v_via_smem && !first_pass) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = sv[j * kEpiThreads];
          } else if (p.v != nullptr && valid && n < p.Cout && (!first_pass || p.v_load_initial)) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 a = *reinterpret_cast<const float4*>(p.v + vrow + (int64_t)h * p.R_alloc * 8);
              const float4 b = *reinterpret_cast<const float4*>(p.v + vrow + (int64_t)h * p.R_alloc * 8 + 4);
              v[8 * h] = a.x; v[8 * h + 1] = a.y; v[8 * h + 2] = a.z; v[8 * h + 3] = a.w;
              v[8 * h + 4] = b.x; v[8 * h + 5] = b.y; v[8 * h + 6] = b.z; v[8 * h + 7] = b.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = p.hard_reset ? p.v_reset : 0.f;
          }
          uint32_t cnt8[4] = {0u, 0u, 0u, 0u};   // i8 path: this pass's counts, one byte per column
#pragma unroll
          for (int k = 0; k < 8; ++k) cnt2[k] = __floats2half2_rn(0.f, 0.f);
          if (!first_pass && want_sum && valid) {   // spike counts of the earlier passes
            const __half* o = p.out_sum + ((int64_t)(n >> 3) * p.R_alloc + p.G + r) * 8;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint4 raw = *reinterpret_cast<const uint4*>(o + (int64_t)h * p.R_alloc * 8);
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
              for (int k = 0; k < 4; ++k) cnt2[4 * h + k] = h2[k];
            }
          }
          // One timestep of 16 columns: BN affine, charge, fire, reset.  The spike is kept as an all-ones / zero mask:
          // the hard reset to 0 is h AND NOT mask (exactly s ? 0 : h) and one LOP3 per spike drops its bit pattern into
          // the packed output word (fp16 1.0 = 0x3C00 per half, or 0x01 per byte for the u8 format).
          auto lif_step = [&](auto fast_tag, const uint32_t (&acc)[16], int tl) {
            constexpr bool kFast = decltype(fast_tag)::value;
            constexpr int kWords = I8 ? 4 : 8;
            uint32_t packed[kWords];
#pragma unroll
            for (int k = 0; k < kWords; ++k) packed[k] = 0u;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x = fmaf(__uint_as_float(acc[j]), sc_[j], sh_[j]);
              uint32_t m;
              if constexpr (kFast) {
                // hard reset to 0, tau a power of two: h = v + (x - v) * (1/tau) is one exact-product FMA
                // (inline PTX: one predicate per neuron-timestep drives the reset select and a predicated OR of the spike's
                // bit pattern into the packed output word -- 3 instructions; left to itself the compiler spends 4-5)
                const float h = fmaf(__fsub_rn(x, v[j]), inv_tau, v[j]);
                const uint32_t pat = I8 ? (1u << (8 * (j & 3))) : (0x3C00u << (16 * (j & 1)));
                asm("{\n\t"
                    ".reg .pred q;\n\t"
                    "setp.ge.f32 q, %2, %3;\n\t"
                    "selp.f32 %0, 0f00000000, %2, q;\n\t"
                    "@q or.b32 %1, %1, %4;\n\t"
                    "}"
                    : "=f"(v[j]), "+r"(packed[I8 ? (j >> 2) : (j >> 1)])
                    : "f"(h), "f"(p.v_th), "r"(pat));
                m = 0u;
              } else {
                const float dv = p.hard_reset ? __fsub_rn(x, __fsub_rn(v[j], p.v_reset)) : __fsub_rn(x, v[j]);
                const float h = __fadd_rn(v[j], tau_pow2 ? __fmul_rn(dv, inv_tau) : __fdiv_rn(dv, p.tau));
                const bool s_ = h >= p.v_th;
                m = s_ ? 0xFFFFFFFFu : 0u;
                v[j] = p.hard_reset ? (s_ ? p.v_reset : h) : (s_ ? __fsub_rn(h, p.v_th) : h);
              }
              if constexpr (!kFast) {
                if constexpr (I8) packed[j >> 2] |= m & (1u << (8 * (j & 3)));
                else packed[j >> 1] |= m & (0x3C00u << (16 * (j & 1)));
              }
            }
            if (want_sum) {
              if constexpr (I8) {
#pragma unroll
                for (int k = 0; k < 4; ++k) cnt8[k] += packed[k];     // four byte counters per word (at most T <= 16)
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) cnt2[k] = __hadd2(cnt2[k], *reinterpret_cast<const __half2*>(&packed[k]));
              }
            }
            if constexpr (I8) {
              if (valid && n < p.Cout && p.out_spk8 != nullptr) {
                // one 16-byte row = 16 channels; the 128*s plane of the same timestep lies Cout/16 planes further
                const int t = tch * c.T_acc + tl;
                const int64_t plane = (int64_t)(p.Cout8 >> 1) * p.R_alloc * 16;
                uint8_t* o = p.out_spk8 + (int64_t)(t * 2) * plane + (((int64_t)(n >> 4)) * p.R_alloc + p.G + r) * 16;
                *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                *reinterpret_cast<uint4*>(o + plane) = make_uint4(packed[0] << 7, packed[1] << 7, packed[2] << 7, packed[3] << 7);
              }
            } else if (valid && n < p.Cout && p.out_spk != nullptr && !(dbg & 1)) {
              const int t = tch * c.T_acc + tl;
              __half* o = p.out_spk + (((int64_t)t * p.Cout8 + (n >> 3)) * p.R_alloc + p.G + r) * 8;
              *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
              *reinterpret_cast<uint4*>(o + p.R_alloc * 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
            }
          };
          // two register sets: the TMEM load of timestep t + 1 is in flight while timestep t is computed
          auto lif_all = [&](auto fast_tag) {
            if constexpr (I8) {
              // one register set: with hi + lo per timestep a second set pushes the epilogue to the 168-register ceiling
              // and the scheduler loses more to the serialised code than the TMEM load latency costs (cycle-stamp
              // trace: 9.5 k -> 8.1 k cycles per pass)
              uint32_t accA[16], loA[16];
              for (int tl = 0; tl < c.T_acc; ++tl) {
                ld_acc(tl, cc, accA, loA);
                conv_value(accA, loA);
                lif_step(fast_tag, accA, tl);
              }
              return;
            }
            uint32_t accA[16], accB[16], loA[16], loB[16];
            const bool ld = !(dbg & 2);
#pragma unroll
            for (int j = 0; j < 16; ++j) accA[j] = accB[j] = loA[j] = loB[j] = 0u;
            if (ld) ld_acc(0, cc, accA, loA);
            for (int tl = 0; tl < c.T_acc; tl += 2) {
              conv_value(accA, loA);
              if (tl + 1 < c.T_acc && ld) ld_acc(tl + 1, cc, accB, loB);
              lif_step(fast_tag, accA, tl);
              if (tl + 1 < c.T_acc) {
                conv_value(accB, loB);
                if (tl + 2 < c.T_acc && ld) ld_acc(tl + 2, cc, accA, loA);
                lif_step(fast_tag, accB, tl + 1);
              }
            }
          };
          if (trace && threadIdx.x == 0 && trace_it < 7 && cc == col_lo) trace[3 + trace_it * 8 + 6] = clock64();   // state loaded
          if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 1] = clock64();
          if (fast_lif) lif_all(std::true_type{}); else lif_all(std::false_type{});
          if (trace && threadIdx.x == 0 && trace_it < 7 && cc == col_lo) trace[3 + trace_it * 8 + 7] = clock64();   // LIF of group 0 done
          if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 2] = clock64();
          if (valid && n < p.Cout) {
            if (want_sum) {
              if constexpr (I8) {   // byte counters of this pass -> fp16, added to the earlier passes' counts (exact)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  const uint32_t w = cnt8[k >> 1] >> (16 * (k & 1));
                  cnt2[k] = __hadd2(cnt2[k], __halves2half2(__ushort2half_rn((unsigned short)(w & 0xFFu)),
                                                            __ushort2half_rn((unsigned short)((w >> 8) & 0xFFu))));
                }
              }
              const uint32_t* pk = reinterpret_cast<const uint32_t*>(cnt2);
              __half* o = p.out_sum + ((int64_t)(n >> 3) * p.R_alloc + p.G + r) * 8;
              *reinterpret_cast<uint4*>(o) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(o + p.R_alloc * 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
            if (v_via_smem && !last_pass) {
#pragma unroll
              for (int j = 0; j < 16; ++j) sv[j * kEpiThreads] = v[j];
            } else if (p.v != nullptr && (v_via_smem ? p.v_store_final : (!last_pass || p.v_store_final))) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                float* o = p.v + vrow + (int64_t)h * p.R_alloc * 8;
                *reinterpret_cast<float4*>(o) = make_float4(v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3]);
                *reinterpret_cast<float4*>(o + 4) = make_float4(v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]);
              }
            }
          }
          if (trace && threadIdx.x == 0 && trace_it == 1) trace[64 + 4 * ((cc - col_lo) >> 4) + 3] = clock64();
        } else {
          // SD_OUT_MEAN_T on a T-summed input: (conv(sum_t s_t) * scale + T * shift) / T, channels-last fp32
          uint32_t acc[16];
          tc_ld16(t_base + (uint32_t)cc, acc);
          tc_ld_wait();
          if (valid && n < p.Cout) {
            const int64_t img = r / p.P;
            float* o = p.out_real + ((img * p.H + py) * p.W + px) * (int64_t)p.Cout + n;
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 y;
              y.x = __fmul_rn(fmaf(__uint_as_float(acc[j]), sc_[j], __fmul_rn(sh_[j], (float)p.T)), inv_T);
              y.y = __fmul_rn(fmaf(__uint_as_float(acc[j + 1]), sc_[j + 1], __fmul_rn(sh_[j + 1], (float)p.T)), inv_T);
              y.z = __fmul_rn(fmaf(__uint_as_float(acc[j + 2]), sc_[j + 2], __fmul_rn(sh_[j + 2], (float)p.T)), inv_T);
              y.w = __fmul_rn(fmaf(__uint_as_float(acc[j + 3]), sc_[j + 3], __fmul_rn(sh_[j + 3], (float)p.T)), inv_T);
              *reinterpret_cast<float4*>(o + j) = y;
            }
          }
        }
      }
      // release the accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR && !leader) mbar_arrive_remote(acc_empty(sc.stage), 0); else mbar_arrive(acc_empty(sc.stage)); }
      if (trace && threadIdx.x == 0 && trace_it < 7) trace[3 + trace_it * 8 + 5] = clock64();
      ++trace_it;
      sc.advance(c.acc_stages);
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (trace && threadIdx.x == 0) trace[2] = clock64();
  if (warp == kWarpAux) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

// Second kernel of the T-parallel mode: the LIF recurrence over all T on the currents written by the convolution
// passes.  One thread = one pixel row x 8 channels (32-byte current loads, 16-byte spike stores, lanes on consecutive
// rows); the arithmetic is the fused epilogue's, so both modes produce the same bits.
// The kernel runs once per layer and diffusion step on a cold instruction cache with one short block per SM, so it is
// built for few instructions and few dependent round trips: FAST (hard reset to 0, tau a power of two) is a template
// parameter, the t loop is rolled in chunks of 4 timesteps, and the currents of the next chunk are in flight while the
// current one is integrated (ncu on the first version: 60 % of the warp samples were instruction-fetch stalls, the rest a
// chain of T dependent L2 round trips; 16-19 us per launch whatever the layer size).
template <bool FAST>
__global__ void __launch_bounds__(256) lif_from_currents_kernel(const TcParams p) {
  constexpr int kChunk = 4;
  pdl_launch_dependents();
  pdl_wait();
  const float inv_tau = 1.0f / p.tau;
  int tau_exp;
  const bool tau_pow2 = frexpf(p.tau, &tau_exp) == 0.5f;
  const int64_t total = (int64_t)p.Cout8 * p.R_valid;
  const int64_t t_stride = (int64_t)p.Cout8 * p.R_alloc * 8;
  const int64_t plane8 = (int64_t)(p.Cout8 >> 1) * p.R_alloc * 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ch = i / p.R_valid, r = i - ch * p.R_valid;
    const int64_t off = (ch * p.R_alloc + p.G + r) * 8;
    const float* cur = p.cur + off;
    float4 ca[kChunk], cb[kChunk], na[kChunk], nb[kChunk];
#pragma unroll
    for (int k = 0; k < kChunk; ++k) {
      if (k < p.T) {
        ca[k] = __ldg(reinterpret_cast<const float4*>(cur + k * t_stride));
        cb[k] = __ldg(reinterpret_cast<const float4*>(cur + k * t_stride + 4));
      }
    }
    float v[8];
    if (p.v != nullptr && p.v_load_initial) {
      const float4 a = *reinterpret_cast<const float4*>(p.v + off), b = *reinterpret_cast<const float4*>(p.v + off + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = p.hard_reset ? p.v_reset : 0.f;
    }
    __half2 cnt2[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt2[k] = __floats2half2_rn(0.f, 0.f);
    __half* o16 = p.out_spk != nullptr ? p.out_spk + off : nullptr;
    uint8_t* o8 = p.out_spk8 != nullptr ? p.out_spk8 + ((ch >> 1) * p.R_alloc + p.G + r) * 16 + (ch & 1) * 8 : nullptr;
#pragma unroll 1
    for (int t0 = 0; t0 < p.T; t0 += kChunk) {
#pragma unroll
      for (int k = 0; k < kChunk; ++k) {
        if (t0 + kChunk + k < p.T) {
          na[k] = __ldg(reinterpret_cast<const float4*>(cur + (t0 + kChunk + k) * t_stride));
          nb[k] = __ldg(reinterpret_cast<const float4*>(cur + (t0 + kChunk + k) * t_stride + 4));
        }
      }
#pragma unroll
      for (int k = 0; k < kChunk; ++k) {
        if (t0 + k < p.T) {
          const float x[8] = {ca[k].x, ca[k].y, ca[k].z, ca[k].w, cb[k].x, cb[k].y, cb[k].z, cb[k].w};
          uint32_t packed[4];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            float sf[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if constexpr (FAST) {
                const float h = fmaf(__fsub_rn(x[j + u], v[j + u]), inv_tau, v[j + u]);
                sf[u] = h >= p.v_th ? 1.f : 0.f;
                v[j + u] = fmaf(-h, sf[u], h);
              } else {
                const float dv = p.hard_reset ? __fsub_rn(x[j + u], __fsub_rn(v[j + u], p.v_reset)) : __fsub_rn(x[j + u], v[j + u]);
                const float h = __fadd_rn(v[j + u], tau_pow2 ? __fmul_rn(dv, inv_tau) : __fdiv_rn(dv, p.tau));
                const bool s = h >= p.v_th;
                sf[u] = s ? 1.f : 0.f;
                v[j + u] = p.hard_reset ? (s ? p.v_reset : h) : (s ? __fsub_rn(h, p.v_th) : h);
              }
            }
            const __half2 s2 = __floats2half2_rn(sf[0], sf[1]);
            packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&s2);
            cnt2[j >> 1] = __hadd2(cnt2[j >> 1], s2);
          }
          if (o16 != nullptr) {
            *reinterpret_cast<uint4*>(o16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            o16 += t_stride;
          }
          if (o8 != nullptr) {
            // STF8: this thread's 8 channels are one half of a 16-byte row of the s plane and of the 128*s plane
            const uint32_t w0 = __byte_perm((packed[0] >> 10) & 0x00010001u, (packed[1] >> 10) & 0x00010001u, 0x6420);
            const uint32_t w1 = __byte_perm((packed[2] >> 10) & 0x00010001u, (packed[3] >> 10) & 0x00010001u, 0x6420);
            *reinterpret_cast<uint2*>(o8) = make_uint2(w0, w1);
            *reinterpret_cast<uint2*>(o8 + plane8) = make_uint2(w0 << 7, w1 << 7);
            o8 += 2 * plane8;
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kChunk; ++k) { ca[k] = na[k]; cb[k] = nb[k]; }
    }
    if (p.out_sum != nullptr) {
      const uint32_t* pk = reinterpret_cast<const uint32_t*>(cnt2);
      *reinterpret_cast<uint4*>(p.out_sum + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    if (p.v != nullptr && p.v_store_final) {
      *reinterpret_cast<float4*>(p.v + off) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p.v + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Configuration (shared by weight packing and launch)
// ---------------------------------------------------------------------------------------------------
// Experiment knobs (SD_TC_*): read from the environment ONCE per process (first plan build); sd_debug_tc_reload_knobs
// re-reads them for the sweep tool (tools/bench_layers.py).  Defaults are the measured best
// (profiles/r01_experiments.md); none of them changes results except through fp32 summation order (KBLK):
//   SD_TC_PAIR=0            single-CTA kernel instead of 2-CTA pairs
//   SD_TC_SMALL_BATCH_SPLIT=0  keep N = 128 for a lone small batch
//   SD_TC_TCHUNK=0 / SD_TC_TACC=n   one pass over all T / n timesteps per pass (n = 2: two TMEM stages)
//   SD_TC_N256, SD_TC_WIDE256       N = 256 tiles with 2 timesteps per pass
//   SD_TC_NTILE, SD_TC_KBLK, SD_TC_ACC_STAGES, SD_TC_ALIGN   tile shape overrides
//   SD_TC_PERSIST=0         one work unit per CTA / cluster
//   SD_TC_TPAR=0            no T-parallel small-batch mode
//   SD_TC_VSMEM=0           multi-pass tiles carry the membrane potential through the L2 state plane instead of smem
//   SD_TC_DBG (-DSD_TRACE build only, with sd_debug_tc_trace)   1: skip spike stores, 2: skip TMEM loads, 4: un-masked
//                           MMAs in the single-CTA kernel (timing experiments only; 4 gives wrong results at borders)
struct TcKnobs {
  int pair, small_batch_split, tchunk, tacc, wide256, n256, ntile, kblk, acc_stages, align, persist, tpar, dbg, v_smem;
};
static TcKnobs g_knobs;
static std::once_flag g_knobs_once;
static void load_knobs() {
  auto env_int = [](const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
  };
  TcKnobs k;
  k.pair = env_int("SD_TC_PAIR", 1);
  k.small_batch_split = env_int("SD_TC_SMALL_BATCH_SPLIT", 1);
  k.tchunk = env_int("SD_TC_TCHUNK", 1);
  k.tacc = env_int("SD_TC_TACC", 0);
  k.wide256 = env_int("SD_TC_WIDE256", 0);
  k.n256 = env_int("SD_TC_N256", 0);
  k.ntile = env_int("SD_TC_NTILE", 0);
  k.kblk = env_int("SD_TC_KBLK", 0);
  k.acc_stages = env_int("SD_TC_ACC_STAGES", -1);
  k.align = env_int("SD_TC_ALIGN", 0);
  k.persist = env_int("SD_TC_PERSIST", 1);
  k.tpar = env_int("SD_TC_TPAR", 1);
  k.dbg = env_int("SD_TC_DBG", 0);
  k.v_smem = env_int("SD_TC_VSMEM", 1);
  g_knobs = k;
}
static const TcKnobs& knobs() {
  std::call_once(g_knobs_once, load_knobs);
  return g_knobs;
}

static int tc_supported(const sd_conv_desc* d, const char** why) {
  *why = "";
  if (d->transposed || d->kh != 3 || d->kw != 3 || d->stride != 1 || d->pad != 1) { *why = "only 3x3 stride 1 pad 1"; return 0; }
  const bool i8 = d->nsplit == 3;
  if (i8) {
    if (d->in_kind != SD_IN_STF8) { *why = "nsplit = 3 (int8 digits) takes STF8 (u8) spikes"; return 0; }
    if (d->out_kind != SD_OUT_LIF8 && d->out_kind != SD_OUT_CURRENT_SEQ) {
      *why = "nsplit = 3 (int8 digits) writes STF8 spikes (out_kind LIF8) or planar currents (CURRENT_SEQ)";
      return 0;
    }
    if (d->T % 2 || d->in_T != d->T || d->T > 16) { *why = "int8 path needs an even T <= 16 and in_T == T"; return 0; }
    if (d->C_in0 != d->C_in || d->C_in % 32 || d->C_out % 16) { *why = "int8 path needs C_in % 32 == 0, C_out % 16 == 0, one input segment"; return 0; }
    if (d->H_in != d->H_out || d->W_in != d->W_out) { *why = "output grid must equal input grid"; return 0; }
    if (d->W_in + 2 > 64) { *why = "grid too wide for the row-shift window"; return 0; }
    return 1;
  }
  if (d->in_kind != SD_IN_STF) { *why = "input must be STF spikes"; return 0; }
  if (d->H_in != d->H_out || d->W_in != d->W_out) { *why = "output grid must equal input grid"; return 0; }
  if (d->out_kind != SD_OUT_LIF && d->out_kind != SD_OUT_MEAN_T) { *why = "out_kind must be LIF or MEAN_T"; return 0; }
  if (d->out_kind == SD_OUT_LIF && (d->in_T != d->T || d->T > 16)) { *why = "LIF needs in_T == T <= 16"; return 0; }
  if (d->out_kind == SD_OUT_MEAN_T && d->in_T != 1) { *why = "MEAN_T needs a T-summed input (in_T == 1)"; return 0; }
  const int c0 = d->C_in0, c1 = d->C_in - d->C_in0;
  if (c0 % 16 || c1 % 16) { *why = "channel segments must be multiples of 16"; return 0; }
  if (d->C_out % 16) { *why = "C_out must be a multiple of 16"; return 0; }
  if (d->nsplit < 1 || d->nsplit > 2) { *why = "nsplit must be 1 or 2 (fp16 terms) or 3 (int8 digits)"; return 0; }
  if (d->W_in + 2 > 64) { *why = "grid too wide for the row-shift window"; return 0; }
  return 1;
}

// kind::i8 configuration: 2 timesteps per pass (hi + lo accumulators of 2 timesteps fill the 512 TMEM columns at
// N = 128), T / 2 passes with the membrane potential carried through the state plane, K block 64 (or 32) channels.
static int tc_config_i8(const sd_conv_desc* d, TcConfig* c) {
  check_device();
  const int sms = sm_count() > 0 ? sm_count() : 148;
  // 2 timesteps per pass fill TMEM with hi + lo accumulators (one stage); SD_TC_TACC=1: one timestep per pass and two
  // accumulator stages (the epilogue of a pass overlaps the MMAs of the next) at twice the weight traffic
  c->T_acc = knobs().tacc == 1 ? 1 : 2;
  c->n_tchunks = d->T / c->T_acc;
  c->ndx = 1;
  const int64_t rows = (int64_t)d->B * d->H_in * d->W_in;
  const int m_tiles = (int)((rows + kTileRows - 1) / kTileRows);
  int n_tile = 128;
  const int conc = d->concurrent > 1 ? d->concurrent : 1;
  const bool can_pair = knobs().pair && m_tiles >= 2;
  // a single M tile cannot be paired: narrower N tiles spread it over more SMs (same rule as the fp16 path)
  while (n_tile > 32 && knobs().small_batch_split && !can_pair &&
         (int64_t)m_tiles * ((d->C_out + n_tile - 1) / n_tile) * 2 * conc <= sms && d->C_out > n_tile / 2)
    n_tile /= 2;
  c->tpar = 0;
  if (c->n_tchunks > 1 && d->concurrent <= 1 && knobs().tpar) {
    const int64_t pair_units = (((int64_t)m_tiles + 1) / 2) * ((d->C_out + 127) / 128);
    if (pair_units * 4 <= sms) { c->tpar = 1; n_tile = 128; }
  }
  // SD_OUT_CURRENT_SEQ (training branch: convolution only): every (tile, pass) is an independent unit whose epilogue
  // writes x[t] = conv * scale + shift, exactly the first half of the T-parallel mode
  if (d->out_kind == SD_OUT_CURRENT_SEQ) c->tpar = 1;
  if (knobs().ntile > 0) n_tile = knobs().ntile;
  while (n_tile > 32 && n_tile / 2 >= d->C_out) n_tile /= 2;
  if (!(n_tile == 32 || n_tile == 64 || n_tile == 128)) { set_error("conv_tc(i8): bad N tile %d", n_tile); return SD_ERR_UNSUPPORTED; }
  c->N_TILE = n_tile;
  c->pair = (knobs().pair && n_tile == 128 && m_tiles >= 2) ? 1 : 0;
  c->acc_stages = 512 / (c->T_acc * 2 * n_tile) >= 2 ? 2 : 1;
  c->halo = d->W_in + 1;
  c->rows_ld = kTileRows + 2 * c->halo;
  bool found = false;
  static const int kblks[2] = {64, 32};
  for (int i = 0; i < 2 && !found; ++i) {
    const int kblk = kblks[i];
    if (knobs().kblk && kblk != knobs().kblk) continue;
    if (d->C_in % kblk) continue;
    c->a_stage_bytes = (uint32_t)c->T_acc * 2 * (kblk / 16) * c->rows_ld * 16;
    c->b_stage_bytes = 3u * (kblk / 16) * (c->pair ? n_tile / 2 : n_tile) * 16;   // per CTA
    const uint32_t b_ref = 3u * (kblk / 16) * 128 * 16;      // batch-independent fit test (see the fp16 path)
    if (2 * c->a_stage_bytes + 4 * b_ref + kHdrBytes > kSmemBudget) continue;
    c->KBLK = kblk;
    found = true;
  }
  if (!found) { set_error("conv_tc(i8): no K block fits shared memory"); return SD_ERR_UNSUPPORTED; }
  // multi-pass tiles keep the membrane potential in shared memory between their passes when that still leaves
  // >= 2 A stages and >= 4 B stages (else it travels through the L2-resident state plane, as in T-parallel mode)
  uint32_t v_smem = (c->n_tchunks > 1 && !c->tpar && knobs().v_smem) ? kVSmemBytes : 0;
  for (;;) {
    c->a_stages = (int)(((int64_t)kSmemBudget - kHdrBytes - v_smem - 4 * (int64_t)c->b_stage_bytes) / c->a_stage_bytes);
    if (c->a_stages > 3) c->a_stages = 3;
    if (c->a_stages >= 2 || v_smem == 0) break;
    v_smem = 0;
  }
  const uint32_t left = kSmemBudget - kHdrBytes - v_smem - c->a_stages * c->a_stage_bytes;
  c->b_stages = (int)(left / c->b_stage_bytes);
  if (c->b_stages > kMaxBStages) c->b_stages = kMaxBStages;
  c->v_smem_off = v_smem ? kHdrBytes + c->a_stages * c->a_stage_bytes + c->b_stages * c->b_stage_bytes : 0;
  c->smem_bytes = kHdrBytes + c->a_stages * c->a_stage_bytes + c->b_stages * c->b_stage_bytes + v_smem;
  if (c->smem_bytes < 117u * 1024u) c->smem_bytes = 117u * 1024u;
  c->n_tiles = (d->C_out + n_tile - 1) / n_tile;
  c->m_tiles = m_tiles;
  c->c0_blocks = d->C_in / c->KBLK;
  c->num_kblocks = d->C_in / c->KBLK;
  return SD_OK;
}

static int tc_config(const sd_conv_desc* d, TcConfig* c) {
  const char* why;
  if (!tc_supported(d, &why)) { set_error("conv_tc: unsupported descriptor: %s", why); return SD_ERR_UNSUPPORTED; }
  const bool i8 = d->nsplit == 3;
  c->i8 = i8 ? 1 : 0;
  c->accs = i8 ? 2 : 1;
  c->rowch = i8 ? 16 : 8;
  if (i8) return tc_config_i8(d, c);
  c->T_acc = d->out_kind == SD_OUT_LIF ? d->T : 1;
  if (d->out_kind == SD_OUT_LIF && d->T > 4 && d->T % 4 == 0 && knobs().tchunk) c->T_acc = 4;
  if (d->out_kind == SD_OUT_LIF) {
    const int tacc = knobs().tacc;           // experiment knob: timesteps per pass
    if (tacc > 0 && tacc <= d->T && d->T % tacc == 0) c->T_acc = tacc;
  }
  // Layers with exactly 256 output channels have only two N = 128 tiles per M tile, which quantises badly on 148 SMs
  // (98 M tiles -> 196 tiles -> 2 waves, the second one a third full).  One N = 256 tile per M tile with 2 timesteps
  // per pass does the same work in a single wave and fetches the A operand half as often per output column.
  bool wide = false;
  if (d->out_kind == SD_OUT_LIF && d->C_out == 256 && d->T % 2 == 0 && d->T <= 4 && knobs().wide256) {
    c->T_acc = 2;
    wide = true;
  }
  c->n_tchunks = d->out_kind == SD_OUT_LIF ? d->T / c->T_acc : 1;
  int n_tile;
  // Larger N amortises the A-operand fetch from shared memory (4 KB per MMA whatever N is): measured on B200,
  // N = 128 without epilogue overlap beats N = 64 with two TMEM stages (profiles/).
  if (c->T_acc * 256 <= 512 && d->C_out >= 256 && (wide || knobs().n256)) n_tile = 256;
  else if (c->T_acc * 128 <= 512) n_tile = 128;
  else if (c->T_acc * 64 <= 512) n_tile = 64;
  else n_tile = 32;
  {
    // Small batches: with N = 128 there are fewer tiles than SMs (the reference samples 16 images per call: 7 M tiles).
    // Halve N while that at least doubles the number of busy SMs; each tile then also gets a second TMEM stage.
    check_device();
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int64_t rows = (int64_t)d->B * d->H_in * d->W_in;
    const int m_tiles = (int)((rows + kTileRows - 1) / kTileRows);
    const int conc = d->concurrent > 1 ? d->concurrent : 1;   // sub-batches on other streams fill the SMs instead
    // Only when the tiles cannot be paired (a single M tile): the cycle-stamp trace shows that an un-paired MMA costs
    // 72-79 cycles whatever N is (reading the 4 KB A operand from shared memory bounds it) while a paired N = 128 MMA does 8x the
    // work of an N = 32 one in 64 cycles, and the K loop per CTA - the critical path of a small batch - is the same.
    const bool can_pair = knobs().pair && m_tiles >= 2 && n_tile == 128;
    while (n_tile > 32 && knobs().small_batch_split && !can_pair &&
           (int64_t)m_tiles * ((d->C_out + n_tile - 1) / n_tile) * 2 * conc <= sms && d->C_out > n_tile / 2)
      n_tile /= 2;
  }
  // T-parallel small-batch mode: a multi-pass layer whose full-size (N = 128, paired) tiles would occupy less than
  // half of the clusters runs its passes as independent work units (x T/4 parallelism at full tile efficiency) and
  // leaves the LIF recurrence to a second kernel.  The per-CTA K loop, which bounds a small batch, gets T/4 x shorter.
  c->tpar = 0;
  if (d->out_kind == SD_OUT_LIF && c->n_tchunks > 1 && d->concurrent <= 1 && knobs().tpar) {
    check_device();
    const int sms = sm_count() > 0 ? sm_count() : 148;
    const int64_t rows = (int64_t)d->B * d->H_in * d->W_in;
    const int64_t m_tiles = (rows + kTileRows - 1) / kTileRows;
    const int64_t pair_units = ((m_tiles + 1) / 2) * ((d->C_out + 127) / 128);
    if (pair_units * 4 <= sms && c->T_acc * 128 <= 512) {
      c->tpar = 1;
      n_tile = 128;
    }
  }
  if (knobs().ntile > 0) n_tile = knobs().ntile;
  while (n_tile > 32 && n_tile / 2 >= d->C_out) n_tile /= 2;
  if (!(n_tile == 32 || n_tile == 64 || n_tile == 128 || n_tile == 256) || c->T_acc * n_tile > 512) {
    set_error("conv_tc: bad N tile %d for T=%d", n_tile, c->T_acc);
    return SD_ERR_UNSUPPORTED;
  }
  c->N_TILE = n_tile;
  {
    const int64_t rows = (int64_t)d->B * d->H_in * d->W_in;
    const int m_tiles = (int)((rows + kTileRows - 1) / kTileRows);
    c->pair = (knobs().pair && n_tile == 128 && m_tiles >= 2) ? 1 : 0;
  }
  c->acc_stages = 512 / (c->T_acc * n_tile) >= 2 ? 2 : 1;
  if (knobs().acc_stages >= 0) c->acc_stages = knobs().acc_stages >= 2 && 512 / (c->T_acc * n_tile) >= 2 ? 2 : 1;
  const int c0 = d->C_in0, c1 = d->C_in - d->C_in0;
  const int Wp = d->W_in;   // row stride of the dense pixel grid
  // Measured on B200 (profiles/r1_layer_sweep.txt): the pre-shifted (128-byte aligned) variant is SLOWER than plain
  // 16-byte-aligned tap starts (its 3x larger A stage forces KBLK = 16), so it is opt-in only.
  const int want_align = knobs().align;
  const int kblk_pref = knobs().kblk;
  bool found = false;
  for (int ndx = (want_align && Wp % 8 == 0) ? 3 : 1; ndx >= 1 && !found; ndx -= 2) {
    c->ndx = ndx;
    c->halo = ndx == 3 ? Wp : Wp + 1;
    c->rows_ld = kTileRows + 2 * c->halo;
    static const int kblks[3] = {64, 32, 16};
    for (int i = 0; i < 3 && !found; ++i) {
      const int kblk = kblks[i];
      if (kblk_pref && kblk != kblk_pref) continue;   // default: the largest K block that leaves >= 2 A + 4 B stages
      if (c0 % kblk || c1 % kblk) continue;
      c->a_stage_bytes = (uint32_t)c->T_acc * (kblk / 8) * ndx * c->rows_ld * 16;
      c->b_stage_bytes = (uint32_t)d->nsplit * (kblk / 8) * (c->pair ? n_tile / 2 : n_tile) * 16;   // per CTA
      // The K block fixes the fp32 accumulation order ((channel block, tap, 16-channel step)), so it must not depend
      // on the batch size: the fit test uses the N = 128 stage size even when a small batch runs narrower tiles.
      // A batch shard then reproduces the unsharded result bit for bit.
      const uint32_t b_ref = (uint32_t)d->nsplit * (kblk / 8) * (n_tile > 128 ? n_tile : 128) * 16;
      if (2 * c->a_stage_bytes + 4 * b_ref + kHdrBytes > kSmemBudget) continue;
      c->KBLK = kblk;
      found = true;
    }
  }
  if (!found) { set_error("conv_tc: no K block fits shared memory"); return SD_ERR_UNSUPPORTED; }
  const int kblk = c->KBLK;
  // A ring first (2..4 K blocks in flight), the rest of the budget to the B ring (one tap per stage)
  uint32_t v_smem = (c->n_tchunks > 1 && !c->tpar && n_tile <= 128 && knobs().v_smem) ? kVSmemBytes : 0;
  for (;;) {
    c->a_stages = (int)(((int64_t)kSmemBudget - kHdrBytes - v_smem - 4 * (int64_t)c->b_stage_bytes) / c->a_stage_bytes);
    if (c->a_stages > kMaxAStages) c->a_stages = kMaxAStages;
    if (c->a_stages > 3) c->a_stages = 3;
    if (c->a_stages >= 2 || v_smem == 0) break;
    v_smem = 0;
  }
  uint32_t left = kSmemBudget - kHdrBytes - v_smem - c->a_stages * c->a_stage_bytes;
  c->b_stages = (int)(left / c->b_stage_bytes);
  if (c->b_stages > kMaxBStages) c->b_stages = kMaxBStages;
  c->v_smem_off = v_smem ? kHdrBytes + c->a_stages * c->a_stage_bytes + c->b_stages * c->b_stage_bytes : 0;
  c->smem_bytes = kHdrBytes + c->a_stages * c->a_stage_bytes + c->b_stages * c->b_stage_bytes + v_smem;
  // every CTA allocates all 512 TMEM columns: ask for more than half of the SM's shared memory so that two CTAs of
  // this kernel are never co-resident (the second would only sit in tcgen05.alloc until the first one exits)
  if (c->smem_bytes < 117u * 1024u) c->smem_bytes = 117u * 1024u;
  c->n_tiles = (d->C_out + n_tile - 1) / n_tile;
  const int64_t R = (int64_t)d->B * d->H_in * d->W_in;
  c->m_tiles = (int)((R + kTileRows - 1) / kTileRows);
  c->c0_blocks = c0 / kblk;
  c->num_kblocks = c0 / kblk + c1 / kblk;
  return SD_OK;
}

// ---- weight packing -------------------------------------------------------------------------------
// exponent e[co] such that max|w[co]| * 2^e lies in [2^13, 2^14); chan_scale = 2^-e
__global__ void tc_chan_exp_kernel(const float* __restrict__ w, int Cin, float* __restrict__ chan_scale, int* __restrict__ e_out) {
  const int co = blockIdx.x;
  float m = 0.f;
  for (int i = threadIdx.x; i < Cin * 9; i += blockDim.x) m = fmaxf(m, fabsf(w[(int64_t)co * Cin * 9 + i]));
  __shared__ float red[256];
  red[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int e = 0;
    if (red[0] > 0.f && isfinite(red[0])) {
      int ex;
      frexpf(red[0], &ex);  // red = f * 2^ex, f in [0.5, 1)  ->  red * 2^(14 - ex) in [2^13, 2^14)
      e = 14 - ex;
    }
    e_out[co] = e;
    chan_scale[co] = ldexpf(1.0f, -e);
  }
}

// layout [n_tile][k_block][tap][half][split][chunk][n (N_TILE / halves)][8]; halves = 2 for the cta_group::2 variant
__global__ void tc_pack_kernel(const float* __restrict__ w, const int* __restrict__ e_in, __half* __restrict__ out,
                               int Cout, int Cin, int N_TILE, int KBLK, int nsplit, int n_tiles, int num_kblocks,
                               int halves) {
  const int chunks = KBLK / 8;
  const int NH = N_TILE / halves;
  const int64_t total = (int64_t)n_tiles * num_kblocks * 9 * nsplit * chunks * N_TILE * 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i % 8);
    int64_t r = i / 8;
    int n = (int)(r % NH); r /= NH;
    int ch = (int)(r % chunks); r /= chunks;
    int sp = (int)(r % nsplit); r /= nsplit;
    int hf = (int)(r % halves); r /= halves;
    int tap = (int)(r % 9); r /= 9;
    int kb = (int)(r % num_kblocks);
    int nt = (int)(r / num_kblocks);
    const int co = nt * N_TILE + hf * NH + n;
    const int ci = kb * KBLK + ch * 8 + j;
    float val = 0.f;
    if (co < Cout) {
      const float ws = ldexpf(w[((int64_t)co * Cin + ci) * 9 + tap], e_in[co]);  // exact power-of-two scaling
      const __half hi = __float2half_rn(ws);
      val = sp == 0 ? __half2float(hi) : __fsub_rn(ws, __half2float(hi));
    }
    out[i] = __float2half_rn(val);
  }
}

// ---- int8 digits -----------------------------------------------------------------------------------------------
// Per output channel the weights become 22-bit fixed point: w_fix = rint(w * 2^e) with e the largest exponent such that
// max|w_fix| < 2^21 and sum_k |w_fix_k| < 2^31 (then 256 * hi + lo = sum of the active w_fix can never overflow an
// int32, whatever fires).  w_fix = 256 * (128 * d0 + d1) + d2 with d0, d2 in [-128, 127], d1 in [-64, 63].
// chan_scale = 2^-e is folded into the BN scale by the caller (exact).
__global__ void tc_chan_exp_i8_kernel(const float* __restrict__ w, int Cin, float* __restrict__ chan_scale, int* __restrict__ e_out) {
  const int co = blockIdx.x;
  float m = 0.f;
  double l1 = 0.0;
  for (int i = threadIdx.x; i < Cin * 9; i += blockDim.x) {
    const float a = fabsf(w[(int64_t)co * Cin * 9 + i]);
    m = fmaxf(m, a);
    l1 += (double)a;
  }
  __shared__ float red[256];
  __shared__ double red1[256];
  red[threadIdx.x] = m;
  red1[threadIdx.x] = l1;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
      red1[threadIdx.x] += red1[threadIdx.x + s];      // fixed tree order: deterministic
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    int e = 0;
    if (red[0] > 0.f && isfinite(red[0])) {
      int ex;
      frexpf(red[0], &ex);              // max = f * 2^ex, f in [0.5, 1)  ->  max * 2^(21 - ex) in [2^20, 2^21)
      e = 21 - ex;
      // L1 bound with the rounding slack of every term: sum|w| * 2^e + 0.5 * K < 2^31
      const double room = 2147483647.0 - 0.5 * (double)Cin * 9.0 - 1.0;
      while (ldexp(red1[0], e) >= room) --e;
    }
    e_out[co] = e;
    chan_scale[co] = ldexpf(1.0f, -e);
  }
}

// layout [n_tile][k_block][tap][half][digit 3][chunk (16 channels)][n (N_TILE / halves)][16 bytes]
__global__ void tc_pack_i8_kernel(const float* __restrict__ w, const int* __restrict__ e_in, int8_t* __restrict__ out,
                                  int Cout, int Cin, int N_TILE, int KBLK, int n_tiles, int num_kblocks, int halves) {
  const int chunks = KBLK / 16;
  const int NH = N_TILE / halves;
  const int64_t total = (int64_t)n_tiles * num_kblocks * 9 * 3 * chunks * N_TILE * 16;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int j = (int)(i % 16);
    int64_t r = i / 16;
    int n = (int)(r % NH); r /= NH;
    int ch = (int)(r % chunks); r /= chunks;
    int dg = (int)(r % 3); r /= 3;
    int hf = (int)(r % halves); r /= halves;
    int tap = (int)(r % 9); r /= 9;
    int kb = (int)(r % num_kblocks);
    int nt = (int)(r / num_kblocks);
    const int co = nt * N_TILE + hf * NH + n;
    const int ci = kb * KBLK + ch * 16 + j;
    int val = 0;
    if (co < Cout) {
      const int wf = (int)rintf(ldexpf(w[((int64_t)co * Cin + ci) * 9 + tap], e_in[co]));   // exact scaling, one rounding
      const int d2 = ((wf + 128) & 255) - 128;                 // balanced low digit in [-128, 127]
      const int hi = (wf - d2) / 256;                          // exact
      const int d1 = ((hi + 64) & 127) - 64;                   // [-64, 63]
      const int d0 = (hi - d1) / 128;                          // [-128, 127] because |wf| < 2^21
      val = dg == 0 ? d0 : (dg == 1 ? d1 : d2);
    }
    out[i] = (int8_t)val;
  }
}

}  // namespace sd

using namespace sd;

#ifdef SD_TRACE
static long long* g_tc_trace = nullptr;   // diagnostics build only (tools/trace_tc.py); not thread-safe by design
#endif

extern "C" {

// Diagnostics: the next sd_conv_lif_tc launches write per-CTA cycle stamps to `buf` (device memory, at least
// grid * 64 int64; pass null to switch tracing off).  Layout per CTA: [0] entry, [1] set-up done, [2] exit, then per
// tile pass i < 7 at 3 + 8 i: MMA warp got the accumulator, first operands landed, last MMA issued, cycles the MMA warp
// waited for operands, epilogue saw the accumulator, epilogue released it.
// Only the -DSD_TRACE build of the library (tools/trace_tc.py builds it) carries the stamps; the shipped one refuses.
int sd_debug_tc_trace(void* buf) {
#ifdef SD_TRACE
  g_tc_trace = (long long*)buf;
  return SD_OK;
#else
  if (buf == nullptr) return SD_OK;
  set_error("sd_debug_tc_trace: this library was built without -DSD_TRACE (tools/trace_tc.py builds the diagnostics variant)");
  return SD_ERR_UNSUPPORTED;
#endif
}

// Diagnostics: re-read the SD_TC_* environment knobs (tools/bench_layers.py sweeps them inside one process).
int sd_debug_tc_reload_knobs(void) {
  knobs();
  load_knobs();
  return SD_OK;
}

int sd_conv_tc_supported(const sd_conv_desc* d) {
  if (!d || validate_conv_desc(d) != SD_OK) return 0;
  const char* why;
  if (!tc_supported(d, &why)) return 0;
  TcConfig c;
  return tc_config(d, &c) == SD_OK ? 1 : 0;
}

int64_t sd_conv_workspace_bytes(const sd_conv_desc* d) {
  TcConfig c;
  if (!d || validate_conv_desc(d) != SD_OK || tc_config(d, &c) != SD_OK) return 0;
  if (d->out_kind == SD_OUT_CURRENT_SEQ) return 0;                     // the currents go straight to args.out
  if (c.n_tchunks <= 1 || (!c.tpar && c.v_smem_off != 0)) return 0;   // potential carried in shared memory
  // one fp32 state plane [C_out/8][R_alloc][8]; in T-parallel mode one plane of currents per timestep instead
  const int64_t plane = (int64_t)c8(d->C_out) * stf_rows(d->B, d->H_out, d->W_out) * 8 * (int64_t)sizeof(float);
  return c.tpar ? plane * d->T : plane;
}

int64_t sd_conv_weight_layout_tc(const sd_conv_desc* d) {
  TcConfig c;
  if (!d || validate_conv_desc(d) != SD_OK || tc_config(d, &c) != SD_OK) return -1;
  return ((int64_t)c.N_TILE << 16) | ((int64_t)c.KBLK << 4) | ((int64_t)c.i8 << 1) | (int64_t)c.pair;
}

int64_t sd_conv_weight_bytes_tc(const sd_conv_desc* d) {
  TcConfig c;
  if (!d || validate_conv_desc(d) != SD_OK || tc_config(d, &c) != SD_OK) return 0;
  // + C_out ints of scratch for the per-channel exponents at the end
  return (int64_t)c.n_tiles * c.num_kblocks * 9 * c.b_stage_bytes * (c.pair ? 2 : 1) + (int64_t)d->C_out * 4;
}

int sd_conv_pack_weights_tc(const sd_conv_desc* d, const float* w, void* packed, float* chan_scale_out, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  TcConfig c;
  rc = tc_config(d, &c);
  if (rc) return rc;
  SD_REQUIRE(w && packed && chan_scale_out, "null pointer argument");
  SD_DEVICE_OR_RETURN();
  cudaStream_t st = as_stream(stream);
  const int64_t main_bytes = (int64_t)c.n_tiles * c.num_kblocks * 9 * c.b_stage_bytes * (c.pair ? 2 : 1);
  int* e_buf = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(packed) + main_bytes);
  if (c.i8) {
    tc_chan_exp_i8_kernel<<<d->C_out, 256, 0, st>>>(w, d->C_in, chan_scale_out, e_buf);
    SD_LAUNCH_CHECK();
    int64_t bl8 = (main_bytes + 255) / 256;
    if (bl8 > (int64_t)sm_count() * 8) bl8 = (int64_t)sm_count() * 8;
    tc_pack_i8_kernel<<<(unsigned)bl8, 256, 0, st>>>(w, e_buf, (int8_t*)packed, d->C_out, d->C_in, c.N_TILE, c.KBLK,
                                                     c.n_tiles, c.num_kblocks, c.pair ? 2 : 1);
    SD_LAUNCH_CHECK();
    return SD_OK;
  }
  tc_chan_exp_kernel<<<d->C_out, 256, 0, st>>>(w, d->C_in, chan_scale_out, e_buf);
  SD_LAUNCH_CHECK();
  const int64_t total = main_bytes / 2;
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  tc_pack_kernel<<<(unsigned)blocks, 256, 0, st>>>(w, e_buf, (__half*)packed, d->C_out, d->C_in, c.N_TILE, c.KBLK,
                                                   d->nsplit, c.n_tiles, c.num_kblocks, c.pair ? 2 : 1);
  SD_LAUNCH_CHECK();
  return SD_OK;
}

int sd_conv_lif_tc(const sd_conv_desc* d, const sd_conv_args* a, void* stream) {
  int rc = validate_conv_desc(d);
  if (rc) return rc;
  TcConfig c;
  rc = tc_config(d, &c);
  if (rc) return rc;
  SD_REQUIRE(a && a->in && a->weights && a->scale && a->shift && a->out, "null pointer argument");
  if (d->C_in0 < d->C_in) SD_REQUIRE(a->in2 != nullptr, "conv_tc: in2 missing for concat");
  SD_REQUIRE((((uintptr_t)a->in | (uintptr_t)a->in2 | (uintptr_t)a->weights | (uintptr_t)a->out | (uintptr_t)a->out_sum |
               (uintptr_t)a->v | (uintptr_t)a->workspace | (uintptr_t)a->scale | (uintptr_t)a->shift) & 15) == 0,
             "conv_tc: every buffer must be 16-byte aligned (cp.async.bulk / 128-bit accesses)");
  SD_DEVICE_OR_RETURN();
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.in0 = (const __half*)a->in;
  p.in1 = (const __half*)a->in2;
  p.wpack = (const __half*)a->weights;
  p.scale = a->scale;
  p.shift = a->shift;
  p.v = a->v ? a->v : (float*)a->workspace;
  p.v_load_initial = a->v != nullptr;
  p.v_store_final = a->v != nullptr;
  const bool currents_only = d->out_kind == SD_OUT_CURRENT_SEQ;
  if (currents_only) {
    p.cur = (float*)a->out;
    p.v = nullptr;
  } else if (c.tpar) {
    SD_REQUIRE(a->workspace != nullptr, "conv_tc: this configuration needs args.workspace (sd_conv_workspace_bytes)");
    p.cur = (float*)a->workspace;
    p.v = a->v;
  }
  if (c.n_tchunks > 1 && !c.tpar && p.v == nullptr && c.v_smem_off == 0) {
    set_error("conv_tc: T=%d runs as %d passes and needs args.v or args.workspace (sd_conv_workspace_bytes)", d->T,
              c.n_tchunks);
    return SD_ERR_INVALID;
  }
  if (d->out_kind == SD_OUT_LIF) { p.out_spk = (__half*)a->out; p.out_sum = (__half*)a->out_sum; }
  else if (d->out_kind == SD_OUT_LIF8) { p.out_spk8 = (uint8_t*)a->out; p.out_sum = (__half*)a->out_sum; }
  else if (!currents_only) p.out_real = (float*)a->out;
  StfGeom g(d->B, d->H_in, d->W_in);
  p.R_alloc = g.R_alloc; p.G = g.G; p.R_valid = (int64_t)d->B * g.P;
  p.C8_0 = d->C_in0 / c.rowch; p.C8_1 = (d->C_in - d->C_in0) / c.rowch;   // 16-byte-row chunks per input segment
  p.Cout = d->C_out; p.Cout8 = c8(d->C_out);
  p.T = d->T; p.H = d->H_in; p.W = d->W_in; p.Wp = g.Wp; p.P = g.P;
  p.nsplit = d->nsplit; p.out_kind = (d->out_kind == SD_OUT_LIF8 || currents_only) ? SD_OUT_LIF : d->out_kind; p.hard_reset = d->hard_reset;
  p.tau = d->tau; p.v_th = d->v_threshold; p.v_reset = d->v_reset;
  // cute::UMMA::InstrDescriptor: c_format F32 (bit 4), a/b F16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
  // kind::i8: c_format S32 (2 at bit 4), a_format u8 (0 at bit 7), b_format s8 (1 at bit 10)
  p.idesc = (c.i8 ? ((2u << 4) | (1u << 10)) : (1u << 4)) | ((uint32_t)(c.N_TILE >> 3) << 17) |
            ((uint32_t)((c.pair ? 2 * kTileRows : kTileRows) >> 4) << 24);
  p.c = c;
#ifdef SD_TRACE
  p.trace = g_tc_trace;
  p.dbg = g_tc_trace ? knobs().dbg : 0;
#endif
  cudaStream_t st = as_stream(stream);
  int grid;
  const bool persist = knobs().persist != 0;   // experiment knob: 0 = one work unit per CTA / cluster
  const int unit_mult = c.tpar ? c.n_tchunks : 1;
  if (c.pair) {
    const int units = ((c.m_tiles + 1) / 2) * c.n_tiles * unit_mult;
    grid = 2 * (units < sm_count() / 2 || !persist ? units : sm_count() / 2);
  } else {
    grid = c.m_tiles * c.n_tiles * unit_mult;
    if (grid > sm_count() && persist) grid = sm_count();
  }
#define SD_TC_LAUNCH_ONE(NS, KS, PR)                                                                               \
  do {                                                                                                             \
    /* the attribute belongs to the function in ONE device's context: set it once per device */                   \
    static std::once_flag once[64];                                                                                \
    static cudaError_t attr_rc[64];                                                                                \
    const int dev_ = current_device_index();                                                                       \
    std::call_once(once[dev_], [dev_] {                                                                            \
      attr_rc[dev_] = cudaFuncSetAttribute(conv3x3_tc_kernel<NS, KS, PR>,                                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);               \
    });                                                                                                            \
    SD_CUDA(attr_rc[dev_]);                                                                                        \
    cudaLaunchConfig_t cfg = {};                                                                                   \
    cfg.gridDim = dim3((unsigned)grid);                                                                            \
    cfg.blockDim = dim3(kTcThreads);                                                                               \
    cfg.dynamicSmemBytes = c.smem_bytes;                                                                           \
    cfg.stream = st;                                                                                               \
    cudaLaunchAttribute attr[2];                                                                                   \
    attr[0].id = cudaLaunchAttributeClusterDimension;                                                              \
    attr[0].val.clusterDim.x = PR ? 2 : 1;                                                                         \
    attr[0].val.clusterDim.y = 1;                                                                                  \
    attr[0].val.clusterDim.z = 1;                                                                                  \
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                               \
    attr[1].val.programmaticStreamSerializationAllowed = 1;                                                        \
    cfg.attrs = attr;                                                                                              \
    cfg.numAttrs = sd_pdl_enabled(d->concurrent) ? 2 : 1;                                                          \
    SD_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NS, KS, PR>, p));                                           \
  } while (0)
#define SD_TC_LAUNCH(NS, KS)                                    \
  do {                                                          \
    if (c.pair) SD_TC_LAUNCH_ONE(NS, KS, true);                 \
    else SD_TC_LAUNCH_ONE(NS, KS, false);                       \
  } while (0)
  const int ks = c.i8 ? c.KBLK / 32 : c.KBLK / 16;
  if (d->nsplit == 3) {
    if (ks == 1) SD_TC_LAUNCH(3, 1); else SD_TC_LAUNCH(3, 2);
  } else if (d->nsplit == 1) {
    if (ks == 1) SD_TC_LAUNCH(1, 1); else if (ks == 2) SD_TC_LAUNCH(1, 2); else SD_TC_LAUNCH(1, 4);
  } else {
    if (ks == 1) SD_TC_LAUNCH(2, 1); else if (ks == 2) SD_TC_LAUNCH(2, 2); else SD_TC_LAUNCH(2, 4);
  }
#undef SD_TC_LAUNCH
#undef SD_TC_LAUNCH_ONE
  SD_LAUNCH_CHECK();
  if (c.tpar && !currents_only) {
    const int64_t n = (int64_t)p.Cout8 * p.R_valid;
    int64_t bl = (n + 255) / 256;
    if (bl > (int64_t)sm_count() * 8) bl = (int64_t)sm_count() * 8;
    int tau_exp;
    const bool fast = frexpf(p.tau, &tau_exp) == 0.5f && p.hard_reset && p.v_reset == 0.f;
    cudaLaunchConfig_t lcfg = {};
    lcfg.gridDim = dim3((unsigned)bl);
    lcfg.blockDim = dim3(256);
    lcfg.stream = st;
    cudaLaunchAttribute lattr[1];
    lattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    lattr[0].val.programmaticStreamSerializationAllowed = 1;
    lcfg.attrs = lattr;
    lcfg.numAttrs = sd_pdl_enabled(d->concurrent) ? 1 : 0;
    if (fast) SD_CUDA(cudaLaunchKernelEx(&lcfg, lif_from_currents_kernel<true>, p));
    else SD_CUDA(cudaLaunchKernelEx(&lcfg, lif_from_currents_kernel<false>, p));
    SD_LAUNCH_CHECK();
  }
  return SD_OK;
}

}  // extern "C"
