// Probe of tcgen05.mma.kind::i8 on sm_100a (GPU box):  nvcc -gencode arch=compute_100a,code=sm_100a -o probe_i8 probe_i8.cu
//  1. correctness of a single M=128, N=128, K=32 u8 x s8 -> s32 MMA on no-swizzle K-major operands (the layout the
//     conv kernel uses: core matrix = 8 rows x 16 B, SBO = 128 B, LBO = plane stride), including the u8 value 128
//     and the disable-output-lane mask;
//  2. issue rate: cycles per MMA for kind::i8 (K = 32) against kind::f16 (K = 16) at the same operand bytes.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         ((uint64_t)1 << 46);
}

// kind: 0 = f16 (K = 16 halves), 1 = i8 (K = 32 bytes).  Operands: A [128 rows][32 B], B [128 rows][32 B] as two planes
// of 16-byte rows each (plane stride 2048 B).  reps MMAs accumulate into the same D.
template <int KIND>
__global__ void __launch_bounds__(128, 1) probe(const uint8_t* __restrict__ a_g, const uint8_t* __restrict__ b_g,
                                                uint32_t* __restrict__ d_g, long long* __restrict__ cyc, int reps,
                                                uint32_t mask0) {
  __shared__ __align__(1024) uint8_t a_s[4096];
  __shared__ __align__(1024) uint8_t b_s[4096];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096 / 16; i += 128) {
    reinterpret_cast<uint4*>(a_s)[i] = reinterpret_cast<const uint4*>(a_g)[i];
    reinterpret_cast<uint4*>(b_s)[i] = reinterpret_cast<const uint4*>(b_g)[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // idesc: c_format [4,6): F32 = 1, S32 = 2;  a_format [7,10), b_format [10,13): f16 = 0 / u8 = 0, s8 = 1;  N>>3 at [17,23), M>>4 at [24,29)
  const uint32_t idesc = KIND == 0 ? ((1u << 4) | (16u << 17) | (8u << 24))
                                   : ((2u << 4) | (0u << 7) | (1u << 10) | (16u << 17) | (8u << 24));
  if (threadIdx.x == 0) {
    const uint64_t ad = make_desc(smem_u32(a_s), 2048, 128), bd = make_desc(smem_u32(b_s), 2048, 128);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t acc = r != 0;
      if (KIND == 0)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(mask0), "r"(0u), "r"(0u), "r"(0u) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                     ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(mask0), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(smem_u32(&bar), 0);
    cyc[0] = clock64() - t0;
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < 128; c0 += 16) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) d_g[(size_t)threadIdx.x * 128 + c0 + j] = r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128u));
}

// canonical no-swizzle K-major offset of byte kb of row r (two 16-byte planes)
static size_t canon(int r, int kb) { return (size_t)(kb / 16) * 2048 + (size_t)r * 16 + kb % 16; }

int main() {
  std::vector<uint8_t> a(4096), b(4096);
  std::vector<int> A(128 * 32), B(128 * 32);
  srand(1);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 32; ++k) {
      const int av = (rand() % 4 == 0) ? ((k & 1) ? 128 : 1) : 0;     // u8 spikes: 1 or 128
      const int bv = rand() % 256 - 128;                              // s8 digits
      A[r * 32 + k] = av; B[r * 32 + k] = bv;
      a[canon(r, k)] = (uint8_t)av; b[canon(r, k)] = (uint8_t)(int8_t)bv;
    }
  uint8_t *a_d, *b_d; uint32_t* d_d; long long* c_d;
  CK(cudaMalloc(&a_d, 4096)); CK(cudaMalloc(&b_d, 4096)); CK(cudaMalloc(&d_d, 128 * 128 * 4)); CK(cudaMalloc(&c_d, 8));
  CK(cudaMemcpy(a_d, a.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(b_d, b.data(), 4096, cudaMemcpyHostToDevice));
  std::vector<int32_t> d(128 * 128);
  for (uint32_t mask : {0u, 0x0000F00Fu}) {
    CK(cudaMemset(d_d, 0xEE, 128 * 128 * 4));
    probe<1><<<1, 128>>>(a_d, b_d, d_d, c_d, 1, mask);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(d.data(), d_d, 128 * 128 * 4, cudaMemcpyDeviceToHost));
    long bad = 0, checked = 0;
    for (int r = 0; r < 128; ++r) {
      if (r < 32 && ((mask >> r) & 1)) continue;       // disabled rows keep whatever TMEM held
      for (int n = 0; n < 128; ++n) {
        long ref = 0;
        for (int k = 0; k < 32; ++k) ref += (long)A[r * 32 + k] * B[n * 32 + k];
        ++checked;
        if (d[r * 128 + n] != (int32_t)ref) { if (bad < 5) printf("  mismatch r=%d n=%d got %d want %ld\n", r, n, d[r * 128 + n], ref); ++bad; }
      }
    }
    printf("i8 M128 N128 K32 u8xs8 mask=%08x: %ld / %ld mismatches\n", mask, bad, checked);
  }
  // fp16 sanity with the same bytes interpreted as halves is meaningless numerically; only time it
  for (int reps : {64, 512, 4096}) {
    long long c8, c16;
    probe<1><<<1, 128>>>(a_d, b_d, d_d, c_d, reps, 0u); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c8, c_d, 8, cudaMemcpyDeviceToHost));
    probe<0><<<1, 128>>>(a_d, b_d, d_d, c_d, reps, 0u); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&c16, c_d, 8, cudaMemcpyDeviceToHost));
    printf("reps %5d: i8 (K=32) %.1f cycles/MMA, f16 (K=16) %.1f cycles/MMA\n", reps, (double)c8 / reps, (double)c16 / reps);
  }
  printf("probe_i8 done\n");
  return 0;
}
