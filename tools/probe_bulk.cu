// Probe of the global -> shared ingest rate of one SM on sm_100a (GPU box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_bulk tools/probe_bulk.cu && build/probe_bulk
// Question behind it ("what limits the dominant kernel"): conv3x3_tc streams one 12 KB weight block per
// (k block, tap) with a single cp.async.bulk and sees ~16 B/clk per SM.  Is that the bulk-copy engine, the L2, or the
// way the copies are issued?  Modes:
//   0  one thread, cp.async.bulk of `bytes` per stage into a ring of `stages` (what the kernel does)
//   1  two threads in two warps, one ring each (two independent bulk streams)
//   2  one warp, cp.async.cg 16 B per lane + cp.async.mbarrier.arrive.noinc (LDGSTS path)
//   3  four warps, ld.global.v4 -> st.shared.v4
//   4  one thread, `pieces` bulk copies per stage (bytes / pieces each)
// All CTAs read the same L2-resident blob (as the kernel's weight stream does) unless `distinct` is set.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

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
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kMaxStages = 16;

__global__ void __launch_bounds__(128, 1) probe(const uint8_t* __restrict__ src, size_t blob, int mode, int bytes, int stages,
                                                int iters, int pieces, int distinct, long long* __restrict__ cyc,
                                                uint32_t* __restrict__ sink, long long* __restrict__ split) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[2 * kMaxStages];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint8_t* base = src + (distinct ? (size_t)blockIdx.x * blob : 0);
  const int rings = mode == 1 ? 2 : 1;
  if (tid == 0) {
    for (int s = 0; s < 2 * kMaxStages; ++s) mbar_init(smem_u32(&bars[s]), mode == 2 ? 32 : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0 || mode == 1 || mode == 4) {
    if (warp < rings && lane == 0) {
      const uint32_t ring = smem_u32(smem) + warp * stages * bytes;
      const uint8_t* b = base + (size_t)warp * (blob / 2);
      const size_t span = blob / rings;
      size_t off = 0;
      long long tw = 0, te = 0, ti = 0;
      for (int i = 0; i < iters; ++i) {
        const int s = i % stages;
        const uint32_t bar = smem_u32(&bars[warp * kMaxStages + s]);
        const long long c0 = clock64();
        if (i >= stages) mbar_wait(bar, ((i / stages) - 1) & 1);
        const long long c1 = clock64();
        mbar_expect_tx(bar, bytes);
        const long long c2 = clock64();
        const int pc = mode == 4 ? pieces : 1;
        const int pb = bytes / pc;
        for (int q = 0; q < pc; ++q) bulk_g2s(ring + s * bytes + q * pb, b + off + q * pb, pb, bar);
        const long long c3 = clock64();
        tw += c1 - c0; te += c2 - c1; ti += c3 - c2;
        off += bytes;
        if (off + bytes > span) off = 0;
      }
      if (blockIdx.x == 0 && warp == 0) { split[0] = tw; split[1] = te; split[2] = ti; }
      for (int i = iters; i < iters + stages; ++i) {
        const int s = i % stages;
        if (i >= stages) mbar_wait(smem_u32(&bars[warp * kMaxStages + s]), ((i / stages) - 1) & 1);
      }
    }
  } else if (mode == 2) {
    if (warp == 0) {
      const uint32_t ring = smem_u32(smem);
      size_t off = 0;
      for (int i = 0; i < iters; ++i) {
        const int s = i % stages;
        const uint32_t bar = smem_u32(&bars[s]);
        if (i >= stages) mbar_wait(bar, ((i / stages) - 1) & 1);
        for (int o = lane * 16; o < bytes; o += 512)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + s * bytes + o), "l"(base + off + o) : "memory");
        asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
        off += bytes;
        if (off + bytes > blob) off = 0;
      }
      for (int i = iters; i < iters + stages; ++i) {
        const int s = i % stages;
        if (i >= stages) mbar_wait(smem_u32(&bars[s]), ((i / stages) - 1) & 1);
      }
    }
  } else if (mode == 3) {
    size_t off = 0;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      uint4* dst = reinterpret_cast<uint4*>(smem + s * bytes);
      const uint4* g = reinterpret_cast<const uint4*>(base + off);
#pragma unroll 4
      for (int o = tid; o < bytes / 16; o += 128) {
        uint4 v;
        asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(g + o));
        dst[o] = v;
      }
      off += bytes;
      if (off + bytes > blob) off = 0;
    }
    __syncthreads();
    acc = reinterpret_cast<uint4*>(smem)[tid];
    if (acc.x == 0x12345678u) sink[0] = acc.y;
  }
  __syncthreads();
  const long long t1 = clock64();
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
  if (tid == 1 && smem[5] == 0x77 && smem[bytes - 1] == 0x55 && smem[77] == 0x11) sink[1] = 1;
}

// burst: n copies of `bytes` issued back to back (one barrier each), then waited for in order; stamps[i] = completion
// time of copy i, stamps[32 + i] = time its issue returned, both relative to the first issue
__global__ void __launch_bounds__(128, 1) burst(const uint8_t* __restrict__ src, int bytes, int n, int one_barrier,
                                                long long* __restrict__ stamps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[32];
  if (threadIdx.x == 0) {
    for (int s = 0; s < 32; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    if (one_barrier) mbar_expect_tx(smem_u32(&bars[0]), bytes * n);
    for (int i = 0; i < n; ++i) {
      const uint32_t bar = smem_u32(&bars[one_barrier ? 0 : i]);
      if (!one_barrier) mbar_expect_tx(bar, bytes);
      bulk_g2s(smem_u32(smem) + i * bytes, src + (size_t)i * bytes, bytes, bar);
      stamps[32 + i] = clock64() - t0;
    }
    for (int i = 0; i < (one_barrier ? 1 : n); ++i) {
      mbar_wait(smem_u32(&bars[i]), 0);
      stamps[i] = clock64() - t0;
    }
  }
}

int main(int argc, char** argv) {
  int dev_sms = 0;
  CK(cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t blob = 1u << 20;   // 1 MB per stream: L2 resident
  uint8_t* src;
  CK(cudaMalloc(&src, blob * dev_sms));
  CK(cudaMemset(src, 1, blob * dev_sms));
  long long* cyc;
  CK(cudaMalloc(&cyc, sizeof(long long) * dev_sms));
  uint32_t* sink;
  CK(cudaMalloc(&sink, 16));
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  long long* split;
  CK(cudaMalloc(&split, 64 * sizeof(long long)));
  CK(cudaMemset(split, 0, 64 * sizeof(long long)));
  CK(cudaFuncSetAttribute(burst, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int one_barrier : {0, 1})
    for (int bytes : {2048, 12288})
      for (int n : {1, 2, 4, 8, 16}) {
        if ((size_t)bytes * n > 196608) continue;
        long long h[64];
        for (int rep = 0; rep < 3; ++rep) {
          burst<<<1, 128, (size_t)bytes * n>>>(src, bytes, n, one_barrier, split);
          CK(cudaDeviceSynchronize());
        }
        CK(cudaMemcpy(h, split, sizeof(h), cudaMemcpyDeviceToHost));
        printf("burst one_barrier=%d bytes=%d n=%d  issue-return:", one_barrier, bytes, n);
        for (int i = 0; i < n; ++i) printf(" %lld", h[32 + i]);
        printf("  complete:");
        for (int i = 0; i < (one_barrier ? 1 : n); ++i) printf(" %lld", h[i]);
        printf("\n");
      }
  struct Case { int mode, bytes, stages, pieces; };
  std::vector<Case> cases;
  for (int bytes : {2048, 4096, 12288, 24576, 49152})
    for (int stages : {2, 4, 8})
      if (bytes * stages <= 196608) cases.push_back({0, bytes, stages, 1});
  for (int bytes : {6144, 12288}) cases.push_back({1, bytes, 4, 1});
  for (int bytes : {4096, 12288}) for (int stages : {2, 4, 8}) cases.push_back({2, bytes, stages, 1});
  for (int bytes : {12288}) cases.push_back({3, bytes, 4, 1});
  for (int pieces : {2, 4, 8}) cases.push_back({4, 12288, 6, pieces});
  printf("mode bytes stages pieces grid distinct  B/clk/SM(min) B/clk/SM(mean)\n");
  for (const Case& c : cases) {
    for (int grid : {1, dev_sms}) {
      for (int distinct : {0}) {
        if (distinct && grid == 1) continue;
        const int iters = (8 << 20) / c.bytes;
        const int rings = c.mode == 1 ? 2 : 1;
        const size_t smem = (size_t)c.bytes * c.stages * rings;
        for (int rep = 0; rep < 2; ++rep)
          probe<<<grid, 128, smem>>>(src, blob, c.mode, c.bytes, c.stages, iters, c.pieces, distinct, cyc, sink, split);
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(grid);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        const double total = (double)c.bytes * iters * rings;
        double worst = 0, sum = 0;
        for (long long v : h) { worst = std::max(worst, (double)v); sum += (double)v; }
        long long sp[3];
        CK(cudaMemcpy(sp, split, sizeof(sp), cudaMemcpyDeviceToHost));
        printf("%4d %6d %5d %5d %5d %6d   %8.2f %8.2f   per-iter wait %lld expect_tx %lld issue %lld\n", c.mode, c.bytes, c.stages,
               c.pieces, grid, distinct, total / worst, total * grid / sum, sp[0] / iters, sp[1] / iters, sp[2] / iters);
      }
    }
  }
  return 0;
}
