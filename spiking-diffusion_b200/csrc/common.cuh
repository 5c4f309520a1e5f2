// Shared helpers for libsd_b200.so (sm_100a only).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/sd_b200.h"

namespace sd {

void set_error(const char* fmt, ...);
int check_device();  // SD_OK or SD_ERR_NO_DEVICE for the CURRENT device (cached per device)
int current_device_index();
int sm_count();
int max_threads_per_sm();
int validate_conv_desc(const sd_conv_desc* d);  // SD_OK or SD_ERR_INVALID (conv_simt.cu)

#define SD_REQUIRE(cond, ...)                 \
  do {                                        \
    if (!(cond)) {                            \
      sd::set_error(__VA_ARGS__);             \
      return SD_ERR_INVALID;                  \
    }                                         \
  } while (0)

#define SD_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      sd::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return SD_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define SD_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t e_ = cudaGetLastError();                                               \
    if (e_ != cudaSuccess) {                                                           \
      sd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
      return SD_ERR_CUDA;                                                              \
    }                                                                                  \
  } while (0)

#define SD_DEVICE_OR_RETURN()            \
  do {                                   \
    int rc_ = sd::check_device();        \
    if (rc_ != SD_OK) return rc_;        \
  } while (0)

// ---- STF geometry (see include/sd_b200.h) ---------------------------------------------------------
constexpr int kTileRows = 128;  // M of one tcgen05 tile

__host__ __device__ inline int64_t stf_guard(int W) {
  int64_t g = (int64_t)W + 1;           // the largest row shift of a 3x3 window
  return (g + 7) / 8 * 8;
}
__host__ __device__ inline int64_t stf_rows(int B, int H, int W) {
  int64_t P = (int64_t)H * W;
  int64_t R = (int64_t)B * P;
  R = (R + kTileRows - 1) / kTileRows * kTileRows;
  return stf_guard(W) * 2 + R;
}
__host__ __device__ inline int c8(int C) { return (C + 7) / 8; }

struct StfGeom {
  int H, W, Wp, P;
  int64_t G, R_alloc;
  __host__ __device__ StfGeom(int B, int H_, int W_)
      : H(H_), W(W_), Wp(W_), P(H_ * W_), G(stf_guard(W_)), R_alloc(stf_rows(B, H_, W_)) {}
  __host__ __device__ int64_t row(int b, int y, int x) const { return G + (int64_t)b * P + y * Wp + x; }
  // half index of (t, c, row) for a tensor with C8 channel chunks
  __host__ __device__ int64_t at(int t, int C8, int c, int64_t r) const {
    return (((int64_t)t * C8 + (c >> 3)) * R_alloc + r) * 8 + (c & 7);
  }
};

struct MemoutCoef { float c[SD_MAX_T]; };  // 0.8^(T-1-t), passed by value (kernel parameter)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (the kernels of one diffusion step form a chain of short launches on one stream).
// pdl_launch_dependents(): the next kernel of the stream may be scheduled now (its blocks run their prologue -- barrier
// set-up, TMEM allocation, weights -- while this grid is still working).  pdl_wait(): blocks until every grid this launch
// depends on has completed and its writes are visible; a kernel launched with the attribute must execute it in every
// thread before that thread touches anything an earlier kernel wrote, and before it writes global memory.  Both are no-ops
// in a launch without the attribute.  SD_PDL=0 disables the attribute (sd_pdl_mode()).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// 0: never, 1 (default): for plans of at most two concurrent chains (sd_conv_desc.concurrent <= 2), 2: always.  Measured
// (profiles/r02_experiments.md): +3 % for the single chain of the as-shipped 16 / 32-image batch, +5.5 % for the two chains
// of a 64-image shard, neutral for two large chains; with 3-5 chains the early blocks only take SMs from the other chains'
// kernels (-3 % at 96 images ... -9 % at cfg2's 256).
inline int sd_pdl_mode() {
  static const int mode = [] { const char* e = getenv("SD_PDL"); return e ? atoi(e) : 1; }();
  return mode;
}
inline bool sd_pdl_enabled(int concurrent) { return sd_pdl_mode() == 2 || (sd_pdl_mode() == 1 && concurrent <= 2); }

}  // namespace sd
