// Shared helpers for libsd_b200.so (sm_100a only).
#pragma once
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

}  // namespace sd
