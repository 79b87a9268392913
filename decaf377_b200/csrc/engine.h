// Process-wide engine state shared by the translation units of
// libdecaf377_b200.so: device, stream, scratch arena, error reporting.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <mutex>
#include <string>

#include "../../include/decaf377_b200.h"

namespace d377 {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

struct Engine {
  bool ready = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::recursive_mutex mu;
  std::atomic<uint64_t> launches{0};
  int msm_window_override = 0;
  // host-API staging
  DevBuf in0, in1, out0, out1;
  // msm workspace
  DevBuf msm_ws;
  // fixed-base table (niels, affine) and its geometry
  void* fb_table = nullptr;
  // small device result + pinned host mirror (8 KiB each; layout in kernels.cu)
  uint8_t* d_small = nullptr;
  uint8_t* h_small = nullptr;
  // pipelined host-buffer MSM (d377_msm_submit / d377_msm_wait)
  static constexpr int kSlots = 2;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[kSlots] = {nullptr, nullptr};
  cudaEvent_t ev_done[kSlots] = {nullptr, nullptr};
  DevBuf slot_sc[kSlots], slot_pt[kSlots];
  bool slot_busy[kSlots] = {false, false};
};

Engine& engine();
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int ensure(DevBuf& b, size_t bytes);

#define D377_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::d377::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define D377_REQUIRE_READY()                                                   \
  do {                                                                         \
    if (!::d377::engine().ready) {                                             \
      ::d377::set_error("d377_init has not been called (or failed): no GPU path available"); \
      return D377_ERR_NOT_INITIALISED;                                         \
    }                                                                          \
  } while (0)

#define D377_LAUNCHED() (::d377::engine().launches.fetch_add(1, std::memory_order_relaxed))

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// msm.cu
int msm_dev(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
            uint8_t* out_element, uint8_t* out_encoding);
int msm_enqueue(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                uint8_t* out_element, uint8_t* out_encoding, uint32_t* flags);
int msm_check_flags(uint32_t flags);
int msm_stage_info(float* ms, int* c, int* W, uint64_t* n);
int element_sum_dev(const uint8_t* elements, size_t n, uint8_t* out_element,
                    uint8_t* out_encoding);

}  // namespace d377
