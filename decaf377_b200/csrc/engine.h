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
  int msm_host_chunks_override = 0;
  int tune_acc_run = 0, tune_reduce_seg = 0;  // D377_ACC_RUN / D377_REDUCE_SEG (experiments)
  int tune_groups = 0;                        // D377_MSM_GROUPS: window groups of the sort/accumulate pipeline
  int tune_sort_ctas = 1;                     // D377_MSM_SORT_CTAS: sort CTAs per SM while pipelined
  int tune_normalize = 0;                     // D377_MSM_NORMALIZE: -1 never, 1 always, 0 auto
  int tune_gcd_inv = 1;                       // D377_GCD_INV: 0 = Fermat inversion in the normalisation (A/B)
  int tune_norm_wave = 3;                     // D377_MSM_NORM_WAVE: normalisation CTAs per SM (one resident wave); 0 = by batch size
  int tune_stitch_warp = 1 << 18;             // D377_MSM_STITCH_WARP: stitch levels with at most this many slots use the warp-scan kernel
  // host-API staging
  DevBuf in0, in1, out0, out1;
  // msm workspace
  DevBuf msm_ws;
  // prefix products of k_normalize
  DevBuf scratch;
  // fixed-base table (niels, affine) and its geometry
  void* fb_table = nullptr;
  void* fb_table_jq = nullptr;   // the same multiples on the Jacobi quartic (encoding output)
  int tune_fb_quartic = 1;       // D377_FB_QUARTIC: 0 = Edwards additions + compress (A/B)
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
  // host-buffer MSMs are cut into up to kMsmHostChunks sub-MSMs so that the upload of
  // chunk k+1 overlaps the Pippenger of chunk k (one event per chunk and slot)
  static constexpr int kMsmHostChunks = 8;
  cudaEvent_t ev_chunk[kSlots][kMsmHostChunks] = {};
  // chunk-pipelined host API of the batch kernels: H2D on copy_stream, kernels on
  // `stream`, D2H on out_stream, two staging sets
  cudaStream_t out_stream = nullptr;
  cudaEvent_t pe_in[2] = {nullptr, nullptr}, pe_k[2] = {nullptr, nullptr}, pe_out[2] = {nullptr, nullptr};
  DevBuf st_in[2][3], st_out[2][2];
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

// codec.cu
void launch_decompress(const uint8_t* enc, size_t n, uint8_t* out, uint8_t* ok, cudaStream_t st);
void launch_compress(const uint8_t* in, size_t n, uint8_t* enc, cudaStream_t st);
void launch_elligator(bool hash, bool encode, const uint8_t* r1, const uint8_t* r2, size_t n,
                      uint8_t* out, cudaStream_t st);
void launch_fq_isqrt(const uint8_t* x, size_t n, uint8_t* out, uint8_t* wsq, cudaStream_t st);
void launch_fq_sqrt_ratio(const uint8_t* num, const uint8_t* den, size_t n, uint8_t* out,
                          uint8_t* wsq, cudaStream_t st);
// scalar.cu
void launch_scalar_mul(int point_format, bool encode, const uint8_t* points, const uint8_t* scalars,
                       size_t n, uint8_t* out, uint8_t* ok, cudaStream_t st);
int ensure_fb_table();
int ensure_fb_table_jq();
void launch_fixed_base(bool encode, const void* table, const void* table_jq, const uint8_t* scalars,
                       size_t n, uint8_t* out, cudaStream_t st);

// msm.cu
void launch_normalize(const uint8_t* el, size_t n, uint8_t* scratch, uint8_t* out, cudaStream_t st);
int msm_dev(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
            uint8_t* out_element, uint8_t* out_encoding);
// `chunk` = 0: one Pippenger (split only above 2^26 pairs).  Otherwise the input is
// processed as ceil(n / chunk) sub-MSMs whose partial sums are added at the end; the
// engine stream waits for chunk_ready[k] (if given) before it touches chunk k.
int msm_enqueue(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                uint8_t* out_element, uint8_t* out_encoding, uint32_t* flags, size_t chunk = 0,
                const cudaEvent_t* chunk_ready = nullptr);
int msm_check_flags(uint32_t flags);
int msm_bases_prepare(const uint8_t* points, int point_format, size_t n, uint8_t* records);
int msm_stage_info(float* ms, int* c, int* W, uint64_t* n);
int msm_timeline(float* ms, int cap, int* ngroups);
bool msm_last_mixed();
void msm_shutdown();
int element_sum_dev(const uint8_t* elements, size_t n, uint8_t* out_element,
                    uint8_t* out_encoding);

}  // namespace d377
