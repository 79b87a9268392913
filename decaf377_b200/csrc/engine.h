// Engine state shared by the translation units of libdecaf377_b200.so.
//
// One `Engine` per CUDA device (streams, scratch arenas, MSM pipeline state).  A process
// may hold several (d377_init_multi); every C-ABI entry point acts on the engine the
// calling thread has selected (d377_set_device) or, by default, on the one d377_init chose.
// Entry points take the engine's lock and make its device current for the duration of the
// call (EngineScope), so the library can be called from any host thread.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>

#include "../../include/decaf377_b200.h"

namespace d377 {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

constexpr int kMsmStages = 8;  // points, count, scan, scatter, accumulate, stitch, bucket_reduce, tail
constexpr int kMaxGroups = 8;

struct MsmGeomInfo {
  int c = 0, W = 0;
};

// Per-device state of the Pippenger pipeline (msm.cu).
struct MsmState {
  bool ready = false;
  // second stream: the scalar side (digit recoding, histogram, scan, scatter) of window
  // group k+1 runs here while the engine stream accumulates group k
  cudaStream_t sort_stream = nullptr;
  // third stream: the latency-bound tail of an MSM (stitch, bucket reduction, weighted tree,
  // Horner, compress) runs here, so that the head of the NEXT MSM (normalisation, first sort,
  // first accumulation) starts on the engine stream right behind the last accumulation.
  // The tail owns one of two workspace sets; see msm_once.
  cudaStream_t tail_stream = nullptr;
  // fourth stream: point conversion / batch normalisation of the NEXT MSM (inputs known ready)
  cudaStream_t points_stream = nullptr;
  cudaEvent_t ev_points_done[2] = {};
  cudaEvent_t ev_fork = nullptr, ev_sorted[kMaxGroups] = {}, ev_sort0 = nullptr, ev_sort1 = nullptr;
  cudaEvent_t ev_acc0[kMaxGroups] = {}, ev_acc[kMaxGroups] = {};
  cudaEvent_t ev_stage[kMsmStages + 1] = {};
  cudaEvent_t ev_acc_done = nullptr;        // engine stream: last accumulation of the current MSM
  cudaEvent_t ev_tail_done[2] = {};         // tail stream: the tail that used workspace set k
  cudaEvent_t ev_join = nullptr;
  bool tail_used[2] = {false, false};       // set k has been used by some tail (event valid)
  bool tail_pending = false;                // a tail has been enqueued since the last join
  int cur_set = 0;                          // the set the most recent MSM used
  DevBuf tail_ws[2];
  MsmGeomInfo last_geom;
  bool last_affine = false;
  int last_groups = 1;
  size_t last_n = 0;
  bool stage_valid = false;
};

struct Engine {
  bool ready = false;
  int device = -1;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::recursive_mutex mu;
  int msm_window_override = 0;
  int msm_host_chunks_override = 0;
  int tune_acc_run = 0, tune_reduce_seg = 0;  // D377_ACC_RUN / D377_REDUCE_SEG (experiments)
  int tune_groups = 0;                        // D377_MSM_GROUPS: window groups of the sort/accumulate pipeline
  int tune_sort_ctas = 1;                     // D377_MSM_SORT_CTAS: sort CTAs per SM while pipelined
  int tune_normalize = 0;                     // D377_MSM_NORMALIZE: -1 never, 1 always, 0 auto
  int tune_gcd_inv = 1;                       // D377_GCD_INV: 0 = Fermat inversion in the normalisation (A/B)
  int tune_norm_wave = 3;                     // D377_MSM_NORM_WAVE: normalisation CTAs per SM (one resident wave); 0 = by batch size
  int tune_stitch_warp = 1 << 18;             // D377_MSM_STITCH_WARP: stitch levels with at most this many slots use the warp-scan kernel
  int tune_tail_overlap = 1;                  // D377_MSM_TAIL_OVERLAP: 0 = tails on the engine stream (A/B)
  int tune_tail_prio = 1;                     // D377_MSM_TAIL_PRIO: 1 = tail stream at the greatest priority
  int tune_norm_min_per = 8;                  // D377_MSM_NORM_MIN_PER: normalise Element inputs when n / 2^17 >= this
  int tune_points_prefetch = 1;               // D377_MSM_POINTS_PREFETCH: 0 = point conversion on the engine stream (A/B)
  int tune_points_prio = 0;                   // D377_MSM_POINTS_PRIO: 1 = points stream at the greatest priority (default: least)
  int tune_acc_tma = 0;                       // D377_MSM_ACC_TMA: 1 = bucket accumulation with the TMA-staged operand stream (experiment)
  int tune_sort_prefetch = 1;                 // D377_MSM_SORT_PREFETCH: 0 = the sort always forks from the engine stream (A/B)
  // host-API staging
  DevBuf in0, in1, out0, out1;
  // prefix products of k_normalize
  DevBuf scratch;
  // canonical copy of scalars handed over in Montgomery form (D377_SCALARS_MONTGOMERY)
  DevBuf sc_canon;
  // d377_batch_msm: products, their decode status, segment sums; host-call staging
  DevBuf bm_prod, bm_ok, bm_sum, bm_sc, bm_pt, bm_off, bm_out, bm_okout;
  // fixed-base table (niels, affine) and its geometry
  void* fb_table = nullptr;
  void* fb_table_jq = nullptr;   // the same multiples on the Jacobi quartic (encoding output)
  int tune_fb_quartic = 1;       // D377_FB_QUARTIC: 0 = Edwards additions + compress (A/B)
  size_t fb_quartic_min = (size_t)1 << 18;  // D377_FB_QUARTIC_MIN: smallest batch that builds the 1.6 GB quartic table
  size_t fb_batch_hint = 0;      // size of the host-buffer batch whose chunks are being launched
  // scratch of element_sum_on: [0] engine stream, [1] result stream
  DevBuf sum_ws[2];
  // small device result + pinned host mirror (8 KiB each; layout in kernels.cu)
  uint8_t* d_small = nullptr;
  uint8_t* h_small = nullptr;
  // pipelined host-buffer MSM (d377_msm_submit / d377_msm_wait)
  static constexpr int kSlots = 4;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_h2d[kSlots] = {};
  cudaEvent_t ev_done[kSlots] = {};
  DevBuf slot_sc[kSlots], slot_pt[kSlots];
  bool slot_busy[kSlots] = {};
  // host-buffer MSMs are cut into up to kMsmHostChunks sub-MSMs so that the upload of
  // chunk k+1 overlaps the Pippenger of chunk k (one event per chunk and slot)
  static constexpr int kMsmHostChunks = 8;
  cudaEvent_t ev_chunk[kSlots][kMsmHostChunks] = {};
  cudaEvent_t ev_chunk_sc[kSlots][kMsmHostChunks] = {};   // the chunk's scalars alone are up
  // chunk-pipelined host API of the batch kernels: H2D on copy_stream, kernels on
  // `stream`, D2H on out_stream, two staging sets
  cudaStream_t out_stream = nullptr;
  cudaEvent_t pe_in[2] = {nullptr, nullptr}, pe_k[2] = {nullptr, nullptr}, pe_out[2] = {nullptr, nullptr};
  DevBuf st_in[2][3], st_out[2][2];
  // prepared MSM bases handed out by d377_msm_bases_create: device pointer -> number of bases
  std::unordered_map<const void*, size_t> bases;
  // asynchronous MSMs (d377_msm_dev_async): sticky status word, reported by d377_sync
  bool async_status_dirty = false;
  MsmState msm;
  // multi-GPU calls (d377_msm_multi*): one persistent host thread per engine
  std::thread worker;
  std::mutex wmu;
  std::condition_variable wcv;
  std::function<int()> wtask;
  bool whas = false, wdone = false, wquit = false;
  int wrc = 0;
  std::string werr;
  cudaEvent_t ev_partial = nullptr;  // this engine's partial sum has landed on the gathering device
  uint64_t peer_mask = 0;            // bit d: this device may store into device d's memory (peer access enabled)
};

// d_small / h_small layout (kSmallBytes each)
constexpr size_t kSmallBytes = 16384;
constexpr size_t kSmallResult = 0;      // [0,160) result of the synchronous calls
constexpr size_t kSmallTmp = 512;       // [512,640) tmp
constexpr size_t kSmallGather = 1024;   // [1024,2048) partial sums of up to 8 devices (d377_msm_multi*)
constexpr size_t kSmallPartials = 2048; // [2048,4096) MSM chunk partials
constexpr size_t kSmallFlags = 4096;    // status word of the synchronous MSM
constexpr size_t kSmallAsyncFlags = 4100;  // sticky status word of d377_msm_dev_async
constexpr size_t kSmallDebug = 4104;    // on-curve debug predicate failures (D377_DEBUG_ON_CURVE builds)
constexpr size_t kSmallSlots = 4352;    // [4352 + 256 k, ...) slot k
constexpr size_t kSmallGatherRing = 8192;  // [8192,12288) four gather areas of d377_msm_multi_dev_async
// [4352, 5376): the four slots; [8192, 11264): three more gather areas of
// d377_msm_multi_dev_async (multi.cu)

Engine& engine();              // the calling thread's engine (selected or default); never null
Engine* engine_for(int device);  // nullptr if that device has not been initialised
void select_engine(Engine* e);   // thread-local selection (nullptr = default)
Engine* selected_engine();       // the calling thread's explicit selection (may be nullptr)
void multi_shutdown();           // multi.cu
void set_error(const char* fmt, ...);
const char* last_error();
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int ensure(DevBuf& b, size_t bytes);
void count_launch();

// Lock the engine and make its device current; restores the previous device on exit.
struct EngineScope {
  Engine& e;
  int prev = -1;
  explicit EngineScope(Engine& en) : e(en) {
    e.mu.lock();
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != e.device) cudaSetDevice(e.device);
  }
  ~EngineScope() {
    if (prev >= 0 && prev != e.device) cudaSetDevice(prev);
    e.mu.unlock();
  }
  EngineScope(const EngineScope&) = delete;
  EngineScope& operator=(const EngineScope&) = delete;
};

#define D377_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t _e = (expr);                                                   \
    if (_e != cudaSuccess) return ::d377::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

// Entry-point prologue: fails when no engine is ready, otherwise locks it and selects its
// device.  The _NOJOIN form leaves MSM tails running on the tail stream; the plain form
// first orders the engine stream behind them (so that anything enqueued by this call sees
// the results of every earlier call).
#define D377_REQUIRE_READY_NOJOIN()                                            \
  ::d377::Engine& _eng = ::d377::engine();                                     \
  if (!_eng.ready) {                                                           \
    ::d377::set_error("d377_init has not been called (or failed): no GPU path available"); \
    return D377_ERR_NOT_INITIALISED;                                           \
  }                                                                            \
  ::d377::EngineScope _scope(_eng)

#define D377_REQUIRE_READY()                                                   \
  D377_REQUIRE_READY_NOJOIN();                                                 \
  do {                                                                         \
    int _jrc = ::d377::msm_join(_eng);                                         \
    if (_jrc) return _jrc;                                                     \
  } while (0)

#define D377_LAUNCHED() (::d377::count_launch())

inline unsigned grid_for(size_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// kernels.cu
void launch_fr_from_mont(const uint8_t* in, size_t n, uint8_t* out, cudaStream_t st);
// codec.cu
void launch_decompress(const uint8_t* enc, size_t n, uint8_t* out, uint8_t* ok, cudaStream_t st,
                       int out_format = D377_PT_ELEMENT);
void launch_compress(const uint8_t* in, size_t n, uint8_t* enc, cudaStream_t st,
                     int in_format = D377_PT_ELEMENT);
void launch_elligator(bool hash, bool encode, const uint8_t* r1, const uint8_t* r2, size_t width, size_t n,
                      uint8_t* out, cudaStream_t st);
void launch_fq_isqrt(const uint8_t* x, size_t n, uint8_t* out, uint8_t* wsq, cudaStream_t st);
void launch_fq_sqrt_ratio(const uint8_t* num, const uint8_t* den, size_t n, uint8_t* out,
                          uint8_t* wsq, cudaStream_t st);
void launch_on_curve(const uint8_t* el, size_t n, int check_order, uint8_t* ok, cudaStream_t st);
// scalar.cu
void launch_scalar_mul(int point_format, bool encode, const uint8_t* points, const uint8_t* scalars,
                       size_t n, uint8_t* out, uint8_t* ok, cudaStream_t st);
int ensure_fb_table();
int ensure_fb_table_jq();
void launch_fixed_base(bool encode, const void* table, const void* table_jq, const uint8_t* scalars,
                       size_t n, uint8_t* out, cudaStream_t st);

// msm.cu
void launch_seg_sum(const uint8_t* prods, const uint8_t* okp, const uint32_t* offs, size_t nseg,
                    uint8_t* out_el, uint8_t* ok_out, cudaStream_t st);
void launch_normalize(const uint8_t* el, size_t n, uint8_t* scratch, uint8_t* out, cudaStream_t st);
// Order the engine stream behind every MSM tail enqueued so far (no host synchronisation).
int msm_join(Engine& e);
int msm_dev(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
            uint8_t* out_element, uint8_t* out_encoding);
int msm_dev_async(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                  uint8_t* out_element, uint8_t* out_encoding, int flags);
// `inputs_ready`: the inputs depend on nothing but chunk_ready[k] (nothing at all without
// chunk_ready): the scalar side may then start under the previous MSM's accumulation.
// `chunk` = 0: one Pippenger (split only above 2^26 pairs).  Otherwise the input is
// processed as ceil(n / chunk) sub-MSMs whose partial sums are added at the end; the
// engine stream waits for chunk_ready[k] (if given) before it touches chunk k.
// The result is complete on result_stream(): the tail stream, or the engine stream when the
// tail overlap is switched off.
int msm_enqueue(const uint8_t* scalars, const uint8_t* points, int point_format, size_t n,
                uint8_t* out_element, uint8_t* out_encoding, uint32_t* flags, size_t chunk = 0,
                const cudaEvent_t* chunk_ready = nullptr, bool inputs_ready = false,
                const cudaEvent_t* scalars_ready_ev = nullptr);
cudaStream_t result_stream(Engine& e);
int msm_check_flags(uint32_t flags);
int check_bases(Engine& e, const uint8_t* points, size_t n);   // kernels.cu
int msm_bases_prepare(const uint8_t* points, int point_format, size_t n, uint8_t* records);
int msm_stage_info(float* ms, int* c, int* W, uint64_t* n);
int msm_timeline(float* ms, int cap, int* ngroups);
bool msm_last_mixed();
void msm_shutdown(Engine& e);
int element_sum_on(Engine& e, cudaStream_t st, const uint8_t* elements, size_t n, uint8_t* out_element,
                   uint8_t* out_encoding);

}  // namespace d377
