// frs.cu -- context, pipeline driver and C ABI of libfreddie_b200.so (see include/freddie_b200.h).
// The pipeline replaces segment() (freddie_segment.py:738-844) for a whole batch of tints.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/freddie_b200.h"
#include "common.cuh"
#include "kernels_signal.cuh"
#include "kernels_dp.cuh"
#include "kernels_finish.cuh"

static thread_local char g_err[512] = "";

#define FRS_SIDE_STREAMS 6  // 0..3: CTA classes of the DP (2..5) and its solver, high priority; 4..5: warp classes
#ifndef DP_BIG_THREADS
#define DP_BIG_THREADS 1024  // CTA size of the DP kernel for subproblems with more than 32 candidates
#endif
#define FRS_SLOTS 6          // batches in flight per context: copy in | head kernels | tail | copy out, and queueing depth
                             // (a batch takes ~5 ms from submit to results at 2 ms per stage: the host must run ahead)

struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <typename T> T* as() const { return (T*)p; }
};

struct Stage {
  const char* name;
  cudaEvent_t ev0, ev1;
  int launches;
  bool used;
};

// One batch in flight: its inputs, its results and its counters.  The intermediates of the pipeline are
// shared by the slots of a context (runs are ordered on the compute stream), so that the copy of the next
// batch and the read-back of the previous one overlap the kernels of the current one.
struct Slot {
  // inputs (device)
  DBuf b_tint_island_off, b_tint_rep_off, b_tint_read_off, b_island_start, b_island_sample_off, b_island_tint,
      b_rep_iv_off, b_rep_weight, b_rep_fs, b_rep_fe, b_rep_tint, b_read_rep, b_read_strand, b_read_len,
      b_read_iv_off, b_read_seq_off, b_read_tint, b_riv_ts, b_riv_te, b_riv_qs, b_riv_qe, b_riv_cig_off, b_cigar,
      b_seq_a, b_seq_t, b_sig_work, b_tiles, b_cov_tiles, b_dig_tiles, b_tint_order, b_params, b_cut_tab;
  // results (device)
  DBuf b_tint_final_off, b_final_pos, b_tint_digit_off, b_digits, b_read_head, b_read_gap_off, b_gap_rec, b_counters;
  // working set of the run's TAIL (clip fetch, poly-A/T scans, head fields): it runs on its own stream beside
  // the head of the next batch, so it owns its buffers
  DBuf b_clip_n, b_clip_words, b_clip_off, b_clip_a, b_clip_t, b_task_order, b_task_res, b_poly_cls, b_poly_flag, b_bsum_tail;
  DBuf b_seq_edge, b_clip_eoff;  // edge store of the batch (input) and the clips' offsets into it
  DBuf b_cigar16, b_cig_n, b_bsum_in;  // compact encodings as they arrive, scratch of the expanding scan
  DBuf b_in_arena;                     // ONE allocation behind every input buffer above (views into it)
  int edge_words = 0;            // > 0: the edge store is in use for this batch
  int n_copies = 0;              // host-to-device copies of the last upload (after merging)
  bool prepped = false, prep_derive_riv = false, prep_cigar16 = false, prep_cig_n = false, prep_qe = false;
  frs_batch hb;  // sizes of the batch; its pointers are not used after the upload
  int n_sig_work = 0, n_sig_direct = 0, n_tiles = 0, n_cov_tiles = 0, n_dig_tiles = 0;
  i64 est_P = 0, est_dig = 0;  // first guesses of the data-dependent capacities (from the batch's shape)
  bool uploaded = false, enqueued = false, ran = false, busy = false, down_pending = false;
  bool seq_resident = false;
  const u32* zc_a = nullptr;  // lazy sequence mode: the caller's (pinned) planes as the device sees them
  const u32* zc_t = nullptr;
  void* h_tab = nullptr;      // pinned staging of the derived work tables
  size_t h_tab_cap = 0;
  i64* h_cnt = nullptr;       // pinned landing area of the counters
  cudaEvent_t ev_up = nullptr, ev_head = nullptr, ev_ran = nullptr, ev_cnt = nullptr, ev_down = nullptr;
  cudaEvent_t tl[12] = {};  // FRS_HOST_PROFILE: timeline of the slot (copy in, head + 5 marks inside it, tail, copy out)
  // parameters of the run (kept for a repeat after a capacity miss)
  frs_params prm;
  std::vector<double> prm_tables;
  std::vector<double> dev_tables;  // what the slot's device tables hold (tables + tp), empty = nothing yet
  Caps caps_used = {0, 0, 0, 0, 0, 0, 0, 0};  // capacities the enqueued run was launched with
  frs_result_sizes sizes;
  i64 n_cand = 0, n_sub = 0, cov_elems = 0, tab_elems = 0, clip_words = 0;
  i64 st_h2d_upload = 0, st_h2d_run = 0, st_d2h_run = 0, st_poly_tasks = 0, st_poly_long = 0;
};

struct frs_context {
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;                 // compute
  cudaStream_t st_in = nullptr, st_out = nullptr;  // host-to-device / device-to-host copies
  cudaStream_t st_tail = nullptr, st_tail_side = nullptr;  // tail of a run (clip fetch + poly scans), beside the next head
  cudaEvent_t ev_tfork = nullptr, ev_tjoin = nullptr;
  cudaStream_t side[FRS_SIDE_STREAMS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[FRS_SIDE_STREAMS] = {};
  cudaEvent_t ev_cov[4] = {};  // coverage chain on side[0]: fork, block offsets ready, matrix ready, threshold done
  char err[512] = "";
  bool profiling = false;
  Slot slot[FRS_SLOTS];
  int cur = 0;       // slot of the synchronous API / of the last submit
  int last_run = 0;  // slot whose intermediates the taps show
  int reruns = 0;    // runs repeated because a capacity was too small (statistics)
  Caps caps = {0, 0, 0, 0, 0, 0, 0, 0};
  // device buffers (grow-only)
  std::vector<DBuf*> all;
  // intermediates shared by the slots
  DBuf b_yraw, b_y, b_sflag, b_bsum, b_bsum_cov, b_cand_flat, b_cand_island, b_island_cand_off, b_tint_cand_off, b_thr, b_vbuf,
      b_leaf_len, b_leaf_sum, b_tint_pos_off, b_tile_state, b_fixed0, b_fixed1, b_sub_start,
      b_sub_n, b_sub_tint, b_sub_info, b_sub_slabs, b_sub_tab_off, b_bases, b_work, b_split_list, b_cursor,
      b_cov_sz, b_tint_cov_off, b_P, b_tab, b_dpfinal, b_ref_list, b_ref_list2, b_gbuf, b_pstate, b_final_flat,
      b_final_island, b_dig_sz, b_seg_ty, b_seg_tn,
      b_run_cnt, b_run_off, b_runs, b_gap_cnt;
  // options (frs_set_option)
  int opt_slab_words = 64, opt_keep_tables = 0, opt_poly_long_class = POLY_LONG_CLASS, opt_lazy_seq = 1;
  cudaEvent_t ev_base = nullptr;  // FRS_HOST_PROFILE: origin of the timelines
  // timing
  Stage stages[FRS_MAX_STAGES];
  int n_stages = 0, cur_stage = -1, launch_count = 0;
};

static int fail(frs_context* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) snprintf(c->err, sizeof c->err, "%s", buf);
  snprintf(g_err, sizeof g_err, "%s", buf);
  return code;
}

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess)                                                                            \
      return fail(c, FRS_ERR_CUDA, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__,   \
                  __LINE__, #call);                                                                   \
  } while (0)

// every stream of the context is idle (before a buffer that kernels in flight may use is replaced)
static int quiesce(frs_context* c) {
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->st_in));
  CK(cudaStreamSynchronize(c->st_out));
  CK(cudaStreamSynchronize(c->st_tail));
  CK(cudaStreamSynchronize(c->st_tail_side));
  for (int i = 0; i < FRS_SIDE_STREAMS; ++i) CK(cudaStreamSynchronize(c->side[i]));
  return 0;
}

static int ensure(frs_context* c, DBuf& b, size_t bytes) {
  if (bytes < 16) bytes = 16;
  if (b.cap >= bytes) return 0;
  if (b.p) {
    int r = quiesce(c);  // grow-only: rare after the first batches
    if (r) return r;
    CK(cudaFree(b.p));
  }
  b.p = nullptr;
  b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  bool known = false;
  for (DBuf* q : c->all) known |= (q == &b);
  if (!known) c->all.push_back(&b);
  return 0;
}
#define ENS(buf, bytes)                          \
  do {                                           \
    int r_ = ensure(c, c->buf, (size_t)(bytes)); \
    if (r_) return r_;                           \
  } while (0)
#define ENSS(buf, bytes)                       \
  do {                                         \
    int r_ = ensure(c, S.buf, (size_t)(bytes)); \
    if (r_) return r_;                         \
  } while (0)

// Every kernel of a run is launched through here: programmatic stream serialization (see pdl_prologue in common.cuh)
// for a kernel that DIRECTLY follows another kernel of this file on the same stream.  Anything else enqueued in
// between (event record / wait, copy, a plain launch) ends the chain and the next kernel is launched plainly: only
// the kernel-after-kernel edge is relaxed, every event edge keeps its full meaning.
// Policy (measured on B200, config 2): a run ALONE on the GPU gains ~2 % (1.62 -> 1.59 ms; the ~60 launch gaps of the
// stream shrink); with another batch of the context in flight (frs_submit pipeline) those gaps are already filled by the
// other batch's kernels and the resident-but-waiting CTAs of a dependent launch only take SM slots from them
// (end to end 78.6 -> 77.0 M reads/s), so enqueue_run switches it off while any other slot is busy.
// FRS_PDL (development): 0 = never, 2 = always.
static int pdl_mode() {
  static const int m = [] { const char* e = getenv("FRS_PDL"); return e ? atoi(e) : 1; }();
  return m;
}
static thread_local bool g_pdl_alone = true;  // set by enqueue_run: no other batch of the context is in flight
static bool pdl_enabled() { return pdl_mode() == 2 || (pdl_mode() == 1 && g_pdl_alone); }
static thread_local cudaStream_t g_pdl_chain = nullptr;  // stream whose last enqueued operation was a launch_k kernel
#define cudaEventRecord(...) (g_pdl_chain = nullptr, cudaEventRecord(__VA_ARGS__))
#define cudaStreamWaitEvent(...) (g_pdl_chain = nullptr, cudaStreamWaitEvent(__VA_ARGS__))
#define cudaMemcpyAsync(...) (g_pdl_chain = nullptr, cudaMemcpyAsync(__VA_ARGS__))
#define cudaMemsetAsync(...) (g_pdl_chain = nullptr, cudaMemsetAsync(__VA_ARGS__))
#define cudaStreamSynchronize(...) (g_pdl_chain = nullptr, cudaStreamSynchronize(__VA_ARGS__))
#define cudaEventSynchronize(...) (g_pdl_chain = nullptr, cudaEventSynchronize(__VA_ARGS__))
template <typename... P, typename... A>
static inline void launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl_enabled() && g_pdl_chain == st && st != nullptr) ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);  // errors are sticky: the run's cudaGetLastError sees them
  g_pdl_chain = st;
}

static void stage_begin(frs_context* c, const char* name, cudaStream_t on = nullptr) {
  // a stage's two events are recorded on the stream its launches go to
  static thread_local cudaStream_t cur_on = nullptr;
  if (!on) on = c->stream;
  int s = -1;
  for (int i = 0; i < c->n_stages; ++i)
    if (c->stages[i].name == name) s = i;
  if (s < 0 && c->n_stages < FRS_MAX_STAGES) {
    s = c->n_stages++;
    c->stages[s].name = name;
    c->stages[s].launches = 0;
    c->stages[s].used = false;
    cudaEventCreate(&c->stages[s].ev0);
    cudaEventCreate(&c->stages[s].ev1);
  }
  if (c->cur_stage >= 0 && c->profiling) cudaEventRecord(c->stages[c->cur_stage].ev1, cur_on ? cur_on : c->stream);
  c->cur_stage = s;
  cur_on = on;
  if (s >= 0) {
    c->stages[s].used = true;
    if (c->profiling) cudaEventRecord(c->stages[s].ev0, on);
  }
}
static void stage_end(frs_context* c, cudaStream_t on = nullptr) {
  if (c->cur_stage >= 0 && c->profiling) cudaEventRecord(c->stages[c->cur_stage].ev1, on ? on : c->stream);
  c->cur_stage = -1;
}
#define LAUNCHED()                                           \
  do {                                                       \
    c->launch_count++;                                       \
    if (c->cur_stage >= 0) c->stages[c->cur_stage].launches++; \
  } while (0)

static inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }
// grid of a grid-stride kernel whose true extent is only known on the device: enough CTAs for the upper
// bound, at most a few waves of the machine
static inline int gs_grid(i64 upper, int threads, int max_ctas = 148 * 8) {
  i64 g = (upper + threads - 1) / threads;
  if (g < 1) g = 1;
  return (int)(g < max_ctas ? g : max_ctas);
}

// device-wide helpers ------------------------------------------------------------------------
// zero fill by a kernel (every buffer has at least 256 bytes of slack behind `bytes`: rounding up to 16 is safe)
static void dev_zero(frs_context* c, cudaStream_t st, void* p, size_t bytes) {
  if (!bytes) return;
  const size_t n16 = (bytes + 15) / 16;
  const size_t g = (n16 + 255) / 256;
  launch_k(k_zero16, (unsigned)(g < (size_t)c->n_sm * 16 ? g : (size_t)c->n_sm * 16), 256, 0, st, (uint4*)p, n16);
  c->launch_count++;
}
static void dev_copy_word(frs_context* c, cudaStream_t st, i64* dst, const void* src, int bytes) {
  CopyWords w;
  w.n = 1; w.dst[0] = dst; w.src[0] = src; w.bytes[0] = bytes;
  launch_k(k_copy_words, 1, 32, 0, st, w);
  c->launch_count++;
}
template <typename TIn, typename TOut>
static int scan_exclusive_on(frs_context* c, cudaStream_t st, DBuf& scratch, const TIn* in, i64 n, TOut* out,
                             i64* total_out = nullptr /* device: also receives out[n] */) {
  if (n <= SCAN_SMALL_MAX && (const void*)in != (const void*)out) {
    launch_k(k_scan_small<TIn, TOut>, 1, 1024, 0, st, in, (int)n, out, total_out); LAUNCHED();
    return 0;
  }
  int nb = cdiv(n > 0 ? n : 1, SCAN_TILE);
  { int r = ensure(c, scratch, (size_t)(nb + 1) * 8); if (r) return r; }
  i64* bs = scratch.as<i64>();
  launch_k(k_scan_block_sums<TIn>, nb, SCAN_THREADS, 0, st, in, n, bs); LAUNCHED();
  launch_k(k_scan_bsums, 1, 1024, 0, st, bs, nb); LAUNCHED();
  launch_k(k_scan_apply<TIn, TOut>, nb, SCAN_THREADS, 0, st, in, n, bs, out, total_out); LAUNCHED();
  return 0;
}
template <typename TIn, typename TOut>
static int scan_exclusive(frs_context* c, const TIn* in, i64 n, TOut* out, i64* total_out = nullptr) {
  return scan_exclusive_on<TIn, TOut>(c, c->stream, c->b_bsum, in, n, out, total_out);
}
// compaction of byte flags; the count ends up in bsum[nb] and is copied to *count_out (device)
static int compact_flags(frs_context* c, const u8* flags, i64 n, int* idx_out, i64* count_out) {
  int nb = cdiv(n > 0 ? n : 1, FLAG_TILE);
  ENS(b_bsum, (size_t)(nb + 1) * 8);
  i64* bs = c->b_bsum.as<i64>();
  launch_k(k_flag_sums, nb, SCAN_THREADS, 0, c->stream, flags, n, bs); LAUNCHED();
  launch_k(k_scan_bsums, 1, 1024, 0, c->stream, bs, nb); LAUNCHED();
  launch_k(k_flag_compact, nb, SCAN_THREADS, 0, c->stream, flags, n, bs, idx_out, count_out); LAUNCHED();
  return 0;
}
static const char* deverr_text(int code) {
  switch (code) {
    case DEVERR_BREAK_LARGE_POS: return "assert max_c_idx_y_v > 0 (freddie_segment.py:643)";
    case DEVERR_BREAK_LARGE_RANGE: return "break_large_problems window leaves the candidate list (freddie_segment.py:640)";
    case DEVERR_RATIO_RANGE: return "assert 0 <= cov_ratio <= 1 (freddie_segment.py:821)";
    case DEVERR_THREAD_CIGAR: return "CIGAR threading failed (freddie_segment.py:303/326/349)";
    case DEVERR_Q_RANGE: return "assert 0 <= q_ssc_pos <= q_esc_pos <= length (freddie_segment.py:389)";
    case DEVERR_GAP_RANGE: return "assert on unaligned gap coordinates (freddie_segment.py:462/466)";
    case DEVERR_POLY_RANGE: return "assert on poly-A/T coordinates (freddie_segment.py:410/441/450)";
    case DEVERR_BACKTRACE: return "internal: DP backtrace left the table";
    default: return "unknown device assert";
  }
}

// C ABI ----------------------------------------------------------------------------------------
extern "C" {

int frs_abi_version(void) { return FRS_ABI_VERSION; }

int frs_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

const char* frs_last_error(const frs_context* ctx) { return ctx ? ctx->err : g_err; }

int frs_mem_info(int device, long long* free_bytes, long long* total_bytes) {
  frs_context* c = nullptr;
  if (!free_bytes || !total_bytes) return fail(c, FRS_ERR_ARG, "frs_mem_info: NULL argument");
  size_t f = 0, t = 0;
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess || cudaMemGetInfo(&f, &t) != cudaSuccess) {
    cudaGetLastError();
    return fail(c, FRS_ERR_CUDA, "frs_mem_info: device %d not available", device);
  }
  cudaSetDevice(prev);
  *free_bytes = (long long)f;
  *total_bytes = (long long)t;
  return 0;
}

int frs_create(int device, frs_context** out) {
  frs_context* c = nullptr;
  if (!out) return fail(c, FRS_ERR_ARG, "frs_create: out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(c, FRS_ERR_CUDA, "frs_create: no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(c, FRS_ERR_ARG, "frs_create: device %d out of range (%d devices)", device, n);
  c = new frs_context();
  c->device = device;
  bool ok = cudaSetDevice(device) == cudaSuccess &&
            cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, -1) == cudaSuccess &&  // above the tail
            cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithPriority(&c->st_tail, cudaStreamNonBlocking, 0) == cudaSuccess &&
            cudaStreamCreateWithPriority(&c->st_tail_side, cudaStreamNonBlocking, 0) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_tfork, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&c->ev_tjoin, cudaEventDisableTiming) == cudaSuccess;
  for (int k = 0; ok && k < FRS_SLOTS; ++k) {
    Slot& S = c->slot[k];
    memset(&S.hb, 0, sizeof S.hb);
    memset(&S.sizes, 0, sizeof S.sizes);
    memset(&S.prm, 0, sizeof S.prm);
    ok = cudaMallocHost((void**)&S.h_cnt, CNT_SLOTS * 8) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_up, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_ran, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_head, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_cnt, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&S.ev_down, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    int r = fail(nullptr, FRS_ERR_CUDA, "frs_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return r;
  }
  if (getenv("FRS_HOST_PROFILE")) {
    cudaEventCreate(&c->ev_base);
    cudaEventRecord(c->ev_base, c->stream);
    for (int k = 0; k < FRS_SLOTS; ++k)
      for (int e = 0; e < 12; ++e) cudaEventCreate(&c->slot[k].tl[e]);
  }
  cudaDeviceGetAttribute(&c->n_sm, cudaDevAttrMultiProcessorCount, device);
  if (c->n_sm < 1) c->n_sm = 148;
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < FRS_SIDE_STREAMS; ++i) {
    // the chain "large DP classes -> solver of the split subproblems" is the critical path of the stage
    cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, i < 4 ? prio_hi : prio_lo);
    cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  for (auto& e : c->ev_cov) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  {
    const int big = 227 * 1024 - 256;  // the kernels also hold a few bytes of static shared memory
    cudaError_t ea[7] = {
        cudaFuncSetAttribute(k_dp_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPW_WARPS * sizeof(DpWarpSmem<8>))),
        cudaFuncSetAttribute(k_dp_warp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPW_WARPS * sizeof(DpWarpSmem<16>))),
        cudaFuncSetAttribute(k_signal, cudaFuncAttributeMaxDynamicSharedMemorySize, SIG_BINS * 4),
        cudaFuncSetAttribute(k_dp<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, big),
        cudaFuncSetAttribute(k_dp<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, big),
        cudaFuncSetAttribute(k_dp<DP_BIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, big),
        cudaFuncSetAttribute(k_dp_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, big)};
    for (int i = 0; i < 7; ++i)
      if (ea[i] != cudaSuccess) {
        int r = fail(nullptr, FRS_ERR_CUDA, "frs_create: shared-memory attribute %d: %s", i, cudaGetErrorString(ea[i]));
        cudaGetLastError();
        frs_destroy(c);
        return r;
      }
  }
  *out = c;
  return 0;
}

void frs_destroy(frs_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->st_in);
  cudaStreamSynchronize(c->st_out);
  if (c->st_tail) cudaStreamSynchronize(c->st_tail);
  if (c->st_tail_side) cudaStreamSynchronize(c->st_tail_side);
  for (DBuf* b : c->all)
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < c->n_stages; ++i) {
    cudaEventDestroy(c->stages[i].ev0);
    cudaEventDestroy(c->stages[i].ev1);
  }
  for (int k = 0; k < FRS_SLOTS; ++k) {
    Slot& S = c->slot[k];
    if (S.h_cnt) cudaFreeHost(S.h_cnt);
    if (S.h_tab) cudaFreeHost(S.h_tab);
    if (S.ev_up) cudaEventDestroy(S.ev_up);
    if (S.ev_ran) cudaEventDestroy(S.ev_ran);
    if (S.ev_head) cudaEventDestroy(S.ev_head);
    if (S.ev_cnt) cudaEventDestroy(S.ev_cnt);
    if (S.ev_down) cudaEventDestroy(S.ev_down);
  }
  for (int i = 0; i < FRS_SIDE_STREAMS; ++i) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  for (auto& e : c->ev_cov)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->stream);
  cudaStreamDestroy(c->st_in);
  cudaStreamDestroy(c->st_out);
  if (c->st_tail) cudaStreamDestroy(c->st_tail);
  if (c->st_tail_side) cudaStreamDestroy(c->st_tail_side);
  if (c->ev_tfork) cudaEventDestroy(c->ev_tfork);
  if (c->ev_tjoin) cudaEventDestroy(c->ev_tjoin);
  delete c;
}

void* frs_stream(frs_context* c) { return c ? (void*)c->stream : nullptr; }

int frs_set_profiling(frs_context* c, int enabled) {
  if (!c) return FRS_ERR_ARG;
  c->profiling = enabled != 0;
  return 0;
}

int frs_last_launch_count(frs_context* c) { return c ? c->launch_count : 0; }

int frs_get_stats(frs_context* c, long long* out, int n) {
  if (!c || !out) return FRS_ERR_ARG;
  const Slot& S = c->slot[c->last_run];
  const long long v[FRS_N_STATS] = {S.st_h2d_upload, S.st_h2d_run, S.st_d2h_run, S.clip_words,
                                    (long long)S.hb.n_seq_words, S.st_poly_tasks, S.st_poly_long, (long long)c->reruns,
                                    (long long)S.n_copies};
  for (int i = 0; i < n && i < FRS_N_STATS; ++i) out[i] = v[i];
  return FRS_N_STATS;
}

int frs_set_option(frs_context* c, int key, long long value) {
  if (!c) return FRS_ERR_ARG;
  switch (key) {
    case FRS_OPT_SLAB_WORDS:
      if (value < 1 || value > (1 << 20)) return fail(c, FRS_ERR_ARG, "frs_set_option: slab words out of range");
      c->opt_slab_words = (int)value;
      return 0;
    case FRS_OPT_KEEP_DP_TABLES:
      c->opt_keep_tables = value != 0;
      return 0;
    case FRS_OPT_LAZY_SEQ:
      c->opt_lazy_seq = value != 0;
      return 0;
    case FRS_OPT_POLY_LONG_CLASS:
      if (value < 1 || value >= POLY_CLASSES) return fail(c, FRS_ERR_ARG, "frs_set_option: poly class out of range");
      c->opt_poly_long_class = (int)value;
      return 0;
    default:
      return fail(c, FRS_ERR_ARG, "frs_set_option: unknown key %d", key);
  }
}

int frs_get_timings(frs_context* c, const char** names, float* ms, int* launches) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->st_tail);
  int k = 0;
  for (int i = 0; i < c->n_stages; ++i) {
    if (!c->stages[i].used) continue;
    float t = 0.f;
    if (c->profiling) cudaEventElapsedTime(&t, c->stages[i].ev0, c->stages[i].ev1);
    names[k] = c->stages[i].name;
    ms[k] = t;
    launches[k] = c->stages[i].launches;
    ++k;
  }
  return k;
}

}  // extern "C"

// ----------------------------------------------------------------------------------------------
// upload: validation, derived work tables, host-to-device copies (all on the copy-in stream)
// ----------------------------------------------------------------------------------------------
#define H2D(buf, src, bytes)                                                                      \
  do {                                                                                            \
    ENSS(buf, bytes);                                                                             \
    if ((bytes) > 0) CK(cudaMemcpyAsync(S.buf.p, src, (size_t)(bytes), cudaMemcpyHostToDevice, c->st_in)); \
    S.st_h2d_upload += (i64)(bytes);                                                              \
  } while (0)

static int stage_upload(frs_context* c, Slot& S, const frs_batch* b) {
  if (b->n_tints <= 0) return fail(c, FRS_ERR_ARG, "frs_upload: empty batch");
  if (b->n_samples <= 0 || b->n_islands <= 0) return fail(c, FRS_ERR_ARG, "frs_upload: batch without islands");
  const int T = b->n_tints, NI = b->n_islands, NR = b->n_reps, N = b->n_reads;
  // ---- validate the offset tables (the reference asserts the same facts while parsing) ----
  if (b->tint_island_off[0] != 0 || b->tint_island_off[T] != NI || b->tint_rep_off[0] != 0 ||
      b->tint_rep_off[T] != NR || b->tint_read_off[0] != 0 || b->tint_read_off[T] != N ||
      b->island_sample_off[0] != 0 || b->island_sample_off[NI] != b->n_samples || b->rep_iv_off[0] != 0 ||
      b->rep_iv_off[NR] != b->n_rep_ivs || b->read_iv_off[0] != 0 || b->read_iv_off[N] != b->n_read_ivs ||
      (b->riv_cig_off && (b->riv_cig_off[0] != 0 || b->riv_cig_off[b->n_read_ivs] != b->n_cigar_ops)) ||
      b->read_seq_off[0] != 0 || b->read_seq_off[N] != b->n_seq_words)
    return fail(c, FRS_ERR_ARG, "frs_upload: inconsistent offset tables");
  if ((!b->cigar && !b->cigar16) || (!b->riv_cig_off && !b->riv_cig_n) || (!b->riv_qe && !b->qe_from_cigar))
    return fail(c, FRS_ERR_ARG, "frs_upload: cigar / riv_cig_off / riv_qe missing without their compact form");
  for (int t = 0; t < T; ++t)
    // a tint without reads is legal (the reference writes a header-only SEGMENT file for it)
    if (b->tint_island_off[t + 1] <= b->tint_island_off[t] || b->tint_rep_off[t + 1] < b->tint_rep_off[t] ||
        b->tint_read_off[t + 1] < b->tint_read_off[t])
      return fail(c, FRS_ERR_ARG, "frs_upload: tint %d has no islands or negative counts", t);
  for (int i = 0; i < NI; ++i)
    if (b->island_sample_off[i + 1] - b->island_sample_off[i] < 2)
      return fail(c, FRS_ERR_ARG, "AssertionError: island %d is empty (freddie_segment.py:140)", i);
  static const bool prof = getenv("FRS_HOST_PROFILE") != nullptr;
  const auto tp0 = std::chrono::steady_clock::now();
  // the slot's previous results must have left the device before its buffers are reused
  if (S.down_pending) { CK(cudaEventSynchronize(S.ev_down)); S.down_pending = false; }
  S.hb = *b;
  // ---- derived host tables ----
  std::vector<SigWork> sig;
  std::vector<std::pair<int, int>> direct_runs;  // rep ranges of the sparse tints (k_signal direct mode)
  std::vector<TileWork> tiles;
  std::vector<RepTile> cov_tiles, dig_tiles;
  const int DIG_REPS = DIG_THREADS;
  i64 est_P = 0, est_dig = 0;
  tiles.reserve((size_t)b->n_samples / TILE_SAMPLES + (size_t)NI + 16);
  cov_tiles.reserve((size_t)NR / COV_THREADS + (size_t)T + 16);
  dig_tiles.reserve((size_t)NR / DIG_REPS + (size_t)T + 16);
  // every read points at a rep of its own tint with the same number of intervals (the dedupe key of
  // read_split, freddie_segment.py:165-170); checked before anything is copied
  {
    long long bad_read = -1;
    int bad_kind = 0;
#pragma omp parallel for schedule(static) if (N > 65536)
    for (int t = 0; t < T; ++t) {
      const int r0 = b->tint_rep_off[t], r1 = b->tint_rep_off[t + 1];
      for (int r = b->tint_read_off[t]; r < b->tint_read_off[t + 1]; ++r) {
        const int rep = b->read_rep[r];
        int kind = 0;
        if (rep < r0 || rep >= r1) kind = 1;
        else if (b->read_iv_off[r + 1] - b->read_iv_off[r] != b->rep_iv_off[rep + 1] - b->rep_iv_off[rep]) kind = 2;
        if (kind) {
#pragma omp critical
          if (bad_read < 0 || r < bad_read) { bad_read = r; bad_kind = kind; }
        }
      }
    }
    if (bad_kind == 1)
      return fail(c, FRS_ERR_ARG, "frs_upload: read %lld points at rep %d of another tint", bad_read, b->read_rep[bad_read]);
    if (bad_kind == 2)
      return fail(c, FRS_ERR_ARG, "frs_upload: read %lld and its rep %d differ in their number of intervals", bad_read,
                  b->read_rep[bad_read]);
  }
  for (int t = 0; t < T; ++t) {
    for (int i = b->tint_island_off[t]; i < b->tint_island_off[t + 1]; ++i) {
      int n = b->island_sample_off[i + 1] - b->island_sample_off[i];
      for (int lo = 0; lo < n; lo += TILE_SAMPLES) tiles.push_back(TileWork{i, lo, b->island_sample_off[i], n});
    }
    int r0 = b->tint_rep_off[t], r1 = b->tint_rep_off[t + 1];
    int s0 = b->island_sample_off[b->tint_island_off[t]], s1 = b->island_sample_off[b->tint_island_off[t + 1]];
    int single = (r1 - r0) <= SIG_REPS;
    const i64 n_endpoints = 2 * (i64)(b->rep_iv_off[r1] - b->rep_iv_off[r0]);
    if (n_endpoints < 8 * (i64)(s1 - s0)) {  // sparse tint: endpoints go straight to the global signal
      // flat samples and reps need no tint: runs of consecutive sparse tints share full CTAs
      if (r1 > r0) {
        if (!direct_runs.empty() && direct_runs.back().second == r0) direct_runs.back().second = r1;
        else direct_runs.push_back(std::make_pair(r0, r1));
      }
    } else {
      for (int w = s0; w < s1; w += SIG_BINS)
        for (int r = r0; r < r1; r += SIG_REPS)
          sig.push_back(SigWork{t, w, w + SIG_BINS < s1 ? w + SIG_BINS : s1, r, r + SIG_REPS < r1 ? r + SIG_REPS : r1, single});
    }
    int R = r1 - r0, Rp = (R + 3) & ~3;
    for (int r = 0; r < Rp; r += COV_THREADS) cov_tiles.push_back(RepTile{t, r});
    for (int r = 0; r < R; r += DIG_REPS) dig_tiles.push_back(RepTile{t, r});
    const int ni = b->tint_island_off[t + 1] - b->tint_island_off[t];
    est_P += (i64)Rp * ((s1 - s0) / 16 + 2 * ni + 8);
    est_dig += (i64)R * ((s1 - s0) / 24 + 2 * ni + 8);
  }
  S.est_P = est_P;
  S.est_dig = est_dig;
  // tints by decreasing sample count (per-tint CTAs: start the long ones first)
  std::vector<int> tint_order(T);
  for (int t = 0; t < T; ++t) tint_order[t] = t;
  {
    const int* io = b->tint_island_off;
    const int* so = b->island_sample_off;
    std::stable_sort(tint_order.begin(), tint_order.end(), [&](int x, int y) {
      return so[io[x + 1]] - so[io[x]] > so[io[y + 1]] - so[io[y]];
    });
  }
  S.n_sig_work = (int)sig.size();  // histogram items first, then the direct ones
  for (const auto& run : direct_runs)
    for (int r = run.first; r < run.second; r += SIG_DIRECT_REPS)
      sig.push_back(SigWork{-1, 0, b->n_samples, r, r + SIG_DIRECT_REPS < run.second ? r + SIG_DIRECT_REPS : run.second, 2});
  S.n_sig_direct = (int)sig.size() - S.n_sig_work;
  S.n_tiles = (int)tiles.size();
  S.n_cov_tiles = (int)cov_tiles.size();
  S.n_dig_tiles = (int)dig_tiles.size();
  const auto tp1 = std::chrono::steady_clock::now();
  // ---- copies ----
  // Every input of the slot lives in ONE device arena, in the order below (256-byte aligned): arrays that are
  // laid out the same way on the host (frs_batch.host_arena: one pinned allocation, same order and alignment --
  // what freddie_b200.pack.PackedBatch.pin builds) cross the bus as a few large copies instead of ~35 small ones
  // (measured: 2.3 ms for 87 MB in 37 copies against 1.6 ms in one).
  if (S.tl[0]) cudaEventRecord(S.tl[0], c->st_in);
  S.st_h2d_upload = 0;
  S.n_copies = 0;
  struct InPlan { DBuf* buf; const void* src; size_t bytes; };
  std::vector<InPlan> plan;
  plan.reserve(48);
  auto IN = [&](DBuf& buf, const void* src, size_t bytes) { plan.push_back(InPlan{&buf, src, bytes}); };
  const bool derive_riv = !b->riv_ts || !b->riv_te;  // NULL: derived on the device from the rep intervals
  IN(S.b_tint_island_off, b->tint_island_off, (size_t)(T + 1) * 4);
  IN(S.b_tint_rep_off, b->tint_rep_off, (size_t)(T + 1) * 4);
  IN(S.b_tint_read_off, b->tint_read_off, (size_t)(T + 1) * 4);
  IN(S.b_island_start, b->island_start, (size_t)NI * 4);
  IN(S.b_island_sample_off, b->island_sample_off, (size_t)(NI + 1) * 4);
  IN(S.b_rep_iv_off, b->rep_iv_off, (size_t)(NR + 1) * 4);
  IN(S.b_rep_weight, b->rep_weight, (size_t)NR * 4);
  IN(S.b_rep_fs, b->rep_iv_fs, (size_t)b->n_rep_ivs * 4);
  IN(S.b_rep_fe, b->rep_iv_fe, (size_t)b->n_rep_ivs * 4);
  IN(S.b_read_rep, b->read_rep, (size_t)N * 4);
  IN(S.b_read_strand, b->read_strand, (size_t)N);
  IN(S.b_read_len, b->read_len, (size_t)N * 4);
  IN(S.b_read_iv_off, b->read_iv_off, (size_t)(N + 1) * 4);
  IN(S.b_read_seq_off, b->read_seq_off, (size_t)(N + 1) * 8);
  if (!derive_riv) {
    IN(S.b_riv_ts, b->riv_ts, (size_t)b->n_read_ivs * 4);
    IN(S.b_riv_te, b->riv_te, (size_t)b->n_read_ivs * 4);
  }
  IN(S.b_riv_qs, b->riv_qs, (size_t)b->n_read_ivs * 4);
  // CIGAR ops, their offsets and the query ends: whole, or in their compact forms (expanded by kernels below)
  if (b->cigar16) IN(S.b_cigar16, b->cigar16, (size_t)b->n_cigar_ops * 2);
  else IN(S.b_cigar, b->cigar, (size_t)b->n_cigar_ops * 4);
  if (b->riv_cig_n) IN(S.b_cig_n, b->riv_cig_n, (size_t)b->n_read_ivs);
  else IN(S.b_riv_cig_off, b->riv_cig_off, (size_t)(b->n_read_ivs + 1) * 4);
  if (b->riv_qe) IN(S.b_riv_qe, b->riv_qe, (size_t)b->n_read_ivs * 4);
  // sequence bit-planes.  The poly-A/T scans only look at the soft clips of a read, and those are known
  // after segmentation.  Lazy mode (default): when the caller's planes are pinned / registered host memory
  // the device can address, nothing of them is copied here but the optional edge store (the first / last plane
  // words of every read, dense), and k_clip_gather fetches the few longer clips during the run.  Pageable
  // planes are copied whole (there is no host-side gather and no round trip inside a run).
  S.zc_a = S.zc_t = nullptr;
  S.seq_resident = true;
  S.edge_words = 0;
  if (c->opt_lazy_seq && b->n_seq_words > 0) {
    cudaPointerAttributes aa, at;
    const bool ok_a = cudaPointerGetAttributes(&aa, b->seq_is_a) == cudaSuccess && aa.type == cudaMemoryTypeHost && aa.devicePointer;
    const bool ok_t = cudaPointerGetAttributes(&at, b->seq_is_t) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer;
    cudaGetLastError();  // an unregistered pointer is not an error here
    if (ok_a && ok_t) {
      S.zc_a = (const u32*)aa.devicePointer;
      S.zc_t = (const u32*)at.devicePointer;
      S.seq_resident = false;
    }
  }
  if (S.seq_resident) {
    IN(S.b_seq_a, b->seq_is_a, (size_t)b->n_seq_words * 4);
    IN(S.b_seq_t, b->seq_is_t, (size_t)b->n_seq_words * 4);
  } else if (b->seq_edge && b->seq_edge_words > 0 && b->seq_edge_words <= 64) {
    S.edge_words = b->seq_edge_words;
    IN(S.b_seq_edge, b->seq_edge, (size_t)N * 4 * (size_t)S.edge_words * 4);
  }
  // the derived tables go through pinned staging owned by the slot, so that their copy is asynchronous
  {
    const size_t bytes[5] = {(size_t)T * 4, sig.size() * sizeof(SigWork), tiles.size() * sizeof(TileWork),
                             cov_tiles.size() * sizeof(RepTile), dig_tiles.size() * sizeof(RepTile)};
    const void* src[5] = {tint_order.data(), sig.data(), tiles.data(), cov_tiles.data(), dig_tiles.data()};
    size_t off[6] = {0};
    for (int k = 0; k < 5; ++k) off[k + 1] = off[k] + ((bytes[k] + 255) & ~(size_t)255);
    if (S.h_tab_cap < off[5]) {
      // the previous copies out of the staging area are done (ev_up of the slot's last upload)
      if (S.h_tab) { CK(cudaStreamSynchronize(c->st_in)); CK(cudaFreeHost(S.h_tab)); }
      S.h_tab = nullptr;
      S.h_tab_cap = 0;
      const size_t want = off[5] + off[5] / 4 + (1u << 20);
      CK(cudaMallocHost(&S.h_tab, want));
      S.h_tab_cap = want;
    } else if (S.uploaded) {
      CK(cudaEventSynchronize(S.ev_up));  // the staging area is free again
    }
    for (int k = 0; k < 5; ++k)
      if (bytes[k]) memcpy((char*)S.h_tab + off[k], src[k], bytes[k]);
    IN(S.b_tint_order, (char*)S.h_tab + off[0], bytes[0]);
    IN(S.b_sig_work, (char*)S.h_tab + off[1], bytes[1]);
    IN(S.b_tiles, (char*)S.h_tab + off[2], bytes[2]);
    IN(S.b_cov_tiles, (char*)S.h_tab + off[3], bytes[3]);
    IN(S.b_dig_tiles, (char*)S.h_tab + off[4], bytes[4]);
  }
  const size_t n_copied = plan.size();
  // device-only inputs (filled by the kernels below), behind the copied ones
  if (derive_riv) {
    IN(S.b_riv_ts, nullptr, (size_t)b->n_read_ivs * 4);
    IN(S.b_riv_te, nullptr, (size_t)b->n_read_ivs * 4);
  }
  if (b->cigar16) IN(S.b_cigar, nullptr, (size_t)b->n_cigar_ops * 4);
  if (b->riv_cig_n) IN(S.b_riv_cig_off, nullptr, (size_t)(b->n_read_ivs + 1) * 4);
  if (!b->riv_qe) IN(S.b_riv_qe, nullptr, (size_t)b->n_read_ivs * 4);
  IN(S.b_island_tint, nullptr, (size_t)NI * 4);
  IN(S.b_rep_tint, nullptr, (size_t)NR * 4);
  IN(S.b_read_tint, nullptr, (size_t)N * 4);
  {
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t total = 0;
    for (const InPlan& p : plan) total += al(p.bytes < 16 ? 16 : p.bytes);
    { int r = ensure(c, S.b_in_arena, total + 256); if (r) return r; }
    char* base = (char*)S.b_in_arena.p;
    size_t o = 0;
    for (const InPlan& p : plan) {
      p.buf->p = base + o;
      p.buf->cap = al(p.bytes < 16 ? 16 : p.bytes);
      o += p.buf->cap;
    }
    // copies: runs of arrays that are adjacent on the host exactly as they are in the arena become one copy
    const bool arena = b->host_arena != 0;
    auto staged = [&](const void* q) { return (const char*)q >= (const char*)S.h_tab && (const char*)q < (const char*)S.h_tab + S.h_tab_cap; };
    size_t k = 0;
    while (k < n_copied) {
      size_t j = k, bytes = plan[k].bytes;
      // extend the run while the next array starts exactly where this one's 256-byte slot ends, inside the
      // same host allocation (the slot's staging area, or the caller's declared arena)
      while (j + 1 < n_copied && plan[j + 1].bytes > 0 &&
             (const char*)plan[j + 1].src == (const char*)plan[j].src + plan[j].buf->cap &&
             ((staged(plan[j].src) && staged(plan[j + 1].src)) || (arena && !staged(plan[j].src) && !staged(plan[j + 1].src)))) {
        ++j;
        bytes = (size_t)(((char*)plan[j].buf->p + plan[j].bytes) - (char*)plan[k].buf->p);
      }
      if (bytes > 0) CK(cudaMemcpyAsync(plan[k].buf->p, plan[k].src, bytes, cudaMemcpyHostToDevice, c->st_in));
      for (size_t q = k; q <= j; ++q) S.st_h2d_upload += (i64)plan[q].bytes;
      S.n_copies++;
      k = j + 1;
    }
  }
  // the kernels that complete the inputs (expansion of the compact forms, owner tables, read intervals) do NOT
  // run on the copy stream: behind the high-priority kernels of the batches in flight they would wait for SM
  // slots and hold up the next batch's copies (measured: 0.7 ms gaps between transfers); enqueue_prep runs
  // them at the start of the slot's first run instead
  S.prep_derive_riv = derive_riv;
  S.prep_cigar16 = b->cigar16 != nullptr;
  S.prep_cig_n = b->riv_cig_n != nullptr;
  S.prep_qe = b->riv_qe == nullptr;
  S.prepped = false;
  CK(cudaEventRecord(S.ev_up, c->st_in));
  if (S.tl[1]) cudaEventRecord(S.tl[1], c->st_in);
  if (prof) {
    const auto tp2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[frs host profile] upload: checks + tables %.3f ms, copies enqueued %.3f ms\n",
            std::chrono::duration<double, std::milli>(tp1 - tp0).count(), std::chrono::duration<double, std::milli>(tp2 - tp1).count());
  }
  S.uploaded = true;
  S.enqueued = false;
  S.ran = false;
  return 0;
}

// ----------------------------------------------------------------------------------------------
// run: every kernel of the pipeline, enqueued without a single host round trip
// ----------------------------------------------------------------------------------------------
// completes the inputs of a freshly uploaded batch on stream `st` (which has waited for the copies)
static int enqueue_prep(frs_context* c, Slot& S, cudaStream_t st) {
  const frs_batch& B = S.hb;
  const int T = B.n_tints, NI = B.n_islands, NR = B.n_reps, N = B.n_reads;
  if (S.prep_cigar16 && B.n_cigar_ops > 0)
    launch_k(k_widen_u16, gs_grid(B.n_cigar_ops, 256), 256, 0, st, S.b_cigar16.as<unsigned short>(), B.n_cigar_ops, S.b_cigar.as<u32>());
  if (S.prep_cig_n) {
    const int saved = c->launch_count;
    int r = scan_exclusive_on<u8, int>(c, st, S.b_bsum_in, S.b_cig_n.as<u8>(), (i64)B.n_read_ivs, S.b_riv_cig_off.as<int>());
    c->launch_count = saved;
    if (r) return r;
  }
  if (S.prep_qe && B.n_read_ivs > 0)
    launch_k(k_derive_qe, gs_grid(B.n_read_ivs, 256), 256, 0, st, B.n_read_ivs, S.b_riv_qs.as<int>(), S.b_riv_cig_off.as<int>(),
                                                          S.b_cigar.as<u32>(), S.b_riv_qe.as<int>());
  // owner tables (tint of every island / rep / read) are derived on the device
  launch_k(k_owner_tables, cdiv((i64)NI + NR + N, 256), 256, 0, st, 
      T, NI, NR, N, S.b_tint_island_off.as<int>(), S.b_tint_rep_off.as<int>(), S.b_tint_read_off.as<int>(),
      S.b_island_tint.as<int>(), S.b_rep_tint.as<int>(), S.b_read_tint.as<int>());
  if (S.prep_derive_riv && N > 0)
    launch_k(k_derive_riv, cdiv((i64)N * 8, 256), 256, 0, st, N, NI, S.b_read_rep.as<int>(), S.b_read_iv_off.as<int>(),
                                               S.b_rep_iv_off.as<int>(), S.b_rep_fs.as<int>(), S.b_rep_fe.as<int>(),
                                               S.b_island_sample_off.as<int>(), S.b_island_start.as<int>(),
                                               S.b_riv_ts.as<int>(), S.b_riv_te.as<int>());
  CK(cudaGetLastError());
  S.prepped = true;
  return 0;
}

// largest tint (in words of 32 read reps) whose small subproblems are solved by one warp each (k_dp_warp);
// larger tints go to the CTA classes, where the warps of a CTA share the words of a chunk.  FRS_DP_WARP_WORDS:
// development knob.
static int dp_warp_words() {
  static int v = -1;
  if (v < 0) {
    v = DP_WARP_MAX_WORDS;
    if (const char* e = getenv("FRS_DP_WARP_WORDS")) v = std::max(1, std::min(2047, atoi(e)));
  }
  return v;
}

static int check_params(frs_context* c, const frs_params* prm) {
  // parse_args asserts (freddie_segment.py:104-109)
  if (!(prm->tp >= 0.5 && prm->tp <= 1.0)) return fail(c, FRS_ERR_ARG, "AssertionError: 1 >= threshold_rate >= 0.5");
  if (!(prm->vf > 0 && prm->vf < 10)) return fail(c, FRS_ERR_ARG, "AssertionError: 10 > variance_factor > 0");
  if (!(prm->sigma > 0 && prm->sigma <= 50)) return fail(c, FRS_ERR_ARG, "AssertionError: 50 >= sigma > 0");
  if (!(prm->mps > 3)) return fail(c, FRS_ERR_ARG, "AssertionError: max_problem_size > 3");
  if (prm->mps < 9)
    return fail(c, FRS_ERR_LIMIT, "max_problem_size < 9 is not supported: the reference's +-5 anchor window "
                                  "(freddie_segment.py:639) indexes outside the problem there (IndexError / wrap)");
  if (!(prm->lo >= 0)) return fail(c, FRS_ERR_ARG, "AssertionError: min_read_support_outside >= 0");
  if (prm->gauss_radius != (int)(4.0 * prm->sigma + 0.5) || prm->refine_radius != (int)(1.0 * prm->sigma + 0.5))
    return fail(c, FRS_ERR_ARG, "frs_run: kernel radii do not match sigma");
  if (prm->gauss_radius > 1000) return fail(c, FRS_ERR_LIMIT, "gauss radius too large");
  if (prm->thr_table_len < 0 || prm->thr_table_len > 4096) return fail(c, FRS_ERR_ARG, "frs_run: threshold table length");
  return 0;
}

static void keep_params(Slot& S, const frs_params* prm) {
  const int lw = prm->gauss_radius, rr = prm->refine_radius;
  const size_t n = (size_t)prm->thr_table_len + (2 * lw + 1) + (2 * rr + 1);
  S.prm_tables.resize(n);
  double* d = S.prm_tables.data();
  if (prm->thr_table_len) memcpy(d, prm->thr_table, (size_t)prm->thr_table_len * 8);
  memcpy(d + prm->thr_table_len, prm->gauss_w, (size_t)(2 * lw + 1) * 8);
  memcpy(d + prm->thr_table_len + (2 * lw + 1), prm->refine_w, (size_t)(2 * rr + 1) * 8);
  S.prm = *prm;
  S.prm.thr_table = d;
  S.prm.gauss_w = d + prm->thr_table_len;
  S.prm.refine_w = d + prm->thr_table_len + (2 * lw + 1);
}

static int enqueue_run(frs_context* c, Slot& S) {
  {
    bool alone = true;
    for (int k = 0; k < FRS_SLOTS; ++k) alone &= (&c->slot[k] == &S) || !c->slot[k].busy;
    g_pdl_alone = alone;
  }
  const frs_params* prm = &S.prm;
  const frs_batch& B = S.hb;
  const int T = B.n_tints, NI = B.n_islands, NR = B.n_reps, N = B.n_reads;
  const i64 L = B.n_samples;
  cudaStream_t st = c->stream;
  c->launch_count = 0;
  for (int i = 0; i < c->n_stages; ++i) { c->stages[i].used = false; c->stages[i].launches = 0; }

  // upper bounds known from the batch alone
  const i64 KMAX = L / 2 + 2 * (i64)NI + 16;   // peaks are >= 2 apart, plus both ends of every island
  const i64 NSUB_MAX = KMAX / 2 + 1;           // every subproblem has an interior candidate of its own
  const i64 NFIN_MAX = KMAX + L / 20 + 16;     // refine adds peaks >= 20 apart inside segments longer than 40
  // capacities of the data-dependent buffers: grow-only, first guesses from the shape of the batch
  Caps& cp = c->caps;
  cp.P = std::max<i64>(cp.P, S.est_P);
  cp.dig = std::max<i64>(cp.dig, S.est_dig);
  cp.work = std::max<i64>(cp.work, std::min<i64>(NSUB_MAX, 1 << 16));
  cp.split = std::max<i64>(cp.split, 1024);
  cp.tab = std::max<i64>(cp.tab, 1 << 16);
  cp.runs = std::max<i64>(cp.runs, (i64)NR * 4 + 64);
  cp.gaps = std::max<i64>(cp.gaps, (i64)N * 3 + 64);
  cp.clipw = std::max<i64>(cp.clipw, S.seq_resident ? 0 : (i64)N * 10 + 64);
  S.caps_used = cp;  // by value in every launch below: a later growth (other slot) does not change this run

  // every buffer of the run, before the first launch (a growing buffer drains the context)
  const auto tr0 = std::chrono::steady_clock::now();
  const int lw = prm->gauss_radius, rr = prm->refine_radius;
  const size_t ptab = (size_t)prm->thr_table_len + (2 * lw + 1) + (2 * rr + 1);
  ENSS(b_params, ptab * 8);
  ENSS(b_cut_tab, CUT_TAB_N * sizeof(int2));
  ENSS(b_counters, CNT_SLOTS * 8);
  ENS(b_yraw, L * 4);
  ENS(b_y, L * 8);
  ENS(b_sflag, L);
  ENS(b_cand_flat, KMAX * 4);
  ENS(b_vbuf, L * 8);
  ENS(b_tint_pos_off, (size_t)(T + 1) * 4);
  const size_t n_groups = (size_t)S.n_tiles / TILE_GROUP + 1;
  ENS(b_tile_state, n_groups * 8 + (size_t)S.n_tiles * (8 + 2 * TILE_WORDS * 4 + 4) + 64);
  ENS(b_thr, (size_t)T * 8);
  {
    size_t nh = (size_t)L / 8 + 64 * (size_t)T + 64;  // heap scratch of the giant tints (see k_threshold)
    ENS(b_leaf_len, nh * 4);
    ENS(b_leaf_sum, nh * 8);
  }
  ENS(b_cand_island, KMAX * 4);
  ENS(b_island_cand_off, (size_t)(NI + 1) * 4);
  ENS(b_fixed0, KMAX);
  ENS(b_fixed1, KMAX);
  ENS(b_sub_start, NSUB_MAX * 4);
  ENS(b_sub_n, NSUB_MAX * 4);
  ENS(b_sub_tint, NSUB_MAX * 4);
  ENS(b_sub_info, NSUB_MAX * 4);
  ENS(b_sub_slabs, NSUB_MAX * 4);
  ENS(b_sub_tab_off, NSUB_MAX * 8);
  ENS(b_tint_cand_off, (size_t)(T + 1) * 4);
  ENS(b_cov_sz, (size_t)(T + 1) * 8);
  ENS(b_tint_cov_off, (size_t)(T + 1) * 8);
  ENS(b_bases, (16 + DP_CLASSES * DP_BUCKETS) * 4);
  ENS(b_cursor, (16 + DP_CLASSES * DP_BUCKETS) * 4);
  ENS(b_P, cp.P * 4);
  ENS(b_dpfinal, KMAX);
  ENS(b_work, cp.work * 8);
  ENS(b_split_list, cp.split * 4);
  ENS(b_tab, cp.tab * 4);
  ENS(b_ref_list, KMAX * 8 + 16);
  ENS(b_ref_list2, KMAX * 8 + 16);
  ENS(b_gbuf, L * 8);
  ENS(b_pstate, L);
  ENS(b_final_flat, KMAX * 4);
  ENSS(b_final_pos, NFIN_MAX * 4);
  ENS(b_final_island, NFIN_MAX * 4);
  ENSS(b_tint_final_off, (size_t)(T + 1) * 4);
  ENS(b_dig_sz, (size_t)(T + 1) * 8);
  ENSS(b_tint_digit_off, (size_t)(T + 1) * 8);
  ENS(b_seg_ty, NFIN_MAX * 4);
  ENS(b_seg_tn, NFIN_MAX * 4);
  ENSS(b_digits, cp.dig);
  ENS(b_run_cnt, (size_t)NR * 4);
  ENS(b_run_off, (size_t)(NR + 1) * 4);
  ENS(b_gap_cnt, (size_t)N * 4);
  ENSS(b_read_gap_off, (size_t)(N + 1) * 4);
  ENS(b_runs, cp.runs * 8);
  ENSS(b_read_head, (size_t)N * 32);
  ENSS(b_gap_rec, cp.gaps * 12);
  ENSS(b_clip_n, (size_t)N * 8);
  ENSS(b_clip_words, (size_t)N * 8);
  ENSS(b_clip_off, (size_t)N * 16 + 8);
  ENSS(b_task_order, (size_t)N * 16);
  ENSS(b_task_res, (size_t)N * 4 * sizeof(PolyRes));
  ENSS(b_poly_cls, (2 * POLY_CLASSES + 1) * 4);
  ENSS(b_poly_flag, (size_t)N * 4);
  ENSS(b_bsum_tail, ((size_t)N * 2 / SCAN_TILE + 2) * 8);
  if (S.edge_words) ENSS(b_clip_eoff, (size_t)N * 8);
  if (!S.seq_resident) {
    ENSS(b_clip_a, cp.clipw * 4);
    ENSS(b_clip_t, cp.clipw * 4);
  }
  // subproblems are at most max_problem_size + 11 candidates long (k_fixed_b: anchors move by < 5 + 5);
  // the solver keeps G / arg of the largest one in shared memory and stages the tables of those that fit
  const int dp_max_n = std::min(prm->mps + 12, 180);
  int dp_stage_n = std::min(dp_max_n, DP_SMEM_MAX_N);
  while (dp_stage_n > 0 && dps_smem_bytes(dp_max_n, dp_stage_n) > 216 * 1024) dp_stage_n -= 4;
  if (dp_stage_n < 0) dp_stage_n = 0;

  static const bool prof_run = getenv("FRS_HOST_PROFILE") != nullptr;
  const auto tr1 = std::chrono::steady_clock::now();
  // the run starts when the batch has arrived and the slot's previous results have been read back
  CK(cudaStreamWaitEvent(st, S.ev_up, 0));
  if (S.down_pending) CK(cudaStreamWaitEvent(st, S.ev_down, 0));

  if (S.tl[2]) cudaEventRecord(S.tl[2], st);
  if (!S.prepped) { int r = enqueue_prep(c, S, st); if (r) return r; }
  // parameter tables
  double* d_tbl = S.b_params.as<double>();
  double* d_gw = d_tbl + prm->thr_table_len;
  double* d_rw = d_gw + (2 * lw + 1);
  const int2* d_cut = S.b_cut_tab.as<int2>();
  std::vector<double> key = S.prm_tables;
  key.push_back(prm->tp);
  key.push_back((double)prm->thr_table_len);
  key.push_back((double)lw);
  if (S.dev_tables != key) {  // else the tables of the slot are those of its last run: nothing to copy
    CK(cudaMemcpyAsync(d_tbl, S.prm_tables.data(), ptab * 8, cudaMemcpyHostToDevice, st));
    launch_k(k_cut_table, CUT_TAB_N / 256, 256, 0, st, d_tbl, prm->thr_table_len, prm->tp, S.b_cut_tab.as<int2>());
    LAUNCHED();
    S.dev_tables.swap(key);
  }

  i64* d_cnt = S.b_counters.as<i64>();
  int* d_err = (int*)(d_cnt + CNT_ERR);
  stage_begin(c, "signal");
  {
    // every zero fill of the run in one launch (largest region first: the raw signal), charged to the stage
    // that needs the largest one
    ZeroList z;
    z.n = 0;
    size_t acc = 0;
    auto add = [&](void* p, size_t bytes) {
      if (!bytes) return;
      const size_t n16 = (bytes + 15) / 16;  // every buffer has at least 256 bytes of slack behind `bytes`
      acc = std::max(acc, n16);
      z.p[z.n] = (uint4*)p;
      z.n16[z.n++] = n16;
    };
    add(c->b_yraw.p, (size_t)L * 4);
    add(c->b_sflag.p, (size_t)L);
    add(c->b_run_cnt.p, (size_t)std::max(NR, 1) * 4);
    add(c->b_tile_state.p, n_groups * 8);  // group sums of k_smooth
    add(d_cnt, CNT_SLOTS * 8);
    const size_t g = (acc + 255) / 256;
    launch_k(k_zero_multi, (unsigned)(g < (size_t)c->n_sm * 16 ? g : (size_t)c->n_sm * 16), 256, 0, st, z);
    LAUNCHED();
  }

  const int* d_tint_island_off = S.b_tint_island_off.as<int>();
  const int* d_tint_rep_off = S.b_tint_rep_off.as<int>();
  const int* d_island_sample_off = S.b_island_sample_off.as<int>();
  const int* d_island_tint = S.b_island_tint.as<int>();

  // ================= phase 1: signal -> smoothed signal -> candidates, threshold =================
  if (S.n_sig_work > 0) {
    launch_k(k_signal, S.n_sig_work, SIG_THREADS, SIG_BINS * 4, st, S.b_sig_work.as<SigWork>(), S.b_rep_iv_off.as<int>(),
                                                               S.b_rep_weight.as<int>(), S.b_rep_fs.as<int>(),
                                                               S.b_rep_fe.as<int>(), prm->ignore_ends, c->b_yraw.as<int>());
    LAUNCHED();
  }
  if (S.n_sig_direct > 0) {  // no histogram: no shared memory, full occupancy
    launch_k(k_signal, S.n_sig_direct, SIG_THREADS, 0, st, S.b_sig_work.as<SigWork>() + S.n_sig_work, S.b_rep_iv_off.as<int>(),
                                                      S.b_rep_weight.as<int>(), S.b_rep_fs.as<int>(),
                                                      S.b_rep_fe.as<int>(), prm->ignore_ends, c->b_yraw.as<int>());
    LAUNCHED();
  }

  stage_begin(c, "smooth");
  {
    // group totals | per tile: candidate / positive ballot words, packed counts
    unsigned long long* d_gsum = c->b_tile_state.as<unsigned long long>();
    int2* d_toff = (int2*)(d_gsum + n_groups);  // (candidates, positives) before every tile (k_tile_prefix)
    u32* d_cmask = (u32*)(d_toff + S.n_tiles);
    u32* d_pmask = d_cmask + (size_t)S.n_tiles * TILE_WORDS;
    u32* d_tcnt = d_pmask + (size_t)S.n_tiles * TILE_WORDS;
    const size_t sm = (size_t)p1_smem_layout(lw).total;
    // FRS_SMOOTH_OCC (dev knob): register budget of k_smooth as CTAs per SM
    static const int occ = [] { const char* e = getenv("FRS_SMOOTH_OCC"); return e ? atoi(e) : 16; }();
#define FRS_LAUNCH_SMOOTH(MINB)                                                                                              \
  do {                                                                                                                       \
    if (sm > 48 * 1024) CK(cudaFuncSetAttribute(k_smooth<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));      \
    launch_k(k_smooth<MINB>, S.n_tiles, GAUSS_THREADS, sm, st, S.b_tiles.as<TileWork>(), d_island_sample_off,           \
             c->b_yraw.as<int>(), d_gw, lw, c->b_y.as<double>(), d_cmask, d_pmask, d_tcnt, d_gsum);                          \
  } while (0)
    if (occ == 10) FRS_LAUNCH_SMOOTH(10);
    else if (occ == 12) FRS_LAUNCH_SMOOTH(12);
    else FRS_LAUNCH_SMOOTH(16);
#undef FRS_LAUNCH_SMOOTH
    LAUNCHED();
    stage_begin(c, "lists");
    launch_k(k_tile_prefix, (unsigned)n_groups, TILE_GROUP, 0, st, S.n_tiles, d_tcnt, d_gsum, d_toff);
    LAUNCHED();
    launch_k(k_tile_lists, S.n_tiles, GAUSS_THREADS, 0, st, S.b_tiles.as<TileWork>(), S.n_tiles, d_island_sample_off,
                                                       d_island_tint, d_tint_island_off, T, d_cmask, d_pmask, d_tcnt,
                                                       d_toff, c->b_y.as<double>(), c->b_cand_flat.as<int>(),
                                                       c->b_vbuf.as<double>(), c->b_tint_pos_off.as<int>(), d_cnt + CNT_K);
    LAUNCHED();
  }

  if (S.tl[7]) cudaEventRecord(S.tl[7], st);
  // ================= phase 2: fixed candidates, subproblems, coverage, DP =================
  // (the number of candidates stays on the device: grid-stride launches sized by its upper bound)
  stage_begin(c, "cand_meta");
  const i64* d_K = d_cnt + CNT_K;
  const int g_cand = gs_grid(KMAX, 256);
  launch_k(k_cand_meta, g_cand, 256, 0, st, c->b_cand_flat.as<int>(), d_K, d_island_sample_off, NI,
                                      c->b_cand_island.as<int>(), c->b_island_cand_off.as<int>());
  LAUNCHED();
  // The coverage matrix only needs the candidate list, and its chain (block offsets per tint, scan, k_coverage) runs
  // on a side stream.  k_threshold fills the SMs by itself (beside it k_coverage only doubled both times), so the
  // tiny offset kernels run beside k_threshold and k_coverage starts when it ends, beside fixed -> subproblems ->
  // plan, which are short latency-bound kernels.  The head stream takes the offsets back before k_plan_finish and
  // the matrix before the DP fork.
  cudaStream_t cs = c->side[0];
  CK(cudaEventRecord(c->ev_cov[0], st));
  CK(cudaStreamWaitEvent(cs, c->ev_cov[0], 0));
  // coverage block offsets per tint (rows = candidates of the tint, stride Rp)
  launch_k(k_tint_cov_sizes, cdiv(T + 1, 256), 256, 0, cs, T, d_tint_island_off, c->b_island_cand_off.as<int>(), d_tint_rep_off,
           c->b_tint_cand_off.as<int>(), c->b_cov_sz.as<i64>());
  LAUNCHED();
  { int r = scan_exclusive_on<i64, i64>(c, cs, c->b_bsum_cov, c->b_cov_sz.as<i64>(), T, c->b_tint_cov_off.as<i64>()); if (r) return r; }
  CK(cudaEventRecord(c->ev_cov[1], cs));
  stage_begin(c, "threshold");
  launch_k(k_threshold, T, THR_THREADS, 0, st, S.b_tint_order.as<int>(), d_tint_island_off, d_island_sample_off,
                                         c->b_tint_pos_off.as<int>(), prm->vf, c->b_vbuf.as<double>(),
                                         c->b_leaf_len.as<int>(), c->b_leaf_sum.as<double>(), c->b_thr.as<double>());
  LAUNCHED();
  CK(cudaEventRecord(c->ev_cov[3], st));
  CK(cudaStreamWaitEvent(cs, c->ev_cov[3], 0));
  stage_begin(c, "coverage", cs);
  launch_k(k_coverage, dim3((unsigned)std::max(S.n_cov_tiles, 1), COV_CHUNKS), COV_THREADS, 0, cs,
           S.b_cov_tiles.as<RepTile>(), d_tint_rep_off, c->b_tint_cand_off.as<int>(), c->b_tint_cov_off.as<i64>(),
           S.b_rep_iv_off.as<int>(), S.b_rep_fs.as<int>(), S.b_rep_fe.as<int>(), c->b_cand_flat.as<int>(), c->b_P.as<u32>(),
           T, S.n_cov_tiles > 0 ? cp.P : -1);
  LAUNCHED();
  CK(cudaEventRecord(c->ev_cov[2], cs));

  stage_begin(c, "fixed");
  launch_k(k_fixed_a, g_cand, 256, 0, st, d_K, c->b_cand_flat.as<int>(), c->b_cand_island.as<int>(),
                                    c->b_island_cand_off.as<int>(), d_island_tint, c->b_y.as<double>(),
                                    c->b_thr.as<double>(), c->b_fixed0.as<u8>(), c->b_fixed1.as<u8>());
  LAUNCHED();
  launch_k(k_fixed_b, g_cand, 256, 0, st, d_K, c->b_cand_flat.as<int>(), c->b_cand_island.as<int>(),
                                    c->b_island_cand_off.as<int>(), c->b_y.as<double>(), prm->mps,
                                    c->b_fixed0.as<u8>(), c->b_fixed1.as<u8>(), d_err);
  LAUNCHED();
  stage_begin(c, "subproblems");
  const int slab_words = c->opt_slab_words;
  const int keep = c->opt_keep_tables;
  launch_k(k_sub_build, g_cand, 256, 0, st, d_K, c->b_fixed1.as<u8>(), c->b_cand_island.as<int>(),
                                      c->b_island_cand_off.as<int>(), d_island_tint, d_tint_rep_off,
                                      S.b_tint_read_off.as<int>(), slab_words, dp_warp_words(), keep,
                                      c->b_sub_start.as<int>(), c->b_sub_n.as<int>(), c->b_sub_tint.as<int>(),
                                      c->b_sub_info.as<int>(), c->b_sub_slabs.as<int>(), c->b_sub_tab_off.as<i64>(),
                                      d_cnt + CNT_PLAN, d_err);
  LAUNCHED();
  CK(cudaStreamWaitEvent(st, c->ev_cov[1], 0));  // the offsets of the coverage blocks
  launch_k(k_plan_finish, 1, 32, 0, st, d_cnt, c->b_tint_cov_off.as<i64>(), T, c->b_bases.as<int>(), c->b_cursor.as<int>());
  LAUNCHED();

  if (S.tl[8]) cudaEventRecord(S.tl[8], st);
  launch_k(k_copy_flags, g_cand, 256, 0, st, d_K, c->b_fixed1.as<u8>(), c->b_dpfinal.as<u8>());
  LAUNCHED();
  {
    stage_begin(c, "dp_plan");
    launch_k(k_sub_fill, gs_grid(NSUB_MAX, 256, 148 * 4), 256, 0, st, d_cnt, cp, c->b_sub_info.as<int>(), c->b_sub_slabs.as<int>(),
                                                                c->b_bases.as<int>(), c->b_cursor.as<int>(),
                                                                c->b_work.as<DpWork>(), c->b_split_list.as<int>());
    LAUNCHED();
    launch_k(k_zero_tab, 148 * 4, 256, 0, st, d_cnt, cp, c->b_tab.as<int>());
    LAUNCHED();
    stage_begin(c, "dp");
    DpArgs A;
    A.sub_start = c->b_sub_start.as<int>(); A.sub_n = c->b_sub_n.as<int>(); A.sub_tint = c->b_sub_tint.as<int>();
    A.sub_info = c->b_sub_info.as<int>(); A.sub_tab_off = c->b_sub_tab_off.as<i64>();
    A.tint_rep_off = d_tint_rep_off; A.tint_cand_off = c->b_tint_cand_off.as<int>();
    A.tint_cov_off = c->b_tint_cov_off.as<i64>(); A.rep_weight = S.b_rep_weight.as<int>();
    A.cand_flat = c->b_cand_flat.as<int>(); A.P = c->b_P.as<u32>();
    A.thr_table = d_tbl; A.thr_table_len = prm->thr_table_len; A.tp = prm->tp; A.cut_tab = d_cut;
    A.lo = prm->lo; A.keep_tables = keep;
    A.tab = c->b_tab.as<int>(); A.final_flag = c->b_dpfinal.as<u8>(); A.err = d_err;
    A.cnt = d_cnt; A.caps = cp; A.bases = c->b_bases.as<int>(); A.cursor = c->b_cursor.as<int>();
    A.m_cap = dp_max_n;
    A.sub_left = c->b_sub_slabs.as<int>();
    const DpWork* wl = c->b_work.as<DpWork>();
    // persistent CTAs per SM of the six classes (development knob: FRS_DP_GRIDS="8,3,6,3,1,1")
    static int dp_cps[DP_CLASSES] = {8, 3, 6, 3, 1, 1};
    static bool dp_cps_read = false;
    if (!dp_cps_read) {
      dp_cps_read = true;
      if (const char* e = getenv("FRS_DP_GRIDS")) sscanf(e, "%d,%d,%d,%d,%d,%d", &dp_cps[0], &dp_cps[1], &dp_cps[2], &dp_cps[3], &dp_cps[4], &dp_cps[5]);
    }
    // the classes are independent: persistent launches on side streams, so that the few long CTAs of the
    // large classes overlap the many short items of the small ones (fork / join with events)
    CK(cudaStreamWaitEvent(st, c->ev_cov[2], 0));  // the coverage matrix (side stream)
    CK(cudaEventRecord(c->ev_fork, st));
    bool used[FRS_SIDE_STREAMS] = {};
    // FRS_DP_TIMELINE=1 (development): end of every class's launch relative to the fork, printed after the run
    static const bool dp_tl = getenv("FRS_DP_TIMELINE") != nullptr;
    static cudaEvent_t tl_ev[8] = {};
    if (dp_tl && !tl_ev[0])
      for (auto& e : tl_ev) cudaEventCreate(&e);
    if (dp_tl) cudaEventRecord(tl_ev[7], st);
    const int top = (A.m_cap > DP_SMEM_MAX_N) ? DP_CLASSES - 1 : DP_CLASSES - 2;  // class 5 needs n > 56
    for (int k = top; k >= 0; --k) {  // longest-running classes first
      const int sidx = k >= 2 ? k - 2 : 4 + k;  // a stream per class
      cudaStream_t ks = c->side[sidx];
      CK(cudaStreamWaitEvent(ks, c->ev_fork, 0));
      g_pdl_chain = nullptr;  // plain launches below
      switch (k) {
        case 0: k_dp_warp<8><<<c->n_sm * dp_cps[0], DPW_WARPS * 32, DPW_WARPS * sizeof(DpWarpSmem<8>), ks>>>(A, wl, 0); break;
        case 1: k_dp_warp<16><<<c->n_sm * dp_cps[1], DPW_WARPS * 32, DPW_WARPS * sizeof(DpWarpSmem<16>), ks>>>(A, wl, 1); break;
        case 2: { const int sm = dp_smem_layout(16, DPT_MAXW, 1).total; k_dp<128><<<c->n_sm * dp_cps[2], 128, sm, ks>>>(A, wl, 2, sm); break; }
        case 3: { const int sm = dp_smem_layout(32, DPT_MAXW, 1).total; k_dp<256><<<c->n_sm * dp_cps[3], 256, sm, ks>>>(A, wl, 3, sm); break; }
        default: { const int sm = 226 * 1024; k_dp<DP_BIG_THREADS><<<c->n_sm * dp_cps[k], DP_BIG_THREADS, sm, ks>>>(A, wl, k, sm); break; }
      }
      LAUNCHED();
      CK(cudaEventRecord(c->ev_join[sidx], ks));
      if (dp_tl) cudaEventRecord(tl_ev[k], ks);
      used[sidx] = true;
    }
    {
      // only CTA classes (streams 0..3) have split subproblems: their solver starts as soon as those are
      // done and runs beside the warp classes
      cudaStream_t ss = c->side[0];
      for (int k = 1; k < 4; ++k)
        if (used[k]) CK(cudaStreamWaitEvent(ss, c->ev_join[k], 0));
      k_dp_solve<<<148, DPS_THREADS, dps_smem_bytes(dp_max_n, dp_stage_n), ss>>>(A, c->b_split_list.as<int>(), dp_max_n,
                                                                                 dp_stage_n);
      LAUNCHED();
      CK(cudaEventRecord(c->ev_join[0], ss));
      if (dp_tl) cudaEventRecord(tl_ev[6], ss);
      used[0] = true;
    }
    for (int k = 0; k < FRS_SIDE_STREAMS; ++k)
      if (used[k]) CK(cudaStreamWaitEvent(st, c->ev_join[k], 0));
    if (dp_tl) {
      cudaStreamSynchronize(st);
      float t[7] = {};
      for (int k = 0; k <= 6; ++k)
        if (k == 6 || k <= top) cudaEventElapsedTime(&t[k], tl_ev[7], tl_ev[k]);
      fprintf(stderr, "[frs dp timeline] ends after the fork (ms): warp8 %.3f  warp16 %.3f  cta128 %.3f  cta256 %.3f  cta1024 %.3f  big %.3f  solve %.3f\n",
              t[0], t[1], t[2], t[3], t[4], t[5], t[6]);
    }
  }

  // ================= phase 3: refine, final positions, digits =================
  if (S.tl[9]) cudaEventRecord(S.tl[9], st);
  stage_begin(c, "refine");
  int* d_ref_cnt = (int*)(d_cnt + CNT_REF);
  int* d_ref_cnt2 = (int*)(d_cnt + CNT_REF2);
  launch_k(k_final_mark, g_cand, 256, 0, st, d_K, c->b_dpfinal.as<u8>(), c->b_cand_flat.as<int>(),
                                       c->b_cand_island.as<int>(), c->b_island_cand_off.as<int>(),
                                       c->b_sflag.as<u8>(), c->b_ref_list.as<int2>(), d_ref_cnt);
  LAUNCHED();
  launch_k(k_refine_filter, 148 * 8, 256, 0, st, c->b_ref_list.as<int2>(), d_ref_cnt, c->b_yraw.as<int>(),
                                           c->b_ref_list2.as<int2>(), d_ref_cnt2);
  LAUNCHED();
  launch_k(k_refine, 148 * 8, REF_THREADS, 0, st, c->b_ref_list2.as<int2>(), d_ref_cnt2, c->b_yraw.as<int>(), d_rw, rr, prm->sigma,
                                            c->b_gbuf.as<double>(), c->b_pstate.as<u8>(), c->b_sflag.as<u8>());
  LAUNCHED();

  stage_begin(c, "finals");
  { int r = compact_flags(c, c->b_sflag.as<u8>(), L, c->b_final_flat.as<int>(), d_cnt + CNT_NFIN); if (r) return r; }
  const i64* d_nfin = d_cnt + CNT_NFIN;
  launch_k(k_final_meta, 148 * 4, 256, 0, st, d_nfin, c->b_final_flat.as<int>(), d_island_sample_off,
                                        S.b_island_start.as<int>(), d_island_tint, d_tint_island_off, NI, T,
                                        S.b_final_pos.as<int>(), c->b_final_island.as<int>(),
                                        S.b_tint_final_off.as<int>());
  LAUNCHED();
  launch_k(k_digit_sizes, cdiv(T, 256), 256, 0, st, T, d_tint_rep_off, S.b_tint_final_off.as<int>(), c->b_dig_sz.as<i64>());
  LAUNCHED();
  { int r = scan_exclusive<i64, i64>(c, c->b_dig_sz.as<i64>(), T, S.b_tint_digit_off.as<i64>(), d_cnt + CNT_NDIG); if (r) return r; }
  launch_k(k_seg_cuts, 148 * 4, 256, 0, st, d_nfin, c->b_final_flat.as<int>(), c->b_final_island.as<int>(), d_cut, d_tbl,
                                      prm->thr_table_len, prm->tp, c->b_seg_ty.as<int>(), c->b_seg_tn.as<int>());
  LAUNCHED();

  if (S.tl[10]) cudaEventRecord(S.tl[10], st);
  stage_begin(c, "digits");
  if (S.n_dig_tiles > 0) {
    launch_k(k_digits, dim3((unsigned)S.n_dig_tiles, DIG_CHUNKS), DIG_THREADS, 0, st, 
        S.b_dig_tiles.as<RepTile>(), d_tint_rep_off, S.b_tint_final_off.as<int>(), S.b_tint_digit_off.as<i64>(),
        S.b_rep_iv_off.as<int>(), S.b_rep_fs.as<int>(), S.b_rep_fe.as<int>(), c->b_final_flat.as<int>(),
        c->b_seg_ty.as<int>(), c->b_seg_tn.as<int>(), S.b_digits.as<u8>(), c->b_run_cnt.as<int>(), d_err, d_cnt, cp);
    LAUNCHED();
  }

  stage_begin(c, "runs");
  { int r = scan_exclusive<int, int>(c, c->b_run_cnt.as<int>(), NR, c->b_run_off.as<int>(), d_cnt + CNT_NRUN); if (r) return r; }
  if (N > 0) {
    launch_k(k_gap_count, cdiv(N, 256), 256, 0, st, N, S.b_read_rep.as<int>(), c->b_run_off.as<int>(), c->b_gap_cnt.as<int>());
    LAUNCHED();
  }
  { int r = scan_exclusive<int, int>(c, c->b_gap_cnt.as<int>(), N, S.b_read_gap_off.as<int>(), d_cnt + CNT_NGAP); if (r) return r; }
  if (NR > 0) {
    launch_k(k_run_fill, cdiv((i64)NR * 32, 256), 256, 0, st, NR, S.b_rep_tint.as<int>(), d_tint_rep_off,
                                                        S.b_tint_final_off.as<int>(), S.b_tint_digit_off.as<i64>(),
                                                        S.b_digits.as<u8>(), c->b_run_off.as<int>(), c->b_runs.as<int2>(),
                                                        d_cnt, cp);
    LAUNCHED();
  }

  stage_begin(c, "gaps");
  if (N > 0) {
    GapArgs G;
    G.n_reads = N; G.read_rep = S.b_read_rep.as<int>(); G.read_strand = S.b_read_strand.as<u8>();
    G.read_len = S.b_read_len.as<int>(); G.read_iv_off = S.b_read_iv_off.as<int>();
    G.read_seq_off = S.b_read_seq_off.as<i64>(); G.read_tint = S.b_read_tint.as<int>();
    G.riv_ts = S.b_riv_ts.as<int>(); G.riv_te = S.b_riv_te.as<int>(); G.riv_qs = S.b_riv_qs.as<int>();
    G.riv_qe = S.b_riv_qe.as<int>(); G.riv_cig_off = S.b_riv_cig_off.as<int>(); G.cigar = S.b_cigar.as<u32>();
    G.seq_a = S.seq_resident ? S.b_seq_a.as<u32>() : S.b_clip_a.as<u32>();
    G.seq_t = S.seq_resident ? S.b_seq_t.as<u32>() : S.b_clip_t.as<u32>();
    G.run_off = c->b_run_off.as<int>();
    G.runs = c->b_runs.as<int2>(); G.tint_final_off = S.b_tint_final_off.as<int>();
    G.final_pos = S.b_final_pos.as<int>(); G.read_gap_off = S.b_read_gap_off.as<int>();
    G.read_head = S.b_read_head.as<int>(); G.gap_rec = S.b_gap_rec.as<int>(); G.err = d_err;
    G.clip_n = S.b_clip_n.as<int>(); G.clip_words = S.b_clip_words.as<int>(); G.clip_off = S.b_clip_off.as<i64>();
    G.seq_resident = S.seq_resident ? 1 : 0;
    G.cls_count = S.b_poly_cls.as<int>();
    G.task_order = S.b_task_order.as<int>(); G.task_res = S.b_task_res.as<PolyRes>();
    G.long_class = c->opt_poly_long_class;
    G.cnt = d_cnt; G.caps = cp;
    G.edge = S.edge_words ? S.b_seq_edge.as<u32>() : nullptr;
    G.edge_words = S.edge_words;
    G.clip_eoff = S.edge_words ? S.b_clip_eoff.as<int>() : nullptr;
    launch_k(k_gap_prep, cdiv((i64)N * 2, 128), 128, 0, st, G); LAUNCHED();
    launch_k(k_gap_sizes, gs_grid((i64)N * 2, 128, 148 * 8), 128, 0, st, G); LAUNCHED();
    stage_end(c);
    // ---- TAIL of the run, on its own stream: it touches only buffers of this slot, so the head of the next
    // batch (compute stream) runs beside it -- the clip fetch is bound by the bus, not by the SMs ----
    cudaStream_t tl = c->st_tail;
    if (S.tl[3]) cudaEventRecord(S.tl[3], st);
    CK(cudaEventRecord(S.ev_head, st));
    CK(cudaStreamWaitEvent(tl, S.ev_head, 0));
    dev_zero(c, tl, S.b_poly_cls.p, (2 * POLY_CLASSES + 1) * 4);
    if (!S.seq_resident) {
      stage_begin(c, "clip_fetch", tl);
      // compact offsets of the clips' plane words (total at [2N]), then the words themselves, straight from
      // the caller's pinned planes
      { int r = scan_exclusive_on<int, i64>(c, tl, S.b_bsum_tail, S.b_clip_words.as<int>(), (i64)N * 2, S.b_clip_off.as<i64>(), d_cnt + CNT_CLIPW); if (r) return r; }
      // a few CTAs only: the kernel waits on the bus (the bus allows a few hundred reads in flight, not tens of
      // thousands), and loads from host memory that are pending for microseconds fill the miss queues of the SM
      // they run on -- the other SMs belong to the head of the next batch meanwhile
      launch_k(k_clip_gather, S.edge_words ? 24 : c->n_sm, 256, 0, tl, G, S.zc_a, S.zc_t, S.b_clip_a.as<u32>(), S.b_clip_t.as<u32>());
      LAUNCHED();
    }
    stage_begin(c, "poly", tl);
    launch_k(k_poly_filter, cdiv((i64)N * 4, 128), 128, 0, tl, G, S.b_poly_flag.as<u8>()); LAUNCHED();
    launch_k(k_poly_bases, 1, 32, 0, tl, G.cls_count, G.long_class, d_err + 2); LAUNCHED();
    launch_k(k_poly_scatter, cdiv((i64)N * 4, 256), 256, 0, tl, N * 4, G.clip_n, S.b_poly_flag.as<u8>(), G.cls_count,
                                                           G.task_order, d_cnt, cp); LAUNCHED();
    // long clips (one warp each) run beside the short ones (one thread each)
    CK(cudaEventRecord(c->ev_tfork, tl));
    CK(cudaStreamWaitEvent(c->st_tail_side, c->ev_tfork, 0));
    launch_k(k_poly_long, 148 * 4, 128, 0, c->st_tail_side, G); LAUNCHED();
    CK(cudaEventRecord(c->ev_tjoin, c->st_tail_side));
    launch_k(k_poly_scan, cdiv((i64)N * 4, 128), 128, 0, tl, G); LAUNCHED();
    CK(cudaStreamWaitEvent(tl, c->ev_tjoin, 0));
    launch_k(k_gap_finish, cdiv(N, 128), 128, 0, tl, G); LAUNCHED();
    stage_end(c, tl);
  } else {
    stage_end(c);
    CK(cudaEventRecord(S.ev_head, st));
    CK(cudaStreamWaitEvent(c->st_tail, S.ev_head, 0));
  }
  // the ONE read-back of the run: every count, the plan and the assert channel
  CK(cudaEventRecord(S.ev_ran, c->st_tail));
  if (S.tl[4]) cudaEventRecord(S.tl[4], c->st_tail);
  CK(cudaMemcpyAsync(S.h_cnt, d_cnt, CNT_SLOTS * 8, cudaMemcpyDeviceToHost, c->st_tail));
  CK(cudaEventRecord(S.ev_cnt, c->st_tail));
  CK(cudaGetLastError());
  if (prof_run) {
    const auto tr2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[frs host profile] run: buffers %.3f ms, launches enqueued %.3f ms\n",
            std::chrono::duration<double, std::milli>(tr1 - tr0).count(), std::chrono::duration<double, std::milli>(tr2 - tr1).count());
  }
  S.enqueued = true;
  S.ran = false;
  return 0;
}

// waits for the run of the slot; repeats it with larger buffers if a capacity was missed
static int finish_run(frs_context* c, Slot& S) {
  if (!S.enqueued) return fail(c, FRS_ERR_STATE, "frs_run: nothing enqueued");
  static const bool prof_fin = getenv("FRS_HOST_PROFILE") != nullptr;
  for (int attempt = 0;; ++attempt) {
    const auto tf0 = std::chrono::steady_clock::now();
    CK(cudaEventSynchronize(S.ev_cnt));
    if (prof_fin)
      fprintf(stderr, "[frs host profile] run: waited %.3f ms for the device (attempt %d)\n",
              std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tf0).count(), attempt);
    const i64* h = S.h_cnt;
    const int* he = (const int*)(h + CNT_ERR);
    Caps& cp = c->caps;
    const i64 K = h[CNT_K], COV = h[CNT_COV], NWORK = h[CNT_NWORK], NSPLIT = h[CNT_PLAN + PLAN_SPLIT],
              TAB = h[CNT_PLAN + PLAN_TAB], NFIN = h[CNT_NFIN], NDIG = h[CNT_NDIG], NRUN = h[CNT_NRUN] & 0xffffffffLL,
              NGAP = h[CNT_NGAP] & 0xffffffffLL, CLIPW = h[CNT_CLIPW];
    if (NWORK > 0x7fffffffLL) return fail(c, FRS_ERR_LIMIT, "too many DP work items in one batch (%lld)", (long long)NWORK);
    bool miss = false;
    const Caps& used = S.caps_used;  // what THIS run was launched with (the context's may have grown since)
    auto grow = [&](i64& cap, i64 had, i64 need) {
      if (need > had) { cap = std::max<i64>(cap, need + need / 8 + 64); miss = true; }
    };
    // stages run in this order; a stage behind a missed capacity reports garbage, so stop at the first miss
    grow(cp.P, used.P, COV); grow(cp.work, used.work, NWORK); grow(cp.split, used.split, NSPLIT); grow(cp.tab, used.tab, TAB);
    if (!miss) grow(cp.dig, used.dig, NDIG);
    if (!miss) grow(cp.runs, used.runs, NRUN);
    if (!miss) grow(cp.gaps, used.gaps, NGAP);
    if (!miss && !S.seq_resident) grow(cp.clipw, used.clipw, CLIPW);
    if (he[0] && !miss) {
      if (he[0] == DEVERR_DP_SMEM)
        return fail(c, FRS_ERR_LIMIT, "subproblem with %d candidates exceeds the DP kernel's shared-memory budget "
                                      "(max_problem_size too large for this build)", he[1]);
      if (he[0] == DEVERR_SCORE_RANGE)
        return fail(c, FRS_ERR_LIMIT, "tint %d: candidates x reads leaves the 30-bit range of the DP scores", he[1]);
      return fail(c, FRS_ERR_ASSERT, "AssertionError: %s [item %d]", deverr_text(he[0]), he[1]);
    }
    if (!miss) {
      S.n_cand = K;
      S.n_sub = h[CNT_PLAN + PLAN_NSUB];
      S.cov_elems = COV;
      S.tab_elems = TAB;
      S.clip_words = S.seq_resident ? 0 : CLIPW;
      S.st_h2d_run = S.seq_resident ? 0 : CLIPW * 8;  // plane words the device fetched from host memory
      S.st_d2h_run = CNT_SLOTS * 8;
      S.st_poly_tasks = he[2];
      S.st_poly_long = he[3];
      S.sizes.n_final = NFIN;
      S.sizes.n_digit_bytes = NDIG;
      S.sizes.n_gap_records = NGAP;
      S.sizes.n_candidates = K;
      S.sizes.n_subproblems = S.n_sub;
      S.sizes.dp_cells = h[CNT_PLAN + PLAN_CELLS];
      S.sizes.dp_read_cells = h[CNT_PLAN + PLAN_RCELLS];
      S.sizes.max_subproblem = (int)h[CNT_PLAN + PLAN_MAXALL];
      S.sizes.pad = 0;
      S.ran = true;
      c->last_run = (int)(&S - c->slot);
      return 0;
    }
    if (attempt >= 8) return fail(c, FRS_ERR_LIMIT, "internal: buffer capacities do not converge");
    c->reruns++;
    { int r = quiesce(c); if (r) return r; }
    { int r = enqueue_run(c, S); if (r) return r; }
  }
}

#define D2H(dst, buf, bytes)                                                                              \
  do {                                                                                                    \
    if ((dst) && (bytes) > 0) CK(cudaMemcpyAsync(dst, S.buf.p, (size_t)(bytes), cudaMemcpyDeviceToHost, c->st_out)); \
  } while (0)

static int enqueue_download(frs_context* c, Slot& S, const frs_result* o) {
  if (!S.ran) return fail(c, FRS_ERR_STATE, "frs_download: no results (call frs_run first)");
  const int T = S.hb.n_tints, N = S.hb.n_reads;
  CK(cudaStreamWaitEvent(c->st_out, S.ev_ran, 0));
  if (S.tl[5]) cudaEventRecord(S.tl[5], c->st_out);
  D2H(o->tint_final_off, b_tint_final_off, (size_t)(T + 1) * 4);
  D2H(o->final_pos, b_final_pos, S.sizes.n_final * 4);
  D2H(o->tint_digit_off, b_tint_digit_off, (size_t)(T + 1) * 8);
  D2H(o->digits, b_digits, S.sizes.n_digit_bytes);
  D2H(o->read_head, b_read_head, (size_t)N * 32);
  D2H(o->read_gap_off, b_read_gap_off, (size_t)(N + 1) * 4);
  D2H(o->gap_rec, b_gap_rec, S.sizes.n_gap_records * 12);
  CK(cudaEventRecord(S.ev_down, c->st_out));
  if (S.tl[6]) cudaEventRecord(S.tl[6], c->st_out);
  S.down_pending = true;
  return 0;
}

extern "C" {

int frs_upload(frs_context* c, const frs_batch* b) {
  if (!c || !b) return fail(c, FRS_ERR_ARG, "frs_upload: NULL argument");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[c->cur];
  if (S.busy) return fail(c, FRS_ERR_STATE, "frs_upload: a submitted batch is in flight in this slot (frs_fetch it first)");
  int r = stage_upload(c, S, b);
  if (r) return r;
  CK(cudaEventSynchronize(S.ev_up));  // synchronous API: the caller's arrays are free on return
  if ((r = enqueue_prep(c, S, c->stream))) return r;  // the inputs are complete before the first frs_run is timed
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int frs_run(frs_context* c, const frs_params* prm, frs_result_sizes* sizes_out) {
  if (!c || !prm) return fail(c, FRS_ERR_ARG, "frs_run: NULL argument");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[c->cur];
  if (!S.uploaded) return fail(c, FRS_ERR_STATE, "frs_run: no batch uploaded");
  int r = check_params(c, prm);
  if (r) return r;
  keep_params(S, prm);
  if ((r = enqueue_run(c, S))) return r;
  if ((r = finish_run(c, S))) return r;
  if (sizes_out) *sizes_out = S.sizes;
  return 0;
}

int frs_segment_batch(frs_context* c, const frs_batch* batch, const frs_params* prm, frs_result_sizes* sizes) {
  int r = frs_upload(c, batch);
  if (r) return r;
  return frs_run(c, prm, sizes);
}

int frs_download(frs_context* c, const frs_result* o) {
  if (!c || !o) return fail(c, FRS_ERR_ARG, "frs_download: NULL argument");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[c->cur];
  int r = enqueue_download(c, S, o);
  if (r) return r;
  CK(cudaEventSynchronize(S.ev_down));
  S.down_pending = false;
  return 0;
}

// ---- pipelined form: two batches in flight per context, one host thread ----
int frs_submit(frs_context* c, const frs_batch* b, const frs_params* prm, int* ticket) {
  if (!c || !b || !prm || !ticket) return fail(c, FRS_ERR_ARG, "frs_submit: NULL argument");
  CK(cudaSetDevice(c->device));
  const int k = (c->cur + 1) % FRS_SLOTS;
  Slot& S = c->slot[k];
  if (S.busy) return fail(c, FRS_ERR_STATE, "frs_submit: %d batches are in flight already (frs_fetch the oldest first)", FRS_SLOTS);
  int r = check_params(c, prm);
  if (r) return r;
  static const bool prof = getenv("FRS_HOST_PROFILE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  if ((r = stage_upload(c, S, b))) return r;
  const auto t1 = std::chrono::steady_clock::now();
  keep_params(S, prm);
  if ((r = enqueue_run(c, S))) return r;
  if (prof) {
    const auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[frs host profile] submit: upload side %.3f ms, run side %.3f ms\n",
            std::chrono::duration<double, std::milli>(t1 - t0).count(), std::chrono::duration<double, std::milli>(t2 - t1).count());
  }
  S.busy = true;
  c->cur = k;
  *ticket = k;
  return 0;
}

int frs_wait(frs_context* c, int ticket, frs_result_sizes* sizes_out) {
  if (!c || ticket < 0 || ticket >= FRS_SLOTS) return fail(c, FRS_ERR_ARG, "frs_wait: bad ticket");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[ticket];
  if (!S.busy) return fail(c, FRS_ERR_STATE, "frs_wait: no batch in flight under this ticket");
  int r = finish_run(c, S);
  if (r) { S.busy = false; return r; }
  if (sizes_out) *sizes_out = S.sizes;
  return 0;
}

int frs_fetch_start(frs_context* c, int ticket, const frs_result* o) {
  if (!c || !o || ticket < 0 || ticket >= FRS_SLOTS) return fail(c, FRS_ERR_ARG, "frs_fetch: bad argument");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[ticket];
  if (!S.busy) return fail(c, FRS_ERR_STATE, "frs_fetch: no batch in flight under this ticket");
  if (S.down_pending) return fail(c, FRS_ERR_STATE, "frs_fetch_start: the results of this ticket are being copied already");
  int r = 0;
  if (!S.ran) r = finish_run(c, S);
  if (!r) r = enqueue_download(c, S, o);
  if (r) S.busy = false;
  return r;
}

int frs_fetch_finish(frs_context* c, int ticket) {
  if (!c || ticket < 0 || ticket >= FRS_SLOTS) return fail(c, FRS_ERR_ARG, "frs_fetch: bad argument");
  CK(cudaSetDevice(c->device));
  Slot& S = c->slot[ticket];
  if (!S.busy || !S.down_pending) return fail(c, FRS_ERR_STATE, "frs_fetch_finish: no copy in flight under this ticket");
  int r = 0;
  if (cudaEventSynchronize(S.ev_down) != cudaSuccess) r = fail(c, FRS_ERR_CUDA, "frs_fetch: %s", cudaGetErrorString(cudaGetLastError()));
  if (!r && S.tl[6] && c->ev_base) {
    cudaEventSynchronize(S.tl[6]);
    float t[11];
    for (int e = 0; e < 11; ++e) cudaEventElapsedTime(&t[e], c->ev_base, S.tl[e]);
    fprintf(stderr, "[frs timeline] slot %d: h2d %.3f-%.3f  head %.3f-%.3f  tail -%.3f  d2h %.3f-%.3f ms | head parts: signal+smooth+lists %.3f  "
            "threshold..plan %.3f  coverage+dp %.3f  refine+finals %.3f  digits..gaps %.3f\n", ticket, t[0], t[1], t[2],
            t[3], t[4], t[5], t[6], t[7] - t[2], t[8] - t[7], t[9] - t[8], t[10] - t[9], t[3] - t[10]);
  }
  S.down_pending = false;
  S.busy = false;
  return r;
}

int frs_fetch(frs_context* c, int ticket, const frs_result* o) {
  int r = frs_fetch_start(c, ticket, o);
  if (r) return r;
  return frs_fetch_finish(c, ticket);
}

int frs_get_intermediate(frs_context* c, int which, void* dst, size_t cap, size_t* bytes) {
  if (!c || !bytes) return fail(c, FRS_ERR_ARG, "frs_get_intermediate: NULL argument");
  const Slot& S = c->slot[c->last_run];
  if (!S.ran) return fail(c, FRS_ERR_STATE, "frs_get_intermediate: nothing has run");
  CK(cudaSetDevice(c->device));
  const void* src = nullptr;
  size_t sz = 0;
  const i64 L = S.hb.n_samples, K = S.n_cand, NS = S.n_sub;
  switch (which) {
    case FRS_TAP_Y_RAW: src = c->b_yraw.p; sz = L * 4; break;
    case FRS_TAP_Y: src = c->b_y.p; sz = L * 8; break;
    case FRS_TAP_THR: src = c->b_thr.p; sz = (size_t)S.hb.n_tints * 8; break;
    case FRS_TAP_CAND: src = c->b_cand_flat.p; sz = K * 4; break;
    case FRS_TAP_FIXED: src = c->b_fixed1.p; sz = K; break;
    case FRS_TAP_DP_FINAL: src = c->b_dpfinal.p; sz = K; break;
    case FRS_TAP_SUB_START: src = c->b_sub_start.p; sz = NS * 4; break;
    case FRS_TAP_SUB_N: src = c->b_sub_n.p; sz = NS * 4; break;
    case FRS_TAP_COVERAGE: src = c->b_P.p; sz = S.cov_elems * 4; break;
    case FRS_TAP_DP_TABLES: src = c->b_tab.p; sz = S.tab_elems * 4; break;
    case FRS_TAP_COV_OFF: src = c->b_tint_cov_off.p; sz = (size_t)(S.hb.n_tints + 1) * 8; break;
    case FRS_TAP_SUB_TAB_OFF: src = c->b_sub_tab_off.p; sz = NS * 8; break;
    case FRS_TAP_FINAL_FLAGS: src = c->b_sflag.p; sz = L; break;
    default: return fail(c, FRS_ERR_ARG, "frs_get_intermediate: unknown tap %d", which);
  }
  *bytes = sz;
  size_t n = sz < cap ? sz : cap;
  if (dst && n > 0) {
    CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  return 0;
}

}  // extern "C"
