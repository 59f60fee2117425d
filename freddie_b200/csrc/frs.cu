// frs.cu -- context, pipeline driver and C ABI of libfreddie_b200.so (see include/freddie_b200.h).
// The pipeline replaces segment() (freddie_segment.py:738-844) for a whole batch of tints.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
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

struct frs_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t side[FRS_SIDE_STREAMS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[FRS_SIDE_STREAMS] = {};
  char err[512] = "";
  bool uploaded = false, ran = false;
  bool profiling = false;
  // host copy of the batch sizes and small offset arrays
  frs_batch hb;  // pointers here are DEVICE pointers after upload
  std::vector<int> h_tint_island_off, h_tint_rep_off, h_tint_read_off, h_island_sample_off;
  int n_sig_work = 0, n_sig_direct = 0, n_tiles = 0, n_cov_tiles = 0, n_dig_tiles = 0;
  // device buffers (grow-only)
  std::vector<DBuf*> all;
  DBuf b_tint_island_off, b_tint_rep_off, b_tint_read_off, b_island_start, b_island_sample_off, b_island_tint,
      b_rep_iv_off, b_rep_weight, b_rep_fs, b_rep_fe, b_rep_tint, b_read_rep, b_read_strand, b_read_len,
      b_read_iv_off, b_read_seq_off, b_read_tint, b_riv_ts, b_riv_te, b_riv_qs, b_riv_qe, b_riv_cig_off, b_cigar,
      b_seq_a, b_seq_t;
  DBuf b_sig_work, b_tiles, b_cov_tiles, b_dig_tiles, b_tint_order;
  DBuf b_params;  // thr table | gauss w | refine w
  DBuf b_yraw, b_y, b_sflag, b_bsum, b_cand_flat, b_cand_island, b_island_cand_off, b_tint_cand_off, b_thr, b_vbuf,
      b_leaf_off, b_leaf_len, b_leaf_sum, b_tint_pos_off, b_tile_state, b_fixed0, b_fixed1, b_fixed_list, b_sub_flag, b_sub_fidx, b_sub_start,
      b_sub_n, b_sub_tint, b_sub_info, b_sub_slabs, b_sz_tab, b_sub_tab_off, b_plan, b_work, b_split_list, b_cursor,
      b_cov_sz, b_tint_cov_off, b_P, b_tab, b_dpfinal, b_ref_list, b_ref_list2, b_gbuf, b_pstate, b_final_flat,
      b_final_pos, b_final_island, b_tint_final_off, b_dig_sz, b_tint_digit_off, b_seg_ty, b_seg_tn, b_digits,
      b_run_cnt, b_run_off, b_runs, b_gap_cnt, b_clip_n, b_clip_words, b_clip_off, b_task_order, b_task_res, b_poly_cls, b_poly_flag, b_read_gap_off, b_read_head, b_gap_rec, b_counters, b_stats, b_err;
  i64* h_pin = nullptr;  // pinned scratch for small D2H reads
  // results of the last run
  frs_result_sizes sizes;
  i64 n_cand = 0, n_fixed = 0, n_sub = 0, cov_elems = 0, tab_elems = 0;
  // options (frs_set_option)
  int opt_slab_words = 64, opt_keep_tables = 0, opt_poly_long_class = POLY_LONG_CLASS, opt_lazy_seq = 1;
  // lazy sequence mode: host copies of the small per-read tables, pinned staging for the clip words
  bool seq_resident = false;
  std::vector<int> h_read_len;
  std::vector<u8> h_read_strand;
  std::vector<i64> h_read_seq_off;
  const u32* h_seq_a = nullptr;  // caller's planes (valid until frs_run returns, see the header)
  const u32* h_seq_t = nullptr;
  void* h_stage = nullptr;       // pinned: clip_n (D2H), clip_off + gathered words (H2D)
  size_t h_stage_cap = 0;
  i64 st_h2d_upload = 0, st_h2d_run = 0, st_d2h_run = 0, st_poly_tasks = 0, st_poly_long = 0, st_clip_words = 0;
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

static int ensure(frs_context* c, DBuf& b, size_t bytes) {
  if (bytes < 16) bytes = 16;
  if (b.cap >= bytes) return 0;
  if (b.p) CK(cudaFree(b.p));
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

static void stage_begin(frs_context* c, const char* name) {
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
  if (c->cur_stage >= 0 && c->profiling) cudaEventRecord(c->stages[c->cur_stage].ev1, c->stream);
  c->cur_stage = s;
  if (s >= 0) {
    c->stages[s].used = true;
    if (c->profiling) cudaEventRecord(c->stages[s].ev0, c->stream);
  }
}
static void stage_end(frs_context* c) {
  if (c->cur_stage >= 0 && c->profiling) cudaEventRecord(c->stages[c->cur_stage].ev1, c->stream);
  c->cur_stage = -1;
}
#define LAUNCHED()                                           \
  do {                                                       \
    c->launch_count++;                                       \
    if (c->cur_stage >= 0) c->stages[c->cur_stage].launches++; \
  } while (0)

static inline int cdiv(i64 a, i64 b) { return (int)((a + b - 1) / b); }

// device-wide helpers ------------------------------------------------------------------------
template <typename TIn, typename TOut>
static int scan_exclusive(frs_context* c, const TIn* in, i64 n, TOut* out) {
  if (n <= SCAN_SMALL_MAX && (const void*)in != (const void*)out) {
    k_scan_small<TIn, TOut><<<1, 1024, 0, c->stream>>>(in, (int)n, out); LAUNCHED();
    return 0;
  }
  int nb = cdiv(n > 0 ? n : 1, SCAN_TILE);
  ENS(b_bsum, (size_t)(nb + 1) * 8);
  i64* bs = c->b_bsum.as<i64>();
  k_scan_block_sums<TIn><<<nb, SCAN_THREADS, 0, c->stream>>>(in, n, bs); LAUNCHED();
  k_scan_bsums<<<1, 1024, 0, c->stream>>>(bs, nb); LAUNCHED();
  k_scan_apply<TIn, TOut><<<nb, SCAN_THREADS, 0, c->stream>>>(in, n, bs, out); LAUNCHED();
  return 0;
}
// compaction of byte flags; the count ends up in bsum[nb] and is copied to counters[slot]
static int compact_flags(frs_context* c, const u8* flags, i64 n, int* idx_out, int counter_slot) {
  int nb = cdiv(n > 0 ? n : 1, FLAG_TILE);
  ENS(b_bsum, (size_t)(nb + 1) * 8);
  i64* bs = c->b_bsum.as<i64>();
  k_flag_sums<<<nb, SCAN_THREADS, 0, c->stream>>>(flags, n, bs); LAUNCHED();
  k_scan_bsums<<<1, 1024, 0, c->stream>>>(bs, nb); LAUNCHED();
  k_flag_compact<<<nb, SCAN_THREADS, 0, c->stream>>>(flags, n, bs, idx_out); LAUNCHED();
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + counter_slot, bs + nb, 8, cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}
static int read_counters(frs_context* c, int n) {
  CK(cudaMemcpyAsync(c->h_pin, c->b_counters.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
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
static int check_dev_err(frs_context* c) {
  int* h = (int*)(c->h_pin + 60);  // pinned scratch: [code, item, poly tasks, long poly tasks]
  CK(cudaMemcpyAsync(h, c->b_err.p, 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->st_poly_tasks = h[2];
  c->st_poly_long = h[3];
  if (h[0]) return fail(c, FRS_ERR_ASSERT, "AssertionError: %s [item %d]", deverr_text(h[0]), h[1]);
  return 0;
}

// Lazy sequence mode: after k_gap_prep the clip lengths are known.  Bring them to the host, gather the
// plane words each clip needs from the caller's (host) bit-planes into pinned staging, and send only
// those to the device -- a few per cent of the reads' bases instead of all of them.
static int fetch_clip_words(frs_context* c, int N) {
  cudaStream_t st = c->stream;
  const size_t n_clip = (size_t)N * 2;
  // device: exclusive scan of the per-clip word counts -> compact offsets (total at [2N])
  { int r = scan_exclusive<int, i64>(c, c->b_clip_words.as<int>(), (i64)n_clip, c->b_clip_off.as<i64>()); if (r) return r; }
  // staging layout: [clip_n: 2N int][clip_off: 2N+1 i64][words A][words T]
  const size_t o_off = (n_clip * 4 + 15) & ~(size_t)15;
  const size_t o_words = o_off + (((n_clip + 1) * 8 + 15) & ~(size_t)15);
  auto grow = [&](size_t need) -> int {
    if (c->h_stage_cap >= need) return 0;
    if (c->h_stage) CK(cudaFreeHost(c->h_stage));
    c->h_stage = nullptr;
    c->h_stage_cap = 0;
    size_t want = need + need / 4 + (16u << 20);
    CK(cudaMallocHost(&c->h_stage, want));
    c->h_stage_cap = want;
    return 0;
  };
  { int r = grow(o_words + 64); if (r) return r; }
  CK(cudaMemcpyAsync((char*)c->h_stage, c->b_clip_n.p, n_clip * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync((char*)c->h_stage + o_off, c->b_clip_off.p, (n_clip + 1) * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  c->st_d2h_run += (i64)n_clip * 4 + (i64)(n_clip + 1) * 8;
  const i64 total = ((const i64*)((char*)c->h_stage + o_off))[n_clip];
  if (c->h_stage_cap < o_words + (size_t)total * 8 + 64) {  // grow, keeping the two tables
    std::vector<char> keep((char*)c->h_stage, (char*)c->h_stage + o_words);
    int r = grow(o_words + (size_t)total * 8 + 64);
    if (r) return r;
    memcpy(c->h_stage, keep.data(), o_words);
  }
  const int* h_n = (const int*)c->h_stage;
  const i64* h_off = (const i64*)((char*)c->h_stage + o_off);
  u32* h_wa = (u32*)((char*)c->h_stage + o_words);
  u32* h_wt = h_wa + total;
  const u32* pa = c->h_seq_a;
  const u32* pt = c->h_seq_t;
  const int* rl = c->h_read_len.data();
  const u8* rs = c->h_read_strand.data();
  const i64* so = c->h_read_seq_off.data();
#pragma omp parallel for schedule(static, 2048)
  for (long long k = 0; k < (long long)n_clip; ++k) {
    const int n = h_n[k];
    if (n < 20) continue;
    const int i = (int)(k >> 1);
    const ClipGeo g = clip_geometry(rl[i], n, (k & 1) == 0, rs[i] != 0);
    const i64 src = so[i] + g.w_first;
    memcpy(h_wa + h_off[k], pa + src, (size_t)g.n_words * 4);
    memcpy(h_wt + h_off[k], pt + src, (size_t)g.n_words * 4);
  }
  int r;
  if ((r = ensure(c, c->b_seq_a, (size_t)total * 4))) return r;
  if ((r = ensure(c, c->b_seq_t, (size_t)total * 4))) return r;
  if (total > 0) {
    CK(cudaMemcpyAsync(c->b_seq_a.p, h_wa, (size_t)total * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->b_seq_t.p, h_wt, (size_t)total * 4, cudaMemcpyHostToDevice, st));
  }
  c->st_h2d_run += total * 8;
  c->st_clip_words = total;
  return 0;
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
  memset(&c->hb, 0, sizeof c->hb);
  memset(&c->sizes, 0, sizeof c->sizes);
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMallocHost((void**)&c->h_pin, 64 * 8) != cudaSuccess) {
    int r = fail(nullptr, FRS_ERR_CUDA, "frs_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return r;
  }
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  for (int i = 0; i < FRS_SIDE_STREAMS; ++i) {
    // the chain "large DP classes -> solver of the split subproblems" is the critical path of the stage
    cudaStreamCreateWithPriority(&c->side[i], cudaStreamNonBlocking, i < 4 ? prio_hi : prio_lo);
    cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming);
  }
  cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming);
  cudaFuncSetAttribute(k_dp_warp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPW_WARPS * sizeof(DpWarpSmem<8>)));
  cudaFuncSetAttribute(k_dp_warp<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(DPW_WARPS * sizeof(DpWarpSmem<16>)));
  cudaFuncSetAttribute(k_signal, cudaFuncAttributeMaxDynamicSharedMemorySize, SIG_BINS * 4);
  cudaFuncSetAttribute(k_dp<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
  cudaFuncSetAttribute(k_dp<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
  cudaFuncSetAttribute(k_dp<DP_BIG_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 256);
  cudaFuncSetAttribute(k_dp_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  *out = c;
  return 0;
}

void frs_destroy(frs_context* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (DBuf* b : c->all)
    if (b->p) cudaFree(b->p);
  for (int i = 0; i < c->n_stages; ++i) {
    cudaEventDestroy(c->stages[i].ev0);
    cudaEventDestroy(c->stages[i].ev1);
  }
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  for (int i = 0; i < FRS_SIDE_STREAMS; ++i) {
    if (c->side[i]) cudaStreamDestroy(c->side[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  cudaStreamDestroy(c->stream);
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
  const long long v[FRS_N_STATS] = {c->st_h2d_upload, c->st_h2d_run, c->st_d2h_run, c->st_clip_words,
                                    (long long)c->hb.n_seq_words, c->st_poly_tasks, c->st_poly_long};
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

#define H2D(buf, src, bytes)                                                                      \
  do {                                                                                            \
    ENS(buf, bytes);                                                                              \
    if ((bytes) > 0) CK(cudaMemcpyAsync(c->buf.p, src, (size_t)(bytes), cudaMemcpyHostToDevice, c->stream)); \
    c->st_h2d_upload += (i64)(bytes);                                                             \
  } while (0)

int frs_upload(frs_context* c, const frs_batch* b) {
  if (!c || !b) return fail(c, FRS_ERR_ARG, "frs_upload: NULL argument");
  CK(cudaSetDevice(c->device));
  if (b->n_tints <= 0) return fail(c, FRS_ERR_ARG, "frs_upload: empty batch");
  if (b->n_samples <= 0 || b->n_islands <= 0) return fail(c, FRS_ERR_ARG, "frs_upload: batch without islands");
  const int T = b->n_tints, NI = b->n_islands, NR = b->n_reps, N = b->n_reads;
  // ---- validate the offset tables (the reference asserts the same facts while parsing) ----
  if (b->tint_island_off[0] != 0 || b->tint_island_off[T] != NI || b->tint_rep_off[0] != 0 ||
      b->tint_rep_off[T] != NR || b->tint_read_off[0] != 0 || b->tint_read_off[T] != N ||
      b->island_sample_off[0] != 0 || b->island_sample_off[NI] != b->n_samples || b->rep_iv_off[0] != 0 ||
      b->rep_iv_off[NR] != b->n_rep_ivs || b->read_iv_off[0] != 0 || b->read_iv_off[N] != b->n_read_ivs ||
      b->riv_cig_off[0] != 0 || b->riv_cig_off[b->n_read_ivs] != b->n_cigar_ops || b->read_seq_off[0] != 0 ||
      b->read_seq_off[N] != b->n_seq_words)
    return fail(c, FRS_ERR_ARG, "frs_upload: inconsistent offset tables");
  for (int t = 0; t < T; ++t)
    if (b->tint_island_off[t + 1] <= b->tint_island_off[t] || b->tint_rep_off[t + 1] <= b->tint_rep_off[t] ||
        b->tint_read_off[t + 1] < b->tint_read_off[t])
      return fail(c, FRS_ERR_ARG, "frs_upload: tint %d has no islands or no read reps", t);
  for (int i = 0; i < NI; ++i)
    if (b->island_sample_off[i + 1] - b->island_sample_off[i] < 2)
      return fail(c, FRS_ERR_ARG, "AssertionError: island %d is empty (freddie_segment.py:140)", i);
  c->hb = *b;
  c->h_tint_island_off.assign(b->tint_island_off, b->tint_island_off + T + 1);
  c->h_tint_rep_off.assign(b->tint_rep_off, b->tint_rep_off + T + 1);
  c->h_tint_read_off.assign(b->tint_read_off, b->tint_read_off + T + 1);
  c->h_island_sample_off.assign(b->island_sample_off, b->island_sample_off + NI + 1);
  // ---- derived host tables ----
  std::vector<SigWork> sig;
  std::vector<std::pair<int, int>> direct_runs;  // rep ranges of the sparse tints (k_signal direct mode)
  std::vector<TileWork> tiles;
  std::vector<RepTile> cov_tiles, dig_tiles;
  const int DIG_REPS = DIG_THREADS;
  for (int t = 0; t < T; ++t) {
    for (int i = b->tint_island_off[t]; i < b->tint_island_off[t + 1]; ++i) {
      int n = b->island_sample_off[i + 1] - b->island_sample_off[i];
      for (int lo = 0; lo < n; lo += TILE_SAMPLES) tiles.push_back(TileWork{i, lo, b->island_sample_off[i], n});
    }
    int r0 = b->tint_rep_off[t], r1 = b->tint_rep_off[t + 1];
    for (int r = b->tint_read_off[t]; r < b->tint_read_off[t + 1]; ++r) {
      const int rep = b->read_rep[r];
      if (rep < r0 || rep >= r1) return fail(c, FRS_ERR_ARG, "frs_upload: read %d points at rep %d of another tint", r, rep);
      if (b->read_iv_off[r + 1] - b->read_iv_off[r] != b->rep_iv_off[rep + 1] - b->rep_iv_off[rep])
        return fail(c, FRS_ERR_ARG, "frs_upload: read %d and its rep %d differ in their number of intervals", r, rep);
    }
    int s0 = b->island_sample_off[b->tint_island_off[t]], s1 = b->island_sample_off[b->tint_island_off[t + 1]];
    int single = (r1 - r0) <= SIG_REPS;
    const i64 n_endpoints = 2 * (i64)(b->rep_iv_off[r1] - b->rep_iv_off[r0]);
    if (n_endpoints < 8 * (i64)(s1 - s0)) {  // sparse tint: endpoints go straight to the global signal
      // flat samples and reps need no tint: runs of consecutive sparse tints share full CTAs
      if (!direct_runs.empty() && direct_runs.back().second == r0) direct_runs.back().second = r1;
      else direct_runs.push_back(std::make_pair(r0, r1));
    } else {
      for (int w = s0; w < s1; w += SIG_BINS)
        for (int r = r0; r < r1; r += SIG_REPS)
          sig.push_back(SigWork{t, w, w + SIG_BINS < s1 ? w + SIG_BINS : s1, r, r + SIG_REPS < r1 ? r + SIG_REPS : r1, single});
    }
    int R = r1 - r0, Rp = (R + 3) & ~3;
    for (int r = 0; r < Rp; r += COV_THREADS) cov_tiles.push_back(RepTile{t, r});
    for (int r = 0; r < R; r += DIG_REPS) dig_tiles.push_back(RepTile{t, r});
  }
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
  c->n_sig_work = (int)sig.size();  // histogram items first, then the direct ones
  for (const auto& run : direct_runs)
    for (int r = run.first; r < run.second; r += SIG_DIRECT_REPS)
      sig.push_back(SigWork{-1, 0, b->n_samples, r, r + SIG_DIRECT_REPS < run.second ? r + SIG_DIRECT_REPS : run.second, 2});
  c->n_sig_direct = (int)sig.size() - c->n_sig_work;
  c->n_tiles = (int)tiles.size();
  c->n_cov_tiles = (int)cov_tiles.size();
  c->n_dig_tiles = (int)dig_tiles.size();
  // ---- copies ----
  c->st_h2d_upload = 0;
  H2D(b_tint_island_off, b->tint_island_off, (size_t)(T + 1) * 4);
  H2D(b_tint_rep_off, b->tint_rep_off, (size_t)(T + 1) * 4);
  H2D(b_tint_read_off, b->tint_read_off, (size_t)(T + 1) * 4);
  H2D(b_island_start, b->island_start, (size_t)NI * 4);
  H2D(b_island_sample_off, b->island_sample_off, (size_t)(NI + 1) * 4);
  H2D(b_rep_iv_off, b->rep_iv_off, (size_t)(NR + 1) * 4);
  H2D(b_rep_weight, b->rep_weight, (size_t)NR * 4);
  H2D(b_rep_fs, b->rep_iv_fs, (size_t)b->n_rep_ivs * 4);
  H2D(b_rep_fe, b->rep_iv_fe, (size_t)b->n_rep_ivs * 4);
  H2D(b_read_rep, b->read_rep, (size_t)N * 4);
  H2D(b_read_strand, b->read_strand, (size_t)N);
  H2D(b_read_len, b->read_len, (size_t)N * 4);
  H2D(b_read_iv_off, b->read_iv_off, (size_t)(N + 1) * 4);
  H2D(b_read_seq_off, b->read_seq_off, (size_t)(N + 1) * 8);
  const bool derive_riv = !b->riv_ts || !b->riv_te;  // NULL: derived on the device from the rep intervals
  if (derive_riv) {
    ENS(b_riv_ts, (size_t)b->n_read_ivs * 4);
    ENS(b_riv_te, (size_t)b->n_read_ivs * 4);
  } else {
    H2D(b_riv_ts, b->riv_ts, (size_t)b->n_read_ivs * 4);
    H2D(b_riv_te, b->riv_te, (size_t)b->n_read_ivs * 4);
  }
  H2D(b_riv_qs, b->riv_qs, (size_t)b->n_read_ivs * 4);
  H2D(b_riv_qe, b->riv_qe, (size_t)b->n_read_ivs * 4);
  H2D(b_riv_cig_off, b->riv_cig_off, (size_t)(b->n_read_ivs + 1) * 4);
  H2D(b_cigar, b->cigar, (size_t)b->n_cigar_ops * 4);
  c->seq_resident = !c->opt_lazy_seq;
  if (c->seq_resident) {
    H2D(b_seq_a, b->seq_is_a, (size_t)b->n_seq_words * 4);
    H2D(b_seq_t, b->seq_is_t, (size_t)b->n_seq_words * 4);
  } else {
    // the poly-A/T scans only look at the clips of a read, and those are known after segmentation:
    // frs_run fetches just the clip words from the caller's planes (k_gap_prep -> host gather -> H2D)
    c->h_seq_a = b->seq_is_a;
    c->h_seq_t = b->seq_is_t;
    c->h_read_len.assign(b->read_len, b->read_len + N);
    c->h_read_strand.assign(b->read_strand, b->read_strand + N);
    c->h_read_seq_off.assign(b->read_seq_off, b->read_seq_off + N + 1);
  }
  // the derived tables live in pageable vectors: stage synchronously before they go out of scope
  // owner tables (tint of every island / rep / read) are derived on the device
  ENS(b_island_tint, (size_t)NI * 4);
  ENS(b_rep_tint, (size_t)NR * 4);
  ENS(b_read_tint, (size_t)N * 4);
  k_owner_tables<<<cdiv((i64)NI + NR + N, 256), 256, 0, c->stream>>>(
      T, NI, NR, N, c->b_tint_island_off.as<int>(), c->b_tint_rep_off.as<int>(), c->b_tint_read_off.as<int>(),
      c->b_island_tint.as<int>(), c->b_rep_tint.as<int>(), c->b_read_tint.as<int>());
  if (derive_riv && N > 0)
    k_derive_riv<<<cdiv(N, 256), 256, 0, c->stream>>>(N, NI, c->b_read_rep.as<int>(), c->b_read_iv_off.as<int>(),
                                                      c->b_rep_iv_off.as<int>(), c->b_rep_fs.as<int>(), c->b_rep_fe.as<int>(),
                                                      c->b_island_sample_off.as<int>(), c->b_island_start.as<int>(),
                                                      c->b_riv_ts.as<int>(), c->b_riv_te.as<int>());
  H2D(b_tint_order, tint_order.data(), (size_t)T * 4);
  H2D(b_sig_work, sig.data(), sig.size() * sizeof(SigWork));
  H2D(b_tiles, tiles.data(), tiles.size() * sizeof(TileWork));
  H2D(b_cov_tiles, cov_tiles.data(), cov_tiles.size() * sizeof(RepTile));
  H2D(b_dig_tiles, dig_tiles.data(), dig_tiles.size() * sizeof(RepTile));
  CK(cudaStreamSynchronize(c->stream));
  c->uploaded = true;
  c->ran = false;
  return 0;
}

int frs_run(frs_context* c, const frs_params* prm, frs_result_sizes* sizes_out) {
  if (!c || !prm) return fail(c, FRS_ERR_ARG, "frs_run: NULL argument");
  if (!c->uploaded) return fail(c, FRS_ERR_STATE, "frs_run: no batch uploaded");
  CK(cudaSetDevice(c->device));
  // parse_args asserts (freddie_segment.py:104-109)
  if (!(prm->tp >= 0.5 && prm->tp <= 1.0)) return fail(c, FRS_ERR_ARG, "AssertionError: 1 >= threshold_rate >= 0.5");
  if (!(prm->vf > 0 && prm->vf < 10)) return fail(c, FRS_ERR_ARG, "AssertionError: 10 > variance_factor > 0");
  if (!(prm->sigma > 0 && prm->sigma <= 50)) return fail(c, FRS_ERR_ARG, "AssertionError: 50 >= sigma > 0");
  if (!(prm->mps > 3)) return fail(c, FRS_ERR_ARG, "AssertionError: max_problem_size > 3");
  if (prm->mps < 11)
    return fail(c, FRS_ERR_LIMIT, "max_problem_size < 11 is not supported: the reference's +-5 anchor window "
                                  "(freddie_segment.py:639) indexes outside the problem there (IndexError / wrap)");
  if (!(prm->lo >= 0)) return fail(c, FRS_ERR_ARG, "AssertionError: min_read_support_outside >= 0");
  if (prm->gauss_radius != (int)(4.0 * prm->sigma + 0.5) || prm->refine_radius != (int)(1.0 * prm->sigma + 0.5))
    return fail(c, FRS_ERR_ARG, "frs_run: kernel radii do not match sigma");
  if (prm->gauss_radius > 1000) return fail(c, FRS_ERR_LIMIT, "gauss radius too large");

  const frs_batch& B = c->hb;
  const int T = B.n_tints, NI = B.n_islands, NR = B.n_reps, N = B.n_reads;
  const i64 L = B.n_samples;
  cudaStream_t st = c->stream;
  c->launch_count = 0;
  for (int i = 0; i < c->n_stages; ++i) { c->stages[i].used = false; c->stages[i].launches = 0; }

  // parameter tables
  const int lw = prm->gauss_radius, rr = prm->refine_radius;
  size_t ptab = (size_t)prm->thr_table_len + (2 * lw + 1) + (2 * rr + 1);
  ENS(b_params, ptab * 8);
  double* d_tbl = c->b_params.as<double>();
  double* d_gw = d_tbl + prm->thr_table_len;
  double* d_rw = d_gw + (2 * lw + 1);
  CK(cudaMemcpyAsync(d_tbl, prm->thr_table, (size_t)prm->thr_table_len * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_gw, prm->gauss_w, (size_t)(2 * lw + 1) * 8, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_rw, prm->refine_w, (size_t)(2 * rr + 1) * 8, cudaMemcpyHostToDevice, st));

  ENS(b_counters, 64 * 8);
  ENS(b_stats, 8 * 8);
  ENS(b_err, 16);
  CK(cudaMemsetAsync(c->b_counters.p, 0, 64 * 8, st));
  CK(cudaMemsetAsync(c->b_stats.p, 0, 8 * 8, st));
  CK(cudaMemsetAsync(c->b_err.p, 0, 16, st));
  int* d_err = c->b_err.as<int>();

  const int* d_tint_island_off = c->b_tint_island_off.as<int>();
  const int* d_tint_rep_off = c->b_tint_rep_off.as<int>();
  const int* d_island_sample_off = c->b_island_sample_off.as<int>();
  const int* d_island_tint = c->b_island_tint.as<int>();

  // ================= phase 1: signal -> smoothed signal -> candidates, threshold =================
  ENS(b_yraw, L * 4);
  CK(cudaMemsetAsync(c->b_yraw.p, 0, L * 4, st));
  stage_begin(c, "signal");
  if (c->n_sig_work > 0) {
    k_signal<<<c->n_sig_work, SIG_THREADS, SIG_BINS * 4, st>>>(c->b_sig_work.as<SigWork>(), c->b_rep_iv_off.as<int>(),
                                                               c->b_rep_weight.as<int>(), c->b_rep_fs.as<int>(),
                                                               c->b_rep_fe.as<int>(), prm->ignore_ends, c->b_yraw.as<int>());
    LAUNCHED();
  }
  if (c->n_sig_direct > 0) {  // no histogram: no shared memory, full occupancy
    k_signal<<<c->n_sig_direct, SIG_THREADS, 0, st>>>(c->b_sig_work.as<SigWork>() + c->n_sig_work, c->b_rep_iv_off.as<int>(),
                                                      c->b_rep_weight.as<int>(), c->b_rep_fs.as<int>(),
                                                      c->b_rep_fe.as<int>(), prm->ignore_ends, c->b_yraw.as<int>());
    LAUNCHED();
  }

  ENS(b_y, L * 8);
  stage_begin(c, "smooth");
  ENS(b_sflag, L);
  ENS(b_cand_flat, (L / 2 + 2 * NI + 16) * 4);  // peaks are >= 2 apart, plus both ends of every island
  ENS(b_vbuf, L * 8);
  ENS(b_tint_pos_off, (size_t)(T + 1) * 4);
  const size_t n_groups = (size_t)c->n_tiles / TILE_GROUP + 1;
  ENS(b_tile_state, n_groups * 8 + (size_t)c->n_tiles * (2 * TILE_WORDS * 4 + 4) + 64);
  {
    // group totals | per tile: candidate / positive ballot words, packed counts
    unsigned long long* d_gsum = c->b_tile_state.as<unsigned long long>();
    u32* d_cmask = (u32*)(d_gsum + n_groups);
    u32* d_pmask = d_cmask + (size_t)c->n_tiles * TILE_WORDS;
    u32* d_tcnt = d_pmask + (size_t)c->n_tiles * TILE_WORDS;
    CK(cudaMemsetAsync(d_gsum, 0, n_groups * 8, st));
    const size_t sm = (size_t)p1_smem_layout(lw).total;
    if (sm > 48 * 1024) CK(cudaFuncSetAttribute(k_smooth, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_smooth<<<c->n_tiles, GAUSS_THREADS, sm, st>>>(c->b_tiles.as<TileWork>(), d_island_sample_off, c->b_yraw.as<int>(),
                                                    d_gw, lw, c->b_y.as<double>(), d_cmask, d_pmask, d_tcnt, d_gsum);
    LAUNCHED();
    stage_begin(c, "lists");
    k_tile_lists<<<c->n_tiles, GAUSS_THREADS, 0, st>>>(c->b_tiles.as<TileWork>(), c->n_tiles, d_island_sample_off,
                                                       d_island_tint, d_tint_island_off, T, d_cmask, d_pmask, d_tcnt,
                                                       d_gsum, c->b_y.as<double>(), c->b_cand_flat.as<int>(),
                                                       c->b_vbuf.as<double>(), c->b_tint_pos_off.as<int>(),
                                                       c->b_counters.as<i64>());
    LAUNCHED();
  }

  stage_begin(c, "threshold");
  ENS(b_thr, (size_t)T * 8);
  {
    size_t nh = (size_t)L / 8 + 64 * (size_t)T + 64;  // heap scratch of the giant tints (see k_threshold)
    ENS(b_leaf_len, nh * 4);
    ENS(b_leaf_sum, nh * 8);
  }
  k_threshold<<<T, THR_THREADS, 0, st>>>(c->b_tint_order.as<int>(), d_tint_island_off, d_island_sample_off,
                                         c->b_tint_pos_off.as<int>(), prm->vf, c->b_vbuf.as<double>(),
                                         c->b_leaf_len.as<int>(), c->b_leaf_sum.as<double>(), c->b_thr.as<double>());
  LAUNCHED();
  stage_end(c);

  { int r = read_counters(c, 1); if (r) return r; }   // sync: number of candidates
  const i64 K = c->h_pin[0];
  c->n_cand = K;

  // ================= phase 2: fixed candidates, subproblems, coverage, DP =================
  stage_begin(c, "fixed");
  ENS(b_cand_island, K * 4);
  ENS(b_island_cand_off, (size_t)(NI + 1) * 4);
  k_cand_meta<<<cdiv(K + 1, 256), 256, 0, st>>>(c->b_cand_flat.as<int>(), (int)K, d_island_sample_off, NI,
                                                c->b_cand_island.as<int>(), c->b_island_cand_off.as<int>());
  LAUNCHED();
  ENS(b_fixed0, K);
  ENS(b_fixed1, K);
  k_fixed_a<<<cdiv(K, 256), 256, 0, st>>>((int)K, c->b_cand_flat.as<int>(), c->b_cand_island.as<int>(),
                                          c->b_island_cand_off.as<int>(), d_island_tint, c->b_y.as<double>(),
                                          c->b_thr.as<double>(), c->b_fixed0.as<u8>(), c->b_fixed1.as<u8>());
  LAUNCHED();
  k_fixed_b<<<cdiv(K, 256), 256, 0, st>>>((int)K, c->b_cand_flat.as<int>(), c->b_cand_island.as<int>(),
                                          c->b_island_cand_off.as<int>(), c->b_y.as<double>(), prm->mps,
                                          c->b_fixed0.as<u8>(), c->b_fixed1.as<u8>(), d_err);
  LAUNCHED();
  stage_begin(c, "subproblems");
  // every subproblem has an interior candidate of its own: at most K/2 of them
  const i64 NSUB_MAX = K / 2 + 1;
  ENS(b_sub_start, NSUB_MAX * 4);
  ENS(b_sub_n, NSUB_MAX * 4);
  ENS(b_sub_tint, NSUB_MAX * 4);
  ENS(b_sub_info, NSUB_MAX * 4);
  ENS(b_sub_slabs, NSUB_MAX * 4);
  ENS(b_sub_tab_off, NSUB_MAX * 8);
  ENS(b_plan, PLAN_SLOTS * 8);
  CK(cudaMemsetAsync(c->b_plan.p, 0, PLAN_SLOTS * 8, st));
  const int slab_words = c->opt_slab_words;
  const int keep = c->opt_keep_tables;
  k_sub_build<<<cdiv(K, 256), 256, 0, st>>>((int)K, c->b_fixed1.as<u8>(), c->b_cand_island.as<int>(),
                                            c->b_island_cand_off.as<int>(), d_island_tint, d_tint_rep_off, slab_words, keep,
                                            c->b_sub_start.as<int>(), c->b_sub_n.as<int>(), c->b_sub_tint.as<int>(),
                                            c->b_sub_info.as<int>(), c->b_sub_slabs.as<int>(), c->b_sub_tab_off.as<i64>(),
                                            c->b_plan.as<i64>());
  LAUNCHED();
  // coverage block offsets per tint (rows = candidates of the tint, stride Rp)
  ENS(b_tint_cand_off, (size_t)(T + 1) * 4);
  ENS(b_cov_sz, (size_t)(T + 1) * 8);
  ENS(b_tint_cov_off, (size_t)(T + 1) * 8);
  k_tint_cov_sizes<<<cdiv(T + 1, 256), 256, 0, st>>>(T, d_tint_island_off, c->b_island_cand_off.as<int>(), d_tint_rep_off,
                                                     c->b_tint_cand_off.as<int>(), c->b_cov_sz.as<i64>());
  LAUNCHED();
  { int r = scan_exclusive<i64, i64>(c, c->b_cov_sz.as<i64>(), T, c->b_tint_cov_off.as<i64>()); if (r) return r; }
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 3, c->b_tint_cov_off.as<i64>() + T, 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 16, c->b_plan.p, PLAN_SLOTS * 8, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 5, d_err, 8, cudaMemcpyDeviceToDevice, st));
  { int r = read_counters(c, 16 + PLAN_SLOTS); if (r) return r; }   // the ONE sync of this phase: plan, sizes, asserts
  {
    const int* he = (const int*)(c->h_pin + 5);
    if (he[0]) return fail(c, FRS_ERR_ASSERT, "AssertionError: %s [item %d]", deverr_text(he[0]), he[1]);
  }
  const i64 NSUB = c->h_pin[16 + PLAN_NSUB];
  const i64 COV = c->h_pin[3];
  c->n_sub = NSUB;
  c->cov_elems = COV;
  i64 tab_total = c->h_pin[16 + PLAN_TAB], n_split = c->h_pin[16 + PLAN_SPLIT], dp_cells = c->h_pin[16 + PLAN_CELLS],
      dp_read_cells = c->h_pin[16 + PLAN_RCELLS], cls_cnt[DP_CLASSES];
  int cls_maxn[DP_CLASSES];
  for (int k = 0; k < DP_CLASSES; ++k) {
    cls_cnt[k] = c->h_pin[16 + PLAN_WORK + k];
    cls_maxn[k] = (int)c->h_pin[16 + PLAN_MAXN + k];
  }
  const int max_n = (int)c->h_pin[16 + PLAN_MAXALL];
  c->tab_elems = tab_total;

  ENS(b_P, COV * 4);
  stage_begin(c, "coverage");
  if (NSUB > 0) {
    k_coverage<<<dim3((unsigned)c->n_cov_tiles, COV_CHUNKS), COV_THREADS, 0, st>>>(c->b_cov_tiles.as<RepTile>(), d_tint_rep_off,
                                                       c->b_tint_cand_off.as<int>(), c->b_tint_cov_off.as<i64>(),
                                                       c->b_rep_iv_off.as<int>(), c->b_rep_fs.as<int>(),
                                                       c->b_rep_fe.as<int>(), c->b_cand_flat.as<int>(), c->b_P.as<u32>());
    LAUNCHED();
  }

  ENS(b_dpfinal, K);
  CK(cudaMemcpyAsync(c->b_dpfinal.p, c->b_fixed1.p, K, cudaMemcpyDeviceToDevice, st));
  if (NSUB > 0) {
    stage_begin(c, "dp_plan");
    i64 n_work = 0;
    DpBases bases;
    for (int k = 0; k < DP_CLASSES; ++k) { bases.base[k] = (int)n_work; n_work += cls_cnt[k]; }
    if (n_work > 0x7fffffff) return fail(c, FRS_ERR_LIMIT, "too many DP work items in one batch (%lld)", (long long)n_work);
    ENS(b_work, n_work * 8);
    ENS(b_split_list, n_split * 4);
    ENS(b_cursor, 8 * 4);
    CK(cudaMemsetAsync(c->b_cursor.p, 0, 8 * 4, st));
    k_sub_fill<<<cdiv(NSUB, 256), 256, 0, st>>>((int)NSUB, c->b_sub_info.as<int>(), c->b_sub_slabs.as<int>(), bases,
                                                c->b_cursor.as<int>(), c->b_work.as<DpWork>(), c->b_split_list.as<int>());
    LAUNCHED();
    ENS(b_tab, tab_total * 4);
    if (tab_total > 0) CK(cudaMemsetAsync(c->b_tab.p, 0, tab_total * 4, st));
    stage_begin(c, "dp");
    DpArgs A;
    A.sub_start = c->b_sub_start.as<int>(); A.sub_n = c->b_sub_n.as<int>(); A.sub_tint = c->b_sub_tint.as<int>();
    A.sub_info = c->b_sub_info.as<int>(); A.sub_tab_off = c->b_sub_tab_off.as<i64>();
    A.tint_rep_off = d_tint_rep_off; A.tint_cand_off = c->b_tint_cand_off.as<int>();
    A.tint_cov_off = c->b_tint_cov_off.as<i64>(); A.rep_weight = c->b_rep_weight.as<int>();
    A.cand_flat = c->b_cand_flat.as<int>(); A.P = c->b_P.as<u32>();
    A.thr_table = d_tbl; A.thr_table_len = prm->thr_table_len; A.tp = prm->tp;
    A.lo = prm->lo; A.keep_tables = keep;
    A.tab = c->b_tab.as<int>(); A.final_flag = c->b_dpfinal.as<u8>(); A.err = d_err;
    const int SMEM_BUDGET = 226 * 1024;  // of the 227 KB a CTA can opt in to: room for 4-word chunks up to n = 55
    // the classes are independent: launch them on side streams so that the few long CTAs of the large
    // classes overlap the many short ones (fork / join with events on the context stream)
    CK(cudaEventRecord(c->ev_fork, st));
    bool used[FRS_SIDE_STREAMS] = {};
    for (int k = DP_CLASSES - 1; k >= 0; --k) {  // longest-running classes first
      if (cls_cnt[k] == 0) continue;
      const int sidx = k >= 2 ? k - 2 : 4 + k;  // a stream per class
      cudaStream_t ks = c->side[sidx];
      CK(cudaStreamWaitEvent(ks, c->ev_fork, 0));
      const DpWork* wl = c->b_work.as<DpWork>() + bases.base[k];
      if (k <= 1) {
        const unsigned g = (unsigned)((cls_cnt[k] + DPW_WARPS - 1) / DPW_WARPS);
        if (k == 0) k_dp_warp<8><<<g, DPW_WARPS * 32, DPW_WARPS * sizeof(DpWarpSmem<8>), ks>>>(A, wl, (int)cls_cnt[k]);
        else k_dp_warp<16><<<g, DPW_WARPS * 32, DPW_WARPS * sizeof(DpWarpSmem<16>), ks>>>(A, wl, (int)cls_cnt[k]);
      } else {
        const int M = cls_maxn[k], on_chip = k < 5 ? 1 : 0;
        int wc = (k <= 3) ? 4 : DPT_MAXW;
        while (wc > 1 && dp_smem_layout(M, wc, on_chip).total > SMEM_BUDGET) wc >>= 1;
        const size_t sm = (size_t)dp_smem_layout(M, wc, on_chip).total;
        if (sm > 227 * 1024 - 256)
          return fail(c, FRS_ERR_LIMIT, "subproblem with %d candidates exceeds the DP kernel's shared-memory budget "
                                        "(max_problem_size too large for this build)", M);
        const unsigned g = (unsigned)cls_cnt[k];
        switch (k) {
          case 2: k_dp<128><<<g, 128, sm, ks>>>(A, wl, M, wc, on_chip); break;
          case 3: k_dp<256><<<g, 256, sm, ks>>>(A, wl, M, wc, on_chip); break;
          default: k_dp<DP_BIG_THREADS><<<g, DP_BIG_THREADS, sm, ks>>>(A, wl, M, wc, on_chip); break;
        }
      }
      LAUNCHED();
      CK(cudaEventRecord(c->ev_join[sidx], ks));
      used[sidx] = true;
    }
    if (n_split > 0) {
      // only CTA classes (streams 0..3) have split subproblems: their solver starts as soon as those are
      // done and runs beside the warp classes
      cudaStream_t ss = c->side[0];
      if (!used[0]) CK(cudaStreamWaitEvent(ss, c->ev_fork, 0));
      for (int k = 1; k < 4; ++k)
        if (used[k]) CK(cudaStreamWaitEvent(ss, c->ev_join[k], 0));
      const int stage_n = max_n < DP_SMEM_MAX_N ? max_n : DP_SMEM_MAX_N;
      size_t sm2 = dps_smem_bytes(max_n, stage_n);
      if (sm2 > 220 * 1024) return fail(c, FRS_ERR_LIMIT, "subproblem with %d candidates exceeds the DP solver's budget", max_n);
      k_dp_solve<<<(unsigned)n_split, DPS_THREADS, sm2, ss>>>(A, c->b_split_list.as<int>(), max_n, stage_n);
      LAUNCHED();
      CK(cudaEventRecord(c->ev_join[0], ss));
      used[0] = true;
    }
    for (int k = 0; k < FRS_SIDE_STREAMS; ++k)
      if (used[k]) CK(cudaStreamWaitEvent(st, c->ev_join[k], 0));
  }

  // ================= phase 3: refine, final positions, digits =================
  stage_begin(c, "refine");
  ENS(b_ref_list, K * 8 + 16);
  CK(cudaMemsetAsync(c->b_sflag.p, 0, L, st));
  CK(cudaMemsetAsync(c->b_counters.as<i64>() + 10, 0, 8, st));
  int* d_ref_cnt = (int*)(c->b_counters.as<i64>() + 10);
  k_final_mark<<<cdiv(K, 256), 256, 0, st>>>((int)K, c->b_dpfinal.as<u8>(), c->b_cand_flat.as<int>(),
                                             c->b_cand_island.as<int>(), c->b_island_cand_off.as<int>(),
                                             c->b_sflag.as<u8>(), c->b_ref_list.as<int2>(), d_ref_cnt);
  LAUNCHED();
  ENS(b_ref_list2, K * 8 + 16);
  CK(cudaMemsetAsync(c->b_counters.as<i64>() + 9, 0, 8, st));
  int* d_ref_cnt2 = (int*)(c->b_counters.as<i64>() + 9);
  k_refine_filter<<<148 * 8, 256, 0, st>>>(c->b_ref_list.as<int2>(), d_ref_cnt, c->b_yraw.as<int>(),
                                           c->b_ref_list2.as<int2>(), d_ref_cnt2);
  LAUNCHED();
  ENS(b_gbuf, L * 8);
  ENS(b_pstate, L);
  k_refine<<<148 * 8, REF_THREADS, 0, st>>>(c->b_ref_list2.as<int2>(), d_ref_cnt2, c->b_yraw.as<int>(), d_rw, rr, prm->sigma,
                                            c->b_gbuf.as<double>(), c->b_pstate.as<u8>(), c->b_sflag.as<u8>());
  LAUNCHED();

  stage_begin(c, "finals");
  // refine adds peaks at least 20 samples apart inside segments longer than 40
  const i64 NFIN_MAX = K + L / 20 + 16;
  ENS(b_final_flat, (L / 2 + 2 * NI + 16) * 4);
  { int r = compact_flags(c, c->b_sflag.as<u8>(), L, c->b_final_flat.as<int>(), 11); if (r) return r; }
  const i64* d_nfin = c->b_counters.as<i64>() + 11;
  ENS(b_final_pos, NFIN_MAX * 4);
  ENS(b_final_island, NFIN_MAX * 4);
  ENS(b_tint_final_off, (size_t)(T + 1) * 4);
  k_final_meta<<<148 * 4, 256, 0, st>>>(d_nfin, c->b_final_flat.as<int>(), d_island_sample_off,
                                        c->b_island_start.as<int>(), d_island_tint, d_tint_island_off, NI, T,
                                        c->b_final_pos.as<int>(), c->b_final_island.as<int>(),
                                        c->b_tint_final_off.as<int>());
  LAUNCHED();
  ENS(b_dig_sz, (size_t)(T + 1) * 8);
  ENS(b_tint_digit_off, (size_t)(T + 1) * 8);
  k_digit_sizes<<<cdiv(T, 256), 256, 0, st>>>(T, d_tint_rep_off, c->b_tint_final_off.as<int>(), c->b_dig_sz.as<i64>());
  LAUNCHED();
  { int r = scan_exclusive<i64, i64>(c, c->b_dig_sz.as<i64>(), T, c->b_tint_digit_off.as<i64>()); if (r) return r; }
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 12, c->b_tint_digit_off.as<i64>() + T, 8, cudaMemcpyDeviceToDevice, st));
  ENS(b_seg_ty, NFIN_MAX * 4);
  ENS(b_seg_tn, NFIN_MAX * 4);
  k_seg_cuts<<<148 * 4, 256, 0, st>>>(d_nfin, c->b_final_flat.as<int>(), c->b_final_island.as<int>(), d_tbl,
                                      prm->thr_table_len, prm->tp, c->b_seg_ty.as<int>(), c->b_seg_tn.as<int>());
  LAUNCHED();
  { int r = read_counters(c, 13); if (r) return r; }  // sync: final positions, digit bytes
  const i64 NFIN = c->h_pin[11];
  const i64 NDIG = c->h_pin[12];
  if (NFIN > NFIN_MAX) return fail(c, FRS_ERR_LIMIT, "internal: more final positions than the refine bound allows");

  ENS(b_digits, NDIG);
  stage_begin(c, "digits");
  ENS(b_run_cnt, (size_t)NR * 4);
  ENS(b_run_off, (size_t)(NR + 1) * 4);
  CK(cudaMemsetAsync(c->b_run_cnt.p, 0, (size_t)NR * 4, st));
  k_digits<<<dim3((unsigned)c->n_dig_tiles, DIG_CHUNKS), DIG_THREADS, 0, st>>>(c->b_dig_tiles.as<RepTile>(), d_tint_rep_off,
                                                   c->b_tint_final_off.as<int>(), c->b_tint_digit_off.as<i64>(),
                                                   c->b_rep_iv_off.as<int>(), c->b_rep_fs.as<int>(), c->b_rep_fe.as<int>(),
                                                   c->b_final_flat.as<int>(), c->b_seg_ty.as<int>(), c->b_seg_tn.as<int>(),
                                                   c->b_digits.as<u8>(), c->b_run_cnt.as<int>(), d_err);
  LAUNCHED();

  stage_begin(c, "runs");
  { int r = scan_exclusive<int, int>(c, c->b_run_cnt.as<int>(), NR, c->b_run_off.as<int>()); if (r) return r; }
  CK(cudaMemsetAsync(c->b_counters.as<i64>() + 13, 0, 16, st));
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 13, c->b_run_off.as<int>() + NR, 4, cudaMemcpyDeviceToDevice, st));
  ENS(b_gap_cnt, (size_t)N * 4);
  ENS(b_read_gap_off, (size_t)(N + 1) * 4);
  k_gap_count<<<cdiv(N, 256), 256, 0, st>>>(N, c->b_read_rep.as<int>(), c->b_run_off.as<int>(), c->b_gap_cnt.as<int>());
  LAUNCHED();
  { int r = scan_exclusive<int, int>(c, c->b_gap_cnt.as<int>(), N, c->b_read_gap_off.as<int>()); if (r) return r; }
  CK(cudaMemcpyAsync(c->b_counters.as<i64>() + 14, c->b_read_gap_off.as<int>() + N, 4, cudaMemcpyDeviceToDevice, st));
  { int r = read_counters(c, 15); if (r) return r; }  // sync: number of runs / gap records
  const i64 NRUN = c->h_pin[13];
  const i64 NGAP = c->h_pin[14];
  ENS(b_runs, NRUN * 8);
  k_run_fill<<<cdiv((i64)NR * 32, 256), 256, 0, st>>>(NR, c->b_rep_tint.as<int>(), d_tint_rep_off,
                                                      c->b_tint_final_off.as<int>(), c->b_tint_digit_off.as<i64>(),
                                                      c->b_digits.as<u8>(), c->b_run_off.as<int>(), c->b_runs.as<int2>());
  LAUNCHED();

  ENS(b_read_head, (size_t)N * 32);
  ENS(b_gap_rec, NGAP * 12);
  stage_begin(c, "gaps");
  {
    GapArgs G;
    G.n_reads = N; G.read_rep = c->b_read_rep.as<int>(); G.read_strand = c->b_read_strand.as<u8>();
    G.read_len = c->b_read_len.as<int>(); G.read_iv_off = c->b_read_iv_off.as<int>();
    G.read_seq_off = c->b_read_seq_off.as<i64>(); G.read_tint = c->b_read_tint.as<int>();
    G.riv_ts = c->b_riv_ts.as<int>(); G.riv_te = c->b_riv_te.as<int>(); G.riv_qs = c->b_riv_qs.as<int>();
    G.riv_qe = c->b_riv_qe.as<int>(); G.riv_cig_off = c->b_riv_cig_off.as<int>(); G.cigar = c->b_cigar.as<u32>();
    G.seq_a = c->b_seq_a.as<u32>(); G.seq_t = c->b_seq_t.as<u32>(); G.run_off = c->b_run_off.as<int>();
    G.runs = c->b_runs.as<int2>(); G.tint_final_off = c->b_tint_final_off.as<int>();
    G.final_pos = c->b_final_pos.as<int>(); G.read_gap_off = c->b_read_gap_off.as<int>();
    G.read_head = c->b_read_head.as<int>(); G.gap_rec = c->b_gap_rec.as<int>(); G.err = d_err;
    ENS(b_clip_n, (size_t)N * 8);
    ENS(b_clip_words, (size_t)N * 8);
    ENS(b_clip_off, (size_t)N * 16 + 8);
    ENS(b_task_order, (size_t)N * 16);
    ENS(b_task_res, (size_t)N * 4 * sizeof(PolyRes));
    ENS(b_poly_cls, (2 * POLY_CLASSES + 1) * 4);
    G.clip_n = c->b_clip_n.as<int>(); G.clip_words = c->b_clip_words.as<int>(); G.clip_off = c->b_clip_off.as<i64>(); G.seq_resident = c->seq_resident ? 1 : 0;
    G.cls_count = c->b_poly_cls.as<int>();
    G.task_order = c->b_task_order.as<int>(); G.task_res = c->b_task_res.as<PolyRes>();
    G.long_class = c->opt_poly_long_class;
    c->st_h2d_run = c->st_d2h_run = c->st_clip_words = 0;
    if (N > 0) {
      CK(cudaMemsetAsync(c->b_poly_cls.p, 0, (2 * POLY_CLASSES + 1) * 4, st));
      k_gap_prep<<<cdiv(N, 128), 128, 0, st>>>(G); LAUNCHED();
      if (NGAP > 0) { k_gap_sizes<<<cdiv(NGAP, 128), 128, 0, st>>>(G, (int)NGAP); LAUNCHED(); }
      if (!c->seq_resident) {
        stage_begin(c, "clip_fetch");
        int r = fetch_clip_words(c, N);
        if (r) return r;
        G.seq_a = c->b_seq_a.as<u32>();
        G.seq_t = c->b_seq_t.as<u32>();
      }
      stage_begin(c, "poly");
      ENS(b_poly_flag, (size_t)N * 4);
      k_poly_filter<<<cdiv((i64)N * 4, 128), 128, 0, st>>>(G, c->b_poly_flag.as<u8>()); LAUNCHED();
      k_poly_bases<<<1, 32, 0, st>>>(G.cls_count, G.long_class, d_err + 2); LAUNCHED();
      k_poly_scatter<<<cdiv((i64)N * 4, 256), 256, 0, st>>>(N * 4, G.clip_n, c->b_poly_flag.as<u8>(), G.cls_count,
                                                             G.task_order); LAUNCHED();
      // long clips (one warp each) run beside the short ones (one thread each)
      CK(cudaEventRecord(c->ev_fork, st));
      CK(cudaStreamWaitEvent(c->side[0], c->ev_fork, 0));
      k_poly_long<<<148 * 4, 128, 0, c->side[0]>>>(G); LAUNCHED();
      CK(cudaEventRecord(c->ev_join[0], c->side[0]));
      k_poly_scan<<<cdiv((i64)N * 4, 128), 128, 0, st>>>(G); LAUNCHED();
      CK(cudaStreamWaitEvent(st, c->ev_join[0], 0));
      k_gap_finish<<<cdiv(N, 128), 128, 0, st>>>(G); LAUNCHED();
    }
  }
  stage_end(c);
  { int r = check_dev_err(c); if (r) return r; }
  CK(cudaGetLastError());

  c->sizes.n_final = NFIN;
  c->sizes.n_digit_bytes = NDIG;
  c->sizes.n_gap_records = NGAP;
  c->sizes.n_candidates = K;
  c->sizes.n_subproblems = NSUB;
  c->sizes.dp_cells = dp_cells;
  c->sizes.dp_read_cells = dp_read_cells;
  c->sizes.max_subproblem = max_n;
  c->sizes.pad = 0;
  if (sizes_out) *sizes_out = c->sizes;
  c->ran = true;
  return 0;
}

int frs_segment_batch(frs_context* c, const frs_batch* batch, const frs_params* prm, frs_result_sizes* sizes) {
  int r = frs_upload(c, batch);
  if (r) return r;
  return frs_run(c, prm, sizes);
}

#define D2H(dst, buf, bytes)                                                                              \
  do {                                                                                                    \
    if ((dst) && (bytes) > 0) CK(cudaMemcpyAsync(dst, c->buf.p, (size_t)(bytes), cudaMemcpyDeviceToHost, c->stream)); \
  } while (0)

int frs_download(frs_context* c, const frs_result* o) {
  if (!c || !o) return fail(c, FRS_ERR_ARG, "frs_download: NULL argument");
  if (!c->ran) return fail(c, FRS_ERR_STATE, "frs_download: no results (call frs_run first)");
  CK(cudaSetDevice(c->device));
  const int T = c->hb.n_tints, N = c->hb.n_reads;
  D2H(o->tint_final_off, b_tint_final_off, (size_t)(T + 1) * 4);
  D2H(o->final_pos, b_final_pos, c->sizes.n_final * 4);
  D2H(o->tint_digit_off, b_tint_digit_off, (size_t)(T + 1) * 8);
  D2H(o->digits, b_digits, c->sizes.n_digit_bytes);
  D2H(o->read_head, b_read_head, (size_t)N * 32);
  D2H(o->read_gap_off, b_read_gap_off, (size_t)(N + 1) * 4);
  D2H(o->gap_rec, b_gap_rec, c->sizes.n_gap_records * 12);
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

int frs_get_intermediate(frs_context* c, int which, void* dst, size_t cap, size_t* bytes) {
  if (!c || !bytes) return fail(c, FRS_ERR_ARG, "frs_get_intermediate: NULL argument");
  if (!c->ran) return fail(c, FRS_ERR_STATE, "frs_get_intermediate: nothing has run");
  CK(cudaSetDevice(c->device));
  const void* src = nullptr;
  size_t sz = 0;
  const i64 L = c->hb.n_samples, K = c->n_cand, NS = c->n_sub;
  switch (which) {
    case FRS_TAP_Y_RAW: src = c->b_yraw.p; sz = L * 4; break;
    case FRS_TAP_Y: src = c->b_y.p; sz = L * 8; break;
    case FRS_TAP_THR: src = c->b_thr.p; sz = (size_t)c->hb.n_tints * 8; break;
    case FRS_TAP_CAND: src = c->b_cand_flat.p; sz = K * 4; break;
    case FRS_TAP_FIXED: src = c->b_fixed1.p; sz = K; break;
    case FRS_TAP_DP_FINAL: src = c->b_dpfinal.p; sz = K; break;
    case FRS_TAP_SUB_START: src = c->b_sub_start.p; sz = NS * 4; break;
    case FRS_TAP_SUB_N: src = c->b_sub_n.p; sz = NS * 4; break;
    case FRS_TAP_COVERAGE: src = c->b_P.p; sz = c->cov_elems * 4; break;
    case FRS_TAP_DP_TABLES: src = c->b_tab.p; sz = c->tab_elems * 4; break;
    case FRS_TAP_COV_OFF: src = c->b_tint_cov_off.p; sz = (size_t)(c->hb.n_tints + 1) * 8; break;
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

