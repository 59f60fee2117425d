// cluster.cu -- C ABI of the cluster-stage pre-processing (include/freddie_b200.h, frs_cprep_*).
// Replaces read_segment's rep merge, preprocess_ilp and partition_reads of freddie_cluster.py (:154-164, :277-328,
// :198-274) on the arrays the segment stage produces.  The quadratic steps are the kernels of
// kernels_cluster.cuh; this file owns the buffers, the launch order and the list bookkeeping between the
// device phases (stable grouping of reps by structure, components -> pieces: linear passes over small arrays).
#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "../../include/freddie_b200.h"
#include "kernels_cluster.cuh"

namespace {

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as() const { return (T*)p; }
};

}  // namespace

struct frs_cprep {
  int device = 0;
  cudaStream_t st = nullptr;
  char err[512] = "";
  int launches = 0;
  // inputs
  Buf d_tint_read_off, d_tint_seg_n, d_tint_digit_off, d_read_row, d_digits, d_head, d_gap_off, d_gap_rec;
  // plan (host-computed offsets)
  Buf d_tint_row_off, d_rowword_off, d_tabr_off, d_tabr_cap, d_tabn_off, d_tabn_cap, d_out_off;
  // phase 1
  Buf d_rowbits, d_row_f, d_row_l, d_row_hash, d_row_tint, d_row_slot, d_row_first, d_tab;
  Buf d_gsort, d_read_tint, d_key_row, d_key_cnt, d_key_pe, d_key_ps, d_key_hash, d_slot, d_first, d_flag, d_count, d_scan,
      d_sums, d_total;
  Buf d_read_rep, d_rep_read, d_rep_tint, d_rep_first_read, d_rep_count, d_rep_fl, d_rep_cat, d_rep_gap, d_rep_row, d_rep_hash,
      d_tint_rep_off;
  Buf d_sslot, d_sfirst, d_sflag, d_scount, d_sscan, d_rep_struct, d_tint_struct_off, d_s_tint, d_s_row, d_s_f, d_s_l, d_s_cat,
      d_s_cnt;
  Buf d_I, d_C;
  // phase 2
  Buf d_sb_off, d_adj_off, d_sbits, d_adj_a, d_adj_b, d_deg, d_active_a, d_active_b, d_any, d_parent, d_label, d_edges;
  // phase 3
  Buf d_q_node, d_q_end, d_mem_off, d_mem, d_q_pairs, d_q_pair_off, d_inc, d_err;
  u32* adj_final = nullptr;
  // host copies kept for the fetch
  frs_cluster_batch hb{};
  frs_cluster_sizes sizes{};
  std::vector<int> h_tint_rep_off, h_tint_struct_off, h_tint_part_off, h_part_rid_off, h_part_rids, h_rep_struct;
  std::vector<i64> h_part_inc_off, h_out_off, h_edges;
  bool ran = false;
  cudaEvent_t ev[6] = {};
  float ms[5] = {};
};

namespace {

int fail(frs_cprep* c, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(c->err, sizeof c->err, fmt, ap);
  va_end(ap);
  return code;
}

#define CPK(call)                                                                                           \
  do {                                                                                                      \
    cudaError_t e_ = (call);                                                                                \
    if (e_ != cudaSuccess) return fail(c, FRS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int ensure(frs_cprep* c, Buf& b, size_t bytes) {
  if (bytes < 16) bytes = 16;
  if (b.cap >= bytes) return 0;
  if (b.p) CPK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  const size_t want = bytes + bytes / 4;
  if (cudaMalloc(&b.p, want) != cudaSuccess) {
    cudaGetLastError();
    CPK(cudaMalloc(&b.p, bytes));
    b.cap = bytes;
  } else {
    b.cap = want;
  }
  return 0;
}
#define ENS(buf, bytes)                        \
  do {                                         \
    int r_ = ensure(c, c->buf, (size_t)(bytes)); \
    if (r_) return r_;                         \
  } while (0)
#define UP(buf, src, bytes)                                                                                  \
  do {                                                                                                       \
    ENS(buf, bytes);                                                                                         \
    if ((bytes) > 0) CPK(cudaMemcpyAsync(c->buf.p, src, (size_t)(bytes), cudaMemcpyHostToDevice, c->st));     \
  } while (0)

inline unsigned blocks(i64 n, int per) { return (unsigned)std::max<i64>(1, (n + per - 1) / per); }
inline int pow2_at_least(i64 n) {
  int c = 2;
  while (c < n) c <<= 1;
  return c;
}

template <class TIn, class TOut>
int scan_exclusive(frs_cprep* c, const TIn* in, i64 n, TOut* out, TOut* h_total) {
  const int per = CP_SCAN_THREADS * CP_SCAN_ITEMS;
  const int nb = (int)blocks(n, per);
  ENS(d_sums, (size_t)nb * sizeof(TOut));
  ENS(d_total, sizeof(TOut));
  if (n > 0) {
    k_cp_scan_sums<TIn, TOut><<<nb, CP_SCAN_THREADS, 0, c->st>>>(in, n, c->d_sums.as<TOut>());
    k_cp_scan_top<TOut><<<1, CP_SCAN_THREADS, 0, c->st>>>(c->d_sums.as<TOut>(), nb, c->d_total.as<TOut>());
    k_cp_scan_apply<TIn, TOut><<<nb, CP_SCAN_THREADS, 0, c->st>>>(in, n, c->d_sums.as<TOut>(), out);
    c->launches += 3;
    CPK(cudaMemcpyAsync(h_total, c->d_total.p, sizeof(TOut), cudaMemcpyDeviceToHost, c->st));
    CPK(cudaStreamSynchronize(c->st));
  } else {
    *h_total = 0;
  }
  return 0;
}

}  // namespace

extern "C" {

int frs_cprep_create(int device, frs_cprep** out) {
  if (!out) return FRS_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return FRS_ERR_CUDA;  // no device: no CPU fallback
  }
  frs_cprep* c = new frs_cprep();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return FRS_ERR_CUDA;
  }
  for (auto& e : c->ev) cudaEventCreate(&e);
  *out = c;
  return 0;
}

void frs_cprep_destroy(frs_cprep* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  Buf* all[] = {&c->d_tint_read_off, &c->d_tint_seg_n, &c->d_tint_digit_off, &c->d_read_row, &c->d_digits, &c->d_head,
                &c->d_gap_off, &c->d_gap_rec, &c->d_tint_row_off, &c->d_rowword_off, &c->d_tabr_off, &c->d_tabr_cap,
                &c->d_tabn_off, &c->d_tabn_cap, &c->d_out_off, &c->d_rowbits, &c->d_row_f, &c->d_row_l, &c->d_row_hash,
                &c->d_row_tint, &c->d_row_slot, &c->d_row_first, &c->d_tab, &c->d_gsort, &c->d_read_tint, &c->d_key_row,
                &c->d_key_cnt, &c->d_key_pe, &c->d_key_ps, &c->d_key_hash, &c->d_slot, &c->d_first, &c->d_flag, &c->d_count,
                &c->d_scan, &c->d_sums, &c->d_total, &c->d_read_rep, &c->d_rep_read, &c->d_rep_tint, &c->d_rep_first_read,
                &c->d_rep_count, &c->d_rep_fl, &c->d_rep_cat, &c->d_rep_gap, &c->d_rep_row, &c->d_rep_hash, &c->d_tint_rep_off,
                &c->d_sslot, &c->d_sfirst, &c->d_sflag, &c->d_scount, &c->d_sscan, &c->d_rep_struct, &c->d_tint_struct_off,
                &c->d_s_tint, &c->d_s_row, &c->d_s_f, &c->d_s_l, &c->d_s_cat, &c->d_s_cnt, &c->d_I, &c->d_C, &c->d_sb_off,
                &c->d_adj_off, &c->d_sbits, &c->d_adj_a, &c->d_adj_b, &c->d_deg, &c->d_active_a, &c->d_active_b, &c->d_any,
                &c->d_parent, &c->d_label, &c->d_edges, &c->d_q_node, &c->d_q_end, &c->d_mem_off, &c->d_mem, &c->d_q_pairs,
                &c->d_q_pair_off, &c->d_inc, &c->d_err};
  for (Buf* b : all)
    if (b->p) cudaFree(b->p);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->st);
  delete c;
}

const char* frs_cprep_last_error(frs_cprep* c) { return c ? c->err : "null context"; }

int frs_cprep_timings(frs_cprep* c, float* ms, int n) {
  if (!c || !ms) return FRS_ERR_ARG;
  for (int k = 0; k < n && k < 5; ++k) ms[k] = c->ms[k];
  return 5;
}

int frs_cprep_run(frs_cprep* c, const frs_cluster_batch* b, int maximum_ilp_size, frs_cluster_sizes* sizes) {
  if (!c || !b || !sizes) return FRS_ERR_ARG;
  c->ran = false;
  c->launches = 0;
  CPK(cudaSetDevice(c->device));
  const int T = b->n_tints, N = b->n_reads;
  if (T < 0 || N < 0) return fail(c, FRS_ERR_ARG, "negative counts");
  if (maximum_ilp_size < 1) return fail(c, FRS_ERR_ARG, "ZeroDivisionError: maximum_ilp_size must be >= 1 (split_list_evenly, :112)");
  if (T > 0 && (!b->tint_read_off || !b->tint_seg_n || !b->tint_digit_off)) return fail(c, FRS_ERR_ARG, "null tint tables");
  if (N > 0 && (!b->read_row || !b->digits || !b->read_head || !b->read_gap_off)) return fail(c, FRS_ERR_ARG, "null read tables");
  c->hb = *b;
  // ---- plan: rows, hash-table regions, output offsets ----
  std::vector<int> tint_row_off(T + 1, 0), tabr_cap(std::max(T, 1)), tabn_cap(std::max(T, 1));
  std::vector<i64> rowword_off(T + 1, 0), tabr_off(T + 1, 0), tabn_off(T + 1, 0);
  if (T > 0 && (b->tint_read_off[0] != 0 || b->tint_read_off[T] != N)) return fail(c, FRS_ERR_ARG, "tint_read_off does not span the reads");
  for (int t = 0; t < T; ++t) {
    const int M = b->tint_seg_n[t];
    const i64 bytes = b->tint_digit_off[t + 1] - b->tint_digit_off[t];
    const int nr = b->tint_read_off[t + 1] - b->tint_read_off[t];
    if (M < 1 || bytes < 0 || bytes % M || nr < 0) return fail(c, FRS_ERR_ARG, "tint %d: %lld digit bytes for M = %d", t, (long long)bytes, M);
    const i64 rows = bytes / M;
    if (rows > 0x7fffffff - tint_row_off[t]) return fail(c, FRS_ERR_LIMIT, "more than 2^31 digit rows");
    tint_row_off[t + 1] = tint_row_off[t] + (int)rows;
    rowword_off[t + 1] = rowword_off[t] + rows * ((M + 31) / 32);
    tabr_cap[t] = pow2_at_least(2 * rows);
    tabn_cap[t] = pow2_at_least(2 * (i64)nr);
    tabr_off[t + 1] = tabr_off[t] + tabr_cap[t];
    tabn_off[t + 1] = tabn_off[t] + tabn_cap[t];
  }
  const int n_rows = tint_row_off[T];
  for (int t = 0; t < T; ++t) {  // read_row inside the tint's rows (one linear pass; the device trusts it)
    const int rows = tint_row_off[t + 1] - tint_row_off[t];
    for (int i = b->tint_read_off[t]; i < b->tint_read_off[t + 1]; ++i)
      if (b->read_row[i] < 0 || b->read_row[i] >= rows) return fail(c, FRS_ERR_ARG, "read %d: digit row %d outside its tint", i, b->read_row[i]);
  }
  const i64 G = N > 0 ? b->read_gap_off[N] : 0;
  if (G < 0 || (G > 0 && !b->gap_rec)) return fail(c, FRS_ERR_ARG, "gap records missing");
  const i64 D = T > 0 ? b->tint_digit_off[T] : 0;

  cudaEventRecord(c->ev[0], c->st);
  UP(d_tint_read_off, b->tint_read_off, (size_t)(T + 1) * 4);
  UP(d_tint_seg_n, b->tint_seg_n, (size_t)T * 4);
  UP(d_tint_digit_off, b->tint_digit_off, (size_t)(T + 1) * 8);
  UP(d_read_row, b->read_row, (size_t)N * 4);
  UP(d_digits, b->digits, (size_t)D);
  UP(d_head, b->read_head, (size_t)N * 32);
  UP(d_gap_off, b->read_gap_off, (size_t)(N + 1) * 4);
  UP(d_gap_rec, b->gap_rec, (size_t)G * 12);
  UP(d_tint_row_off, tint_row_off.data(), (size_t)(T + 1) * 4);
  UP(d_rowword_off, rowword_off.data(), (size_t)(T + 1) * 8);
  UP(d_tabr_off, tabr_off.data(), (size_t)(T + 1) * 8);
  UP(d_tabr_cap, tabr_cap.data(), (size_t)T * 4);
  UP(d_tabn_off, tabn_off.data(), (size_t)(T + 1) * 8);
  UP(d_tabn_cap, tabn_cap.data(), (size_t)T * 4);
  ENS(d_err, 16);
  CPK(cudaMemsetAsync(c->d_err.p, 0, 16, c->st));

  // ---- phase 1a: digit rows -> bits, merged by content ----
  const i64 tab_words = std::max(tabr_off[T], tabn_off[T]);
  ENS(d_tab, (size_t)tab_words * 4);
  ENS(d_rowbits, (size_t)rowword_off[T] * 4);
  ENS(d_row_f, (size_t)n_rows * 4);
  ENS(d_row_l, (size_t)n_rows * 4);
  ENS(d_row_hash, (size_t)n_rows * 8);
  ENS(d_row_tint, (size_t)n_rows * 4);
  ENS(d_row_slot, (size_t)n_rows * 4);
  ENS(d_row_first, (size_t)n_rows * 4);
  int* d_err = c->d_err.as<int>();
  if (n_rows > 0) {
    k_cp_fill<int><<<std::min(blocks(tabr_off[T], 256), 4096u), 256, 0, c->st>>>(c->d_tab.as<int>(), tabr_off[T], CP_EMPTY);
    k_cp_row_bits<<<blocks((i64)n_rows * 32, 256), 256, 0, c->st>>>(
        n_rows, T, c->d_tint_row_off.as<int>(), c->d_tint_seg_n.as<int>(), c->d_tint_digit_off.as<i64>(),
        c->d_rowword_off.as<i64>(), c->d_digits.as<u8>(), c->d_rowbits.as<u32>(), c->d_row_f.as<int>(), c->d_row_l.as<int>(),
        c->d_row_hash.as<u64>(), c->d_row_tint.as<int>(), d_err);
    k_cp_row_insert<<<blocks(n_rows, 128), 128, 0, c->st>>>(
        n_rows, c->d_row_tint.as<int>(), c->d_tint_row_off.as<int>(), c->d_tint_seg_n.as<int>(), c->d_rowword_off.as<i64>(),
        c->d_rowbits.as<u32>(), c->d_row_hash.as<u64>(), c->d_tabr_off.as<i64>(), c->d_tabr_cap.as<int>(), c->d_tab.as<int>(),
        c->d_row_slot.as<int>());
    k_cp_resolve<<<blocks(n_rows, 256), 256, 0, c->st>>>(n_rows, c->d_row_tint.as<int>(), c->d_tabr_off.as<i64>(),
                                                         c->d_tab.as<int>(), c->d_row_slot.as<int>(), c->d_row_first.as<int>());
    c->launches += 4;
  }
  // ---- phase 1b: read keys, reps in first-seen order ----
  ENS(d_gsort, (size_t)G * 12);
  ENS(d_read_tint, (size_t)N * 4);
  ENS(d_key_row, (size_t)N * 4);
  ENS(d_key_cnt, (size_t)N * 4);
  ENS(d_key_pe, (size_t)N * 4);
  ENS(d_key_ps, (size_t)N * 4);
  ENS(d_key_hash, (size_t)N * 8);
  ENS(d_slot, (size_t)N * 4);
  ENS(d_first, (size_t)N * 4);
  ENS(d_flag, (size_t)N * 4);
  ENS(d_count, (size_t)N * 4);
  ENS(d_scan, (size_t)(N + 1) * 4);
  ENS(d_read_rep, (size_t)N * 4);
  ENS(d_tint_rep_off, (size_t)(T + 1) * 4);
  ENS(d_tint_struct_off, (size_t)(T + 1) * 4);
  int U = 0, S_tot = 0;
  if (N > 0) {
    k_cp_fill<int><<<std::min(blocks(tabn_off[T], 256), 4096u), 256, 0, c->st>>>(c->d_tab.as<int>(), tabn_off[T], CP_EMPTY);
    CPK(cudaMemsetAsync(c->d_count.p, 0, (size_t)N * 4, c->st));
    k_cp_read_key<<<blocks(N, 128), 128, 0, c->st>>>(
        N, T, c->d_tint_read_off.as<int>(), c->d_tint_row_off.as<int>(), c->d_read_row.as<int>(), c->d_row_first.as<int>(),
        c->d_head.as<int>(), c->d_gap_off.as<int>(), c->d_gap_rec.as<int>(), c->d_tint_seg_n.as<int>(), c->d_gsort.as<int>(),
        c->d_read_tint.as<int>(), c->d_key_row.as<int>(), c->d_key_cnt.as<int>(), c->d_key_pe.as<int>(), c->d_key_ps.as<int>(),
        c->d_key_hash.as<u64>(), d_err);
    k_cp_read_insert<<<blocks(N, 128), 128, 0, c->st>>>(
        N, c->d_read_tint.as<int>(), c->d_key_row.as<int>(), c->d_key_cnt.as<int>(), c->d_key_pe.as<int>(), c->d_key_ps.as<int>(),
        c->d_key_hash.as<u64>(), c->d_gap_off.as<int>(), c->d_gsort.as<int>(), c->d_tabn_off.as<i64>(), c->d_tabn_cap.as<int>(),
        c->d_tab.as<int>(), c->d_slot.as<int>());
    k_cp_resolve<<<blocks(N, 256), 256, 0, c->st>>>(N, c->d_read_tint.as<int>(), c->d_tabn_off.as<i64>(), c->d_tab.as<int>(),
                                                    c->d_slot.as<int>(), c->d_first.as<int>());
    k_cp_flag_count<<<blocks(N, 256), 256, 0, c->st>>>(N, c->d_first.as<int>(), c->d_flag.as<int>(), c->d_count.as<int>());
    c->launches += 5;
    int r = scan_exclusive<int, int>(c, c->d_flag.as<int>(), N, c->d_scan.as<int>(), &U);
    if (r) return r;
  }
  // ---- phase 1c: preprocess_ilp per rep, structures in first-seen order ----
  ENS(d_rep_read, (size_t)U * 4);
  ENS(d_rep_tint, (size_t)U * 4);
  ENS(d_rep_first_read, (size_t)U * 4);
  ENS(d_rep_count, (size_t)U * 4);
  ENS(d_rep_fl, (size_t)U * 8);
  ENS(d_rep_cat, (size_t)U);
  ENS(d_rep_gap, (size_t)U * 12);
  ENS(d_rep_row, (size_t)U * 4);
  ENS(d_rep_hash, (size_t)U * 8);
  ENS(d_sslot, (size_t)U * 4);
  ENS(d_sfirst, (size_t)U * 4);
  ENS(d_sflag, (size_t)U * 4);
  ENS(d_scount, (size_t)U * 4);
  ENS(d_sscan, (size_t)(U + 1) * 4);
  ENS(d_rep_struct, (size_t)U * 4);
  if (N > 0) {
    k_cp_rep_prep<<<blocks(N, 256), 256, 0, c->st>>>(
        N, c->d_read_tint.as<int>(), c->d_tint_read_off.as<int>(), c->d_tint_seg_n.as<int>(), c->d_first.as<int>(),
        c->d_scan.as<int>(), c->d_count.as<int>(), c->d_key_row.as<int>(), c->d_row_f.as<int>(), c->d_row_l.as<int>(),
        c->d_head.as<int>(), c->d_read_rep.as<int>(), c->d_rep_read.as<int>(), c->d_rep_tint.as<int>(),
        c->d_rep_first_read.as<int>(), c->d_rep_count.as<int>(), c->d_rep_fl.as<int>(), c->d_rep_cat.as<u8>(),
        c->d_rep_gap.as<int>(), c->d_rep_row.as<int>(), c->d_rep_hash.as<u64>());
    c->launches += 1;
  }
  k_cp_offsets<<<blocks(T + 1, 256), 256, 0, c->st>>>(T, c->d_tint_read_off.as<int>(), c->d_scan.as<int>(), N, U,
                                                      c->d_tint_rep_off.as<int>());
  c->launches += 1;
  if (U > 0) {
    // the read tables' regions are large enough for the reps of a tint (U_t <= N_t)
    k_cp_fill<int><<<std::min(blocks(tabn_off[T], 256), 4096u), 256, 0, c->st>>>(c->d_tab.as<int>(), tabn_off[T], CP_EMPTY);
    CPK(cudaMemsetAsync(c->d_scount.p, 0, (size_t)U * 4, c->st));
    k_cp_struct_insert<<<blocks(U, 128), 128, 0, c->st>>>(U, c->d_rep_tint.as<int>(), c->d_rep_row.as<int>(), c->d_rep_cat.as<u8>(),
                                                         c->d_rep_hash.as<u64>(), c->d_tabn_off.as<i64>(), c->d_tabn_cap.as<int>(),
                                                         c->d_tab.as<int>(), c->d_sslot.as<int>());
    k_cp_resolve<<<blocks(U, 256), 256, 0, c->st>>>(U, c->d_rep_tint.as<int>(), c->d_tabn_off.as<i64>(), c->d_tab.as<int>(),
                                                    c->d_sslot.as<int>(), c->d_sfirst.as<int>());
    k_cp_flag_count<<<blocks(U, 256), 256, 0, c->st>>>(U, c->d_sfirst.as<int>(), c->d_sflag.as<int>(), c->d_scount.as<int>());
    c->launches += 4;
    int r = scan_exclusive<int, int>(c, c->d_sflag.as<int>(), U, c->d_sscan.as<int>(), &S_tot);
    if (r) return r;
  }
  ENS(d_s_tint, (size_t)S_tot * 4);
  ENS(d_s_row, (size_t)S_tot * 4);
  ENS(d_s_f, (size_t)S_tot * 4);
  ENS(d_s_l, (size_t)S_tot * 4);
  ENS(d_s_cat, (size_t)S_tot);
  ENS(d_s_cnt, (size_t)S_tot * 4);
  if (U > 0) {
    k_cp_struct_fill<<<blocks(U, 256), 256, 0, c->st>>>(
        U, c->d_rep_tint.as<int>(), c->d_tint_rep_off.as<int>(), c->d_sfirst.as<int>(), c->d_sscan.as<int>(),
        c->d_scount.as<int>(), c->d_rep_row.as<int>(), c->d_rep_fl.as<int>(), c->d_rep_cat.as<u8>(), c->d_rep_struct.as<int>(),
        c->d_s_tint.as<int>(), c->d_s_row.as<int>(), c->d_s_f.as<int>(), c->d_s_l.as<int>(), c->d_s_cat.as<u8>(),
        c->d_s_cnt.as<int>());
    c->launches += 1;
  }
  k_cp_offsets<<<blocks(T + 1, 256), 256, 0, c->st>>>(T, c->d_tint_rep_off.as<int>(), c->d_sscan.as<int>(), U, S_tot,
                                                      c->d_tint_struct_off.as<int>());
  c->launches += 1;
  // sizes of the per-tint blocks (one small read-back: the adjacency matrices are sized from it)
  c->h_tint_rep_off.assign(T + 1, 0);
  c->h_tint_struct_off.assign(T + 1, 0);
  CPK(cudaMemcpyAsync(c->h_tint_rep_off.data(), c->d_tint_rep_off.p, (size_t)(T + 1) * 4, cudaMemcpyDeviceToHost, c->st));
  CPK(cudaMemcpyAsync(c->h_tint_struct_off.data(), c->d_tint_struct_off.p, (size_t)(T + 1) * 4, cudaMemcpyDeviceToHost, c->st));
  int h_err[4] = {0, 0, 0, 0};
  CPK(cudaMemcpyAsync(h_err, c->d_err.p, 16, cudaMemcpyDeviceToHost, c->st));
  CPK(cudaStreamSynchronize(c->st));
  if (h_err[0] == CPERR_DIGIT) return fail(c, FRS_ERR_ARG, "a digit row holds a character other than 0, 1, 2");
  if (h_err[0] == CPERR_GAP) return fail(c, FRS_ERR_ARG, "AssertionError: 0 <= g[0] < g[1] < len(read['data']) (freddie_cluster.py:164)");
  // I / C rows
  c->h_out_off.assign(T + 1, 0);
  std::vector<i64> sb_off(T + 1, 0), adj_off(T + 1, 0);
  for (int t = 0; t < T; ++t) {
    const i64 Ut = c->h_tint_rep_off[t + 1] - c->h_tint_rep_off[t], St = c->h_tint_struct_off[t + 1] - c->h_tint_struct_off[t];
    const i64 M = b->tint_seg_n[t];
    c->h_out_off[t + 1] = c->h_out_off[t] + Ut * M;
    sb_off[t + 1] = sb_off[t] + St * ((M + 31) / 32);
    adj_off[t + 1] = adj_off[t] + St * ((St + 31) / 32);
  }
  UP(d_out_off, c->h_out_off.data(), (size_t)(T + 1) * 8);
  ENS(d_I, (size_t)c->h_out_off[T]);
  ENS(d_C, (size_t)c->h_out_off[T]);
  if (U > 0) {
    k_cp_rows_out<<<blocks((i64)U * 32, 256), 256, 0, c->st>>>(
        U, c->d_rep_tint.as<int>(), c->d_tint_rep_off.as<int>(), c->d_rep_read.as<int>(), c->d_read_row.as<int>(),
        c->d_tint_seg_n.as<int>(), c->d_tint_digit_off.as<i64>(), c->d_digits.as<u8>(), c->d_rep_fl.as<int>(),
        c->d_out_off.as<i64>(), c->d_I.as<u8>(), c->d_C.as<u8>());
    c->launches += 1;
  }
  cudaEventRecord(c->ev[1], c->st);

  // ---- phase 2: pair test, pruning rounds, components ----
  UP(d_sb_off, sb_off.data(), (size_t)(T + 1) * 8);
  UP(d_adj_off, adj_off.data(), (size_t)(T + 1) * 8);
  ENS(d_sbits, (size_t)sb_off[T] * 4);
  ENS(d_adj_a, (size_t)adj_off[T] * 4);
  ENS(d_adj_b, (size_t)adj_off[T] * 4);
  ENS(d_deg, (size_t)S_tot * 4);
  ENS(d_active_a, (size_t)std::max(T, 1) * 4);
  ENS(d_active_b, (size_t)std::max(T, 1) * 4);
  ENS(d_any, 4);
  ENS(d_parent, (size_t)S_tot * 4);
  ENS(d_label, (size_t)S_tot * 4);
  ENS(d_edges, (size_t)std::max(T, 1) * 16);
  CPK(cudaMemsetAsync(c->d_edges.p, 0, (size_t)std::max(T, 1) * 16, c->st));
  int rounds = 0;
  u32* A = c->d_adj_a.as<u32>();
  u32* B = c->d_adj_b.as<u32>();
  if (S_tot > 0) {
    const unsigned wg = blocks((i64)S_tot * 32, 256);
    k_cp_struct_bits<<<blocks(S_tot, 128), 128, 0, c->st>>>(
        S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(), c->d_s_row.as<int>(), c->d_tint_row_off.as<int>(),
        c->d_tint_seg_n.as<int>(), c->d_rowword_off.as<i64>(), c->d_rowbits.as<u32>(), c->d_sb_off.as<i64>(), c->d_sbits.as<u32>());
    k_cp_pair_test<<<wg, 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(), c->d_s_f.as<int>(),
                                          c->d_s_l.as<int>(), c->d_s_cat.as<u8>(), c->d_sb_off.as<i64>(), c->d_sbits.as<u32>(),
                                          c->d_adj_off.as<i64>(), A, c->d_edges.as<i64>());
    c->launches += 2;
  }
  cudaEventRecord(c->ev[2], c->st);
  if (S_tot > 0) {
    const unsigned wg = blocks((i64)S_tot * 32, 256);
    // B starts as a copy, so that the rows of tints that stop changing are valid in both buffers
    CPK(cudaMemcpyAsync(B, A, (size_t)adj_off[T] * 4, cudaMemcpyDeviceToDevice, c->st));
    k_cp_fill<int><<<blocks(T, 256), 256, 0, c->st>>>(c->d_active_a.as<int>(), T, 1);
    c->launches += 1;
    int* act = c->d_active_a.as<int>();
    int* act_next = c->d_active_b.as<int>();
    while (true) {
      CPK(cudaMemsetAsync(act_next, 0, (size_t)T * 4, c->st));
      CPK(cudaMemsetAsync(c->d_any.p, 0, 4, c->st));
      k_cp_degree<<<wg, 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(), c->d_adj_off.as<i64>(), A,
                                         act, c->d_deg.as<int>());
      k_cp_prune_round<<<wg, 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(), c->d_adj_off.as<i64>(),
                                              A, B, c->d_deg.as<int>(), act, act_next, c->d_any.as<int>());
      c->launches += 2;
      ++rounds;
      int any = 0;
      CPK(cudaMemcpyAsync(&any, c->d_any.p, 4, cudaMemcpyDeviceToHost, c->st));
      CPK(cudaStreamSynchronize(c->st));
      if (!any) break;  // nothing removed: A == B for every tint
      // tints that changed: B is their new graph; A is stale for them -> bring A up to date lazily by swapping
      // roles, after copying the rows of the changed tints is avoided by keeping both buffers equal for the
      // tints that did not change (they were equal before the round and were not written)
      std::swap(A, B);
      std::swap(act, act_next);
      // the new B (old A) is stale for the tints that changed this round: they are active in the next round and
      // every row of an active tint is rewritten, so it is valid again before anyone reads it
    }
  }
  c->adj_final = A;
  cudaEventRecord(c->ev[3], c->st);
  if (S_tot > 0) {
    const unsigned wg = blocks((i64)S_tot * 32, 256);
    k_cp_parent_init<<<blocks(S_tot, 256), 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(),
                                                            c->d_parent.as<int>());
    k_cp_union<<<wg, 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(), c->d_adj_off.as<i64>(), A,
                                      c->d_parent.as<int>(), c->d_edges.as<i64>());
    k_cp_labels<<<blocks(S_tot, 256), 256, 0, c->st>>>(S_tot, c->d_s_tint.as<int>(), c->d_tint_struct_off.as<int>(),
                                                       c->d_parent.as<int>(), c->d_label.as<int>());
    c->launches += 3;
  }
  cudaEventRecord(c->ev[4], c->st);
  // ---- host bookkeeping: reps grouped by structure, components -> pieces -> partitions ----
  std::vector<int> label(S_tot), s_cnt(S_tot);
  c->h_rep_struct.assign(U, 0);
  c->h_edges.assign((size_t)std::max(T, 1) * 2, 0);
  if (S_tot > 0) {
    CPK(cudaMemcpyAsync(label.data(), c->d_label.p, (size_t)S_tot * 4, cudaMemcpyDeviceToHost, c->st));
    CPK(cudaMemcpyAsync(s_cnt.data(), c->d_s_cnt.p, (size_t)S_tot * 4, cudaMemcpyDeviceToHost, c->st));
    CPK(cudaMemcpyAsync(c->h_rep_struct.data(), c->d_rep_struct.p, (size_t)U * 4, cudaMemcpyDeviceToHost, c->st));
  }
  CPK(cudaMemcpyAsync(c->h_edges.data(), c->d_edges.p, (size_t)std::max(T, 1) * 16, cudaMemcpyDeviceToHost, c->st));
  CPK(cudaStreamSynchronize(c->st));
  std::vector<int> mem_off(S_tot + 1, 0), mem(U);
  for (int s = 0; s < S_tot; ++s) mem_off[s + 1] = mem_off[s] + s_cnt[s];
  if (mem_off[S_tot] != U) return fail(c, FRS_ERR_ASSERT, "internal: structure sizes do not add up");
  {
    std::vector<int> cur(mem_off.begin(), mem_off.end() - 1);
    for (int t = 0; t < T; ++t)
      for (int u = c->h_tint_rep_off[t]; u < c->h_tint_rep_off[t + 1]; ++u)  // ascending rep ids: the order of unique_data[i][1]
        mem[cur[c->h_tint_struct_off[t] + c->h_rep_struct[u]]++] = u - c->h_tint_rep_off[t];
  }
  std::vector<int> q_node, q_end;
  q_node.reserve(S_tot);
  q_end.reserve(S_tot);
  c->h_tint_part_off.assign(T + 1, 0);
  c->h_part_rid_off.assign(1, 0);
  c->h_part_rids.clear();
  c->h_part_rids.reserve(U);
  std::vector<int> order, comp_start, piece_q0;
  for (int t = 0; t < T; ++t) {
    const int s0 = c->h_tint_struct_off[t], S = c->h_tint_struct_off[t + 1] - s0;
    // nodes grouped by the root of their component (= its smallest node), ascending inside: counting sort
    order.assign(S, 0);
    comp_start.assign(S + 1, 0);
    for (int i = 0; i < S; ++i) comp_start[label[s0 + i] + 1]++;
    for (int i = 0; i < S; ++i) comp_start[i + 1] += comp_start[i];
    {
      std::vector<int> cur(comp_start.begin(), comp_start.end() - 1);
      for (int i = 0; i < S; ++i) order[cur[label[s0 + i]]++] = i;
    }
    int parts = 0;
    for (int root = 0; root < S; ++root) {
      const int n = comp_start[root + 1] - comp_start[root];
      if (n == 0) continue;
      const int* comp = order.data() + comp_start[root];
      // split_list_evenly (:112-116)
      const int p = (n + maximum_ilp_size - 1) / maximum_ilp_size;
      const int s = (n + p - 1) / p;
      for (int idx = 0; idx < p * s; idx += s) {
        const int lo = std::min(idx, n), hi = std::min(idx + s, n);
        const int q0 = (int)q_node.size();
        piece_q0.push_back(q0);  // a piece may be empty: it owns no position then
        for (int k = lo; k < hi; ++k) {
          q_node.push_back(s0 + comp[k]);
          const int gs = s0 + comp[k];
          for (int x = mem_off[gs]; x < mem_off[gs + 1]; ++x) c->h_part_rids.push_back(mem[x]);
        }
        for (int k = lo; k < hi; ++k) q_end.push_back(q0 + (hi - lo));
        c->h_part_rid_off.push_back((int)c->h_part_rids.size());
        ++parts;
      }
    }
    c->h_tint_part_off[t + 1] = c->h_tint_part_off[t] + parts;
  }
  const int P = c->h_tint_part_off[T];
  const int Q = (int)q_node.size();
  piece_q0.push_back(Q);
  // ---- phase 3: incompatible pairs ----
  UP(d_q_node, q_node.data(), (size_t)Q * 4);
  UP(d_q_end, q_end.data(), (size_t)Q * 4);
  UP(d_mem_off, mem_off.data(), (size_t)(S_tot + 1) * 4);
  UP(d_mem, mem.data(), (size_t)U * 4);
  ENS(d_q_pairs, (size_t)Q * 8);
  ENS(d_q_pair_off, (size_t)(Q + 1) * 8);
  i64 n_inc = 0;
  std::vector<i64> q_pair_off(Q + 1, 0);
  if (Q > 0) {
    const unsigned wg = blocks((i64)Q * 32, 256);
    k_cp_incomp<0><<<wg, 256, 0, c->st>>>(Q, c->d_q_node.as<int>(), c->d_q_end.as<int>(), c->d_s_tint.as<int>(),
                                          c->d_tint_struct_off.as<int>(), c->d_adj_off.as<i64>(), A, c->d_mem_off.as<int>(),
                                          c->d_mem.as<int>(), c->d_q_pairs.as<i64>(), nullptr, nullptr);
    c->launches += 1;
    int r = scan_exclusive<i64, i64>(c, c->d_q_pairs.as<i64>(), Q, c->d_q_pair_off.as<i64>(), &n_inc);
    if (r) return r;
    if (n_inc > ((i64)1 << 40)) return fail(c, FRS_ERR_LIMIT, "%lld incompatible pairs", (long long)n_inc);
    ENS(d_inc, (size_t)n_inc * 8);
    k_cp_incomp<1><<<wg, 256, 0, c->st>>>(Q, c->d_q_node.as<int>(), c->d_q_end.as<int>(), c->d_s_tint.as<int>(),
                                          c->d_tint_struct_off.as<int>(), c->d_adj_off.as<i64>(), A, c->d_mem_off.as<int>(),
                                          c->d_mem.as<int>(), nullptr, c->d_q_pair_off.as<i64>(), c->d_inc.as<int>());
    c->launches += 1;
    CPK(cudaMemcpyAsync(q_pair_off.data(), c->d_q_pair_off.p, (size_t)Q * 8, cudaMemcpyDeviceToHost, c->st));
  }
  cudaEventRecord(c->ev[5], c->st);
  CPK(cudaStreamSynchronize(c->st));
  CPK(cudaGetLastError());
  q_pair_off[Q] = n_inc;
  c->h_part_inc_off.assign(P + 1, 0);
  for (int p = 0; p <= P; ++p) c->h_part_inc_off[p] = q_pair_off[piece_q0[p]];
  for (int k = 0; k < 5; ++k) cudaEventElapsedTime(&c->ms[k], c->ev[k], c->ev[k + 1]);

  frs_cluster_sizes& z = c->sizes;
  z.n_reps = U;
  z.n_structs = S_tot;
  z.n_parts = P;
  z.n_incomp = n_inc;
  z.n_row_bytes = c->h_out_off[T];
  z.edges_before = z.edges_after = 0;
  for (int t = 0; t < T; ++t) {
    c->h_edges[2 * t] /= 2;
    c->h_edges[2 * t + 1] /= 2;
    z.edges_before += c->h_edges[2 * t];
    z.edges_after += c->h_edges[2 * t + 1];
  }
  z.prune_rounds = rounds;
  z.launches = c->launches;
  *sizes = z;
  c->ran = true;
  return 0;
}

int frs_cprep_fetch(frs_cprep* c, const frs_cluster_result* o) {
  if (!c || !o) return FRS_ERR_ARG;
  if (!c->ran) return fail(c, FRS_ERR_STATE, "frs_cprep_fetch before a successful frs_cprep_run");
  CPK(cudaSetDevice(c->device));
  const int T = c->hb.n_tints, N = c->hb.n_reads;
  const i64 U = c->sizes.n_reps;
#define DOWN(dst, buf, bytes)                                                                                   \
  do {                                                                                                          \
    if ((dst) && (bytes) > 0) CPK(cudaMemcpyAsync(dst, c->buf.p, (size_t)(bytes), cudaMemcpyDeviceToHost, c->st)); \
  } while (0)
  DOWN(o->read_rep, d_read_rep, (size_t)N * 4);
  DOWN(o->rep_first_read, d_rep_first_read, (size_t)U * 4);
  DOWN(o->rep_count, d_rep_count, (size_t)U * 4);
  DOWN(o->rep_fl, d_rep_fl, (size_t)U * 8);
  DOWN(o->rep_cat, d_rep_cat, (size_t)U);
  DOWN(o->rep_gap, d_rep_gap, (size_t)U * 12);
  DOWN(o->I, d_I, (size_t)c->sizes.n_row_bytes);
  DOWN(o->C, d_C, (size_t)c->sizes.n_row_bytes);
  DOWN(o->inc, d_inc, (size_t)c->sizes.n_incomp * 8);
#undef DOWN
  if (o->tint_rep_off) memcpy(o->tint_rep_off, c->h_tint_rep_off.data(), (size_t)(T + 1) * 4);
  if (o->tint_struct_off) memcpy(o->tint_struct_off, c->h_tint_struct_off.data(), (size_t)(T + 1) * 4);
  if (o->tint_row_off) memcpy(o->tint_row_off, c->h_out_off.data(), (size_t)(T + 1) * 8);
  if (o->rep_struct && U > 0) memcpy(o->rep_struct, c->h_rep_struct.data(), (size_t)U * 4);
  if (o->tint_part_off) memcpy(o->tint_part_off, c->h_tint_part_off.data(), (size_t)(T + 1) * 4);
  if (o->part_rid_off) memcpy(o->part_rid_off, c->h_part_rid_off.data(), c->h_part_rid_off.size() * 4);
  if (o->part_rids && U > 0) memcpy(o->part_rids, c->h_part_rids.data(), (size_t)U * 4);
  if (o->part_inc_off) memcpy(o->part_inc_off, c->h_part_inc_off.data(), c->h_part_inc_off.size() * 8);
  if (o->tint_edges && T > 0) memcpy(o->tint_edges, c->h_edges.data(), (size_t)T * 16);
  CPK(cudaStreamSynchronize(c->st));
  return 0;
}

}  // extern "C"
