// split.cu -- C ABI of the split-stage tint construction (include/freddie_b200.h, frs_split_*).
// Replaces get_transcriptional_intervals (freddie_split.py:295-364) and break_tint (:246-293) for a batch of read
// groups (what read_sam yields, :207-244).  Sorting, sweeps, union-find and the set unions of break_tint run in
// the kernels of kernels_split.cuh; this file owns the buffers, the launch order and the final list assembly.
#include <algorithm>
#include <cstdarg>
#include <cstring>
#include <vector>

#include "../../include/freddie_b200.h"
#include "kernels_split.cuh"

namespace {
struct SBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as() const { return (T*)p; }
};
}  // namespace

struct frs_split {
  int device = 0;
  cudaStream_t st = nullptr;
  char err[512] = "";
  int launches = 0;
  std::vector<SBuf*> all;
  SBuf d_group_read_off, d_read_iv_off, d_iv_s, d_iv_e, d_read_group, d_iv_read, d_keys_a, d_keys_b, d_hist, d_hist_scan,
      d_delta, d_open, d_first, d_sid_before, d_simple_key, d_simple_end, d_iv_sid, d_parent, d_root, d_read_comp, d_err,
      d_sums, d_total, d_node_of, d_big_reads, d_off_a, d_off_b, d_flag, d_pos, d_pairs, d_comp_reads, d_cnt64, d_off64,
      d_civ, d_node_root;
  frs_split_sizes sizes{};
  int n_groups = 0;
  std::vector<int> h_group_tint_off, h_tint_iv_off, h_tint_iv_s, h_tint_iv_e, h_tint_rid_off, h_tint_rids;
  bool ran = false;
  cudaEvent_t ev[2] = {};
  float ms = 0;
  frs_split() {
    SBuf* bs[] = {&d_group_read_off, &d_read_iv_off, &d_iv_s, &d_iv_e, &d_read_group, &d_iv_read, &d_keys_a, &d_keys_b, &d_hist,
                  &d_hist_scan, &d_delta, &d_open, &d_first, &d_sid_before, &d_simple_key, &d_simple_end, &d_iv_sid, &d_parent,
                  &d_root, &d_read_comp, &d_err, &d_sums, &d_total, &d_node_of, &d_big_reads, &d_off_a, &d_off_b, &d_flag, &d_pos,
                  &d_pairs, &d_comp_reads, &d_cnt64, &d_off64, &d_civ, &d_node_root};
    all.assign(bs, bs + sizeof bs / sizeof bs[0]);
  }
};

namespace {

int fail(frs_split* c, int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(c->err, sizeof c->err, fmt, ap);
  va_end(ap);
  return code;
}
#define SPK(call)                                                                                                  \
  do {                                                                                                             \
    cudaError_t e_ = (call);                                                                                       \
    if (e_ != cudaSuccess) return fail(c, FRS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

int ensure(frs_split* c, SBuf& b, size_t bytes) {
  if (bytes < 16) bytes = 16;
  if (b.cap >= bytes) return 0;
  if (b.p) SPK(cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  SPK(cudaMalloc(&b.p, bytes + bytes / 4));
  b.cap = bytes + bytes / 4;
  return 0;
}
#define SENS(buf, bytes)                           \
  do {                                             \
    int r_ = ensure(c, c->buf, (size_t)(bytes));   \
    if (r_) return r_;                             \
  } while (0)
#define SUP(buf, src, bytes)                                                                              \
  do {                                                                                                    \
    SENS(buf, bytes);                                                                                     \
    if ((bytes) > 0) SPK(cudaMemcpyAsync(c->buf.p, src, (size_t)(bytes), cudaMemcpyHostToDevice, c->st)); \
  } while (0)

inline unsigned nblk(i64 n, int per) { return (unsigned)std::max<i64>(1, (n + per - 1) / per); }
inline int bits_of(unsigned long long v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return std::max(b, 1);
}

template <class TIn, class TOut>
int scan_excl(frs_split* c, const TIn* in, i64 n, TOut* out, TOut* h_total) {
  const int nb = (int)nblk(n, CP_SCAN_THREADS * CP_SCAN_ITEMS);
  SENS(d_sums, (size_t)nb * sizeof(TOut));
  SENS(d_total, sizeof(TOut));
  if (h_total) *h_total = 0;
  if (n <= 0) return 0;
  k_cp_scan_sums<TIn, TOut><<<nb, CP_SCAN_THREADS, 0, c->st>>>(in, n, c->d_sums.as<TOut>());
  k_cp_scan_top<TOut><<<1, CP_SCAN_THREADS, 0, c->st>>>(c->d_sums.as<TOut>(), nb, c->d_total.as<TOut>());
  k_cp_scan_apply<TIn, TOut><<<nb, CP_SCAN_THREADS, 0, c->st>>>(in, n, c->d_sums.as<TOut>(), out);
  c->launches += 3;
  if (h_total) {
    SPK(cudaMemcpyAsync(h_total, c->d_total.p, sizeof(TOut), cudaMemcpyDeviceToHost, c->st));
    SPK(cudaStreamSynchronize(c->st));
  }
  return 0;
}

// stable LSD radix sort of n keys on their low `bits` bits; the sorted keys end in *keys (buffers are swapped)
int radix_sort(frs_split* c, u64** keys, u64** tmp, i64 n, int bits) {
  if (n <= 1) return 0;
  const int n_tiles = (int)nblk(n, SP_SORT_TILE);
  SENS(d_hist, (size_t)256 * n_tiles * 4);
  SENS(d_hist_scan, (size_t)256 * n_tiles * 4);
  for (int shift = 0; shift < bits; shift += 8) {
    k_sp_hist<<<n_tiles, SP_SORT_THREADS, 0, c->st>>>(*keys, n, shift, n_tiles, c->d_hist.as<int>());
    c->launches += 1;
    int r = scan_excl<int, int>(c, c->d_hist.as<int>(), (i64)256 * n_tiles, c->d_hist_scan.as<int>(), (int*)nullptr);
    if (r) return r;
    k_sp_scatter<<<n_tiles, SP_SORT_THREADS, 0, c->st>>>(*keys, n, shift, n_tiles, c->d_hist_scan.as<int>(), *tmp);
    c->launches += 1;
    std::swap(*keys, *tmp);
  }
  return 0;
}

// sorted keys -> the distinct ones, in order; returns their number
int unique_keys(frs_split* c, const u64* sorted, i64 n, SBuf& out, int* n_out) {
  *n_out = 0;
  if (n <= 0) return 0;
  SENS(d_flag, (size_t)n * 4);
  SENS(d_pos, (size_t)n * 4);
  k_sp_unique_flags<<<nblk(n, 256), 256, 0, c->st>>>(n, sorted, c->d_flag.as<int>());
  c->launches += 1;
  int r = scan_excl<int, int>(c, c->d_flag.as<int>(), n, c->d_pos.as<int>(), n_out);
  if (r) return r;
  r = ensure(c, out, (size_t)*n_out * 8);
  if (r) return r;
  k_sp_compact<<<nblk(n, 256), 256, 0, c->st>>>(n, sorted, c->d_flag.as<int>(), c->d_pos.as<int>(), out.as<u64>());
  c->launches += 1;
  return 0;
}

}  // namespace

extern "C" {

int frs_split_create(int device, frs_split** out) {
  if (!out) return FRS_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return FRS_ERR_CUDA;  // no device: no CPU fallback
  }
  frs_split* c = new frs_split();
  c->device = device;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return FRS_ERR_CUDA;
  }
  for (auto& e : c->ev) cudaEventCreate(&e);
  *out = c;
  return 0;
}

void frs_split_destroy(frs_split* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->st);
  for (SBuf* b : c->all)
    if (b->p) cudaFree(b->p);
  for (auto& e : c->ev)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(c->st);
  delete c;
}

const char* frs_split_last_error(frs_split* c) { return c ? c->err : "null context"; }

int frs_split_run(frs_split* c, const frs_split_batch* b, frs_split_sizes* sizes) {
  if (!c || !b || !sizes) return FRS_ERR_ARG;
  c->ran = false;
  c->launches = 0;
  SPK(cudaSetDevice(c->device));
  const int G = b->n_groups, N = b->n_reads;
  if (G < 0 || N < 0) return fail(c, FRS_ERR_ARG, "negative counts");
  if (G > 0 && !b->group_read_off) return fail(c, FRS_ERR_ARG, "null group table");
  if (N > 0 && (!b->read_iv_off || G == 0)) return fail(c, FRS_ERR_ARG, "null read table");
  if (G > 0 && (b->group_read_off[0] != 0 || b->group_read_off[G] != N)) return fail(c, FRS_ERR_ARG, "group_read_off does not span the reads");
  for (int g = 0; g < G; ++g)
    if (b->group_read_off[g + 1] < b->group_read_off[g]) return fail(c, FRS_ERR_ARG, "group_read_off decreases");
  const i64 n_iv = N > 0 ? b->read_iv_off[N] : 0;
  if (N > 0 && b->read_iv_off[0] != 0) return fail(c, FRS_ERR_ARG, "read_iv_off[0] != 0");
  for (int r = 0; r < N; ++r)
    if (b->read_iv_off[r + 1] <= b->read_iv_off[r]) return fail(c, FRS_ERR_ARG, "read %d has no alignment interval (read_sam keeps none such, :223-227)", r);
  if (n_iv > 0 && (!b->iv_s || !b->iv_e)) return fail(c, FRS_ERR_ARG, "null interval arrays");
  if (n_iv >= (1ll << 30)) return fail(c, FRS_ERR_LIMIT, "more than 2^30 alignment intervals in one batch");
  const int max_iv = b->max_intervals > 0 ? b->max_intervals : 100, max_rd = b->max_reads > 0 ? b->max_reads : 1500;
  c->n_groups = G;
  c->h_group_tint_off.assign(G + 1, 0);
  c->h_tint_iv_off.assign(1, 0);
  c->h_tint_rid_off.assign(1, 0);
  c->h_tint_iv_s.clear();
  c->h_tint_iv_e.clear();
  c->h_tint_rids.clear();
  frs_split_sizes& z = c->sizes;
  memset(&z, 0, sizeof z);
  cudaEventRecord(c->ev[0], c->st);
  int n_simple = 0;
  std::vector<int> root, read_comp, simple_end;
  std::vector<u64> simple_key;
  if (n_iv > 0) {
    SUP(d_group_read_off, b->group_read_off, (size_t)(G + 1) * 4);
    SUP(d_read_iv_off, b->read_iv_off, (size_t)(N + 1) * 4);
    SUP(d_iv_s, b->iv_s, (size_t)n_iv * 4);
    SUP(d_iv_e, b->iv_e, (size_t)n_iv * 4);
    SENS(d_read_group, (size_t)N * 4);
    SENS(d_iv_read, (size_t)n_iv * 4);
    SENS(d_keys_a, (size_t)n_iv * 16);
    SENS(d_keys_b, (size_t)n_iv * 16);
    SENS(d_err, 16);
    SPK(cudaMemsetAsync(c->d_err.p, 0, 16, c->st));
    k_sp_read_owner<<<nblk(N, 256), 256, 0, c->st>>>(N, G, c->d_group_read_off.as<int>(), c->d_read_iv_off.as<int>(),
                                                    c->d_read_group.as<int>(), c->d_iv_read.as<int>());
    k_sp_events<<<nblk(n_iv, 256), 256, 0, c->st>>>(n_iv, c->d_iv_read.as<int>(), c->d_read_group.as<int>(), c->d_iv_s.as<int>(),
                                                   c->d_iv_e.as<int>(), c->d_keys_a.as<u64>(), c->d_err.as<int>());
    c->launches += 2;
    // ---- the sweep (:296-321) ----
    const i64 n_ev = 2 * n_iv;
    u64* keys = c->d_keys_a.as<u64>();
    u64* tmp = c->d_keys_b.as<u64>();
    { int r = radix_sort(c, &keys, &tmp, n_ev, 33 + bits_of((unsigned long long)std::max(G - 1, 1))); if (r) return r; }
    SENS(d_delta, (size_t)n_ev * 4);
    SENS(d_open, (size_t)n_ev * 4);
    SENS(d_first, (size_t)n_ev * 4);
    SENS(d_sid_before, (size_t)n_ev * 4);
    k_sp_delta<<<nblk(n_ev, 256), 256, 0, c->st>>>(n_ev, keys, c->d_delta.as<int>());
    c->launches += 1;
    int open_end = 0;
    { int r = scan_excl<int, int>(c, c->d_delta.as<int>(), n_ev, c->d_open.as<int>(), &open_end); if (r) return r; }
    k_sp_first<<<nblk(n_ev, 256), 256, 0, c->st>>>(n_ev, keys, c->d_open.as<int>(), c->d_first.as<int>());
    c->launches += 1;
    { int r = scan_excl<int, int>(c, c->d_first.as<int>(), n_ev, c->d_sid_before.as<int>(), &n_simple); if (r) return r; }
    int h_err = 0;
    SPK(cudaMemcpyAsync(&h_err, c->d_err.p, 4, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaStreamSynchronize(c->st));
    if (h_err) return fail(c, FRS_ERR_ARG, "an alignment interval has a negative start or ends before it starts");
    SENS(d_simple_key, (size_t)n_simple * 8);
    SENS(d_simple_end, (size_t)n_simple * 4);
    k_sp_simple<<<nblk(n_ev, 256), 256, 0, c->st>>>(n_ev, keys, c->d_open.as<int>(), c->d_first.as<int>(), c->d_sid_before.as<int>(),
                                                   c->d_simple_key.as<u64>(), c->d_simple_end.as<int>());
    // ---- groups of simple intervals joined by reads (:322-343) ----
    SENS(d_iv_sid, (size_t)n_iv * 4);
    SENS(d_parent, (size_t)n_simple * 4);
    SENS(d_root, (size_t)n_simple * 4);
    SENS(d_read_comp, (size_t)N * 4);
    k_sp_iv_simple<<<nblk(n_iv, 256), 256, 0, c->st>>>(n_iv, n_simple, c->d_iv_read.as<int>(), c->d_read_group.as<int>(),
                                                      c->d_iv_s.as<int>(), c->d_simple_key.as<u64>(), c->d_iv_sid.as<int>());
    k_sp_iota<<<nblk(n_simple, 256), 256, 0, c->st>>>(n_simple, c->d_parent.as<int>());
    k_sp_join_reads<<<nblk(N, 256), 256, 0, c->st>>>(N, c->d_read_iv_off.as<int>(), c->d_iv_sid.as<int>(), c->d_parent.as<int>());
    k_sp_roots<<<nblk(n_simple, 256), 256, 0, c->st>>>(n_simple, c->d_parent.as<int>(), c->d_root.as<int>());
    k_sp_read_comp<<<nblk(N, 256), 256, 0, c->st>>>(N, c->d_read_iv_off.as<int>(), c->d_iv_sid.as<int>(), c->d_root.as<int>(),
                                                   c->d_read_comp.as<int>());
    c->launches += 6;
    root.resize(n_simple);
    read_comp.resize(N);
    simple_end.resize(n_simple);
    simple_key.resize(n_simple);
    SPK(cudaMemcpyAsync(root.data(), c->d_root.p, (size_t)n_simple * 4, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaMemcpyAsync(read_comp.data(), c->d_read_comp.p, (size_t)N * 4, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaMemcpyAsync(simple_end.data(), c->d_simple_end.p, (size_t)n_simple * 4, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaMemcpyAsync(simple_key.data(), c->d_simple_key.p, (size_t)n_simple * 8, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaStreamSynchronize(c->st));
  }
  z.n_simple = n_simple;
  // ---- host: the groups in the order of their smallest simple interval; lists by counting sort ----
  std::vector<int> comp_iv_off(n_simple + 1, 0), comp_rd_off(n_simple + 1, 0);
  for (int s = 0; s < n_simple; ++s) comp_iv_off[root[s] + 1]++;
  for (int r = 0; r < N; ++r) comp_rd_off[read_comp[r] + 1]++;
  for (int s = 0; s < n_simple; ++s) {
    comp_iv_off[s + 1] += comp_iv_off[s];
    comp_rd_off[s + 1] += comp_rd_off[s];
  }
  std::vector<int> comp_ivs(n_simple), comp_rds(N);
  {
    std::vector<int> cur(comp_iv_off.begin(), comp_iv_off.end() - 1);
    for (int s = 0; s < n_simple; ++s) comp_ivs[cur[root[s]]++] = s;
    std::vector<int> cur2(comp_rd_off.begin(), comp_rd_off.end() - 1);
    for (int r = 0; r < N; ++r) comp_rds[cur2[read_comp[r]]++] = r;
  }
  // big groups (:357): nodes = their simple intervals, numbered group by group
  std::vector<int> node_of(n_simple, -1), node_sid, big_reads, big_comp_of_root(n_simple, -1);
  std::vector<i64> junc_off(1, 0), ivk_off(1, 0);
  int n_big = 0;
  for (int s = 0; s < n_simple; ++s) {
    if (root[s] != s) continue;
    const int ni = comp_iv_off[s + 1] - comp_iv_off[s], nr = comp_rd_off[s + 1] - comp_rd_off[s];
    if (nr < 3) continue;  // (:345)
    if (ni < max_iv && nr < max_rd) continue;
    big_comp_of_root[s] = n_big++;
    for (int k = comp_iv_off[s]; k < comp_iv_off[s + 1]; ++k) {
      node_of[comp_ivs[k]] = (int)node_sid.size();
      node_sid.push_back(comp_ivs[k]);
    }
    for (int k = comp_rd_off[s]; k < comp_rd_off[s + 1]; ++k) {
      const int r = comp_rds[k];
      const int n = b->read_iv_off[r + 1] - b->read_iv_off[r];
      big_reads.push_back(r);
      junc_off.push_back(junc_off.back() + (n - 1));
      ivk_off.push_back(ivk_off.back() + n);
    }
  }
  z.n_big = n_big;
  const int n_nodes = (int)node_sid.size(), n_big_reads = (int)big_reads.size();
  std::vector<u64> pairs, civ;
  std::vector<int> comp_reads;
  if (n_big > 0) {
    SUP(d_node_of, node_of.data(), (size_t)n_simple * 4);
    SUP(d_big_reads, big_reads.data(), (size_t)n_big_reads * 4);
    SUP(d_off_a, junc_off.data(), (size_t)(n_big_reads + 1) * 8);
    SUP(d_off_b, ivk_off.data(), (size_t)(n_big_reads + 1) * 8);
    const int node_bits = bits_of((unsigned long long)std::max(n_nodes - 1, 1));
    // junction support (:265-278)
    SENS(d_node_root, (size_t)n_nodes * 4);
    SENS(d_parent, (size_t)std::max(n_nodes, n_simple) * 4);
    k_sp_iota<<<nblk(n_nodes, 256), 256, 0, c->st>>>(n_nodes, c->d_parent.as<int>());
    c->launches += 1;
    const i64 n_j = junc_off.back();
    if (n_j > 0) {
      SENS(d_keys_a, (size_t)n_j * 8);
      SENS(d_keys_b, (size_t)n_j * 8);
      u64* keys = c->d_keys_a.as<u64>();
      u64* tmp = c->d_keys_b.as<u64>();
      k_sp_junctions<<<nblk(n_big_reads, 128), 128, 0, c->st>>>(n_big_reads, c->d_big_reads.as<int>(), c->d_read_iv_off.as<int>(),
                                                              c->d_iv_sid.as<int>(), c->d_node_of.as<int>(), c->d_off_a.as<i64>(), keys);
      c->launches += 1;
      { int r = radix_sort(c, &keys, &tmp, n_j, 32 + node_bits); if (r) return r; }
      k_sp_join_edges<<<nblk(n_j, 256), 256, 0, c->st>>>(n_j, keys, c->d_parent.as<int>());
      c->launches += 1;
    }
    k_sp_roots<<<nblk(n_nodes, 256), 256, 0, c->st>>>(n_nodes, c->d_parent.as<int>(), c->d_node_root.as<int>());
    c->launches += 1;
    // reads of every component (:279-282)
    const i64 n_k = ivk_off.back();
    SENS(d_keys_a, (size_t)n_k * 8);
    SENS(d_keys_b, (size_t)n_k * 8);
    u64* keys = c->d_keys_a.as<u64>();
    u64* tmp = c->d_keys_b.as<u64>();
    k_sp_comp_read_keys<<<nblk(n_big_reads, 128), 128, 0, c->st>>>(n_big_reads, c->d_big_reads.as<int>(), c->d_read_iv_off.as<int>(),
                                                                 c->d_iv_sid.as<int>(), c->d_node_of.as<int>(),
                                                                 c->d_node_root.as<int>(), c->d_off_b.as<i64>(), keys);
    c->launches += 1;
    { int r = radix_sort(c, &keys, &tmp, n_k, 32 + node_bits); if (r) return r; }
    int n_pairs = 0;
    { int r = unique_keys(c, keys, n_k, c->d_pairs, &n_pairs); if (r) return r; }
    SENS(d_comp_reads, (size_t)n_nodes * 4);
    SPK(cudaMemsetAsync(c->d_comp_reads.p, 0, (size_t)n_nodes * 4, c->st));
    k_sp_count_comp<<<nblk(n_pairs, 256), 256, 0, c->st>>>(n_pairs, c->d_pairs.as<u64>(), c->d_comp_reads.as<int>());
    c->launches += 1;
    // intervals of every kept component (:283-291)
    SENS(d_cnt64, (size_t)n_pairs * 8);
    SENS(d_off64, (size_t)(n_pairs + 1) * 8);
    k_sp_comp_iv_keys<0><<<nblk(n_pairs, 128), 128, 0, c->st>>>(n_pairs, c->d_pairs.as<u64>(), c->d_comp_reads.as<int>(),
                                                              c->d_big_reads.as<int>(), c->d_read_iv_off.as<int>(),
                                                              c->d_iv_sid.as<int>(), c->d_node_of.as<int>(), c->d_cnt64.as<i64>(),
                                                              nullptr, nullptr);
    c->launches += 1;
    i64 n_civ_keys = 0;
    { int r = scan_excl<i64, i64>(c, c->d_cnt64.as<i64>(), n_pairs, c->d_off64.as<i64>(), &n_civ_keys); if (r) return r; }
    int n_civ = 0;
    if (n_civ_keys > 0) {
      SENS(d_keys_a, (size_t)n_civ_keys * 8);
      SENS(d_keys_b, (size_t)n_civ_keys * 8);
      keys = c->d_keys_a.as<u64>();
      tmp = c->d_keys_b.as<u64>();
      k_sp_comp_iv_keys<1><<<nblk(n_pairs, 128), 128, 0, c->st>>>(n_pairs, c->d_pairs.as<u64>(), c->d_comp_reads.as<int>(),
                                                                c->d_big_reads.as<int>(), c->d_read_iv_off.as<int>(),
                                                                c->d_iv_sid.as<int>(), c->d_node_of.as<int>(), nullptr,
                                                                c->d_off64.as<i64>(), keys);
      c->launches += 1;
      { int r = radix_sort(c, &keys, &tmp, n_civ_keys, 32 + node_bits); if (r) return r; }
      { int r = unique_keys(c, keys, n_civ_keys, c->d_civ, &n_civ); if (r) return r; }
    }
    pairs.resize(n_pairs);
    civ.resize(n_civ);
    comp_reads.resize(n_nodes);
    SPK(cudaMemcpyAsync(pairs.data(), c->d_pairs.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, c->st));
    if (n_civ) SPK(cudaMemcpyAsync(civ.data(), c->d_civ.p, (size_t)n_civ * 8, cudaMemcpyDeviceToHost, c->st));
    SPK(cudaMemcpyAsync(comp_reads.data(), c->d_comp_reads.p, (size_t)n_nodes * 4, cudaMemcpyDeviceToHost, c->st));
  }
  cudaEventRecord(c->ev[1], c->st);
  SPK(cudaStreamSynchronize(c->st));
  SPK(cudaGetLastError());
  cudaEventElapsedTime(&c->ms, c->ev[0], c->ev[1]);
  // ---- assemble the tints in the reference's order ----
  auto emit_iv = [&](int sid) {
    c->h_tint_iv_s.push_back((int)(simple_key[sid] & 0xffffffffu));
    c->h_tint_iv_e.push_back(simple_end[sid]);
  };
  auto close_tint = [&](int g) {
    c->h_tint_iv_off.push_back((int)c->h_tint_iv_s.size());
    c->h_tint_rid_off.push_back((int)c->h_tint_rids.size());
    c->h_group_tint_off[g + 1]++;
  };
  size_t pi = 0, ci = 0;  // cursors into pairs / civ (both sorted by component = root node)
  for (int s = 0; s < n_simple; ++s) {
    if (root[s] != s) continue;
    const int nr = comp_rd_off[s + 1] - comp_rd_off[s];
    if (nr < 3) continue;
    const int g = (int)(simple_key[s] >> 32);
    const int g0 = b->group_read_off[g];
    if (big_comp_of_root[s] < 0) {
      for (int k = comp_iv_off[s]; k < comp_iv_off[s + 1]; ++k) emit_iv(comp_ivs[k]);
      for (int k = comp_rd_off[s]; k < comp_rd_off[s + 1]; ++k) c->h_tint_rids.push_back(comp_rds[k] - g0);
      close_tint(g);
      continue;
    }
    // break_tint: the components whose root node lies in this group's node range, in order
    const int node_lo = node_of[comp_ivs[comp_iv_off[s]]], node_hi = node_lo + (comp_iv_off[s + 1] - comp_iv_off[s]);
    while (pi < pairs.size() && (int)(pairs[pi] >> 32) < node_hi) {
      const int comp = (int)(pairs[pi] >> 32);
      if (comp < node_lo) return fail(c, FRS_ERR_ASSERT, "internal: component outside its group");
      size_t pe = pi;
      while (pe < pairs.size() && (int)(pairs[pe] >> 32) == comp) ++pe;
      if (comp_reads[comp] > 2) {
        while (ci < civ.size() && (int)(civ[ci] >> 32) < comp) ++ci;
        for (; ci < civ.size() && (int)(civ[ci] >> 32) == comp; ++ci) emit_iv(node_sid[(int)(civ[ci] & 0xffffffffu)]);
        for (size_t k = pi; k < pe; ++k) c->h_tint_rids.push_back(big_reads[(int)(pairs[k] & 0xffffffffu)] - g0);
        close_tint(g);
      }
      pi = pe;
    }
  }
  for (int g = 0; g < G; ++g) c->h_group_tint_off[g + 1] += c->h_group_tint_off[g];
  z.n_tints = (int64_t)c->h_tint_iv_off.size() - 1;
  z.n_tint_ivs = (int64_t)c->h_tint_iv_s.size();
  z.n_tint_rids = (int64_t)c->h_tint_rids.size();
  z.launches = c->launches;
  *sizes = z;
  c->ran = true;
  return 0;
}

int frs_split_fetch(frs_split* c, const frs_split_result* o) {
  if (!c || !o) return FRS_ERR_ARG;
  if (!c->ran) return fail(c, FRS_ERR_STATE, "frs_split_fetch before a successful frs_split_run");
  if (o->group_tint_off) memcpy(o->group_tint_off, c->h_group_tint_off.data(), c->h_group_tint_off.size() * 4);
  if (o->tint_iv_off) memcpy(o->tint_iv_off, c->h_tint_iv_off.data(), c->h_tint_iv_off.size() * 4);
  if (o->tint_iv_s && !c->h_tint_iv_s.empty()) memcpy(o->tint_iv_s, c->h_tint_iv_s.data(), c->h_tint_iv_s.size() * 4);
  if (o->tint_iv_e && !c->h_tint_iv_e.empty()) memcpy(o->tint_iv_e, c->h_tint_iv_e.data(), c->h_tint_iv_e.size() * 4);
  if (o->tint_rid_off) memcpy(o->tint_rid_off, c->h_tint_rid_off.data(), c->h_tint_rid_off.size() * 4);
  if (o->tint_rids && !c->h_tint_rids.empty()) memcpy(o->tint_rids, c->h_tint_rids.data(), c->h_tint_rids.size() * 4);
  return 0;
}

float frs_split_last_ms(frs_split* c) { return c ? c->ms : 0.f; }

}  // extern "C"
