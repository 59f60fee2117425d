// kernels_split.cuh -- SURVEY.md 8f-4: tint construction of freddie_split.py (get_transcriptional_intervals
// :295-364, break_tint :246-293) for a batch of read groups.
//   * the union of all alignment intervals as a sorted sweep over 2 x intervals EVENTS (key = group | position | end?),
//     starts before ends at the same position so that touching intervals merge like `s > end` (:303) says;
//   * simple intervals joined through reads = union-find with the smallest member as the root (the BFS of :325-337
//     discovers groups in that order);
//   * break_tint: an alignment interval lies inside ONE simple interval, so pos_to_intrv[...] of its start, of its
//     last base and of the next interval's start (:263, :268-269) are the simple intervals of the two alignment
//     intervals -- no position table; junction support by sorting (u, v) pairs, run lengths >= 2 are the edges;
//     the reads / intervals of a component by sorting and uniquing (component, read) and (component, interval).
// Sorting: a plain stable LSD radix sort of 64-bit keys (8 bits per pass, only the passes the keys need).
#pragma once
#include <stdio.h>

#include "scan.cuh"

#define SP_SORT_THREADS 256
#define SP_SORT_ITEMS 8
#define SP_SORT_TILE (SP_SORT_THREADS * SP_SORT_ITEMS)

// ---- radix sort pass: per-tile digit histograms, hist[digit * n_tiles + tile] ----
__global__ void __launch_bounds__(SP_SORT_THREADS) k_sp_hist(const u64* __restrict__ keys, i64 n, int shift, int n_tiles,
                                                            int* __restrict__ hist) {
  __shared__ int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const i64 base = (i64)blockIdx.x * SP_SORT_TILE;
  for (int k = 0; k < SP_SORT_ITEMS; ++k) {
    const i64 i = base + (i64)k * SP_SORT_THREADS + threadIdx.x;
    if (i < n) atomicAdd(&h[(int)((keys[i] >> shift) & 255u)], 1);
  }
  __syncthreads();
  hist[(i64)threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];
}

// stable scatter of one tile: rank of a key = keys of the same digit in earlier tiles (scanned histogram) + in
// earlier steps of this tile + in earlier warps of this step + in lower lanes of its warp
__global__ void __launch_bounds__(SP_SORT_THREADS) k_sp_scatter(const u64* __restrict__ keys, i64 n, int shift, int n_tiles,
                                                               const int* __restrict__ hist_scan, u64* __restrict__ out) {
  __shared__ int run[256];                          // keys of digit d placed so far by this tile
  __shared__ int wcnt[SP_SORT_THREADS / 32][256];   // per warp: keys of digit d in this step
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  run[threadIdx.x] = hist_scan[(i64)threadIdx.x * n_tiles + blockIdx.x];
  const i64 base = (i64)blockIdx.x * SP_SORT_TILE;
  for (int k = 0; k < SP_SORT_ITEMS; ++k) {
    for (int w = 0; w < SP_SORT_THREADS / 32; ++w) wcnt[w][threadIdx.x] = 0;
    __syncthreads();
    const i64 i = base + (i64)k * SP_SORT_THREADS + threadIdx.x;
    const bool have = i < n;
    const u64 key = have ? keys[i] : 0;
    const int d = have ? (int)((key >> shift) & 255u) : 256 + lane;  // absent keys match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    const int below = __popc(peers & ((1u << lane) - 1u));
    if (have && below == 0) wcnt[warp][d] = __popc(peers);
    __syncthreads();
    int before = 0;
    if (have) {
      for (int w = 0; w < warp; ++w) before += wcnt[w][d];
      out[run[d] + before + below] = key;
    }
    __syncthreads();
    {
      int tot = 0;
      for (int w = 0; w < SP_SORT_THREADS / 32; ++w) tot += wcnt[w][threadIdx.x];
      run[threadIdx.x] += tot;
    }
    __syncthreads();
  }
}

// ---- events of the sweep ----
__global__ void k_sp_read_owner(int N, int G, const int* __restrict__ group_read_off, const int* __restrict__ read_iv_off,
                                int* __restrict__ read_group, int* __restrict__ iv_read) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  int lo = 0, hi = G;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (group_read_off[mid] <= r) lo = mid; else hi = mid;
  }
  read_group[r] = lo;
  for (int k = read_iv_off[r]; k < read_iv_off[r + 1]; ++k) iv_read[k] = r;
}
__global__ void k_sp_events(i64 n_iv, const int* __restrict__ iv_read, const int* __restrict__ read_group,
                            const int* __restrict__ iv_s, const int* __restrict__ iv_e, u64* __restrict__ keys, int* err) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_iv) return;
  const u64 g = (u64)read_group[iv_read[k]];
  const int s = iv_s[k], e = iv_e[k];
  if (s < 0 || e < s) atomicCAS(err, 0, 1);
  keys[2 * k] = (g << 33) | ((u64)(u32)s << 1);
  keys[2 * k + 1] = (g << 33) | ((u64)(u32)e << 1) | 1u;
}
__global__ void k_sp_delta(i64 n, const u64* __restrict__ keys, int* __restrict__ delta) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) delta[i] = (keys[i] & 1u) ? -1 : 1;
}
// open intervals BEFORE event i (exclusive scan of the deltas): a start with none open begins a simple interval
__global__ void k_sp_first(i64 n, const u64* __restrict__ keys, const int* __restrict__ open_before, int* __restrict__ first) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) first[i] = (!(keys[i] & 1u) && open_before[i] == 0) ? 1 : 0;
}
// simple interval table: key (group << 32 | start) and end; sid_before = exclusive scan of `first`
__global__ void k_sp_simple(i64 n, const u64* __restrict__ keys, const int* __restrict__ open_before,
                            const int* __restrict__ first, const int* __restrict__ sid_before, u64* __restrict__ simple_key,
                            int* __restrict__ simple_end) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = keys[i];
  const u64 g = key >> 33;
  const u32 pos = (u32)((key >> 1) & 0xffffffffu);
  if (first[i]) simple_key[sid_before[i]] = (g << 32) | pos;
  if ((key & 1u) && open_before[i] == 1) simple_end[sid_before[i] - 1] = (int)pos;  // the last open interval closes
}
// simple interval of every alignment interval: the last one of its group that starts at or before it
__global__ void k_sp_iv_simple(i64 n_iv, int n_simple, const int* __restrict__ iv_read, const int* __restrict__ read_group,
                               const int* __restrict__ iv_s, const u64* __restrict__ simple_key, int* __restrict__ iv_sid) {
  const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_iv) return;
  const u64 want = ((u64)read_group[iv_read[k]] << 32) | (u32)iv_s[k];
  int lo = 0, hi = n_simple;  // simple_key[lo] <= want < simple_key[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (simple_key[mid] <= want) lo = mid; else hi = mid;
  }
  iv_sid[k] = lo;
}

// ---- union-find, smaller root wins ----
__device__ __forceinline__ int sp_find(int* parent, int x) {
  int p = *(volatile int*)&parent[x];
  while (p != x) {
    x = p;
    p = *(volatile int*)&parent[x];
  }
  return x;
}
__device__ __forceinline__ void sp_union(int* parent, int a, int b) {
  while (true) {
    a = sp_find(parent, a);
    b = sp_find(parent, b);
    if (a == b) return;
    if (a < b) { const int x = a; a = b; b = x; }
    if (atomicCAS(&parent[a], a, b) == a) return;
  }
}
__global__ void k_sp_iota(int n, int* __restrict__ p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}
// (:325-337) a read joins the simple intervals of its consecutive alignment intervals
__global__ void k_sp_join_reads(int N, const int* __restrict__ read_iv_off, const int* __restrict__ iv_sid, int* parent) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  for (int k = read_iv_off[r]; k + 1 < read_iv_off[r + 1]; ++k)
    if (iv_sid[k] != iv_sid[k + 1]) sp_union(parent, iv_sid[k], iv_sid[k + 1]);
}
__global__ void k_sp_roots(int n, int* parent, int* __restrict__ root) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) root[i] = sp_find(parent, i);
}
__global__ void k_sp_read_comp(int N, const int* __restrict__ read_iv_off, const int* __restrict__ iv_sid,
                               const int* __restrict__ root, int* __restrict__ read_comp) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  read_comp[r] = read_iv_off[r] < read_iv_off[r + 1] ? root[iv_sid[read_iv_off[r]]] : -1;
}

// ---- break_tint (:246-293) over the reads of the big groups ----
// node = index of a simple interval among the intervals of the big groups (node_of[sid], -1 elsewhere)
// junction keys (u << 32 | v) of consecutive alignment intervals (:265-276); 0xffff... for intervals outside big groups
__global__ void k_sp_junctions(int n_big_reads, const int* __restrict__ big_reads, const int* __restrict__ read_iv_off,
                               const int* __restrict__ iv_sid, const int* __restrict__ node_of, const i64* __restrict__ out_off,
                               u64* __restrict__ keys) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_big_reads) return;
  const int r = big_reads[q];
  i64 o = out_off[q];
  for (int k = read_iv_off[r]; k + 1 < read_iv_off[r + 1]; ++k)
    keys[o++] = ((u64)(u32)node_of[iv_sid[k]] << 32) | (u32)node_of[iv_sid[k + 1]];
}
// edges = keys that occur at least twice (:278): the first key of a run of length >= 2 joins its two nodes
__global__ void k_sp_join_edges(i64 n, const u64* __restrict__ sorted, int* parent) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const u64 key = sorted[i];
  if (i > 0 && sorted[i - 1] == key) return;         // not the first of its run
  if (i + 1 >= n || sorted[i + 1] != key) return;    // support 1
  sp_union(parent, (int)(key >> 32), (int)(key & 0xffffffffu));
}
// (component, read) keys: one per alignment interval of a big read (:261-264: the read STARTS an alignment there)
__global__ void k_sp_comp_read_keys(int n_big_reads, const int* __restrict__ big_reads, const int* __restrict__ read_iv_off,
                                    const int* __restrict__ iv_sid, const int* __restrict__ node_of,
                                    const int* __restrict__ node_root, const i64* __restrict__ out_off, u64* __restrict__ keys) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_big_reads) return;
  const int r = big_reads[q];
  i64 o = out_off[q];
  for (int k = read_iv_off[r]; k < read_iv_off[r + 1]; ++k) keys[o++] = ((u64)(u32)node_root[node_of[iv_sid[k]]] << 32) | (u32)q;
}
// flags the first key of every run (sorted keys)
__global__ void k_sp_unique_flags(i64 n, const u64* __restrict__ sorted, int* __restrict__ flag) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || sorted[i - 1] != sorted[i]) ? 1 : 0;
}
__global__ void k_sp_compact(i64 n, const u64* __restrict__ sorted, const int* __restrict__ flag, const int* __restrict__ pos,
                             u64* __restrict__ out) {
  const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && flag[i]) out[pos[i]] = sorted[i];
}
// reads per component: pairs are sorted by (component, read); count per component root node
__global__ void k_sp_count_comp(int n_pairs, const u64* __restrict__ pairs, int* __restrict__ comp_reads) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pairs) atomicAdd(&comp_reads[(int)(pairs[i] >> 32)], 1);
}
// (component, interval) keys of the kept components: every interval in which one of the component's reads starts
// an alignment (:284-287); sizes first (pass 0), then the keys at the scanned offsets
template <int WRITE>
__global__ void k_sp_comp_iv_keys(int n_pairs, const u64* __restrict__ pairs, const int* __restrict__ comp_reads,
                                  const int* __restrict__ big_reads, const int* __restrict__ read_iv_off,
                                  const int* __restrict__ iv_sid, const int* __restrict__ node_of, i64* __restrict__ cnt,
                                  const i64* __restrict__ off, u64* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pairs) return;
  const int c = (int)(pairs[i] >> 32), r = big_reads[(int)(pairs[i] & 0xffffffffu)];
  const bool kept = comp_reads[c] > 2;  // (:283)
  const int n = kept ? read_iv_off[r + 1] - read_iv_off[r] : 0;
  if (!WRITE) {
    cnt[i] = n;
    return;
  }
  i64 o = off[i];
  for (int k = 0; k < n; ++k) keys[o++] = ((u64)(u32)c << 32) | (u32)node_of[iv_sid[read_iv_off[r] + k]];
}
