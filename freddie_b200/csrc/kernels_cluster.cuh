// kernels_cluster.cuh -- SURVEY.md 8f-3: the Gurobi-free front of freddie_cluster.py on bit-packed rows.
//   read_segment's rep merge (freddie_cluster.py:154-164), preprocess_ilp (:277-328), partition_reads (:198-274).
// Layouts (all per tint t, M = segments, W = ceil(M/32)):
//   rowbits  [rowword_off[t] + r*W + w]       digit rows as bits (bit = digit '1'), row-major (dedupe, find_segment_read)
//   sbits    [sb_off[t] + w*S + s]            the rows of the S structures, WORD-major: the pair test reads word w of
//                                             32 consecutive structures as one coalesced load
//   adj      [adj_off[t] + s*SW + jw]         adjacency of the compatibility graph, SW = ceil(S/32) words per row
#pragma once
#include <stdio.h>

#include "scan.cuh"

#define CP_EMPTY (-1)
enum { CPERR_DIGIT = 1, CPERR_GAP = 2 };

__device__ __forceinline__ u64 cp_mix(u64 h, u64 v) {
  h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
  h *= 0xff51afd7ed558ccdULL;
  h ^= h >> 32;
  return h;
}

// owner of item i in a CSR offset table off[0..n]
__device__ __forceinline__ int cp_owner(const int* __restrict__ off, int n, int i) {
  int lo = 0, hi = n;  // off[lo] <= i < off[hi]
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid;
  }
  return lo;
}

// First-seen dedupe: every item probes the table region of its tint; a slot holds the SMALLEST item index of one
// key class (equal keys only ever replace each other, by atomicMin), so that "first seen in index order" -- the
// insertion order of the reference's dicts -- needs no sort.  Returns the slot; the class representative is read
// from it by a later launch.
template <class Eq>
__device__ __forceinline__ int cp_insert(int* tab, int cap, u64 hash, int item, Eq eq) {
  int slot = (int)(hash & (u64)(cap - 1));
  while (true) {
    int cur = *(volatile int*)&tab[slot];
    if (cur == CP_EMPTY) {
      int old = atomicCAS(&tab[slot], CP_EMPTY, item);
      if (old == CP_EMPTY) return slot;
      cur = old;
    }
    if (cur == item || eq(cur)) {
      atomicMin(&tab[slot], item);
      return slot;
    }
    slot = (slot + 1) & (cap - 1);
  }
}

// ---- digit rows -> bits, first / last 1 (find_segment_read :175-184), hash.  One warp per row. ----
__global__ void k_cp_row_bits(int n_rows, int T, const int* __restrict__ tint_row_off, const int* __restrict__ tint_seg_n,
                              const i64* __restrict__ tint_digit_off, const i64* __restrict__ rowword_off,
                              const u8* __restrict__ digits, u32* __restrict__ rowbits, int* __restrict__ row_f,
                              int* __restrict__ row_l, u64* __restrict__ row_hash, int* __restrict__ row_tint, int* err) {
  const int gr = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gr >= n_rows) return;
  const int t = cp_owner(tint_row_off, T, gr);
  const int r = gr - tint_row_off[t], M = tint_seg_n[t], W = (M + 31) >> 5;
  const u8* src = digits + tint_digit_off[t] + (i64)r * M;
  u32* dst = rowbits + rowword_off[t] + (i64)r * W;
  int f = -1, l = M - 1;
  bool any = false;
  u64 h = 0x243f6a8885a308d3ULL;
  for (int w = 0; w < W; ++w) {
    const int j = w * 32 + lane;
    int d = 0;
    if (j < M) {
      d = (int)src[j] - '0';
      if (d < 0 || d > 2) atomicCAS(err, 0, CPERR_DIGIT);
    }
    const u32 word = __ballot_sync(0xffffffffu, d == 1);
    if (word) {
      if (!any) f = w * 32 + __ffs(word) - 1;
      l = w * 32 + 31 - __clz(word);
      any = true;
    }
    h = cp_mix(h, word);
    if (lane == 0) dst[w] = word;
  }
  if (lane == 0) {
    row_f[gr] = f;
    row_l[gr] = l;
    row_hash[gr] = h;
    row_tint[gr] = t;
  }
}

__global__ void k_cp_row_insert(int n_rows, const int* __restrict__ row_tint, const int* __restrict__ tint_row_off,
                                const int* __restrict__ tint_seg_n, const i64* __restrict__ rowword_off,
                                const u32* __restrict__ rowbits, const u64* __restrict__ row_hash,
                                const i64* __restrict__ tab_off, const int* __restrict__ tab_cap, int* tab,
                                int* __restrict__ row_slot) {
  const int gr = blockIdx.x * blockDim.x + threadIdx.x;
  if (gr >= n_rows) return;
  const int t = row_tint[gr], W = (tint_seg_n[t] + 31) >> 5;
  const int r0 = tint_row_off[t];
  const u32* base = rowbits + rowword_off[t];
  const u32* mine = base + (i64)(gr - r0) * W;
  auto eq = [&](int other) {
    const u32* o = base + (i64)(other - r0) * W;
    for (int w = 0; w < W; ++w)
      if (o[w] != mine[w]) return false;
    return true;
  };
  row_slot[gr] = cp_insert(tab + tab_off[t], tab_cap[t], row_hash[gr], gr, eq);
}

// class representative of every item (rows, reads or reps): the occupant of its slot once all inserts are done
__global__ void k_cp_resolve(int n, const int* __restrict__ item_tint, const i64* __restrict__ tab_off,
                             const int* __restrict__ tab, const int* __restrict__ slot, int* __restrict__ first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  first[i] = tab[tab_off[item_tint[i]] + slot[i]];
}

// ---- read keys (read_segment :154-160) ----
// renders "a-b:c" the way the SEGMENT file holds it (freddie_segment.py:472 sorts these strings)
__device__ __forceinline__ int cp_render(char* s, int a, int b, int c) {
  int n = 0;
  const int v[3] = {a, b, c};
  for (int k = 0; k < 3; ++k) {
    char tmp[12];
    int m = 0, x = v[k];
    do { tmp[m++] = (char)('0' + x % 10); x /= 10; } while (x > 0);
    while (m > 0) s[n++] = tmp[--m];
    if (k == 0) s[n++] = '-';
    if (k == 1) s[n++] = ':';
  }
  return n;
}
__device__ __forceinline__ bool cp_str_less(const int* x, const int* y) {
  char a[40], b[40];
  const int na = cp_render(a, x[0], x[1], x[2]), nb = cp_render(b, y[0], y[1], y[2]);
  const int n = na < nb ? na : nb;
  for (int k = 0; k < n; ++k)
    if (a[k] != b[k]) return a[k] < b[k];
  return na < nb;
}

// One thread per read: the read's internal gaps in the order of the file (sorted as strings, duplicates dropped),
// sizes <= 10 as 0; the two poly-tail entries (E before S in the sorted file); hash of the whole key.
__global__ void k_cp_read_key(int N, int T, const int* __restrict__ tint_read_off, const int* __restrict__ tint_row_off,
                              const int* __restrict__ read_row, const int* __restrict__ row_first,
                              const int* __restrict__ read_head, const int* __restrict__ read_gap_off,
                              const int* __restrict__ gap_rec, const int* __restrict__ tint_seg_n, int* __restrict__ gsort,
                              int* __restrict__ read_tint, int* __restrict__ key_row, int* __restrict__ key_cnt,
                              int* __restrict__ key_pe, int* __restrict__ key_ps, u64* __restrict__ key_hash, int* err) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int t = cp_owner(tint_read_off, T, i);
  read_tint[i] = t;
  const int flags = read_head[8 * (i64)i];
  const bool has = flags & 1;
  const int sk = (flags >> 8) & 3, ek = (flags >> 16) & 3;
  const int g0 = read_gap_off[i], g1 = read_gap_off[i + 1];
  int* g = gsort + 3 * (i64)g0;
  int n = has ? g1 - g0 : 0;
  const int M = tint_seg_n[t];
  for (int k = 0; k < n; ++k) {  // insertion sort by the strings (reads have a handful of gaps)
    int cur[3] = {gap_rec[3 * (i64)(g0 + k)], gap_rec[3 * (i64)(g0 + k) + 1], gap_rec[3 * (i64)(g0 + k) + 2]};
    if (!(0 <= cur[0] && cur[0] < cur[1] && cur[1] < M) || cur[2] < 0) atomicCAS(err, 0, CPERR_GAP);  // :164
    int p = k;
    while (p > 0 && cp_str_less(cur, g + 3 * (p - 1))) {
      g[3 * p] = g[3 * (p - 1)];
      g[3 * p + 1] = g[3 * (p - 1) + 1];
      g[3 * p + 2] = g[3 * (p - 1) + 2];
      --p;
    }
    g[3 * p] = cur[0];
    g[3 * p + 1] = cur[1];
    g[3 * p + 2] = cur[2];
  }
  int m = 0;  // drop exact duplicates (the file is written from a set), threshold the sizes
  for (int k = 0; k < n; ++k) {
    if (k > 0 && g[3 * k] == g[3 * (k - 1)] && g[3 * k + 1] == g[3 * (k - 1) + 1] && g[3 * k + 2] == g[3 * (k - 1) + 2]) continue;
    g[3 * m] = g[3 * k];
    g[3 * m + 1] = g[3 * k + 1];
    g[3 * m + 2] = g[3 * k + 2];
    ++m;
  }
  const int row = row_first[tint_row_off[t] + read_row[i]];
  const int pe = (has && ek) ? (read_head[8 * (i64)i + 5] > 10 ? read_head[8 * (i64)i + 5] : 0) : -1;
  const int ps = (has && sk) ? (read_head[8 * (i64)i + 2] > 10 ? read_head[8 * (i64)i + 2] : 0) : -1;
  u64 h = cp_mix(0x13198a2e03707344ULL, (u64)(u32)row);
  h = cp_mix(h, (u64)m);
  for (int k = 0; k < m; ++k) h = cp_mix(h, (u64)(g[3 * k + 2] > 10 ? g[3 * k + 2] : 0));
  h = cp_mix(h, ((u64)(u32)pe << 32) | (u32)ps);
  key_row[i] = row;
  key_cnt[i] = m;
  key_pe[i] = pe;
  key_ps[i] = ps;
  key_hash[i] = h;
}

__global__ void k_cp_read_insert(int N, const int* __restrict__ read_tint, const int* __restrict__ key_row,
                                 const int* __restrict__ key_cnt, const int* __restrict__ key_pe,
                                 const int* __restrict__ key_ps, const u64* __restrict__ key_hash,
                                 const int* __restrict__ read_gap_off, const int* __restrict__ gsort,
                                 const i64* __restrict__ tab_off, const int* __restrict__ tab_cap, int* tab,
                                 int* __restrict__ slot) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int t = read_tint[i];
  const int m = key_cnt[i];
  const int* mine = gsort + 3 * (i64)read_gap_off[i];
  auto eq = [&](int o) {
    if (key_row[o] != key_row[i] || key_cnt[o] != m || key_pe[o] != key_pe[i] || key_ps[o] != key_ps[i]) return false;
    const int* other = gsort + 3 * (i64)read_gap_off[o];
    for (int k = 0; k < m; ++k) {
      const int a = mine[3 * k + 2] > 10 ? mine[3 * k + 2] : 0, b = other[3 * k + 2] > 10 ? other[3 * k + 2] : 0;
      if (a != b) return false;
    }
    return true;
  };
  slot[i] = cp_insert(tab + tab_off[t], tab_cap[t], key_hash[i], i, eq);
}

// flags of the class representatives and the sizes of the classes
__global__ void k_cp_flag_count(int n, const int* __restrict__ first, int* __restrict__ flag, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  flag[i] = first[i] == i;
  atomicAdd(&count[first[i]], 1);
}

// ---- preprocess_ilp per rep (:277-328): one thread per read; the rep's first read does the work ----
__global__ void k_cp_rep_prep(int N, const int* __restrict__ read_tint, const int* __restrict__ tint_read_off,
                              const int* __restrict__ tint_seg_n, const int* __restrict__ read_first,
                              const int* __restrict__ rep_scan, const int* __restrict__ class_count,
                              const int* __restrict__ key_row, const int* __restrict__ row_f, const int* __restrict__ row_l,
                              const int* __restrict__ read_head, int* __restrict__ o_read_rep, int* __restrict__ rep_read,
                              int* __restrict__ rep_tint, int* __restrict__ rep_first_read, int* __restrict__ rep_count,
                              int* __restrict__ rep_fl, u8* __restrict__ rep_cat, int* __restrict__ rep_gap,
                              int* __restrict__ rep_row, u64* __restrict__ rep_hash) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int t = read_tint[i], r0 = tint_read_off[t];
  const int base = rep_scan[r0];
  o_read_rep[i] = rep_scan[read_first[i]] - base;
  if (read_first[i] != i) return;
  const int u = rep_scan[i], M = tint_seg_n[t];
  const int row = key_row[i];
  int f = row_f[row], l = row_l[row];
  const int flags = read_head[8 * (i64)i];
  const bool has = flags & 1;
  const int sk = has ? (flags >> 8) & 3 : 0, ek = has ? (flags >> 16) & 3 : 0;
  u8 cat = 'N';
  int ga = 0, gb = 0, gv = 0;
  if ((sk != 0) + (ek != 0) == 1) {  // len(read['poly_tail']) == 1 (:296)
    if (sk && read_head[8 * (i64)i + 1] > 10) {
      cat = 'S';
      ga = -1; gb = f; gv = read_head[8 * (i64)i + 2];
      f = 0;
    } else if (ek && read_head[8 * (i64)i + 4] > 10) {
      cat = 'E';
      ga = l; gb = M; gv = read_head[8 * (i64)i + 5];
      l = M - 1;
    }
  }
  rep_read[u] = i;
  rep_tint[u] = t;
  rep_first_read[u] = i - r0;
  rep_count[u] = class_count[i];
  rep_fl[2 * (i64)u] = f;
  rep_fl[2 * (i64)u + 1] = l;
  rep_cat[u] = cat;
  rep_gap[3 * (i64)u] = ga;
  rep_gap[3 * (i64)u + 1] = gb;
  rep_gap[3 * (i64)u + 2] = gv;
  rep_row[u] = row;
  // structure key (:211): the I row and (f, l, category); f and l follow from the row and the category
  rep_hash[u] = cp_mix(cp_mix(0x0a4093822299f31dULL, (u64)(u32)row), (u64)cat);
}

__global__ void k_cp_struct_insert(int U, const int* __restrict__ rep_tint, const int* __restrict__ rep_row,
                                   const u8* __restrict__ rep_cat, const u64* __restrict__ rep_hash,
                                   const i64* __restrict__ tab_off, const int* __restrict__ tab_cap, int* tab,
                                   int* __restrict__ slot) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  auto eq = [&](int o) { return rep_row[o] == rep_row[u] && rep_cat[o] == rep_cat[u]; };
  const int t = rep_tint[u];
  slot[u] = cp_insert(tab + tab_off[t], tab_cap[t], rep_hash[u], u, eq);
}

__global__ void k_cp_struct_fill(int U, const int* __restrict__ rep_tint, const int* __restrict__ tint_rep_off,
                                 const int* __restrict__ struct_first, const int* __restrict__ struct_scan,
                                 const int* __restrict__ class_count, const int* __restrict__ rep_row,
                                 const int* __restrict__ rep_fl, const u8* __restrict__ rep_cat,
                                 int* __restrict__ rep_struct, int* __restrict__ s_tint,
                                 int* __restrict__ s_row, int* __restrict__ s_f, int* __restrict__ s_l, u8* __restrict__ s_cat,
                                 int* __restrict__ s_cnt) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  const int t = rep_tint[u];
  const int base = struct_scan[tint_rep_off[t]];
  rep_struct[u] = struct_scan[struct_first[u]] - base;
  if (struct_first[u] != u) return;
  const int s = struct_scan[u];
  s_tint[s] = t;
  s_row[s] = rep_row[u];
  s_f[s] = rep_fl[2 * (i64)u];
  s_l[s] = rep_fl[2 * (i64)u + 1];
  s_cat[s] = rep_cat[u];
  s_cnt[s] = class_count[u];
}
// CSR offsets of the classes per tint from the offsets of the items: out[t] = scan[in_off[t]] (tints without
// items at the end: the total), t = 0..T
__global__ void k_cp_offsets(int T, const int* __restrict__ in_off, const int* __restrict__ scan, int n_items, int total,
                             int* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > T) return;
  const int r0 = in_off[t];
  out[t] = r0 < n_items ? scan[r0] : total;
}

// I and C rows of every rep (:289-291, :312-316).  One warp per rep.
__global__ void k_cp_rows_out(int U, const int* __restrict__ rep_tint, const int* __restrict__ tint_rep_off,
                              const int* __restrict__ rep_read, const int* __restrict__ read_row,
                              const int* __restrict__ tint_seg_n, const i64* __restrict__ tint_digit_off,
                              const u8* __restrict__ digits, const int* __restrict__ rep_fl, const i64* __restrict__ out_off,
                              u8* __restrict__ I, u8* __restrict__ C) {
  const int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (u >= U) return;
  const int t = rep_tint[u], M = tint_seg_n[t];
  const u8* src = digits + tint_digit_off[t] + (i64)read_row[rep_read[u]] * M;
  const i64 o = out_off[t] + (i64)(u - tint_rep_off[t]) * M;
  const int f = rep_fl[2 * (i64)u], l = rep_fl[2 * (i64)u + 1];
  for (int j = lane; j < M; j += 32) {
    const int d = (int)src[j] - '0';
    I[o + j] = (u8)(d & 1);
    C[o + j] = (u8)(j >= f && j <= l && d == 0);
  }
}

// ---- partition_reads (:198-274) ----
__global__ void k_cp_struct_bits(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                                 const int* __restrict__ s_row, const int* __restrict__ tint_row_off,
                                 const int* __restrict__ tint_seg_n, const i64* __restrict__ rowword_off,
                                 const u32* __restrict__ rowbits, const i64* __restrict__ sb_off, u32* __restrict__ sbits) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S_tot) return;
  const int t = s_tint[s], W = (tint_seg_n[t] + 31) >> 5;
  const int S = tint_struct_off[t + 1] - tint_struct_off[t], sl = s - tint_struct_off[t];
  const u32* src = rowbits + rowword_off[t] + (i64)(s_row[s] - tint_row_off[t]) * W;
  u32* dst = sbits + sb_off[t] + sl;
  for (int w = 0; w < W; ++w) dst[(i64)w * S] = src[w];
}

// The pair test (:219-236): one warp per structure i, lane = structure j of the current adjacency word.
// A pair is compatible if the categories agree (or one is 'N'), the rows share a 1 inside the overlap [F, L] of
// their segment ranges, and they differ in < 3 segments there (0 if the overlap is 1-3 segments long).  F = -1
// (both rows without a 1: the reference slices from the END then) cannot have a common 1: incompatible.
__global__ void k_cp_pair_test(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                               const int* __restrict__ s_f, const int* __restrict__ s_l, const u8* __restrict__ s_cat,
                               const i64* __restrict__ sb_off, const u32* __restrict__ sbits,
                               const i64* __restrict__ adj_off, u32* __restrict__ adj, i64* __restrict__ tint_edges) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gi >= S_tot) return;
  const int t = s_tint[gi], s0 = tint_struct_off[t];
  const int S = tint_struct_off[t + 1] - s0, SW = (S + 31) >> 5, i = gi - s0;
  const u32* bits = sbits + sb_off[t];
  u32* row = adj + adj_off[t] + (i64)i * SW;
  const int fi = s_f[gi], li = s_l[gi];
  const u8 ci = s_cat[gi];
  int edges = 0;
  for (int jw = 0; jw < SW; ++jw) {
    const int j = jw * 32 + lane;
    bool ok = false;
    if (j < S && j != i) {
      const u8 cj = s_cat[s0 + j];
      const int F = max(fi, s_f[s0 + j]), L = min(li, s_l[s0 + j]);
      if (!(ci != 'N' && cj != 'N' && ci != cj) && F >= 0 && L >= F) {
        const int o = L - F + 1;
        const int dmax = o > 3 ? 2 : 0;
        int wsum = 0, d = 0;
        for (int w = F >> 5; w <= (L >> 5) && d <= dmax; ++w) {
          u32 m = 0xffffffffu;
          if (w == (F >> 5)) m &= 0xffffffffu << (F & 31);
          if (w == (L >> 5)) m &= 0xffffffffu >> (31 - (L & 31));
          const u32 a = bits[(i64)w * S + i], b = bits[(i64)w * S + j];
          wsum += __popc(a & b & m);
          d += __popc((a ^ b) & m);
        }
        ok = wsum >= 1 && d <= dmax;
      }
    }
    const u32 word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) row[jw] = word;
    edges += __popc(word);
  }
  if (lane == 0 && edges) atomicAdd((u64*)&tint_edges[2 * t], (u64)edges);  // both directions: halved by the host
}

__global__ void k_cp_degree(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                            const i64* __restrict__ adj_off, const u32* __restrict__ adj, const int* __restrict__ tint_active,
                            int* __restrict__ deg) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gi >= S_tot) return;
  const int t = s_tint[gi];
  if (!tint_active[t]) return;
  const int s0 = tint_struct_off[t], S = tint_struct_off[t + 1] - s0, SW = (S + 31) >> 5;
  const u32* row = adj + adj_off[t] + (i64)(gi - s0) * SW;
  int d = 0;
  for (int w = lane; w < SW; w += 32) d += __popc(row[w]);
  for (int o = 16; o; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) deg[gi] = d;
}

// One synchronous pruning round (:242-255): an edge (i, j) stays if i or j has no other neighbour or if they
// have a common neighbour; decided on the graph `a` of the round's start, written to `b`.  One warp per row i:
// for every edge the lanes AND the two rows 32 words at a time.  Both directions of an edge are decided
// independently and identically.
__global__ void k_cp_prune_round(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                                 const i64* __restrict__ adj_off, const u32* __restrict__ a, u32* __restrict__ b,
                                 const int* __restrict__ deg, const int* __restrict__ tint_active,
                                 int* __restrict__ tint_active_next, int* __restrict__ any_change) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gi >= S_tot) return;
  const int t = s_tint[gi];
  if (!tint_active[t]) return;
  const int s0 = tint_struct_off[t], S = tint_struct_off[t + 1] - s0, SW = (S + 31) >> 5, i = gi - s0;
  const u32* A = a + adj_off[t];
  const u32* ri = A + (i64)i * SW;
  u32* out = b + adj_off[t] + (i64)i * SW;
  const int di = deg[gi];
  bool changed = false;
  for (int jw = 0; jw < SW; ++jw) {
    const u32 word = ri[jw];
    u32 keep = word;
    if (di != 1) {
      u32 rest = word;
      while (rest) {
        const int bit = __ffs(rest) - 1;
        rest &= rest - 1;
        const int j = jw * 32 + bit;
        if (deg[s0 + j] == 1) continue;
        const u32* rj = A + (i64)j * SW;
        bool common = false;
        for (int w0 = 0; w0 < SW && !common; w0 += 32) {
          const int w = w0 + lane;
          const u32 x = w < SW ? (ri[w] & rj[w]) : 0u;
          common = __any_sync(0xffffffffu, x != 0);
        }
        if (!common) keep &= ~(1u << bit);
      }
    }
    if (lane == 0) out[jw] = keep;
    changed |= keep != word;
  }
  if (lane == 0 && changed) {
    tint_active_next[t] = 1;
    *any_change = 1;
  }
}

// connected components (:257): union-find with the smaller root as the parent, so that the root of a component
// is its smallest structure -- the order networkx yields components in
__device__ __forceinline__ int cp_find(int* parent, int x) {
  int p = *(volatile int*)&parent[x];
  while (p != x) {
    x = p;
    p = *(volatile int*)&parent[x];
  }
  return x;
}
__global__ void k_cp_parent_init(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                                 int* __restrict__ parent) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < S_tot) parent[s] = s - tint_struct_off[s_tint[s]];
}
__global__ void k_cp_union(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off,
                           const i64* __restrict__ adj_off, const u32* __restrict__ adj, int* parent,
                           i64* __restrict__ tint_edges) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gi >= S_tot) return;
  const int t = s_tint[gi], s0 = tint_struct_off[t];
  const int S = tint_struct_off[t + 1] - s0, SW = (S + 31) >> 5, i = gi - s0;
  const u32* row = adj + adj_off[t] + (i64)i * SW;
  int* par = parent + s0;
  int edges = 0;
  for (int w = lane; w < SW; w += 32) {
    u32 word = row[w];
    edges += __popc(word);
    if (w * 32 > i) continue;
    while (word) {
      const int j = w * 32 + __ffs(word) - 1;
      word &= word - 1;
      if (j >= i) break;
      while (true) {
        int a = cp_find(par, i), b = cp_find(par, j);
        if (a == b) break;
        if (a < b) { const int x = a; a = b; b = x; }
        if (atomicCAS(&par[a], a, b) == a) break;
      }
    }
  }
  for (int o = 16; o; o >>= 1) edges += __shfl_xor_sync(0xffffffffu, edges, o);
  if (lane == 0 && edges) atomicAdd((u64*)&tint_edges[2 * t + 1], (u64)edges);
}
__global__ void k_cp_labels(int S_tot, const int* __restrict__ s_tint, const int* __restrict__ tint_struct_off, int* parent,
                            int* __restrict__ label) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S_tot) return;
  const int s0 = tint_struct_off[s_tint[s]];
  label[s] = cp_find(parent + s0, s - s0);
}

// Incompatible pairs of a piece (:266-273).  q = a position of the flat piece list; its structure i is paired
// with every LATER structure j of the same piece that is not adjacent to it; the pair contributes the product of
// their rep lists, r1-major.  One warp per q; pass 0 counts, pass 1 writes at the scanned offsets.
template <int WRITE>
__global__ void k_cp_incomp(int Q, const int* __restrict__ q_node, const int* __restrict__ q_end, const int* __restrict__ s_tint,
                            const int* __restrict__ tint_struct_off, const i64* __restrict__ adj_off,
                            const u32* __restrict__ adj, const int* __restrict__ mem_off, const int* __restrict__ mem,
                            i64* __restrict__ q_pairs, const i64* __restrict__ q_pair_off, int* __restrict__ inc) {
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (q >= Q) return;
  const int gi = q_node[q], t = s_tint[gi], s0 = tint_struct_off[t];
  const int S = tint_struct_off[t + 1] - s0, SW = (S + 31) >> 5;
  const u32* row = adj + adj_off[t] + (i64)(gi - s0) * SW;
  const int mi = mem_off[gi + 1] - mem_off[gi];
  const int end = q_end[q];
  i64 total = 0;
  i64 base = WRITE ? q_pair_off[q] : 0;
  for (int p0 = q + 1; p0 < end; p0 += 32) {
    const int p = p0 + lane;
    i64 c = 0;
    int gj = 0;
    if (p < end) {
      gj = q_node[p];
      const int j = gj - s0;
      if (!((row[j >> 5] >> (j & 31)) & 1u)) c = (i64)mi * (mem_off[gj + 1] - mem_off[gj]);
    }
    i64 incl = c;
    for (int o = 1; o < 32; o <<= 1) {
      const i64 v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (WRITE && c) {
      const int mj = mem_off[gj + 1] - mem_off[gj];
      i64 w = base + incl - c;
      for (int x = 0; x < mi; ++x) {
        const int r1 = mem[mem_off[gi] + x];
        for (int y = 0; y < mj; ++y, ++w) {
          inc[2 * w] = r1;
          inc[2 * w + 1] = mem[mem_off[gj] + y];
        }
      }
    }
    const i64 chunk = __shfl_sync(0xffffffffu, incl, 31);
    total += chunk;
    base += chunk;
  }
  if (!WRITE && lane == 0) q_pairs[q] = total;
}
