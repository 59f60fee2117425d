// kernels_dp.cuh -- cumulative coverage, fixed candidates, subproblem tables, DP tables + solve.
// Reference steps: get_cumulative_coverage (freddie_segment.py:188-246), fixed set (:776-788),
// break_large_problems (:623-645), optimize / run_optimize (:475-596).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// K5 cumulative coverage.  P[q][r] = number of samples of read rep r strictly before candidate q
// (flat coordinates), one row per candidate of the tint, one column per rep (row stride Rp, a
// multiple of 4 so that row segments can be fetched with 16-byte TMA bulk copies).  Inside an
// island  P[j]-P[i]  equals the reference's  C[j]-C[i]  (:493): earlier islands add the same
// constant to both rows.  One thread per rep walks the candidates and its (sorted) intervals with
// two pointers; a warp writes 128 contiguous bytes of one row per step.  The candidate rows of a tint
// are cut into up to COV_CHUNKS chunks (grid.y) so that the serial walk of a thread is bounded by
// Kt / COV_CHUNKS candidates (+ its own few intervals, re-walked from the first one per chunk).
// ---------------------------------------------------------------------------------------------
#define COV_THREADS 128
#define COV_CHUNKS 8
#define COV_MIN_CHUNK 32
struct RepTile { int tint; int rep_lo; };  // rep_lo: tint-local first rep of the tile

__global__ void __launch_bounds__(COV_THREADS) k_coverage(const RepTile* __restrict__ tiles,
                                                         const int* __restrict__ tint_rep_off,
                                                         const int* __restrict__ tint_cand_off,
                                                         const i64* __restrict__ tint_cov_off,
                                                         const int* __restrict__ rep_iv_off,
                                                         const int* __restrict__ iv_fs, const int* __restrict__ iv_fe,
                                                         const int* __restrict__ cand_flat, u32* __restrict__ P,
                                                         int n_tints, i64 cap_P) {
  pdl_prologue();
  // the matrix does not fit its buffer: the host grows it and repeats the run (the total is read from the scan,
  // not from the counters: the kernel runs on a side stream before k_plan_finish copies it there)
  if (tint_cov_off[n_tints] > cap_P) return;
  const RepTile tl = tiles[blockIdx.x];
  const int r0 = tint_rep_off[tl.tint];
  const int R = tint_rep_off[tl.tint + 1] - r0;
  const int Rp = (R + 3) & ~3;
  const int r = tl.rep_lo + threadIdx.x;
  if (r >= Rp) return;
  const int q0 = tint_cand_off[tl.tint], q1 = tint_cand_off[tl.tint + 1];
  const int chunk = max(COV_MIN_CHUNK, (q1 - q0 + COV_CHUNKS - 1) / COV_CHUNKS);
  const int qa = q0 + (int)blockIdx.y * chunk, qb = min(q1, qa + chunk);
  if (qa >= q1) return;
  u32* out = P + tint_cov_off[tl.tint] + r;
  if (r >= R) {  // padding columns
    for (int q = qa; q < qb; ++q) out[(i64)(q - q0) * Rp] = 0u;
    return;
  }
  int a = rep_iv_off[r0 + r];
  const int b = rep_iv_off[r0 + r + 1];
  u32 acc = 0;
  int fs = (a < b) ? iv_fs[a] : 0x7fffffff;
  int fe = (a < b) ? iv_fe[a] : 0x7fffffff;
  for (int q = qa; q < qb; ++q) {
    const int cf = cand_flat[q];
    while (a < b && fe < cf) {  // interval entirely before the candidate (te is an inclusive sample)
      acc += (u32)(fe - fs + 1);
      ++a;
      fs = (a < b) ? iv_fs[a] : 0x7fffffff;
      fe = (a < b) ? iv_fe[a] : 0x7fffffff;
    }
    u32 part = (a < b && cf > fs) ? (u32)(cf - fs) : 0u;
    out[(i64)(q - q0) * Rp] = acc + part;
  }
}

// per tint: first candidate rank and size of the coverage block (rows = candidates, stride Rp)
__global__ void k_tint_cov_sizes(int T, const int* __restrict__ tint_island_off, const int* __restrict__ island_cand_off,
                                 const int* __restrict__ tint_rep_off, int* __restrict__ tint_cand_off,
                                 i64* __restrict__ cov_sz) {
  pdl_prologue();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > T) return;
  int q = island_cand_off[tint_island_off[t]];
  tint_cand_off[t] = q;
  if (t < T) {
    int q1 = island_cand_off[tint_island_off[t + 1]];
    int R = tint_rep_off[t + 1] - tint_rep_off[t];
    cov_sz[t] = (i64)(q1 - q) * ((R + 3) & ~3);
  }
}

// ---------------------------------------------------------------------------------------------
// K6 fixed candidates.  a: ends of each island and candidates above the tint's threshold.
// b: break_large_problems over the SNAPSHOT of consecutive fixed pairs; additions go to fixed1.
// ---------------------------------------------------------------------------------------------
__global__ void k_fixed_a(const i64* __restrict__ n_cand_p, const int* __restrict__ cand_flat, const int* __restrict__ cand_island,
                          const int* __restrict__ island_cand_off, const int* __restrict__ island_tint,
                          const double* __restrict__ y, const double* __restrict__ thr, u8* __restrict__ fixed0,
                          u8* __restrict__ fixed1) {
  pdl_prologue();
  const int n_cand = (int)*n_cand_p;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_cand; q += gridDim.x * blockDim.x) {
    int isl = cand_island[q];
    bool f = (q == island_cand_off[isl]) || (q == island_cand_off[isl + 1] - 1);
    if (!f) f = y[cand_flat[q]] > thr[island_tint[isl]];
    fixed0[q] = f;
    fixed1[q] = f;
  }
}

__global__ void k_fixed_b(const i64* __restrict__ n_cand_p, const int* __restrict__ cand_flat, const int* __restrict__ cand_island,
                          const int* __restrict__ island_cand_off, const double* __restrict__ y, int mps,
                          const u8* __restrict__ fixed0, u8* __restrict__ fixed1, int* __restrict__ err) {
  pdl_prologue();
  const int n_cand = (int)*n_cand_p;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_cand; q += gridDim.x * blockDim.x) {
    if (!fixed0[q]) continue;
    int isl = cand_island[q];
    int c0 = island_cand_off[isl], c1 = island_cand_off[isl + 1];
    if (q == c1 - 1) continue;
    int e = q + 1;
    while (!fixed0[e]) ++e;  // the island's last candidate is fixed
    int size = e - q + 1;
    if (size <= mps) continue;
    int cnt = (int)ceil((double)size / (double)mps);
    double ps = __ddiv_rn((double)size, (double)cnt);
    int s_local = q - c0;
    for (int i = 1; i < cnt; ++i) {
      int mid = (int)__dadd_rn((double)s_local, __dmul_rn((double)i, ps));
      double best = -INFINITY;
      int best_c = -1;
      bool bad = false;
      for (int c = mid - 5; c < mid + 5; ++c) {
        if (c < 0 || c0 + c >= c1) { dev_fail(err, DEVERR_BREAK_LARGE_RANGE, q); bad = true; break; }
        double v = y[cand_flat[c0 + c]];
        if (v > best) { best = v; best_c = c; }
      }
      if (bad) break;
      if (!(best > 0.0)) { dev_fail(err, DEVERR_BREAK_LARGE_POS, q); break; }
      fixed1[c0 + best_c] = 1;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Subproblem plan.  Every subproblem gets a size class (by its candidate count n) and a mode:
//   fused : all read reps of the tint fit one slab (words <= slab_words) and n <= DP_SMEM_MAX_N.
//           ONE CTA builds ins/out in shared memory and solves the DP in place -- the tables
//           never touch HBM ("one CTA per tint-sized problem").
//   split : giant tints.  The reps are cut into slabs of slab_words x 32; every (subproblem, slab)
//           is a CTA that accumulates its partial tables in shared memory and adds them to the
//           global tables with RED once per slab ("multi-CTA cooperative mode"); k_dp_solve then
//           runs on the summed tables.
// Classes: n <= 8 / 16 / 32 / DP_SMEM_MAX_N share a kernel body with different CTA sizes and
// shared-memory carve-ups; class 4 (n > DP_SMEM_MAX_N, only reachable with a large -mps) keeps its
// out table in global memory and is always split.
// ---------------------------------------------------------------------------------------------
#define DP_CLASSES 6
#define DP_SMEM_MAX_N 56
#ifndef DP_WARP_MAX_WORDS
#define DP_WARP_MAX_WORDS 16
#endif
// counter slots (i64) written by k_sub_plan
#define PLAN_WORK 0     // [6] work items per class
#define PLAN_MAXN 6     // [6] largest n per class
#define PLAN_SPLIT 12   // subproblems that need k_dp_solve
#define PLAN_CELLS 13   // sum C(n,3)
#define PLAN_RCELLS 14  // sum C(n,3) * R
#define PLAN_MAXALL 15  // largest n
#define PLAN_NSUB 16    // number of subproblems
#define PLAN_TAB 17     // int32 elements of the global DP tables
#define PLAN_SLOTS 20
// Work items of a class are listed heaviest first: every subproblem falls into one of DP_BUCKETS cost buckets
// (log2 of n^2 x words of an item), a class's list is the concatenation of its buckets in descending order, and
// the persistent CTAs / warps take items from the front.  Without it the order is arbitrary and a heavy item
// taken last is the tail of the launch (measured on config 2: 14 % of the warp slots busy on average).
#define DP_BUCKETS 8
__host__ __device__ inline int dp_bucket_of(int n, int item_words) {
  const long long cost = (long long)n * n * (item_words > 0 ? item_words : 1);
  int lg = 0;
  while ((cost >> (lg + 1)) > 0) ++lg;
  const int b = (lg - 6) / 2;
  return b < 0 ? 0 : b >= DP_BUCKETS ? DP_BUCKETS - 1 : b;
}

struct DpWork { int sub; int slab; };

// classes: 0 / 1 = one WARP per subproblem (n <= 8 / 16, tint of at most 512 reps);
//          2 / 3 / 4 = one CTA of 128 / 256 / 512 threads per (subproblem, slab), n <= 16 / 32 / 56;
//          5 = n > 56 (only reachable with a large -mps): out table in global memory, always split.
__host__ __device__ inline int dp_class_of(int n, int words, int fused, int warp_max_words) {
  if (fused && n <= 16 && words <= warp_max_words) return n <= 8 ? 0 : 1;
  return n <= 16 ? 2 : n <= 32 ? 3 : n <= DP_SMEM_MAX_N ? 4 : 5;
}

// read-rep words per CTA of a subproblem: bounded work per CTA (~ n^3 x words) so that a few large
// subproblems of a mid-size tint do not become the tail of the launch, at least ~8 slabs for tints
// well above that bound, and at most `cap` words (giant tints: fewer, fatter CTAs -> fewer REDs).
__host__ __device__ inline int dp_slab_words(int n, int words, int cap) {
  int lat = (1 << 20) / (n * n * n);
  lat = max(4, min(64, lat)) & ~3;
  int spread = (((words + 7) / 8) + 3) & ~3;
  if (words <= 256) cap = min(cap, 16);  // mid-size tints: more, shorter CTAs (swept on B200); giant tints keep fat slabs
  return max(1, min(cap, max(lat, spread)));
}

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Subproblems = consecutive fixed candidates (a, b) of one island with at least one interior
// candidate (:571-596).  One thread per candidate: a fixed candidate looks for the next fixed one
// (at most ~mps steps after break_large_problems) and, if there is an interior, appends the
// subproblem to the list (slot from a warp-aggregated atomic: the ORDER of the list is arbitrary, the
// results do not depend on it).  Per subproblem: class, slab size, CTA count, table block.
// sub_info[p] = class | fused << 8 | slab_words << 16;  sub_slabs[p] = CTAs of the subproblem.
// sub_tab_off[p] = first int32 of the subproblem's global table block (pair-indexed ins [n(n-1)/2]
// followed by out [C(n,3)]); only subproblems whose tables leave the chip own one.
__global__ void k_sub_build(const i64* __restrict__ n_cand_p, const u8* __restrict__ fixed, const int* __restrict__ cand_island,
                            const int* __restrict__ island_cand_off, const int* __restrict__ island_tint,
                            const int* __restrict__ tint_rep_off, const int* __restrict__ tint_read_off, int slab_cap,
                            int warp_max_words, int keep_tables, int* __restrict__ sub_start, int* __restrict__ sub_n,
                            int* __restrict__ sub_tint, int* __restrict__ sub_info, int* __restrict__ sub_slabs,
                            i64* __restrict__ sub_tab_off, i64* __restrict__ plan, int* __restrict__ err) {
  pdl_prologue();
  const int n_cand = (int)*n_cand_p;
  const int lane = threadIdx.x & 31;
  // CTA-uniform trip count: every lane of a warp reaches the ballots
  for (int q0 = blockIdx.x * blockDim.x; q0 < n_cand; q0 += gridDim.x * blockDim.x) {
    const int q = q0 + threadIdx.x;
    int cls = -1, slabs = 0, n = 0, split = 0, info = 0, t = 0, bucket = 0;
    long long t3 = 0, rc = 0, sz = 0;
    bool has = false;
    if (q < n_cand && fixed[q]) {
      const int isl = cand_island[q];
      const int c1 = island_cand_off[isl + 1];
      if (q < c1 - 1) {
        int e = q + 1;
        while (!fixed[e]) ++e;  // the island's last candidate is fixed
        if (e - q >= 2) {
          has = true;
          n = e - q + 1;
          t = island_tint[isl];
          const int R = tint_rep_off[t + 1] - tint_rep_off[t];
          const int words = (R + 31) >> 5;
          const int sw = dp_slab_words(n, words, slab_cap);
          const int fused = (n <= DP_SMEM_MAX_N && words <= sw) ? 1 : 0;
          cls = dp_class_of(n, words, fused, warp_max_words);
          slabs = fused ? 1 : (words + sw - 1) / sw;
          bucket = dp_bucket_of(n, fused ? words : sw);
          split = (!fused && cls == DP_CLASSES - 1) ? 1 : 0;  // only n > DP_SMEM_MAX_N needs the separate solver
          t3 = (long long)n * (n - 1) * (n - 2) / 6;
          rc = t3 * R;
          info = cls | (fused << 8) | (bucket << 12) | (sw << 16);
          sz = (fused && !keep_tables) ? 0 : (long long)n * (n - 1) / 2 + t3;
          // |score| <= (segments of a path) x (reads of the tint): must stay inside the 30-bit range of FRS_NEG_INF
          if ((long long)n * (tint_read_off[t + 1] - tint_read_off[t]) >= 0x3fffffffLL) dev_fail(err, DEVERR_SCORE_RANGE, t);
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, has);
    if (m == 0) continue;
    int base = 0;
    if (lane == __ffs(m) - 1) base = (int)atomicAdd((unsigned long long*)&plan[PLAN_NSUB], (unsigned long long)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (has) {
      const int p = base + __popc(m & ((1u << lane) - 1u));
      sub_start[p] = q;
      sub_n[p] = n;
      sub_tint[p] = t;
      sub_info[p] = info;
      sub_slabs[p] = slabs;
      sub_tab_off[p] = sz ? (i64)atomicAdd((unsigned long long*)&plan[PLAN_TAB], (unsigned long long)sz) : 0;
      // items per (class, bucket), aggregated over the lanes that share the key
      const int key = cls * DP_BUCKETS + bucket;
      const unsigned grp = __match_any_sync(m, key);
      const int tot = __reduce_add_sync(grp, slabs);
      if (lane == __ffs(grp) - 1)
        atomicAdd((unsigned long long*)&plan[CNT_BUCKET - CNT_PLAN + key], (unsigned long long)tot);
    }
    // warp-aggregated statistics
#pragma unroll
    for (int c = 0; c < DP_CLASSES; ++c) {
      long long s = warp_sum_ll(cls == c ? slabs : 0);
      int mx = warp_max_i(cls == c ? n : 0);
      if (lane == 0 && s) {
        atomicAdd((unsigned long long*)&plan[PLAN_WORK + c], (unsigned long long)s);
        atomicMax((long long*)&plan[PLAN_MAXN + c], (long long)mx);
      }
    }
    long long s_split = warp_sum_ll(split), s_t3 = warp_sum_ll(t3), s_rc = warp_sum_ll(rc);
    int m_all = warp_max_i(n);
    if (lane == 0) {
      if (s_split) atomicAdd((unsigned long long*)&plan[PLAN_SPLIT], (unsigned long long)s_split);
      if (s_t3) atomicAdd((unsigned long long*)&plan[PLAN_CELLS], (unsigned long long)s_t3);
      if (s_rc) atomicAdd((unsigned long long*)&plan[PLAN_RCELLS], (unsigned long long)s_rc);
      atomicMax((long long*)&plan[PLAN_MAXALL], (long long)m_all);
    }
  }
}

// One thread, after k_sub_build and the coverage-offset scan: first work item of every class, totals, and
// the cursors of the work lists (fill cursors [0..DP_CLASSES], work-stealing cursors [8..8+DP_CLASSES]).
__global__ void k_plan_finish(i64* __restrict__ cnt, const i64* __restrict__ tint_cov_off, int n_tints,
                              int* __restrict__ bases /* [16 + DP_CLASSES * DP_BUCKETS]: classes, then (class, bucket) */,
                              int* __restrict__ cursor /* [16 + DP_CLASSES * DP_BUCKETS] */) {
  pdl_prologue();
  // one warp (launched <<<1, 32>>>): exclusive prefix of the (class, bucket) counts in list order -- class by class,
  // heaviest bucket first -- with the loads of all entries in flight together (one thread walking the 48 counters
  // was a chain of dependent loads: 8 us)
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  constexpr int NE = DP_CLASSES * DP_BUCKETS, NJ = (NE + 31) / 32;
  i64 v[NJ], ex[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = lane + 32 * j;  // list position: class e / DP_BUCKETS, bucket DP_BUCKETS - 1 - e % DP_BUCKETS
    v[j] = e < NE ? cnt[CNT_BUCKET + (e / DP_BUCKETS) * DP_BUCKETS + (DP_BUCKETS - 1 - e % DP_BUCKETS)] : 0;
  }
  i64 carry = 0;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    i64 x = v[j];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const i64 y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    ex[j] = carry + x - v[j];
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = lane + 32 * j;
    if (e < NE) {
      const int k = e / DP_BUCKETS, b = DP_BUCKETS - 1 - e % DP_BUCKETS;
      const int base = (int)(ex[j] < 0x7fffffffLL ? ex[j] : 0x7fffffffLL);
      bases[16 + k * DP_BUCKETS + b] = base;
      if (b == DP_BUCKETS - 1) bases[k] = base;  // a class starts with its heaviest bucket
    }
  }
  if (lane == 0) {
    cnt[CNT_NWORK] = carry;
    cnt[CNT_COV] = tint_cov_off[n_tints];
  }
  for (int k = lane; k < 16 + NE; k += 32) cursor[k] = 0;
}

// every data-dependent buffer of the DP stage fits its capacity (else the stage is skipped and repeated)
__device__ __forceinline__ bool dp_caps_ok(const i64* __restrict__ cnt, const Caps& cp) {
  return cnt[CNT_COV] <= cp.P && cnt[CNT_PLAN + PLAN_TAB] <= cp.tab && cnt[CNT_NWORK] <= cp.work &&
         cnt[CNT_PLAN + PLAN_SPLIT] <= cp.split;
}

// work lists: class c owns work[base[c] .. base[c] + count[c]); cursor[c] starts at 0.  The order of
// the items inside a class is arbitrary (atomics) -- results do not depend on it (integer sums).
__global__ void k_sub_fill(const i64* __restrict__ cnt, Caps caps, const int* __restrict__ sub_info,
                           const int* __restrict__ sub_slabs, const int* __restrict__ bases, int* __restrict__ cursor,
                           DpWork* __restrict__ work, int* __restrict__ split_list) {
  pdl_prologue();
  if (!dp_caps_ok(cnt, caps)) return;
  const int n_sub = (int)cnt[CNT_PLAN + PLAN_NSUB];
  const int lane = threadIdx.x & 31;
  // warp-uniform loop; the single-slab items of a warp that share a (class, bucket) key bump its cursor with ONE
  // atomic (32 k atomics on ~48 addresses serialised in L2: 25 us for 0.2 M instructions); multi-slab items keep
  // their own (a unique match key)
  for (int p0 = blockIdx.x * blockDim.x + threadIdx.x - lane; p0 < n_sub; p0 += gridDim.x * blockDim.x) {
    const int p = p0 + lane;
    const bool valid = p < n_sub;
    int info = 0, cls = 0, slabs = 0, key = 0;
    if (valid) {
      info = sub_info[p];
      cls = info & 0xff;
      slabs = sub_slabs[p];
      key = 16 + cls * DP_BUCKETS + ((info >> 12) & 7);
    }
    const bool single = valid && slabs == 1;
    const unsigned peers = __match_any_sync(0xffffffffu, single ? key : -1 - lane);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (valid && lane == leader) base = atomicAdd(&cursor[key], single ? __popc(peers) : slabs);
    base = __shfl_sync(0xffffffffu, base, leader);
    if (!valid) continue;
    const int off = bases[key] + base + (single ? __popc(peers & ((1u << lane) - 1u)) : 0);
    for (int s = 0; s < slabs; ++s) work[off + s] = DpWork{p, s};
    if (!((info >> 8) & 1) && cls == DP_CLASSES - 1) split_list[atomicAdd(&cursor[DP_CLASSES], 1)] = p;
  }
}

// zeroes the global DP tables of the run (their size is only known on the device)
__global__ void k_zero_tab(const i64* __restrict__ cnt, Caps caps, int* __restrict__ tab) {
  pdl_prologue();
  if (!dp_caps_ok(cnt, caps)) return;
  const i64 n = cnt[CNT_PLAN + PLAN_TAB];
  for (i64 e = (i64)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (i64)gridDim.x * blockDim.x) tab[e] = 0;
}

// ---------------------------------------------------------------------------------------------
// index helpers.  pairs (i<j): row-major upper triangle.  triples (i<j<k): [j][k-j-1][i], j = 1..n-2
// -> exactly C(n,3) entries; the left candidate i is the fastest index because the lanes of the triple
// phase run over i (conflict-free shared-memory accumulation).
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int pair_index(int i, int j, int n) { return i * (2 * n - i - 1) / 2 + (j - i - 1); }
__host__ __device__ __forceinline__ int triple_mid_off(int j, int n) {
  // sum_{j'=1}^{j-1} j' (n-1-j')
  int m = j - 1;
  return (n - 1) * m * (m + 1) / 2 - m * (m + 1) * (2 * m + 1) / 6;
}

__host__ __device__ __forceinline__ int triple_index(int i, int j, int k, int n) {
  return triple_mid_off(j, n) + (k - j - 1) * j + i;
}

__device__ __forceinline__ u32 smem_u32(const void* p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, u32 parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, u32 bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------------------------------------
// K8 DP solve (block-cooperative device function; tables in shared OR global memory).
//   G(j,k) = best continuation after committing segment (j,k):
//   G(j,E) = ins(j,E);  G(j,k) = max_{k'>k} D(j,k,k')  (ascending k', first maximum wins)
//   D(i,j,k) = ins(i,j) + out(i,j,k) + G(j,k)  if both segments span >= 5 samples, out >= lo and
//              G(j,k) is finite, else -inf                                   (:532-558)
// Top level (:560-566): lexicographic (j,k), strict improvement over ins(0,E).  Backtrace (:592-594)
// marks the chosen candidates.  amb = pair-indexed POSITIVE ambiguous counts: ins(i,j) = -amb[pair(i,j)].
// ---------------------------------------------------------------------------------------------
#define DPS_MAX_WARPS 32
template <bool WARP>
__device__ __forceinline__ void dp_sync() {
  if (WARP) __syncwarp(); else __syncthreads();
}
// WARP = true: executed by one warp (lanes), else by the whole CTA.  For a fixed j every (k, k2) is
// independent: groups of SUB lanes own a k, the lanes of a group stride over k2 (ascending, strict >
// keeps the first maximum per lane) and a shuffle reduction picks the maximum with the smallest k2.
template <bool WARP, int SUB>
__device__ void dp_solve(const int n, const int* cf, const int* amb, const int* out, const int lo,
                         int* G /*[n*n]*/, short* arg /*[n*n]*/, int* red /*[2*DPS_MAX_WARPS], CTA mode only*/,
                         u8* __restrict__ final_flag /* + qs */, int* __restrict__ err, int p) {
  const int tid = WARP ? (int)(threadIdx.x & 31) : (int)threadIdx.x;
  const int nt = WARP ? 32 : (int)blockDim.x;
  const int E = n - 1;
  const int grp = tid / SUB, gl = tid % SUB, ngrp = nt / SUB;
  for (int e = tid; e < n * n; e += nt) { G[e] = FRS_NEG_INF; arg[e] = -1; }
  dp_sync<WARP>();
  for (int j = tid; j < E; j += nt) G[j * n + E] = -amb[pair_index(j, E, n)];
  dp_sync<WARP>();
  for (int j = E - 2; j >= 0; --j) {
    for (int kk = j + 1; kk < E; kk += ngrp) {  // uniform trip count: the shuffles below need every lane
      const int k = kk + grp;
      int best = FRS_NEG_INF, bk = 0x7fff;
      if (k < E && cf[k] - cf[j] >= 5) {
        const int base = -amb[pair_index(j, k, n)];
        const int* o = out + triple_mid_off(k, n) + j;  // out(j, k, k2) = o[(k2-k-1)*k]
        for (int k2 = k + 1 + gl; k2 <= E; k2 += SUB) {
          if (cf[k2] - cf[k] < 5) continue;
          int ov = o[(k2 - k - 1) * k];
          if (ov < lo) continue;
          int g = G[k * n + k2];
          if (g == FRS_NEG_INF) continue;
          int d = base + ov + g;
          if (d > best) { best = d; bk = k2; }
        }
      }
#pragma unroll
      for (int s = SUB / 2; s > 0; s >>= 1) {
        int ob = __shfl_xor_sync(0xffffffffu, best, s);
        int ok = __shfl_xor_sync(0xffffffffu, bk, s);
        if (ob > best || (ob == best && ok < bk)) { best = ob; bk = ok; }
      }
      if (gl == 0 && k < E) {
        G[j * n + k] = best;
        arg[j * n + k] = (short)(best == FRS_NEG_INF ? -1 : bk);
      }
    }
    dp_sync<WARP>();
  }
  // top level: D(0,j,k) over 1 <= j < k <= E, first maximum in lexicographic order
  int my_best = FRS_NEG_INF, my_e = 0x7fffffff;
  for (int e = tid; e < n * n; e += nt) {
    int j = e / n, k = e - j * n;
    if (j < 1 || k <= j) continue;
    if (cf[j] - cf[0] < 5 || cf[k] - cf[j] < 5) continue;
    int ov = out[triple_index(0, j, k, n)];
    if (ov < lo) continue;
    int g = G[j * n + k];
    if (g == FRS_NEG_INF) continue;
    int d = -amb[pair_index(0, j, n)] + ov + g;
    if (d > my_best || (d == my_best && e < my_e)) { my_best = d; my_e = e; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int ov = __shfl_xor_sync(0xffffffffu, my_best, o);
    int oe = __shfl_xor_sync(0xffffffffu, my_e, o);
    if (ov > my_best || (ov == my_best && oe < my_e)) { my_best = ov; my_e = oe; }
  }
  if (!WARP) {
    if ((tid & 31) == 0) { red[tid >> 5] = my_best; red[DPS_MAX_WARPS + (tid >> 5)] = my_e; }
    __syncthreads();
  }
  if (tid == 0) {
    if (!WARP) {
      for (int w = 1; w < (nt >> 5); ++w) {
        int bv = red[w], be = red[DPS_MAX_WARPS + w];
        if (bv > my_best || (bv == my_best && be < my_e)) { my_best = bv; my_e = be; }
      }
    }
    int none = -amb[pair_index(0, E, n)];
    if (my_best != FRS_NEG_INF && my_best > none) {
      int j = my_e / n, k = my_e - j * n;
      final_flag[j] = 1;
      final_flag[k] = 1;
      int guard = 0;
      while (k != E) {
        int k2 = arg[j * n + k];
        if (k2 <= k || ++guard > n) { dev_fail(err, DEVERR_BACKTRACE, p); break; }
        final_flag[k2] = 1;
        j = k;
        k = k2;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K7 DP tables (+ fused solve).  One CTA = (subproblem, slab of read-rep words).  Per chunk of wc words:
//   1. TMA bulk copies stage the n coverage rows x 32*wc reps of the chunk into shared memory
//      (the next chunk's copies fly during the triple phase of the current one);
//   2. mask phase: a warp owns a row i, a lane owns a rep; for every j>i the lanes compare
//      cov = P[j]-P[i] with the pair's integer cuts and __ballot_sync packs the yea / nay bits of
//      32 reps into one word each (:488-497);
//   3. ins pass: amb(i,j) = sum_w W . ~(yea|nay)   (:500-506), one thread per pair;
//   4. triple phase: out(i,j,k) = sum_w W . [(yea_ij & nay_jk) | (nay_ij & yea_jk)]  (:509-528) as
//      weighted popcounts over the packed words; a warp owns the middle candidate j, its lanes the
//      left candidate i, the loop runs over k with broadcast shared-memory loads.
// Weights: bit-planes of W per word (plane b = reps whose weight has bit b); a word whose reps all
// have weight 1 costs one POPC.  Partial sums accumulate in shared memory over the chunks of the slab.
// ---------------------------------------------------------------------------------------------
#define DPT_MAXW 4
#define DP_MASK_J 8   // pairs (i, j0..j0+7) per mask-phase unit
#define DP_TRI_K 16   // right candidates k per triple-phase unit

struct DpArgs {
  const int* sub_start; const int* sub_n; const int* sub_tint; const int* sub_info; const i64* sub_tab_off;
  const int* tint_rep_off; const int* tint_cand_off; const i64* tint_cov_off;
  const int* rep_weight; const int* cand_flat; const u32* P;
  const double* thr_table; int thr_table_len; double tp;
  const int2* cut_tab;  // integer cuts of the lengths below CUT_TAB_N (k_cut_table)
  int lo; int keep_tables;
  int* tab;        // global tables of the split / kept subproblems
  u8* final_flag;  // [n_cand]
  int* err;
  const i64* cnt;  // device counters of the run (work counts per class, capacities check)
  Caps caps;
  const int* bases;  // [DP_CLASSES] first work item of each class
  int* cursor;       // [8 + class] work-stealing cursors, [8 + DP_CLASSES] the solver's
  int m_cap;         // upper bound of a subproblem's size (max_problem_size + 12, see k_fixed_b)
  int* sub_left;     // slabs of every split subproblem still to be added to its global tables (starts as sub_slabs)
};

// shared-memory carve-up of k_dp for (M = largest n of the launch, wc, out table on chip?)
// yea / nay bits of the two synthetic columns for a pair of span `d` samples with cuts (ty, tn): a rep that covers
// the whole window has coverage d on the pair, a rep that misses it has 0
__device__ __forceinline__ u8 dp_pair_flags(int d, int ty, int tn) {
  return (u8)((d >= ty ? 1 : 0) | (d <= tn ? 2 : 0) | (0 >= ty ? 4 : 0) | (0 <= tn ? 8 : 0));
}
// bit 0 / bit 2: the (spanning / missing) column is in out(i,j,k) = (yea_ij & nay_jk) | (nay_ij & yea_jk)
__device__ __forceinline__ u32 dp_flag_cross(u32 fij, u32 fjk) { return ((fij & (fjk >> 1)) | ((fij >> 1) & fjk)) & 5u; }

#define DP_PASS_WORDS 64                       // read-rep words classified per pass (see k_dp)
#define DP_LIVE_CAP (DP_PASS_WORDS * 32 + DPT_MAXW * 32)  // queue of live reps: a pass + the carry of the previous one
struct DpSmem {
  int tile, ynm, cf, ty, tn, pf, planes, nplanes, vmask, munit, cumu, amb, out, G, arg, live, red, total;
};
__host__ __device__ inline DpSmem dp_smem_layout(int M, int wc, int out_on_chip) {
  DpSmem s;
  int p2 = M * (M - 1) / 2;
  int c3 = M * (M - 1) * (M - 2) / 6;
  int o = 16;  // mbarrier
  s.tile = o; o += M * 32 * wc * 4;
  s.ynm = o; o += p2 * wc * 8;
  s.cf = o; o += M * 4;
  s.ty = o; o += p2 * 4;
  s.tn = o; o += p2 * 4;
  s.pf = o; o += (p2 + 3) & ~3;  // dp_pair_flags of every pair
  s.planes = o; o += wc * 32 * 4;
  s.nplanes = o; o += wc * 4;
  s.vmask = o; o += wc * 4;
  s.munit = o; o += (p2 / DP_MASK_J + M + 1) * 4;  // mask-phase units (row i, first j), ushort2
  s.cumu = o; o += (M + 1) * 4;                    // triple-phase units: prefix of j * ceil((n-1-j)/DP_TRI_K)
  s.amb = o; o += out_on_chip ? p2 * 4 : 0;
  s.out = o; o += out_on_chip ? c3 * 4 : 0;
  // G and arg are written by the solver, after the last chunk: the queue of live reps shares their bytes
  s.G = o;
  s.live = o;
  s.arg = o + (out_on_chip ? M * M * 4 : 0);
  {
    const int solver = out_on_chip ? M * M * 6 : 0, queue = DP_LIVE_CAP * 4;
    o += solver > queue ? solver : queue;
  }
  o = (o + 3) & ~3;
  s.red = o; o += 2 * DPS_MAX_WARPS * 4;
  s.total = (o + 15) & ~15;
  return s;
}

__device__ __forceinline__ int wpopc(u32 m, const u32* __restrict__ pl, int np) {
  if (np <= 1) return __popc(m);  // every rep of the word has weight 1 (m only holds valid reps)
  int acc = 0;
  for (int b = 0; b < np; ++b) acc += __popc(m & pl[b]) << b;
  return acc;
}

// mask phase of one chunk: yea / nay words of every pair (i<j) for the nw words staged in `tile`.
template <int THREADS, bool MASKED>
__device__ __forceinline__ void dp_mask_phase(const int n, const int nw, const int wc, const int CW,
                                              const u32* __restrict__ tile, const int* __restrict__ ty,
                                              const int* __restrict__ tn, uint2* __restrict__ ynm,
                                              const u32* __restrict__ vmask, const ushort2* __restrict__ munit,
                                              const int n_munit) {
  constexpr int NW = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 vl = 0xffffffffu;  // bit w: this lane's rep of word w is a real rep
  if (MASKED) {
    vl = 0;
    for (int w = 0; w < nw; ++w) vl |= ((vmask[w] >> lane) & 1u) << w;
  }
  // units = (row i, block of DP_MASK_J consecutive j): equal cost, dealt round-robin to the warps
  for (int u = warp; u < n_munit; u += NW) {
    const ushort2 un = munit[u];
    const int i = un.x, j0 = un.y, j1 = min(n, j0 + DP_MASK_J);
    const u32* rowi = tile + (size_t)i * CW + lane;
    u32 ri[DPT_MAXW];
#pragma unroll
    for (int w = 0; w < DPT_MAXW; ++w) ri[w] = (w < nw) ? rowi[w * 32] : 0u;
    int e = pair_index(i, j0, n);
    for (int j = j0; j < j1; ++j, ++e) {
      const int cy = ty[e], cn = tn[e];
      const u32* rowj = tile + (size_t)j * CW + lane;
      uint2* dst = ynm + (size_t)e * wc;
#pragma unroll
      for (int w = 0; w < DPT_MAXW; ++w) {
        if (w < nw) {
          const int cov = (int)(rowj[w * 32] - ri[w]);
          bool py = cov >= cy, pn = cov <= cn;
          if (MASKED) { const bool valid = (vl >> w) & 1u; py = py && valid; pn = pn && valid; }
          const u32 by = __ballot_sync(0xffffffffu, py);
          const u32 bn = __ballot_sync(0xffffffffu, pn);
          if (lane == 0) dst[w] = make_uint2(by, bn);
        }
      }
    }
  }
}

// triple phase of one chunk: out(i,j,k) += sum_w W . [(yea_ij & nay_jk) | (nay_ij & yea_jk)].  The
// (j, i) pairs are flattened over the threads (a warp holds consecutive i of one or two j, so the
// yn_jk loads are broadcasts and the accumulation is conflict-free); the loop runs over k.
template <int THREADS, bool W1>
__device__ __forceinline__ void dp_triple_phase(const int n, const int nw, const int wc, const uint2* __restrict__ ynm,
                                                const u32* __restrict__ planes, const int* __restrict__ nplanes,
                                                const int* __restrict__ cumu, const int out_on_chip,
                                                int* __restrict__ dst) {
  // units (j, block of DP_TRI_K right candidates, i), i fastest: cumu[j] = first unit of middle j
  const int U = cumu[n - 1];
  for (int q = threadIdx.x; q < U; q += THREADS) {
    int lo = 1, hi = n - 1;  // largest j in [1, n-2] with cumu[j] <= q
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (cumu[mid] <= q) lo = mid; else hi = mid;
    }
    const int j = lo;
    const int rem = q - cumu[j];
    const int kb = rem / j, i = rem - kb * j;
    const int k0 = j + 1 + kb * DP_TRI_K, k1 = min(n, k0 + DP_TRI_K);
    const uint2* yn_ij = ynm + (size_t)pair_index(i, j, n) * wc;
    uint2 rij[DPT_MAXW];
#pragma unroll
    for (int w = 0; w < DPT_MAXW; ++w) rij[w] = (w < nw) ? yn_ij[w] : make_uint2(0u, 0u);
    const uint2* yn_jk = ynm + (size_t)pair_index(j, k0, n) * wc;
    int o = triple_mid_off(j, n) + (k0 - j - 1) * j + i;
    for (int k = k0; k < k1; ++k, yn_jk += wc, o += j) {
      int acc = 0;
#pragma unroll
      for (int w = 0; w < DPT_MAXW; ++w) {
        if (w < nw) {
          const uint2 jk = yn_jk[w];
          const u32 m = (rij[w].x & jk.y) | (rij[w].y & jk.x);
          if (W1) acc += __popc(m);
          else if (m) acc += wpopc(m, planes + w * 32, nplanes[w]);
        }
      }
      if (acc) {
        if (out_on_chip) dst[o] += acc;
        else atomicAdd(&dst[o], acc);
      }
    }
  }
}

// Persistent: the CTAs of a class take the class's work items from a shared cursor (work stealing: items
// differ in cost by orders of magnitude) until the list is empty; the number of items is only known on
// the device.  smem_bytes = dynamic shared memory of the launch; every item carves it for its own n.
template <int THREADS>
__global__ void __launch_bounds__(THREADS, (THREADS <= 128 ? 6 : THREADS == 256 ? 3 : 1)) k_dp(DpArgs A, const DpWork* __restrict__ work_all,
                                                int cls, int smem_bytes) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ int s_n_munit, s_item, s_nq, s_wfull, s_wdead, s_last;
  if (!dp_caps_ok(A.cnt, A.caps)) return;
  const int n_work = (int)A.cnt[CNT_PLAN + PLAN_WORK + cls];
  if (n_work == 0) return;
  const DpWork* work = work_all + A.bases[cls];
  const int out_on_chip = cls < 5 ? 1 : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = THREADS / 32;
  unsigned long long* bar = (unsigned long long*)dsm;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  u32 phase = 0;
  for (;;) {
    __syncthreads();  // the previous item is done with the shared memory (and the barrier is initialised)
    if (tid == 0) s_item = atomicAdd(&A.cursor[8 + cls], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_work) break;
    const DpWork wk = work[item];
    const int p = wk.sub;
    const int n = A.sub_n[p];
    int wc = DPT_MAXW;
    while (wc > 1 && dp_smem_layout(n, wc, out_on_chip).total > smem_bytes) wc >>= 1;
    const DpSmem L = dp_smem_layout(n, wc, out_on_chip);
    if (L.total > smem_bytes || n > A.m_cap) {
      if (tid == 0) dev_fail(A.err, DEVERR_DP_SMEM, n);
      continue;
    }
    const int qs = A.sub_start[p];
    const int t = A.sub_tint[p];
    const bool fused = (A.sub_info[p] >> 8) & 1;
    const int r0 = A.tint_rep_off[t];
    const int R = A.tint_rep_off[t + 1] - r0;
    const int Rp = (R + 3) & ~3;
    const int words = (R + 31) >> 5;
    const int sw = A.sub_info[p] >> 16;  // read-rep words per CTA of this subproblem
    const int w_lo = fused ? 0 : wk.slab * sw;
    const int w_hi = fused ? words : min(words, w_lo + sw);
    const int p2 = n * (n - 1) / 2;
    const int c3 = n * (n - 1) * (n - 2) / 6;
    const int CW = 32 * wc;
    u32* tile = (u32*)(dsm + L.tile);      // [n][CW]   (16-byte aligned rows)
    uint2* ynm = (uint2*)(dsm + L.ynm);    // [p2][wc]  x = yea, y = nay
    int* cf = (int*)(dsm + L.cf);
    int* ty = (int*)(dsm + L.ty);
    int* tn = (int*)(dsm + L.tn);
    u32* planes = (u32*)(dsm + L.planes);  // [wc][32]
    int* nplanes = (int*)(dsm + L.nplanes);
    u32* vmask = (u32*)(dsm + L.vmask);
    ushort2* munit = (ushort2*)(dsm + L.munit);
    int* cumu = (int*)(dsm + L.cumu);
    int* amb_s = (int*)(dsm + L.amb);      // [p2]
    int* out_s = (int*)(dsm + L.out);      // [c3]
    const u32* Prow0 = A.P + A.tint_cov_off[t] + (i64)(qs - A.tint_cand_off[t]) * Rp;
    int* tab_g = A.tab + A.sub_tab_off[p];  // only dereferenced when the subproblem owns a block

    for (int i = tid; i < n; i += THREADS) cf[i] = A.cand_flat[qs + i];
    if (tid == 32 % THREADS) {  // unit tables of the two phases (a few hundred entries)
      int u = 0;
      for (int i = 0; i < n - 1; ++i)
        for (int j0 = i + 1; j0 < n; j0 += DP_MASK_J) munit[u++] = make_ushort2((unsigned short)i, (unsigned short)j0);
      s_n_munit = u;
      int acc = 0;
      cumu[0] = 0;
      for (int j = 1; j <= n - 1; ++j) {
        cumu[j] = acc;
        acc += j * ((n - 1 - j + DP_TRI_K - 1) / DP_TRI_K);
      }
    }
    if (out_on_chip) {
      for (int e = tid; e < p2; e += THREADS) amb_s[e] = 0;
      for (int e = tid; e < c3; e += THREADS) out_s[e] = 0;
    }
    __syncthreads();
    // cuts of every pair (the TMA copies of the first pass fly meanwhile)
    // ---- Rep compaction.  For a read rep r let c = P[last][r] - P[first][r], the samples of the whole window
    // [cf[0], cf[n-1]) it covers.  c == 0: every pair of the subproblem sees coverage 0; c == window: every pair
    // (i, j) sees cf[j] - cf[i].  All reps of one of these two kinds have the same coverage row: two synthetic
    // columns, weighted with the sums of their weights, whose yea / nay bits per pair follow from the cuts alone
    // (dp_pair_flags) and whose contribution is added in closed form after the chunk loop (exact for every -tp).
    // Only the reps in between (an alignment boundary inside the window) need their own bit.  Passes of
    // DP_PASS_WORDS words: TMA bulk copies stage the first and the last coverage row of the pass, the warps
    // classify and queue the partial reps, and the chunk loop gathers the n rows for queued reps only
    // (typically 40 % of the reps on the synthetic configs; all sums are integer, the order is free). ----
    int* live = (int*)(dsm + L.live);  // queue of tint-local rep indices
    u8* pf = dsm + L.pf;
    auto issue_rows = [&](int ws, int we) {  // first / last row of words [ws, we) -> tile rows 0 / 1 (as 2 x 2048 u32)
      const int col0 = ws * 32;
      const int cols = min(we * 32, Rp) - col0;
      const u32 bytes = (u32)cols * 4u;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(bar, bytes * 2u);
      tma_bulk_g2s(tile, Prow0 + col0, bytes, bar);
      tma_bulk_g2s(tile + DP_PASS_WORDS * 32, Prow0 + (i64)(n - 1) * Rp + col0, bytes, bar);
    };
    const bool tma_rows = 2 * DP_PASS_WORDS * 32 <= n * CW;  // the two staged rows fit the tile (n >= 16 at wc = 4 ...)
    if (tid == 0) {
      s_nq = 0;
      s_wfull = 0;
      s_wdead = 0;
      if (tma_rows && w_lo < w_hi) issue_rows(w_lo, min(w_hi, w_lo + DP_PASS_WORDS));
    }
    for (int e = tid; e < p2; e += THREADS) {
      // decode pair e -> (i, j): rows are short, a linear walk is fine (done once per item)
      int i = 0, rem = e;
      while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
      int j = i + 1 + rem;
      int a, b;
      length_cuts_t(cf[j] - cf[i] + 1, A.cut_tab, A.thr_table, A.thr_table_len, A.tp, a, b);
      ty[e] = a;
      tn[e] = b;
      pf[e] = dp_pair_flags(cf[j] - cf[i], a, b);
    }
    const u32 win = (u32)(cf[n - 1] - cf[0]);  // samples of the window: coverage of a rep that spans it
    int head = 0;                              // first queued rep not yet processed (CTA-uniform)

    // one chunk: the queued reps live[head .. head + ncol)
    auto run_chunk = [&](const int ncol) {
      const int nw = (ncol + 31) >> 5;
      // weight planes + valid mask of the chunk's words, gather of the n coverage rows
      for (int w = warp; w < nw; w += NW) {
        const int c = w * 32 + lane;
        int wt = 0;
        if (c < ncol) {
          wt = A.rep_weight[r0 + live[head + c]];
        }
        const int mx = warp_max_i(wt);
        const int np = 32 - __clz(mx);
        for (int b = 0; b < np; ++b) {
          const u32 m = __ballot_sync(0xffffffffu, (wt >> b) & 1);
          if (lane == 0) planes[w * 32 + b] = m;
        }
        const u32 vm = __ballot_sync(0xffffffffu, c < ncol);
        if (lane == 0) { nplanes[w] = np; vmask[w] = vm; }
      }
      for (int u = warp; u < n * nw; u += NW) {  // a warp fetches word w of row i: 32 queued reps
        const int i = u / nw, c = (u - i * nw) * 32 + lane;
        u32 v = 0;
        if (c < ncol) v = Prow0[(i64)i * Rp + live[head + c]];
        tile[(size_t)i * CW + c] = v;
      }
      __syncthreads();
      if (ncol & 31) dp_mask_phase<THREADS, true>(n, nw, wc, CW, tile, ty, tn, ynm, vmask, munit, s_n_munit);
      else dp_mask_phase<THREADS, false>(n, nw, wc, CW, tile, ty, tn, ynm, vmask, munit, s_n_munit);
      __syncthreads();  // masks complete, tile free
      // ---- ins pass: ambiguous reps per pair ----
      for (int e = tid; e < p2; e += THREADS) {
        int acc = 0;
        for (int w = 0; w < nw; ++w) {
          const uint2 yn = ynm[(size_t)e * wc + w];
          const u32 am = vmask[w] & ~(yn.x | yn.y);
          if (am) acc += wpopc(am, planes + w * 32, nplanes[w]);
        }
        if (acc) {
          if (out_on_chip) amb_s[e] += acc;
          else atomicAdd(&tab_g[e], acc);
        }
      }
      // ---- triple phase ----
      {
        bool w1 = true;  // every word of the chunk has unit weights only
        for (int w = 0; w < nw; ++w) w1 = w1 && nplanes[w] <= 1;
        int* dst = out_on_chip ? out_s : (tab_g + p2);
        if (w1) dp_triple_phase<THREADS, true>(n, nw, wc, ynm, planes, nplanes, cumu, out_on_chip, dst);
        else dp_triple_phase<THREADS, false>(n, nw, wc, ynm, planes, nplanes, cumu, out_on_chip, dst);
      }
      __syncthreads();  // the next chunk rewrites the weight planes, the masks and the tile
    };

    for (int ws = w_lo; ws < w_hi; ws += DP_PASS_WORDS) {
      const int we = min(w_hi, ws + DP_PASS_WORDS);
      __syncthreads();  // cuts visible (first pass); queue carry in place, tile free (later passes)
      if (tma_rows) {
        mbar_wait(bar, phase);
        phase ^= 1;
      }
      // ---- classify the reps of the pass ----
      for (int w = ws + warp; w < we; w += NW) {
        const int rep = w * 32 + lane;
        u32 c = 0;
        int wt = 0, wd = 0;
        bool part = false;
        if (rep < R) {
          if (tma_rows) c = tile[DP_PASS_WORDS * 32 + rep - ws * 32] - tile[rep - ws * 32];
          else c = Prow0[(i64)(n - 1) * Rp + rep] - Prow0[rep];
          part = c != 0u && c != win;
          if (!part) {
            const int x = A.rep_weight[r0 + rep];
            if (c == win) wt = x; else wd = x;
          }
        }
        const u32 m = __ballot_sync(0xffffffffu, part);
        int wsum = wt, dsum = wd;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
          dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        }
        int base = 0;
        if (lane == 0) {
          if (m) base = atomicAdd(&s_nq, __popc(m));
          if (wsum) atomicAdd(&s_wfull, wsum);
          if (dsum) atomicAdd(&s_wdead, dsum);
        }
        base = __shfl_sync(0xffffffffu, base, 0);
        if (part) live[base + __popc(m & ((1u << lane) - 1u))] = rep;
      }
      __syncthreads();  // queue complete, staged rows free
      const int nq = s_nq;
      const bool last = we >= w_hi;
      while (nq - head >= CW || (last && nq - head > 0)) {
        const int ncol = min(CW, nq - head);
        run_chunk(ncol);
        head += ncol;
      }
      if (!last) {  // carry the tail of the queue (< CW reps) to its front; the next pass's rows may fly now
        if (tid == 0 && tma_rows) issue_rows(we, min(w_hi, we + DP_PASS_WORDS));
        const int rest = nq - head;
        int v = 0;
        if (tid < rest) v = live[head + tid];
        __syncthreads();
        if (tid < rest) live[tid] = v;
        if (tid == 0) s_nq = rest;
        head = 0;
      }
    }
    __syncthreads();
    // ---- the two synthetic columns in closed form ----
    if (s_wfull > 0 || s_wdead > 0) {
      const int wf = s_wfull, wd = s_wdead;
      for (int e = tid; e < p2; e += THREADS) {
        const u32 f = pf[e];
        const int v = ((f & 3u) ? 0 : wf) + ((f & 12u) ? 0 : wd);
        if (v) {
          if (out_on_chip) amb_s[e] += v;
          else atomicAdd(&tab_g[e], v);
        }
      }
      int* dst = out_on_chip ? out_s : (tab_g + p2);
      for (int q = tid; q < (n - 1) * (n - 2) / 2; q += THREADS) {  // (j, i): q = j (j - 1) / 2 + i
        int j = 1, i = q;
        while (i >= j) { i -= j; ++j; }
        const u32 fij = pf[pair_index(i, j, n)];
        const u8* fjk = pf + pair_index(j, j + 1, n);
        int o = triple_mid_off(j, n) + i;
        for (int k = j + 1; k < n; ++k, ++fjk, o += j) {
          const u32 m = dp_flag_cross(fij, *fjk);
          if (m) {
            const int v = ((m & 1u) ? wf : 0) + ((m & 4u) ? wd : 0);
            if (out_on_chip) dst[o] += v;
            else atomicAdd(&dst[o], v);
          }
        }
      }
      __syncthreads();
    }

    if (!out_on_chip) continue;  // class 5: tables are already in global memory
    if (!fused) {
      // split mode: add this slab's partial tables to the global ones; the CTA that finishes the LAST slab of the
      // subproblem reads the sums back and solves the DP right away (no separate solver launch behind all the
      // table kernels: the solve of one subproblem runs beside the slabs of the others)
      for (int e = tid; e < p2; e += THREADS) { int v = amb_s[e]; if (v) atomicAdd(&tab_g[e], v); }
      for (int e = tid; e < c3; e += THREADS) { int v = out_s[e]; if (v) atomicAdd(&tab_g[p2 + e], v); }
      __threadfence();
      __syncthreads();
      if (tid == 0) s_last = atomicSub(&A.sub_left[p], 1) == 1;
      __syncthreads();
      if (!s_last) continue;
      __threadfence();
      for (int e = tid; e < p2; e += THREADS) amb_s[e] = __ldcg(&tab_g[e]);
      for (int e = tid; e < c3; e += THREADS) out_s[e] = __ldcg(&tab_g[p2 + e]);
      __syncthreads();
    } else if (A.keep_tables) {
      for (int e = tid; e < p2; e += THREADS) tab_g[e] = amb_s[e];
      for (int e = tid; e < c3; e += THREADS) tab_g[p2 + e] = out_s[e];
    }
    dp_solve<false, (THREADS <= 256 ? 8 : 16)>(n, cf, amb_s, out_s, A.lo, (int*)(dsm + L.G), (short*)(dsm + L.arg), (int*)(dsm + L.red),
                   A.final_flag + qs, A.err, p);
  }
}

// ---------------------------------------------------------------------------------------------
// K7+K8 for the many small subproblems of typical tints: ONE WARP per subproblem (n <= MAXN, at most
// DP_WARP_MAX_WORDS x 32 read reps), eight subproblems per CTA, warp-synchronous only.  Per word of
// 32 reps: lanes = reps load the n coverage rows (coalesced), every pair's yea/nay word is a pair of
// ballots, ambiguous counts and the out table accumulate in the warp's shared-memory slice, then the
// warp solves the DP in place.  No table ever leaves the SM.
// ---------------------------------------------------------------------------------------------
#define DPW_WARPS 8
template <int MAXN>
struct DpWarpSmem {
  static constexpr int P2 = MAXN * (MAXN - 1) / 2;
  static constexpr int C3 = MAXN * (MAXN - 1) * (MAXN - 2) / 6;
  u32 tile[MAXN][32];
  uint2 yn[P2];
  int2 cut[P2];  // x = ty, y = tn
  int amb[P2];
  int out[C3];
  int G[MAXN * MAXN];
  short arg[MAXN * MAXN];
  int cf[MAXN];
  int mid[MAXN];  // triple_mid_off(j, n) of the current subproblem
  u32 planes[32];
  unsigned short queue[64];  // queued reps (tint-local index)
  u8 pf[(P2 + 3) & ~3];      // per pair: bit 0 / 1 = yea / nay of a rep that spans the window, bit 2 / 3 = of one that misses it
};

template <int MAXN>
__global__ void __launch_bounds__(DPW_WARPS * 32) k_dp_warp(DpArgs A, const DpWork* __restrict__ work_all, int cls) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char dsm[];
  __shared__ uchar2 s_ji[(MAXN - 1) * (MAXN - 2) / 2];  // (middle j, left i) of every lane slot of the triple phase
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (!dp_caps_ok(A.cnt, A.caps)) return;
  const int n_work = (int)A.cnt[CNT_PLAN + PLAN_WORK + cls];
  const DpWork* work = work_all + A.bases[cls];
  for (int q = threadIdx.x; q < (MAXN - 1) * (MAXN - 2) / 2; q += DPW_WARPS * 32) {
    int j = 1, rem = q;
    while (rem >= j) { rem -= j; ++j; }
    s_ji[q] = make_uchar2((unsigned char)j, (unsigned char)rem);
  }
  __syncthreads();
 for (;;) {  // persistent warps: items from the class's cursor until the list is empty
  int item = 0;
  if (lane == 0) item = atomicAdd(&A.cursor[8 + cls], 1);
  item = __shfl_sync(0xffffffffu, item, 0);
  if (item >= n_work) break;
  __syncwarp();
  DpWarpSmem<MAXN>& S = reinterpret_cast<DpWarpSmem<MAXN>*>(dsm)[warp];
  const int p = work[item].sub;
  const int n = A.sub_n[p];
  const int qs = A.sub_start[p];
  const int t = A.sub_tint[p];
  const int r0 = A.tint_rep_off[t];
  const int R = A.tint_rep_off[t + 1] - r0;
  const int Rp = (R + 3) & ~3;
  const int words = (R + 31) >> 5;
  const int p2 = n * (n - 1) / 2;
  const int c3 = n * (n - 1) * (n - 2) / 6;
  const u32* Prow0 = A.P + A.tint_cov_off[t] + (i64)(qs - A.tint_cand_off[t]) * Rp;
  if (lane < n) S.cf[lane] = A.cand_flat[qs + lane];
  __syncwarp();
  for (int e = lane; e < p2; e += 32) {
    int i = 0, rem = e;
    while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
    int j = i + 1 + rem;
    int a, b;
    length_cuts_t(S.cf[j] - S.cf[i] + 1, A.cut_tab, A.thr_table, A.thr_table_len, A.tp, a, b);
    S.cut[e] = make_int2(a, b);
    S.amb[e] = 0;
    S.pf[e] = dp_pair_flags(S.cf[j] - S.cf[i], a, b);
  }
  if (lane < n) S.mid[lane] = triple_mid_off(lane, n);
  for (int e = lane; e < c3; e += 32) S.out[e] = 0;
  const int npair_ij = (n - 2) * (n - 1) / 2;  // (j, i) with 1 <= j <= n-2, i < j
  // Rep compaction (see k_dp): the reps that miss the window [cf[0], cf[n-1]) and the reps that span it are two
  // synthetic columns (added in closed form after the loop), only the reps in between are queued; a word of
  // 32 QUEUED reps is processed whenever the queue holds one.
  const u32 win = (u32)(S.cf[n - 1] - S.cf[0]);
  const u32* Plast = Prow0 + (i64)(n - 1) * Rp;
  int nq = 0, wfull = 0, wdead = 0;
  // one compacted word: the queued reps queue[0 .. cnt)
  auto process = [&](const int cnt) {
    const bool valid = lane < cnt;
    const int rep = valid ? (int)S.queue[lane] : 0;
    const int wt = valid ? A.rep_weight[r0 + rep] : 0;
    const int np = 32 - __clz(warp_max_i(wt));
    const u32 vm = __ballot_sync(0xffffffffu, valid);
    __syncwarp();  // previous word's readers are done with tile / yn / planes
    if (np > 1)
      for (int b = 0; b < np; ++b) {
        u32 m = __ballot_sync(0xffffffffu, (wt >> b) & 1);
        if (lane == 0) S.planes[b] = m;
      }
    for (int i = 0; i < n; ++i) S.tile[i][lane] = valid ? Prow0[(i64)i * Rp + rep] : 0u;
    __syncwarp();
    // masks
    int e = 0;
    for (int i = 0; i < n - 1; ++i) {
      const u32 ri = S.tile[i][lane];
#pragma unroll 4
      for (int j = i + 1; j < n; ++j, ++e) {
        const int cov = (int)(S.tile[j][lane] - ri);
        const int2 cut = S.cut[e];
        const u32 by = __ballot_sync(0xffffffffu, valid && cov >= cut.x);
        const u32 bn = __ballot_sync(0xffffffffu, valid && cov <= cut.y);
        if (lane == 0) S.yn[e] = make_uint2(by, bn);
      }
    }
    __syncwarp();
    // ambiguous counts
    for (int q = lane; q < p2; q += 32) {
      const uint2 yn = S.yn[q];
      const u32 am = vm & ~(yn.x | yn.y);
      if (am) S.amb[q] += wpopc(am, S.planes, np);
    }
    // triples: lanes over (j, i), loop over k
    for (int q = lane; q < npair_ij; q += 32) {
      const uchar2 ji = s_ji[q];  // q = j (j - 1) / 2 + i, independent of n
      const int j = ji.x, i = ji.y;
      const uint2 ij = S.yn[pair_index(i, j, n)];
      const uint2* jk = S.yn + pair_index(j, j + 1, n);
      int* o = S.out + S.mid[j] + i;
      for (int k = j + 1; k < n; ++k, ++jk, o += j) {
        const uint2 v = *jk;
        const u32 m = (ij.x & v.y) | (ij.y & v.x);
        if (m) *o += wpopc(m, S.planes, np);
      }
    }
    __syncwarp();
  };
  // removes the first cnt entries of the queue
  auto pop = [&](const int cnt) {
    const int rest = nq - cnt;
    const unsigned short v = lane < rest ? S.queue[cnt + lane] : (unsigned short)0;
    __syncwarp();
    if (lane < rest) S.queue[lane] = v;
    nq = rest;
    __syncwarp();
  };
  for (int w = 0; w < words; ++w) {
    const int rep = w * 32 + lane;
    int wt = 0, wd = 0;
    bool part = false;
    if (rep < R) {
      const u32 c = Plast[rep] - Prow0[rep];
      part = c != 0u && c != win;
      if (!part) {
        const int x = A.rep_weight[r0 + rep];
        if (c == win) wt = x; else wd = x;
      }
    }
    const u32 m = __ballot_sync(0xffffffffu, part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wt += __shfl_xor_sync(0xffffffffu, wt, o);
      wd += __shfl_xor_sync(0xffffffffu, wd, o);
    }
    wfull += wt;
    wdead += wd;
    if (part) S.queue[nq + __popc(m & ((1u << lane) - 1u))] = (unsigned short)rep;
    nq += __popc(m);
    __syncwarp();
    if (nq >= 32) {
      process(32);
      pop(32);
    }
  }
  while (nq > 0) {
    const int cnt = min(32, nq);
    process(cnt);
    pop(cnt);
  }
  // the two synthetic columns in closed form: their yea / nay bits per pair are the flags of S.pf
  if (wfull > 0 || wdead > 0) {
    __syncwarp();
    for (int q = lane; q < p2; q += 32) {
      const u32 f = S.pf[q];
      S.amb[q] += ((f & 3u) ? 0 : wfull) + ((f & 12u) ? 0 : wdead);
    }
    for (int q = lane; q < npair_ij; q += 32) {
      const uchar2 ji = s_ji[q];
      const int j = ji.x, i = ji.y;
      const u32 fij = S.pf[pair_index(i, j, n)];
      const u8* fjk = S.pf + pair_index(j, j + 1, n);
      int* o = S.out + S.mid[j] + i;
      for (int k = j + 1; k < n; ++k, ++fjk, o += j) {
        const u32 m = dp_flag_cross(fij, *fjk);
        if (m) *o += ((m & 1u) ? wfull : 0) + ((m & 4u) ? wdead : 0);
      }
    }
  }
  __syncwarp();
  if (A.keep_tables) {
    int* tab_g = A.tab + A.sub_tab_off[p];
    for (int e = lane; e < p2; e += 32) tab_g[e] = S.amb[e];
    for (int e = lane; e < c3; e += 32) tab_g[p2 + e] = S.out[e];
  }
  dp_solve<true, (MAXN <= 8 ? 4 : 2)>(n, S.cf, S.amb, S.out, A.lo, S.G, S.arg, nullptr, A.final_flag + qs, A.err, p);
  __syncwarp();
 }
}

// K8 for split subproblems: tables summed in global memory by the slab CTAs of k_dp; staged into
// shared memory when they fit (n <= DP_SMEM_MAX_N), the sweep is latency-bound on table reads.
#define DPS_THREADS 256
__global__ void __launch_bounds__(DPS_THREADS) k_dp_solve(DpArgs A, const int* __restrict__ split_list, int max_n,
                                                          int stage_max_n) {
  pdl_prologue();
  extern __shared__ __align__(16) int ssm[];
  __shared__ int s_item;
  if (!dp_caps_ok(A.cnt, A.caps)) return;
  const int n_split = (int)A.cnt[CNT_PLAN + PLAN_SPLIT];
  int* G = ssm;                         // [n][n]
  int* cf = G + max_n * max_n;          // [n]
  int* red = cf + max_n;                // [2*DPS_MAX_WARPS]
  short* arg = (short*)(red + 2 * DPS_MAX_WARPS);  // [n][n]
  int* stab = (int*)(arg + ((max_n * max_n + 1) & ~1));
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_item = atomicAdd(&A.cursor[8 + DP_CLASSES], 1);
    __syncthreads();
    const int item = s_item;
    if (item >= n_split) break;
    const int p = split_list[item];
    const int n = A.sub_n[p], qs = A.sub_start[p];
    if (n > max_n) {
      if (threadIdx.x == 0) dev_fail(A.err, DEVERR_DP_SMEM, n);
      continue;
    }
    const int* tab = A.tab + A.sub_tab_off[p];
    const int p2 = n * (n - 1) / 2;
    for (int i = threadIdx.x; i < n; i += DPS_THREADS) cf[i] = A.cand_flat[qs + i];
    if (n <= stage_max_n) {
      const int tot = p2 + n * (n - 1) * (n - 2) / 6;
      for (int e = threadIdx.x; e < tot; e += DPS_THREADS) stab[e] = tab[e];
      tab = stab;
    }
    __syncthreads();
    dp_solve<false, 4>(n, cf, tab, tab + p2, A.lo, G, arg, red, A.final_flag + qs, A.err, p);
  }
}
__host__ inline size_t dps_smem_bytes(int max_n, int stage_max_n) {
  size_t b = (size_t)max_n * max_n * 4 + (size_t)max_n * 4 + 2 * DPS_MAX_WARPS * 4 + (size_t)((max_n * max_n + 1) & ~1) * 2;
  size_t m = stage_max_n;
  b += (m * (m - 1) / 2 + m * (m - 1) * (m - 2) / 6) * 4;
  return b + 16;
}
