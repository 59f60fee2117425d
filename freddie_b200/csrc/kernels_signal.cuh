// kernels_signal.cuh -- splice signal, Gaussian stencil, variance threshold, candidate peaks.
// Reference steps: process_splicing_data (freddie_segment.py:648-678), gaussian_filter1d (:755),
// variance threshold (:757-759), candidates_from_peaks (:615-621).
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// K1 signal: warp-aggregated shared-memory histogram.
// One CTA = (tint, window of <= SIG_BINS samples, chunk of <= SIG_REPS read reps).  Every endpoint
// of the chunk that falls in the window is counted in shared memory (lanes that hit the same bin
// with weight 1 are merged with __match_any_sync and counted once), then the non-zero bins are
// flushed to the global int32 signal (plain store when the tint has a single chunk, else RED).
// ---------------------------------------------------------------------------------------------
#define SIG_BINS 16384
#define SIG_REPS 2048
#define SIG_THREADS 256

struct SigWork { int tint; int win_lo; int win_hi; int rep_lo; int rep_hi; int single; };

__device__ __forceinline__ void hist_add(int* hist, int bin, int w, bool active) {
  unsigned m = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  bool uni = __all_sync(m, w == 1);
  if (uni) {
    unsigned peers = __match_any_sync(m, bin);
    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
  } else {
    atomicAdd(&hist[bin], w);
  }
}

__global__ void __launch_bounds__(SIG_THREADS) k_signal(const SigWork* __restrict__ work,
                                                       const int* __restrict__ rep_iv_off,
                                                       const int* __restrict__ rep_weight,
                                                       const int* __restrict__ iv_fs, const int* __restrict__ iv_fe,
                                                       int ignore_ends, int* __restrict__ y_raw) {
  extern __shared__ int hist[];  // SIG_BINS ints (dynamic: above the 48 KB static limit)
  const SigWork wk = work[blockIdx.x];
  const int nb = wk.win_hi - wk.win_lo;
  for (int b = threadIdx.x; b < nb; b += SIG_THREADS) hist[b] = 0;
  __syncthreads();
  // one lane per rep, intervals walked in lock step so that lanes hit the same splice site together
  const int n_rep = wk.rep_hi - wk.rep_lo;
  const int n_round = (n_rep + SIG_THREADS - 1) / SIG_THREADS;
  for (int rd = 0; rd < n_round; ++rd) {
    int r = wk.rep_lo + rd * SIG_THREADS + threadIdx.x;
    int a = 0, b = 0, w = 0;
    if (r < wk.rep_hi) { a = rep_iv_off[r]; b = rep_iv_off[r + 1]; w = rep_weight[r]; }
    int m = b - a;
    int mmax = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mmax = max(mmax, __shfl_xor_sync(0xffffffffu, mmax, o));
    for (int k = 0; k < mmax; ++k) {
      bool have = k < m;
      int fs = have ? iv_fs[a + k] : -1;
      int fe = have ? iv_fe[a + k] : -1;
      bool use_s = have && !(ignore_ends && k == 0) && fs >= wk.win_lo && fs < wk.win_hi;
      bool use_e = have && !(ignore_ends && k == m - 1) && fe >= wk.win_lo && fe < wk.win_hi;
      hist_add(hist, fs - wk.win_lo, w, use_s);
      hist_add(hist, fe - wk.win_lo, w, use_e);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += SIG_THREADS) {
    int v = hist[b];
    if (wk.single) y_raw[wk.win_lo + b] = v;
    else if (v) atomicAdd(&y_raw[wk.win_lo + b], v);
  }
}

// ---------------------------------------------------------------------------------------------
// K2 Gaussian: shared-memory tiled fp64 stencil over one island tile.
// Evaluation order of scipy's correlate1d symmetric branch:  acc = y[l]*w[c];
// for jj=-lw..-1: acc = acc + (y[l+jj] + y[l-jj]) * w[c+jj], each op rounded separately (no FMA).
// Reflect extension 'd c b a | a b c d | d c b a', valid when the island is shorter than lw.
// ---------------------------------------------------------------------------------------------
#define TILE_SAMPLES 1024
#define GAUSS_THREADS 128
#define GAUSS_OPT 8  // consecutive outputs per thread (register sliding window)
#define GAUSS_DB 4   // tap distances per register block

struct TileWork { int island; int lo; };  // lo = island-local first sample of the tile

// Taps are evaluated outermost pair first, like scipy.  The radius is padded to a multiple of
// GAUSS_DB with zero weights: those pairs come first and add (a+b)*0 = +0 to a non-negative
// accumulator, which leaves every bit unchanged.  Thread t owns outputs 8t..8t+7; for a block of 4
// distances it needs 11 + 11 consecutive inputs, so a tap pair costs ~0.7 shared-memory loads instead
// of 3.  The tile is stored with a skew of one double per 8 so that the stride-8 accesses of a warp
// are bank-conflict free.
__host__ __device__ __forceinline__ int gauss_pad_radius(int lw) { return (lw + GAUSS_DB - 1) / GAUSS_DB * GAUSS_DB; }
__host__ __device__ __forceinline__ int gauss_skew(int q) { return q + (q >> 3); }
__host__ __device__ inline size_t gauss_smem_bytes(int lw) {
  int lwp = gauss_pad_radius(lw);
  return (size_t)(lwp + 1 + gauss_skew(TILE_SAMPLES + 2 * lwp) + 2) * 8;
}

__global__ void __launch_bounds__(GAUSS_THREADS) k_gauss(const TileWork* __restrict__ tiles,
                                                        const int* __restrict__ island_sample_off,
                                                        const int* __restrict__ y_raw,
                                                        const double* __restrict__ gw, int lw,
                                                        double* __restrict__ y) {
  extern __shared__ double gsm[];
  const int lwp = gauss_pad_radius(lw);
  double* wd = gsm;              // wd[d] = weight of the pair at distance d (0 for the padding)
  double* ext = gsm + lwp + 1;   // skewed tile + halo
  const TileWork tw = tiles[blockIdx.x];
  const int f0 = island_sample_off[tw.island];
  const int n = island_sample_off[tw.island + 1] - f0;
  const int cnt = min(TILE_SAMPLES, n - tw.lo);
  for (int d = threadIdx.x; d <= lwp; d += GAUSS_THREADS) wd[d] = (d <= lw) ? gw[lw - d] : 0.0;
  const int span = cnt + 2 * lwp;
  const int first = tw.lo - lwp;
  if (first >= 0 && first + span <= n) {  // interior tile: no reflection
    const int* src = y_raw + f0 + first;
    for (int s = threadIdx.x; s < span; s += GAUSS_THREADS) ext[gauss_skew(s)] = (double)src[s];
  } else {
    const int n2 = 2 * n;
    for (int s = threadIdx.x; s < span; s += GAUSS_THREADS) {
      int j = (first + s) % n2;
      if (j < 0) j += n2;
      if (j >= n) j = n2 - 1 - j;
      ext[gauss_skew(s)] = (double)y_raw[f0 + j];
    }
  }
  __syncthreads();
  const int x0 = threadIdx.x * GAUSS_OPT;
  if (x0 >= cnt) return;
  // logical index of output i's centre: lwp + x0 + i ; all offsets below are warp-uniform + 8*t
  auto at = [&](int off) -> double { return ext[x0 + threadIdx.x + off + (off >> 3)]; };  // skew(8t+off) = 9t+off+(off>>3)
  double acc[GAUSS_OPT];
  const double w0 = wd[0];
#pragma unroll
  for (int i = 0; i < GAUSS_OPT; ++i) acc[i] = __dmul_rn(at(lwp + i), w0);
  for (int d0 = lwp; d0 >= GAUSS_DB; d0 -= GAUSS_DB) {
    double bl[GAUSS_OPT + GAUSS_DB - 1], br[GAUSS_OPT + GAUSS_DB - 1];
#pragma unroll
    for (int m = 0; m < GAUSS_OPT + GAUSS_DB - 1; ++m) {
      bl[m] = at(lwp - d0 + m);                   // input  x0 + m - d0
      br[m] = at(lwp + d0 - (GAUSS_DB - 1) + m);  // input  x0 + m + d0 - (DB-1)
    }
#pragma unroll
    for (int sft = 0; sft < GAUSS_DB; ++sft) {
      const double w = wd[d0 - sft];
#pragma unroll
      for (int i = 0; i < GAUSS_OPT; ++i)
        acc[i] = __dadd_rn(acc[i], __dmul_rn(__dadd_rn(bl[i + sft], br[i - sft + GAUSS_DB - 1]), w));
    }
  }
  double* dst = y + f0 + tw.lo + x0;
  if (x0 + GAUSS_OPT <= cnt && ((((size_t)dst) & 15) == 0)) {
#pragma unroll
    for (int i = 0; i < GAUSS_OPT; i += 2) *reinterpret_cast<double2*>(dst + i) = make_double2(acc[i], acc[i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < GAUSS_OPT; ++i)
      if (x0 + i < cnt) dst[i] = acc[i];
  }
}

// ---------------------------------------------------------------------------------------------
// K4 candidates: strict local maxima with plateau -> floor midpoint, ends never peaks, plus the
// first and last sample of every island (scipy _local_maxima_1d; candidates_from_peaks :615-621).
// Writes byte flags; the ordered list comes from the generic compaction.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GAUSS_THREADS) k_peaks(const TileWork* __restrict__ tiles,
                                                        const int* __restrict__ island_sample_off,
                                                        const double* __restrict__ y, u8* __restrict__ flag) {
  const TileWork tw = tiles[blockIdx.x];
  const int f0 = island_sample_off[tw.island];
  const int n = island_sample_off[tw.island + 1] - f0;
  const int cnt = min(TILE_SAMPLES, n - tw.lo);
  const double* yi = y + f0;
  for (int t = threadIdx.x; t < cnt; t += GAUSS_THREADS) {
    int x = tw.lo + t;
    if (x == 0 || x == n - 1) { flag[f0 + x] = 1; continue; }
    double v = yi[x];
    if (yi[x - 1] < v) {
      int ia = x + 1;
      while (ia < n - 1 && yi[ia] == v) ++ia;
      if (yi[ia] < v) flag[f0 + ((x + ia - 1) >> 1)] = 1;
    }
  }
}

// after compaction: per candidate rank q -> island id, and the island / tint offset tables
__global__ void k_cand_meta(const int* __restrict__ cand_flat, int n_cand, const int* __restrict__ island_sample_off,
                            int n_islands, int* __restrict__ cand_island, int* __restrict__ island_cand_off) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_cand) {
    if (q == n_cand) island_cand_off[n_islands] = n_cand;
    return;
  }
  int f = cand_flat[q];
  int isl = upper_row(island_sample_off, n_islands, f);
  cand_island[q] = isl;
  if (f == island_sample_off[isl]) island_cand_off[isl] = q;
}

// ---------------------------------------------------------------------------------------------
// K3 variance threshold (one CTA per tint): ordered compaction of the positive smoothed samples,
// then numpy's pairwise summation tree (DOUBLE_pairwise_sum) for mean and variance:
//   n < 8: sequential; n <= 128: eight strided accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//   remainder added sequentially; else split at n/2 rounded down to a multiple of 8.
// Leaves (<=128 elements) are summed in parallel, the tree is combined in numpy's order by thread 0.
// ---------------------------------------------------------------------------------------------
#define THR_THREADS 256

// SQ: sum of (a[i]-mean)^2 instead of a[i] (numpy evaluates x = a - mean, then x*x, then the same tree)
template <bool SQ>
__device__ __forceinline__ double pw_term(double v, double mean) {
  if (!SQ) return v;
  double d = __dsub_rn(v, mean);
  return __dmul_rn(d, d);
}
template <bool SQ>
__device__ double pw_leaf(const double* a, int n, double mean) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res = __dadd_rn(res, pw_term<SQ>(a[i], mean));
    return res;
  }
  double r[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = pw_term<SQ>(a[k], mean);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = __dadd_rn(r[k], pw_term<SQ>(a[i + k], mean));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; ++i) res = __dadd_rn(res, pw_term<SQ>(a[i], mean));
  return res;
}

// thread 0: enumerate leaves of the pairwise tree in order
__device__ int pw_leaves(int n, int* leaf_off, int* leaf_len) {
  int stack_off[40], stack_len[40];
  int sp = 0, nl = 0;
  stack_off[0] = 0; stack_len[0] = n; sp = 1;
  while (sp > 0) {
    --sp;
    int o = stack_off[sp], l = stack_len[sp];
    if (l <= 128) { leaf_off[nl] = o; leaf_len[nl] = l; ++nl; continue; }
    int n2 = l / 2;
    n2 -= n2 % 8;
    // right pushed first so that the left half is expanded first (in-order leaves)
    stack_off[sp] = o + n2; stack_len[sp] = l - n2; ++sp;
    stack_off[sp] = o; stack_len[sp] = n2; ++sp;
  }
  return nl;
}

// thread 0: combine leaf sums following the recursion  pw(l) = pw(left) + pw(right)
__device__ double pw_combine(int n, const double* leaf_sum) {
  // iterative post-order: frames hold (len, state, left value)
  int f_len[40];
  int f_state[40];
  double f_left[40];
  int sp = 0, next_leaf = 0;
  double ret = 0.0;
  f_len[0] = n; f_state[0] = 0; sp = 1;
  while (sp > 0) {
    int t = sp - 1;
    int l = f_len[t];
    if (f_state[t] == 0) {
      if (l <= 128) { ret = leaf_sum[next_leaf++]; --sp; continue; }
      int n2 = l / 2;
      n2 -= n2 % 8;
      f_state[t] = 1;
      f_len[sp] = n2; f_state[sp] = 0; ++sp;
    } else if (f_state[t] == 1) {
      f_left[t] = ret;
      int n2 = l / 2;
      n2 -= n2 % 8;
      f_state[t] = 2;
      f_len[sp] = l - n2; f_state[sp] = 0; ++sp;
    } else {
      ret = __dadd_rn(f_left[t], ret);
      --sp;
    }
  }
  return ret;
}

__global__ void __launch_bounds__(THR_THREADS) k_threshold(const int* __restrict__ tint_order,
                                                          const int* __restrict__ tint_island_off,
                                                          const int* __restrict__ island_sample_off,
                                                          const double* __restrict__ y, double vf,
                                                          double* __restrict__ vbuf, int* __restrict__ leaf_off,
                                                          int* __restrict__ leaf_len, double* __restrict__ leaf_sum,
                                                          double* __restrict__ thr) {
  __shared__ int sm_scan[40];
  __shared__ int sm_nl;
  __shared__ double sm_mean;
  const int t = tint_order[blockIdx.x];  // largest tints first: the longest CTA must not start last
  const int s0 = island_sample_off[tint_island_off[t]];
  const int s1 = island_sample_off[tint_island_off[t + 1]];
  // ordered compaction of positives into vbuf[s0 ...]: 8 consecutive samples per thread, one block
  // scan per 2048 samples
  int base = 0;
  for (int off = s0; off < s1; off += THR_THREADS * 8) {
    const int i0 = off + threadIdx.x * 8;
    double v[8];
    int p = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] = (i0 + k < s1) ? y[i0 + k] : 0.0;
      p += (v[k] > 0.0) ? 1 : 0;
    }
    int tot;
    int ex = block_exclusive_scan<int>(p, &tot, sm_scan);
    double* dst = vbuf + s0 + base + ex;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (v[k] > 0.0) *dst++ = v[k];
    base += tot;
  }
  const int n = base;
  if (n == 0) {
    if (threadIdx.x == 0) thr[t] = __longlong_as_double(0x7ff8000000000000LL);  // NaN (:757-759, empty mean)
    return;
  }
  // per-tint leaf scratch: leaves have > 64 elements once n > 128, so n/64 + 2 slots suffice
  const int lbase = s0 / 64 + 2 * t;
  int* lo = leaf_off + lbase;
  int* ll = leaf_len + lbase;
  double* ls = leaf_sum + lbase;
  double* v = vbuf + s0;
  __syncthreads();
  if (threadIdx.x == 0) sm_nl = pw_leaves(n, lo, ll);
  __syncthreads();
  const int nl = sm_nl;
  for (int k = threadIdx.x; k < nl; k += THR_THREADS) ls[k] = pw_leaf<false>(v + lo[k], ll[k], 0.0);
  __syncthreads();
  if (threadIdx.x == 0) sm_mean = __ddiv_rn(pw_combine(n, ls), (double)n);
  __syncthreads();
  const double mean = sm_mean;
  for (int k = threadIdx.x; k < nl; k += THR_THREADS) ls[k] = pw_leaf<true>(v + lo[k], ll[k], mean);
  __syncthreads();
  if (threadIdx.x == 0) {
    double var = __ddiv_rn(pw_combine(n, ls), (double)n);
    thr[t] = __dadd_rn(mean, __dmul_rn(vf, __dsqrt_rn(var)));
  }
}
