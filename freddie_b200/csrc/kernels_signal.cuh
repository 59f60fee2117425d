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

#define SIG_DIRECT_REPS 256  // reps per CTA in direct mode (one per thread)

// single: 0 = histogram, RED flush (the tint has several rep chunks); 1 = histogram, plain store;
// 2 = DIRECT: the tints have fewer than 8 endpoints per sample (typical tints: ~0.1), so zeroing and
// flushing a 16 k-bin histogram per window costs more than the endpoints themselves -- the CTA adds the
// endpoints of 256 reps (any tints: flat samples need no tint) straight to the zeroed global signal,
// with the same warp aggregation (one RED per distinct site and warp).  Launched without shared memory.
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
  pdl_prologue();
  extern __shared__ int hist[];  // SIG_BINS ints (dynamic: above the 48 KB static limit)
  const SigWork wk = work[blockIdx.x];
  const int nb = wk.win_hi - wk.win_lo;
  const bool direct = wk.single == 2;
  int* const yg = y_raw + wk.win_lo;
  if (!direct) {
    for (int b = threadIdx.x; b < nb; b += SIG_THREADS) hist[b] = 0;
    __syncthreads();
  }
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
    for (int k0 = 0; k0 < mmax; k0 += 4) {  // four intervals' loads in flight per round trip
      int fs4[4], fe4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const bool have = k0 + u < m;
        fs4[u] = have ? iv_fs[a + k0 + u] : -1;
        fe4[u] = have ? iv_fe[a + k0 + u] : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u;
        if (k >= mmax) break;  // warp-uniform
        const bool have = k < m;
        const int fs = fs4[u], fe = fe4[u];
        bool use_s = have && !(ignore_ends && k == 0) && fs >= wk.win_lo && fs < wk.win_hi;
        bool use_e = have && !(ignore_ends && k == m - 1) && fe >= wk.win_lo && fe < wk.win_hi;
        if (direct) {  // CTA-uniform; kept apart so that the histogram path compiles to shared-memory atomics
          hist_add(yg, fs - wk.win_lo, w, use_s);
          hist_add(yg, fe - wk.win_lo, w, use_e);
        } else {
          hist_add(hist, fs - wk.win_lo, w, use_s);
          hist_add(hist, fe - wk.win_lo, w, use_e);
        }
      }
    }
  }
  if (direct) return;
  __syncthreads();
  for (int b = threadIdx.x; b < nb; b += SIG_THREADS) {
    int v = hist[b];
    if (wk.single) y_raw[wk.win_lo + b] = v;
    else if (v) atomicAdd(&y_raw[wk.win_lo + b], v);
  }
}

// ---------------------------------------------------------------------------------------------
// K2 + K4 + first half of K3 in one pass over the signal (k_smooth), one CTA per island tile:
//   1. fp64 Gaussian of the tile, evaluation order of scipy's correlate1d symmetric branch:
//      acc = y[l]*w[c];  for jj=-lw..-1: acc = acc + (y[l+jj] + y[l-jj]) * w[c+jj], each op rounded
//      separately (no FMA).  Reflect extension 'd c b a | a b c d | d c b a', also valid when the
//      island is shorter than lw.  The raw signal is SPARSE (splice sites): a pair whose two inputs
//      are zero adds (0+0)*w = +0 to a non-negative accumulator and changes no bit, so every output
//      only visits the pairs that have a non-zero input, outermost first, found with a bit mask of
//      the non-zero staged samples (~3 pairs instead of 20 at sigma = 5).
//   2. candidates of the tile: strict local maxima with plateau -> floor midpoint, island ends never
//      peaks, plus the first and last sample of every island (scipy _local_maxima_1d;
//      candidates_from_peaks :615-621), decided per OWNING sample so that tiles never write into
//      each other:  m is a peak  <=>  y[m] > 0, the maximal plateau [a, b] of value y[m] around m has
//      1 <= a, b <= n-2, y[a-1] < y[m] > y[b+1], and m == (a+b)>>1.
//   3. per 32 samples one ballot word of candidates and one of positive samples (the input of the
//      variance threshold, :757-759), plus the two counts of the tile.
// k_tile_lists then writes the ordered candidate list and the ordered positive samples from the
// masks (its offsets: k_tile_prefix, a two-level sum of the tile counts, no device-wide scan).  HBM traffic
// of the two: 4 B read + 8 B written per sample, 1/4 B of masks, and the positive samples once more.
// (Tried on B200 and dropped, byte-identical but slower: ordered lists by decoupled look-back inside k_smooth;
// one persistent launch with a grid barrier between the two phases, 382 us against 263 + 91 us; the same with
// warp-level quarter tiles, 430-470 us -- the per-tile fixed cost, ~400 warp instructions, is what dominates,
// and smaller units pay it more often.)
// ---------------------------------------------------------------------------------------------
#define TILE_SAMPLES 1024
#define GAUSS_THREADS 128
#define TILE_WORDS (TILE_SAMPLES / 32)

// lo = island-local first sample of the tile; f0 / n = first flat sample / length of the island (copied
// here so that a CTA learns its geometry from ONE 16-byte load instead of a chain of two)
struct __align__(16) TileWork { int island; int lo; int f0; int n; };

// staged window: logical index s = island sample lo - lw - 1 + s (one extra sample on both sides: the
// neighbours of the tile's first and last output); centre of tile sample x = lw + 1 + x
struct P1Smem { int wd, ext, yout, nz, red, total; };
__host__ __device__ inline P1Smem p1_smem_layout(int lw) {
  const int span = TILE_SAMPLES + 2 * lw + 2;
  P1Smem s;
  int o = 0;
  s.wd = o; o += (lw + 1) * 8;
  s.ext = o; o += ((span + 2) * 4 + 7) & ~7;
  s.yout = o; o += (TILE_SAMPLES + 2) * 8;
  s.nz = o; o += ((span + 127) / 128 * 4 + 2) * 4;
  s.red = o; o += 16 * 4;
  s.total = (o + 15) & ~15;
  return s;
}

__device__ __forceinline__ u32 nz_bit(const u32* nz, int s) { return (nz[s >> 5] >> (s & 31)) & 1u; }
// 32 mask bits starting at bit position s (s >= 0)
__device__ __forceinline__ u32 nz_word(const u32* nz, int s) {
  const int w = s >> 5, b = s & 31;
  return b ? __funnelshift_r(nz[w], nz[w + 1], b) : nz[w];
}

// y of the sample whose staged centre index is c.  wd[d] = weight of the pair at distance d.
// The staged samples are the raw int32 counts: (double)a + (double)b == (double)(a + b) exactly for counts,
// so a pair costs one conversion, and the tile takes half the shared memory of doubles (16 CTAs per SM).
__device__ __forceinline__ double gauss_sparse(const int* ext, const u32* nz, const double* wd, int lw, int c) {
  double acc = __dmul_rn((double)ext[c], wd[0]);
  if (lw == 0) return acc;
  if (lw <= 32) {
    // bit (d-1) of m: the pair at distance d has a non-zero input
    const u32 keep = lw == 32 ? 0xffffffffu : ((1u << lw) - 1u);
    const u32 right = nz_word(nz, c + 1) & keep;                     // bit d-1 = sample c+d
    const u32 left = __brev(nz_word(nz, c - lw)) >> (32 - lw);       // bit lw-d -> bit d-1 = sample c-d
    u32 m = right | (left & keep);
    while (m) {
      const int d = 32 - __clz(m);  // outermost remaining pair first
      m &= ~(1u << (d - 1));
      acc = __dadd_rn(acc, __dmul_rn((double)(ext[c - d] + ext[c + d]), wd[d]));
    }
  } else {
    for (int d = lw; d >= 1; --d)
      if (nz_bit(nz, c - d) | nz_bit(nz, c + d))
        acc = __dadd_rn(acc, __dmul_rn((double)(ext[c - d] + ext[c + d]), wd[d]));
  }
  return acc;
}

// y at island-local sample x straight from the raw signal in global memory (plateaus that leave the
// tile: rare).  Same operation order; zero pairs are not skipped, which gives the same bits.
__device__ double gauss_global(const int* __restrict__ yr, int n, int x, const double* wd, int lw) {
  const int n2 = 2 * n;
  auto g = [&](int i) -> double {
    int j = i % n2;
    if (j < 0) j += n2;
    if (j >= n) j = n2 - 1 - j;
    return (double)yr[j];
  };
  double acc = __dmul_rn(g(x), wd[0]);
  for (int d = lw; d >= 1; --d) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(g(x - d), g(x + d)), wd[d]));
  return acc;
}

#define TILE_GROUP 1024  // tiles per group of the two-level count prefix

template <int MINB>  // CTAs per SM the register budget is cut for (16 -> 32 registers, 12 -> 40, 10 -> 48)
__global__ void __launch_bounds__(GAUSS_THREADS, MINB) k_smooth(const TileWork* __restrict__ tiles,
                                                         const int* __restrict__ island_sample_off,
                                                         const int* __restrict__ y_raw,
                                                         const double* __restrict__ gw, int lw,
                                                         double* __restrict__ y, u32* __restrict__ cmask,
                                                         u32* __restrict__ pmask, u32* __restrict__ tile_cnt,
                                                         unsigned long long* __restrict__ group_sum /* zeroed */) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char p1sm[];
  const P1Smem Lo = p1_smem_layout(lw);
  double* wd = (double*)(p1sm + Lo.wd);
  int* ext = (int*)(p1sm + Lo.ext);          // raw tile + halo (int32 counts)
  double* yout = (double*)(p1sm + Lo.yout);  // yout[1 + x] = y of tile sample x; [0], [cnt+1] = neighbours
  u32* nz = (u32*)(p1sm + Lo.nz);            // bit s: staged sample s is non-zero
  int* red = (int*)(p1sm + Lo.red);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const TileWork tw = tiles[blockIdx.x];
  const int f0 = tw.f0;
  const int n = tw.n;
  const int cnt = min(TILE_SAMPLES, n - tw.lo);
  const int* yr = y_raw + f0;
  for (int d = tid; d <= lw; d += GAUSS_THREADS) wd[d] = gw[lw - d];
  const int span = cnt + 2 * lw + 2;
  const int first = tw.lo - lw - 1;
  const int span_r = (span + GAUSS_THREADS - 1) / GAUSS_THREADS * GAUSS_THREADS;
  const bool interior = first >= 0 && first + span <= n;
  const int n2 = 2 * n;
  // Staging in two passes.  (1) every lane issues ALL its 4-byte global -> shared copies asynchronously (cp.async):
  // with a register in between, the store of chunk k stalled the warp until its load returned and the 5 - 9 loads of a
  // lane went out one latency after the other (29 % of the kernel's stall samples).  (2) after cp.async.wait_all each
  // lane reads back its OWN elements (no cross-lane visibility needed) and the warp ballots the non-zero mask.
  const unsigned ext_sh = (unsigned)__cvta_generic_to_shared(ext);
  for (int s0 = warp * 32; s0 < span_r; s0 += GAUSS_THREADS) {
    const int s = s0 + lane;
    const int j00 = first + s0;
    if (j00 >= 0 && j00 + 31 < n && s0 + 31 < span) {
      // warp-uniform fast path: the 32 samples lie inside the island (most tiles touch an island end, so a
      // CTA-wide `interior` test sent 94 % of them through the reflect arithmetic below)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ext_sh + 4u * (unsigned)s), "l"(yr + j00 + lane) : "memory");
    } else if (s < span) {
      const int j0 = j00 + lane;
      int j = j0;
      if (!interior) {
        // scipy's reflect (d c b a | a b c d | d c b a): one fold covers every island longer than the halo
        if (j0 < 0) j = -1 - j0;
        else if (j0 >= n) j = n2 - 1 - j0;
        if ((unsigned)j >= (unsigned)n) {  // island shorter than the halo: general period-2n fold
          j = j0 % n2;
          if (j < 0) j += n2;
          if (j >= n) j = n2 - 1 - j;
        }
      }
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(ext_sh + 4u * (unsigned)s), "l"(yr + j) : "memory");
    }
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  for (int s0 = warp * 32; s0 < span_r; s0 += GAUSS_THREADS) {  // whole warps: the ballot needs every lane
    const int s = s0 + lane;
    const int v = s < span ? ext[s] : 0;
    const u32 m = __ballot_sync(0xffffffffu, v != 0);
    if (lane == 0) nz[s0 >> 5] = m;
  }
  if (tid < 2) nz[(span_r >> 5) + tid] = 0u;
  __syncthreads();
  // bit w of wm: mask word w of the staged window has a non-zero sample (at most 50 words at sigma = 50)
  unsigned long long wm;
  {
    const int n_nz = (span_r >> 5) + 2;
    const u32 lo32 = __ballot_sync(0xffffffffu, lane < n_nz && nz[lane] != 0u);
    const u32 hi32 = __ballot_sync(0xffffffffu, 32 + lane < n_nz && nz[32 + lane] != 0u);
    wm = ((unsigned long long)hi32 << 32) | lo32;
  }
  // ---- Gaussian: 32 samples per step (coalesced 256-byte stores); the tile's steps are dealt ROUND ROBIN to the four
  // warps (warp w: steps w, w + 4, ...).  Live steps come in runs (splice sites cluster) and the median tile is half
  // empty: with 256 consecutive samples per warp one warp carried the run while the others waited at the barrier
  // (19 % of the stall samples); dealt round robin the barrier wait shrank and the kernel went 197 -> 178 us ----
  // bit it of live: step `it` of this warp has a non-zero input near its window (else its y is all 0).  Lane `it`
  // tests its step once (64-bit shifts), one ballot hands the eight answers to the warp.
  u32 live;
  {
    bool any = false;
    const int xb = (lane * 4 + warp) * 32;  // step `it` of warp w = 32-sample step it * 4 + w of the tile, see below
    if (lane < TILE_WORDS / 4 && xb < cnt) {
      // inputs of the step: staged samples [xb + 1, xb + 32 + 2*lw]; testing the whole mask words that
      // hold them is conservative (a false positive only runs the sparse filter over zeros: same bits)
      const int w_a = (xb + 1) >> 5, w_b = (xb + 32 + 2 * lw) >> 5;  // w_b - w_a <= 14
      any = ((wm >> w_a) & ((2ull << (w_b - w_a)) - 1ull)) != 0ull;
    }
    live = __ballot_sync(0xffffffffu, any);
  }
  for (int it = 0; it < TILE_WORDS / 4; ++it) {
    const int xb = (it * 4 + warp) * 32;
    if (xb >= cnt) break;
    const int x = xb + lane;
    double v = 0.0;
    if (((live >> it) & 1u) && x < cnt) v = gauss_sparse(ext, nz, wd, lw, lw + 1 + x);
    if (x < cnt) {
      y[f0 + tw.lo + x] = v;
      yout[1 + x] = v;
    }
  }
  // the neighbours of the tile's first and last sample: only read when the tile does not start / end its island
  // (an island end is a candidate by rule and never looks at its neighbours)
  if (tid >= GAUSS_THREADS - 2) {
    const bool left = tid == GAUSS_THREADS - 2;
    if (left ? tw.lo > 0 : tw.lo + cnt < n)
      yout[left ? 0 : cnt + 1] = gauss_sparse(ext, nz, wd, lw, left ? lw : lw + 1 + cnt);
  }
  __syncthreads();
  // ---- candidates and positives ----
  auto Y = [&](int X) -> double {  // island-local sample, 0 <= X < n
    const int x = X - tw.lo;
    return (x >= -1 && x <= cnt) ? yout[1 + x] : gauss_global(yr, n, X, wd, lw);
  };
  int nc = 0, np = 0;
  u32* cm_out = cmask + (size_t)blockIdx.x * TILE_WORDS;
  u32* pm_out = pmask + (size_t)blockIdx.x * TILE_WORDS;
  // dead steps first (the same warp smoothed these samples: a dead step is all zeros): no positives, and only an
  // island's first / last sample can be a candidate -- lane `it` writes the two words of step `it`, no ballots
  {
    const int wi = lane * 4 + warp;
    u32 cm = 0u;
    if (lane < TILE_WORDS / 4 && wi * 32 < cnt && !((live >> lane) & 1u)) {
      const int X0 = tw.lo + wi * 32, e = n - 1 - X0;
      cm = (X0 == 0 ? 1u : 0u) | ((unsigned)e < 32u ? 1u << e : 0u);
      cm_out[wi] = cm;
      pm_out[wi] = 0u;
    }
    nc = __popc(cm);  // summed over the lanes below
  }
  for (u32 lv = live; lv; lv &= lv - 1u) {
    const int it = __ffs(lv) - 1;
    const int wi = it * 4 + warp;
    const int x = wi * 32 + lane, X = tw.lo + x;
    bool is_c = false, is_p = false;
    if (x < cnt) {
      is_c = (X == 0 || X == n - 1);
      const double v = yout[1 + x];
      is_p = v > 0.0;
      if (!is_c && is_p) {
        const double l = yout[x], r = yout[x + 2];
        if (l < v && r < v) is_c = true;
        else if (!(l > v) && !(r > v)) {  // a neighbour equals v: walk the plateau
          int a = X, b = X;
          while (a - 1 >= 0 && Y(a - 1) == v) --a;
          while (b + 1 <= n - 1 && Y(b + 1) == v) ++b;
          if (a >= 1 && b <= n - 2 && Y(a - 1) < v && Y(b + 1) < v && X == ((a + b) >> 1)) is_c = true;
        }
      }
    }
    const u32 cm = __ballot_sync(0xffffffffu, is_c), pm = __ballot_sync(0xffffffffu, is_p);
    if (lane == 0) { cm_out[wi] = cm; pm_out[wi] = pm; nc += __popc(cm); np += __popc(pm); }
  }
  nc += np << 16;
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) nc += __shfl_xor_sync(0xffffffffu, nc, o);  // lanes 0..7 hold the dead steps' counts
  if (lane == 0) red[warp] = nc;
  __syncthreads();
  if (tid == 0) {
    const u32 v = (u32)(red[0] + red[1] + red[2] + red[3]);  // candidates | positives << 16
    tile_cnt[blockIdx.x] = v;
    // totals of every group of TILE_GROUP consecutive tiles: positives << 32 | candidates
    atomicAdd(&group_sum[blockIdx.x / TILE_GROUP], ((unsigned long long)(v >> 16) << 32) | (unsigned long long)(v & 0xffffu));
  }
}

// (candidates, positives) before every tile: one CTA per group of TILE_GROUP tiles adds the totals of the groups before
// it (k_smooth's atomics) and scans its own tiles' counts.  k_tile_lists then reads its two offsets with one load; when
// every tile summed up to TILE_GROUP counts by itself, that prefix was two thirds of its instructions.
__global__ void __launch_bounds__(TILE_GROUP) k_tile_prefix(int n_tiles, const u32* __restrict__ tile_cnt,
                                                            const unsigned long long* __restrict__ group_sum,
                                                            int2* __restrict__ tile_off) {
  pdl_prologue();
  __shared__ i64 sm[40];
  __shared__ i64 before_sm;
  const int g = blockIdx.x, tile = g * TILE_GROUP + threadIdx.x;
  i64 s = 0;
  for (int k = threadIdx.x; k < g; k += TILE_GROUP) s += (i64)group_sum[k];  // positives << 32 | candidates, no carries
  i64 tot;
  block_exclusive_scan<i64>(s, &tot, sm);
  if (threadIdx.x == 0) before_sm = tot;
  i64 v = 0;
  if (tile < n_tiles) {
    const u32 c = tile_cnt[tile];
    v = ((i64)(c >> 16) << 32) | (i64)(c & 0xffffu);
  }
  const i64 ex = block_exclusive_scan<i64>(v, (i64*)nullptr, sm);  // (its first barrier publishes before_sm)
  const i64 off = before_sm + ex;
  if (tile < n_tiles) tile_off[tile] = make_int2((int)(off & 0xffffffffll), (int)(off >> 32));
}

// Ordered candidate list and ordered positive samples of one tile from its ballot words.  The tile's
// offsets into the two lists come from k_tile_prefix.  Also: per-tint offsets of the positive list and the totals.
__global__ void __launch_bounds__(GAUSS_THREADS) k_tile_lists(const TileWork* __restrict__ tiles, int n_tiles,
                                                             const int* __restrict__ island_sample_off,
                                                             const int* __restrict__ island_tint,
                                                             const int* __restrict__ tint_island_off, int n_tints,
                                                             const u32* __restrict__ cmask, const u32* __restrict__ pmask,
                                                             const u32* __restrict__ tile_cnt,
                                                             const int2* __restrict__ tile_off,
                                                             const double* __restrict__ y, int* __restrict__ cand_flat,
                                                             double* __restrict__ vbuf, int* __restrict__ tint_pos_off,
                                                             i64* __restrict__ n_cand_out) {
  pdl_prologue();
  __shared__ int pre_c[TILE_WORDS], pre_p[TILE_WORDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x;
  const u32* cm_in = cmask + (size_t)tile * TILE_WORDS;
  const u32* pm_in = pmask + (size_t)tile * TILE_WORDS;
  const int2 off = tile_off[tile];  // (candidates, positives) before this tile (k_tile_prefix)
  const TileWork tw = tiles[tile];
  const int nwords = (min(TILE_SAMPLES, tw.n - tw.lo) + 31) >> 5;  // k_smooth wrote only these
  if (warp == 0) {
    int c = lane < nwords ? __popc(cm_in[lane]) : 0, p = lane < nwords ? __popc(pm_in[lane]) : 0;
    const int c0 = c, p0 = p;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int a = __shfl_up_sync(0xffffffffu, c, o), b = __shfl_up_sync(0xffffffffu, p, o);
      if (lane >= o) { c += a; p += b; }
    }
    pre_c[lane] = c - c0;
    pre_p[lane] = p - p0;
  }
  __syncthreads();
  const int off_c = off.x, off_p = off.y;
  if (tid == 0) {
    if (tw.lo == 0) {
      const int t = island_tint[tw.island];
      if (tw.island == tint_island_off[t]) tint_pos_off[t] = off_p;
    }
    if (tile == n_tiles - 1) {
      const u32 v = tile_cnt[tile];
      tint_pos_off[n_tints] = off_p + (int)(v >> 16);
      *n_cand_out = (i64)off_c + (i64)(v & 0xffffu);
    }
  }
  const int fbase = tw.f0 + tw.lo;
  const u32 lt = (1u << lane) - 1u;
  for (int it = 0; it < TILE_WORDS / 4; ++it) {
    const int wi = it * 4 + warp;  // round robin: a half-empty tile still gives every warp words
    if (wi >= nwords) break;
    const u32 cm = cm_in[wi], pm = pm_in[wi];
    const int f = fbase + wi * 32 + lane;
    if ((cm >> lane) & 1u) cand_flat[off_c + pre_c[wi] + __popc(cm & lt)] = f;
    if ((pm >> lane) & 1u) vbuf[off_p + pre_p[wi] + __popc(pm & lt)] = y[f];
  }
}

// after compaction: per candidate rank q -> island id, and the island / tint offset tables.  The number of
// candidates is read on the device (grid-stride: the launch does not depend on it).
__global__ void k_cand_meta(const int* __restrict__ cand_flat, const i64* __restrict__ n_cand_p,
                            const int* __restrict__ island_sample_off, int n_islands, int* __restrict__ cand_island,
                            int* __restrict__ island_cand_off) {
  pdl_prologue();
  const int n_cand = (int)*n_cand_p;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q <= n_cand; q += gridDim.x * blockDim.x) {
    if (q == n_cand) { island_cand_off[n_islands] = n_cand; break; }
    int f = cand_flat[q];
    int isl = upper_row(island_sample_off, n_islands, f);
    cand_island[q] = isl;
    if (f == island_sample_off[isl]) island_cand_off[isl] = q;
  }
}

// ---------------------------------------------------------------------------------------------
// K3 variance threshold (one CTA per tint) over the tint's slice of the compacted positive samples
// (written in sample order by k_tile_lists): numpy's pairwise summation tree (DOUBLE_pairwise_sum) for
// mean and variance:
//   n < 8: sequential; n <= 128: eight strided accumulators, combined ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//   remainder added sequentially; else split at n/2 rounded down to a multiple of 8.
// Leaves (<=128 elements) are summed in parallel by groups of 8 lanes, inner nodes level by level.
// ---------------------------------------------------------------------------------------------
#define THR_THREADS 256

// SQ: sum of (a[i]-mean)^2 instead of a[i] (numpy evaluates x = a - mean, then x*x, then the same tree)
template <bool SQ>
__device__ __forceinline__ double pw_term(double v, double mean) {
  if (!SQ) return v;
  double d = __dsub_rn(v, mean);
  return __dmul_rn(d, d);
}
// one leaf by a group of 8 lanes: lane k owns numpy's accumulator r[k] (coalesced 64-byte reads), the
// combine ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is three shuffle steps, the remainder is added by lane 0
template <bool SQ>
__device__ __forceinline__ double pw_leaf8(const double* __restrict__ a, int n, double mean, int sub) {
  // every lane of the warp reaches the shuffles, whatever its group's n
  const int n8 = n - (n % 8);
  double r = 0.0;
  if (n >= 8) {
    r = pw_term<SQ>(a[sub], mean);
    for (int i = 8; i < n8; i += 8) r = __dadd_rn(r, pw_term<SQ>(a[i + sub], mean));
  }
  r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 1));
  r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 2));
  r = __dadd_rn(r, __shfl_down_sync(0xffffffffu, r, 4));
  if (sub == 0) {
    if (n < 8) {
      r = 0.0;
      for (int i = 0; i < n; ++i) r = __dadd_rn(r, pw_term<SQ>(a[i], mean));
    } else {
      for (int i = n8; i < n; ++i) r = __dadd_rn(r, pw_term<SQ>(a[i], mean));
    }
  }
  return r;  // valid on sub == 0
}

// The pairwise tree as an implicit heap (root 1, children 2i / 2i+1): a node of length l splits at
// n2 = l/2 - (l/2)%8 while l > 128, so all leaves sit at the last two or three levels and the heap of a
// tint with n positives has at most ~n/14 slots.  Every slot finds its (offset, length) by walking
// down from the root (<= 24 steps); leaves are summed by groups of 8 lanes, inner nodes level by
// level from the bottom -- no serial pass over the leaves.
#define THR_HEAP_SMEM 2048  // heap slots held in shared memory (n <= ~28 k positives); else global scratch
__device__ __forceinline__ bool pw_node(int n, int idx, int& off, int& len) {
  // idx >= 1.  false if the slot is not a node of the tree (an ancestor is already a leaf)
  int o = 0, l = n;
  const int depth = 31 - __clz(idx);
  for (int b = depth - 1; b >= 0; --b) {
    if (l <= 128) return false;
    int n2 = l / 2;
    n2 -= n2 % 8;
    if ((idx >> b) & 1) { o += n2; l -= n2; } else { l = n2; }
  }
  off = o;
  len = l;
  return true;
}

template <bool SQ>
__device__ __forceinline__ double pw_tree(const double* __restrict__ v, int n, double mean, int levels, int* h_len,
                                          double* h_val) {
  // 1. leaves (groups of 8 lanes; whole warps run the shuffles)
  const int slots = 1 << levels;
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;
  for (int i0 = 0; i0 < slots; i0 += THR_THREADS / 8) {
    const int idx = i0 + grp;
    int off = 0, len = 0;
    const bool node = idx >= 1 && idx < slots && pw_node(n, idx, off, len);
    const bool leaf = node && len <= 128;
    const double r = pw_leaf8<SQ>(v + (leaf ? off : 0), leaf ? len : 0, mean, sub);
    if (sub == 0 && idx < slots) {
      h_len[idx] = node ? len : 0;
      if (leaf) h_val[idx] = r;
    }
  }
  __syncthreads();
  // 2. inner nodes, bottom level first
  for (int d = levels - 2; d >= 0; --d) {
    for (int idx = (1 << d) + threadIdx.x; idx < (2 << d); idx += THR_THREADS)
      if (h_len[idx] > 128) h_val[idx] = __dadd_rn(h_val[2 * idx], h_val[2 * idx + 1]);
    __syncthreads();
  }
  return h_val[1];
}

__global__ void __launch_bounds__(THR_THREADS) k_threshold(const int* __restrict__ tint_order,
                                                          const int* __restrict__ tint_island_off,
                                                          const int* __restrict__ island_sample_off,
                                                          const int* __restrict__ tint_pos_off, double vf,
                                                          const double* __restrict__ vbuf, int* __restrict__ heap_len,
                                                          double* __restrict__ heap_val, double* __restrict__ thr) {
  pdl_prologue();
  __shared__ int s_len[THR_HEAP_SMEM];
  __shared__ double s_val[THR_HEAP_SMEM];
  const int t = tint_order[blockIdx.x];  // largest tints first: the longest CTA must not start last
  const int s0 = island_sample_off[tint_island_off[t]];
  const int n = tint_pos_off[t + 1] - tint_pos_off[t];
  if (n == 0) {
    if (threadIdx.x == 0) thr[t] = __longlong_as_double(0x7ff8000000000000LL);  // NaN (:757-759, empty mean)
    return;
  }
  // depth of the tree: the right child (l - n2 >= l/2) is the longer one, up to 15 elements longer
  // than half; one spare level covers paths that alternate sides
  int levels = 1;
  for (int l = n; l > 128; ++levels) { int n2 = l / 2; n2 -= n2 % 8; l -= n2; }
  levels += 1;
  const bool small = (1 << levels) <= THR_HEAP_SMEM;
  // global heap scratch of tint t: slots [s0/8 + 64 t, ...), at most n/14 + 4 of them (see THR_HEAP_ELEMS)
  int* h_len = small ? s_len : heap_len + (s0 / 8 + 64 * t);
  double* h_val = small ? s_val : heap_val + (s0 / 8 + 64 * t);
  const double* v = vbuf + tint_pos_off[t];
  const double mean = __ddiv_rn(pw_tree<false>(v, n, 0.0, levels, h_len, h_val), (double)n);
  __syncthreads();
  const double var = __ddiv_rn(pw_tree<true>(v, n, mean, levels, h_len, h_val), (double)n);
  if (threadIdx.x == 0) thr[t] = __dadd_rn(mean, __dmul_rn(vf, __dsqrt_rn(var)));
}
