// kernels_finish.cuh -- refine, final positions, digit matrix, 1-runs, gaps / poly-A.
// Reference steps: refine_segmentation (freddie_segment.py:249-266), digits (:808-838),
// get_unaligned_gaps_and_polyA (:370-472) with get_interval_start/end (:307-349),
// forward_thread_cigar (:289-304), find_longest_poly (:352-367).
#pragma once
#include "common.cuh"

// Pre-refine finals = candidates kept by fixed | DP.  One thread per candidate: mark its sample in the
// per-sample flag array and, if the segment up to the next final of the island is longer than 40
// samples (:252), append it to the refine work list (arbitrary order).
__global__ void k_final_mark(const i64* __restrict__ n_cand_p, const u8* __restrict__ dpfinal, const int* __restrict__ cand_flat,
                             const int* __restrict__ cand_island, const int* __restrict__ island_cand_off,
                             u8* __restrict__ sflag, int2* __restrict__ ref_list, int* __restrict__ ref_cnt) {
  pdl_prologue();
  const int n_cand = (int)*n_cand_p;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < n_cand; q += gridDim.x * blockDim.x) {
    if (!dpfinal[q]) continue;
    const int a = cand_flat[q];
    sflag[a] = 1;
    const int c1 = island_cand_off[cand_island[q] + 1];
    if (q >= c1 - 1) continue;
    int e = q + 1;
    while (!dpfinal[e]) ++e;  // the island's last candidate is final
    const int b = cand_flat[e];
    if (b - a > 2 * 20) ref_list[atomicAdd(ref_cnt, 1)] = make_int2(a, b);
  }
}

// ---------------------------------------------------------------------------------------------
// K9 refine: one CTA per pre-refine segment (a, b) of an island (:249-266).
//   v = raw[a:b] with 20 samples zeroed at both ends; skip if b-a <= 40 or sum(v) < 20;
//   g = Gaussian(v), radius int(sigma+.5), zero outside (mode='constant'), same pair order as K2;
//   peaks = strict local maxima of g (plateau midpoint), min-distance 20 suppression by descending
//   height (ties: the later peak first = stable ascending argsort walked from the end);
//   keep peak i iff the sequential sum of g[round(i-sigma) : round(i+sigma+1)] (python slice) >= 20.
// g and the peak states live in global scratch indexed by flat sample (segments are disjoint).
// ---------------------------------------------------------------------------------------------
#define REF_THREADS 128
#define REF_SKIP 20

// refine, stage A: one warp per candidate segment sums the inner raw signal (:256-257: skip if the sum
// is below 20).  Splice sites are sparse, so almost every segment stops here; the survivors go to the
// list of stage B (one CTA per segment).
__global__ void k_refine_filter(const int2* __restrict__ in_list, const int* __restrict__ in_cnt,
                                const int* __restrict__ y_raw, int2* __restrict__ out_list, int* __restrict__ out_cnt) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  const int n = *in_cnt;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < n; e += nw) {
    const int2 seg = in_list[e];
    long long s = 0;
    for (int x = seg.x + REF_SKIP + lane; x < seg.y - REF_SKIP; x += 32) s += y_raw[x];
    int si = (int)min(s, (long long)20);  // non-negative terms: only "total < 20" matters
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) si += __shfl_xor_sync(0xffffffffu, si, o);
    if (lane == 0 && si >= 20) out_list[atomicAdd(out_cnt, 1)] = seg;
  }
}

__device__ void k_refine_segment(const int a, const int b, const int* __restrict__ y_raw,
                                 const double* __restrict__ rw, int rad, double sigma, double* __restrict__ gbuf,
                                 u8* __restrict__ pstate, u8* __restrict__ sflag, int* sm_red, int* sm_flag_p) {
  const int len = b - a;
#define sm_flag (*sm_flag_p)
  const int tid = threadIdx.x;
  // sum of the inner raw signal (integers)
  long long s = 0;
  for (int x = REF_SKIP + tid; x < len - REF_SKIP; x += REF_THREADS) s += y_raw[a + x];
  int si = (int)min(s, (long long)20);  // non-negative terms: only "total < 20" matters, so clamp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) si += __shfl_xor_sync(0xffffffffu, si, o);
  if ((tid & 31) == 0) sm_red[tid >> 5] = si;
  __syncthreads();
  long long tot = 0;
  for (int w = 0; w < REF_THREADS / 32; ++w) tot += sm_red[w];
  if (tot < 20) return;
  // g = constant-mode Gaussian of v
  double* g = gbuf + a;
  u8* ps = pstate + a;
  for (int x = tid; x < len; x += REF_THREADS) {
    auto v = [&](int i) -> double {
      return (i >= REF_SKIP && i < len - REF_SKIP) ? (double)y_raw[a + i] : 0.0;
    };
    double acc = __dmul_rn(v(x), rw[rad]);
    for (int jj = -rad; jj < 0; ++jj) acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(v(x + jj), v(x - jj)), rw[rad + jj]));
    g[x] = acc;
    ps[x] = 0;
  }
  __syncthreads();
  // peaks (1 = undecided)
  for (int x = 1 + tid; x < len - 1; x += REF_THREADS) {
    double vx = g[x];
    if (g[x - 1] < vx) {
      int ia = x + 1;
      while (ia < len - 1 && g[ia] == vx) ++ia;
      if (g[ia] < vx) ps[(x + ia - 1) >> 1] = 1;
    }
  }
  __syncthreads();
  // min-distance suppression in rounds: an undecided peak with no higher-priority undecided peak
  // within distance < 20 is kept (2); undecided peaks next to a kept one are removed (3)
  for (int round = 0; round < len; ++round) {
    if (tid == 0) sm_flag = 0;
    __syncthreads();
    for (int x = tid; x < len; x += REF_THREADS) {
      if (ps[x] != 1) continue;
      bool top = true;
      double hx = g[x];
      for (int d = 1; d < REF_SKIP && top; ++d) {
        int l = x - d, r = x + d;
        // state 4 = marked "keep" earlier in this same round by another thread: still a competitor
        if (l >= 0 && (ps[l] == 1 || ps[l] == 4) && g[l] > hx) top = false;    // earlier peak wins only if strictly higher
        if (r < len && (ps[r] == 1 || ps[r] == 4) && g[r] >= hx) top = false;  // later peak wins ties
      }
      if (top) ps[x] = 4;  // provisional keep
    }
    __syncthreads();
    for (int x = tid; x < len; x += REF_THREADS) {
      if (ps[x] != 1) continue;
      bool rm = false;
      for (int d = 1; d < REF_SKIP && !rm; ++d) {
        int l = x - d, r = x + d;
        if ((l >= 0 && ps[l] == 4) || (r < len && ps[r] == 4)) rm = true;
      }
      if (rm) ps[x] = 3;
    }
    __syncthreads();
    for (int x = tid; x < len; x += REF_THREADS) {
      if (ps[x] == 4) ps[x] = 2;
      else if (ps[x] == 1) sm_flag = 1;
    }
    __syncthreads();
    if (!sm_flag) break;
    __syncthreads();
  }
  // window test on kept peaks (python slice semantics, sequential left-to-right sum)
  for (int x = tid; x < len; x += REF_THREADS) {
    if (ps[x] != 2) continue;
    long long lo_i = (long long)rint(__dsub_rn((double)x, sigma));
    long long hi_i = (long long)rint(__dadd_rn(__dadd_rn((double)x, sigma), 1.0));
    if (lo_i < 0) { lo_i += len; if (lo_i < 0) lo_i = 0; }
    if (hi_i < 0) { hi_i += len; if (hi_i < 0) hi_i = 0; }
    if (lo_i > len) lo_i = len;
    if (hi_i > len) hi_i = len;
    double sum = 0.0;
    for (long long i = lo_i; i < hi_i; ++i) sum = __dadd_rn(sum, g[i]);
    if (!(sum < 20.0)) sflag[a + x] = 1;
  }
}
#undef sm_flag

__global__ void __launch_bounds__(REF_THREADS) k_refine(const int2* __restrict__ ref_list, const int* __restrict__ ref_cnt,
                                                       const int* __restrict__ y_raw, const double* __restrict__ rw,
                                                       int rad, double sigma, double* __restrict__ gbuf,
                                                       u8* __restrict__ pstate, u8* __restrict__ sflag) {
  pdl_prologue();
  __shared__ int sm_red[REF_THREADS / 32];
  __shared__ int sm_flag;
  const int n_work = *ref_cnt;
  for (int e = blockIdx.x; e < n_work; e += gridDim.x) {
    k_refine_segment(ref_list[e].x, ref_list[e].y, y_raw, rw, rad, sigma, gbuf, pstate, sflag, sm_red, &sm_flag);
    __syncthreads();
  }
}

// after compaction of the per-sample final flags: positions + per-tint offsets + island of each final
__global__ void k_final_meta(const i64* __restrict__ n_final_p, const int* __restrict__ final_flat,
                             const int* __restrict__ island_sample_off, const int* __restrict__ island_start,
                             const int* __restrict__ island_tint, const int* __restrict__ tint_island_off, int n_islands,
                             int n_tints, int* __restrict__ final_pos, int* __restrict__ final_island,
                             int* __restrict__ tint_final_off) {
  pdl_prologue();
  const int n_final = (int)*n_final_p;  // device-side count: no host round trip before this launch
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e <= n_final; e += gridDim.x * blockDim.x) {
    if (e == n_final) { tint_final_off[n_tints] = n_final; break; }
    int f = final_flat[e];
    int isl = upper_row(island_sample_off, n_islands, f);
    final_island[e] = isl;
    final_pos[e] = island_start[isl] + (f - island_sample_off[isl]);
    int t = island_tint[isl];
    if (f == island_sample_off[isl] && isl == tint_island_off[t]) tint_final_off[t] = e;
  }
}

// per tint: digit block size = n_reps * (n_final - 1)
__global__ void k_digit_sizes(int n_tints, const int* __restrict__ tint_rep_off, const int* __restrict__ tint_final_off,
                              i64* __restrict__ sz) {
  pdl_prologue();
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tints) return;
  i64 S = tint_final_off[t + 1] - tint_final_off[t] - 1;
  sz[t] = S * (tint_rep_off[t + 1] - tint_rep_off[t]);
}

// per final e (segment e -> e+1): integer cuts of the segment, or a separator marker
__global__ void k_seg_cuts(const i64* __restrict__ n_final_p, const int* __restrict__ final_flat,
                           const int* __restrict__ final_island, const int2* __restrict__ cut_tab,
                           const double* __restrict__ tbl, int tbl_len, double tp,
                           int* __restrict__ seg_ty, int* __restrict__ seg_tn) {
  pdl_prologue();
  const int n_final = (int)*n_final_p;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_final; e += gridDim.x * blockDim.x) {
    int ty = 0x7fffffff, tn = -2;  // tn == -2 marks "no segment" (island separator or tint end)
    if (e + 1 < n_final && final_island[e] == final_island[e + 1])
      length_cuts_t(final_flat[e + 1] - final_flat[e] + 1, cut_tab, tbl, tbl_len, tp, ty, tn);
    seg_ty[e] = ty;
    seg_tn[e] = tn;
  }
}

// Capacity guards of the output stages (see Caps in common.cuh): a stage whose data-dependent buffer is
// too small is skipped -- and so is everything that would read its output -- and the host repeats the run.
__device__ __forceinline__ bool digits_fit(const i64* __restrict__ cnt, const Caps& cp) { return cnt[CNT_NDIG] <= cp.dig; }
__device__ __forceinline__ bool runs_fit(const i64* __restrict__ cnt, const Caps& cp) {
  return digits_fit(cnt, cp) && cnt[CNT_NRUN] <= cp.runs;
}
__device__ __forceinline__ bool gaps_fit(const i64* __restrict__ cnt, const Caps& cp) {
  return runs_fit(cnt, cp) && cnt[CNT_NGAP] <= cp.gaps;
}
__device__ __forceinline__ bool clips_fit(const i64* __restrict__ cnt, const Caps& cp) {
  return gaps_fit(cnt, cp) && cnt[CNT_CLIPW] <= cp.clipw;
}

// ---------------------------------------------------------------------------------------------
// K10 digits (:808-838).  One thread per read rep walks the tint's final positions and its own
// (ordered) intervals with two pointers: P(x) = samples of the rep strictly before flat x, and the
// coverage of segment s is P(f_{s+1}) - P(f_s).  32 segments at a time go through a shared-memory
// transpose so that the ASCII digit rows (rep-major, what the formatter copies) are written as
// contiguous 32-byte pieces.  The same walk counts the rep's runs of '1' digits.
// ---------------------------------------------------------------------------------------------
#define DIG_THREADS 128
#define DIG_SEGS 32
#define DIG_CHUNKS 8      // grid.y: the segments of a tint are cut into chunks (bounded serial walk per thread)
#define DIG_MIN_CHUNK 64  // segments, a multiple of DIG_SEGS
__global__ void __launch_bounds__(DIG_THREADS) k_digits(const RepTile* __restrict__ tiles,
                                                       const int* __restrict__ tint_rep_off,
                                                       const int* __restrict__ tint_final_off,
                                                       const i64* __restrict__ tint_digit_off,
                                                       const int* __restrict__ rep_iv_off,
                                                       const int* __restrict__ iv_fs, const int* __restrict__ iv_fe,
                                                       const int* __restrict__ final_flat,
                                                       const int* __restrict__ seg_ty, const int* __restrict__ seg_tn,
                                                       u8* __restrict__ digits, int* __restrict__ run_cnt /* zeroed */,
                                                       int* __restrict__ err, const i64* __restrict__ cnt, Caps caps) {
  pdl_prologue();
  __shared__ u8 tile[DIG_THREADS][DIG_SEGS + 1];
  if (!digits_fit(cnt, caps)) return;
  const RepTile tl = tiles[blockIdx.x];
  const int r0 = tint_rep_off[tl.tint];
  const int R = tint_rep_off[tl.tint + 1] - r0;
  const int f0 = tint_final_off[tl.tint];
  const int S = tint_final_off[tl.tint + 1] - f0 - 1;
  const int chunk = max(DIG_MIN_CHUNK, ((S + DIG_CHUNKS - 1) / DIG_CHUNKS + DIG_SEGS - 1) / DIG_SEGS * DIG_SEGS);
  const int sa = (int)blockIdx.y * chunk, sb = min(S, sa + chunk);
  if (sa >= S) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = tl.rep_lo + threadIdx.x;
  const bool have = r < R;
  int a = 0, b = 0;
  if (have) { a = rep_iv_off[r0 + r]; b = rep_iv_off[r0 + r + 1]; }
  int fs = (a < b) ? iv_fs[a] : 0x7fffffff;
  int fe = (a < b) ? iv_fe[a] : 0x7fffffff;
  u32 acc = 0;
  auto P = [&](int x) -> u32 {  // x is non-decreasing over the calls
    while (a < b && fe < x) {
      acc += (u32)(fe - fs + 1);
      ++a;
      fs = (a < b) ? iv_fs[a] : 0x7fffffff;
      fe = (a < b) ? iv_fe[a] : 0x7fffffff;
    }
    return acc + ((a < b && x > fs) ? (u32)(x - fs) : 0u);
  };
  auto digit = [&](int s, u32 p_lo, u32 p_hi) -> u8 {
    const int tn = seg_tn[f0 + s];
    if (tn == -2) return '0';  // island separator
    const int cov = (int)(p_hi - p_lo);
    if (cov > final_flat[f0 + s + 1] - final_flat[f0 + s] + 1) dev_fail(err, DEVERR_RATIO_RANGE, r0 + r);
    return (cov >= seg_ty[f0 + s]) ? '1' : ((cov <= tn) ? '0' : '2');
  };
  // the digit before the chunk decides whether the chunk's first '1' starts a run
  bool prev1 = false;
  u32 p_prev = 0u;
  if (have) {
    if (sa > 0) {
      const u32 p0 = P(final_flat[f0 + sa - 1]);
      p_prev = P(final_flat[f0 + sa]);
      prev1 = digit(sa - 1, p0, p_prev) == '1';
    } else {
      p_prev = P(final_flat[f0]);
    }
  }
  int runs = 0;
  u8* out0 = digits + tint_digit_off[tl.tint];
  for (int s0 = sa; s0 < sb; s0 += DIG_SEGS) {
    const int ns = min(DIG_SEGS, sb - s0);
    if (have) {
      for (int k = 0; k < ns; ++k) {
        const int s = s0 + k;
        const u32 p_next = P(final_flat[f0 + s + 1]);
        const u8 d = digit(s, p_prev, p_next);
        p_prev = p_next;
        const bool is1 = d == '1';
        runs += (is1 && !prev1) ? 1 : 0;
        prev1 = is1;
        tile[threadIdx.x][k] = d;
      }
    }
    __syncthreads();
    // rows of this warp's 32 reps, 32 contiguous bytes each
    for (int rr = 0; rr < 32; ++rr) {
      const int row = tl.rep_lo + warp * 32 + rr;
      if (row >= R) break;
      if (lane < ns) out0[(i64)row * S + s0 + lane] = tile[warp * 32 + rr][lane];
    }
    __syncthreads();
  }
  if (have && runs) atomicAdd(&run_cnt[r0 + r], runs);
}

// 1-runs per rep (shared by all reads of the rep): one warp per rep, 32 segments per step, starts and
// ends found with ballots -- the k-th start and the k-th end of a row belong to the same run.
__global__ void k_run_fill(int n_reps, const int* __restrict__ rep_tint, const int* __restrict__ tint_rep_off,
                           const int* __restrict__ tint_final_off, const i64* __restrict__ tint_digit_off,
                           const u8* __restrict__ digits, const int* __restrict__ run_off, int2* __restrict__ runs,
                           const i64* __restrict__ cnt, Caps caps) {
  pdl_prologue();
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= n_reps || !runs_fit(cnt, caps)) return;
  int t = rep_tint[r];
  int S = tint_final_off[t + 1] - tint_final_off[t] - 1;
  const u8* row = digits + tint_digit_off[t] + (i64)(r - tint_rep_off[t]) * S;
  int* out = (int*)(runs + run_off[r]);
  if (run_off[r + 1] == run_off[r]) return;
  int base_s = 0, base_e = 0;
  bool carry1 = false;  // digit before this chunk is '1'
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool is1 = s < S && row[s] == '1';
    const unsigned m1 = __ballot_sync(0xffffffffu, is1);
    const bool after1 = (s0 + 32 < S) && row[s0 + 32] == '1';  // same address for the whole warp
    const unsigned prevm = (m1 << 1) | (carry1 ? 1u : 0u);
    const unsigned nextm = (m1 >> 1) | (after1 ? 0x80000000u : 0u);
    const unsigned ms = m1 & ~prevm, me = m1 & ~nextm;
    const unsigned lt = (1u << lane) - 1u;
    if ((ms >> lane) & 1u) out[2 * (base_s + __popc(ms & lt))] = s;
    if ((me >> lane) & 1u) out[2 * (base_e + __popc(me & lt)) + 1] = s;
    base_s += __popc(ms);
    base_e += __popc(me);
    carry1 = (m1 >> 31) & 1u;
  }
}

__global__ void k_gap_count(int n_reads, const int* __restrict__ read_rep, const int* __restrict__ run_off,
                            int* __restrict__ gap_cnt) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_reads) return;
  int r = read_rep[i];
  int c = run_off[r + 1] - run_off[r];
  gap_cnt[i] = c > 0 ? c - 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// find_longest_poly (:352-367) as independent scan TASKS.  A read has up to four: (start | end clip)
// x (A | T).  Scan position t of a clip of n bases maps to read index first + t*step in the bit-plane
// of the (possibly complemented) target base.  Score recurrence sc_t = max(0, sc_{t-1} + (+1 | -2)),
// sc_{-1} = 0 (the reference's special case for t = 0 gives the same value).  For each maximal run
// of positive scores: i* = LAST index of the run's maximum, len = i*+1-i0; inside a run the score is
// matches - 2*mismatches, so matches = (peak + 2*len)/3 exactly and purity p = matches/len.  A run
// counts if len >= 20 and p >= 0.85; the FIRST maximum of p wins.  Clips shorter than 20 bases cannot
// hold such a run and get no task.
// The scans are serial recurrences of very different lengths (0 .. whole read), so a thread per read
// leaves ~5 of 32 lanes busy.  Instead every task is a thread and the tasks are bucketed by length
// class (4 classes per octave, longest first), which keeps the lanes of a warp in step.
// ---------------------------------------------------------------------------------------------
#define POLY_CLASSES 96
#define POLY_LONG_CLASS 36  // poly_class(512): longer clips are scanned by a whole warp (k_poly_long); swept on B200
#define POLY_C 64            // bases per lane and window in k_poly_long
struct PolyRes { double p; int i0; int len; };  // len == 0: no qualifying run

__device__ __forceinline__ int poly_class_dev(int n) {
  // n >= 20.  class grows with n: octave * 4 + the next two mantissa bits
  int o = 31 - __clz(n);
  return min(POLY_CLASSES - 1, o * 4 + ((n >> (o - 2)) & 3));
}

struct GapArgs {
  int n_reads;
  const int* read_rep; const u8* read_strand; const int* read_len; const int* read_iv_off;
  const i64* read_seq_off; const int* read_tint;
  const int* riv_ts; const int* riv_te; const int* riv_qs; const int* riv_qe; const int* riv_cig_off;
  const u32* cigar; const u32* seq_a; const u32* seq_t;
  const int* run_off; const int2* runs;
  const int* tint_final_off; const int* final_pos;
  const int* read_gap_off;
  int* read_head; int* gap_rec; int* err;
  // poly tasks: slot = read*4 + (0 start-A, 1 start-T, 2 end-A, 3 end-T)
  int* clip_n;        // [2N] bases of the read's start (2i) / end (2i+1) clip, 0 = nothing to scan (< 20)
  int* clip_words;    // [2N] plane words the clip needs (lazy mode: scanned into clip_off)
  i64* clip_off;      // [2N+1] word offset of the clip's plane words inside seq_a / seq_t (see clip_geometry)
  int seq_resident;   // 1: seq_a/seq_t hold the whole reads (clip_off written by k_gap_prep);
                      // 0: they hold only the clip words, gathered by the host after k_gap_prep
  int* cls_count;     // [POLY_CLASSES] (+ [POLY_CLASSES] cursors, + 1 total) zeroed before k_poly_filter
  int* task_order;    // [4N] slots, longest class first
  PolyRes* task_res;  // [4N]
  int long_class;     // tasks of classes >= long_class go to k_poly_long (default POLY_LONG_CLASS)
  const i64* cnt;     // device counters + capacities of the run (guards)
  Caps caps;
  // edge store (lazy sequence mode, optional): first / last edge_words plane words of every read, resident
  const u32* edge;    // [n_reads][2 sides][2 planes][edge_words] or NULL
  int edge_words;
  int* clip_eoff;     // [2N] >= 0: the clip's words start at edge + clip_eoff (plane A; plane T at + edge_words);
                      //       -1: they are in seq_a / seq_t at clip_off
};

// first plane word of a clip: the edge store when the clip fits one of its blocks, else the (gathered or
// resident) plane arrays
__device__ __forceinline__ const u32* clip_plane(const GapArgs& A, int clip, bool plane_a) {
  if (A.edge) {
    const int eo = A.clip_eoff[clip];
    if (eo >= 0) return A.edge + eo + (plane_a ? 0 : A.edge_words);
  }
  return (plane_a ? A.seq_a : A.seq_t) + A.clip_off[clip];
}

// forward_thread_cigar (:289-304): every op length, insertions included, is clipped by the remaining
// target distance.  Returns false if the CIGAR is exhausted before reaching t_goal.
__device__ __forceinline__ bool thread_cigar(const u32* __restrict__ cig, int c0, int c1, int t_goal, int t_pos,
                                             int q_pos, int& q_out) {
  int k = c0;
  while (t_pos < t_goal) {
    if (k >= c1) return false;
    u32 op = cig[k++];
    int c = (int)(op >> 4);
    int ty = (int)(op & 15u);
    c = min(c, t_goal - t_pos);
    if (ty == 0) { t_pos += c; q_pos += c; }
    else if (ty == 2) t_pos += c;
    else if (ty == 1) q_pos += c;
  }
  q_out = q_pos;
  return t_pos == t_goal;
}

// get_interval_start (:307-326): first interval with t_end >= p.  Intervals are ordered and disjoint
// (asserted at parse time, :158-161), so the linear search of the reference is a binary search.
__device__ bool interval_start(const GapArgs& A, int i0, int i1, int p, int& q, int& slack) {
  int lo = i0, hi = i1;  // first k in [i0, i1) with te[k] >= p
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (A.riv_te[mid] < p) lo = mid + 1; else hi = mid;
  }
  if (lo >= i1) return false;
  const int k = lo, ts = A.riv_ts[k];
  if (p < ts) { q = A.riv_qs[k]; slack = p - ts; return true; }
  slack = 0;
  if (!thread_cigar(A.cigar, A.riv_cig_off[k], A.riv_cig_off[k + 1], p, ts, A.riv_qs[k], q)) return false;
  return q >= A.riv_qs[k] && q <= A.riv_qe[k];
}
// get_interval_end (:329-349): last interval with t_start <= p
__device__ bool interval_end(const GapArgs& A, int i0, int i1, int p, int& q, int& slack) {
  int lo = i0, hi = i1;  // first k with ts[k] > p
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (A.riv_ts[mid] <= p) lo = mid + 1; else hi = mid;
  }
  if (lo <= i0) return false;
  const int k = lo - 1, ts = A.riv_ts[k], te = A.riv_te[k];
  if (te < p) { q = A.riv_qe[k]; slack = te - p; return true; }
  slack = 0;
  if (!thread_cigar(A.cigar, A.riv_cig_off[k], A.riv_cig_off[k + 1], p, ts, A.riv_qs[k], q)) return false;
  return q >= 0 && q <= A.riv_qe[k];
}

// A clip of n bases of a read of L bases covers the read's first n bases (start clip on '+', end clip
// on '-') or its last n (start clip on '-', end clip on '+'), because '-' reads are scanned from the
// other end (:393-401, :423-431).  Only the 32-bit plane words that overlap that range are needed:
// first word w_first, n_words words; scan position t reads bit (idx0 + t*step) RELATIVE to w_first.
struct ClipGeo { int w_first, n_words, idx0, step; };
__host__ __device__ __forceinline__ ClipGeo clip_geometry(int L, int n, bool is_start, bool minus) {
  ClipGeo g;
  const bool at_begin = (is_start != minus);
  g.step = minus ? -1 : 1;
  g.w_first = at_begin ? 0 : ((L - n) >> 5);
  g.n_words = at_begin ? ((n + 31) >> 5) : (((L + 31) >> 5) - g.w_first);
  const int abs0 = is_start ? (minus ? L - 1 : 0) : (minus ? n - 1 : L - n);
  g.idx0 = abs0 - (g.w_first << 5);
  return g;
}

// K11a: per read, everything of get_unaligned_gaps_and_polyA (:370-472) except the poly scans:
// clip bounds by CIGAR threading, unaligned gaps between consecutive 1-runs, and the scan tasks.
// head[3] = q_ssc and head[6] = q_esc are provisional; k_gap_finish rewrites them.
__global__ void __launch_bounds__(128) k_gap_prep(GapArgs A) {
  pdl_prologue();
  // two lanes per read: lane 0 threads the CIGAR to the start of the first 1-run, lane 1 to the end of the last one
  // (two independent chains of dependent loads), each then owns its side's clip and half of the gap records
  const int gt = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = gt >> 1, side = gt & 1;
  const bool act = i < A.n_reads && gaps_fit(A.cnt, A.caps);
  int q = 0, L = 0, ra = 0, rb = 0;
  bool ok = true, has = false;
  if (act) {
    int* head = A.read_head + (i64)i * 8 + side * 4;
    head[0] = head[1] = head[2] = head[3] = 0;
    A.clip_n[(i64)i * 2 + side] = 0;
    A.clip_words[(i64)i * 2 + side] = 0;
    const int rep = A.read_rep[i];
    ra = A.run_off[rep];
    rb = A.run_off[rep + 1];
    has = ra != rb;  // else: no '1' digit, empty gaps (:372)
    if (has) {
      const int t = A.read_tint[i];
      const int* fpos = A.final_pos + A.tint_final_off[t];  // segs[s] = (fpos[s], fpos[s+1])
      const int i0 = A.read_iv_off[i], i1 = A.read_iv_off[i + 1];
      L = A.read_len[i];
      int slack;
      ok = side == 0 ? interval_start(A, i0, i1, fpos[A.runs[ra].x], q, slack)
                     : interval_end(A, i0, i1, fpos[A.runs[rb - 1].y + 1], q, slack);
    }
  }
  const int q_other = __shfl_xor_sync(0xffffffffu, q, 1);
  const bool ok_other = __shfl_xor_sync(0xffffffffu, (int)ok, 1) != 0;
  if (!act || !has) return;
  const int q_ssc = side == 0 ? q : q_other, q_esc = side == 0 ? q_other : q;
  if (!(ok && ok_other)) { if (side == 0) dev_fail(A.err, DEVERR_THREAD_CIGAR, i); return; }
  if (!(0 <= q_ssc && q_ssc <= q_esc && q_esc <= L)) { if (side == 0) dev_fail(A.err, DEVERR_Q_RANGE, i); return; }
  int* head = A.read_head + (i64)i * 8;
  if (side == 0) { head[0] = 1; head[3] = q_ssc; }
  else head[6] = q_esc;
  {
    const int nb = side == 0 ? q_ssc : L - q_esc;  // 0: start clip, 1: end clip
    if (nb >= 20) {
      const bool minus = A.read_strand[i] != 0;
      const int nwr = (L + 31) >> 5;  // plane words of the read
      A.clip_n[(i64)i * 2 + side] = nb;
      const ClipGeo g = clip_geometry(L, nb, side == 0, minus);
      if (A.edge && g.n_words <= A.edge_words) {
        // the clip lies inside the read's first or last edge_words plane words: read it from the edge store
        const bool at_begin = (side == 0) != minus;
        const int idx = at_begin ? 0 : g.w_first - max(0, nwr - A.edge_words);
        A.clip_eoff[2 * (i64)i + side] = (int)((((i64)i * 2 + (at_begin ? 0 : 1)) * 2) * A.edge_words + idx);
      } else {
        if (A.edge) A.clip_eoff[2 * (i64)i + side] = -1;
        A.clip_words[(i64)i * 2 + side] = g.n_words;
        if (A.seq_resident) A.clip_off[2 * (i64)i + side] = A.read_seq_off[i] + g.w_first;
      }
    }
  }
  // unaligned gaps between consecutive 1-runs (:455-471): (l1, f2, owner); k_gap_sizes fills the size
  int* rec = A.gap_rec + ((i64)A.read_gap_off[i] + side) * 3;
  for (int k = ra + side; k + 1 < rb; k += 2, rec += 6) { rec[0] = A.runs[k].y; rec[1] = A.runs[k + 1].x; rec[2] = i; }
}

// K11a': one thread per unaligned-gap record: "{l1}-{f2}:{size}" (:455-471)
__global__ void k_gap_sizes(GapArgs A) {
  pdl_prologue();
  if (!gaps_fit(A.cnt, A.caps)) return;
  const int n_gaps = (int)A.cnt[CNT_NGAP];  // device-side count: the launch is a grid-stride loop
  for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < n_gaps; g += gridDim.x * blockDim.x) {
    int* rec = A.gap_rec + (i64)g * 3;
    const int l1 = rec[0], f2 = rec[1], i = rec[2];
    if (i < 0 || i >= A.n_reads) continue;  // record of a read that failed an assert in k_gap_prep (never written)
    const int* fpos = A.final_pos + A.tint_final_off[A.read_tint[i]];
    const int i0 = A.read_iv_off[i], i1 = A.read_iv_off[i + 1];
    const int L = A.read_len[i];
    int qa, sa, qb, sb;
    if (!interval_end(A, i0, i1, fpos[l1 + 1], qa, sa)) { dev_fail(A.err, DEVERR_THREAD_CIGAR, i); continue; }
    if (!interval_start(A, i0, i1, fpos[f2], qb, sb)) { dev_fail(A.err, DEVERR_THREAD_CIGAR, i); continue; }
    if (!(0 < qa && qa <= qb && qb < L)) { dev_fail(A.err, DEVERR_GAP_RANGE, i); continue; }
    int size = max(0, qb - qa + sa + sb);
    if (!(size < L)) { dev_fail(A.err, DEVERR_GAP_RANGE, i); continue; }
    rec[2] = size;
  }
}

// Lazy sequence mode: the poly-A/T scans only read the soft clips, a few per cent of the reads' bases, and
// the clips are only known after segmentation.  The caller's bit-planes stay in (pinned, device-mapped)
// HOST memory; one thread per needed plane word fetches it over the bus into the compact device arrays the
// scan kernels read (clip k owns words [clip_off[k], clip_off[k+1])).  Neighbouring threads read
// neighbouring words (the end clip of a read and the start clip of the next one are adjacent in the
// planes), so the requests coalesce into whole sectors.  No host round trip, no host gather.
__global__ void __launch_bounds__(256) k_clip_gather(GapArgs A, const u32* __restrict__ host_a, const u32* __restrict__ host_t,
                                                     u32* __restrict__ out_a, u32* __restrict__ out_t) {
  pdl_prologue();
  if (!clips_fit(A.cnt, A.caps)) return;
  const i64 total = A.cnt[CNT_CLIPW];
  const int n_clip = A.n_reads * 2;
  for (i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (i64)gridDim.x * blockDim.x) {
    const int k = upper_row64(A.clip_off, n_clip, j);  // clip that owns compact word j
    const int i = k >> 1;
    const ClipGeo g = clip_geometry(A.read_len[i], A.clip_n[k], (k & 1) == 0, A.read_strand[i] != 0);
    const i64 src = A.read_seq_off[i] + g.w_first + (j - A.clip_off[k]);
    out_a[j] = host_a[src];
    out_t[j] = host_t[src];
  }
}

// Scan-task filter, one thread per slot.  A run that find_longest_poly keeps has len >= 20 and at most
// 0.15*len mismatches, i.e. at most 0.15*len + 1 stretches of matches holding >= 0.85*len matches, so
// one stretch has >= ceil(0.85*len / (0.15*len + 1)) >= 5 consecutive matches.  A clip whose plane has
// no 5 consecutive set bits therefore yields "no run" without being scanned (exact; ~90 % of the
// clips of random sequence).  The test is word-parallel: v & v>>1 & v>>2 & v>>3 & v>>4 over the
// plane words, carried across word boundaries.  Survivors are counted per length class.
__global__ void __launch_bounds__(128) k_poly_filter(GapArgs A, u8* __restrict__ pass_flag) {
  pdl_prologue();
  __shared__ int sh_cnt[POLY_CLASSES];
  for (int k = threadIdx.x; k < POLY_CLASSES; k += blockDim.x) sh_cnt[k] = 0;
  __syncthreads();
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot < A.n_reads * 4 && clips_fit(A.cnt, A.caps)) {
    const int i = slot >> 2, which = slot & 3, clip = slot >> 1;
    const int n = A.clip_n[clip];
    bool pass = false;
    if (n >= 20) {
      const bool minus = A.read_strand[i] != 0;
      const bool want_a = (which & 1) == 0;
      const u32* pl = clip_plane(A, clip, want_a != minus);
      const ClipGeo geo = clip_geometry(A.read_len[i], n, which < 2, minus);
      const int last = geo.idx0 + geo.step * (n - 1);
      const int b_lo = min(geo.idx0, last), b_hi = max(geo.idx0, last);  // clip bits, relative to the first word
      const int w_lo = b_lo >> 5, w_hi = b_hi >> 5;
      u32 prev = 0u;
      for (int w = w_lo; w <= w_hi + 1 && !pass; ++w) {
        u32 cur = 0u;
        if (w <= w_hi) {
          cur = pl[w];
          if (w == w_lo) cur &= 0xffffffffu << (b_lo & 31);
          if (w == w_hi && (b_hi & 31) != 31) cur &= 0xffffffffu >> (31 - (b_hi & 31));
        }
        const unsigned long long v = ((unsigned long long)cur << 32) | prev;
        const unsigned long long r = v & (v >> 1) & (v >> 2) & (v >> 3) & (v >> 4);
        pass = (u32)r != 0u;  // stretches that start in `prev` (and may end in `cur`)
        prev = cur;
      }
      if (pass) atomicAdd(&sh_cnt[poly_class_dev(n)], 1);
      else { PolyRes none; none.p = 0.0; none.i0 = 0; none.len = 0; A.task_res[slot] = none; }
    }
    pass_flag[slot] = pass ? 1 : 0;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < POLY_CLASSES; k += blockDim.x)
    if (sh_cnt[k]) atomicAdd(&A.cls_count[k], sh_cnt[k]);
}

// one warp: class bases, LONGEST class first; cls_count[c] becomes the base, cursors start at 0.  Three classes per
// lane and a warp scan (one thread walking the 96 counters was a chain of dependent loads).
__global__ void k_poly_bases(int* __restrict__ cls_count, int long_class, int* __restrict__ stat /* [2] */) {
  pdl_prologue();
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  constexpr int PER = (POLY_CLASSES + 31) / 32;
  int v[PER], sum = 0;
#pragma unroll
  for (int j = 0; j < PER; ++j) {  // list position e = lane * PER + j holds class POLY_CLASSES - 1 - e
    const int e = lane * PER + j;
    v[j] = e < POLY_CLASSES ? cls_count[POLY_CLASSES - 1 - e] : 0;
    sum += v[j];
  }
  int x = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  const int total = __shfl_sync(0xffffffffu, x, 31);
  int acc = x - sum;
  __syncwarp();  // every count is read before any base is written
#pragma unroll
  for (int j = 0; j < PER; ++j) {
    const int e = lane * PER + j;
    if (e < POLY_CLASSES) {
      const int c = POLY_CLASSES - 1 - e;
      cls_count[c] = acc;
      if (c == long_class - 1) stat[1] = acc;  // tasks that are warp-scanned (long clips) come before this class
      acc += v[j];
    }
  }
  if (lane == 0) {
    cls_count[2 * POLY_CLASSES] = total;  // total tasks
    stat[0] = total;                       // scan tasks that survived the filter
  }
}

// order[] = the surviving tasks, longest class first.  The cursor of a class is bumped once per CTA and class
// (shared-memory counters first): per-warp bumps were ~50 k atomics on a dozen addresses, serialised in L2.
__global__ void __launch_bounds__(256) k_poly_scatter(int n_slots, const int* __restrict__ clip_n, const u8* __restrict__ pass_flag,
                               int* __restrict__ cls_count, int* __restrict__ order, const i64* __restrict__ cnt, Caps caps) {
  pdl_prologue();
  __shared__ int sh_cnt[POLY_CLASSES], sh_base[POLY_CLASSES];
  for (int k = threadIdx.x; k < POLY_CLASSES; k += blockDim.x) sh_cnt[k] = 0;
  __syncthreads();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const bool pass = s < n_slots && clips_fit(cnt, caps) && pass_flag[s];
  int c = 0, rank = 0;
  if (pass) {
    c = poly_class_dev(clip_n[s >> 1]);
    rank = atomicAdd(&sh_cnt[c], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < POLY_CLASSES; k += blockDim.x)
    if (sh_cnt[k]) sh_base[k] = atomicAdd(&cls_count[POLY_CLASSES + k], sh_cnt[k]);
  __syncthreads();
  if (pass) order[cls_count[c] + sh_base[c] + rank] = s;
}

// K11b: one thread per scan task (clips shorter than 1024 bases)
__global__ void __launch_bounds__(128) k_poly_scan(GapArgs A) {
  pdl_prologue();
  // tasks of the long classes (front of the order) belong to k_poly_long
  const int e = A.cls_count[A.long_class - 1] + blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= A.cls_count[2 * POLY_CLASSES] || !clips_fit(A.cnt, A.caps)) return;
  const int slot = A.task_order[e];
  const int i = slot >> 2, which = slot & 3;
  const int clip = slot >> 1;
  const int n = A.clip_n[clip];
  const bool minus = A.read_strand[i] != 0;
  const bool want_a = (which & 1) == 0;
  // '+': seq[..] == ch; '-': reversed read, complemented target base (:392-401, :422-431)
  const u32* pl = clip_plane(A, clip, want_a != minus);
  const ClipGeo geo = clip_geometry(A.read_len[i], n, which < 2, minus);
  const int step = geo.step;
  int idx = geo.idx0;
  PolyRes best; best.p = 0.0; best.i0 = 0; best.len = 0;
  int sc = 0, run_i0 = 0, run_best = 0, run_best_i = 0;
  // words are consumed in scan order; the next one is requested a whole word ahead of its use
  const int wlast = (idx + step * (n - 1)) >> 5;  // last word the scan touches
  int wi = idx >> 5;
  u32 w = pl[wi];
  u32 wnext = (wi != wlast) ? pl[wi + step] : 0u;
  for (int t = 0; t < n; ++t) {
    const int b = idx & 31;
    const int m = (int)((w >> b) & 1u);
    const int nsc = max(0, sc + (m ? 1 : -2));
    if (nsc > 0) {
      if (sc == 0) { run_i0 = t; run_best = 0; }
      if (nsc >= run_best) { run_best = nsc; run_best_i = t; }
    } else if (sc > 0) {  // the run [run_i0 .. t-1] closes
      const int len = run_best_i + 1 - run_i0;
      if (len >= 20) {
        const double p = __ddiv_rn((double)((run_best + 2 * len) / 3), (double)len);
        if (p >= 0.85 && (best.len == 0 || p > best.p)) { best.p = p; best.i0 = run_i0; best.len = len; }
      }
    }
    sc = nsc;
    idx += step;
    if ((step > 0) ? (b == 31) : (b == 0)) {
      w = wnext;
      wi += step;
      wnext = (wi != wlast && t + 1 < n) ? pl[wi + step] : 0u;
    }
  }
  if (sc > 0) {
    const int len = run_best_i + 1 - run_i0;
    if (len >= 20) {
      const double p = __ddiv_rn((double)((run_best + 2 * len) / 3), (double)len);
      if (p >= 0.85 && (best.len == 0 || p > best.p)) { best.p = p; best.i0 = run_i0; best.len = len; }
    }
  }
  A.task_res[slot] = best;
}

// K11b': one WARP per long scan task (n >= 1024): the serial recurrence is cut into 64-base chunks, one
// per lane.  A chunk acts on the incoming score as f(x) = max(x + a, b) (a = sum of its deltas, b = its
// final score after the last reset); these maps compose associatively, so a warp scan gives every lane
// its exact incoming score.  Each lane then rescans its chunk with absolute scores and reports: the
// piece of a run continuing from the left (its maximum, last argmax, whether it closes), the best run
// that lies inside the chunk, and the run still open at its right end.  The warp merges the 32 reports
// in position order, which reproduces the serial scan exactly (first maximum of p, last index of a
// run's maximum).  ~1 instruction per base per warp instead of ~10 per base on one thread.

__device__ __forceinline__ unsigned long long poly_fetch64(const u32* __restrict__ pl, int nwords, int idx0, int step,
                                                           int cnt) {
  // bit k (k < cnt) of the result = plane bit at read index idx0 + k*step
  int j0 = step > 0 ? idx0 : idx0 - 63;  // lowest index of the ascending 64-bit window
  int shl = 0;
  if (j0 < 0) { shl = -j0; j0 = 0; }
  const int wq = j0 >> 5, sh = j0 & 31;
  const u32 w0 = wq < nwords ? pl[wq] : 0u;
  const u32 w1 = wq + 1 < nwords ? pl[wq + 1] : 0u;
  const u32 w2 = (sh && wq + 2 < nwords) ? pl[wq + 2] : 0u;
  unsigned long long asc = ((((unsigned long long)w1) << 32) | w0) >> sh;
  if (sh) asc |= ((unsigned long long)w2) << (64 - sh);
  asc <<= shl;
  unsigned long long x = step > 0 ? asc : __brevll(asc);
  if (cnt < 64) x &= ((1ull << cnt) - 1ull);
  return x;
}

__device__ __forceinline__ void poly_offer(PolyRes& best, int i0, int pk, int pk_t) {
  const int len = pk_t + 1 - i0;
  if (len >= 20) {
    const double p = __ddiv_rn((double)((pk + 2 * len) / 3), (double)len);
    if (p >= 0.85 && (best.len == 0 || p > best.p)) { best.p = p; best.i0 = i0; best.len = len; }
  }
}

__global__ void __launch_bounds__(128) k_poly_long(GapArgs A) {
  pdl_prologue();
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int n_long = clips_fit(A.cnt, A.caps) ? A.cls_count[A.long_class - 1] : 0;
  for (int e = gw; e < n_long; e += nwarps) {
    const int slot = A.task_order[e];
    const int i = slot >> 2, which = slot & 3;
    const int clip = slot >> 1;
    const int n = A.clip_n[clip];
    const bool minus = A.read_strand[i] != 0;
    const bool want_a = (which & 1) == 0;
    const u32* pl = clip_plane(A, clip, want_a != minus);
    const ClipGeo geo = clip_geometry(A.read_len[i], n, which < 2, minus);
    const int nwords = geo.n_words, step = geo.step, idx_start = geo.idx0;
    bool open = false;
    int r_i0 = 0, r_best = 0, r_best_t = 0, sc_carry = 0;
    PolyRes best; best.p = 0.0; best.i0 = 0; best.len = 0;
    for (int W0 = 0; W0 < n; W0 += 32 * POLY_C) {
      const int t0 = W0 + lane * POLY_C;
      const int cnt = max(0, min(POLY_C, n - t0));
      const unsigned long long bits = cnt > 0 ? poly_fetch64(pl, nwords, idx_start + t0 * step, step, cnt) : 0ull;
      // pass 1: the chunk as a map x -> max(x + fa, fb)
      int fa = 0, fb = 0;
      {
        int s = 0, mn = 0x3fffffff;
        for (int k = 0; k < cnt; ++k) {
          s += ((bits >> k) & 1ull) ? 1 : -2;
          mn = min(mn, s);
        }
        if (cnt > 0) { fa = s; fb = s - mn; }
      }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {  // inclusive scan of map composition (earlier chunk applied first)
        const int pa = __shfl_up_sync(FULL, fa, o), pb = __shfl_up_sync(FULL, fb, o);
        if (lane >= o) { fb = max(pb + fa, fb); fa = pa + fa; }
      }
      int ea = __shfl_up_sync(FULL, fa, 1), eb = __shfl_up_sync(FULL, fb, 1);
      const int sc_in = lane == 0 ? sc_carry : max(sc_carry + ea, eb);
      const int sc_out_w = max(sc_carry + __shfl_sync(FULL, fa, 31), __shfl_sync(FULL, fb, 31));
      // pass 2: rescan with absolute scores
      int head_max = 0, head_max_t = 0, head_closed = 0;
      PolyRes loc; loc.p = 0.0; loc.i0 = 0; loc.len = 0;
      int cur_i0 = -1, cur_best = 0, cur_best_t = 0, sc = sc_in;
      for (int k = 0; k < cnt; ++k) {
        const int t = t0 + k;
        const int nsc = max(0, sc + (((bits >> k) & 1ull) ? 1 : -2));
        if (nsc > 0) {
          if (sc == 0) { cur_i0 = t; cur_best = 0; }
          if (nsc >= cur_best) { cur_best = nsc; cur_best_t = t; }
        } else if (sc > 0) {
          if (cur_i0 < 0) { head_closed = 1; head_max = cur_best; head_max_t = cur_best_t; }
          else poly_offer(loc, cur_i0, cur_best, cur_best_t);
        }
        sc = nsc;
      }
      const int tail_open = (cnt > 0 && sc > 0) ? 1 : 0;
      const int tail_i0 = cur_i0;
      if (tail_open && cur_i0 < 0) { head_max = cur_best; head_max_t = cur_best_t; }  // run spans the whole chunk
      // merge the 32 reports in position order (every lane computes the same state)
      for (int l = 0; l < 32; ++l) {
        const int c_l = __shfl_sync(FULL, cnt, l);
        const int hm = __shfl_sync(FULL, head_max, l), hmt = __shfl_sync(FULL, head_max_t, l);
        const int hc = __shfl_sync(FULL, head_closed, l);
        const double lp = __shfl_sync(FULL, loc.p, l);
        const int li0 = __shfl_sync(FULL, loc.i0, l), ll = __shfl_sync(FULL, loc.len, l);
        const int to = __shfl_sync(FULL, tail_open, l), ti0 = __shfl_sync(FULL, tail_i0, l);
        const int tm = __shfl_sync(FULL, cur_best, l), tmt = __shfl_sync(FULL, cur_best_t, l);
        if (c_l == 0) continue;
        if (open) {
          if (hm >= r_best) { r_best = hm; r_best_t = hmt; }
          if (hc) { poly_offer(best, r_i0, r_best, r_best_t); open = false; }
        }
        if (ll > 0 && (best.len == 0 || lp > best.p)) { best.p = lp; best.i0 = li0; best.len = ll; }
        if (to && ti0 >= 0) { open = true; r_i0 = ti0; r_best = tm; r_best_t = tmt; }
      }
      sc_carry = sc_out_w;
    }
    if (open) poly_offer(best, r_i0, r_best, r_best_t);
    if (lane == 0) A.task_res[slot] = best;
  }
}

// K11c: per read, pick the poly candidates (A offered before T, first maximum of p wins, :392-408)
// and write the final head fields (:407-420, :438-454).
__global__ void k_gap_finish(GapArgs A) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n_reads || !clips_fit(A.cnt, A.caps)) return;
  int* head = A.read_head + (i64)i * 8;
  if (!(head[0] & 1)) return;
  const int L = A.read_len[i];
  const int q_ssc = head[3], q_esc = head[6];
  const int* cn = A.clip_n + (i64)i * 2;
  const PolyRes* rs = A.task_res + (i64)i * 4;
  int flags = 1;
  {
    int kind = 0; PolyRes b; b.p = 0; b.i0 = 0; b.len = 0;
    if (cn[0] >= 20 && rs[0].len > 0) { b = rs[0]; kind = 1; }
    if (cn[0] >= 20 && rs[1].len > 0 && (kind == 0 || rs[1].p > b.p)) { b = rs[1]; kind = 2; }
    if (kind) {
      int gap = q_ssc - b.i0 - b.len;
      if (!(0 <= gap && gap < q_ssc)) { dev_fail(A.err, DEVERR_POLY_RANGE, i); return; }
      flags |= kind << 8;
      head[1] = b.len; head[2] = gap; head[3] = b.i0;
    }
  }
  {
    int kind = 0; PolyRes b; b.p = 0; b.i0 = 0; b.len = 0;
    if (cn[1] >= 20 && rs[2].len > 0) { b = rs[2]; kind = 1; }
    if (cn[1] >= 20 && rs[3].len > 0 && (kind == 0 || rs[3].p > b.p)) { b = rs[3]; kind = 2; }
    if (kind) {
      int esc = L - q_esc - b.i0;
      if (!(b.i0 >= 0 && b.i0 < L - q_esc && esc > 0)) { dev_fail(A.err, DEVERR_POLY_RANGE, i); return; }
      flags |= kind << 16;
      head[4] = b.len; head[5] = b.i0; head[6] = esc;
    } else {
      head[6] = L - q_esc;
    }
  }
  head[0] = flags;
}
