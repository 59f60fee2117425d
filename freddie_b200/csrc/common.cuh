// common.cuh -- shared helpers: error handling, device scans / compaction, threshold arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef long long i64;
typedef unsigned int u32;
typedef unsigned char u8;

#define FRS_NEG_INF (-0x3fffffff)  // "-inf" of the DP (scores fit in 30 bits)

// Programmatic dependent launch (sm_90+): first statement of every kernel.  `wait` holds the grid until the grid
// before it in the stream has completed and its writes are visible; `launch_dependents` then lets the NEXT kernel of
// the stream (when it was launched with cudaLaunchAttributeProgrammaticStreamSerialization, see launch_k in frs.cu)
// become resident as soon as every CTA of this grid has passed that point.  The stream keeps its sequential
// meaning -- what overlaps is the launch latency and the CTA ramp-up of a kernel with the execution of its
// predecessor (a run is ~60 short launches).  The order matters: with the trigger BEFORE the wait a grid's grandchild
// can become resident while the grandparent still runs, and that faulted on B200 (illegal address in the run);
// wait-then-trigger keeps at most two grids of a stream in flight and is bit-exact on the whole GPU suite.  Both
// instructions are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
}

// device-side assert channel: first failing code wins
enum {
  DEVERR_NONE = 0,
  DEVERR_BREAK_LARGE_POS = 1,   // freddie_segment.py:643  assert max_c_idx_y_v > 0
  DEVERR_BREAK_LARGE_RANGE = 2, // freddie_segment.py:640  window leaves the candidate list
  DEVERR_RATIO_RANGE = 3,       // freddie_segment.py:821  assert 0 <= cov_ratio <= 1
  DEVERR_THREAD_CIGAR = 4,      // freddie_segment.py:303,326,349  CIGAR threading failed
  DEVERR_Q_RANGE = 5,           // freddie_segment.py:389  0 <= q_ssc <= q_esc <= length
  DEVERR_GAP_RANGE = 6,         // freddie_segment.py:462,466
  DEVERR_POLY_RANGE = 7,        // freddie_segment.py:410,441,450
  DEVERR_BACKTRACE = 8,         // internal: DP backtrace ran off the table
  DEVERR_ISLAND = 9,            // freddie_segment.py:668  interval ends in different islands
  DEVERR_DP_SMEM = 10,          // limit: a subproblem does not fit the DP kernel's shared memory
  DEVERR_SCORE_RANGE = 11,      // limit: DP scores of a tint could leave the 30-bit range
};

// ---------------------------------------------------------------------------------------------
// Device-side counters of one run: i64 slots written by the kernels and read by the host ONCE, at the
// end of the run.  No launch of the pipeline depends on a host round trip: grids are sized from upper
// bounds known at upload time (or are persistent) and read their true extents here; buffers whose size
// is data-dependent and not linearly bounded (coverage matrix, DP tables, digits, ...) have a capacity
// (Caps) -- a kernel that would leave it returns without writing, and the host grows the buffer and
// repeats the run (first batches of a context only; capacities are grow-only).
// ---------------------------------------------------------------------------------------------
enum {
  CNT_K = 0,       // candidates
  CNT_NWORK = 2,   // DP work items (sum over classes)
  CNT_COV = 3,     // coverage elements
  CNT_REF2 = 9,    // refine work list after the filter (int)
  CNT_REF = 10,    // refine work list (int)
  CNT_NFIN = 11,   // final positions
  CNT_NDIG = 12,   // digit bytes
  CNT_NRUN = 13,   // 1-runs over all reps
  CNT_NGAP = 14,   // gap records
  CNT_CLIPW = 15,  // plane words of the soft clips (lazy sequence mode)
  CNT_PLAN = 16,   // PLAN_SLOTS slots of the subproblem plan (kernels_dp.cuh)
  CNT_ERR = 40,    // 4 ints: first device assert (code, item), poly tasks, long poly tasks
  CNT_BUCKET = 48, // DP work items per (class, cost bucket): DP_CLASSES x DP_BUCKETS slots (kernels_dp.cuh)
  CNT_SLOTS = 96
};
struct Caps { i64 P, tab, work, split, dig, runs, gaps, clipw; };  // capacities in elements

__device__ __forceinline__ void dev_fail(int* err, int code, int where) {
  if (atomicCAS(&err[0], 0, code) == 0) err[1] = where;
}

// ---------------------------------------------------------------------------------------------
// thresholds: the reference compares the IEEE quotient  double(cov)/double(len)  against h and
// l = 1-h (freddie_segment.py:490-495, :816-828).  The quotient is monotone in cov, so the two
// compares are equivalent to integer cuts computed once per length with true fp64 divides:
//   yea  <=>  cov >= ty      nay  <=>  cov <= tn
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void length_cuts(int len, const double* __restrict__ tbl, int tbl_len, double tp,
                                            int& ty, int& tn) {
  double h = (len < tbl_len) ? tbl[len] : tp;
  double l = __dsub_rn(1.0, h);
  double dl = (double)len;
  int c = (int)floor(h * dl);
  if (c < 0) c = 0;
  while (c > 0 && __ddiv_rn((double)(c - 1), dl) > h) --c;
  while (!(__ddiv_rn((double)c, dl) > h)) ++c;  // ends at len+1 at the latest (ratio > 1 >= h)
  ty = c;
  int d = (int)floor(l * dl);
  if (d < 0) d = 0;
  while (__ddiv_rn((double)d, dl) < l) ++d;  // first d with d/len >= l
  while (d > 0 && !(__ddiv_rn((double)(d - 1), dl) < l)) --d;
  tn = d - 1;
}

// The cuts depend on the length alone: a table for the lengths below CUT_TAB_N (one tiny launch per run)
// replaces the fp64 divides in the kernels that need cuts per candidate pair / per segment.
#define CUT_TAB_N 8192
__global__ void k_cut_table(const double* __restrict__ tbl, int tbl_len, double tp, int2* __restrict__ cut_tab) {
  pdl_prologue();
  const int len = blockIdx.x * blockDim.x + threadIdx.x;
  if (len >= CUT_TAB_N) return;
  int ty = 0x7fffffff, tn = -1;
  if (len >= 1) length_cuts(len, tbl, tbl_len, tp, ty, tn);
  cut_tab[len] = make_int2(ty, tn);
}
__device__ __forceinline__ void length_cuts_t(int len, const int2* __restrict__ cut_tab, const double* __restrict__ tbl,
                                              int tbl_len, double tp, int& ty, int& tn) {
  if (len >= 1 && len < CUT_TAB_N) {
    const int2 v = __ldg(&cut_tab[len]);
    ty = v.x;
    tn = v.y;
  } else {
    length_cuts(len, tbl, tbl_len, tp, ty, tn);
  }
}

// ---------------------------------------------------------------------------------------------
// block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total_out, T* smem /* >= 33 entries */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  T x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) smem[warp] = x;
  __syncthreads();
  if (warp == 0) {
    T s = (lane < nw) ? smem[lane] : (T)0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T y = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += y;
    }
    if (lane < nw) smem[lane] = s;  // inclusive warp totals
    if (lane == 31) smem[32] = s;
  }
  __syncthreads();
  T base = (warp > 0) ? smem[warp - 1] : (T)0;
  if (total_out) *total_out = smem[32 > nw ? nw - 1 : 31];
  T r = base + x - v;
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------------
// device-wide exclusive scan / flag compaction: 3 launches (block sums, scan of sums, apply)
// ---------------------------------------------------------------------------------------------
#define SCAN_THREADS 512
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

template <typename TIn>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block_sums(const TIn* __restrict__ in, i64 n,
                                                                 i64* __restrict__ bsum) {
  pdl_prologue();
  __shared__ i64 sm[40];
  i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
  i64 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k)
    if (base + k < n) s += (i64)in[base + k];
  i64 tot;
  block_exclusive_scan<i64>(s, &tot, sm);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

// single CTA: in-place exclusive scan of bsum[0..nb), total written to bsum[nb]
__global__ void __launch_bounds__(1024) k_scan_bsums(i64* __restrict__ bsum, int nb) {
  pdl_prologue();
  __shared__ i64 sm[40];
  __shared__ i64 carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += 1024) {
    int i = base + threadIdx.x;
    i64 v = (i < nb) ? bsum[i] : 0;
    i64 tot;
    i64 ex = block_exclusive_scan<i64>(v, &tot, sm);
    i64 c = carry;
    if (i < nb) bsum[i] = c + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry = c + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry;
}

// out[i] = exclusive prefix (TOut), out[n] = total
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TIn* __restrict__ in, i64 n,
                                                            const i64* __restrict__ bsum, TOut* __restrict__ out,
                                                            i64* __restrict__ total_out /* or NULL */) {
  pdl_prologue();
  __shared__ i64 sm[40];
  i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
  i64 v[SCAN_ITEMS];
  i64 s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    v[k] = (base + k < n) ? (i64)in[base + k] : 0;
    s += v[k];
  }
  i64 ex = block_exclusive_scan<i64>(s, (i64*)nullptr, sm) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) {
    if (base + k < n) out[base + k] = (TOut)ex;
    ex += v[k];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
    out[n] = (TOut)bsum[gridDim.x];
    if (total_out) *total_out = bsum[gridDim.x];  // the total also goes to the run's counter block (no copy launch)
  }
}

// short arrays (per-tint tables): one CTA, one launch.  out[i] = exclusive prefix, out[n] = total
#define SCAN_SMALL_MAX 65536
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(1024) k_scan_small(const TIn* __restrict__ in, int n, TOut* __restrict__ out,
                                                     i64* __restrict__ total_out /* or NULL */) {
  pdl_prologue();
  __shared__ i64 sm[40];
  const int per = (n + 1023) / 1024;
  const int i0 = min(n, (int)threadIdx.x * per), i1 = min(n, i0 + per);
  i64 s = 0;
  for (int i = i0; i < i1; ++i) s += (i64)in[i];
  i64 tot;
  i64 ex = block_exclusive_scan<i64>(s, &tot, sm);
  for (int i = i0; i < i1; ++i) {
    const i64 v = (i64)in[i];
    out[i] = (TOut)ex;
    ex += v;
  }
  if (threadIdx.x == 0) {
    out[n] = (TOut)tot;
    if (total_out) *total_out = tot;
  }
}

// byte-flag compaction, idx_out[rank] = i for every i with flags[i] != 0 (ascending): 16 flags per
// thread from one 16-byte load
#define FLAG_ITEMS 16
#define FLAG_TILE (SCAN_THREADS * FLAG_ITEMS)
__device__ __forceinline__ uint4 flag_load16(const u8* __restrict__ flags, i64 base, i64 n) {
  if (base + FLAG_ITEMS <= n) return *reinterpret_cast<const uint4*>(flags + base);  // base is a multiple of 16
  u32 w[4] = {0u, 0u, 0u, 0u};
  for (int k = 0; k < FLAG_ITEMS; ++k)
    if (base + k < n && flags[base + k]) w[k >> 2] |= 1u << ((k & 3) * 8);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ u32 flag_bits16(uint4 v) {  // bit k = flag k != 0
  auto nzb = [](u32 x) -> u32 {  // 4 bytes -> 4 bits
    u32 m = (x | (x >> 4)) & 0x0f0f0f0fu;
    m = (m | (m >> 2)) & 0x03030303u;
    m = (m | (m >> 1)) & 0x01010101u;
    return (m | (m >> 7) | (m >> 14) | (m >> 21)) & 0xfu;
  };
  return nzb(v.x) | (nzb(v.y) << 4) | (nzb(v.z) << 8) | (nzb(v.w) << 12);
}
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_sums(const u8* __restrict__ flags, i64 n, i64* __restrict__ bsum) {
  pdl_prologue();
  __shared__ int sm[40];
  const i64 base = (i64)blockIdx.x * FLAG_TILE + (i64)threadIdx.x * FLAG_ITEMS;
  const int s = (base < n) ? __popc(flag_bits16(flag_load16(flags, base, n))) : 0;
  int tot;
  block_exclusive_scan<int>(s, &tot, sm);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_flag_compact(const u8* __restrict__ flags, i64 n,
                                                              const i64* __restrict__ bsum, int* __restrict__ idx_out,
                                                              i64* __restrict__ count_out /* or NULL */) {
  pdl_prologue();
  __shared__ int sm[40];
  if (count_out && blockIdx.x == 0 && threadIdx.x == 0) *count_out = bsum[gridDim.x];
  const i64 base = (i64)blockIdx.x * FLAG_TILE + (i64)threadIdx.x * FLAG_ITEMS;
  u32 bits = (base < n) ? flag_bits16(flag_load16(flags, base, n)) : 0u;
  i64 ex = (i64)block_exclusive_scan<int>(__popc(bits), (int*)nullptr, sm) + bsum[blockIdx.x];
  while (bits) {
    const int k = __ffs(bits) - 1;
    bits &= bits - 1;
    idx_out[ex++] = (int)(base + k);
  }
}

// Zero fill and small copies as KERNELS: a run must not touch the copy engines, which carry the next batch in
// and the previous one out (a cudaMemsetAsync / device-to-device cudaMemcpyAsync of a run queues behind
// those transfers: measured, the head of a pipelined run took 2.7 ms instead of 1.9 ms).
__global__ void k_zero16(uint4* __restrict__ p, size_t n16) {
  pdl_prologue();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    p[i] = make_uint4(0u, 0u, 0u, 0u);
}
// several zero fills in ONE launch: the buffers a run needs zeroed (counters, raw signal, group sums, refine flags,
// run counts) are all free when the run starts, so they share the large fill's launch instead of paying ~4 us each
struct ZeroList { uint4* p[6]; size_t n16[6]; int n; };  // n16[k] = 16-byte words of region k
__global__ void k_zero_multi(ZeroList z) {
  pdl_prologue();
  const size_t t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
#pragma unroll  // constant indices: the list stays in the parameter bank (a runtime index copies it to local memory)
  for (int k = 0; k < 6; ++k)
    if (k < z.n)
      for (size_t i = t0; i < z.n16[k]; i += stride) z.p[k][i] = make_uint4(0u, 0u, 0u, 0u);
}
// dst[k] = (i64) value at src[k] for up to 4 scattered words: the totals of scans into the counter block
struct CopyWords { i64* dst[4]; const void* src[4]; int bytes[4]; int n; };
__global__ void k_copy_words(CopyWords w) {
  pdl_prologue();
  const int k = threadIdx.x;
  if (k < w.n) *w.dst[k] = w.bytes[k] == 8 ? *(const i64*)w.src[k] : (i64)*(const int*)w.src[k];
}
// final flags of the DP start as the fixed flags (only the candidates that exist)
__global__ void k_copy_flags(const i64* __restrict__ n_p, const u8* __restrict__ src, u8* __restrict__ dst) {
  pdl_prologue();
  const i64 n = *n_p;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) dst[i] = src[i];
}

// compact batch encodings -> the arrays the kernels read (frs_batch.cigar16 / riv_cig_n / qe_from_cigar)
__global__ void k_widen_u16(const unsigned short* __restrict__ in, i64 n, u32* __restrict__ out) {
  pdl_prologue();
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) out[i] = in[i];
}
__global__ void k_derive_qe(i64 n_ivs, const int* __restrict__ qs, const int* __restrict__ cig_off,
                            const u32* __restrict__ cigar, int* __restrict__ qe) {
  pdl_prologue();
  for (i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x; k < n_ivs; k += (i64)gridDim.x * blockDim.x) {
    int q = qs[k];
    for (int c = cig_off[k]; c < cig_off[k + 1]; ++c) {
      const u32 op = cigar[c];
      if ((op & 15u) <= 1u) q += (int)(op >> 4);  // 0: M/X/=, 1: I consume query bases
    }
    qe[k] = q;
  }
}

// largest i in [0, n) with off[i] <= x  (off ascending, off[0] <= x)
__device__ __forceinline__ int upper_row(const int* __restrict__ off, int n, int x) {
  int lo = 0, hi = n;  // invariant: off[lo] <= x, (hi == n or off[hi] > x)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_row64(const i64* __restrict__ off, int n, i64 x) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// tint of every island / read rep / read (binary search in the tint offset tables)
__global__ void k_owner_tables(int T, int n_islands, int n_reps, int n_reads, const int* __restrict__ tint_island_off,
                               const int* __restrict__ tint_rep_off, const int* __restrict__ tint_read_off,
                               int* __restrict__ island_tint, int* __restrict__ rep_tint, int* __restrict__ read_tint) {
  pdl_prologue();
  long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n_islands) { island_tint[e] = upper_row(tint_island_off, T, (int)e); return; }
  e -= n_islands;
  if (e < n_reps) { rep_tint[e] = upper_row(tint_rep_off, T, (int)e); return; }
  e -= n_reps;
  if (e < n_reads) read_tint[e] = upper_row(tint_read_off, T, (int)e);
}

// Genomic target coordinates of every read's intervals from its rep's flat-sample intervals: a read's
// target intervals ARE its rep's (the dedupe key of read_split, freddie_segment.py:165-170), so a caller
// may leave frs_batch.riv_ts / riv_te NULL and save their host-to-device copy.  Eight lanes per read (a read has
// ~8 intervals): the lanes' loads of a read's intervals are in flight together instead of one dependent chain
// per read (one thread per read: 54 us for 193 k reads).
__global__ void k_derive_riv(int n_reads, int n_islands, const int* __restrict__ read_rep, const int* __restrict__ read_iv_off,
                             const int* __restrict__ rep_iv_off, const int* __restrict__ rep_fs, const int* __restrict__ rep_fe,
                             const int* __restrict__ island_sample_off, const int* __restrict__ island_start,
                             int* __restrict__ riv_ts, int* __restrict__ riv_te) {
  pdl_prologue();
  const int g = threadIdx.x & 7;
  const int r = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
  if (r >= n_reads) return;
  const int k0 = read_iv_off[r], k1 = read_iv_off[r + 1];
  const int q0 = rep_iv_off[read_rep[r]];
  int lo = 0, hi = -1, base = 0;
  for (int k = k0 + g; k < k1; k += 8) {
    const int q = q0 + (k - k0);
    const int fs = rep_fs[q], fe = rep_fe[q];
    if (fs < lo || fs > hi) {  // intervals of a read are sorted: a lane's next interval often stays in its island
      const int isl = upper_row(island_sample_off, n_islands, fs);
      lo = island_sample_off[isl];
      hi = island_sample_off[isl + 1] - 1;
      base = island_start[isl] - lo;
    }
    riv_ts[k] = fs + base;
    riv_te[k] = fe + base;
  }
}
