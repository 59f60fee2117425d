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
// two pointers; a warp writes 128 contiguous bytes of one row per step.
// ---------------------------------------------------------------------------------------------
#define COV_THREADS 128
struct RepTile { int tint; int rep_lo; };  // rep_lo: tint-local first rep of the tile

__global__ void __launch_bounds__(COV_THREADS) k_coverage(const RepTile* __restrict__ tiles,
                                                         const int* __restrict__ tint_rep_off,
                                                         const int* __restrict__ tint_cand_off,
                                                         const i64* __restrict__ tint_cov_off,
                                                         const int* __restrict__ rep_iv_off,
                                                         const int* __restrict__ iv_fs, const int* __restrict__ iv_fe,
                                                         const int* __restrict__ cand_flat, u32* __restrict__ P) {
  const RepTile tl = tiles[blockIdx.x];
  const int r0 = tint_rep_off[tl.tint];
  const int R = tint_rep_off[tl.tint + 1] - r0;
  const int Rp = (R + 3) & ~3;
  const int r = tl.rep_lo + threadIdx.x;
  if (r >= Rp) return;
  const int q0 = tint_cand_off[tl.tint], q1 = tint_cand_off[tl.tint + 1];
  u32* out = P + tint_cov_off[tl.tint] + r;
  if (r >= R) {  // padding columns
    for (int q = q0; q < q1; ++q) out[(i64)(q - q0) * Rp] = 0u;
    return;
  }
  int a = rep_iv_off[r0 + r];
  const int b = rep_iv_off[r0 + r + 1];
  u32 acc = 0;
  int fs = (a < b) ? iv_fs[a] : 0x7fffffff;
  int fe = (a < b) ? iv_fe[a] : 0x7fffffff;
  for (int q = q0; q < q1; ++q) {
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
__global__ void k_fixed_a(int n_cand, const int* __restrict__ cand_flat, const int* __restrict__ cand_island,
                          const int* __restrict__ island_cand_off, const int* __restrict__ island_tint,
                          const double* __restrict__ y, const double* __restrict__ thr, u8* __restrict__ fixed0,
                          u8* __restrict__ fixed1) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_cand) return;
  int isl = cand_island[q];
  bool f = (q == island_cand_off[isl]) || (q == island_cand_off[isl + 1] - 1);
  if (!f) f = y[cand_flat[q]] > thr[island_tint[isl]];
  fixed0[q] = f;
  fixed1[q] = f;
}

__global__ void k_fixed_b(int n_cand, const int* __restrict__ cand_flat, const int* __restrict__ cand_island,
                          const int* __restrict__ island_cand_off, const double* __restrict__ y, int mps,
                          const u8* __restrict__ fixed0, u8* __restrict__ fixed1, int* __restrict__ err) {
  int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_cand || !fixed0[q]) return;
  int isl = cand_island[q];
  int c0 = island_cand_off[isl], c1 = island_cand_off[isl + 1];
  if (q == c1 - 1) return;
  int e = q + 1;
  while (!fixed0[e]) ++e;  // the island's last candidate is fixed
  int size = e - q + 1;
  if (size <= mps) return;
  int cnt = (int)ceil((double)size / (double)mps);
  double ps = __ddiv_rn((double)size, (double)cnt);
  int s_local = q - c0;
  for (int i = 1; i < cnt; ++i) {
    int mid = (int)__dadd_rn((double)s_local, __dmul_rn((double)i, ps));
    double best = -INFINITY;
    int best_c = -1;
    for (int c = mid - 5; c < mid + 5; ++c) {
      if (c < 0 || c0 + c >= c1) { dev_fail(err, DEVERR_BREAK_LARGE_RANGE, q); return; }
      double v = y[cand_flat[c0 + c]];
      if (v > best) { best = v; best_c = c; }
    }
    if (!(best > 0.0)) { dev_fail(err, DEVERR_BREAK_LARGE_POS, q); return; }
    fixed1[c0 + best_c] = 1;
  }
}

// subproblems: consecutive fixed candidates (a, b) of one island with at least one interior
// candidate.  flag over the fixed list, then compaction gives the subproblem list.
__global__ void k_sub_flag(int n_fixed, const int* __restrict__ fixed_list, const int* __restrict__ cand_island,
                           u8* __restrict__ flag) {
  int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_fixed) return;
  u8 v = 0;
  if (f + 1 < n_fixed) {
    int a = fixed_list[f], b = fixed_list[f + 1];
    v = (cand_island[a] == cand_island[b] && b - a >= 2) ? 1 : 0;
  }
  flag[f] = v;
}

// per subproblem sizes for the offset scans.  slab_words: words of 32 reps handled by one CTA.
__global__ void k_sub_sizes(int n_sub, const int* __restrict__ sub_fidx, const int* __restrict__ fixed_list,
                            const int* __restrict__ cand_island, const int* __restrict__ island_tint,
                            const int* __restrict__ tint_rep_off, int slab_words, int* __restrict__ sub_start,
                            int* __restrict__ sub_n, int* __restrict__ sub_tint, int* __restrict__ sz_pair,
                            int* __restrict__ sz_triple, int* __restrict__ sz_work, i64* __restrict__ stats) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_sub) return;
  int f = sub_fidx[p];
  int a = fixed_list[f], b = fixed_list[f + 1];
  int n = b - a + 1;
  int t = island_tint[cand_island[a]];
  int R = tint_rep_off[t + 1] - tint_rep_off[t];
  int words = (R + 31) >> 5;
  sub_start[p] = a;
  sub_n[p] = n;
  sub_tint[p] = t;
  sz_pair[p] = n * n;
  i64 t3 = (i64)n * (n - 1) * (n - 2) / 6;
  sz_triple[p] = (int)t3;
  sz_work[p] = (words + slab_words - 1) / slab_words;
  atomicAdd((unsigned long long*)&stats[0], (unsigned long long)t3);
  atomicAdd((unsigned long long*)&stats[1], (unsigned long long)(t3 * R));
  atomicMax((int*)&stats[2], n);
}

// ---------------------------------------------------------------------------------------------
// K7 DP tables.  One CTA = (subproblem, slab of read-rep words).  Per chunk of Wc words:
//   1. TMA bulk copies stage the n coverage rows x 32*Wc reps of the chunk into shared memory;
//   2. mask phase: a warp owns a row i, a lane owns a rep; for every j>i the lanes compare
//      cov = P[j]-P[i] with the pair's integer cuts and __ballot_sync packs the yea / nay bits of
//      32 reps into one word each; ambiguous counts (ins, :500-506) are accumulated on the fly;
//   3. triple phase: out(i,j,k) = sum_w W . [(yea_ij & nay_jk) | (nay_ij & yea_jk)]  (:509-528) as
//      weighted popcounts over the packed words; a warp owns the middle candidate j, its lanes the
//      left candidate i, the loop runs over k with broadcast shared-memory loads.
// Weights: bit-planes of W per word (plane b = reps whose weight has bit b), so weight-1 data costs
// one AND + POPC.  Non-zero partial sums are added to the global tables with RED.
// out layout per subproblem: [j][i][k-j-1], j = 1..n-2, i < j < k  -> exactly C(n,3) entries.
// ---------------------------------------------------------------------------------------------
#define DPT_THREADS 512
#define DPT_MAXW 8

__device__ __forceinline__ int pair_index(int i, int j, int n) { return i * (2 * n - i - 1) / 2 + (j - i - 1); }
__device__ __forceinline__ int triple_mid_off(int j, int n) {
  // sum_{j'=1}^{j-1} j' (n-1-j')
  int m = j - 1;
  return (n - 1) * m * (m + 1) / 2 - m * (m + 1) * (2 * m + 1) / 6;
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

struct DptArgs {
  const int* sub_start; const int* sub_n; const int* sub_tint; const int* sub_work_off;
  const i64* sub_pair_off; const i64* sub_triple_off;
  int n_sub;
  const int* tint_rep_off; const int* tint_cand_off; const i64* tint_cov_off;
  const int* rep_weight; const int* cand_flat; const u32* P;
  const double* thr_table; int thr_table_len; double tp;
  int slab_words; int max_n;
  int* ins; int* out;
};

// shared-memory carve-up for a given (n_max, Wc)
__host__ __device__ inline size_t dpt_smem_bytes(int n, int wc) {
  size_t p2 = (size_t)n * (n - 1) / 2;
  size_t b = 0;
  b += 16;                             // mbarrier
  b += (size_t)n * 4;                  // cf
  b += p2 * 8;                         // ty, tn
  b += (size_t)n * 32 * wc * 4;        // coverage tile
  b += p2 * wc * 8;                    // yea/nay words (uint2)
  b += (size_t)wc * 32 * 4 + wc * 4;   // weight planes + plane counts
  return (b + 15) & ~(size_t)15;
}

__global__ void __launch_bounds__(DPT_THREADS, 1) k_dp_tables(DptArgs A, int wc) {
  extern __shared__ __align__(16) unsigned char dsm[];
  // work item -> (subproblem, slab)
  const int p = upper_row(A.sub_work_off, A.n_sub, (int)blockIdx.x);
  const int slab = blockIdx.x - A.sub_work_off[p];
  const int n = A.sub_n[p];
  const int qs = A.sub_start[p];
  const int t = A.sub_tint[p];
  const int r0 = A.tint_rep_off[t];
  const int R = A.tint_rep_off[t + 1] - r0;
  const int Rp = (R + 3) & ~3;
  const int words = (R + 31) >> 5;
  const int w_lo = slab * A.slab_words;
  const int w_hi = min(words, w_lo + A.slab_words);
  const int p2 = n * (n - 1) / 2;
  const int CW = 32 * wc;
  // carve shared memory (sized for max_n by the host)
  unsigned long long* bar = (unsigned long long*)dsm;
  u32* tile = (u32*)(dsm + 16);                      // [n][CW]   (16-byte aligned rows)
  uint2* ynm = (uint2*)(tile + (size_t)A.max_n * CW); // [p2][wc]  x = yea, y = nay
  int* cf = (int*)(ynm + (size_t)(A.max_n * (A.max_n - 1) / 2) * wc);
  int* ty = cf + A.max_n;
  int* tn = ty + (A.max_n * (A.max_n - 1) / 2);
  u32* planes = (u32*)(tn + (A.max_n * (A.max_n - 1) / 2));  // [wc][32]
  int* nplanes = (int*)(planes + wc * 32);                   // [wc]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = DPT_THREADS / 32;
  const u32* Prow0 = A.P + A.tint_cov_off[t] + (i64)(qs - A.tint_cand_off[t]) * Rp;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < n; i += DPT_THREADS) cf[i] = A.cand_flat[qs + i];
  __syncthreads();
  // first chunk's TMA can fly while the cuts are computed
  u32 phase = 0;
  auto issue = [&](int w0) {
    int col0 = w0 * 32;
    int cols = min(CW, Rp - col0);
    u32 bytes = (u32)cols * 4u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar, bytes * (u32)n);
    for (int i = 0; i < n; ++i) tma_bulk_g2s(tile + (size_t)i * CW, Prow0 + (i64)i * Rp + col0, bytes, bar);
  };
  if (tid == 0 && w_lo < w_hi) issue(w_lo);
  for (int e = tid; e < p2; e += DPT_THREADS) {
    // decode pair e -> (i, j): rows are short, a linear walk is fine (done once per CTA)
    int i = 0, rem = e;
    while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
    int j = i + 1 + rem;
    int a, b;
    length_cuts(cf[j] - cf[i] + 1, A.thr_table, A.thr_table_len, A.tp, a, b);
    ty[e] = a;
    tn[e] = b;
  }
  int* ins_g = A.ins + A.sub_pair_off[p];
  int* out_g = A.out + A.sub_triple_off[p];

  for (int w0 = w_lo; w0 < w_hi; w0 += wc) {
    const int nw = min(wc, w_hi - w0);
    // weight planes of the chunk
    for (int w = warp; w < nw; w += NW) {
      int rep = (w0 + w) * 32 + lane;
      int wt = (rep < R) ? A.rep_weight[r0 + rep] : 0;
      int mx = wt;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      int np = 32 - __clz(mx);
      for (int b = 0; b < np; ++b) {
        u32 m = __ballot_sync(0xffffffffu, (wt >> b) & 1);
        if (lane == 0) planes[w * 32 + b] = m;
      }
      if (lane == 0) nplanes[w] = np;
    }
    __syncthreads();  // cuts + planes visible; previous chunk's triple phase done
    mbar_wait(bar, phase);
    phase ^= 1;
    // ---- mask phase: rows i paired from both ends for balance ----
    const int nrows = n - 1;                 // rows i = 0..n-2
    const int nrp = (nrows + 1) / 2;         // row a is processed together with row nrows-1-a
    for (int a = warp; a < nrp; a += NW) {
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        const int i = side ? (nrows - 1 - a) : a;
        if (side && i == a) break;
        const u32* rowi = tile + (size_t)i * CW;
        const int ebase = pair_index(i, i + 1, n);
        for (int j = i + 1; j < n; ++j) {
          const int e = ebase + (j - i - 1);
          const int cy = ty[e], cn = tn[e];
          const u32* rowj = tile + (size_t)j * CW;
          int amb = 0;
          for (int w = 0; w < nw; ++w) {
            const int rep = (w0 + w) * 32 + lane;
            const bool valid = rep < R;
            const int cov = (int)(rowj[w * 32 + lane] - rowi[w * 32 + lane]);
            const u32 by = __ballot_sync(0xffffffffu, valid && cov >= cy);
            const u32 bn = __ballot_sync(0xffffffffu, valid && cov <= cn);
            if (lane == 0) {
              ynm[(size_t)e * wc + w] = make_uint2(by, bn);
              const int left = R - (w0 + w) * 32;
              const u32 vm = (left >= 32) ? 0xffffffffu : ((1u << left) - 1u);
              const u32 am = vm & ~(by | bn);
              if (am) {
                const int np = nplanes[w];
                for (int b = 0; b < np; ++b) amb += __popc(am & planes[w * 32 + b]) << b;
              }
            }
          }
          if (lane == 0 && amb) atomicAdd(&ins_g[i * n + j], amb);
        }
      }
    }
    __syncthreads();  // masks complete, tile free
    if (tid == 0 && w0 + wc < w_hi) issue(w0 + wc);
    // ---- triple phase ----
    for (int j = 1 + warp; j <= n - 2; j += NW) {
      const int cols = n - 1 - j;
      int* outj = out_g + triple_mid_off(j, n);
      for (int i = lane; i < j; i += 32) {
        const uint2* yn_ij = ynm + (size_t)pair_index(i, j, n) * wc;
        uint2 rij[DPT_MAXW];
#pragma unroll
        for (int w = 0; w < DPT_MAXW; ++w) rij[w] = (w < nw) ? yn_ij[w] : make_uint2(0u, 0u);
        const uint2* yn_jk = ynm + (size_t)pair_index(j, j + 1, n) * wc;
        for (int k = j + 1; k < n; ++k, yn_jk += wc) {
          int acc = 0;
#pragma unroll
          for (int w = 0; w < DPT_MAXW; ++w) {
            if (w < nw) {
              uint2 jk = yn_jk[w];
              u32 m = (rij[w].x & jk.y) | (rij[w].y & jk.x);
              if (m) {
                int np = nplanes[w];
                for (int b = 0; b < np; ++b) acc += __popc(m & planes[w * 32 + b]) << b;
              }
            }
          }
          if (acc) atomicAdd(&outj[i * cols + (k - j - 1)], acc);
        }
      }
    }
    __syncthreads();  // the next chunk rewrites the weight planes and the masks: wait for every warp's triple phase
  }
}

// ---------------------------------------------------------------------------------------------
// K8 DP solve (one CTA per subproblem).  G(j,k) = best continuation after committing segment (j,k):
//   G(j,E) = ins(j,E);  G(j,k) = max_{k'>k} D(j,k,k')  (ascending k', first maximum wins)
//   D(i,j,k) = ins(i,j) + out(i,j,k) + G(j,k)  if both segments span >= 5 samples, out >= lo and
//              G(j,k) is finite, else -inf                                   (:532-558)
// Top level (:560-566): lexicographic (j,k), strict improvement over ins(0,E).  Backtrace (:592-594)
// marks the chosen candidates.
// ---------------------------------------------------------------------------------------------
#define DPS_THREADS 128

struct DpsArgs {
  const int* sub_start; const int* sub_n; const i64* sub_pair_off; const i64* sub_triple_off;
  const int* cand_flat; const int* ins; const int* out; int lo; int max_n;
  u8* final_flag; int* err;
};

__global__ void __launch_bounds__(DPS_THREADS) k_dp_solve(DpsArgs A) {
  extern __shared__ int ssm[];
  const int p = blockIdx.x;
  const int n = A.sub_n[p], qs = A.sub_start[p], E = n - 1;
  int* G = ssm;                          // [n][n]
  int* cf = G + A.max_n * A.max_n;       // [n]
  short* arg = (short*)(cf + A.max_n);   // [n][n]
  __shared__ int best_v[DPS_THREADS / 32];
  __shared__ int best_e[DPS_THREADS / 32];
  const int* ins = A.ins + A.sub_pair_off[p];  // positive ambiguous counts; ins(i,j) = -ins[i*n+j]
  const int* out = A.out + A.sub_triple_off[p];
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += DPS_THREADS) cf[i] = A.cand_flat[qs + i];
  for (int e = tid; e < n * n; e += DPS_THREADS) { G[e] = FRS_NEG_INF; arg[e] = -1; }
  __syncthreads();
  for (int j = tid; j < E; j += DPS_THREADS) G[j * n + E] = -ins[j * n + E];
  __syncthreads();
  for (int j = E - 2; j >= 0; --j) {
    // all k in (j, E) are independent given rows k > j
    for (int k = j + 1 + tid; k < E; k += DPS_THREADS) {
      int best = FRS_NEG_INF, bk = -1;
      if (cf[k] - cf[j] >= 5) {
        const int base = -ins[j * n + k];
        const int* o = out + triple_mid_off(k, n) + j * (n - 1 - k);
        for (int k2 = k + 1; k2 <= E; ++k2) {
          if (cf[k2] - cf[k] < 5) continue;
          int ov = o[k2 - k - 1];
          if (ov < A.lo) continue;
          int g = G[k * n + k2];
          if (g == FRS_NEG_INF) continue;
          int d = base + ov + g;
          if (d > best) { best = d; bk = k2; }
        }
      }
      G[j * n + k] = best;
      arg[j * n + k] = (short)bk;
    }
    __syncthreads();
  }
  // top level: D(0,j,k) over 1 <= j < k <= E, first maximum in lexicographic order
  int my_best = FRS_NEG_INF, my_e = 0x7fffffff;
  for (int e = tid; e < n * n; e += DPS_THREADS) {
    int j = e / n, k = e - j * n;
    if (j < 1 || k <= j) continue;
    if (cf[j] - cf[0] < 5 || cf[k] - cf[j] < 5) continue;
    int ov = out[triple_mid_off(j, n) + 0 * (n - 1 - j) + (k - j - 1)];
    if (ov < A.lo) continue;
    int g = G[j * n + k];
    if (g == FRS_NEG_INF) continue;
    int d = -ins[0 * n + j] + ov + g;
    if (d > my_best || (d == my_best && e < my_e)) { my_best = d; my_e = e; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    int ov = __shfl_xor_sync(0xffffffffu, my_best, o);
    int oe = __shfl_xor_sync(0xffffffffu, my_e, o);
    if (ov > my_best || (ov == my_best && oe < my_e)) { my_best = ov; my_e = oe; }
  }
  if ((tid & 31) == 0) { best_v[tid >> 5] = my_best; best_e[tid >> 5] = my_e; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < DPS_THREADS / 32; ++w)
      if (best_v[w] > my_best || (best_v[w] == my_best && best_e[w] < my_e)) { my_best = best_v[w]; my_e = best_e[w]; }
    int none = -ins[0 * n + E];
    if (my_best != FRS_NEG_INF && my_best > none) {
      int j = my_e / n, k = my_e - j * n;
      A.final_flag[qs + j] = 1;
      A.final_flag[qs + k] = 1;
      int guard = 0;
      while (k != E) {
        int k2 = arg[j * n + k];
        if (k2 <= k || ++guard > n) { dev_fail(A.err, DEVERR_BACKTRACE, p); break; }
        A.final_flag[qs + k2] = 1;
        j = k;
        k = k2;
      }
    }
  }
}
