/*
 * freddie_b200.h -- C ABI of the B200-native segment stage (libfreddie_b200.so)
 *
 * The reference (vpc-ccg/freddie, py/freddie_segment.py) has no FFI: its boundary is the stage's CLI
 * plus the in-process seam  segment(tint, sigma, smoothed_threshold, threshold_rate, variance_factor,
 * max_problem_size, min_read_support_outside, ignore_ends)  (freddie_segment.py:738-747) that
 * run_segment() (:681-735) calls once per tint.  This header is what a ctypes binding of that seam
 * binds to: the Python host packs a batch of tints into flat arrays (frs_batch), one call runs every
 * step of segment() for the whole batch on one GPU, and the results come back as flat arrays
 * (frs_result) from which the SEGMENT rows (:715-731) are formatted.  INTEGRATION.md shows the stub.
 *
 * Conventions: every function returns 0 on success, <0 on error (message via frs_last_error).  The
 * caller owns every host buffer; the library keeps no host pointer after a call returns.  One context
 * per GPU, used by one host thread at a time; all device work is ordered on the context's streams.
 * No torch types, no C++ types.  There is NO CPU fallback: without a CUDA device frs_create fails.
 */
#ifndef FREDDIE_B200_H
#define FREDDIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRS_ABI_VERSION 2

/* error codes */
#define FRS_OK 0
#define FRS_ERR_CUDA (-1)        /* CUDA runtime failure */
#define FRS_ERR_ARG (-2)         /* malformed batch / parameter (the reference would assert) */
#define FRS_ERR_ASSERT (-3)      /* a data-dependent assert of the reference fired on the device */
#define FRS_ERR_LIMIT (-4)       /* a documented size limit of this build was exceeded */
#define FRS_ERR_STATE (-5)       /* calls out of order */
#define FRS_ERR_IO (-6)          /* host file I/O (parser / formatter) */

typedef struct frs_context frs_context;

/* Flags of parse_args() (freddie_segment.py:53-110) plus the host-computed tables.
 * thr_table      = smooth_threshold(tp) (:277-286), thr_table_len entries.
 * gauss_w        = scipy's normalised kernel for truncate=4.0, 2*gauss_radius+1 doubles (:755).
 * refine_w       = same for truncate=1.0 (refine_segmentation, :260-261).
 * Both kernels are computed by numpy on the host so their bits match scipy's. */
typedef struct {
  double sigma;
  double tp;
  double vf;
  int32_t mps;
  int32_t lo;
  int32_t ignore_ends;
  int32_t thr_table_len;
  const double* thr_table;
  const double* gauss_w;
  const double* refine_w;
  int32_t gauss_radius;
  int32_t refine_radius;
} frs_params;

/* A packed batch of tints (what read_split + read_sequence build, :121-185, as CSR arrays).
 * "sample" = one position of an island, islands iterated s..e inclusive (:652-659); samples of all
 * islands of all tints are concatenated ("flat" index).  Interval ends (te) are flat indices of the
 * INCLUSIVE sample the reference uses (:205,:666-673). */
typedef struct {
  int32_t n_tints, n_islands, n_reps, n_rep_ivs, n_reads, n_read_ivs, n_cigar_ops, n_samples;
  int64_t n_seq_words;
  /* tints */
  const int32_t* tint_island_off; /* [n_tints+1] */
  const int32_t* tint_rep_off;    /* [n_tints+1] */
  const int32_t* tint_read_off;   /* [n_tints+1] */
  /* islands */
  const int32_t* island_start;      /* [n_islands] genomic s */
  const int32_t* island_sample_off; /* [n_islands+1] */
  /* read reps (:165-170), any order inside a tint */
  const int32_t* rep_iv_off; /* [n_reps+1] */
  const int32_t* rep_weight; /* [n_reps] number of reads of the rep */
  const int32_t* rep_iv_fs;  /* [n_rep_ivs] flat sample of ts */
  const int32_t* rep_iv_fe;  /* [n_rep_ivs] flat sample of te */
  /* reads, file order */
  const int32_t* read_rep;     /* [n_reads] batch-global rep id */
  const uint8_t* read_strand;  /* [n_reads] 0 '+', 1 '-' */
  const int32_t* read_len;     /* [n_reads] len(seq) */
  const int32_t* read_iv_off;  /* [n_reads+1] */
  const int64_t* read_seq_off; /* [n_reads+1] word offsets into the bit-planes */
  const int32_t* riv_ts;       /* [n_read_ivs] genomic target start; riv_ts / riv_te may be NULL: a read's target */
  const int32_t* riv_te;       /* [n_read_ivs] genomic target end     intervals are its rep's (:165-170), derived on the device */
  const int32_t* riv_qs;       /* [n_read_ivs] query start */
  const int32_t* riv_qe;       /* [n_read_ivs] query end */
  const int32_t* riv_cig_off;  /* [n_read_ivs+1] */
  const uint32_t* cigar;       /* [n_cigar_ops] (len << 4) | op, op: 0 M/X/=, 1 I, 2 D, 3 other */
  const uint32_t* seq_is_a;    /* [n_seq_words] bit b of word w of a read = (seq[32w+b]=='A') */
  const uint32_t* seq_is_t;    /* [n_seq_words] same for 'T' */
  /* optional EDGE STORE (may be NULL / 0): the first and the last seq_edge_words plane words of every read, dense:
   * seq_edge[((read*2 + side)*2 + plane)*E + w], side 0 = words [0, min(E, nw)) of the read's planes, side 1 =
   * words [max(0, nw-E), nw), both left-aligned and zero-padded, plane 0 = isA, 1 = isT, nw = words of the read.
   * The poly-A/T scans only read the soft clips, which sit at the two ends of a read: with pinned planes and an
   * edge store, frs_upload copies the store whole (one dense DMA) and only the few clips longer than 32*E bases
   * are fetched from the planes afterwards. */
  int32_t seq_edge_words;
  int32_t host_arena; /* 1: the arrays of this struct that are copied (all but the two sequence planes) lie in ONE pinned
                         allocation, in the order of the fields of this struct (compact forms in the place of the
                         arrays they stand for, seq_edge last), each starting at the next multiple of 256 bytes:
                         frs_upload then moves them with a few large copies.  0: no assumption. */
  const uint32_t* seq_edge;
  /* optional COMPACT encodings of three read-interval arrays (each may be NULL; a quarter of the batch's bytes on
   * the bus): expanded on the device into the arrays above, which may then be NULL themselves.
   *   cigar16    [n_cigar_ops]  the same (len << 4) | op values as `cigar` when every len < 4096
   *   riv_cig_n  [n_read_ivs]   CIGAR ops of each interval (< 256) instead of the offsets `riv_cig_off`
   *   riv_qe == NULL (with qe_from_cigar = 1): qe = qs + query bases the interval's CIGAR consumes (M/X/=/I);
   *              the packer sets it only after checking that equality on every interval */
  const uint16_t* cigar16;
  const uint8_t* riv_cig_n;
  int32_t qe_from_cigar;
  int32_t reserved1;
} frs_batch;

/* Sizes of the variable-length results of the batch that was just run. */
typedef struct {
  int64_t n_final;       /* total final positions */
  int64_t n_digit_bytes; /* sum over tints of n_reps(t) * n_segs(t) */
  int64_t n_gap_records; /* total "a-b:n" records */
  int64_t n_candidates;  /* statistics */
  int64_t n_subproblems;
  int64_t dp_cells;      /* sum over subproblems of C(n,3)  (BASELINE metric "DP cell updates") */
  int64_t dp_read_cells; /* sum of C(n,3) * n_reps(tint) */
  int32_t max_subproblem;
  int32_t pad;
} frs_result_sizes;

/* Caller-allocated result buffers (sizes from frs_result_sizes and the batch). */
typedef struct {
  int32_t* tint_final_off; /* [n_tints+1] */
  int32_t* final_pos;      /* [n_final] genomic positions = tint['final_positions'] (:806) */
  int64_t* tint_digit_off; /* [n_tints+1] byte offset of the tint's digit block */
  uint8_t* digits;         /* [n_digit_bytes] ASCII '0'/'1'/'2'; row r_local of tint t starts at
                              tint_digit_off[t] + r_local * (n_final(t)-1)   (:815-838) */
  int32_t* read_head;      /* [n_reads*8] fixed part of every read's gaps field, see FRS_HEAD_* */
  int32_t* read_gap_off;   /* [n_reads+1] */
  int32_t* gap_rec;        /* [n_gap_records*3] (l1, f2, size) -> "{l1}-{f2}:{size}" (:455-471) */
} frs_result;

/* read_head layout (8 x int32 per read) */
#define FRS_HEAD_FLAGS 0 /* bit0: read has a '1' digit (else the gaps field is empty, :372);
                            bits 8-9: start poly kind (0 none, 1 'A', 2 'T'); bits 16-17: end poly kind */
#define FRS_HEAD_S_LEN 1 /* "S{A|T}_{s_len}:{s_gap}" + "SSC:{ssc}"  or just "SSC:{ssc}" (:407-420) */
#define FRS_HEAD_S_GAP 2
#define FRS_HEAD_SSC 3
#define FRS_HEAD_E_LEN 4 /* "E{A|T}_{e_len}:{e_gap}" + "ESC:{esc}"  or just "ESC:{esc}" (:438-454) */
#define FRS_HEAD_E_GAP 5
#define FRS_HEAD_ESC 6
#define FRS_HEAD_RESERVED 7

/* ---- context ---- */
int frs_abi_version(void);
int frs_device_count(void);
/* free / total device memory in bytes (the directory driver bounds its batches with it) */
int frs_mem_info(int device, long long* free_bytes, long long* total_bytes);
int frs_create(int device, frs_context** out);
void frs_destroy(frs_context* ctx);
const char* frs_last_error(const frs_context* ctx); /* ctx may be NULL: last global error */
void* frs_stream(frs_context* ctx);                 /* cudaStream_t of the context */

/* ---- the hot path (replaces segment(), freddie_segment.py:738-844, for a batch of tints) ---- */
/* H2D copy of the batch (host arrays may be pinned or pageable); returns when the arrays have been read.
 * With FRS_OPT_LAZY_SEQ (default) pinned sequence bit-planes are not copied here, see the option. */
int frs_upload(frs_context* ctx, const frs_batch* batch);
/* Runs every kernel of the pipeline on the uploaded batch; may be called repeatedly. */
int frs_run(frs_context* ctx, const frs_params* prm, frs_result_sizes* sizes);
/* D2H copy of the results into caller buffers; synchronises the stream. */
int frs_download(frs_context* ctx, const frs_result* out);
/* upload + run + (caller allocates between) is the usual sequence; this convenience call does
 * upload+run and leaves the download to the caller once it has sized its buffers. */
int frs_segment_batch(frs_context* ctx, const frs_batch* batch, const frs_params* prm,
                      frs_result_sizes* sizes);

/* ---- the same hot path, pipelined: up to six batches in flight per context (copy in | kernels | tail | copy out), ONE host thread.
 * The reference overlaps tints with a process pool (imap_unordered, freddie_segment.py:871-876); here the
 * copy of batch k+1 and the read-back of batch k-1 overlap the kernels of batch k on the copy engines.
 *   frs_submit  enqueues the host-to-device copies and every kernel of the run and returns at once (no
 *               host round trip inside a run: all counts stay on the device).  The batch arrays (and with
 *               FRS_OPT_LAZY_SEQ the two sequence planes) must stay valid until frs_wait returns.
 *   frs_wait    blocks until the run is complete and returns the result sizes (if a data-dependent buffer
 *               was too small it is grown and the run repeated first: first batches of a context only).
 *   frs_fetch   copies the results into caller buffers, blocks until they have arrived, frees the ticket.
  * Tickets are slots: at most six may be outstanding, and they complete in submission order.  The tail of a
 * run (clip fetch from pinned host memory, poly-A/T scans) executes on its own stream beside the head of the
 * next batch. */
int frs_submit(frs_context* ctx, const frs_batch* batch, const frs_params* prm, int* ticket);
int frs_wait(frs_context* ctx, int ticket, frs_result_sizes* sizes);
int frs_fetch(frs_context* ctx, int ticket, const frs_result* out);
/* frs_fetch in two halves, for a host loop that never blocks on a copy: _start enqueues the device-to-host
 * copies (after frs_wait), _finish blocks until they have arrived and frees the ticket. */
int frs_fetch_start(frs_context* ctx, int ticket, const frs_result* out);
int frs_fetch_finish(frs_context* ctx, int ticket);

/* ---- debug taps for per-step parity tests (values of the LAST frs_run) ---- */
enum {
  FRS_TAP_Y_RAW = 1,      /* int32 [n_samples]   process_splicing_data (:648-678) */
  FRS_TAP_Y = 2,          /* f64   [n_samples]   gaussian_filter1d (:755) */
  FRS_TAP_THR = 3,        /* f64   [n_tints]     variance threshold (:757-759) */
  FRS_TAP_CAND = 4,       /* int32 [n_candidates] flat sample index of each candidate (:615-621) */
  FRS_TAP_FIXED = 5,      /* u8    [n_candidates] fixed flags after break_large_problems (:776-788) */
  FRS_TAP_DP_FINAL = 6,   /* u8    [n_candidates] fixed | chosen by the DP (:793-801) */
  FRS_TAP_SUB_START = 7,  /* int32 [n_subproblems] first candidate rank of each subproblem */
  FRS_TAP_SUB_N = 8,      /* int32 [n_subproblems] size n */
  FRS_TAP_COVERAGE = 9,   /* u32   cumulative coverage rows, see DESIGN.md */
  FRS_TAP_DP_TABLES = 10, /* int32 per-subproblem blocks: pair-indexed ambiguous counts (= -ins, :500-506),
                             n(n-1)/2 entries, then out (:509-528) as [j][k-j-1][i], C(n,3) entries.  Blocks
                             exist for subproblems of giant tints (summed over their rep slabs) and, with
                             FRS_OPT_KEEP_DP_TABLES, for every subproblem */
  FRS_TAP_COV_OFF = 12,   /* int64 [n_tints+1] element offset of each tint's coverage block */
  FRS_TAP_SUB_TAB_OFF = 13, /* int64 [n_subproblems] first element of each subproblem's table block (the
                               subproblem list and the blocks are in no particular order) */
  FRS_TAP_FINAL_FLAGS = 14, /* u8 [n_samples] 1 at every final position after refine_segmentation (:249-266,
                               :803-805): the DP-final candidates plus the positions refine added */
};
/* Copies min(cap_bytes, size) bytes of the tap to dst (host) and stores the full size in *bytes. */
int frs_get_intermediate(frs_context* ctx, int which, void* dst, size_t cap_bytes, size_t* bytes);

/* ---- tuning / test options ---- */
enum {
  FRS_OPT_SLAB_WORDS = 1,     /* read reps per DP CTA, in words of 32 (default 64): tints with more reps are
                                 processed by several CTAs per subproblem (multi-CTA mode) */
  FRS_OPT_KEEP_DP_TABLES = 2, /* also store the on-chip ins/out tables of small tints for FRS_TAP_DP_TABLES */
  FRS_OPT_POLY_LONG_CLASS = 3, /* length class (4 per octave: 36 = 512 bases, default) from which a poly-A/T
                                 clip scan is done by a whole warp instead of one thread; 1 = every scan */
  FRS_OPT_LAZY_SEQ = 4,       /* 1 (default): when the caller's seq_is_a / seq_is_t are pinned (or registered) host
                                 memory, frs_upload does NOT copy them; after segmentation a kernel fetches only
                                 the plane words of the soft clips straight from that memory (a few per cent of
                                 the bases).  The two arrays must then stay valid until the run has finished.
                                 Pageable planes are copied whole.
                                 0: frs_upload always copies both planes whole (inputs fully resident in HBM). */
};
int frs_set_option(frs_context* ctx, int key, long long value);

/* transfer statistics of the last frs_upload / frs_run; returns the number of statistics */
#define FRS_N_STATS 9
enum { FRS_STAT_H2D_UPLOAD = 0, FRS_STAT_H2D_RUN = 1 /* bytes the device fetched from the caller's pinned planes */,
       FRS_STAT_D2H_RUN = 2, FRS_STAT_CLIP_WORDS = 3,
       FRS_STAT_SEQ_WORDS = 4, FRS_STAT_POLY_TASKS = 5 /* poly-A/T scan tasks that survived the 5-stretch filter */,
       FRS_STAT_POLY_LONG_TASKS = 6 /* of which scanned by a whole warp */,
       FRS_STAT_RERUNS = 7 /* runs of this context repeated because a buffer capacity was missed */,
       FRS_STAT_H2D_COPIES = 8 /* host-to-device copies the last upload issued (adjacent arrays are merged) */ };
int frs_get_stats(frs_context* ctx, long long* out, int n);

/* ---- per-kernel device timing of the last frs_run (CUDA events on the context stream) ---- */
#define FRS_MAX_STAGES 32
int frs_set_profiling(frs_context* ctx, int enabled);
/* names[i] points to a static string; ms[i] is the summed duration of the stage's launches;
 * launches[i] the number of kernel launches in it.  Returns the number of stages (<= FRS_MAX_STAGES). */
int frs_get_timings(frs_context* ctx, const char** names, float* ms, int* launches);
/* total kernel launches issued by the last frs_run */
int frs_last_launch_count(frs_context* ctx);

/* ---- host side: native SPLIT parser and SEGMENT formatter (read_split/read_sequence :121-185,
 *      output :715-731).  See host_io.cpp. ---- */
typedef struct frs_parsed frs_parsed;
/* Parses the tints named by (split_paths[i], reads_paths[i]) with n_threads host threads into one
 * packed batch owned by the returned object. */
int frs_parse_tints(const char* const* split_paths, const char* const* reads_paths, int n, int n_threads,
                    frs_parsed** out, char* err, size_t err_cap);
/* Fills *batch with pointers into the parsed object (valid until frs_parsed_free). */
int frs_parsed_batch(const frs_parsed* p, frs_batch* batch);
void frs_parsed_free(frs_parsed* p);
/* Writes segment_<contig>_<id>.tsv (+ empty .log) for every tint of the parsed batch into
 * out_paths[i] (tsv) / log_paths[i]. */
int frs_format_tints(const frs_parsed* p, const frs_result* res, const char* const* out_paths,
                     const char* const* log_paths, int n_threads, char* err, size_t err_cap);

/* ---- packed side-channel: a parsed batch as one binary file ("FRSBATC1"), so that a
 *      pipeline can hand tints to this stage without the TSV round trip of split_*.tsv / reads_*.tsv
 *      (freddie_split.py:445-481 writes them, freddie_segment.py:121-185 reads them); text stays the
 *      default.  frs_packed_read returns the same object frs_parse_tints does. ---- */
int frs_packed_write(const frs_parsed* p, const char* path, char* err, size_t err_cap);
int frs_packed_read(const char* path, frs_parsed** out, char* err, size_t err_cap);
/* The SEGMENT twin ("FRSSEGM1"): the results of a batch as arrays, for a consumer that would otherwise
 * re-parse segment_*.tsv (freddie_cluster.py:119-172); freddie_b200/packed.py reads it back. */
int frs_packed_write_segment(const frs_parsed* p, const frs_result* res, const char* path, char* err, size_t err_cap);

/* =========================================================================================================
 * Next row of the scope table: the Gurobi-free front of freddie_cluster.py, on the arrays
 * the segment stage produces (the FRSSEGM1 arrays = frs_result + the batch's read tables).  Replaces, per tint:
 *   read_segment's read-rep merge          freddie_cluster.py:154-164  (key: digits with 2->0, internal gap
 *                                           sizes and poly-tail lengths, sizes <= 10 as 0, in the file's order)
 *   preprocess_ilp                         :277-328  (I, C, FL, poly-tail category, the added tail gap;
 *                                           garbage_cost of the `constant` model = 3 x reads of the rep)
 *   partition_reads                        :198-274  (structures (I row, FL, category) merged in first-seen
 *                                           order, the O(U^2 M) pair test, synchronous pruning rounds of the
 *                                           compatibility graph, connected components, even pieces of at most
 *                                           maximum_ilp_size structures, incompatible rep pairs per piece)
 * Everything quadratic runs in CUDA kernels on bit-packed rows (kernels_cluster.cuh); the list bookkeeping
 * between them (stable grouping, pieces) is native host code inside the library.  No CPU fallback.
 * ========================================================================================================= */
typedef struct frs_cprep frs_cprep;

typedef struct {
  int32_t n_tints, n_reads;
  const int32_t* tint_read_off;  /* [n_tints+1] reads of tint t */
  const int32_t* tint_seg_n;     /* [n_tints]   M = segments of the tint (final positions - 1), >= 1 */
  const int64_t* tint_digit_off; /* [n_tints+1] first byte of the tint's digit rows in `digits` (rows of M bytes) */
  const int32_t* read_row;       /* [n_reads]   digit row of the read inside its tint (FRSSEGM1: read_rep - tint_rep_off) */
  const uint8_t* digits;         /* ASCII '0' '1' '2', frs_result.digits */
  const int32_t* read_head;      /* [8*n_reads] frs_result.read_head */
  const int32_t* read_gap_off;   /* [n_reads+1] frs_result.read_gap_off */
  const int32_t* gap_rec;        /* [3*n_gaps]  frs_result.gap_rec (seg a, seg b, size), any order inside a read */
} frs_cluster_batch;

typedef struct {
  int64_t n_reps;       /* read reps over all tints (= entries of part_rids) */
  int64_t n_structs;    /* distinct structures */
  int64_t n_parts;      /* partitions */
  int64_t n_incomp;     /* incompatible rep pairs over all partitions */
  int64_t n_row_bytes;  /* sum over tints of reps x M (size of I and of C) */
  int64_t edges_before, edges_after; /* compatibility graph edges over all tints, before / after pruning */
  int32_t prune_rounds; /* pruning rounds executed (the last one removes nothing) */
  int32_t launches;     /* kernel launches of this run */
} frs_cluster_sizes;

/* caller-allocated host arrays (sizes from frs_cluster_sizes); any pointer may be NULL (skipped) */
typedef struct {
  int32_t* tint_rep_off;    /* [n_tints+1] */
  int32_t* read_rep;        /* [n_reads]   rep of the read, local to its tint, in first-seen order (read_segment) */
  int32_t* rep_first_read;  /* [n_reps]    first read of the rep (local index): the read preprocess_ilp looks at */
  int32_t* rep_count;       /* [n_reps]    reads of the rep (garbage_cost `constant` = 3 x this) */
  int32_t* rep_fl;          /* [2*n_reps]  FL (:311) */
  uint8_t* rep_cat;         /* [n_reps]    'N' 'S' 'E' (:297-308) */
  int32_t* rep_gap;         /* [3*n_reps]  for cat != 'N': key (a, b) and value of the gap preprocess_ilp adds (:302,:306) */
  int64_t* tint_row_off;    /* [n_tints+1] first byte of the tint's rows in I / C (rows of M bytes, rep order) */
  uint8_t* I;               /* [n_row_bytes] 0/1 (:289-291) */
  uint8_t* C;               /* [n_row_bytes] 0/1 (:312-316) */
  int32_t* tint_struct_off; /* [n_tints+1] */
  int32_t* rep_struct;      /* [n_reps]    structure of the rep, local to its tint, first-seen order (:207-218) */
  int32_t* tint_part_off;   /* [n_tints+1] */
  int32_t* part_rid_off;    /* [n_parts+1] into part_rids */
  int32_t* part_rids;       /* [n_reps]    rep ids (local to the tint) of every partition (:263-265) */
  int64_t* part_inc_off;    /* [n_parts+1] into inc (pairs) */
  int32_t* inc;             /* [2*n_incomp] (rid_1, rid_2) in the reference's order (:266-273) */
  int64_t* tint_edges;      /* [2*n_tints] edges before / after pruning per tint */
} frs_cluster_result;

int frs_cprep_create(int device, frs_cprep** out);
void frs_cprep_destroy(frs_cprep* c);
const char* frs_cprep_last_error(frs_cprep* c);
/* copies the batch in and runs every step; the results stay on the device / in the object until the next run */
int frs_cprep_run(frs_cprep* c, const frs_cluster_batch* batch, int maximum_ilp_size, frs_cluster_sizes* sizes);
int frs_cprep_fetch(frs_cprep* c, const frs_cluster_result* out);
/* device milliseconds of the last run: [0] dedupe + preprocess, [1] pair test, [2] pruning, [3] components,
 * [4] incompatible pairs; returns the number of entries */
int frs_cprep_timings(frs_cprep* c, float* ms, int n);

/* =========================================================================================================
 * Last row of the scope table: tint construction of freddie_split.py on decoded alignments.
 * Replaces, for a batch of read groups (what read_sam yields, freddie_split.py:207-244: reads of one contig whose
 * alignments chain into each other; BAM decoding itself stays with the caller):
 *   get_transcriptional_intervals  :295-364  (simple intervals = union of the alignment intervals, touching ones
 *                                   merged; groups joined through multi-interval reads, in the order of their
 *                                   smallest interval; fewer than 3 reads dropped)
 *   break_tint                     :246-293  (groups with >= max_intervals intervals or >= max_reads reads: intervals
 *                                   joined by junctions that >= 2 reads support; per component the reads that
 *                                   start an alignment in it and all intervals those reads start in)
 * Sorting (radix), sweeps (scans), union-find and the set unions run in CUDA kernels (kernels_split.cuh); the
 * final lists are assembled by native host code inside the library.  No CPU fallback.
 * ========================================================================================================= */
typedef struct frs_split frs_split;

typedef struct {
  int32_t n_groups, n_reads;
  const int32_t* group_read_off; /* [n_groups+1] reads of group g (read id inside a group = index - group_read_off[g]) */
  const int32_t* read_iv_off;    /* [n_reads+1]  alignment intervals of a read, in target order, at least one */
  const int32_t* iv_s;           /* [n_intervals] target start (>= 0) */
  const int32_t* iv_e;           /* [n_intervals] target end (exclusive, > start) */
  int32_t max_intervals;         /* 100 in the reference (:357); <= 0: that default */
  int32_t max_reads;             /* 1500 in the reference (:357); <= 0: that default */
} frs_split_batch;

typedef struct {
  int64_t n_tints, n_tint_ivs, n_tint_rids;
  int64_t n_simple;  /* simple intervals over all groups */
  int32_t n_big;     /* groups that went through break_tint */
  int32_t launches;  /* kernel launches of this run */
} frs_split_sizes;

/* caller-allocated host arrays (sizes from frs_split_sizes); tints of group g: group_tint_off[g] .. [g+1], in the
 * order get_transcriptional_intervals returns them */
typedef struct {
  int32_t* group_tint_off; /* [n_groups+1] */
  int32_t* tint_iv_off;    /* [n_tints+1] */
  int32_t* tint_iv_s;      /* [n_tint_ivs] tint['intervals'], sorted */
  int32_t* tint_iv_e;
  int32_t* tint_rid_off;   /* [n_tints+1] */
  int32_t* tint_rids;      /* [n_tint_rids] tint['rids'] (ids inside the group), sorted */
} frs_split_result;

int frs_split_create(int device, frs_split** out);
void frs_split_destroy(frs_split* c);
const char* frs_split_last_error(frs_split* c);
int frs_split_run(frs_split* c, const frs_split_batch* batch, frs_split_sizes* sizes);
int frs_split_fetch(frs_split* c, const frs_split_result* out);
/* device milliseconds of the last run (first kernel to last kernel) */
float frs_split_last_ms(frs_split* c);

#ifdef __cplusplus
}
#endif
#endif /* FREDDIE_B200_H */
