"""ctypes binding of ``libfreddie_b200.so`` (C ABI in ``include/freddie_b200.h``).

There is no CPU fallback: if the shared library is missing, cannot be loaded, or no CUDA device is
present, the product path raises.  Build with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C freddie_b200/csrc``.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FRS_LIB") or os.path.join(_HERE, "libfreddie_b200.so")  # FRS_LIB: development builds

FRS_MAX_STAGES = 32

# taps (frs_get_intermediate)
TAP_Y_RAW, TAP_Y, TAP_THR, TAP_CAND, TAP_FIXED, TAP_DP_FINAL, TAP_SUB_START, TAP_SUB_N = 1, 2, 3, 4, 5, 6, 7, 8
TAP_COVERAGE, TAP_DP_TABLES, TAP_COV_OFF, TAP_SUB_TAB_OFF, TAP_FINAL_FLAGS = 9, 10, 12, 13, 14
OPT_SLAB_WORDS, OPT_KEEP_DP_TABLES, OPT_POLY_LONG_CLASS, OPT_LAZY_SEQ = 1, 2, 3, 4
STAT_NAMES = ["h2d_upload", "h2d_run", "d2h_run", "clip_words", "seq_words", "poly_tasks", "poly_long_tasks", "reruns", "h2d_copies"]

_p = C.c_void_p


class FrsParams(C.Structure):
    _fields_ = [
        ("sigma", C.c_double), ("tp", C.c_double), ("vf", C.c_double),
        ("mps", C.c_int32), ("lo", C.c_int32), ("ignore_ends", C.c_int32), ("thr_table_len", C.c_int32),
        ("thr_table", _p), ("gauss_w", _p), ("refine_w", _p),
        ("gauss_radius", C.c_int32), ("refine_radius", C.c_int32),
    ]


BATCH_COUNTS = ["n_tints", "n_islands", "n_reps", "n_rep_ivs", "n_reads", "n_read_ivs", "n_cigar_ops", "n_samples"]
BATCH_ARRAYS = [
    "tint_island_off", "tint_rep_off", "tint_read_off", "island_start", "island_sample_off",
    "rep_iv_off", "rep_weight", "rep_iv_fs", "rep_iv_fe",
    "read_rep", "read_strand", "read_len", "read_iv_off", "read_seq_off",
    "riv_ts", "riv_te", "riv_qs", "riv_qe", "riv_cig_off", "cigar", "seq_is_a", "seq_is_t",
]


class FrsBatch(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in BATCH_COUNTS] + [("n_seq_words", C.c_int64)]
                + [(n, _p) for n in BATCH_ARRAYS]
                + [("seq_edge_words", C.c_int32), ("host_arena", C.c_int32), ("seq_edge", _p),
                   ("cigar16", _p), ("riv_cig_n", _p), ("qe_from_cigar", C.c_int32), ("reserved1", C.c_int32)])


class FrsResultSizes(C.Structure):
    _fields_ = [
        ("n_final", C.c_int64), ("n_digit_bytes", C.c_int64), ("n_gap_records", C.c_int64),
        ("n_candidates", C.c_int64), ("n_subproblems", C.c_int64), ("dp_cells", C.c_int64),
        ("dp_read_cells", C.c_int64), ("max_subproblem", C.c_int32), ("pad", C.c_int32),
    ]


RESULT_ARRAYS = ["tint_final_off", "final_pos", "tint_digit_off", "digits", "read_head", "read_gap_off", "gap_rec"]


class FrsResult(C.Structure):
    _fields_ = [(n, _p) for n in RESULT_ARRAYS]


CLUSTER_BATCH_ARRAYS = ["tint_read_off", "tint_seg_n", "tint_digit_off", "read_row", "digits", "read_head", "read_gap_off", "gap_rec"]
CLUSTER_RESULT_ARRAYS = [
    ("tint_rep_off", "i4"), ("read_rep", "i4"), ("rep_first_read", "i4"), ("rep_count", "i4"), ("rep_fl", "i4"), ("rep_cat", "u1"),
    ("rep_gap", "i4"), ("tint_row_off", "i8"), ("I", "u1"), ("C", "u1"), ("tint_struct_off", "i4"), ("rep_struct", "i4"),
    ("tint_part_off", "i4"), ("part_rid_off", "i4"), ("part_rids", "i4"), ("part_inc_off", "i8"), ("inc", "i4"), ("tint_edges", "i8"),
]


class FrsClusterBatch(C.Structure):
    _fields_ = [("n_tints", C.c_int32), ("n_reads", C.c_int32)] + [(n, _p) for n in CLUSTER_BATCH_ARRAYS]


class FrsClusterSizes(C.Structure):
    _fields_ = [("n_reps", C.c_int64), ("n_structs", C.c_int64), ("n_parts", C.c_int64), ("n_incomp", C.c_int64),
                ("n_row_bytes", C.c_int64), ("edges_before", C.c_int64), ("edges_after", C.c_int64),
                ("prune_rounds", C.c_int32), ("launches", C.c_int32)]


class FrsClusterResult(C.Structure):
    _fields_ = [(n, _p) for n, _ in CLUSTER_RESULT_ARRAYS]


class FrsSplitBatch(C.Structure):
    _fields_ = [("n_groups", C.c_int32), ("n_reads", C.c_int32), ("group_read_off", _p), ("read_iv_off", _p), ("iv_s", _p),
                ("iv_e", _p), ("max_intervals", C.c_int32), ("max_reads", C.c_int32)]


class FrsSplitSizes(C.Structure):
    _fields_ = [("n_tints", C.c_int64), ("n_tint_ivs", C.c_int64), ("n_tint_rids", C.c_int64), ("n_simple", C.c_int64),
                ("n_big", C.c_int32), ("launches", C.c_int32)]


SPLIT_RESULT_ARRAYS = ["group_tint_off", "tint_iv_off", "tint_iv_s", "tint_iv_e", "tint_rid_off", "tint_rids"]


class FrsSplitResult(C.Structure):
    _fields_ = [(n, _p) for n in SPLIT_RESULT_ARRAYS]


class FrsError(RuntimeError):
    """Raised for every non-zero status of the library.  ``code`` is the FRS_ERR_* value."""

    def __init__(self, code, msg):
        super().__init__("[libfreddie_b200 %d] %s" % (code, msg))
        self.code = code
        self.msg = msg


_lib = None


def load():
    """Loads the library (once) and declares every prototype of the header."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FrsError(-100, "%s not found: the CUDA extension is not built (run __graft_entry__.build()); "
                             "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.frs_abi_version.restype = C.c_int
    lib.frs_device_count.restype = C.c_int
    lib.frs_mem_info.argtypes = [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
    lib.frs_mem_info.restype = C.c_int
    lib.frs_create.argtypes = [C.c_int, C.POINTER(_p)]
    lib.frs_destroy.argtypes = [_p]
    lib.frs_destroy.restype = None
    lib.frs_last_error.argtypes = [_p]
    lib.frs_last_error.restype = C.c_char_p
    lib.frs_stream.argtypes = [_p]
    lib.frs_stream.restype = _p
    lib.frs_upload.argtypes = [_p, C.POINTER(FrsBatch)]
    lib.frs_run.argtypes = [_p, C.POINTER(FrsParams), C.POINTER(FrsResultSizes)]
    lib.frs_download.argtypes = [_p, C.POINTER(FrsResult)]
    lib.frs_segment_batch.argtypes = [_p, C.POINTER(FrsBatch), C.POINTER(FrsParams), C.POINTER(FrsResultSizes)]
    lib.frs_submit.argtypes = [_p, C.POINTER(FrsBatch), C.POINTER(FrsParams), C.POINTER(C.c_int)]
    lib.frs_wait.argtypes = [_p, C.c_int, C.POINTER(FrsResultSizes)]
    lib.frs_fetch.argtypes = [_p, C.c_int, C.POINTER(FrsResult)]
    lib.frs_fetch_start.argtypes = [_p, C.c_int, C.POINTER(FrsResult)]
    lib.frs_fetch_finish.argtypes = [_p, C.c_int]
    lib.frs_get_intermediate.argtypes = [_p, C.c_int, _p, C.c_size_t, C.POINTER(C.c_size_t)]
    lib.frs_set_profiling.argtypes = [_p, C.c_int]
    lib.frs_get_timings.argtypes = [_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.POINTER(C.c_int)]
    lib.frs_last_launch_count.argtypes = [_p]
    lib.frs_set_option.argtypes = [_p, C.c_int, C.c_longlong]
    lib.frs_get_stats.argtypes = [_p, C.POINTER(C.c_longlong), C.c_int]
    lib.frs_get_stats.restype = C.c_int
    for fn in ("frs_create", "frs_upload", "frs_run", "frs_download", "frs_segment_batch", "frs_get_intermediate",
               "frs_submit", "frs_wait", "frs_fetch", "frs_fetch_start", "frs_fetch_finish",
               "frs_set_profiling", "frs_get_timings", "frs_last_launch_count", "frs_set_option"):
        getattr(lib, fn).restype = C.c_int
    if hasattr(lib, "frs_parse_tints"):
        lib.frs_parse_tints.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_int, C.c_int,
                                        C.POINTER(_p), C.c_char_p, C.c_size_t]
        lib.frs_parse_tints.restype = C.c_int
        lib.frs_parsed_batch.argtypes = [_p, C.POINTER(FrsBatch)]
        lib.frs_parsed_batch.restype = C.c_int
        lib.frs_parsed_free.argtypes = [_p]
        lib.frs_parsed_free.restype = None
        lib.frs_format_tints.argtypes = [_p, C.POINTER(FrsResult), C.POINTER(C.c_char_p), C.POINTER(C.c_char_p),
                                         C.c_int, C.c_char_p, C.c_size_t]
        lib.frs_format_tints.restype = C.c_int
        lib.frs_packed_write.argtypes = [_p, C.c_char_p, C.c_char_p, C.c_size_t]
        lib.frs_packed_write.restype = C.c_int
        lib.frs_packed_read.argtypes = [C.c_char_p, C.POINTER(_p), C.c_char_p, C.c_size_t]
        lib.frs_packed_read.restype = C.c_int
        lib.frs_packed_write_segment.argtypes = [_p, C.POINTER(FrsResult), C.c_char_p, C.c_char_p, C.c_size_t]
        lib.frs_packed_write_segment.restype = C.c_int
    lib.frs_cprep_create.argtypes = [C.c_int, C.POINTER(_p)]
    lib.frs_cprep_destroy.argtypes = [_p]
    lib.frs_cprep_destroy.restype = None
    lib.frs_cprep_last_error.argtypes = [_p]
    lib.frs_cprep_last_error.restype = C.c_char_p
    lib.frs_cprep_run.argtypes = [_p, C.POINTER(FrsClusterBatch), C.c_int, C.POINTER(FrsClusterSizes)]
    lib.frs_cprep_fetch.argtypes = [_p, C.POINTER(FrsClusterResult)]
    lib.frs_cprep_timings.argtypes = [_p, C.POINTER(C.c_float), C.c_int]
    for fn in ("frs_cprep_create", "frs_cprep_run", "frs_cprep_fetch", "frs_cprep_timings"):
        getattr(lib, fn).restype = C.c_int
    lib.frs_split_create.argtypes = [C.c_int, C.POINTER(_p)]
    lib.frs_split_destroy.argtypes = [_p]
    lib.frs_split_destroy.restype = None
    lib.frs_split_last_error.argtypes = [_p]
    lib.frs_split_last_error.restype = C.c_char_p
    lib.frs_split_run.argtypes = [_p, C.POINTER(FrsSplitBatch), C.POINTER(FrsSplitSizes)]
    lib.frs_split_fetch.argtypes = [_p, C.POINTER(FrsSplitResult)]
    lib.frs_split_last_ms.argtypes = [_p]
    lib.frs_split_last_ms.restype = C.c_float
    for fn in ("frs_split_create", "frs_split_run", "frs_split_fetch"):
        getattr(lib, fn).restype = C.c_int
    if lib.frs_abi_version() != 2:
        raise FrsError(-101, "ABI version mismatch")
    _lib = lib
    return lib


EXPORTED = [
    "frs_abi_version", "frs_device_count", "frs_mem_info", "frs_create", "frs_destroy", "frs_last_error", "frs_stream",
    "frs_upload", "frs_run", "frs_download", "frs_segment_batch", "frs_submit", "frs_wait", "frs_fetch", "frs_fetch_start",
    "frs_fetch_finish",
    "frs_get_intermediate", "frs_set_profiling",
    "frs_get_timings", "frs_last_launch_count", "frs_set_option", "frs_get_stats", "frs_parse_tints", "frs_parsed_batch", "frs_parsed_free",
    "frs_format_tints", "frs_packed_write", "frs_packed_read", "frs_packed_write_segment",
    "frs_split_create", "frs_split_destroy", "frs_split_last_error", "frs_split_run", "frs_split_fetch", "frs_split_last_ms",
    "frs_cprep_create", "frs_cprep_destroy", "frs_cprep_last_error", "frs_cprep_run", "frs_cprep_fetch", "frs_cprep_timings",
]
