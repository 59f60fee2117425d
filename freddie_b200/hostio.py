"""Native host I/O: SPLIT parser and SEGMENT formatter of ``libfreddie_b200.so`` (csrc/host_io.cpp).

Replaces the reference's regex parsing (read_split / read_sequence, freddie_segment.py:121-185) and
its row formatting (:715-731) with multi-threaded C++ that reads the SPLIT files straight into the
packed batch and writes the SEGMENT files straight from the result arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
from typing import List, Sequence, Tuple

from . import _lib
from .engine import Engine, SegmentParams


def available() -> bool:
    return hasattr(_lib.load(), "frs_parse_tints")


def _paths(split_dir: str, outdir: str, chunk: Sequence[Tuple[str, int]]):
    sp = [("{}/{}/split_{}_{}.tsv".format(split_dir, c, c, t)).encode() for c, t in chunk]
    rp = [("{}/{}/reads_{}_{}.tsv".format(split_dir, c, c, t)).encode() for c, t in chunk]
    op = [("{}/{}/segment_{}_{}.tsv".format(outdir, c, c, t)).encode() for c, t in chunk]
    lp = [("{}/{}/segment_{}_{}.log".format(outdir, c, c, t)).encode() for c, t in chunk]
    return sp, rp, op, lp


class ParsedBatch:
    """A batch parsed by the native parser; owns the C++ object."""

    def __init__(self, split_paths: List[bytes], reads_paths: List[bytes], threads: int, packed: bytes = None):
        """Parses the tints' SPLIT files -- or, with ``packed``, loads a batch written by ``write_packed``."""
        self.lib = _lib.load()
        self.handle = C.c_void_p()
        err = C.create_string_buffer(1024)
        if packed is not None:
            rc = self.lib.frs_packed_read(packed, C.byref(self.handle), err, len(err))
        else:
            n = len(split_paths)
            a = (C.c_char_p * n)(*split_paths)
            b = (C.c_char_p * n)(*reads_paths)
            rc = self.lib.frs_parse_tints(a, b, n, threads, C.byref(self.handle), err, len(err))
        if rc != 0:
            msg = err.value.decode(errors="replace")
            if msg.startswith("AssertionError"):
                raise AssertionError(msg)
            raise _lib.FrsError(rc, msg)
        self.struct = _lib.FrsBatch()
        self.lib.frs_parsed_batch(self.handle, C.byref(self.struct))
        # a read's target intervals are its rep's: the library derives them on the device (no copy)
        self.lean = _lib.FrsBatch()
        C.memmove(C.byref(self.lean), C.byref(self.struct), C.sizeof(_lib.FrsBatch))
        self.lean.riv_ts = None
        self.lean.riv_te = None
        self.n_tints = self.struct.n_tints
        self.n_reads = self.struct.n_reads

    def as_struct(self):
        return self.lean

    @classmethod
    def from_packed(cls, path: str) -> "ParsedBatch":
        return cls([], [], 1, packed=path.encode())

    def write_packed(self, path: str):
        """The batch as one binary file (``frs_packed_write``): the side-channel that skips the TSV round trip."""
        err = C.create_string_buffer(1024)
        rc = self.lib.frs_packed_write(self.handle, path.encode(), err, len(err))
        if rc != 0:
            raise _lib.FrsError(rc, err.value.decode(errors="replace"))

    def format(self, res, out_paths: List[bytes], log_paths: List[bytes], threads: int):
        n = len(out_paths)
        err = C.create_string_buffer(1024)
        r = res.as_struct()
        rc = self.lib.frs_format_tints(self.handle, C.byref(r), (C.c_char_p * n)(*out_paths),
                                       (C.c_char_p * n)(*log_paths), threads, err, len(err))
        if rc != 0:
            raise _lib.FrsError(rc, err.value.decode(errors="replace"))

    def write_packed_segment(self, res, path: str):
        """The results of the batch as one binary file (``frs_packed_write_segment``, "FRSSEGM1");
        ``freddie_b200.packed.PackedSegment`` reads it back."""
        err = C.create_string_buffer(1024)
        r = res.as_struct()
        rc = self.lib.frs_packed_write_segment(self.handle, C.byref(r), path.encode(), err, len(err))
        if rc != 0:
            raise _lib.FrsError(rc, err.value.decode(errors="replace"))

    def close(self):
        if self.handle:
            self.lib.frs_parsed_free(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def parse_batch_native(split_dir: str, outdir: str, chunk: Sequence[Tuple[str, int]], threads: int,
                       packed_file: str = None):
    """Host half of a batch that needs no GPU: parse the tints of ``chunk`` (or load ``packed_file``, a batch of
    exactly those tints in that order).  Returns (ParsedBatch, output paths, log paths, seconds)."""
    t0 = time.perf_counter()
    sp, rp, op, lp = _paths(split_dir, outdir, chunk)
    pb = ParsedBatch.from_packed(packed_file) if packed_file else ParsedBatch(sp, rp, threads)
    if pb.n_tints != len(chunk):
        pb.close()
        raise _lib.FrsError(-2, "%s holds %d tints, its index lists %d" % (packed_file, pb.n_tints, len(chunk)))
    return pb, op, lp, time.perf_counter() - t0


def run_parsed_native(eng: Engine, prm: SegmentParams, parsed, chunk, threads: int, packed_segment: str = None):
    """GPU half + formatter of a batch parsed by ``parse_batch_native``; frees the parsed batch."""
    prof = os.environ.get("FRS_CLI_PROFILE")
    pb, op, lp, t_parse = parsed
    t1 = time.perf_counter()
    try:
        res = eng.segment_batch(pb, prm)
        t2 = time.perf_counter()
        pb.format(res, op, lp, threads)
        if packed_segment:
            pb.write_packed_segment(res, packed_segment)
        t3 = time.perf_counter()
        if prof:
            sys.stderr.write("[frs cli profile] batch of %d tints / %d reads: parse %.3f s  upload+kernels+download %.3f s  "
                             "format+write %.3f s\n" % (len(chunk), pb.n_reads, t_parse, t2 - t1, t3 - t2))
        return pb.n_reads, int(res.sizes["dp_cells"])
    finally:
        pb.close()


def run_batch_native(eng: Engine, prm: SegmentParams, split_dir: str, outdir: str,
                     chunk: Sequence[Tuple[str, int]], threads: int, packed_file: str = None,
                     packed_segment: str = None):
    """One batch, files to files: parse (or load ``packed_file``), segment on ``eng``'s GPU, format."""
    parsed = parse_batch_native(split_dir, outdir, chunk, threads, packed_file)
    return run_parsed_native(eng, prm, parsed, chunk, threads, packed_segment)
