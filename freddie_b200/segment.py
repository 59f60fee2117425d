#!/usr/bin/env python3
"""Drop-in replacement of Freddie's segment stage (``py/freddie_segment.py``) on B200 GPUs.

Same CLI, same SPLIT-in / SEGMENT-out formats, same in-process seam::

    segment(tint, sigma, smoothed_threshold, threshold_rate, variance_factor,
            max_problem_size, min_read_support_outside, ignore_ends)          (reference :738-747)

but every step of the hot path runs in the CUDA kernels of ``libfreddie_b200.so`` for a whole batch of
tints at a time.  Tints are bin-packed by estimated cost across the visible GPUs (two host threads with
one library context each per GPU, no collective), batches are bounded by a read count and by an estimate
of their device footprint against the GPU's free memory, and the host side only parses, packs and formats.  New flags are additive (``--gpus``, ``--batch-reads``); ``-t``
keeps its meaning of host worker threads (parsing / formatting).
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import os
import re
import sys
import threading
import time
from math import ceil
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import engine as _engine
from .engine import Engine, SegmentParams, smooth_threshold, apply_result, format_tint  # noqa: F401
from .pack import pack_tints
from . import schedule


# ------------------------------------------------------------------------------------------------
# CLI (reference parse_args, :53-110)
# ------------------------------------------------------------------------------------------------
def str_to_bool(value):
    if isinstance(value, bool):
        return value
    v = value.lower()
    if v in {"false", "f", "0", "no", "n"}:
        return False
    if v in {"true", "t", "1", "yes", "y"}:
        return True
    raise ValueError(f"{value} is not a valid boolean value")


def parse_args(argv=None):
    parser = argparse.ArgumentParser(description="Cluster aligned reads into isoforms")
    parser.add_argument("-s", "--split-dir", type=str, required=True,
                        help="Path to Freddie split directory of the reads")
    parser.add_argument("--consider-ends", type=str_to_bool, nargs="?", const=True, default=False,
                        help="Consider the start and end splice sites in segmentation")
    parser.add_argument("-o", "--outdir", type=str, default="freddie_segment/",
                        help="Path to output directory. Default: freddie_segment/")
    parser.add_argument("-t", "--threads", type=int, default=1,
                        help="Number of host threads (parsing / formatting). Default: 1")
    parser.add_argument("-sd", "--sigma", type=float, default=5.0, help="Sigma value for gaussian_filter1d")
    parser.add_argument("-tp", "--threshold-rate", type=float, default=0.90,
                        help="Threshold rate above which the read will be considered as covering a segment. "
                             "Low threshold is 1-threshold_rate. Default: 0.9.")
    parser.add_argument("-vf", "--variance-factor", type=float, default=3.0,
                        help="The stdev factor to fix a candidate peak. Default 3.0")
    parser.add_argument("-mps", "--max-problem-size", type=int, default=50,
                        help="Maximum number of candidate breakpoints allowed per segmentation problem")
    parser.add_argument("-lo", "--min-read-support-outside", type=int, default=3,
                        help="Minimum reads support for splice site to support a breakpoint")
    # additive flags of the B200 implementation
    parser.add_argument("--gpus", type=int, default=0, help="GPUs to use (0 = all visible)")
    parser.add_argument("--batch-reads", type=int, default=131072, help="Upper bound of reads per GPU batch")
    parser.add_argument("--packed-segment", action="store_true",
                        help="Also write every batch's results as a binary file under <outdir>/packed_segment/ "
                             "(read back with freddie_b200.packed.PackedSegment); the TSV files are written as always")
    args = parser.parse_args(argv)
    assert 1 >= args.threshold_rate >= 0.5
    assert 10 > args.variance_factor > 0
    assert 50 >= args.sigma > 0
    assert args.max_problem_size > 3
    assert args.min_read_support_outside >= 0
    assert args.threads > 0
    return args


# ------------------------------------------------------------------------------------------------
# SPLIT parsing in Python (reference read_split :121-171 / read_sequence :174-185).  The CLI prefers
# the native parser (csrc/host_io.cpp); this one serves the in-process seam and small runs.
# ------------------------------------------------------------------------------------------------
_CHR = r"[0-9A-Za-z!#$%&+./:;?@^_|~-][0-9A-Za-z!#$%&*+./:;=?@^_|~-]*"
_IV = r"[0-9]+-[0-9]+"
_RIV = _IV + ":" + _IV + r":(?:[0-9]+[MIDNSHPX=])+"
_TINT_RE = re.compile(r"#(%s)\t([0-9]+)\t(%s(?:,%s)*)\t([0-9]+)\n$" % (_CHR, _IV, _IV))
_READ_RE = re.compile(r"([0-9]+)\t([!-?A-~]{1,254})\t(%s)\t([+-])\t([0-9]+)\t(%s(?:\t%s)*)\n$" % (_CHR, _RIV, _RIV))
_RIV_RE = re.compile(r"([0-9]+)-([0-9]+):([0-9]+)-([0-9]+):((?:[0-9]+[MIDNSHPX=])+)")
_OP_RE = re.compile(r"([0-9]+)([MIDNSHPX=])")


def read_split(split_tsv: str) -> List[dict]:
    tints: Dict[int, dict] = {}
    with open(split_tsv) as fh:
        for line in fh:
            if line[0] == "#":
                m = _TINT_RE.match(line)
                if m is None:
                    raise AttributeError("malformed tint header: %r" % line[:80])
                tid = int(m.group(2))
                islands = [(int(a), int(b)) for a, b in (x.split("-") for x in m.group(3).split(","))]
                assert tid not in tints, "Transcriptional interval with id {} is repeated!".format(tid)
                assert all(a[1] < b[0] for a, b in zip(islands[:-1], islands[1:])), islands
                assert all(s < e for s, e in islands)
                tints[tid] = dict(id=tid, chr=m.group(1), intervals=islands, read_count=int(m.group(4)),
                                  reads=[], read_reps=None)
            else:
                m = _READ_RE.match(line)
                if m is None:
                    raise AttributeError("malformed read line: %r" % line[:80])
                ivs = [(int(a), int(b), int(c), int(d), [(int(n), op) for n, op in _OP_RE.findall(cg)])
                       for a, b, c, d, cg in _RIV_RE.findall(m.group(6))]
                assert all(x[1] <= y[0] and x[3] <= y[2] for x, y in zip(ivs[:-1], ivs[1:]))
                assert all(x[0] < x[1] and x[2] < x[3] for x in ivs)
                read = dict(id=int(m.group(1)), name=m.group(2), chr=m.group(3), strand=m.group(4),
                            tint=int(m.group(5)), intervals=ivs)
                tints[read["tint"]]["reads"].append(read)
    for tint in tints.values():
        assert len(tint["reads"]) == tint["read_count"]
    return list(tints.values())


def read_sequence(tint: dict, reads_tsv: str) -> None:
    rid_to_seq = {}
    with open(reads_tsv) as fh:
        for line in fh:
            cols = line.rstrip().split("\t")
            rid_to_seq[int(cols[0])] = cols[3]
    assert len(rid_to_seq) == len(tint["reads"]), tint["id"]
    for read in tint["reads"]:
        read["seq"] = rid_to_seq[read["id"]]
        read["length"] = len(read["seq"])


# ------------------------------------------------------------------------------------------------
# engines (one per GPU, created lazily)
# ------------------------------------------------------------------------------------------------
_ENGINES: Dict[tuple, Engine] = {}
_ENGINES_LOCK = threading.Lock()


def get_engine(device: int = 0, lane: int = 0) -> Engine:
    """One library context per (GPU, lane); a lane is one host thread of the directory driver."""
    with _ENGINES_LOCK:
        e = _ENGINES.get((device, lane))
        if e is None:
            e = Engine(device)
            _ENGINES[(device, lane)] = e
        return e


def segment_batch(tints: Sequence[dict], sigma, smoothed_threshold, threshold_rate, variance_factor,
                  max_problem_size, min_read_support_outside, ignore_ends, device: int = 0) -> List[int]:
    """``segment`` for many tints in one GPU batch; mutates every tint like the reference does."""
    prm = SegmentParams(sigma, threshold_rate, variance_factor, max_problem_size, min_read_support_outside,
                        ignore_ends, smoothed_threshold)
    batch = pack_tints(tints)
    res = get_engine(device).segment_batch(batch, prm)
    apply_result(batch, res)
    return [t["id"] for t in tints]


def segment(tint, sigma, smoothed_threshold, threshold_rate, variance_factor, max_problem_size,
            min_read_support_outside, ignore_ends):
    """Same signature and side effects as the reference's ``segment`` (:738-844): sets
    ``tint['final_positions']``, ``tint['segs']``, ``read['data']``, ``read['gaps']``; returns the id."""
    return segment_batch([tint], sigma, smoothed_threshold, threshold_rate, variance_factor, max_problem_size,
                         min_read_support_outside, ignore_ends)[0]


def run_segment(segment_args):
    """Same argument tuple as the reference's ``run_segment`` (:681-735): one tint, files in/out."""
    (split_dir, outdir, contig, tint_id, sigma, smoothed_threshold, threshold_rate, variance_factor,
     max_problem_size, min_read_support_outside, ignore_ends) = segment_args
    log = open("{}/{}/segment_{}_{}.log".format(outdir, contig, contig, tint_id), "w+")
    tints = read_split("{}/{}/split_{}_{}.tsv".format(split_dir, contig, contig, tint_id))
    assert len(tints) == 1
    tint = tints[0]
    read_sequence(tint, "{}/{}/reads_{}_{}.tsv".format(split_dir, contig, contig, tint_id))
    prm = SegmentParams(sigma, threshold_rate, variance_factor, max_problem_size, min_read_support_outside,
                        ignore_ends, smoothed_threshold)
    batch = pack_tints([tint])
    res = get_engine(0).segment_batch(batch, prm)
    with open("{}/{}/segment_{}_{}.tsv".format(outdir, contig, contig, tint_id), "w+") as out:
        out.write(format_tint(batch, res, 0))
    log.close()
    return contig, tint_id


# ------------------------------------------------------------------------------------------------
# directory driver (reference main, :847-885)
# ------------------------------------------------------------------------------------------------
def list_tints(split_dir: str) -> List[Tuple[str, int]]:
    jobs = []
    for contig in os.listdir(split_dir):
        if not os.path.isdir("{}/{}".format(split_dir, contig)):
            continue
        for path in glob.iglob("{}/{}/split_*.tsv".format(split_dir, contig)):
            jobs.append((contig, int(path[:-4].split("/")[-1].split("_")[-1])))
    return jobs


def _load_tint_py(split_dir, contig, tint_id):
    tints = read_split("{}/{}/split_{}_{}.tsv".format(split_dir, contig, contig, tint_id))
    assert len(tints) == 1
    read_sequence(tints[0], "{}/{}/reads_{}_{}.tsv".format(split_dir, contig, contig, tint_id))
    return tints[0]


def run_directory(split_dir: str, outdir: str, prm: SegmentParams, threads: int = 1, gpus: int = 0,
                  batch_reads: int = 131072, native: Optional[bool] = None, progress: bool = True,
                  lanes: int = 2, packed_segment: bool = False) -> dict:
    """Segments every tint of a SPLIT directory; returns counters.  Output is independent of
    ``threads``, ``gpus``, ``lanes`` and batch composition.  Every GPU is fed by ``lanes`` host threads
    (one library context each) that take the GPU's batches in turn, so that parsing and formatting of
    one batch overlap the copies and kernels of another."""
    from . import hostio
    lib = _engine._lib.load()
    n_dev = lib.frs_device_count()
    if n_dev <= 0:
        raise _engine._lib.FrsError(-1, "no CUDA device visible; freddie_b200 has no CPU fallback")
    n_gpus = n_dev if gpus <= 0 else min(gpus, n_dev)
    from . import packed as _packed
    packed_batches = _packed.read_index(split_dir) if _packed.is_packed_dir(split_dir) else None
    if packed_batches is not None:  # the packed side-channel: batches exist already, no TSV parse
        native = True
        jobs = [ct for b in packed_batches for ct in b["tints"]]
    else:
        jobs = list_tints(split_dir)
    for contig in {c for c, _ in jobs} | {c for c in os.listdir(split_dir) if os.path.isdir(os.path.join(split_dir, c))}:
        os.makedirs("{}/{}".format(outdir, contig), exist_ok=True)
    if native is None:
        native = hostio.available()
    if packed_batches is not None:
        costs = [(schedule.estimate_cost(b["reads"]), float(b["reads"])) for b in packed_batches]
    else:
        costs = [schedule.estimate_cost_from_files(split_dir, c, t) for c, t in jobs]
    shards = schedule.lpt_partition(costs, n_gpus)
    # a batch may take about a third of the free device memory of its GPU (two lanes per GPU in flight).  The
    # query needs a CUDA context, so it is made by the first lane whose context exists (never on this thread,
    # before the lanes have started): until then only the read-count bound applies.
    mem_bound = {}

    def batch_bytes():
        return mem_bound.get("bytes")

    def learn_memory(dev):
        if "bytes" in mem_bound:
            return
        free_b, total_b = C.c_longlong(0), C.c_longlong(0)
        if lib.frs_mem_info(dev, C.byref(free_b), C.byref(total_b)) == 0 and free_b.value > 0:
            mem_bound["bytes"] = free_b.value / (1.5 * max(1, lanes))
    done = [0]
    total = len(jobs)
    step = max(1, ceil(total / 100)) if total else 1
    lock = threading.Lock()
    stats = dict(tints=total, reads=0, dp_cells=0, gpus=n_gpus)
    errors: List[BaseException] = []

    def tick(n_tints, n_reads, cells):
        with lock:
            for _ in range(n_tints):
                if progress and done[0] % step == 0:
                    print("[freddie_segment] Done with {}/{} tints ({:.1%})".format(done[0], total, done[0] / total))
                done[0] += 1
            stats["reads"] += n_reads
            stats["dp_cells"] += cells

    if packed_batches is not None:
        feeds = [schedule.BatchFeed([], [], batch_reads, ready=[(packed_batches[i]["tints"], packed_batches[i]["file"])
                                                                for i in shards[d]]) for d in range(n_gpus)]
    else:
        feeds = [schedule.BatchFeed([], [], batch_reads,
                                    ready=((chunk, None) for chunk in schedule.batches(
                                        [jobs[i] for i in shards[d]], [costs[i] for i in shards[d]], batch_reads,
                                        batch_bytes)))
                 for d in range(n_gpus)]

    seg_index: List[dict] = []
    if packed_segment:
        os.makedirs(os.path.join(outdir, "packed_segment"), exist_ok=True)

    def segment_file(chunk):
        if not packed_segment:
            return None
        with lock:
            name = "segment_batch_%05d.frsg" % len(seg_index)
            seg_index.append(dict(file=name, tints=[[c, t] for c, t in chunk]))
        return os.path.join(outdir, "packed_segment", name)

    def worker(dev, lane):
        try:
            if n_gpus > 1:  # this lane's threads, parse teams and pinned buffers next to its GPU
                from . import affinity
                affinity.bind_to_gpu(dev)
            t_eng = time.perf_counter()
            # the CUDA context of the lane is created on a helper thread while the lane parses its first batch
            # (context creation costs as much as parsing ~100 k reads and needs nothing from the host side)
            box = {}

            def make_engine():
                try:
                    box["eng"] = get_engine(dev, lane)
                except BaseException as e:  # noqa: BLE001
                    box["err"] = e

            maker = threading.Thread(target=make_engine)
            maker.start()
            eng = None
            for chunk, packed_file in feeds[dev]:
                parsed = hostio.parse_batch_native(split_dir, outdir, chunk, threads, packed_file) if native else None
                if eng is None:
                    maker.join()
                    if "err" in box:
                        raise box["err"]
                    eng = box["eng"]
                    learn_memory(dev)
                    if os.environ.get("FRS_CLI_PROFILE"):
                        sys.stderr.write("[frs cli profile] context of GPU %d lane %d ready after %.3f s\n"
                                         % (dev, lane, time.perf_counter() - t_eng))
                if native:
                    n_reads, cells = hostio.run_parsed_native(eng, prm, parsed, chunk, threads, segment_file(chunk))
                else:
                    tints = [_load_tint_py(split_dir, c, t) for c, t in chunk]
                    batch = pack_tints(tints)
                    res = eng.segment_batch(batch, prm)
                    for k, (c, t) in enumerate(chunk):
                        open("{}/{}/segment_{}_{}.log".format(outdir, c, c, t), "w+").close()
                        with open("{}/{}/segment_{}_{}.tsv".format(outdir, c, c, t), "w+") as fh:
                            fh.write(format_tint(batch, res, k))
                    n_reads, cells = batch.n_reads, int(res.sizes["dp_cells"])
                tick(len(chunk), n_reads, cells)
            maker.join()
        except BaseException as e:  # propagate like the reference: the whole run aborts
            errors.append(e)
            for f in feeds:
                f.stop()

    ths = [threading.Thread(target=worker, args=(d, k)) for d in range(n_gpus) for k in range(max(1, lanes))]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    if errors:
        raise errors[0]
    if packed_segment:
        import json
        with open(os.path.join(outdir, "packed_segment", "index.json"), "w") as fh:
            json.dump(dict(format="freddie-b200-packed-segment-1", batches=seg_index), fh)
    return stats


def main(argv=None):
    args = parse_args(argv)
    args.split_dir = args.split_dir.rstrip("/")
    prm = SegmentParams(args.sigma, args.threshold_rate, args.variance_factor, args.max_problem_size,
                        args.min_read_support_outside, not args.consider_ends)
    run_directory(args.split_dir, args.outdir, prm, threads=args.threads, gpus=args.gpus,
                  batch_reads=args.batch_reads, packed_segment=args.packed_segment)


if __name__ == "__main__":
    main()
