"""Tint scheduling: cost estimates, LPT bin-packing across GPUs, memory-bounded batches.

The reference's only parallelism is ``Pool.imap_unordered`` over tints (freddie_segment.py:871-876);
tints never interact, so they are sharded across GPUs with no exchange step.
"""
from __future__ import annotations

import os
import threading
from typing import Iterable, List, Optional, Sequence, Tuple


def estimate_reads_from_bytes(split_bytes: int) -> float:
    """A split row is ~60 B of fixed fields plus ~25 B per interval."""
    return max(1.0, split_bytes / 220.0)


def estimate_cost(n_reads: float) -> float:
    """Relative GPU cost of a tint, from the only quantity known before parsing: its read count.

    SURVEY.md 8e proposes a model in L, K and R; measured on B200 (``profiles/cost_model.py``, one tint per
    batch, 1 ... 50 k reads, ``profiles/r02_cost_model_*.txt``) the device time of a tint is close to LINEAR in
    its reads (~0.05 ms per 1 000 reads) on top of a small per-tint constant -- every stage but the DP streams
    over reads / samples, and the DP's work per read saturates once the candidate count does.  The first
    model of this file (reads x candidates, quadratic up to 40 k reads) ranked tints correctly (Spearman 0.97)
    but overweighted large tints by up to 58x, which starves the GPU that holds a giant tint of other work."""
    return float(n_reads) + 50.0


def estimate_device_bytes(n_reads: float) -> float:
    """Rough device footprint of a tint inside a batch: packed inputs and per-read results (~2 KB per read
    with resident sequence planes), the coverage matrix 4 x candidates x reps and the digit matrix
    reps x segments -- the two blocks that grow faster than linearly with the tint."""
    k = min(n_reads, 40000.0) / 40.0 + 20.0
    return n_reads * 2048.0 + 4.0 * k * n_reads + n_reads * (k + 64.0)


def estimate_cost_from_files(split_dir: str, contig: str, tint_id: int) -> Tuple[float, float]:
    """(cost, estimated reads) from the split file size alone -- no parsing."""
    p = "{}/{}/split_{}_{}.tsv".format(split_dir, contig, contig, tint_id)
    try:
        b = os.path.getsize(p)
    except OSError:
        b = 0
    n = estimate_reads_from_bytes(b)
    return estimate_cost(n), n


def lpt_partition(costs: Sequence[Tuple[float, float]], n_bins: int) -> List[List[int]]:
    """Longest-processing-time-first: items by decreasing cost, each to the least-loaded bin.
    Deterministic (ties broken by index)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i][0], i))
    load = [0.0] * n_bins
    bins: List[List[int]] = [[] for _ in range(n_bins)]
    for i in order:
        b = min(range(n_bins), key=lambda j: (load[j], j))
        bins[b].append(i)
        load[b] += costs[i][0]
    return bins


def batches(jobs: Sequence, costs: Sequence[Tuple[float, float]], batch_reads: int,
            batch_bytes: Optional[float] = None) -> Iterable[List]:
    """Consecutive jobs grouped so that the estimated reads of a batch stay under ``batch_reads`` and, with
    ``batch_bytes``, its estimated device footprint (``estimate_device_bytes``) under that many bytes
    (``batch_bytes`` may be a callable that returns the bound, or None while it is not known yet).  A
    single tint above either bound forms its own batch -- a tint is never split."""
    cur: List = []
    acc, mem = 0.0, 0.0
    for job, (_, n) in zip(jobs, costs):
        m = estimate_device_bytes(n)
        bb = batch_bytes() if callable(batch_bytes) else batch_bytes
        if cur and (acc + n > batch_reads or (bb is not None and mem + m > bb)):
            yield cur
            cur, acc, mem = [], 0.0, 0.0
        cur.append(job)
        acc += n
        mem += m
    if cur:
        yield cur


class BatchFeed:
    """The batches of one GPU, handed out to that GPU's lanes (host threads) one at a time: every batch
    goes to exactly one lane, in order, and ``stop()`` (an error in any lane) ends the feed for all."""

    def __init__(self, jobs: Sequence, costs: Sequence[Tuple[float, float]], batch_reads: int, ready: Optional[Iterable] = None):
        """``ready``: batches that already exist (a packed directory) instead of grouping ``jobs``."""
        self._it = iter(ready) if ready is not None else iter(batches(jobs, costs, batch_reads))
        self._lock = threading.Lock()
        self._stopped = False

    def stop(self) -> None:
        self._stopped = True

    def next(self) -> Optional[List]:
        with self._lock:
            if self._stopped:
                return None
            return next(self._it, None)

    def __iter__(self):
        while True:
            chunk = self.next()
            if chunk is None:
                return
            yield chunk
