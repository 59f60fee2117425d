"""TEST INFRASTRUCTURE -- CPU restatement of the tint construction of freddie_split.py, the last row of the
hot-path scope table.  Nothing under ``freddie_b200/`` imports this file.

Follows ``/root/reference/py/freddie_split.py``:

* ``get_transcriptional_intervals`` (:295-364): the union of all alignment intervals of a group of reads cut into
  "simple" intervals (a new one starts where an interval begins strictly after everything seen so far: touching
  intervals merge), the simple intervals joined through the reads that have alignments in several of them (the BFS of
  :325-337 = connected components, discovered in the order of their smallest simple interval), groups of fewer
  than three reads dropped, and every group with >= 100 intervals or >= 1500 reads handed to ``break_tint``;
* ``break_tint`` (:246-293): inside a big group, intervals joined by splice junctions that at least two reads
  support; every component becomes a tint made of the reads that START an alignment in it and of every interval
  in which one of those reads starts an alignment (components with fewer than three such reads are dropped).

A group is what ``read_sam`` (:207-244) yields: reads as lists of (start, end) target intervals in order, read id =
position in the list.  Pinned: ``oracle/pin_split_tints.py`` runs the UNMODIFIED reference functions (pysam stubbed:
they never touch it) on seeded groups and stores a SHA-256 per group in ``tests/golden/split_tints.json``.
"""
from __future__ import annotations

import hashlib
import json
from typing import List, Sequence, Tuple

import numpy as np

Group = Sequence[Sequence[Tuple[int, int]]]


def _find(parent: np.ndarray, x: int) -> int:
    while parent[x] != x:
        parent[x] = parent[parent[x]]
        x = parent[x]
    return x


def _union(parent: np.ndarray, a: int, b: int) -> None:
    ra, rb = _find(parent, a), _find(parent, b)
    if ra != rb:
        parent[max(ra, rb)] = min(ra, rb)  # the root of a component is its smallest member


def simple_intervals(reads: Group):
    """(:296-321) sorted sweep over every (start, end): returns (starts, ends) of the simple intervals and, per read,
    the simple interval of each of its alignment intervals."""
    n_iv = sum(len(r) for r in reads)
    if n_iv == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), [[] for _ in reads]
    s = np.fromiter((iv[0] for r in reads for iv in r), dtype=np.int64, count=n_iv)
    e = np.fromiter((iv[1] for r in reads for iv in r), dtype=np.int64, count=n_iv)
    order = np.lexsort((e, s))
    ss, ee = s[order], e[order]
    run_end = np.maximum.accumulate(ee)
    new = np.ones(n_iv, dtype=bool)
    new[1:] = ss[1:] > run_end[:-1]
    sid_sorted = np.cumsum(new) - 1
    starts = ss[new]
    ends = np.maximum.reduceat(ee, np.flatnonzero(new))
    sid = np.empty(n_iv, dtype=np.int64)
    sid[order] = sid_sorted
    per_read, k = [], 0
    for r in reads:
        per_read.append(sid[k:k + len(r)].tolist())
        k += len(r)
    return starts, ends, per_read


def break_tint(intervals: List[Tuple[int, int]], rids: List[int], reads: Group):
    """(:246-293) on the sorted, disjoint ``intervals`` of a big group."""
    st = np.array([a for a, _ in intervals], dtype=np.int64)
    n = len(intervals)

    def at(pos: int) -> int:  # pos_to_intrv (:251-255): the interval that holds the position
        k = int(np.searchsorted(st, pos, side="right")) - 1
        assert 0 <= k < n and intervals[k][0] <= pos < intervals[k][1], (pos, k)
        return k

    starts_in = {rid: sorted({at(a) for a, _ in reads[rid]}) for rid in rids}
    weight = {}
    for rid in rids:
        alns = reads[rid]
        for (a1s, a1e), (a2s, a2e) in zip(alns[:-1], alns[1:]):
            v1, v2 = at(a1e - 1), at(a2s)
            assert v1 <= v2 < n
            weight[(v1, v2)] = weight.get((v1, v2), 0) + 1
    parent = np.arange(n)
    for (u, v), w in weight.items():
        if w >= 2:
            _union(parent, u, v)
    root = np.array([_find(parent, i) for i in range(n)])
    out = []
    for c in sorted(set(root.tolist())):  # components in the order of their smallest interval
        members = set(np.flatnonzero(root == c).tolist())
        c_rids = sorted(rid for rid in rids if members.intersection(starts_in[rid]))
        if len(c_rids) > 2:
            ivs = sorted({i for rid in c_rids for i in starts_in[rid]})
            out.append(([intervals[i] for i in ivs], c_rids))
    return out


def transcriptional_intervals(reads: Group, max_intervals: int = 100, max_reads: int = 1500):
    """``get_transcriptional_intervals`` (:295-364): [(intervals, rids)] in the reference's order."""
    starts, ends, per_read = simple_intervals(reads)
    n = len(starts)
    if n == 0:
        return []
    parent = np.arange(n)
    for ids in per_read:
        for a, b in zip(ids[:-1], ids[1:]):
            _union(parent, a, b)
    root = np.array([_find(parent, i) for i in range(n)])
    read_root = np.array([root[ids[0]] if ids else -1 for ids in per_read])
    out = []
    for c in sorted(set(root.tolist())):
        rids = np.flatnonzero(read_root == c).tolist()
        if len(rids) < 3:
            continue
        ivs = [(int(starts[i]), int(ends[i])) for i in np.flatnonzero(root == c)]
        if len(ivs) < max_intervals and len(rids) < max_reads:
            out.append((ivs, rids))
        else:
            out.extend(break_tint(ivs, rids, reads))
    return out


def canonical(tints) -> str:
    return json.dumps([[[[int(s), int(e)] for s, e in ivs], [int(r) for r in rids]] for ivs, rids in tints],
                      separators=(",", ":"))


def digest_of(tints) -> str:
    return hashlib.sha256(canonical(tints).encode()).hexdigest()


# ------------------------------------------------------------------------------------------------------------
# seeded groups (shared by the pin script and the tests)
# ------------------------------------------------------------------------------------------------------------
def make_group(seed: int, n_loci: int, reads_per_locus: int, exons: Tuple[int, int] = (3, 12), chain: float = 0.15,
               big: bool = False) -> List[List[Tuple[int, int]]]:
    """Reads of one group: ``n_loci`` genes next to each other (some reads bridge two genes with probability ``chain``),
    exon boundaries jittered per read, a few unspliced and a few touching intervals.  ``big``: one locus with more
    than 100 exons so that ``break_tint`` runs, with junctions seen once, twice and many times."""
    rng = np.random.default_rng(seed)
    loci, pos = [], 1000
    for _ in range(n_loci):
        ne = int(rng.integers(110, 140)) if big else int(rng.integers(exons[0], exons[1] + 1))
        ex = []
        for _ in range(ne):
            ln = int(rng.integers(40, 400))
            ex.append((pos, pos + ln))
            pos += ln + int(rng.integers(30, 2000))
        loci.append(ex)
        pos += int(rng.integers(0, 3000))
    reads = []
    for li, ex in enumerate(loci):
        for _ in range(reads_per_locus):
            a = int(rng.integers(0, len(ex)))
            span = int(rng.integers(1, 9 if big else len(ex) + 1))
            pick = [k for k in range(a, min(len(ex), a + span)) if k == a or rng.random() < 0.85]
            ivs = []
            for k in pick:
                s, e = ex[k]
                s += int(rng.integers(-15, 16)) if rng.random() < 0.3 else 0
                e += int(rng.integers(-15, 16)) if rng.random() < 0.3 else 0
                if e <= s:
                    e = s + 1
                if ivs and s <= ivs[-1][1]:
                    s = ivs[-1][1] + (0 if rng.random() < 0.2 else 1)  # touching (merged by the sweep) or one apart
                    if e <= s:
                        e = s + 5
                ivs.append((s, e))
            if rng.random() < chain and li + 1 < len(loci):  # a read that continues into the next gene
                s, e = loci[li + 1][0]
                if s > ivs[-1][1]:
                    ivs.append((s, e))
            reads.append(ivs)
    order = sorted(range(len(reads)), key=lambda i: reads[i][0][0])  # read_sam yields reads by start
    return [reads[i] for i in order]


GOLDEN_GROUPS = {
    "tiny": dict(seed=1, n_loci=1, reads_per_locus=2),
    "three_reads": dict(seed=2, n_loci=1, reads_per_locus=3),
    "small": dict(seed=3, n_loci=4, reads_per_locus=25),
    "chained": dict(seed=4, n_loci=12, reads_per_locus=40, chain=0.5),
    "isolated": dict(seed=5, n_loci=30, reads_per_locus=4, chain=0.0),
    "many_reads": dict(seed=6, n_loci=2, reads_per_locus=900, chain=1.0),
    "big": dict(seed=7, n_loci=1, reads_per_locus=700, big=True),
    "big_chained": dict(seed=8, n_loci=3, reads_per_locus=400, big=True, chain=0.3),
    "big_sparse": dict(seed=9, n_loci=2, reads_per_locus=160, big=True, chain=0.0),
    "big_thin": dict(seed=10, n_loci=1, reads_per_locus=70, big=True),
}
