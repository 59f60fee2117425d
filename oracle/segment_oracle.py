"""CPU ORACLE for Freddie's segment stage -- TEST INFRASTRUCTURE, NOT THE PRODUCT.

A from-scratch restatement (numpy + plain Python) of the algorithm in the reference's
``py/freddie_segment.py`` (vpc-ccg/freddie).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` leg may import this file; the product path
(``freddie_b200``) never does and fails loudly when its CUDA library is missing.

Parity pin: the reference ships **no tests or golden vectors** (``test/.gitignore:1`` is ``*``).
This oracle is pinned against outputs of the reference itself, run unmodified in the authoring
container (``oracle/pin_against_reference.py``; byte-identical SEGMENT directories and bit-identical
intermediates), and the resulting golden vectors are committed under ``tests/golden/`` together with
the generating script.  The floating-point sub-steps live in third-party compiled code that is not
under ``/root/reference`` (scipy 1.18.1 ``ndimage.correlate1d`` / ``signal._peak_finding_utils``,
numpy 2.3.5 pairwise ``add.reduce``); their published algorithms are restated here and checked
bit-for-bit against the installed libraries in ``tests/test_oracle_steps.py``.

Every function cites the reference lines it follows (``freddie_segment.py:<lines>`` unless noted).
"""
from __future__ import annotations

import math
import os
import re
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NEG_INF = float("-inf")

# --------------------------------------------------------------------------------------------
# parameters (freddie_segment.py:53-110, :269-286)
# --------------------------------------------------------------------------------------------


def smooth_threshold(threshold: float) -> List[float]:
    """Per-length high-threshold table; follows ``smooth_threshold`` (:277-286)."""
    table: List[float] = []
    while True:
        x = len(table)
        y = threshold / (1 + ((threshold - 0.5) / 0.5) * math.exp(-0.05 * x))
        if x > 5 and x * (threshold - y) < 0.5:
            break
        table.append(round(y, 2))
        assert len(table) < 1000
    return table


def high_threshold(seg_len: int, table: Sequence[float], tp: float) -> float:
    """``get_high_threshold`` (:269-274)."""
    return table[seg_len] if seg_len < len(table) else tp


class Params:
    """CLI defaults of ``parse_args`` (:53-110)."""

    def __init__(self, sigma=5.0, tp=0.90, vf=3.0, mps=50, lo=3, consider_ends=False):
        assert 1 >= tp >= 0.5
        assert 10 > vf > 0
        assert 50 >= sigma > 0
        assert mps > 3
        assert lo >= 0
        self.sigma = float(sigma)
        self.tp = float(tp)
        self.vf = float(vf)
        self.mps = int(mps)
        self.lo = int(lo)
        self.ignore_ends = not consider_ends
        self.table = smooth_threshold(self.tp)


# --------------------------------------------------------------------------------------------
# SPLIT parsing (grammar: freddie_segment.py:17-38, :121-185; writer freddie_split.py:445-481)
# --------------------------------------------------------------------------------------------
_CHR = r"[0-9A-Za-z!#$%&+./:;?@^_|~-][0-9A-Za-z!#$%&*+./:;=?@^_|~-]*"
_TINT_LINE = re.compile(r"#(%s)\t([0-9]+)\t([0-9]+-[0-9]+(?:,[0-9]+-[0-9]+)*)\t([0-9]+)\n$" % _CHR)
_RIV = r"[0-9]+-[0-9]+:[0-9]+-[0-9]+:(?:[0-9]+[MIDNSHPX=])+"
_READ_LINE = re.compile(
    r"([0-9]+)\t([!-?A-~]{1,254})\t(%s)\t([+-])\t([0-9]+)\t(%s(?:\t%s)*)\n$" % (_CHR, _RIV, _RIV))
_RIV_PARTS = re.compile(r"([0-9]+)-([0-9]+):([0-9]+)-([0-9]+):((?:[0-9]+[MIDNSHPX=])+)")
_CIG_OP = re.compile(r"([0-9]+)([MIDNSHPX=])")


def parse_split(split_tsv: str) -> dict:
    """One tint per file (``run_segment`` asserts it, :699).  Follows ``read_split`` (:121-171)."""
    tint = None
    with open(split_tsv) as fh:
        for line in fh:
            if line[0] == "#":
                m = _TINT_LINE.match(line)
                assert m is not None, line
                assert tint is None, "more than one tint in %s" % split_tsv
                islands = [tuple(int(v) for v in x.split("-")) for x in m.group(3).split(",")]
                assert all(a[1] < b[0] for a, b in zip(islands[:-1], islands[1:])), islands
                assert all(s < e for s, e in islands)
                tint = dict(id=int(m.group(2)), chr=m.group(1), intervals=islands,
                            read_count=int(m.group(4)), reads=[])
            else:
                m = _READ_LINE.match(line)
                assert m is not None, line
                ivs = []
                for p in _RIV_PARTS.findall(m.group(6)):
                    ivs.append((int(p[0]), int(p[1]), int(p[2]), int(p[3]),
                                [(int(c), t) for c, t in _CIG_OP.findall(p[4])]))
                assert all(a[1] <= b[0] and a[3] <= b[2] for a, b in zip(ivs[:-1], ivs[1:]))
                assert all(iv[0] < iv[1] and iv[2] < iv[3] for iv in ivs)
                read = dict(id=int(m.group(1)), name=m.group(2), chr=m.group(3), strand=m.group(4),
                            tint=int(m.group(5)), intervals=ivs)
                assert tint is not None and read["tint"] == tint["id"]
                tint["reads"].append(read)
    assert tint is not None
    assert len(tint["reads"]) == tint["read_count"]
    return tint


def parse_reads(tint: dict, reads_tsv: str) -> None:
    """``read_sequence`` (:174-185): columns 0 (rid) and 3 (sequence) only."""
    seqs = {}
    with open(reads_tsv) as fh:
        for line in fh:
            cols = line.rstrip().split("\t")
            seqs[int(cols[0])] = cols[3]
    assert len(seqs) == len(tint["reads"])
    for r in tint["reads"]:
        r["seq"] = seqs[r["id"]]
        r["length"] = len(r["seq"])


def build_reps(tint: dict) -> Tuple[List[tuple], List[List[int]]]:
    """Read reps keyed by the tuple of target intervals, first-seen order (:165-170)."""
    index: Dict[tuple, int] = {}
    keys: List[tuple] = []
    members: List[List[int]] = []
    for ridx, r in enumerate(tint["reads"]):
        k = tuple((iv[0], iv[1]) for iv in r["intervals"])
        j = index.get(k)
        if j is None:
            j = len(keys)
            index[k] = j
            keys.append(k)
            members.append([])
        members[j].append(ridx)
    return keys, members


# --------------------------------------------------------------------------------------------
# A.1 splice signal (process_splicing_data, :648-678)
# --------------------------------------------------------------------------------------------
def island_of(islands: Sequence[Tuple[int, int]], p: int) -> int:
    """Island index containing position ``p`` (inclusive ends, :652-659); KeyError if none."""
    lo, hi = 0, len(islands) - 1
    while lo <= hi:
        mid = (lo + hi) // 2
        s, e = islands[mid]
        if p < s:
            hi = mid - 1
        elif p > e:
            lo = mid + 1
        else:
            return mid
    raise KeyError(p)


def splice_signal(islands, rep_keys, rep_w, ignore_ends: bool) -> List[np.ndarray]:
    Y_raw = [np.zeros(e - s + 1) for s, e in islands]
    for key, w in zip(rep_keys, rep_w):
        m = len(key)
        for idx, (ts, te) in enumerate(key):
            a = island_of(islands, ts)
            b = island_of(islands, te)
            assert a == b, (ts, te)
            s = islands[a][0]
            if not (ignore_ends and idx == 0):
                Y_raw[a][ts - s] += w
            if not (ignore_ends and idx == m - 1):
                Y_raw[a][te - s] += w
    return Y_raw


# --------------------------------------------------------------------------------------------
# A.2 Gaussian (scipy.ndimage.gaussian_filter1d -> correlate1d, symmetric branch)
# --------------------------------------------------------------------------------------------
def gaussian_weights(sigma: float, truncate: float) -> np.ndarray:
    """scipy ``_gaussian_kernel1d`` (``_filters.py:656-666``) with ``lw = int(truncate*sd+0.5)``
    (``_filters.py:745``); computed with numpy exactly as scipy does."""
    lw = int(truncate * float(sigma) + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / sigma2 * x ** 2)
    return phi / phi.sum()


def _reflect_index(idx: np.ndarray, n: int) -> np.ndarray:
    """scipy ``mode='reflect'`` (d c b a | a b c d | d c b a), valid for any overhang."""
    j = np.mod(idx, 2 * n)
    return np.where(j < n, j, 2 * n - 1 - j)


def gaussian_filter(y: np.ndarray, weights: np.ndarray, mode: str = "reflect") -> np.ndarray:
    """Symmetric-pair evaluation order of scipy's ``correlate1d`` (``ni_filters.c``, symmetric
    case): ``acc = y[l]*w[c]; for jj=-lw..-1: acc += (y[l+jj] + y[l-jj]) * w[c+jj]`` with separate
    roundings (numpy elementwise ops never fuse).  ``mode``: 'reflect' (:755) or 'constant' (:260)."""
    y = np.asarray(y, dtype=np.float64)
    n = len(y)
    lw = (len(weights) - 1) // 2
    idx = np.arange(-lw, n + lw)
    if mode == "reflect":
        ext = y[_reflect_index(idx, n)]
    else:
        ext = np.zeros(n + 2 * lw)
        ext[lw:lw + n] = y
    acc = ext[lw:lw + n] * weights[lw]
    for jj in range(-lw, 0):
        left = ext[lw + jj: lw + jj + n]
        right = ext[lw - jj: lw - jj + n]
        acc = acc + (left + right) * weights[lw + jj]
    return acc


# --------------------------------------------------------------------------------------------
# A.3 variance threshold (numpy pairwise add.reduce; _methods.py mean/std)
# --------------------------------------------------------------------------------------------
def pairwise_sum(a: Sequence[float]) -> float:
    """numpy ``DOUBLE_pairwise_sum`` (``loops_utils.h.src``): n<8 sequential; n<=128 eight
    strided accumulators; else split at ``n/2`` rounded down to a multiple of 8."""
    n = len(a)
    if n < 8:
        res = 0.0
        for i in range(n):
            res += float(a[i])
        # numpy starts from -0.0 for an empty/short sum?  It starts at 0. for n<8 via `res = 0.`
        return res
    if n <= 128:
        r = [float(a[k]) for k in range(8)]
        i = 8
        while i < n - (n % 8):
            for k in range(8):
                r[k] += float(a[i + k])
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += float(a[i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def variance_threshold(Y: Sequence[np.ndarray], vf: float) -> float:
    """``mean + vf*std`` of all positive smoothed samples, islands in order (:757-759).
    numpy: ``mean = add.reduce(v)/n``; ``std = sqrt(add.reduce((v-mean)**2)/n)``.  NaN if empty."""
    v = np.concatenate([y[y > 0] for y in Y]) if len(Y) else np.zeros(0)
    n = len(v)
    if n == 0:
        return float("nan")
    mean = pairwise_sum(v) / n
    d = v - mean
    var = pairwise_sum(d * d) / n
    return mean + vf * math.sqrt(var)


# --------------------------------------------------------------------------------------------
# A.4 candidates (scipy _local_maxima_1d; candidates_from_peaks :615-621)
# --------------------------------------------------------------------------------------------
def local_maxima(x: Sequence[float]) -> List[int]:
    """Strict local maxima, plateaus -> floor midpoint, ends never peaks
    (scipy ``_peak_finding_utils.pyx:_local_maxima_1d``)."""
    n = len(x)
    out = []
    i = 1
    i_max = n - 1
    while i < i_max:
        if x[i - 1] < x[i]:
            ia = i + 1
            while ia < i_max and x[ia] == x[i]:
                ia += 1
            if x[ia] < x[i]:
                out.append((i + ia - 1) // 2)
                i = ia
        i += 1
    return out


def local_maxima_np(x: np.ndarray) -> np.ndarray:
    """Vectorised equivalent of :func:`local_maxima` (used for large inputs)."""
    x = np.asarray(x)
    n = len(x)
    if n < 3:
        return np.zeros(0, dtype=np.int64)
    # run-length encode equal values
    change = np.flatnonzero(x[1:] != x[:-1]) + 1
    starts = np.concatenate([[0], change])
    ends = np.concatenate([change, [n]]) - 1  # inclusive
    vals = x[starts]
    k = len(starts)
    if k < 3:
        return np.zeros(0, dtype=np.int64)
    mid = np.arange(1, k - 1)
    ok = (vals[mid - 1] < vals[mid]) & (vals[mid + 1] < vals[mid])
    m = mid[ok]
    return (starts[m] + ends[m]) // 2


def candidates(y: np.ndarray) -> List[int]:
    c = set(int(v) for v in local_maxima_np(y))
    c.add(0)
    c.add(len(y) - 1)
    return sorted(c)


# --------------------------------------------------------------------------------------------
# A.5 cumulative coverage, closed form of get_cumulative_coverage (:188-246)
# --------------------------------------------------------------------------------------------
def coverage_matrix(rep_iv_in_island: Sequence[Sequence[Tuple[int, int]]], cand: Sequence[int]) -> np.ndarray:
    """``C[c][r] = sum_intervals clamp(cand[c]-ys, 0, ye-ys+1)`` for c<K, ``C[K][r]`` = total;
    (``te`` is used as an INCLUSIVE sample, :205,:225-235).  ``rep_iv_in_island[r]`` lists the
    rep's (ys, ye) island-local intervals."""
    K = len(cand)
    R = len(rep_iv_in_island)
    C = np.zeros((K + 1, R), dtype=np.uint32)
    cv = np.asarray(cand, dtype=np.int64)
    for r, ivs in enumerate(rep_iv_in_island):
        if not ivs:
            continue
        col = np.zeros(K + 1, dtype=np.int64)
        for ys, ye in ivs:
            col[:K] += np.clip(cv - ys, 0, ye - ys + 1)
            col[K] += ye - ys + 1
        C[:, r] = col
    return C


# --------------------------------------------------------------------------------------------
# A.6 fixed candidates (:776-788) and break_large_problems (:623-645)
# --------------------------------------------------------------------------------------------
def fixed_candidates(y: np.ndarray, cand: Sequence[int], thr: float, mps: int) -> List[int]:
    K = len(cand)
    fixed = {0, K - 1}
    for c, yi in enumerate(cand):
        if y[yi] > thr:
            fixed.add(c)
    snap = sorted(fixed)
    for s, e in zip(snap[:-1], snap[1:]):
        size = e - s + 1
        if size <= mps:
            continue
        cnt = math.ceil(size / mps)
        ps = size / cnt
        for i in range(1, cnt):
            mid = int(s + i * ps)
            best_v = NEG_INF
            best_c = None
            for c in range(mid - 5, mid + 5):
                if y[cand[c]] > best_v:
                    best_v = y[cand[c]]
                    best_c = c
            assert best_v > 0
            fixed.add(best_c)
    return sorted(fixed)


# --------------------------------------------------------------------------------------------
# A.7 interval-scoring DP (optimize :475-568, run_optimize :571-596), table form
# --------------------------------------------------------------------------------------------
def pair_thresholds(seg_len: int, table: Sequence[float], tp: float) -> Tuple[float, float]:
    h = high_threshold(seg_len, table, tp)
    return h, 1 - h


def dp_tables(cand, C, W, start, end, table, tp):
    """yea/nay masks (:488-497), ``ins`` (:500-506) and ``out`` (:509-528, before the ``lo`` cut).
    Returns (ins[n][n] int64, out[n][n][n] int64) indexed by local offsets (i-start, ...)."""
    n = end - start + 1
    yea = {}
    nay = {}
    ins = np.zeros((n, n), dtype=np.int64)
    Wv = np.asarray(W, dtype=np.int64)
    for i in range(start, end):
        for j in range(i, end + 1):
            seg_len = cand[j] - cand[i] + 1
            h, l = pair_thresholds(seg_len, table, tp)
            c = (C[j] - C[i]) / seg_len
            y = c > h
            z = c < l
            yea[(i, j)] = y
            nay[(i, j)] = z
            if i != j:
                ins[i - start, j - start] = -int((Wv * ~(y | z)).sum())
    out = np.zeros((n, n, n), dtype=np.int64)
    for i in range(start, end):
        for j in range(i + 1, end):
            for k in range(j + 1, end + 1):
                x = (yea[(i, j)] & nay[(j, k)]) | (nay[(i, j)] & yea[(j, k)])
                out[i - start, j - start, k - start] = int((Wv * x).sum())
    return ins, out


def dp_solve(cand, start, end, ins, out, lo) -> List[int]:
    """Recurrence of ``dp`` (:532-558) + top level (:560-566) + backtrace (:592-594), restated as
    a 2-D table ``G(j,k)`` = best continuation after committing segment (j,k):
    ``D(i,j,k) = ins(i,j) + out(i,j,k) + G(j,k)`` when both segments are >= 5 long and
    ``out >= lo``; ``G(j,end) = ins(j,end)``; ``G(j,k) = max_{k'>k} D(j,k,k')`` (first max wins)."""
    n = end - start + 1
    cv = [cand[start + t] for t in range(n)]
    E = n - 1

    def ok(a, b):
        return cv[b] - cv[a] >= 5

    G = [[NEG_INF] * n for _ in range(n)]
    arg = [[-1] * n for _ in range(n)]

    def D(i, j, k):
        if not ok(i, j) or not ok(j, k):
            return NEG_INF
        o = out[i, j, k]
        if o < lo:
            return NEG_INF
        g = G[j][k]
        if g == NEG_INF:
            return NEG_INF
        return int(ins[i, j]) + int(o) + g

    for j in range(E - 1, -1, -1):
        G[j][E] = int(ins[j, E])
        for k in range(E - 1, j, -1):
            best = NEG_INF
            bk = -1
            for k2 in range(k + 1, E + 1):
                d = D(j, k, k2)
                if d > best:
                    best = d
                    bk = k2
            G[j][k] = best
            arg[j][k] = bk
    best = int(ins[0, E])
    choice = None
    for j in range(1, E):
        for k in range(j + 1, E + 1):
            d = D(0, j, k)
            if d > best:
                best = d
                choice = (0, j, k)
    chosen = []
    if choice is not None:
        i, j, k = choice
        chosen.extend([i, j, k])
        while k != E:
            k2 = arg[j][k]
            assert k2 > 0
            chosen.append(k2)
            j, k = k, k2
    return sorted(set(start + t for t in chosen))


def run_dp(cand, fixed, C, W, table, tp, lo, stats: Optional[dict] = None) -> List[int]:
    final = set(fixed)
    for s, e in zip(fixed[:-1], fixed[1:]):
        if e - s < 2:
            continue
        ins, out = dp_tables(cand, C, W, s, e, table, tp)
        final.update(dp_solve(cand, s, e, ins, out, lo))
        if stats is not None:
            n = e - s + 1
            stats["cells"] = stats.get("cells", 0) + n * (n - 1) * (n - 2) // 6
            stats["subproblems"] = stats.get("subproblems", 0) + 1
            stats["max_n"] = max(stats.get("max_n", 0), n)
    return sorted(final)


# --------------------------------------------------------------------------------------------
# A.9 refine (refine_segmentation :249-266; scipy find_peaks(distance=20))
# --------------------------------------------------------------------------------------------
def select_by_distance(peaks: Sequence[int], priority: Sequence[float], distance: int,
                       tie_log: Optional[list] = None) -> List[int]:
    """scipy ``_select_by_peak_distance``: visit peaks by descending priority, each kept peak
    removes neighbours closer than ``distance``.  The reference inherits ``np.argsort``'s
    (platform-dependent, unstable) order among bit-equal heights; this restatement defines the
    order as a STABLE ascending argsort (equal heights: the later peak is visited first) and
    records interacting ties in ``tie_log``."""
    m = len(peaks)
    keep = [True] * m
    order = sorted(range(m), key=lambda t: priority[t])  # stable
    if TIE_EXPLORER is not None:
        order = TIE_EXPLORER.reorder(order, priority)
    if tie_log is not None:
        for a in range(m - 1):
            if priority[a] == priority[a + 1] and peaks[a + 1] - peaks[a] < distance:
                tie_log.append((peaks[a], peaks[a + 1]))
    for t in range(m - 1, -1, -1):
        j = order[t]
        if not keep[j]:
            continue
        k = j - 1
        while k >= 0 and peaks[j] - peaks[k] < distance:
            keep[k] = False
            k -= 1
        k = j + 1
        while k < m and peaks[k] - peaks[j] < distance:
            keep[k] = False
            k += 1
    return [peaks[t] for t in range(m) if keep[t]]


class TieExplorer:
    """Test helper for the one platform-dependent corner of the reference: scipy's
    ``_select_by_peak_distance`` visits peaks in ``np.argsort(priority)`` order, and numpy's default sort is
    NOT stable (AVX-512 / AVX2 sorting networks): among bit-equal heights the order is an artefact of the
    network, not of the data.  The restatement (and the CUDA kernel) define it as the stable order.  To show
    that a SEGMENT file of the reference that differs from ours differs ONLY by such a choice, this class
    re-orders every group of equal priorities by a chosen permutation; ``explain_by_ties`` searches the
    permutations for one that reproduces the reference's bytes."""

    def __init__(self):
        self.sizes: List[int] = []   # probe mode: sizes of the equal-priority groups, in call order
        self.choice = None           # list of permutations (tuples), one per group; None = probe

    def reorder(self, order, priority):
        out, i = [], 0
        while i < len(order):
            j = i
            while j + 1 < len(order) and priority[order[j + 1]] == priority[order[i]]:
                j += 1
            grp = order[i:j + 1]
            if len(grp) > 1:
                if self.choice is None:
                    self.sizes.append(len(grp))
                else:
                    perm = self.choice[self._k] if self._k < len(self.choice) else None
                    self._k += 1
                    if perm is not None:
                        grp = [grp[x] for x in perm]
            out.extend(grp)
            i = j + 1
        return out


TIE_EXPLORER: Optional[TieExplorer] = None


def explain_by_ties(tint: dict, prm: "Params", want_text: Optional[str] = None, limit: int = 20000,
                    want_sha16: Optional[str] = None) -> Optional[list]:
    """Searches the visiting orders among equal-height refine peaks for one under which this restatement
    writes ``want_text`` (or a file whose SHA-256 starts with ``want_sha16``) for ``tint``.  Returns the chosen
    permutations (one per group of equal heights), or None if no order within ``limit`` combinations does --
    i.e. the difference is NOT a tie artefact."""
    import copy
    import hashlib
    import itertools
    global TIE_EXPLORER
    probe = TieExplorer()
    TIE_EXPLORER = probe
    try:
        segment_tint(copy.deepcopy(tint), prm)
    finally:
        TIE_EXPLORER = None
    options = [list(itertools.permutations(range(n))) if n <= 4 else [tuple(range(n)), tuple(reversed(range(n)))]
               for n in probe.sizes]
    total = 1
    for o in options:
        total *= len(o)
        if total > limit:
            return None
    for combo in itertools.product(*options):
        ex = TieExplorer()
        ex.choice, ex._k = list(combo), 0
        TIE_EXPLORER = ex
        try:
            t = copy.deepcopy(tint)
            segment_tint(t, prm)
        finally:
            TIE_EXPLORER = None
        text = format_segment(t)
        if (want_text is not None and text == want_text) or (
                want_sha16 is not None and hashlib.sha256(text.encode()).hexdigest()[:16] == want_sha16):
            return list(combo)
    return None


def refine(y_raw: np.ndarray, final: Sequence[int], sigma: float, weights1: np.ndarray,
           tie_log: Optional[list] = None) -> List[int]:
    skip = 20
    need = 20
    extra = []
    for a, b in zip(final[:-1], final[1:]):
        if b - a <= 2 * skip:
            continue
        v = np.array(y_raw[a:b], dtype=np.float64)
        v[:skip] = 0.0
        v[len(v) - skip:] = 0.0
        if float(v.sum()) < need:  # integers: exact in any order
            continue
        g = gaussian_filter(v, weights1, mode="constant")
        peaks = [int(p) for p in local_maxima_np(g)]
        peaks = select_by_distance(peaks, [g[p] for p in peaks], skip, tie_log)
        for p in peaks:
            lo_i = int(round(p - sigma))
            hi_i = int(round(p + sigma + 1))
            s = 0
            for val in g[lo_i:hi_i].tolist():  # python slice semantics, sequential sum from int 0
                s = s + val
            if s < need:
                continue
            extra.append(p + a)
    return extra


# --------------------------------------------------------------------------------------------
# A.10 digits (:808-838)
# --------------------------------------------------------------------------------------------
def digits_for_island(final: Sequence[int], Cf: np.ndarray, table, tp) -> np.ndarray:
    """(F-1) x R matrix of 1/0/2."""
    F = len(final)
    R = Cf.shape[1]
    out = np.zeros((max(F - 1, 0), R), dtype=np.uint8)
    for t in range(F - 1):
        seg_len = final[t + 1] - final[t] + 1
        h, l = pair_thresholds(seg_len, table, tp)
        ratio = (Cf[t + 1] - Cf[t]) / seg_len
        assert np.all((ratio >= 0) & (ratio <= 1)), "cov_ratio out of [0,1] (:821)"
        out[t] = np.where(ratio > h, 1, np.where(ratio < l, 0, 2))
    return out


# --------------------------------------------------------------------------------------------
# A.11 gaps / poly-A (:289-472)
# --------------------------------------------------------------------------------------------
def thread_cigar(cigar, t_goal, t_pos, q_pos):
    """``forward_thread_cigar`` (:289-304) incl. the quirk that ``I`` lengths are also clipped by
    the remaining target distance."""
    assert t_pos <= t_goal
    idx = 0
    while t_pos < t_goal:
        c, t = cigar[idx]
        c = min(c, t_goal - t_pos)
        if t in "MX=":
            t_pos += c
            q_pos += c
        elif t == "D":
            t_pos += c
        elif t == "I":
            q_pos += c
        idx += 1
    assert t_pos == t_goal
    return q_pos


def interval_start(p, intervals):
    """``get_interval_start`` (:307-326)."""
    for (ts, te, qs, qe, cig) in intervals:
        if te < p:
            continue
        if p < ts:
            return qs, p - ts
        q = thread_cigar(cig, p, ts, qs)
        assert qs <= q <= qe
        return q, 0
    raise AssertionError("interval_start: position beyond read")


def interval_end(p, intervals):
    """``get_interval_end`` (:329-349)."""
    for (ts, te, qs, qe, cig) in reversed(intervals):
        if ts > p:
            continue
        if te < p:
            return qe, te - p
        q = thread_cigar(cig, p, ts, qs)
        assert 0 <= q <= qe
        return q, 0
    raise AssertionError("interval_end: position before read")


def longest_poly(bases: Sequence[bool]):
    """``find_longest_poly`` (:352-367) on a pre-extracted clip: ``bases[t]`` says whether scan
    position t equals the target char.  Yields (i0, len, purity) per maximal positive-score run."""
    n = len(bases)
    if n == 0:
        return
    sc = 1 if bases[0] else 0
    scores = [sc]
    for t in range(1, n):
        sc = max(0, sc + (1 if bases[t] else -2))
        scores.append(sc)
    t = 0
    while t < n:
        if scores[t] <= 0:
            t += 1
            continue
        i0 = t
        best = -1
        best_i = t
        while t < n and scores[t] > 0:
            if scores[t] >= best:  # max(zip(S,i)) -> largest index among equal scores
                best = scores[t]
                best_i = t
            t += 1
        ln = best_i + 1 - i0
        cnt = sum(1 for u in range(i0, i0 + ln) if bases[u])
        yield i0, ln, cnt / ln


_COMP = {"A": "T", "T": "A"}


def _clip_matches(seq: str, strand: str, n: int, offset: int, ch: str) -> List[bool]:
    """Scan position t of a clip of ``n`` bases that starts ``offset`` bases from the read start
    (in read orientation): '+' looks at ``seq[offset+t]``; '-' at ``seq[L-1-offset-t]`` against the
    complemented char (:393-401, :423-431)."""
    L = len(seq)
    if strand == "+":
        return [seq[offset + t] == ch for t in range(n)]
    c = _COMP[ch]
    return [seq[L - 1 - offset - t] == c for t in range(n)]


def read_gaps(read: dict, data: Sequence[int], segs: Sequence[Tuple[int, int]]) -> List[str]:
    """``get_unaligned_gaps_and_polyA`` (:370-472).  Returns the sorted gap strings."""
    if 1 not in data:
        return []
    runs = []
    t = 0
    S = len(data)
    while t < S:
        if data[t] != 1:
            t += 1
            continue
        f = t
        while t < S and data[t] == 1:
            t += 1
        runs.append((f, t - 1))
    ivs = read["intervals"]
    L = read["length"]
    seq = read["seq"]
    strand = read["strand"]
    q_ssc, _ = interval_start(segs[runs[0][0]][0], ivs)
    q_esc, _ = interval_end(segs[runs[-1][1]][1], ivs)
    assert 0 <= q_ssc <= q_esc <= L
    gaps = set()
    polys = []
    for ch in "AT":
        for i0, ln, p in longest_poly(_clip_matches(seq, strand, q_ssc, 0, ch)):
            if ln < 20 or p < 0.85:
                continue
            polys.append((i0, ln, p, ch))
    if polys:
        i0, ln, p, ch = max(polys, key=lambda x: x[2])  # first max purity
        gaps.add("S%s_%d:%d" % (ch, ln, q_ssc - i0 - ln))
        gaps.add("SSC:%d" % i0)
    else:
        gaps.add("SSC:%d" % q_ssc)
    polys = []
    for ch in "AT":
        for i0, ln, p in longest_poly(_clip_matches(seq, strand, L - q_esc, q_esc, ch)):
            if ln < 20 or p < 0.85:
                continue
            polys.append((i0, ln, p, ch))
    if polys:
        i0, ln, p, ch = max(polys, key=lambda x: x[2])
        gaps.add("E%s_%d:%d" % (ch, ln, i0))
        gaps.add("ESC:%d" % (L - q_esc - i0))
        assert L - q_esc - i0 > 0
    else:
        gaps.add("ESC:%d" % (L - q_esc))
    for (_, l1), (f2, _) in zip(runs[:-1], runs[1:]):
        qa, sa = interval_end(segs[l1][1], ivs)
        qb, sb = interval_start(segs[f2][0], ivs)
        assert 0 < qa <= qb < L, (qa, qb, L)
        size = max(0, qb - qa + sa + sb)
        assert 0 <= size < L
        gaps.add("%d-%d:%d" % (l1, f2, size))
    return sorted(gaps)


# --------------------------------------------------------------------------------------------
# whole tint (segment :738-844) and directory driver (run_segment :681-735, main :847-885)
# --------------------------------------------------------------------------------------------
def segment_tint(tint: dict, prm: Params, keep: bool = False, stats: Optional[dict] = None) -> dict:
    """Runs every step on one parsed tint (reads need ``seq``/``length``).  Sets
    ``tint['final_positions']``, ``tint['segs']``, ``read['data']``, ``read['gaps']`` like the
    reference's ``segment`` (:738-844).  With ``keep`` the per-step intermediates are returned."""
    islands = tint["intervals"]
    rep_keys, rep_members = build_reps(tint)
    W = [len(m) for m in rep_members]
    R = len(rep_keys)
    w4 = gaussian_weights(prm.sigma, 4.0)
    w1 = gaussian_weights(prm.sigma, 1.0)
    Y_raw = splice_signal(islands, rep_keys, W, prm.ignore_ends)
    Y = [gaussian_filter(y, w4, "reflect") for y in Y_raw]
    thr = variance_threshold(Y, prm.vf)
    # rep intervals per island, island-local and inclusive ends
    rep_isl: List[List[List[Tuple[int, int]]]] = [[[] for _ in range(R)] for _ in islands]
    for r, key in enumerate(rep_keys):
        for ts, te in key:
            a = island_of(islands, ts)
            rep_isl[a][r].append((ts - islands[a][0], te - islands[a][0]))
    inter = dict(Y_raw=Y_raw, Y=Y, thr=thr, cand=[], fixed=[], dp_final=[], refine=[], final=[], W=W,
                 ties=[])
    final_positions: List[int] = []
    rows = [[] for _ in range(R)]
    for a, (s, e) in enumerate(islands):
        y = Y[a]
        cand = candidates(y)
        C = coverage_matrix(rep_isl[a], cand)
        fixed = fixed_candidates(y, cand, thr, prm.mps)
        chosen = run_dp(cand, fixed, C, W, prm.table, prm.tp, prm.lo, stats)
        final = [cand[c] for c in chosen]
        extra = refine(Y_raw[a], final, prm.sigma, w1, inter["ties"])
        final = sorted(final + extra)
        final_positions.extend(s + f for f in final)
        Cf = coverage_matrix(rep_isl[a], final)
        dg = digits_for_island(final, Cf, prm.table, prm.tp)
        for r in range(R):
            rows[r].extend(dg[:, r].tolist())
            rows[r].append(0)
        if keep:
            inter["cand"].append(cand)
            inter["fixed"].append(fixed)
            inter["dp_final"].append(chosen)
            inter["refine"].append(extra)
            inter["final"].append(final)
    tint["final_positions"] = final_positions
    tint["segs"] = list(zip(final_positions[:-1], final_positions[1:]))
    for r, members in enumerate(rep_members):
        rows[r].pop()
        for ridx in members:
            tint["reads"][ridx]["data"] = rows[r]
    for read in tint["reads"]:
        assert len(read["data"]) == len(tint["segs"])
        read["gaps"] = read_gaps(read, read["data"], tint["segs"])
    inter["rep_members"] = rep_members
    inter["rows"] = rows
    return inter


def format_segment(tint: dict) -> str:
    """Output grammar of ``run_segment`` (:715-731)."""
    out = ["#%s\t%d\t%s\n" % (tint["chr"], tint["id"], ",".join(map(str, tint["final_positions"])))]
    for r in tint["reads"]:
        out.append("\t".join([str(r["id"]), r["name"], r["chr"], r["strand"], str(r["tint"]),
                              "".join(map(str, r["data"])), "".join("%s," % g for g in r["gaps"])]) + "\n")
    return "".join(out)


def list_tints(split_dir: str) -> List[Tuple[str, int]]:
    """Directory walk of ``main`` (:852-857): sub-directories are contigs; ``split_*.tsv``."""
    out = []
    for contig in sorted(os.listdir(split_dir)):
        d = os.path.join(split_dir, contig)
        if not os.path.isdir(d):
            continue
        for fn in sorted(os.listdir(d)):
            if fn.startswith("split_") and fn.endswith(".tsv"):
                out.append((contig, int(fn[:-4].split("_")[-1])))
    return out


def run_tint_files(split_dir: str, outdir: str, contig: str, tint_id: int, prm: Params) -> int:
    tint = parse_split(os.path.join(split_dir, contig, "split_%s_%d.tsv" % (contig, tint_id)))
    parse_reads(tint, os.path.join(split_dir, contig, "reads_%s_%d.tsv" % (contig, tint_id)))
    segment_tint(tint, prm)
    os.makedirs(os.path.join(outdir, contig), exist_ok=True)
    open(os.path.join(outdir, contig, "segment_%s_%d.log" % (contig, tint_id)), "w").close()
    with open(os.path.join(outdir, contig, "segment_%s_%d.tsv" % (contig, tint_id)), "w") as f:
        f.write(format_segment(tint))
    return len(tint["reads"])


def _worker(args):
    return run_tint_files(*args)


def run_dir(split_dir: str, outdir: str, prm: Params, threads: int = 1) -> int:
    jobs = [(split_dir, outdir, c, t, prm) for c, t in list_tints(split_dir)]
    if threads > 1:
        from multiprocessing import Pool
        with Pool(threads) as p:
            return sum(p.imap_unordered(_worker, jobs, chunksize=1))
    return sum(_worker(j) for j in jobs)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="oracle run of the segment stage (test infrastructure)")
    ap.add_argument("-s", "--split-dir", required=True)
    ap.add_argument("-o", "--outdir", default="freddie_segment/")
    ap.add_argument("-t", "--threads", type=int, default=1)
    ap.add_argument("-sd", "--sigma", type=float, default=5.0)
    ap.add_argument("-tp", "--threshold-rate", type=float, default=0.90)
    ap.add_argument("-vf", "--variance-factor", type=float, default=3.0)
    ap.add_argument("-mps", "--max-problem-size", type=int, default=50)
    ap.add_argument("-lo", "--min-read-support-outside", type=int, default=3)
    ap.add_argument("--consider-ends", action="store_true")
    a = ap.parse_args()
    n = run_dir(a.split_dir.rstrip("/"), a.outdir,
                Params(a.sigma, a.threshold_rate, a.variance_factor, a.max_problem_size,
                       a.min_read_support_outside, a.consider_ends), a.threads)
    print("[oracle] %d reads" % n)
