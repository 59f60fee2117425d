"""TEST INFRASTRUCTURE -- CPU restatement of the Gurobi-free front of freddie_cluster.py, the next row of
the hot-path scope table: what happens to a SEGMENT tint between
``read_segment`` and the first ILP.  Nothing under ``freddie_b200/`` imports this file.

Follows ``/root/reference/py/freddie_cluster.py``:

* ``preprocess_ilp`` (:277-328) with ``find_segment_read`` (:175-184) and the garbage costs (:187-196):
  per read rep the incidence row ``I`` (digit % 2), the first / last covered segment ``FL`` (poly tails
  longer than 10 extend it to the tint's end and add a gap), the correction mask ``C`` and the
  poly-tail category;
* ``partition_reads`` (:198-274): reps with the same ``(I row, FL, category)`` are merged (first-seen
  order over the sorted rep ids), every pair of distinct structures is tested for compatibility on the
  overlap of their segment ranges (the O(U^2 M) step), the compatibility graph is pruned until no edge
  without a common neighbour is left between two nodes of degree > 1, and every connected component --
  cut into even pieces of at most ``maximum_ilp_size`` -- becomes a partition ``(rids, incompatible pairs)``.

Pinned: ``oracle/pin_cluster_prep.py`` runs the UNMODIFIED reference functions (gurobipy stubbed: these
functions never touch it) on the SEGMENT files of golden sets and stores a SHA-256 per tint of the canonical
serialisation below in ``tests/golden/cluster_prep.json``; ``tests/test_oracle_cluster_prep.py`` holds this
file to those digests (and to the live functions when ``/root/reference`` is present).
"""
from __future__ import annotations

import hashlib
import json
from math import ceil
from typing import Dict, List, Tuple

import numpy as np


def find_segment_read(row: np.ndarray) -> Tuple[int, int]:
    """(:175-184) first / last segment with a 1; (-1, M-1) when the row has none."""
    ones = np.flatnonzero(row == 1)
    if len(ones) == 0:
        return -1, len(row) - 1
    return int(ones[0]), int(ones[-1])


def preprocess(tint: dict, recycle_model: str = "constant") -> dict:
    """``preprocess_ilp`` (:277-328) on arrays.  ``tint`` is what ``read_segment`` returns for one tint.
    Returns I, C (U x M uint8), FL (U x 2), cat (U one-letter strings), garbage_cost, and the gaps dict
    of every rep after the poly-tail gap was added (the reference stores it back on every read of the rep)."""
    reps = tint["read_reps"]
    U, M = len(reps), len(tint["segs"])
    I = np.zeros((U, M), dtype=np.uint8)
    C = np.zeros((U, M), dtype=np.uint8)
    FL = np.zeros((U, 2), dtype=np.int64)
    cat: List[str] = []
    gaps: List[dict] = []
    for i, idxs in enumerate(reps):
        read = tint["reads"][idxs[0]]
        data = np.asarray(read["data"], dtype=np.int64)
        I[i] = data % 2
        lo, hi = find_segment_read(I[i])
        c = "N"
        g = dict(read["gaps"])
        if len(read["poly_tail"]) == 1:
            key = next(iter(read["poly_tail"]))
            val = read["poly_tail"][key]
            if key in ("SA", "ST") and val[0] > 10:
                c = "S"
                g[(-1, lo)] = val[1]
                lo = 0
            elif key in ("EA", "ET") and val[0] > 10:
                c = "E"
                g[(hi, M)] = val[1]
                hi = M - 1
        FL[i] = (lo, hi)
        j = np.arange(M)
        C[i] = ((j >= lo) & (j <= hi) & (data == 0)).astype(np.uint8)
        cat.append(c)
        gaps.append(g)
    cost = {}
    for i in range(U):
        if recycle_model in ("exons", "introns"):
            # garbage_cost_exons / garbage_cost_introns (:187-196) call .values() on the LIST rows built
            # above (:283-287): the reference raises for these two models; so does the restatement
            raise AttributeError("'list' object has no attribute 'values'")
        elif recycle_model == "constant":
            cost[i] = len(reps[i]) * 3
    return dict(I=I, C=C, FL=FL, cat=cat, garbage_cost=cost, gaps=gaps)


def unique_structures(I: np.ndarray, FL: np.ndarray, cat: List[str]):
    """(:210-218) structures in first-seen order: [(row tuple, (f, l, category)), [rep ids]]."""
    seen: Dict[tuple, int] = {}
    out: List[Tuple[tuple, List[int]]] = []
    for i in range(len(I)):
        d = (tuple(int(x) for x in I[i]), (int(FL[i][0]), int(FL[i][1]), cat[i]))
        k = seen.get(d)
        if k is None:
            seen[d] = len(out)
            out.append((d, [i]))
        else:
            out[k][1].append(i)
    return out


def compatible(d1, d2) -> bool:
    """(:222-236) the pair test.  Ranges use Python slice semantics on purpose: a row without any 1 has
    f = -1, and ``row[-1:l+1]`` is what the reference evaluates."""
    r1, (f1, l1, t1) = d1
    r2, (f2, l2, t2) = d2
    if t1 != "N" and t2 != "N" and t1 != t2:
        return False
    f, l = max(f1, f2), min(l1, l2)
    o = l - f + 1
    a, b = r1[f:l + 1], r2[f:l + 1]
    w = sum(x == y == 1 for x, y in zip(a, b))
    if w < 1:
        return False
    d = sum(x != y for x, y in zip(a, b))
    return (o > 3 and d < 3) or (1 <= o <= 3 and d == 0)


def compatibility_matrix_scalar(structs) -> np.ndarray:
    """N x N symmetric boolean adjacency of the compatibility graph before pruning, pair by pair with the
    reference's own expressions (the O(N^2 M) step; minutes in Python for a 2 000-read tint)."""
    n = len(structs)
    A = np.zeros((n, n), dtype=bool)
    for i in range(n):
        for j in range(i + 1, n):
            if compatible(structs[i][0], structs[j][0]):
                A[i, j] = A[j, i] = True
    return A


def compatibility_matrix(structs) -> np.ndarray:
    """The same matrix with numpy, one row against all others at a time.  The slice ``row[f:l+1]`` is
    resolved the way Python resolves it: f = -1 (a row without any 1) starts at the LAST segment."""
    n = len(structs)
    A = np.zeros((n, n), dtype=bool)
    if n == 0:
        return A
    R = np.array([d[0] for d, _ in structs], dtype=np.int8).reshape(n, -1)
    M = R.shape[1]
    f = np.array([d[1][0] for d, _ in structs], dtype=np.int64)
    l = np.array([d[1][1] for d, _ in structs], dtype=np.int64)
    t = np.array([d[1][2] for d, _ in structs])
    pos = np.arange(M)
    for i in range(n):
        F = np.maximum(f[i], f)
        L = np.minimum(l[i], l)
        o = L - F + 1
        start = np.where(F >= 0, F, M + F)               # Python: negative start counts from the end
        stop = np.where(L + 1 >= 0, L + 1, M + L + 1)     # (L + 1 >= 1 always; kept for the same rule)
        start = np.clip(start, 0, M)
        stop = np.clip(stop, 0, M)
        m = (pos[None, :] >= start[:, None]) & (pos[None, :] < stop[:, None])
        both = (R[i][None, :] == 1) & (R == 1) & m
        diff = (R[i][None, :] != R) & m
        w = both.sum(axis=1)
        d = diff.sum(axis=1)
        cat_ok = ~((t[i] != "N") & (t != "N") & (t[i] != t))
        ok = cat_ok & (w >= 1) & (((o > 3) & (d < 3)) | ((o >= 1) & (o <= 3) & (d == 0)))
        ok[i] = False
        A[i] = ok
    assert (A == A.T).all()
    return A


def prune(A: np.ndarray) -> np.ndarray:
    """(:242-255) rounds of synchronous edge removal: an edge (i, j) stays if i or j has no other neighbour
    or if they have a common neighbour; all removals of a round are decided on the graph of its start."""
    A = A.copy()
    while True:
        deg = A.sum(axis=1)
        Af = A.astype(np.float32)  # BLAS; counts < 2^24 are exact
        common = (Af @ Af) > 0.5
        keep = (deg[:, None] == 1) | (deg[None, :] == 1) | common
        remove = A & ~keep
        if not remove.any():
            return A
        A &= ~remove


def components(A: np.ndarray) -> List[List[int]]:
    """Connected components in the order networkx yields them (by smallest node), each sorted (:257-259)."""
    n = len(A)
    seen = np.zeros(n, dtype=bool)
    out = []
    for s in range(n):
        if seen[s]:
            continue
        comp, stack = [], [s]
        seen[s] = True
        while stack:
            v = stack.pop()
            comp.append(v)
            for u in np.flatnonzero(A[v]):
                if not seen[u]:
                    seen[u] = True
                    stack.append(int(u))
        out.append(sorted(comp))
    return out


def split_list_evenly(l: List[int], m: int):
    """(:112-116)"""
    p = ceil(len(l) / m)
    s = ceil(len(l) / p)
    for idx in range(0, p * s, s):
        yield l[idx:idx + s]


def partitions(structs, A: np.ndarray, maximum_ilp_size: int) -> List[Tuple[List[int], List[Tuple[int, int]]]]:
    """(:257-274) per piece of a component: its rep ids, and every pair of reps whose structures are in the
    piece but not adjacent in the pruned graph."""
    out = []
    for comp in components(A):
        for c in split_list_evenly(comp, maximum_ilp_size):
            rids: List[int] = []
            incomp: List[Tuple[int, int]] = []
            for idx, i in enumerate(c):
                rids.extend(structs[i][1])
                for j in c[idx + 1:]:
                    if A[i, j]:
                        continue
                    for r1 in structs[i][1]:
                        for r2 in structs[j][1]:
                            incomp.append((r1, r2))
            out.append((rids, incomp))
    return out


def cluster_prep(tint: dict, recycle_model: str = "constant", maximum_ilp_size: int = 1000) -> dict:
    """preprocess_ilp + partition_reads of one tint."""
    pre = preprocess(tint, recycle_model)
    structs = unique_structures(pre["I"], pre["FL"], pre["cat"])
    A0 = compatibility_matrix(structs)
    A = prune(A0)
    pre.update(structs=structs, edges_before=int(A0.sum()) // 2, edges_after=int(A.sum()) // 2,
               partitions=partitions(structs, A, maximum_ilp_size))
    return pre


def canonical(I, C, FL, cat, garbage_cost, gaps, parts) -> str:
    """Canonical JSON of the results (same function for the reference's dicts and this file's arrays)."""
    U = len(cat)
    doc = dict(
        I=[[int(x) for x in I[i]] for i in range(U)],
        C=[[int(x) for x in C[i]] for i in range(U)],
        FL=[[int(FL[i][0]), int(FL[i][1])] for i in range(U)],
        cat=list(cat),
        garbage_cost=[float(garbage_cost[i]) for i in range(U)] if len(garbage_cost) else [],
        gaps=[sorted([int(k[0]), int(k[1]), int(v)] for k, v in gaps[i].items()) for i in range(U)],
        partitions=[[[int(r) for r in rids], [[int(a), int(b)] for a, b in inc]] for rids, inc in parts],
    )
    return json.dumps(doc, sort_keys=True, separators=(",", ":"))


def digest_of(res: dict) -> str:
    s = canonical(res["I"], res["C"], res["FL"], res["cat"], res["garbage_cost"], res["gaps"], res["partitions"])
    return hashlib.sha256(s.encode()).hexdigest()
