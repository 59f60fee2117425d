"""Seeded synthetic SPLIT data (ONT-like) for the segment stage.

Produces in-memory tints shaped like what ``read_split`` + ``read_sequence`` build in the
reference (``py/freddie_segment.py:121-185``) and can write them as a SPLIT directory that
follows the writer's grammar (``py/freddie_split.py:445-481``) and the input
contract the reference asserts.  The five ``BASELINE.json`` configs are
realised by :func:`make_config`.

Nothing here is on the hot path; it only feeds tests and ``bench.py``.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


# --------------------------------------------------------------------------------------------
# gene model
# --------------------------------------------------------------------------------------------
def _gene_model(rng: np.random.Generator, base: int, locus_len: int, n_exons: int):
    """Exons (start, end) half-open, sorted, separated by introns >= 80 bp, plus alt sites."""
    n_exons = max(1, n_exons)
    # exon lengths: log-uniform 120..3000, scaled down if the locus is too small
    lens = np.exp(rng.uniform(math.log(120), math.log(3000), size=n_exons)).astype(np.int64)
    min_intron = 80
    budget = locus_len - min_intron * (n_exons + 1)
    if lens.sum() > 0.6 * budget:
        lens = np.maximum(60, (lens * (0.6 * budget / lens.sum())).astype(np.int64))
    slack = locus_len - lens.sum() - min_intron * (n_exons + 1)
    slack = max(int(slack), 0)
    cuts = np.sort(rng.integers(0, slack + 1, size=n_exons))
    gaps = np.diff(np.concatenate([[0], cuts])) + min_intron
    exons = []
    pos = base
    for i in range(n_exons):
        pos += int(gaps[i])
        exons.append((pos, pos + int(lens[i])))
        pos += int(lens[i])
    alt = []
    for (s, e) in exons:
        n_alt_s = int(rng.integers(0, 3))
        n_alt_e = int(rng.integers(0, 3))
        span = max(1, min(60, (e - s) // 4))
        a_s = [s] + [s + int(rng.integers(5, 5 + span)) for _ in range(n_alt_s)]
        a_e = [e] + [e - int(rng.integers(5, 5 + span)) for _ in range(n_alt_e)]
        alt.append((a_s, a_e))
    return exons, alt


def _isoforms(rng: np.random.Generator, exons, alt, n_iso: int):
    isos = []
    n = len(exons)
    for _ in range(n_iso):
        if n <= 2:
            keep = list(range(n))
        else:
            keep = [i for i in range(n) if i in (0, n - 1) or rng.random() < 0.75]
            if rng.random() < 0.25:  # alternative first / last exon
                keep = keep[1:] if rng.random() < 0.5 else keep[:-1]
            if not keep:
                keep = [0]
        iso = []
        for i in keep:
            a_s, a_e = alt[i]
            s = a_s[int(rng.integers(0, len(a_s)))]
            e = a_e[int(rng.integers(0, len(a_e)))]
            if e - s < 30:
                s, e = exons[i]
            iso.append((s, e))
        isos.append(iso)
    return isos


# --------------------------------------------------------------------------------------------
# reads
# --------------------------------------------------------------------------------------------
def _cigar_for(rng: np.random.Generator, tlen: int, indel_p: float):
    """CIGAR ops [(len, op)] consuming exactly ``tlen`` target bases; returns (ops, qlen)."""
    if tlen < 12 or rng.random() >= indel_p:
        return [(tlen, "M")], tlen
    ops = []
    q = 0
    left = tlen
    n_ev = int(rng.integers(1, 4))
    for _ in range(n_ev):
        if left < 8:
            break
        m = int(rng.integers(2, max(3, left // 2)))
        ops.append((m, "M" if rng.random() < 0.8 else ("=" if rng.random() < 0.5 else "X")))
        q += m
        left -= m
        if rng.random() < 0.5:
            k = int(rng.integers(1, 6))
            ops.append((k, "I"))
            q += k
        else:
            k = int(rng.integers(1, min(15, left - 1) + 1)) if left > 2 else 0
            if k > 0:
                ops.append((k, "D"))
                left -= k
    if left > 0:
        ops.append((left, "M"))
        q += left
    else:  # a CIGAR must not end in D (split trims them, freddie_split.py:113-128)
        while ops and ops[-1][1] == "D":
            left += ops.pop()[0]
        ops.append((left, "M"))
        q += left
    return ops, q


def _poly(rng: np.random.Generator, n: int, ch: int, err: float) -> np.ndarray:
    out = np.full(n, ch, dtype=np.uint8)
    if n and err > 0:
        bad = rng.random(n) < err
        out[bad] = _BASES[rng.integers(0, 4, size=int(bad.sum()))]
    return out


def _make_read(rng, rid, contig, tint_id, iso, opts, name_prefix):
    """One read: target intervals from an isoform with ONT-like noise + query side."""
    ivs = [list(x) for x in iso]
    # 5' truncation (drop leading exons, start inside one)
    if len(ivs) > 1 and rng.random() < opts["trunc5"]:
        k = int(rng.integers(0, len(ivs)))
        ivs = ivs[k:]
        s, e = ivs[0]
        if e - s > 40:
            ivs[0][0] = s + int(rng.integers(0, e - s - 30))
    # gross 3' truncation
    if len(ivs) > 1 and rng.random() < opts["trunc3"]:
        k = int(rng.integers(1, len(ivs) + 1))
        ivs = ivs[:k]
        s, e = ivs[-1]
        if e - s > 40:
            ivs[-1][1] = e - int(rng.integers(0, e - s - 30))
    # ragged first start / last end
    ivs[0][0] += int(rng.integers(-opts["ragged"], opts["ragged"] + 1))
    ivs[-1][1] += int(rng.integers(-opts["ragged"], opts["ragged"] + 1))
    # per-boundary jitter
    for iv in ivs:
        if rng.random() < opts["jitter_p"]:
            iv[0] += int(rng.integers(-opts["jitter"], opts["jitter"] + 1))
        if rng.random() < opts["jitter_p"]:
            iv[1] += int(rng.integers(-opts["jitter"], opts["jitter"] + 1))
    # spurious >20bp deletion splitting an interval
    out = []
    for s, e in ivs:
        if e - s > 140 and rng.random() < opts["split_p"]:
            d = int(rng.integers(21, 61))
            c = int(rng.integers(s + 30, e - 30 - d))
            out.append([s, c])
            out.append([c + d, e])
        else:
            out.append([s, e])
    # extra noise endpoints (cfg-3 style: raises the candidate count)
    if opts["noise_p"] > 0:
        out2 = []
        for s, e in out:
            while e - s > 120 and rng.random() < opts["noise_p"]:
                d = int(rng.integers(21, 40))
                c = int(rng.integers(s + 30, e - 30 - d))
                out2.append([s, c])
                s = c + d
            out2.append([s, e])
        out = out2
    # sanitise: strictly increasing, non-empty, separated
    clean = []
    last_e = None
    for s, e in out:
        if last_e is not None and s < last_e + 1:
            s = last_e + 1
        if e - s < 2:
            continue
        clean.append((s, e))
        last_e = e
    if not clean:
        s, e = iso[0]
        clean = [(s, max(e, s + 2))]
    # query side
    strand = "+" if rng.random() < 0.5 else "-"
    clip5 = int(rng.integers(0, opts["clip"] + 1))
    clip3 = int(rng.integers(0, opts["clip"] + 1))
    # poly tails: poly-A at the 3' end of '+' reads, poly-T at the 5' end of '-' reads;
    # a fraction is swapped to exercise every branch of the scan
    tail_n = int(rng.integers(0, opts["polya"] + 1)) if rng.random() < opts["polya_p"] else 0
    tail_at_end = strand == "+"
    tail_ch = ord("A") if strand == "+" else ord("T")
    if rng.random() < 0.1:
        tail_at_end = not tail_at_end
    if rng.random() < 0.05:
        tail_ch = ord("T") if tail_ch == ord("A") else ord("A")
    q = clip5 + (0 if tail_at_end else tail_n)
    ivals = []
    for idx, (s, e) in enumerate(clean):
        ops, qlen = _cigar_for(rng, e - s, opts["indel_p"])
        if idx > 0 and rng.random() < opts["qgap_p"]:
            q += int(rng.integers(1, 31))
        ivals.append((s, e, q, q + qlen, ops))
        q += qlen
    length = q + clip3 + (tail_n if tail_at_end else 0)
    seq = _BASES[rng.integers(0, 4, size=length)]
    if tail_n:
        tail = _poly(rng, tail_n, tail_ch, 0.04)
        if tail_at_end:
            gap = int(rng.integers(0, min(6, clip3) + 1))
            st = ivals[-1][3] + gap
            seq[st:st + tail_n] = tail[: max(0, min(tail_n, length - st))]
        else:
            gap = int(rng.integers(0, min(6, clip5) + 1))
            st = ivals[0][2] - gap - tail_n
            st = max(0, st)
            seq[st:st + tail_n] = tail
    return dict(
        id=rid,
        name="%s%d" % (name_prefix, rid),
        chr=contig,
        strand=strand,
        tint=tint_id,
        intervals=ivals,
        seq=seq.tobytes().decode("ascii"),
        length=length,
    )


_DEFAULT_OPTS = dict(
    trunc5=0.30, trunc3=0.025, ragged=25, jitter_p=0.15, jitter=12, split_p=0.03, noise_p=0.0,
    clip=40, polya=40, polya_p=0.7, indel_p=0.3, qgap_p=0.04, dup_p=0.0,
)


def merge_islands(reads: Sequence[dict]) -> List[Tuple[int, int]]:
    """Islands exactly as split merges them (``py/freddie_split.py:297-320``): a new island
    starts only when ``s > end`` (touching intervals merge)."""
    ivs = sorted((iv[0], iv[1]) for r in reads for iv in r["intervals"])
    islands = []
    start = end = None
    for s, e in ivs:
        if start is None:
            start, end = s, e
        if s > end:
            islands.append((start, end))
            start, end = s, e
        end = max(end, e)
    if start is not None:
        islands.append((start, end))
    return islands


def make_tint(
    rng: np.random.Generator,
    contig: str,
    tint_id: int,
    n_reads: int,
    *,
    base: int = 100000,
    locus_len: int = 40000,
    n_exons: int = 10,
    n_iso: int = 7,
    rid0: int = 0,
    opts: Optional[dict] = None,
) -> dict:
    """A tint dict with ``reads`` (each carrying ``seq``/``length``) and header fields."""
    o = dict(_DEFAULT_OPTS)
    if opts:
        o.update(opts)
    exons, alt = _gene_model(rng, base, locus_len, n_exons)
    isos = _isoforms(rng, exons, alt, n_iso)
    iso_p = rng.dirichlet(np.ones(len(isos)) * 1.5)
    reads: List[dict] = []
    prefix = "r%d_" % tint_id
    for k in range(n_reads):
        iso = isos[int(rng.choice(len(isos), p=iso_p))]
        rd = _make_read(rng, rid0 + k, contig, tint_id, iso, o, prefix)
        if reads and o["dup_p"] > 0 and rng.random() < o["dup_p"]:
            # copy the TARGET intervals of an earlier read (=> read rep with weight > 1),
            # with plain-M CIGARs so query coordinates stay consistent
            src = reads[int(rng.integers(0, len(reads)))]
            q = int(rng.integers(0, o["clip"] + 1))
            ivals = []
            for (s, e, _, _, _) in src["intervals"]:
                ivals.append((s, e, q, q + (e - s), [(e - s, "M")]))
                q += e - s
            length = q + int(rng.integers(0, o["clip"] + 1))
            seq = _BASES[rng.integers(0, 4, size=length)].tobytes().decode("ascii")
            rd = dict(rd, intervals=ivals, seq=seq, length=length)
        reads.append(rd)
    return dict(
        id=tint_id, chr=contig, intervals=merge_islands(reads), read_count=len(reads), reads=reads,
    )


# --------------------------------------------------------------------------------------------
# SPLIT directory writer (grammar of freddie_split.py:445-481)
# --------------------------------------------------------------------------------------------
def interval_field(iv) -> str:
    return "%d-%d:%d-%d:%s" % (iv[0], iv[1], iv[2], iv[3], "".join("%d%s" % (c, t) for c, t in iv[4]))


def write_split_dir(tints: Sequence[dict], split_dir: str) -> None:
    for t in tints:
        d = os.path.join(split_dir, t["chr"])
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "split_%s_%d.tsv" % (t["chr"], t["id"])), "w") as f:
            f.write("#%s\t%d\t%s\t%d\n" % (
                t["chr"], t["id"], ",".join("%d-%d" % x for x in t["intervals"]), len(t["reads"])))
            for r in t["reads"]:
                f.write("\t".join([str(r["id"]), r["name"], r["chr"], r["strand"], str(r["tint"])]
                                  + [interval_field(iv) for iv in r["intervals"]]) + "\n")
        with open(os.path.join(d, "reads_%s_%d.tsv" % (t["chr"], t["id"])), "w") as f:
            for r in t["reads"]:
                f.write("%d\t%s\t%d\t%s\n" % (r["id"], t["chr"], t["id"], r["seq"]))


# --------------------------------------------------------------------------------------------
# BASELINE.json configs
# --------------------------------------------------------------------------------------------
def _lognormal_sizes(rng, n_tints, median, sigma, lo, hi):
    x = np.exp(rng.normal(math.log(median), sigma, size=n_tints))
    return np.clip(x.astype(np.int64), lo, hi)


def _powerlaw_sizes(rng, total, alpha, lo, hi):
    sizes = []
    s = 0
    a1 = 1.0 - alpha
    while s < total:
        u = rng.random()
        x = ((hi ** a1 - lo ** a1) * u + lo ** a1) ** (1.0 / a1)
        x = int(max(lo, min(hi, x)))
        sizes.append(x)
        s += x
    return np.array(sizes, dtype=np.int64)


def config_plan(cfg: int, scale: float = 1.0, seed: Optional[int] = None) -> dict:
    """Tint-size plan of config ``cfg`` (1..5).  ``scale`` < 1 shrinks the read total (tests)."""
    seed = cfg if seed is None else seed
    rng = np.random.default_rng(seed)
    if cfg == 1:
        sizes = np.array([max(3, int(2000 * scale))])
        kind = ["locus40k"]
    elif cfg == 2:
        n_tints = max(1, int(round(3000 * scale)))
        sizes = _lognormal_sizes(rng, n_tints, 30, 1.27, 3, 1500)
        kind = ["typical"] * n_tints
    elif cfg == 3:
        n_tints = max(1, int(round(20 * min(1.0, scale * 4))))
        sizes = np.full(n_tints, max(50, int(100000 * scale)))
        kind = ["giant"] * n_tints
    elif cfg == 4:
        hi = max(10, int(200000 * scale))
        sizes = _powerlaw_sizes(rng, int(2_000_000 * scale), 1.6, 1, hi)
        sizes[0] = hi
        kind = ["giant" if s > 20000 else "typical" for s in sizes]
    elif cfg == 5:
        n_tints = max(1, int(round(60000 * scale)))
        sizes = _lognormal_sizes(rng, n_tints, 30, 1.27, 3, 1500)
        n_tail = max(1, n_tints // 400)
        tail = _powerlaw_sizes(rng, int(6_000_000 * scale), 1.6, 1500, max(2000, int(200000 * scale)))
        sizes = np.concatenate([sizes, tail[: max(n_tail, len(tail))]])
        kind = ["giant" if s > 20000 else "typical" for s in sizes]
    else:
        raise ValueError("cfg must be 1..5")
    return dict(cfg=cfg, seed=seed, sizes=sizes, kind=kind)


_TINT_SPACING = 300000  # genomic distance between tint loci on a contig (loci are <= 120 kb)
_TINTS_PER_CONTIG = 2500


def _tint_job(job):
    """Builds tint ``i`` of a plan from its own child seed (so tints can be generated in parallel)."""
    seed, cfg, i, n, kind, contig, slot, rid0 = job
    rng = np.random.default_rng([seed, 7919, i])
    base = 100000 + slot * _TINT_SPACING
    if kind == "locus40k":
        kw = dict(locus_len=40000, n_exons=int(rng.integers(8, 13)), n_iso=int(rng.integers(6, 9)))
        opts = None
    elif kind == "giant":
        kw = dict(locus_len=int(rng.integers(60000, 120000)), n_exons=int(rng.integers(30, 41)),
                  n_iso=int(rng.integers(8, 16)))
        opts = dict(noise_p=0.55, jitter_p=0.5, jitter=30, ragged=60, split_p=0.1)
    else:
        kw = dict(locus_len=int(rng.integers(2000, 100000)), n_exons=int(rng.integers(2, 26)),
                  n_iso=int(rng.integers(1, 9)))
        opts = None
    if cfg == 3 and i % 4 == 3:  # heavy-duplication variant (weighted read reps)
        opts = dict(opts or {}, dup_p=0.9)
    return make_tint(rng, contig, i, n, base=base, rid0=rid0, opts=opts, **kw)


def config_jobs(cfg: int, scale: float = 1.0, seed: Optional[int] = None) -> list:
    """Per-tint generation jobs of a config (each tint has its own child seed)."""
    plan = config_plan(cfg, scale, seed)
    sizes = [int(n) for n in plan["sizes"]]
    n_contigs = max(1, (len(sizes) + _TINTS_PER_CONTIG - 1) // _TINTS_PER_CONTIG)
    if cfg == 5:
        n_contigs = max(n_contigs, 22)
    jobs = []
    rid = 0
    for i, (n, kind) in enumerate(zip(sizes, plan["kind"])):
        jobs.append((plan["seed"], cfg, i, n, kind, "chr%d" % (1 + i % n_contigs), i // n_contigs, rid))
        rid += n
    return jobs


def iter_config(cfg: int, scale: float = 1.0, seed: Optional[int] = None, workers: int = 1, chunk_reads: int = 200000,
                limit: Optional[int] = None, select: Optional[Sequence[int]] = None):
    """Streams a config as lists of tints of about ``chunk_reads`` reads (same tints, same order as
    ``make_config``), so that whole-transcriptome configs never have to sit in host memory at once.
    ``limit``: only the first ``limit`` tints (every tint has its own child seed, so a prefix of a
    config is byte-identical to the same tints of the full config).  ``select``: only the tints with
    these indices (in the given order)."""
    jobs = config_jobs(cfg, scale, seed)
    if limit is not None:
        jobs = jobs[:limit]
    if select is not None:
        jobs = [jobs[i] for i in select]
    pool = None
    if workers > 1 and len(jobs) > 1:
        from multiprocessing import Pool
        pool = Pool(workers)
    try:
        i = 0
        while i < len(jobs):
            j, acc = i, 0
            while j < len(jobs) and (j == i or acc + jobs[j][3] <= chunk_reads):
                acc += jobs[j][3]
                j += 1
            part = jobs[i:j]
            # largest tints first inside the chunk so the pool stays busy
            if pool is not None and len(part) > 1:
                order = sorted(range(len(part)), key=lambda k: -part[k][3])
                res = pool.map(_tint_job, [part[k] for k in order], chunksize=1 if len(part) < 64 else 8)
                out = [None] * len(part)
                for k, r in zip(order, res):
                    out[k] = r
                yield out
            else:
                yield [_tint_job(x) for x in part]
            i = j
    finally:
        if pool is not None:
            pool.close()
            pool.join()


def run_jobs(jobs: Sequence[tuple], workers: int = 1) -> List[dict]:
    """The tints of ``jobs`` (``config_jobs`` entries), generated by a process pool, in job order."""
    if workers > 1 and len(jobs) > 1:
        from multiprocessing import Pool
        order = sorted(range(len(jobs)), key=lambda k: -jobs[k][3])
        with Pool(workers) as p:
            res = p.map(_tint_job, [jobs[k] for k in order], chunksize=1 if len(jobs) < 256 else 8)
        out = [None] * len(jobs)
        for k, r in zip(order, res):
            out[k] = r
        return out
    return [_tint_job(j) for j in jobs]


def _write_job(arg):
    job, split_dir = arg
    t = _tint_job(job)
    write_split_dir([t], split_dir)
    return t["chr"], t["id"], len(t["reads"])


def write_jobs(jobs: Sequence[tuple], split_dir: str, workers: int = 1) -> List[Tuple[str, int, int]]:
    """Generates the tints of ``jobs`` (``config_jobs`` entries) and writes their SPLIT files from the worker
    processes (largest first), so that neither the generation nor the text formatting of a giant tint
    serialises the others.  Returns ``(contig, tint id, reads)`` per job, in job order."""
    order = sorted(range(len(jobs)), key=lambda k: -jobs[k][3])
    args = [(jobs[k], split_dir) for k in order]
    os.makedirs(split_dir, exist_ok=True)
    if workers > 1 and len(jobs) > 1:
        from multiprocessing import Pool
        with Pool(workers) as p:
            res = p.map(_write_job, args, chunksize=1)
    else:
        res = [_write_job(a) for a in args]
    out = [None] * len(jobs)
    for k, r in zip(order, res):
        out[k] = r
    return out


def make_config(cfg: int, scale: float = 1.0, seed: Optional[int] = None, workers: int = 1) -> List[dict]:
    """Realise a config as a list of tints (seeded, deterministic, independent of ``workers``)."""
    plan = config_plan(cfg, scale, seed)
    sizes = [int(n) for n in plan["sizes"]]
    n_contigs = max(1, (len(sizes) + _TINTS_PER_CONTIG - 1) // _TINTS_PER_CONTIG)
    if cfg == 5:
        n_contigs = max(n_contigs, 22)
    jobs = []
    rid = 0
    for i, (n, kind) in enumerate(zip(sizes, plan["kind"])):
        jobs.append((plan["seed"], cfg, i, n, kind, "chr%d" % (1 + i % n_contigs), i // n_contigs, rid))
        rid += n
    if workers > 1 and len(jobs) > 1:
        from multiprocessing import Pool
        with Pool(workers) as p:
            return p.map(_tint_job, jobs, chunksize=max(1, len(jobs) // (workers * 8)))
    return [_tint_job(j) for j in jobs]


def describe(tints: Sequence[dict]) -> Dict[str, int]:
    n = sum(len(t["reads"]) for t in tints)
    iv = sum(len(r["intervals"]) for t in tints for r in t["reads"])
    L = sum(e - s + 1 for t in tints for (s, e) in t["intervals"])
    return dict(tints=len(tints), reads=n, intervals=iv, positions=L,
                islands=sum(len(t["intervals"]) for t in tints))


# --------------------------------------------------------------------------------------------
# hand-built degenerate tints and named golden sets
# --------------------------------------------------------------------------------------------
def _simple_read(rid, contig, tint_id, strand, ivs, clip5, clip3, seq_fill="C", tail=""):
    q = clip5
    ivals = []
    for s, e in ivs:
        ivals.append((s, e, q, q + (e - s), [(e - s, "M")]))
        q += e - s
    body = (seq_fill * (q + clip3))
    seq = body + tail
    return dict(id=rid, name="d%d" % rid, chr=contig, strand=strand, tint=tint_id, intervals=ivals,
                seq=seq, length=len(seq))


def make_degenerate(contig: str = "chrD") -> List[dict]:
    """Edge tints: one spliced read; one unspliced read (empty signal => NaN threshold); three
    identical reads with a poly-A tail (one rep, weight 3); a 2-sample island; a read with no '1'."""
    tints = []

    def add(reads):
        tid = len(tints)
        for r in reads:
            r["tint"] = tid
        tints.append(dict(id=tid, chr=contig, intervals=merge_islands(reads), read_count=len(reads),
                          reads=reads))

    add([_simple_read(0, contig, 0, "+", [(1000, 1200), (1500, 1800)], 5, 7)])
    add([_simple_read(1, contig, 0, "-", [(5000, 5400)], 3, 0)])
    add([_simple_read(2 + k, contig, 0, "+", [(9000, 9300), (9600, 9900)], 5, 0, "G", "A" * 30)
         for k in range(3)])
    add([_simple_read(5, contig, 0, "+", [(20000, 20001), (20100, 20400)], 2, 2),
         _simple_read(6, contig, 0, "-", [(20000, 20001), (20100, 20390)], 0, 4, "T"),
         _simple_read(7, contig, 0, "+", [(20100, 20400)], 1, 1)])
    add([_simple_read(8, contig, 0, "+", [(30000, 31000)], 4, 4),
         _simple_read(9, contig, 0, "-", [(30000, 31000)], 0, 0),
         _simple_read(10, contig, 0, "+", [(30400, 30450)], 9, 2, "A")])
    return tints


def make_plateau_tint(contig: str = "chrP", tint_id: int = 0) -> dict:
    """Equal-weight splice sites an odd distance apart: the smoothed signal has bit-equal plateau
    samples, so the candidate step must pick the floor midpoint."""
    reads = []
    rid = 0
    for rep, (a, b) in enumerate([(1300, 1607), (1300, 1616), (1309, 1607), (1309, 1616)]):
        for _ in range(6):
            reads.append(_simple_read(rid, contig, tint_id, "+" if rid % 2 else "-",
                                      [(1000, a), (b, 2000), (2300, 2600)], 3, 5,
                                      "ACGT"[rid % 4]))
            rid += 1
    return dict(id=tint_id, chr=contig, intervals=merge_islands(reads), read_count=len(reads), reads=reads)


def make_refine_tie_tint(contig: str = "chrR", tint_id: int = 0) -> dict:
    """Equal-height refine peaks closer than find_peaks' distance of 20 (refine_segmentation,
    freddie_segment.py:249-266): three groups of 30 identical reads put splice sites of equal
    weight 12 apart -- a pair, a triple and a quadruple.  With ``-vf 9.9 -lo 100000`` nothing is fixed and
    the DP keeps no interior candidate, so the whole island is one segment and refine decides alone; which
    of the bit-equal peaks survive depends on the visiting order among equal priorities."""
    reads = []
    rid = 0
    groups = [
        [(1000, 1400), (1412, 2400)],                              # pair: 1400 | 1412
        [(1000, 1700), (1712, 1724), (2024, 2400)],                # triple: 1700 | 1712 | 1724  (+ 2024 alone)
        [(1000, 1100), (1112, 1124), (1136, 2400)],                # quadruple: 1100 | 1112 | 1124 | 1136
    ]
    for ivs in groups:
        for _ in range(30):
            reads.append(_simple_read(rid, contig, tint_id, "+" if rid % 2 else "-", ivs, 3, 5, "ACGT"[rid % 4]))
            rid += 1
    return dict(id=tint_id, chr=contig, intervals=merge_islands(reads), read_count=len(reads), reads=reads)


GOLDEN_SETS = {
    # name: (builder kwargs, CLI flags)
    "cfg1": (dict(cfg=1), []),
    "cfg2_small": (dict(cfg=2, scale=0.02), []),
    "cfg2_flagsA": (dict(cfg=2, scale=0.004, seed=12),
                    ["-sd", "2.5", "-tp", "0.8", "-vf", "1.5", "-mps", "12", "-lo", "1", "--consider-ends"]),
    "cfg2_flagsB": (dict(cfg=2, scale=0.004, seed=13), ["-sd", "12", "-tp", "1.0", "-vf", "0.5", "-lo", "0"]),
    "cfg3_mini": (dict(cfg=3, scale=0.03), []),
    "cfg4_mini": (dict(cfg=4, scale=0.002), []),
    "cfg5_mini": (dict(cfg=5, scale=0.0005), []),
    "cfg2_sigma50": (dict(cfg=2, scale=0.003, seed=14), ["-sd", "50"]),
    "cfg2_mps11": (dict(cfg=2, scale=0.004, seed=15), ["-mps", "11", "-vf", "9.5"]),
    "dup_heavy": (dict(special="dup_heavy"), []),
    "degenerate": (dict(special="degenerate"), []),
    "plateau": (dict(special="plateau"), []),
    "refine_tie": (dict(special="refine_tie"), ["-vf", "9.9", "-lo", "100000"]),
    "cfg2_mps9": (dict(cfg=2, scale=0.004, seed=16), ["-mps", "9", "-vf", "9.5"]),
    "empty_tint": (dict(special="empty_tint"), []),
}


def make_golden_set(name: str) -> Tuple[List[dict], List[str]]:
    kw, flags = GOLDEN_SETS[name]
    if kw.get("special") == "degenerate":
        return make_degenerate(), flags
    if kw.get("special") == "dup_heavy":
        rng = np.random.default_rng([99, 1])
        return [make_tint(rng, "chrW", 0, 1500, locus_len=30000, n_exons=9, n_iso=5,
                          opts=dict(dup_p=0.9, jitter_p=0.3)),
                make_tint(rng, "chrW", 1, 400, base=500000, locus_len=20000, n_exons=6, n_iso=3,
                          opts=dict(dup_p=0.97, noise_p=0.3))], flags
    if kw.get("special") == "plateau":
        return [make_plateau_tint()], flags
    if kw.get("special") == "refine_tie":
        return [make_refine_tie_tint()], flags
    if kw.get("special") == "empty_tint":
        # a tint without reads between two ordinary ones: the reference writes a header-only SEGMENT file
        # whose positions are the ends of the islands (NaN threshold, no candidates but the ends)
        return [make_plateau_tint("chrE", 0),
                dict(id=1, chr="chrE", intervals=[(5000, 5400), (6000, 6100)], read_count=0, reads=[]),
                make_refine_tie_tint("chrE", 2)], flags
    return make_config(**kw), flags
