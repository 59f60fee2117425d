#!/usr/bin/env python3
"""Benchmark of the segment stage hot path (contract: see the task's bench.py section).

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port on host cores)

A *step* = one pass of the whole hot path (signal -> ... -> gaps) over one batch of synthetic SPLIT
data.  Workload at every N: BASELINE.json configs[1] -- "synthetic chromosome-scale SPLIT: 200k reads
across ~3k tints" -- per GPU (weak scaling: rank r segments its own seeded realisation; tints never
interact, so there is no data-path collective, only the timing reduction).

value  = reads/s with the packed batch already resident in HBM (K x frs_run, CUDA events).
e2e    = reads/s through the C ABI with pinned HOST buffers: frs_upload + frs_run + frs_download.
roofline = dominant kernel of the step, algorithmic bytes / CUDA-event time.
cli    = files to files on a bounded prefix of the workload (SPLIT text -> native parser -> CUDA -> native
         formatter -> SEGMENT files), the scope the reference itself is measured at (rank 0, N = 1 only).
cpu_baseline = the oracle port (numpy restatement of the reference, NOT the product) on all host
cores over a bounded sample of the same workload (rank 0, N = 1 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads/sec through freddie_segment"
UNIT = "reads/s"
KERNEL_OF = {"signal": "k_signal", "smooth": "k_smooth", "lists": "k_tile_lists", "coverage": "k_coverage",
             "refine": "k_refine_filter+k_refine", "digits": "k_digits", "gaps": "k_gap_prep+k_gap_sizes", "dp": "k_dp*"}
WORKLOADS = {
    "cfg2": "BASELINE configs[1]: synthetic chromosome-scale SPLIT, 200k reads across ~3k tints per GPU (one seeded "
            "dataset of N x that size, its tints bin-packed over the N GPUs by estimated cost)",
    "cfg3": "BASELINE configs[2]: DP-dominated giant tints, 20 x 100k reads (x --scale), per GPU",
    "cfg4": "BASELINE configs[3]: power-law tint sizes 1..200k reads, ~2.2 M reads (x --scale): ONE dataset bin-packed "
            "over the GPUs by estimated cost",
    "cfg5": "BASELINE configs[4]: whole-transcriptome scale, 10 M reads across ~60k tints (x --scale): ONE dataset "
            "bin-packed over the GPUs by estimated cost",
}


# ------------------------------------------------------------------------------------------------
def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(workload: str, scale: float, seed: int, workers: int):
    from freddie_b200 import synth
    cfg = {"cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[workload]
    return synth.make_config(cfg, scale=scale, seed=seed, workers=workers)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def algorithmic_bytes(counts: dict, sizes: dict, cov_elems: int) -> dict:
    """Per-stage ALGORITHMIC bytes of one step (DESIGN.md 'Bytes per unit')."""
    L, I, R, N = counts["n_samples"], counts["n_rep_ivs"], counts["n_reps"], counts["n_reads"]
    Ir, K = counts["n_read_ivs"], sizes["n_candidates"]
    dig = sizes["n_digit_bytes"]
    return {
        "signal": 8 * I + 8 * L,
        # k_smooth fuses the Gaussian (a4: 16 L) and the candidate peaks (a6: 8 L) into one pass;
        # k_tile_lists writes the candidate list and gathers the positive samples the variance
        # threshold reads (a5: 8 L)
        "smooth": 24 * L,
        "lists": 8 * L,
        "coverage": 8 * I + 4 * cov_elems,
        "dp": 4 * cov_elems + 4 * R,
        "refine": 8 * L,
        "digits": 8 * I + dig,
        "gaps": 16 * Ir + 4 * N,
    }


# ------------------------------------------------------------------------------------------------
def oracle_rate(tints, threads: int, target_reads: int):
    """Oracle (CPU port of the reference algorithm) on a bounded, size-stratified sample of tints."""
    from multiprocessing import Pool
    sample = stratified_sample(tints, target_reads)
    n = sum(len(t["reads"]) for t in sample)
    t0 = time.perf_counter()
    with Pool(threads) as p:
        res = p.map(_oracle_one, sample, chunksize=1)
    dt = time.perf_counter() - t0
    cells = sum(r[1] for r in res)
    return n / dt, cells / dt, dict(tints=len(sample), reads=n, seconds=round(dt, 2), dp_cells=cells)


def _oracle_one(tint):
    from oracle import segment_oracle as orc
    st = {}
    orc.segment_tint(tint, orc.Params(), stats=st)
    return len(tint["reads"]), st.get("cells", 0)


def stratified_sample(tints, target_reads: int):
    """Every k-th tint of the workload ordered by size: the sample keeps the size mix of the workload."""
    order = sorted(range(len(tints)), key=lambda i: len(tints[i]["reads"]))
    stride = max(1, int(len(order) * (sum(len(t["reads"]) for t in tints) / len(tints)) / max(target_reads, 1)))
    pick = order[stride // 2::stride]
    return [tints[i] for i in pick]


class ReferenceRunner:
    """The UNMODIFIED reference CLI (oracle/_ref, see oracle/build_ref.py) on a SPLIT sample in tmpfs:
    ``freddie_segment.py -s SPLIT -o OUT -t <cores>``, wall clock around the subprocess."""

    def __init__(self, sample, cores):
        import tempfile
        from freddie_b200 import synth
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        self.work = tempfile.mkdtemp(prefix="frs_ref_", dir=base)
        self.split = os.path.join(self.work, "split")
        synth.write_split_dir(sample, self.split)
        self.reads = sum(len(t["reads"]) for t in sample)
        self.tints = len(sample)
        self.cores = cores

    def run(self):
        import shutil
        from oracle import build_ref
        out = os.path.join(self.work, "out")
        shutil.rmtree(out, ignore_errors=True)
        t0 = time.perf_counter()
        r = subprocess.run(build_ref.command(self.split, out, self.cores), stdout=subprocess.DEVNULL,
                           stderr=subprocess.PIPE, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0:
            raise RuntimeError("reference CLI failed: " + r.stderr[-500:])
        n_out = sum(len(fs) for _, _, fs in os.walk(out))
        if n_out != 2 * self.tints:
            raise RuntimeError("reference CLI wrote %d files for %d tints" % (n_out, self.tints))
        return dt

    def close(self):
        import shutil
        shutil.rmtree(self.work, ignore_errors=True)


def reference_rate(tints, cores: int, seconds: float):
    """reads/s of the reference CLI on a size-stratified sample sized for about ``seconds`` of wall clock
    (calibrated by a small first run).  Returns (rate, sample description)."""
    probe = ReferenceRunner(stratified_sample(tints, 250 * cores), cores)
    try:
        dt = probe.run()
        rate = probe.reads / dt
    finally:
        probe.close()
    total = sum(len(t["reads"]) for t in tints)
    run = ReferenceRunner(stratified_sample(tints, int(min(total, max(500, rate * seconds)))), cores)
    try:
        dt = run.run()
        return run.reads / dt, dict(tints=run.tints, reads=run.reads, seconds=round(dt, 2), threads=cores)
    finally:
        run.close()


def _ref_cluster_one(path):
    """read_segment + preprocess_ilp + partition_reads of the UNMODIFIED reference on one SEGMENT file
    (freddie_cluster.py:119-172, :277-328, :198-274) -> (reads, digest of the canonical serialisation)."""
    import contextlib
    import hashlib
    import io
    from oracle import build_ref
    from oracle import cluster_prep_oracle as cpo
    fc = build_ref.reference_cluster_module()
    tint = list(fc.read_segment(path).values())[0]
    fc.preprocess_ilp(tint, dict(recycle_model="constant"))
    with contextlib.redirect_stdout(io.StringIO()):  # partition_reads prints every piece
        fc.partition_reads(tint, 1000)
    d = tint["ilp_data"]
    U = len(tint["read_reps"])
    gaps = [tint["reads"][tint["read_reps"][i][0]]["gaps"] for i in range(U)]
    cat = [tint["reads"][tint["read_reps"][i][0]]["poly_tail_category"] for i in range(U)]
    s = cpo.canonical(d["I"], d["C"], d["FL"], cat, d["garbage_cost"], gaps, tint["partitions"])
    return len(tint["reads"]), hashlib.sha256(s.encode()).hexdigest()


def cluster_prep_scope(device, batch, res, cores):
    """read-rep merge + preprocess_ilp + partition_reads of every tint of the workload in ONE
    frs_cprep_run (host arrays in, host arrays out), beside the unmodified reference functions on a bounded
    sample of the same tints (all host cores, one tint per task) whose results are compared digest by digest."""
    import hashlib
    import shutil
    import tempfile
    from freddie_b200.cluster_prep import ClusterPrep, batch_from_segment
    from freddie_b200.engine import format_tint
    cb = batch_from_segment(batch.arrays, res.arrays)
    ctx = ClusterPrep(device)
    out = None
    for _ in range(2):
        out = ctx.run(cb, 1000)
    K = 5
    t0 = time.perf_counter()
    for _ in range(K):
        out = ctx.run(cb, 1000)
    dt = (time.perf_counter() - t0) / K
    in_bytes = int(sum(np.asarray(v).nbytes for v in out.batch.values()))
    out_bytes = int(sum(v.nbytes for v in out.a.values()))
    obj = dict(value=batch.n_reads / dt, unit="reads/s", seconds=round(dt, 5),
               scope="host arrays of the segment stage in -> frs_cprep_run + frs_cprep_fetch -> host arrays out (copies and "
                     "host bookkeeping inside the wall clock), maximum_ilp_size 1000",
               device_ms={k: round(v, 4) for k, v in out.timings_ms.items()}, launches=out.sizes["launches"],
               h2d_bytes=in_bytes, d2h_bytes=out_bytes,
               counts={k: out.sizes[k] for k in ("n_reps", "n_structs", "n_parts", "n_incomp", "edges_before", "edges_after", "prune_rounds")},
               reads=batch.n_reads, tints=batch.n_tints)
    from oracle import build_ref
    if build_ref.cluster_available():
        from multiprocessing import Pool
        from oracle import cluster_prep_oracle as cpo
        # bounded sample: every k-th tint by size, cost ~ reads^2 (the reference's pair loop is pure Python)
        tro = np.asarray(batch.arrays["tint_read_off"])
        n = np.diff(tro)
        order = np.argsort(n, kind="stable")
        budget = float(os.environ.get("FRS_CLUSTER_PAIRS", 1.2e6)) * cores
        stride = max(1, int(np.ceil(float((n.astype(np.float64) ** 2).sum()) / budget)))
        pick = [int(t) for t in order[stride // 2::stride] if n[t] > 0]
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        work = tempfile.mkdtemp(prefix="frs_refc_", dir=base)
        try:
            paths = []
            for t in pick:
                paths.append(os.path.join(work, "segment_%d.tsv" % t))
                with open(paths[-1], "w") as fh:
                    fh.write(format_tint(batch, res, t))
            t0 = time.perf_counter()
            with Pool(cores) as p:
                ref = p.map(_ref_cluster_one, paths, chunksize=1)
            rdt = time.perf_counter() - t0
        finally:
            shutil.rmtree(work, ignore_errors=True)
        same = 0
        for t, (_, dg) in zip(pick, ref):
            r = out.tint(t, "constant")
            mine = hashlib.sha256(cpo.canonical(r["I"], r["C"], r["FL"], r["cat"], r["garbage_cost"], r["gaps"],
                                                r["partitions"]).encode()).hexdigest()
            same += mine == dg
        rn = int(sum(x[0] for x in ref))
        obj["cpu_baseline"] = dict(value=rn / rdt, unit="reads/s", cores=cores, kind="reference",
                                   sample="unmodified read_segment + preprocess_ilp + partition_reads (oracle/_ref/freddie_cluster.bin), "
                                          "one tint per task on %d processes, every %d-th tint by size: %d tints, %d reads, %.2f s"
                                          % (cores, stride, len(pick), rn, rdt),
                                   parity="%d of %d sampled tints equal the CUDA path's canonical digest" % (same, len(pick)))
    ctx.close()
    return obj


def _ref_split_one(group):
    """get_transcriptional_intervals (+ break_tint) of the UNMODIFIED reference on one group (freddie_split.py:295-364,
    :246-293) -> digest of the canonical serialisation."""
    from oracle import build_ref
    from oracle import split_tints_oracle as sto
    fs = build_ref.reference_split_module()
    reads = [dict(id=i, name="r%d" % i, contig="c", strand="+", simple_tints=list(), tint=None,
                  intervals=[(s, e, 0, e - s, [(0, e - s)]) for s, e in ivs]) for i, ivs in enumerate(group)]
    return sto.digest_of([(t["intervals"], t["rids"]) for t in fs.get_transcriptional_intervals(reads=reads)])


def split_tints_scope(device, tints, cores):
    """tint construction of the split stage for the reads of the workload, re-grouped the way read_sam
    would hand them over (runs of 8 neighbouring tints of the synthetic contig form one group, reads by start), in
    ONE frs_split_run; beside it the unmodified reference function on the same groups (one group per task, all host
    cores), compared digest by digest."""
    from freddie_b200.split_tints import SplitTints
    groups = []
    for k in range(0, len(tints), 8):
        reads = [[(iv[0], iv[1]) for iv in r["intervals"]] for t in tints[k:k + 8] for r in t["reads"]]
        reads.sort(key=lambda ivs: ivs[0][0])
        groups.append(reads)
    gro, rio, s, e = [0], [0], [], []
    for reads in groups:
        for ivs in reads:
            for a, b in ivs:
                s.append(a)
                e.append(b)
            rio.append(len(s))
        gro.append(len(rio) - 1)
    arrs = [np.asarray(x, np.int32) for x in (gro, rio, s, e)]
    ctx = SplitTints(device)
    for _ in range(2):
        res, info, _ = ctx.run_arrays(*arrs)
    K = 5
    t0 = time.perf_counter()
    for _ in range(K):
        res, info, raw = ctx.run_arrays(*arrs)
    dt = (time.perf_counter() - t0) / K
    n_reads = len(rio) - 1
    obj = dict(value=n_reads / dt, unit="reads/s", seconds=round(dt, 5), device_ms=round(info["device_ms"], 3),
               scope="host arrays of decoded alignments in -> frs_split_run + frs_split_fetch -> tint lists out (copies, host "
                     "assembly and the Python list building of the mirror inside the wall clock)",
               groups=len(groups), reads=n_reads, intervals=len(s), tints=info["n_tints"], simple_intervals=info["n_simple"],
               broken_groups=info["n_big"], launches=info["launches"])
    from oracle import build_ref
    if build_ref.split_available():
        from multiprocessing import Pool
        from oracle import split_tints_oracle as sto
        t0 = time.perf_counter()
        with Pool(cores) as p:
            ref = p.map(_ref_split_one, groups, chunksize=1)
        rdt = time.perf_counter() - t0
        same = sum(sto.digest_of(res[g]) == ref[g] for g in range(len(groups)))
        obj["cpu_baseline"] = dict(value=n_reads / rdt, unit="reads/s", cores=cores, kind="reference",
                                   sample="unmodified get_transcriptional_intervals + break_tint (oracle/_ref/freddie_split.bin), "
                                          "one group per task on %d processes, all %d groups, %.2f s" % (cores, len(groups), rdt),
                                   parity="%d of %d groups equal the CUDA path's canonical digest" % (same, len(groups)))
    ctx.close()
    return obj


def config_dict(args):
    """The same object from both arms (the driver compares them)."""
    return dict(workload=WORKLOADS[args.workload], scale=args.scale, params="defaults (sd=5 tp=0.9 vf=3 mps=50 lo=3)",
                l2="cuda arm: flushed between timed steps (256 MiB memset)")


def run_reference_arm(args):
    """Times the reference's own CPU implementation on the box's host cores: the unmodified
    ``freddie_segment.py`` (``oracle/_ref``) with ``-t <cores>``; each step is one run of the CLI over the
    same bounded, size-stratified sample of the workload (SPLIT files in tmpfs)."""
    rank, _, world = rank_env()
    if rank != 0:
        return
    from oracle import build_ref
    cores = os.cpu_count() or 1
    tints = make_workload(args.workload, args.scale, 2, min(cores, 16))
    total_steps = args.steps + args.warmup
    budget = 170.0 / max(total_steps, 1)  # seconds of wall clock per step
    if build_ref.available():
        probe = ReferenceRunner(stratified_sample(tints, 250 * cores), cores)
        try:
            rate = probe.reads / probe.run()
        finally:
            probe.close()
        total = sum(len(t["reads"]) for t in tints)
        runner = ReferenceRunner(stratified_sample(tints, int(min(total, max(500, rate * budget)))), cores)
        try:
            secs = [runner.run() for _ in range(total_steps)][args.warmup:]
        finally:
            runner.close()
        v = runner.reads / float(np.mean(secs))
        kind = "reference"
        sample = ("unmodified freddie_segment.py (oracle/_ref/freddie_segment.bin, byte-compiled from the reference "
                  "source) -t %d, files to files in tmpfs, on a size-stratified sample of the workload, the same "
                  "sample every step: %d tints, %d reads, %.2f s per run" % (cores, runner.tints, runner.reads, float(np.mean(secs))))
        cells = None
    else:  # oracle/_ref was not built (no /root/reference at build time): time the port and say so
        rates, cell_rates, secs, smp = [], [], [], None
        rate_guess = 300.0 * cores
        for s in range(total_steps):
            target = int(min(sum(len(t["reads"]) for t in tints), max(500, rate_guess * budget)))
            r, c, smp = oracle_rate(tints, cores, target)
            rate_guess = r
            if s >= args.warmup:
                rates.append(r)
                cell_rates.append(c)
                secs.append(smp["seconds"])
        v = float(np.mean(rates))
        kind = "port"
        sample = "oracle/segment_oracle.py (oracle/_ref not built) on a size-stratified sample per step: %s" % smp
        cells = float(np.mean(cell_rates))
    line = dict(
        impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
        ms_per_step=float(np.mean(secs)) * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int32+f64",
        data="synthetic", config=config_dict(args),
        cpu_baseline=dict(value=v, unit=UNIT, cores=cores, kind=kind, sample=sample),
        e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
        dp_cells_per_sec=cells,
    )
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def make_shard(args, rank, world, workers):
    """The tints of THIS rank.  cfg2 (weak scaling): ONE dataset of world x 200 k reads / world x 3 k tints,
    its tints bin-packed over the GPUs by estimated cost (schedule.lpt_partition, longest first), every rank
    generates only its own shard.  cfg4 / cfg5 (strong scaling): one dataset of the named size, sharded the
    same way.  cfg3: every rank its own realisation (20 equal giant tints do not need packing).
    Returns (list of tint lists = the rank's batches, info)."""
    from freddie_b200 import schedule, synth
    cfg = {"cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 5}[args.workload]
    if args.workload == "cfg3":
        tints = synth.make_config(3, scale=args.scale, seed=3 + 1000 * rank, workers=workers)
        return [[t] for t in tints] if args.scale >= 0.5 else [tints], dict(sharding="replicas (own seed per rank)")
    scale = args.scale * (world if args.workload == "cfg2" else 1)
    jobs = synth.config_jobs(cfg, scale=scale)
    costs = [(schedule.estimate_cost(j[3]), float(j[3])) for j in jobs]
    shards = schedule.lpt_partition(costs, world)
    mine = sorted(shards[rank])
    loads = [sum(costs[i][0] for i in sh) for sh in shards]
    info = dict(sharding="one dataset, LPT bin-packing by estimated cost (freddie_b200.schedule)",
                dataset_tints=len(jobs), dataset_reads=int(sum(j[3] for j in jobs)),
                est_cost_imbalance=round(max(loads) / (sum(loads) / len(loads)), 4))
    my_jobs = [jobs[i] for i in mine]
    if args.workload == "cfg2":
        return [synth.run_jobs(my_jobs, workers)], info
    groups = list(schedule.batches(my_jobs, [costs[i] for i in mine], args.batch_reads))
    return groups, info  # job lists: generated and packed group by group (a whole config does not fit in dicts)


def run_cuda_arm(args):
    rank, local_rank, world = rank_env()
    cores = os.cpu_count() or 1
    # a context drives ~12 streams (copy in / out, head, tail, DP classes): give each its own hardware queue
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    # torchrun pins OMP_NUM_THREADS=1 in its children; the native parser uses OpenMP
    os.environ["OMP_NUM_THREADS"] = str(max(1, cores // max(world, 1)))
    # NCCL logs (version banner, INFO lines the driver may ask for with NCCL_DEBUG) go to stderr: stdout is ONE line
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    # The files-to-files scope runs FIRST, before this process touches the GPU: a user's cold `freddie_segment` run
    # has the GPU to itself, while a child started at the end of the bench shared it with this process's contexts,
    # ~10 GB of device buffers and pinned arenas (measured: 2.7 - 7.4 s cold there, 1.6 s for the same command alone).
    cli_early = None
    if world == 1 and not args.no_cli and args.workload == "cfg2":
        cli_early = cli_scope(cores)
    # host threads and pinned buffers of this rank next to its GPU (first-touch allocation follows the CPU affinity)
    from freddie_b200 import affinity
    all_cpus = os.sched_getaffinity(0)
    numa = affinity.bind_to_gpu(local_rank)
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from freddie_b200 import _lib, synth
    from freddie_b200.engine import Engine, SegmentParams
    from freddie_b200.pack import pack_tints

    t0 = time.time()
    workers = max(1, min(32, cores // world))
    groups, shard_info = make_shard(args, rank, world, workers)
    tints_for_cpu = groups[0] if args.workload == "cfg2" else None
    batches = []
    for g in groups:
        tl = g if (g and isinstance(g[0], dict)) else synth.run_jobs(g, workers)
        batches.append(pack_tints(tl).pin())
        if tints_for_cpu is None:
            batches[-1].tints = []  # keep the arrays only
    del groups
    t_gen = time.time() - t0
    n_reads = sum(b.n_reads for b in batches)
    n_tints = sum(b.n_tints for b in batches)
    prm = SegmentParams()
    eng = Engine(local_rank)
    stream = torch.cuda.ExternalStream(eng.lib.frs_stream(eng.ctx), device=torch.device("cuda", local_rank))
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        with torch.cuda.stream(stream):
            flush.zero_()

    # ---- warm-up: full end-to-end steps (also sizes every device buffer) ----
    eng.set_option(_lib.OPT_LAZY_SEQ, 0)  # `value`: every input, sequence planes included, resident in HBM
    res_resident, sizes = [], []
    for w in range(max(args.warmup, 1)):
        res_resident, sizes = [], []
        for b in batches:
            res_resident.append(eng.segment_batch(b, prm, pinned=True))
            sizes.append(dict(res_resident[-1].sizes, cov_elems=int(eng.tap(_lib.TAP_COV_OFF, np.int64)[-1])))

    # ---- timed: K x frs_run per batch with the batch resident in HBM (upload outside the events) ----
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank)
    stage_ms, stage_launch = {}, {}
    ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in batches]
          for _ in range(args.steps)]
    if len(batches) == 1:
        eng.upload(batches[0])
    barrier()
    sampler.start()
    launches = 0
    for k in range(args.steps):
        for i, b in enumerate(batches):
            if len(batches) > 1:
                eng.upload(b)
            flush_l2()
            ev[k][i][0].record(stream)
            eng.run(prm)
            ev[k][i][1].record(stream)
            launches += eng.launch_count()
            for name, ms, ln in eng.timings():
                stage_ms[name] = stage_ms.get(name, 0.0) + ms
                stage_launch[name] = stage_launch.get(name, 0) + ln
    barrier()
    clocks = sampler.stop()
    t_dev = sum(a.elapsed_time(b) for row in ev for a, b in row) / 1e3
    eng.set_profiling(False)
    reruns_resident = eng.stats()["reruns"]
    stage_launch = {k: v // (args.steps * len(batches)) if len(batches) == 1 else v // args.steps for k, v in stage_launch.items()}

    # ---- timed: end to end through the C ABI with pinned HOST buffers, ONE context, one host thread.
    # Every step = frs_submit (host-to-device copies + every kernel, enqueued without a host round trip) +
    # frs_wait + frs_fetch (device-to-host copies into pinned memory); two batches are in flight, so the
    # copies of one step overlap the kernels of its neighbours on the copy engines.  `serial_value` is the
    # same work with the synchronous calls (upload, run, download one after the other). ----
    eng.set_option(_lib.OPT_LAZY_SEQ, 1)  # pinned planes stay on the host; the clip words are fetched by a kernel
    n_ctx = max(1, int(os.environ.get("FRS_E2E_CONTEXTS", "1")))  # more contexts only add contention (measured: 2 -> 0.65x)
    ctxs = [eng] + [Engine(local_rank) for _ in range(n_ctx - 1)]
    pong = []
    for e2 in ctxs:
        for _ in range(max(args.warmup, 1)):
            rr = [e2.segment_batch(b, prm, pinned=True) for b in batches]
        pong.append([[e2.new_result(_sizes_obj(r), b, pinned=True) for r, b in zip(rr, batches)] for _ in range(2)])
    res_lazy = rr
    st = ctxs[0].stats()
    h2d = sum(_upload_bytes(ctxs[0], b, prm) for b in batches)
    d2h = int(sum(v.nbytes for r in res_lazy for v in r.arrays.values())) + len(batches) * 48 * 8

    def pipelined(e2, bufs, work):
        """work: list of batch indices in order.  Up to five batches in flight (copy in | head kernels | tail |
        copy out, one queued behind them); the host thread only ever blocks in frs_wait: the read-back of a batch is started
        (frs_fetch_start) and collected one iteration later (frs_fetch_finish), behind the submit of the next."""
        from collections import deque
        fly = deque()
        pending = None
        for n, i in enumerate(work):
            fly.append((e2.submit(batches[i], prm), i, n))
            if pending is not None:
                e2.fetch_finish(pending)
                pending = None
            if len(fly) == 4:
                pt, pi, pn = fly.popleft()
                e2.wait(pt)
                e2.fetch_start(pt, bufs[pn & 1][pi])
                pending = pt
        while fly:
            if pending is not None:
                e2.fetch_finish(pending)
            pt, pi, pn = fly.popleft()
            e2.wait(pt)
            e2.fetch_start(pt, bufs[pn & 1][pi])
            pending = pt
        if pending is not None:
            e2.fetch_finish(pending)

    def serial(e2, bufs, work):
        for n, i in enumerate(work):
            e2.upload(batches[i])
            e2.run(prm)
            r = bufs[n & 1][i].as_struct()
            e2._check(e2.lib.frs_download(e2.ctx, __import__("ctypes").byref(r)))

    def timed(fn, n_used):
        work_all = [i for _ in range(args.steps) for i in range(len(batches))]
        parts = [work_all[k::n_used] for k in range(n_used)]
        ths = [threading.Thread(target=fn, args=(ctxs[k], pong[k], parts[k])) for k in range(n_used)]
        barrier()
        w0 = time.perf_counter()
        if n_used == 1:
            fn(ctxs[0], pong[0], parts[0])
        else:
            for t in ths:
                t.start()
            for t in ths:
                t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - w0

    timed(pipelined, 1)  # untimed pass: capacities of the lazy mode
    t_serial = timed(serial, 1)
    t_pipe1 = timed(pipelined, 1)
    t_pipe2 = timed(pipelined, n_ctx) if n_ctx > 1 else t_pipe1
    barrier()
    for i in range(len(batches)):
        for k in res_resident[i].arrays:
            if not np.array_equal(pong[0][0][i].arrays[k], res_resident[i].arrays[k]) and \
               not np.array_equal(pong[0][1][i].arrays[k], res_resident[i].arrays[k]):
                raise RuntimeError("end-to-end (pipelined, lazy sequence) results differ from the resident run: %s" % k)
    t_e2e = min(t_pipe1, t_pipe2)

    # ---- reduce over ranks: max time, sum of units; per-rank busy times for the balance of the shard ----
    tot_reads = n_reads
    tot_cells = int(sum(s["dp_cells"] for s in sizes))
    tot_rcells = int(sum(s["dp_read_cells"] for s in sizes))
    dp_ms = stage_ms.get("dp", 0.0)
    busy = [t_dev / args.steps]
    busy_e2e = [t_e2e / args.steps]
    reads_per_rank = [n_reads]
    if world > 1:
        g = torch.zeros(world, 3, device="cuda", dtype=torch.float64)
        g[rank, 0], g[rank, 1], g[rank, 2] = t_dev / args.steps, t_e2e / args.steps, n_reads
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        busy, busy_e2e, reads_per_rank = g[:, 0].tolist(), g[:, 1].tolist(), [int(x) for x in g[:, 2].tolist()]
        t = torch.tensor([t_dev, t_e2e, dp_ms, t_serial, t_pipe1, t_pipe2], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, dp_ms, t_serial, t_pipe1, t_pipe2 = [float(x) for x in t.tolist()]
        u = torch.tensor([n_reads, tot_cells, launches, h2d, d2h, tot_rcells, n_tints], device="cuda", dtype=torch.int64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        tot_reads, tot_cells, launches, h2d, d2h, tot_rcells, n_tints = [int(x) for x in u.tolist()]  # whole-job totals
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = tot_reads * args.steps / t_dev
    e2e_v = tot_reads * args.steps / t_e2e
    peak, peak_src = peaks()
    alg = {}
    for b, sz, r in zip(batches, sizes, res_resident):
        cov = int(sz.get("cov_elems", 0))
        for k, v in algorithmic_bytes(b.counts(), sz, cov).items():
            alg[k] = alg.get(k, 0) + v
    # dominant KERNEL of the step = the longest single launch.  Stages that are one launch on the context
    # stream are timed exactly by their CUDA events; the DP stage is several persistent kernels running
    # concurrently on side streams and is issue-bound, not HBM-bound: it is reported in `dp_stage`.
    per_step = lambda k: stage_ms[k] / args.steps  # noqa: E731
    single = [k for k in stage_ms if k in alg and stage_launch.get(k) == len(batches)]
    dom = max(single or [k for k in stage_ms if k in alg], key=lambda k: stage_ms[k])
    dom_ms = per_step(dom)
    ach = alg[dom] / (dom_ms * 1e-3) / 1e9
    stages = {k: dict(ms=round(per_step(k), 4), launches=stage_launch[k],
                      alg_GBps=(round(alg[k] / (per_step(k) * 1e-3) / 1e9, 1) if k in alg and v > 0 else None),
                      hbm_frac=(round(alg[k] / (per_step(k) * 1e-3) / 1e9 / peak, 4) if k in alg and v > 0 else None))
              for k, v in stage_ms.items()}
    stream_ms = sum(per_step(k) for k in stage_ms if k in alg and k != "dp")
    stream_bytes = sum(bb for k, bb in alg.items() if k != "dp" and k in stage_ms)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "cfg2" and args.scale == 1.0:
        traffic = json.load(open(tp)).get(dom)
    # DP: read-cell updates against the POPC-limited ceiling: one POPC covers the predicate
    # lanes of 32 read reps; 16 POPC / clk / SM (the quarter-rate integer pipe of sm_100, B300_MICROARCH.md)
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    popc_ceiling = 16.0 * 32.0 * 148.0 * sm_mhz * 1e6 * world
    rcu = tot_rcells * args.steps / max(dp_ms * 1e-3, 1e-12)
    mean = lambda xs: sum(xs) / len(xs)  # noqa: E731
    cfg = config_dict(args)
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=t_dev / args.steps * 1e3, higher_is_better=True,
        scaling="weak" if args.workload in ("cfg2", "cfg3") else "strong", vs_baseline=None,
        dtype="int32+f64", data="synthetic", config=cfg,
        workload_detail=dict(reads=tot_reads, tints=n_tints, batches_per_gpu=len(batches), reads_per_gpu=reads_per_rank,
                             **shard_info),
        balance=dict(device_ms_per_gpu=[round(x * 1e3, 4) for x in busy], device_max_over_mean=round(max(busy) / mean(busy), 4),
                     e2e_ms_per_gpu=[round(x * 1e3, 4) for x in busy_e2e], e2e_max_over_mean=round(max(busy_e2e) / mean(busy_e2e), 4)),
        e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 timing="wall clock around K steps, synchronize on both sides; frs_submit / frs_wait / frs_fetch with pinned "
                        "host buffers, up to five batches in flight per context (copy in | head kernels | tail | copy out, one queued)",
                 one_context_value=tot_reads * args.steps / t_pipe1,
                 contexts_value=tot_reads * args.steps / t_pipe2, contexts=n_ctx,
                 serial_value=tot_reads * args.steps / t_serial,
                 clip_words_per_step=st["clip_words"], seq_words_in_batch=st["seq_words"],
                 reruns=sum(e2.stats()["reruns"] for e2 in ctxs)),
        gpu_launches=launches,
        clocks=clocks,
        roofline=dict(bound="hbm", kernel=KERNEL_OF.get(dom, dom), stage=dom, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                      peak_source=peak_src, alg_bytes_per_launch=alg[dom] / len(batches), ms_per_launch=dom_ms / len(batches),
                      note="longest single launch of the step, timed by its own CUDA events; the concurrent DP kernels "
                           "are issue-bound and reported in dp_stage",
                      streaming_stages=dict(achieved=round(stream_bytes / (stream_ms * 1e-3) / 1e9, 1),
                                            frac=round(stream_bytes / (stream_ms * 1e-3) / 1e9 / peak, 4),
                                            ms=round(stream_ms, 4), alg_bytes=stream_bytes,
                                            what="all HBM-streaming stages together (every stage but dp)")),
        dp_stage=dict(bound="issue (VOTE/LOP3/POPC), not HBM: coverage rows are read once (TMA) and reused on chip",
                      ms=round(dp_ms / args.steps, 4), launches=stage_launch.get("dp"),
                      alg_GBps=round(alg["dp"] / (dp_ms / args.steps * 1e-3) / 1e9, 1) if dp_ms > 0 else None,
                      rcu_per_sec=rcu, ceiling=popc_ceiling, frac=rcu / popc_ceiling,
                      ceiling_how="16 POPC/clk/SM x 32 read reps per word x 148 SMs x %.0f MHz (median SM clock under load) x %d GPUs" % (sm_mhz, world),
                      evidence="profiles/: issue-slot utilisation and stall breakdown of every DP kernel"),
        dp_cells_per_sec=tot_cells * args.steps / max(dp_ms * 1e-3, 1e-12),
        dp_read_cells_per_sec=rcu,
        dp=dict(cells=tot_cells, subproblems=int(sum(s["n_subproblems"] for s in sizes)),
                max_n=int(max(s["max_subproblem"] for s in sizes)), candidates=int(sum(s["n_candidates"] for s in sizes))),
        stages=stages,
        setup_seconds=round(t_gen, 1),
    )
    line["numa"] = numa
    os.sched_setaffinity(0, all_cpus)  # the CPU arms below use every host core
    if cli_early is not None:
        line["cli"] = cli_early
    if world == 1 and tints_for_cpu is not None and len(batches) == 1 and not os.environ.get("FRS_NO_CLUSTER_PREP"):
        try:  # the next row of the scope table; never costs the bench line
            line["cluster_prep"] = cluster_prep_scope(local_rank, batches[0], res_lazy[0], cores)
        except Exception as e:  # noqa: BLE001
            line["cluster_prep"] = dict(error=repr(e)[:300])
        try:  # the last row of the scope table
            line["split_tints"] = split_tints_scope(local_rank, tints_for_cpu, cores)
        except Exception as e:  # noqa: BLE001
            line["split_tints"] = dict(error=repr(e)[:300])
    if world == 1 and not args.no_cpu_baseline and tints_for_cpu is not None:
        from oracle import build_ref
        r, c, sample = oracle_rate(tints_for_cpu, cores, int(os.environ.get("FRS_CPU_SAMPLE_READS", 300 * cores * 8)))
        port = dict(value=r, unit=UNIT, cores=cores, kind="port", dp_cells_per_sec=c,
                    sample="oracle/segment_oracle.py over a size-stratified sample of the same workload: %s" % sample)
        if build_ref.available():
            rr_, smp = reference_rate(tints_for_cpu, cores, float(os.environ.get("FRS_CPU_SECONDS", "15")))
            line["cpu_baseline"] = dict(value=rr_, unit=UNIT, cores=cores, kind="reference",
                                        sample="unmodified freddie_segment.py (oracle/_ref) -t %d, files to files in tmpfs, "
                                               "size-stratified sample of the same workload: %s" % (cores, smp))
            line["cpu_baseline_port"] = port
        else:
            line["cpu_baseline"] = port
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _sizes_obj(res):
    """BatchResult.sizes (dict) -> an object with the attributes BatchResult's constructor reads."""
    from freddie_b200 import _lib
    s = _lib.FrsResultSizes()
    for k, v in res.sizes.items():
        setattr(s, k, v)
    return s


def _upload_bytes(eng, batch, prm):
    """Bytes one step moves host -> device: the copies of frs_upload plus the plane words the clip-fetch kernel
    reads from pinned host memory (lazy sequence mode)."""
    st = eng.stats()
    eng.segment_batch(batch, prm)
    st = eng.stats()
    return st["h2d_upload"] + st["h2d_run"]


def cli_scope(cores):
    """The files-to-files scope of SURVEY.md 8d (what the reference is measured at): SPLIT text on disk ->
    native parser -> CUDA pipeline -> native formatter -> SEGMENT files, over the WHOLE workload of one GPU.
    `value` is the cold number a user sees: wall clock of a fresh ``python -m freddie_b200.segment`` process
    (interpreter start, imports, CUDA context creation, every allocation included); `warm_value` is a second
    pass inside one process (contexts and buffers warm).  Child processes with a time limit, never fatal: the
    kernel-path numbers do not depend on it."""
    import shutil
    import tempfile
    try:
        from freddie_b200 import synth
        work = tempfile.mkdtemp(prefix="frs_bench_cli_")
        try:
            sd = os.path.join(work, "split")
            jobs = synth.config_jobs(2)
            made = synth.write_jobs(jobs, sd, workers=max(1, min(32, cores)))
            n_reads = sum(m[2] for m in made)
            split_bytes = sum(os.path.getsize(os.path.join(b, f)) for b, _, fs in os.walk(sd) for f in fs)
            od = os.path.join(work, "seg_cold")
            t0 = time.perf_counter()
            r = subprocess.run([sys.executable, "-m", "freddie_b200.segment", "-s", sd, "-o", od, "-t", str(cores)],
                               cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=300)
            cold = time.perf_counter() - t0
            if r.returncode != 0:
                return dict(value=None, unit=UNIT, error=(r.stderr.strip().splitlines() or ["exit %d" % r.returncode])[-1][:300])
            n_files = sum(len(fs) for _, _, fs in os.walk(od))
            out = dict(value=n_reads / cold, unit=UNIT, seconds=round(cold, 3), host_threads=cores,
                       sample="whole workload: %d tints, %d reads, %.0f MB of SPLIT text in, %d files out" % (
                           len(made), n_reads, split_bytes / 1e6, n_files),
                       scope="files to files, cold: a fresh `python -m freddie_b200.segment -s SPLIT -o OUT -t %d` process "
                             "(interpreter, imports, CUDA context creation and all allocations inside the wall clock), measured before the "
                             "bench process itself creates a CUDA context" % cores)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cli-run", work, "--cli-threads", str(cores)],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
            if r.returncode == 0:
                w = json.loads(r.stdout.strip().splitlines()[-1])
                out.update(warm_value=w["value"], warm_seconds=w["seconds"], in_process_first_run_seconds=w["first_run_seconds"])
            return out
        finally:
            shutil.rmtree(work, ignore_errors=True)
    except Exception as e:  # noqa: BLE001
        return dict(value=None, unit=UNIT, error="%s: %s" % (type(e).__name__, e))


def cli_run(work, threads):
    """Child of cli_scope: the directory driver twice over work/split inside ONE process (the second pass has
    warm CUDA contexts and buffers), one JSON object on stdout."""
    import shutil
    from freddie_b200.engine import SegmentParams
    from freddie_b200.segment import run_directory
    sd, od = os.path.join(work, "split"), os.path.join(work, "seg")
    t0 = time.perf_counter()
    run_directory(sd, od, SegmentParams(), threads=threads, gpus=1, progress=False)
    first = time.perf_counter() - t0
    shutil.rmtree(od, ignore_errors=True)
    t0 = time.perf_counter()
    st = run_directory(sd, od, SegmentParams(), threads=threads, gpus=1, progress=False)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(value=st["reads"] / dt, unit=UNIT, seconds=round(dt, 4), first_run_seconds=round(first, 4))))


def _download_into(eng, res):
    """frs_download into the already allocated pinned result buffers of the warm-up step."""
    import ctypes as C
    r = res.as_struct()
    eng._check(eng.lib.frs_download(eng.ctx, C.byref(r)))
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--batch-reads", type=int, default=131072, help="reads per GPU batch (cfg4 / cfg5 shards)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the files-to-files scope")
    ap.add_argument("--cli-run", default=None, help=argparse.SUPPRESS)
    ap.add_argument("--cli-threads", type=int, default=1, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.cli_run:
        cli_run(args.cli_run, args.cli_threads)
        return
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_cuda_arm(args)


if __name__ == "__main__":
    main()
