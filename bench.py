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
    "cfg2": "BASELINE configs[1]: synthetic chromosome-scale SPLIT, 200k reads across ~3k tints (seeded, per GPU)",
    "cfg3": "BASELINE configs[2]: DP-dominated giant tints (scaled by --scale), per GPU",
}


# ------------------------------------------------------------------------------------------------
def rank_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_workload(workload: str, scale: float, seed: int, workers: int):
    from freddie_b200 import synth
    cfg = {"cfg2": 2, "cfg3": 3}[workload]
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
def run_cuda_arm(args):
    rank, local_rank, world = rank_env()
    cores = os.cpu_count() or 1
    # torchrun pins OMP_NUM_THREADS=1 in its children; the library's host gather (clip words) and the
    # native parser use OpenMP, so give every rank its share of the cores
    os.environ["OMP_NUM_THREADS"] = str(max(1, cores // max(world, 1)))
    # keep stdout to the ONE JSON line: NCCL prints its version banner there at NCCL_DEBUG >= VERSION
    if "FRS_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["FRS_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from freddie_b200.engine import Engine, SegmentParams
    from freddie_b200.pack import pack_tints

    t0 = time.time()
    tints = make_workload(args.workload, args.scale, 2 + 1000 * rank, max(1, min(16, cores // world)))
    batch = pack_tints(tints).pin()
    t_gen = time.time() - t0
    n_reads = batch.n_reads
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
    from freddie_b200 import _lib
    eng.set_option(_lib.OPT_LAZY_SEQ, 0)  # `value`: every input, sequence planes included, resident in HBM
    res = None
    for _ in range(max(args.warmup, 1)):
        res = eng.segment_batch(batch, prm, pinned=True)
    sizes = res.sizes

    # ---- timed: K x frs_run on the resident batch ----
    eng.upload(batch)
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank)
    stage_ms, stage_launch = {}, {}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    launches = 0
    for k in range(args.steps):
        flush_l2()
        ev[k][0].record(stream)
        eng.run(prm)
        ev[k][1].record(stream)
        launches += eng.launch_count()
        for name, ms, ln in eng.timings():
            stage_ms[name] = stage_ms.get(name, 0.0) + ms
            stage_launch[name] = ln
    barrier()
    clocks = sampler.stop()
    t_dev = sum(a.elapsed_time(b) for a, b in ev) / 1e3
    eng.set_profiling(False)

    # ---- timed: end to end through the C ABI with HOST buffers.  Every step uploads its inputs from
    # pinned host memory, runs, and downloads the results into pinned host memory.  Several library
    # contexts (one host thread each, up to 6: tests/e2e_probe.py shows the rate levelling off there)
    # take alternate steps so that the copies and host round trips of one step overlap the kernels of
    # the others, exactly as the CLI driver runs consecutive batches. ----
    n_lanes = max(1, min(int(os.environ.get("FRS_E2E_LANES", "6")), args.steps))
    lanes = []
    for _ in range(n_lanes):
        e2 = Engine(local_rank)
        r2 = None
        for _ in range(max(args.warmup, 1)):
            r2 = e2.segment_batch(batch, prm, pinned=True)
        lanes.append((e2, r2))
    st = lanes[0][0].stats()
    h2d = st["h2d_upload"] + st["h2d_run"]
    d2h = int(sum(v.nbytes for v in lanes[0][1].arrays.values())) + st["d2h_run"]

    def lane_work(idx, steps):
        e2, r2 = lanes[idx]
        for _ in range(steps):
            e2.upload(batch)
            e2.run(prm)
            _download_into(e2, r2)

    def timed_e2e(n_used):
        per = [args.steps // n_used + (1 if i < args.steps % n_used else 0) for i in range(n_used)]
        ths = [threading.Thread(target=lane_work, args=(i, per[i])) for i in range(n_used)]
        barrier()
        w0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        return time.perf_counter() - w0

    t_e2e_serial = timed_e2e(1)
    t_e2e = timed_e2e(n_lanes)
    barrier()
    same = all(np.array_equal(lanes[0][1].arrays[k], res.arrays[k]) for k in res.arrays)
    if not same:
        raise RuntimeError("end-to-end (lazy sequence) results differ from the resident run")

    # ---- reduce over ranks: max time, sum of units ----
    tot_reads, tot_cells = n_reads, int(sizes["dp_cells"])
    dp_ms = stage_ms.get("dp", 0.0) + stage_ms.get("dp_solve", 0.0)
    if world > 1:
        t = torch.tensor([t_dev, t_e2e, dp_ms, t_e2e_serial], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev, t_e2e, dp_ms, t_e2e_serial = [float(x) for x in t.tolist()]
        u = torch.tensor([n_reads, tot_cells, launches, h2d, d2h], device="cuda", dtype=torch.int64)
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        tot_reads, tot_cells, launches, h2d, d2h = [int(x) for x in u.tolist()]  # whole-job totals
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = tot_reads * args.steps / t_dev
    e2e_v = tot_reads * args.steps / t_e2e
    counts = batch.counts()
    cov_elems = int(eng.tap(12, np.int64)[-1])  # FRS_TAP_COV_OFF: last entry = total coverage elements
    alg = algorithmic_bytes(counts, sizes, cov_elems)
    peak, peak_src = peaks()
    # dominant KERNEL of the step = the longest single launch.  Stages that are one launch on the context
    # stream are timed exactly by their CUDA events; the DP stage is six kernels running concurrently on
    # side streams (the longest of them is shorter than k_smooth, see profiles/*_launches.txt) and is
    # issue-bound, not HBM-bound: it is reported in `dp_stage` instead.
    single = [k for k in stage_ms if k in alg and stage_launch.get(k) == 1]
    dom = max(single or [k for k in stage_ms if k in alg], key=lambda k: stage_ms[k])
    dom_ms = stage_ms[dom] / args.steps
    ach = alg[dom] / (dom_ms * 1e-3) / 1e9
    stages = {k: dict(ms=round(v / args.steps, 4), launches=stage_launch[k],
                      alg_GBps=(round(alg[k] / (v / args.steps * 1e-3) / 1e9, 1) if k in alg and v > 0 else None),
                      hbm_frac=(round(alg[k] / (v / args.steps * 1e-3) / 1e9 / peak, 4) if k in alg and v > 0 else None))
              for k, v in stage_ms.items()}
    stream_ms = sum(v for k, v in stage_ms.items() if k in alg and k != "dp") / args.steps
    stream_bytes = sum(b for k, b in alg.items() if k != "dp" and k in stage_ms)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp) and args.workload == "cfg2" and args.scale == 1.0:
        traffic = json.load(open(tp)).get(dom)
    line = dict(
        metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=t_dev / args.steps * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None,
        dtype="int32+f64", data="synthetic",
        config=dict(workload=WORKLOADS[args.workload], scale=args.scale, reads_per_gpu=n_reads, tints_per_gpu=len(tints),
                    l2="flushed between timed steps (256 MiB memset)", params="defaults (sd=5 tp=0.9 vf=3 mps=50 lo=3)"),
        e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 timing="wall clock around K steps, synchronize on both sides; %d contexts take alternate steps" % n_lanes,
                 serial_value=tot_reads * args.steps / t_e2e_serial,
                 clip_words_per_step=st["clip_words"], seq_words_in_batch=st["seq_words"]),
        gpu_launches=launches,
        clocks=clocks,
        roofline=dict(bound="hbm", kernel=KERNEL_OF.get(dom, dom), stage=dom, achieved=ach, peak=peak, unit="GB/s", frac=ach / peak, traffic=traffic,
                      peak_source=peak_src, alg_bytes_per_launch=alg[dom], ms_per_launch=dom_ms,
                      note="longest single launch of the step, timed by its own CUDA events; the concurrent DP kernels "
                           "are issue-bound and reported in dp_stage",
                      streaming_stages=dict(achieved=round(stream_bytes / (stream_ms * 1e-3) / 1e9, 1),
                                            frac=round(stream_bytes / (stream_ms * 1e-3) / 1e9 / peak, 4),
                                            ms=round(stream_ms, 4), alg_bytes=stream_bytes,
                                            what="all HBM-streaming stages together (every stage but dp)")),
        dp_stage=dict(bound="issue (VOTE/LOP3/POPC), not HBM: coverage rows are read once (TMA) and reused on chip",
                      ms=round(dp_ms / args.steps, 4), launches=stage_launch.get("dp"),
                      alg_GBps=round(alg["dp"] / (dp_ms / args.steps * 1e-3) / 1e9, 1) if dp_ms > 0 else None,
                      evidence="profiles/: issue-slot utilisation and stall breakdown of every DP kernel"),
        dp_cells_per_sec=tot_cells * args.steps / max(dp_ms * 1e-3, 1e-12),
        dp_read_cells_per_sec=int(sizes["dp_read_cells"]) * args.steps * world / max(dp_ms * 1e-3, 1e-12),
        dp=dict(cells=int(sizes["dp_cells"]), subproblems=int(sizes["n_subproblems"]),
                max_n=int(sizes["max_subproblem"]), candidates=int(sizes["n_candidates"])),
        stages=stages,
        setup_seconds=round(t_gen, 1),
    )
    if world == 1 and not args.no_cli:
        line["cli"] = cli_scope(tints, cores)
    if world == 1 and not args.no_cpu_baseline:
        r, c, sample = oracle_rate(tints, cores, int(os.environ.get("FRS_CPU_SAMPLE_READS", 300 * cores * 15)))
        line["cpu_baseline"] = dict(value=r, unit=UNIT, cores=cores, kind="port", dp_cells_per_sec=c,
                                    sample="oracle/segment_oracle.py over a size-stratified sample of the same "
                                           "workload: %s" % sample)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cli_scope(tints, cores):
    """The files-to-files scope of SURVEY.md 8d (what the reference is measured at): SPLIT text on disk ->
    native parser -> CUDA pipeline -> native formatter -> SEGMENT files, through the drop-in directory driver
    (`freddie_b200.segment.run_directory`), on a bounded prefix of the workload's tints.  Runs in a child
    process with a time limit and is never fatal: the kernel-path numbers do not depend on it."""
    import shutil
    import tempfile
    try:
        from freddie_b200 import synth
        n_target = int(os.environ.get("FRS_CLI_SAMPLE_READS", "50000"))
        sample, n = [], 0
        for t in tints:
            sample.append(t)
            n += len(t["reads"])
            if n >= n_target:
                break
        work = tempfile.mkdtemp(prefix="frs_bench_cli_")
        try:
            sd = os.path.join(work, "split")
            synth.write_split_dir(sample, sd)
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cli-run", work, "--cli-threads", str(cores)],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=240)
            if r.returncode != 0:
                return dict(value=None, unit=UNIT, error=(r.stderr.strip().splitlines() or ["exit %d" % r.returncode])[-1][:300])
            out = json.loads(r.stdout.strip().splitlines()[-1])
            out["sample"] = "first %d tints of the workload: %s" % (len(sample), out.pop("sample"))
            return out
        finally:
            shutil.rmtree(work, ignore_errors=True)
    except Exception as e:  # noqa: BLE001
        return dict(value=None, unit=UNIT, error="%s: %s" % (type(e).__name__, e))


def cli_run(work, threads):
    """Child of cli_scope: the directory driver twice over work/split (first run warms the CUDA contexts and
    the library's buffers), one JSON object on stdout."""
    import shutil
    from freddie_b200.engine import SegmentParams
    from freddie_b200.segment import run_directory
    sd, od = os.path.join(work, "split"), os.path.join(work, "seg")
    split_bytes = sum(os.path.getsize(os.path.join(b, f)) for b, _, fs in os.walk(sd) for f in fs)
    t0 = time.perf_counter()
    run_directory(sd, od, SegmentParams(), threads=threads, gpus=1, progress=False)
    cold = time.perf_counter() - t0
    shutil.rmtree(od, ignore_errors=True)
    t0 = time.perf_counter()
    st = run_directory(sd, od, SegmentParams(), threads=threads, gpus=1, progress=False)
    dt = time.perf_counter() - t0
    n_files = sum(len(fs) for _, _, fs in os.walk(od))
    print(json.dumps(dict(
        value=st["reads"] / dt, unit=UNIT, seconds=round(dt, 4), first_run_seconds=round(cold, 4), host_threads=threads,
        sample="%d reads, %.0f MB of SPLIT text in, %d files out" % (st["reads"], split_bytes / 1e6, n_files),
        scope="files to files: native parser -> CUDA pipeline -> native formatter (second run of the process: CUDA "
              "contexts warm; first_run_seconds includes their creation)")))


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
