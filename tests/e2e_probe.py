#!/usr/bin/env python3
"""Dev tool (GPU box): where the end-to-end step of bench.py spends its wall clock.

1. serial breakdown of one context: frs_upload / frs_run / frs_download, wall clock, synchronised;
2. stage timings of a run in lazy-sequence mode (incl. the clip-fetch kernel that reads pinned host memory);
3. pipelined steps (frs_submit / frs_wait / frs_fetch) of ONE context: host time inside each call;
4. raw pinned PCIe copy rates for the same byte counts (the floor of the e2e step).
"""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freddie_b200 import synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

tints = synth.make_config(2, scale=float(os.environ.get("SCALE", "1")), seed=2, workers=16)
batch = pack_tints(tints).pin(edge_words=int(os.environ["EDGE"]) if "EDGE" in os.environ else None)
prm = SegmentParams()
e = Engine(0)
r = None
for _ in range(3):
    r = e.segment_batch(batch, prm, pinned=True)
st = e.stats()
h2d = st["h2d_upload"] + st["h2d_run"]
d2h = int(sum(v.nbytes for v in r.arrays.values()))
print("bytes per step: h2d %.1f MB (of which fetched by the clip kernel %.1f MB)  d2h %.1f MB  reruns %d" % (
    h2d / 1e6, st["h2d_run"] / 1e6, d2h / 1e6, st["reruns"]))

for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e.upload(batch)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    e.run(prm)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    rs = r.as_struct()
    e._check(e.lib.frs_download(e.ctx, C.byref(rs)))
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("serial: upload %.3f ms  run %.3f ms  download %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))

e.set_profiling(True)
e.run(prm)
print("stages (lazy sequence): " + "  ".join("%s %.3f" % (n, ms) for n, ms, _ in e.timings()))
e.set_profiling(False)

tk0 = e.submit(batch, prm)
s0 = e.wait(tk0)
bufs = [e.new_result(s0, batch, pinned=True) for _ in range(2)]
e.fetch(tk0, bufs[0])

from collections import deque  # noqa: E402


def pipelined(K):
    acc = dict(submit=0.0, wait=0.0, fetch=0.0)
    torch.cuda.synchronize()
    w0 = time.perf_counter()
    fly = deque()
    pending = None
    for k in range(K + 6):
        t0 = time.perf_counter()
        if k < K:
            fly.append(e.submit(batch, prm))
        ta = time.perf_counter()
        if pending is not None:
            e.fetch_finish(pending)
            pending = None
        t1 = time.perf_counter()
        acc["submit"] += ta - t0
        acc["fetch"] += t1 - ta
        if len(fly) == int(os.environ.get("DEPTH", "4")) or (k >= K and fly):
            pt = fly.popleft()
            e.wait(pt)
            t2 = time.perf_counter()
            e.fetch_start(pt, bufs[k & 1])
            pending = pt
            t3 = time.perf_counter()
            acc["wait"] += t2 - t1
            acc["fetch"] += t3 - t2
    if pending is not None:
        e.fetch_finish(pending)
    torch.cuda.synchronize()
    return time.perf_counter() - w0, acc


pipelined(12)  # every slot has its buffers now
K = 20
e.set_profiling(True)
dt, acc = pipelined(K)
print("stages of the last pipelined run (its head ran beside the previous tail): " + "  ".join("%s %.3f" % (n, ms) for n, ms, _ in e.timings()))
e.set_profiling(False)
dt, acc = pipelined(K)
print("pipelined, one context: %.3f ms/step (%.1f M reads/s); host time per step: submit %.3f  wait %.3f  fetch %.3f ms" % (
    dt / K * 1e3, batch.n_reads * K / dt / 1e6, acc["submit"] / K * 1e3, acc["wait"] / K * 1e3, acc["fetch"] / K * 1e3))

# PCIe floor for the same byte counts
a = torch.empty(h2d, dtype=torch.uint8).pin_memory()
d = torch.empty(max(h2d, d2h), dtype=torch.uint8, device="cuda")
b = torch.empty(d2h, dtype=torch.uint8).pin_memory()
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    d[:h2d].copy_(a, non_blocking=True)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    b.copy_(d[:d2h], non_blocking=True)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
print("pcie: h2d %.3f ms (%.1f GB/s)  d2h %.3f ms (%.1f GB/s)" % (
    (t1 - t0) * 1e3, h2d / (t1 - t0) / 1e9, (t2 - t1) * 1e3, d2h / (t2 - t1) / 1e9))
