#!/usr/bin/env python3
"""Dev tool (GPU box): where the end-to-end step of bench.py spends its wall clock.

1. serial breakdown of one lane: frs_upload / frs_run / frs_download, wall clock, synchronised;
2. pipelined throughput with 1..4 lanes (library contexts on their own host threads), as bench.py's e2e leg;
3. raw pinned PCIe copy rates for the same byte counts (the floor of the e2e step).
"""
import ctypes as C
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freddie_b200 import synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

tints = synth.make_config(2, scale=float(os.environ.get("SCALE", "1")), seed=2, workers=16)
batch = pack_tints(tints).pin()
prm = SegmentParams()
lanes = []
N_LANES = int(os.environ.get('LANES_MAX', '8'))
for _ in range(N_LANES):
    e = Engine(0)
    r = None
    for _ in range(3):
        r = e.segment_batch(batch, prm, pinned=True)
    lanes.append((e, r))
st = lanes[0][0].stats()
h2d = st["h2d_upload"] + st["h2d_run"]
d2h = int(sum(v.nbytes for v in lanes[0][1].arrays.values())) + st["d2h_run"]
print("bytes per step: h2d %.1f MB  d2h %.1f MB" % (h2d / 1e6, d2h / 1e6))

e, r = lanes[0]
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e.upload(batch); torch.cuda.synchronize()
    t1 = time.perf_counter(); e.run(prm); torch.cuda.synchronize()
    t2 = time.perf_counter()
    rs = r.as_struct(); e._check(e.lib.frs_download(e.ctx, C.byref(rs))); torch.cuda.synchronize()
    t3 = time.perf_counter()
    print("serial: upload %.2f ms  run %.2f ms  download %.2f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))


def lane_work(idx, steps):
    e2, r2 = lanes[idx]
    for _ in range(steps):
        e2.upload(batch)
        e2.run(prm)
        rs = r2.as_struct()
        e2._check(e2.lib.frs_download(e2.ctx, C.byref(rs)))


for n_used in [n for n in (1, 2, 3, 4, 6, 8) if n <= N_LANES]:
    for rep in range(2):
        steps = 24
        per = [steps // n_used + (1 if i < steps % n_used else 0) for i in range(n_used)]
        ths = [threading.Thread(target=lane_work, args=(i, per[i])) for i in range(n_used)]
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for t in ths:
            t.start()
        for t in ths:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - w0
        print("lanes %d: %.2f ms/step  %.1f M reads/s" % (n_used, dt / steps * 1e3, batch.n_reads * steps / dt / 1e6))

# raw copies
hb = torch.empty(h2d, dtype=torch.uint8).pin_memory()
db = torch.empty(h2d, dtype=torch.uint8, device="cuda")
ho = torch.empty(d2h, dtype=torch.uint8).pin_memory()
do = torch.empty(d2h, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, both in (("h2d alone", False), ("h2d + d2h concurrently", True)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1):
            db.copy_(hb, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 10
    print("%s: %.2f ms per step's bytes (h2d %.1f GB/s)" % (name, dt * 1e3, h2d / dt / 1e9))
