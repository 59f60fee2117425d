#!/usr/bin/env python3
"""Dev tool: wall-clock breakdown of one end-to-end step (upload / run / download) on the GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from freddie_b200 import synth
from freddie_b200.engine import Engine, SegmentParams
from freddie_b200.pack import pack_tints
import ctypes as C

tints = synth.make_config(2, scale=float(os.environ.get("SCALE", "1")), seed=2, workers=16)
batch = pack_tints(tints).pin()
eng = Engine(0)
prm = SegmentParams()
res = eng.segment_batch(batch, prm, pinned=True)
res = eng.segment_batch(batch, prm, pinned=True)
eng.set_profiling(True)
for it in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); eng.upload(batch); torch.cuda.synchronize()
    t1 = time.perf_counter(); eng.run(prm); torch.cuda.synchronize()
    t2 = time.perf_counter()
    r = res.as_struct(); eng._check(eng.lib.frs_download(eng.ctx, C.byref(r)))
    t3 = time.perf_counter()
    tm = eng.timings()
    print("upload %.2f ms  run %.2f ms (kernels %.2f ms)  download %.2f ms" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, sum(x[1] for x in tm), (t3 - t2) * 1e3))
print({k: round(v, 3) for k, v, _ in tm})
print(eng.stats())
