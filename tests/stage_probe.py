#!/usr/bin/env python3
"""Dev tool (GPU box): per-stage CUDA-event times of the cfg2 step for the library named by FRS_LIB
(default: the in-tree build), resident sequence planes, L2 flushed between runs.  The packed batch is cached
in /tmp so that several builds can be compared in one call."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freddie_b200 import _lib, synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import PackedBatch, pack_tints  # noqa: E402

cache = "/tmp/frs_cfg2_batch.npz"
if os.path.exists(cache):
    z = np.load(cache)
    batch = PackedBatch({k: z[k] for k in z.files}, [])
else:
    batch = pack_tints(synth.make_config(2, seed=2, workers=16))
    np.savez(cache, **batch.arrays)
batch.pin()
prm = SegmentParams()
e = Engine(0)
e.set_option(_lib.OPT_LAZY_SEQ, 0)
for _ in range(3):
    e.segment_batch(batch, prm)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e.upload(batch)
e.set_profiling(True)
acc, K = {}, 10
tot = 0.0
for _ in range(K):
    flush.zero_()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = torch.cuda.ExternalStream(e.lib.frs_stream(e.ctx))
    a.record(st)
    e.run(prm)
    b.record(st)
    torch.cuda.synchronize()
    tot += a.elapsed_time(b)
    for n, ms, _ in e.timings():
        acc[n] = acc.get(n, 0.0) + ms
print("%s: step %.3f ms | %s" % (os.path.basename(_lib.LIB_PATH), tot / K, "  ".join("%s %.3f" % (n, v / K) for n, v in acc.items())))
