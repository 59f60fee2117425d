#!/usr/bin/env python3
"""Dev tool (GPU box): DP-stage time of a giant-tint workload (config 3 at SCALE) for several slab sizes
(FRS_OPT_SLAB_WORDS) in one process; the batch is generated once."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freddie_b200 import _lib, synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

scale = float(os.environ.get("SCALE", "0.25"))
tints = synth.make_config(3, scale=scale, seed=3, workers=16)
batch = pack_tints(tints).pin()
del tints
prm = SegmentParams()
e = Engine(0)
e.set_option(_lib.OPT_LAZY_SEQ, 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for slab in [int(x) for x in os.environ.get("SLABS", "64,128,256,512").split(",")]:
    e.set_option(_lib.OPT_SLAB_WORDS, slab)
    for _ in range(2):
        res = e.segment_batch(batch, prm)
    e.upload(batch)
    e.set_profiling(True)
    acc, K = {}, 3
    for _ in range(K):
        flush.zero_()
        torch.cuda.synchronize()
        e.run(prm)
        torch.cuda.synchronize()
        for n, ms, _ in e.timings():
            acc[n] = acc.get(n, 0.0) + ms
    e.set_profiling(False)
    tot = sum(acc.values()) / K
    print("slab %4d words: stages %.3f ms | dp %.3f  coverage %.3f  digits %.3f  runs %.3f  gaps %.3f | %.3g RCU/s" % (
        slab, tot, acc["dp"] / K, acc["coverage"] / K, acc["digits"] / K, acc["runs"] / K, acc["gaps"] / K,
        res.sizes["dp_read_cells"] / (acc["dp"] / K * 1e-3)), flush=True)
