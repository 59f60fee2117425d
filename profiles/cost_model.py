#!/usr/bin/env python3
"""Dev tool (GPU box): the scheduler's cost estimate against measured GPU time, tint by tint.

`freddie_b200.schedule.estimate_cost(n_reads)` is what the LPT bin-packing of the directory driver and of
bench.py shards by (SURVEY.md 8e asks for a cost from L, K and R; only the read count -- from the split file
size -- is known before parsing, so the estimate is a function of the reads alone).  This script runs tints of
a power-law size mix (BASELINE configs[3], scaled) ONE PER BATCH, records the device time of frs_run, and
prints estimate, time and their ratio per size decade plus the rank correlation.

    python profiles/cost_model.py [scale] > profiles/rNN_cost_model.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from freddie_b200 import schedule, synth  # noqa: E402
from freddie_b200.engine import Engine, SegmentParams  # noqa: E402
from freddie_b200.pack import pack_tints  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.25
jobs = synth.config_jobs(4, scale=scale)
# a size-stratified pick: every decade of the read count, up to 8 tints each
by_dec = {}
for j in sorted(jobs, key=lambda j: -j[3]):
    by_dec.setdefault(int(np.log10(max(j[3], 1))), []).append(j)
pick = [j for d in sorted(by_dec) for j in by_dec[d][:: max(1, len(by_dec[d]) // 8)][:8]]
tints = synth.run_jobs(pick, workers=min(16, os.cpu_count() or 1))
eng = Engine(0)
prm = SegmentParams()
st = torch.cuda.ExternalStream(eng.lib.frs_stream(eng.ctx))
rows = []
for t in tints:
    b = pack_tints([t]).pin()
    eng.segment_batch(b, prm)
    eng.upload(b)
    ms = []
    for _ in range(3):
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        eng.run(prm)
        z.record(st)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(z))
    rows.append((len(t["reads"]), schedule.estimate_cost(len(t["reads"])), min(ms)))
rows.sort()
n = np.array([r[0] for r in rows], float)
est = np.array([r[1] for r in rows], float)
ms = np.array([r[2] for r in rows], float)
floor = ms.min()  # launch-bound floor of a one-tint batch (~50 launches)
print("# one tint per batch; device ms of frs_run (best of 3); launch-bound floor %.3f ms" % floor)
print("%10s %14s %10s %16s" % ("reads", "estimate", "ms", "(ms-floor)/est"))
for r in rows:
    print("%10d %14.0f %10.3f %16.3e" % (r[0], r[1], r[2], max(r[2] - floor, 0) / r[1]))
rk = lambda x: np.argsort(np.argsort(x))  # noqa: E731
print("# Spearman rank correlation estimate ~ time: %.4f" % np.corrcoef(rk(est), rk(ms))[0, 1])
big = n >= 1000
if big.sum() > 2:
    ratio = (ms[big] - floor) / est[big]
    print("# tints of >= 1000 reads: (ms - floor) / estimate spans %.2e .. %.2e (x%.1f)" % (ratio.min(), ratio.max(), ratio.max() / ratio.min()))
