#!/usr/bin/env python3
"""verify_batch latency of small batches through the host API over forced Pippenger window widths (test hook
schnorr_b200_set_msm_geometry): is the planner's width (chosen for total work) also the best for latency?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, cref
import schnorr_sig_b200 as sb
eng = sb.default_engine(0)
def t(fn, reps=15):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)) * 1e3
for n in (4, 32, 128, 1024, 4096):
    w = cref.workload(2, n, 80, 8)
    rand = sb.synth.scalars(5, 1, n)
    row = []
    for c in (0, 4, 5, 6, 7, 8, 9, 10, 12):
        eng.set_msm_geometry(c, 0)
        v = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand)[0]
        assert v == 0
        ms = t(lambda: eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand))
        row.append("c=%s %.3f" % (c if c else "plan(%d)" % eng.last_batch_plan()[0], ms))
    eng.set_msm_geometry(0, 0)
    print("n=%5d  " % n + "  ".join(row), flush=True)
