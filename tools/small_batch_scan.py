#!/usr/bin/env python3
"""verify_batch latency through the host API, one thread block per signature (k_batch_small) against the Pippenger
pipeline: locates the crossover behind schnorr_b200_set_batch_small_threshold."""
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
for n in (1, 4, 32, 128, 256, 296, 400, 512, 1024):
    w = cref.workload(2, n, 80, 8)
    rand = sb.synth.scalars(5, 1, n)
    row = []
    for thr in (2**62, 0):
        eng.set_batch_small_threshold(thr)
        assert eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand)[0] == 0
        row.append(t(lambda: eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand)))
    print("n=%5d  block-per-signature %.3f ms   Pippenger pipeline %.3f ms" % (n, row[0], row[1]), flush=True)
