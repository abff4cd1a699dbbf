#!/usr/bin/env python3
"""Latency of small verify_many calls through the host API with the block-per-signature kernel (k_verify_one) against the
six-lane kernel (k_verify_dist): locates the crossover behind schnorr_b200_set_one_threshold."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np, cref
import schnorr_sig_b200 as sb
eng = sb.default_engine(0)
def t(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)) * 1e3
for n in (1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048):
    w = cref.workload(3, n, 8, 8)
    want = cref.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], 8)
    row = []
    for thr in (2**62, 0):
        eng.set_one_threshold(thr)
        got = eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])
        assert np.array_equal(got, want), (n, thr)
        row.append(t(lambda: eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])))
    print("n=%5d  block-per-signature %.3f ms   six-lane %.3f ms" % (n, row[0], row[1]), flush=True)
