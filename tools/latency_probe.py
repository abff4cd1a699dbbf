#!/usr/bin/env python3
"""Latency of small calls through the host API (the reference's Criterion cases, benches/schnorr.rs:22-24,67-96):
single verify at 8/80/160-byte messages and verify_batch at 4..128 signatures over one 80-byte message,
next to the single-thread CPU restatement (oracle/cref.c)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np  # noqa: E402
import cref  # noqa: E402
import schnorr_sig_b200 as sb  # noqa: E402


def timeit(fn, reps=20):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts)) * 1e3, float(np.std(ts)) * 1e3


def main():
    eng = sb.default_engine(0)
    print("| case | GPU host-API ms (mean ± sd of 20) | CPU oracle 1 thread ms |")
    print("|---|---|---|")
    for L in (8, 80, 160):
        w = cref.workload(1, 1, L, 1)
        g = timeit(lambda: eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]))
        c = timeit(lambda: cref.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], 1))
        print("| Verify - %d bytes | %.3f ± %.3f | %.3f |" % (L, g[0], g[1], c[0]))
    for n in (4, 16, 32, 64, 128, 1024, 16384):
        w = cref.workload(2, n, 80, 8)
        rand = sb.synth.scalars(5, 1, n)
        g = timeit(lambda: eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand), 10)
        c = timeit(lambda: cref.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], rand, 1), 3 if n > 128 else 5) if n <= 1024 else (float("nan"), 0)
        g2 = timeit(lambda: eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]), 10)
        print("| Verify batch - %d signatures | %.3f ± %.3f (independent verify_many: %.3f) | %.3f |" % (n, g[0], g[1], g2[0], c[0]))


if __name__ == "__main__":
    main()
