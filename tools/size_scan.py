#!/usr/bin/env python3
"""Wall-clock scan of odd call sizes through the host API (pageable numpy buffers): looks for timing cliffs around the
kernel-choice threshold, the wave size, the pipeline chunking threshold and the MSM planner's window widths."""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import schnorr_sig_b200 as s
eng = s.default_engine(0)
w = s.synth.signed_workload(eng, 7, 700000, msg_len=8)
for n in (1, 5, 1000, 10240, 10241, 20000, 37888, 37889, 75776, 100000, 300000, 530000, 540000, 700000):
    sig, pk, inf, off = w["sigs"][:n], w["pk"][:n], w["inf"][:n], w["off"][:n + 1]
    blob = w["blob"][:8 * n]
    eng.verify_many(sig, pk, inf, blob, off)
    t0 = time.perf_counter(); v = eng.verify_many(sig, pk, inf, blob, off); dt = time.perf_counter() - t0
    assert int(v.max()) == 0
    print("verify_many n=%7d  %8.3f ms  %.3e /s" % (n, dt * 1e3, n / dt), flush=True)
for n in (3, 100, 5000, 12345, 33333, 100000, 131072, 300000):
    sig, pk, inf, off = w["sigs"][:n], w["pk"][:n], w["inf"][:n], w["off"][:n + 1]
    blob = w["blob"][:8 * n]
    eng.verify_batch(sig, pk, inf, blob, off, w["rand"][:n])
    t0 = time.perf_counter(); v, _, _ = eng.verify_batch(sig, pk, inf, blob, off, w["rand"][:n]); dt = time.perf_counter() - t0
    assert v == 0
    print("verify_batch n=%7d  %8.3f ms  %.3e /s" % (n, dt * 1e3, n / dt), flush=True)
