#!/usr/bin/env python3
"""Single-call latency (one Signature::verify through the host API) of library builds: python tools/one_ab.py lib.so ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time
sys.path[:0] = [%(root)r, %(root)r + "/oracle"]
import ctypes as C, numpy as np, cref
import schnorr_sig_b200 as sb
_lib = sys.modules["schnorr_sig_b200._lib"]
L = C.CDLL(%(so)r)
for name in list(_lib._SIGNATURES):
    if not hasattr(L, name): _lib._SIGNATURES.pop(name)
for name, (res, args) in _lib._SIGNATURES.items():
    fn = getattr(L, name); fn.restype = res; fn.argtypes = args
_lib._LIB = L
eng = sb.Engine(0)
out = []
for n in (1, 64, 512):
    w = cref.workload(3, n, 8, 8)
    want = cref.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], 8)
    for _ in range(3):
        got = eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])
    assert np.array_equal(got, want)
    ts = []
    for _ in range(30):
        t0 = time.perf_counter(); eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]); ts.append(time.perf_counter() - t0)
    out.append("n=%%d %%.3f ms (min %%.3f)" %% (n, np.mean(ts) * 1e3, np.min(ts) * 1e3))
print("RESULT %%s: %%s" %% (%(so)r.split("/")[-1], "   ".join(out)))
'''
for so in sys.argv[1:]:
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "so": os.path.abspath(so)}], capture_output=True, text=True)
    lines = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    print(lines[0] if lines else "FAILED %s\n%s" % (so, r.stderr[-1500:]), flush=True)
