#!/usr/bin/env python3
"""A/B timing of kernel-variant builds of the CUDA library (never under ncu).

    python tools/ab_variants.py [--log2n 20] [--reps 3] lib_a.so lib_b.so ...

Each library is loaded in its own process (fresh CUDA context); prints the device time of k_verify_fast over
2^log2n device-resident signatures (CUDA events around the dominant kernel, schnorr_b200_last_kernel_ms) and checks
that every verdict is 0 and nothing was handed to the exact kernel.  Experiment tooling only: the product loads
schnorr-sig_b200/csrc/libschnorr_b200.so."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
import schnorr_sig_b200 as sb
_lib = sys.modules["schnorr_sig_b200._lib"]
so = %(so)r
import ctypes as C
L = C.CDLL(so)
for name in list(_lib._SIGNATURES):          # older builds lack newer test hooks
    if not hasattr(L, name):
        _lib._SIGNATURES.pop(name)
for name, (res, args) in _lib._SIGNATURES.items():
    fn = getattr(L, name); fn.restype = res; fn.argtypes = args
_lib._LIB = L
n = 1 << %(log2n)d
dev = torch.device("cuda", 0)
eng = sb.Engine(0)
st = torch.cuda.Stream(device=dev)
eng.set_stream(st.cuda_stream)
hin = sb.synth.host_inputs(sb.synth.DEFAULT_SEED, n, %(msg_len)d)
with torch.cuda.stream(st):
    d_sk = torch.from_numpy(hin["sk"]).to(dev); d_nonce = torch.from_numpy(hin["nonce"]).to(dev)
    d_blob = torch.from_numpy(hin["blob"]).to(dev); d_off = torch.from_numpy(hin["off"].view(np.int64)).to(dev)
    d_pk = torch.empty((n, 96), dtype=torch.uint8, device=dev); d_inf = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_sigs = torch.empty((n, 81), dtype=torch.uint8, device=dev); d_out = torch.empty(n, dtype=torch.uint8, device=dev)
    eng.keygen_dev(n, d_sk, d_pk, d_inf)
    eng.sign_many_dev(n, d_sk, d_pk, d_inf, d_blob, d_off, d_nonce, d_sigs)
    st.synchronize()
    eng.set_dist_threshold(0)
    ms = []
    for r in range(%(reps)d + 1):
        eng.verify_many_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_out)
        st.synchronize()
        ms.append(eng.last_kernel_ms())
    ok = int(d_out.max().item()) == 0
    print("RESULT " + json.dumps({"so": os.path.basename(so), "ms": ms[1:], "best_ms": min(ms[1:]), "verdicts_ok": ok,
                                  "exact": int(eng.last_exact_count()), "n": n}))
'''


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--msg-len", type=int, default=8)
    ap.add_argument("libs", nargs="+")
    a = ap.parse_args()
    rows = []
    for so in a.libs:
        so = os.path.abspath(so)
        code = CHILD % {"root": ROOT, "so": so, "log2n": a.log2n, "reps": a.reps, "msg_len": a.msg_len}
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if r.returncode != 0 or not line:
            print("FAILED %s\n%s\n%s" % (so, r.stdout[-2000:], r.stderr[-3000:]), flush=True)
            continue
        row = json.loads(line[0][7:])
        rows.append(row)
        print("%-32s best %.3f ms  (%s)  %.3f M/s  verdicts_ok=%s exact=%d" % (
            row["so"], row["best_ms"], ", ".join("%.3f" % m for m in row["ms"]), row["n"] / row["best_ms"] / 1e3,
            row["verdicts_ok"], row["exact"]), flush=True)
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
