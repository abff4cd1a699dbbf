#!/usr/bin/env python3
"""Device time of one batch verification (verify_batch_dev) over batch sizes, with the challenges hashed per thread and
on six lanes per signature: locates the crossover behind schnorr_b200_set_batch_dist_threshold."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import schnorr_sig_b200 as sb
dev = torch.device("cuda", 0)
eng = sb.Engine(0)
st = torch.cuda.Stream(device=dev)
eng.set_stream(st.cuda_stream)
nmax = 1 << 20
hin = sb.synth.host_inputs(7, nmax, 8)
with torch.cuda.stream(st):
    d_sk = torch.from_numpy(hin["sk"]).to(dev); d_nonce = torch.from_numpy(hin["nonce"]).to(dev)
    d_blob = torch.from_numpy(hin["blob"]).to(dev); d_off = torch.from_numpy(hin["off"].view(np.int64)).to(dev)
    d_rand = torch.from_numpy(hin["rand"]).to(dev)
    d_pk = torch.empty((nmax, 96), dtype=torch.uint8, device=dev); d_inf = torch.zeros(nmax, dtype=torch.uint8, device=dev)
    d_sigs = torch.empty((nmax, 81), dtype=torch.uint8, device=dev); res = torch.zeros(216, dtype=torch.uint8, device=dev)
    eng.keygen_dev(nmax, d_sk, d_pk, d_inf)
    eng.sign_many_dev(nmax, d_sk, d_pk, d_inf, d_blob, d_off, d_nonce, d_sigs)
    st.synchronize()
for lg in (12, 14, 15, 16, 17, 18, 19, 20):
    n = 1 << lg
    row = []
    for thr in (0, 2**62):
        eng.set_batch_dist_threshold(thr)
        best = 1e9
        for r in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            eng.verify_batch_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_rand, res)
            b.record(st)
            st.synchronize()
            if r:
                best = min(best, a.elapsed_time(b))
        assert int(res[0].item()) == 0
        row.append(best)
    print("n=2^%d  per-thread hash %.3f ms   six-lane hash %.3f ms   -> %.2f M sigs/s" % (lg, row[0], row[1], n / min(row) / 1e3), flush=True)
