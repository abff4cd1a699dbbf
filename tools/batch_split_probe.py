#!/usr/bin/env python3
"""Probe: one 2^k batch as S sub-batches on S streams of the same GPU (batch_partial_dev per slice, one batch_finish_dev),
against the single-launch-chain verify_batch_dev.  Experiment tooling for the pipelined batch path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import schnorr_sig_b200 as sb
dev = torch.device("cuda", 0)
SMAX = 8
engs = [sb.Engine(0) for _ in range(SMAX)]
sts = [torch.cuda.Stream(device=dev) for _ in range(SMAX)]
for e, s in zip(engs, sts):
    e.set_stream(s.cuda_stream)
eng, st = engs[0], sts[0]
nmax = 1 << 18
hin = sb.synth.host_inputs(7, nmax, 8)
with torch.cuda.stream(st):
    d_sk = torch.from_numpy(hin["sk"]).to(dev); d_nonce = torch.from_numpy(hin["nonce"]).to(dev)
    d_blob = torch.from_numpy(hin["blob"]).to(dev); d_off = torch.from_numpy(hin["off"].view(np.int64)).to(dev)
    d_rand = torch.from_numpy(hin["rand"]).to(dev)
    d_pk = torch.empty((nmax, 96), dtype=torch.uint8, device=dev); d_inf = torch.zeros(nmax, dtype=torch.uint8, device=dev)
    d_sigs = torch.empty((nmax, 81), dtype=torch.uint8, device=dev)
    res = torch.zeros(216, dtype=torch.uint8, device=dev); res2 = torch.zeros(216, dtype=torch.uint8, device=dev)
    parts = torch.zeros((SMAX, 192), dtype=torch.uint8, device=dev)
    eng.keygen_dev(nmax, d_sk, d_pk, d_inf)
    eng.sign_many_dev(nmax, d_sk, d_pk, d_inf, d_blob, d_off, d_nonce, d_sigs)
    st.synchronize()
for lg in (14, 16, 17, 18):
    n = 1 << lg
    best = 1e9
    for r in range(4):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); eng.verify_batch_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_rand, res); b.record(st); st.synchronize()
        if r: best = min(best, a.elapsed_time(b))
    line = "n=2^%d single chain %.3f ms" % (lg, best)
    for S in (2, 4, 8):
        m = n // S
        bestS = 1e9
        for r in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            start = torch.cuda.Event(); start.record(st)
            evs = []
            for k in range(S):
                lo = k * m
                if k: sts[k].wait_event(start)
                # message offsets are absolute into the blob: slices of the offset table work as they are
                engs[k].batch_partial_dev(m, d_sigs[lo:], d_pk[lo:], d_inf[lo:], d_blob, d_off[lo:], d_rand[lo:], parts[k])
                if k:
                    ev = torch.cuda.Event(); ev.record(sts[k]); evs.append(ev)
            for ev in evs: st.wait_event(ev)
            eng.batch_finish_dev(S, parts, res2)
            b.record(st); st.synchronize()
            if r: bestS = min(bestS, a.elapsed_time(b))
        ok = bool((res2[:1] == res[:1]).all().item()) and bool((res2[1:98] == res[1:98]).all().item())
        line += "   S=%d %.3f ms%s" % (S, bestS, "" if ok else " (MISMATCH)")
    print(line, flush=True)
