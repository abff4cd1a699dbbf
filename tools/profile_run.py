#!/usr/bin/env python3
"""Small driver for ncu captures: runs each hot path a couple of times on device-resident inputs.

    python tools/profile_run.py [--log2n 16] [--paths verify,hash,batch] [--reps 2]

Prints CUDA-event timings per path (never quote numbers from a run under ncu)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import schnorr_sig_b200 as sb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=16)
    ap.add_argument("--paths", default="verify,hash,batch")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--msg-len", type=int, default=8)
    a = ap.parse_args()
    n, L = 1 << a.log2n, a.msg_len
    dev = torch.device("cuda", 0)
    eng = sb.Engine(0)
    st = torch.cuda.Stream(device=dev)
    eng.set_stream(st.cuda_stream)
    hin = sb.synth.host_inputs(sb.synth.DEFAULT_SEED, n, L)
    with torch.cuda.stream(st):
        d_sk = torch.from_numpy(hin["sk"]).to(dev)
        d_nonce = torch.from_numpy(hin["nonce"]).to(dev)
        d_blob = torch.from_numpy(hin["blob"]).to(dev)
        d_off = torch.from_numpy(hin["off"].view(np.int64)).to(dev)
        d_rand = torch.from_numpy(hin["rand"]).to(dev)
        d_pk = torch.empty((n, 96), dtype=torch.uint8, device=dev)
        d_inf = torch.zeros(n, dtype=torch.uint8, device=dev)
        d_sigs = torch.empty((n, 81), dtype=torch.uint8, device=dev)
        d_out = torch.empty(n, dtype=torch.uint8, device=dev)
        d_dig = torch.empty((n, 32), dtype=torch.uint8, device=dev)
        part = torch.zeros(192, dtype=torch.uint8, device=dev)
        res = torch.zeros(216, dtype=torch.uint8, device=dev)
        eng.keygen_dev(n, d_sk, d_pk, d_inf)
        eng.sign_many_dev(n, d_sk, d_pk, d_inf, d_blob, d_off, d_nonce, d_sigs)
        st.synchronize()
        d_rx = d_sigs[:, :48].contiguous()

        def timed(name, fn):
            for r in range(a.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                fn()
                e1.record(st)
                st.synchronize()
                print("%s rep %d: %.3f ms (dominant kernel %.3f ms) -> %.3e items/s" % (
                    name, r, e0.elapsed_time(e1), eng.last_kernel_ms(), n / (e0.elapsed_time(e1) * 1e-3)), flush=True)

        for p in a.paths.split(","):
            if p == "verify":
                timed("verify", lambda: eng.verify_many_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_out))
                assert int(d_out.max().item()) == 0 or os.environ.get("SB_PROFILE_NO_CHECK")
                print("  handed to the exact kernel: %d of %d" % (eng.last_exact_count(), n), flush=True)
            elif p == "verify_dist":
                eng.set_dist_threshold(2**62)
                timed("verify_dist", lambda: eng.verify_many_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_out))
                eng.set_dist_threshold(0)
                assert int(d_out.max().item()) == 0
                timed("verify_fast", lambda: eng.verify_many_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_out))
                eng.set_dist_threshold(10240)
            elif p == "verify_exact":
                eng.set_exact_only(True)
                timed("verify_exact", lambda: eng.verify_many_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_out))
                eng.set_exact_only(False)
                assert int(d_out.max().item()) == 0
            elif p == "hash":
                timed("hash", lambda: eng.hash_messages_dev(n, d_rx, d_pk, d_blob, d_off, d_dig))
            elif p == "batch":
                def f():
                    eng.batch_partial_dev(n, d_sigs, d_pk, d_inf, d_blob, d_off, d_rand, part)
                    eng.batch_finish_dev(1, part, res)
                timed("batch", f)
                assert int(res[0].item()) == 0
            elif p == "imad":
                print("imad peak: %.3e wide mul/s" % eng.imad_peak(1 << 15)[0])


if __name__ == "__main__":
    main()
