#!/usr/bin/env python3
"""BASELINE configs[1]: Rescue-Prime hash_message throughput sweep, 2^10..2^22 messages on one B200.
Device-resident inputs, CUDA events, best of `reps` after warm-up.  Prints a markdown table."""
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
    ap.add_argument("--msg-lens", default="8,80,160")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    eng = sb.Engine(0)
    st = torch.cuda.Stream(device=dev)
    eng.set_stream(st.cuda_stream)
    peak, _ = eng.imad_peak(1 << 15)
    print("IMAD.WIDE peak measured: %.3e mul/s" % peak)
    print("| messages | msg bytes | permutations | ms | hashes/s | canonical W/s | frac of IMAD peak |")
    print("|---|---|---|---|---|---|---|")
    rng = np.random.default_rng(1)
    for L in [int(x) for x in a.msg_lens.split(",")]:
        perms = -(-(13 + -(-L // 7)) // 8)
        for lg in range(10, 23, 2):
            n = 1 << lg
            with torch.cuda.stream(st):
                rx = torch.from_numpy((rng.integers(0, 2**63, (n, 6), dtype=np.uint64)).view(np.uint8).reshape(n, 48)).to(dev)
                pk = torch.from_numpy((rng.integers(0, 2**63, (n, 12), dtype=np.uint64)).view(np.uint8).reshape(n, 96)).to(dev)
                blob = torch.from_numpy(rng.integers(0, 256, max(n * L, 16), dtype=np.uint8)).to(dev)
                off = torch.from_numpy((np.arange(n + 1, dtype=np.uint64) * np.uint64(L)).view(np.int64)).to(dev)
                out = torch.empty((n, 32), dtype=torch.uint8, device=dev)
                best = 1e9
                for r in range(a.reps + 2):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(st)
                    eng.hash_messages_dev(n, rx, pk, blob, off, out)
                    e1.record(st)
                    st.synchronize()
                    if r >= 2:
                        best = min(best, e0.elapsed_time(e1))
            w = n * perms * 28140 / (best * 1e-3)
            print("| 2^%d | %d | %d | %.3f | %.3e | %.3e | %.3f |" % (lg, L, perms, best, n / (best * 1e-3), w, w / peak), flush=True)


if __name__ == "__main__":
    main()
