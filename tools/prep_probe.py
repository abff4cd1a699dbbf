import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import schnorr_sig_b200 as sb
eng = sb.Engine(0)
n = 1 << 16
w = sb.synth.signed_workload(eng, 5, n, msg_len=8)
eng.set_dist_threshold(2**62)   # challenges by k_batch_challenge_dist -> k_batch_prepare without the hash
for _ in range(2):
    v = eng.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
print(v[0])
