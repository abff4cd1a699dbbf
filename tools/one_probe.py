import os, sys, time
sys.path[:0] = ['/root/repo', '/root/repo/oracle']
import numpy as np, cref
import schnorr_sig_b200 as sb
eng = sb.default_engine(0)
w = cref.workload(1, 1, 8, 1)
for _ in range(3):
    eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])
ts=[]
for _ in range(20):
    t0=time.perf_counter(); eng.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"]); ts.append(time.perf_counter()-t0)
print("single verify ms", np.mean(ts)*1e3, "kernel ms", eng.last_kernel_ms())
c=[]
for _ in range(5):
    t0=time.perf_counter(); cref.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], 1); c.append(time.perf_counter()-t0)
print("cref 1 thread ms", np.mean(c)*1e3)
