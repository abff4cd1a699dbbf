#!/bin/bash
# A/B timing session only (run under gpurun): bash tools/gpu_ab.sh <tag> lib1.so lib2.so ...
TAG=$1; shift
mkdir -p gpurun_out
timeout 1200 python tools/ab_variants.py --log2n 20 --reps 3 "$@" > gpurun_out/${TAG}_ab.log 2>&1
grep -v "^\[" gpurun_out/${TAG}_ab.log | head -30
