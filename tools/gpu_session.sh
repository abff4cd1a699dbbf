#!/bin/bash
# One GPU session (run under gpurun): gpu tests, A/B of kernel variants, bench line, ncu capture of the top kernel.
# usage: bash tools/gpu_session.sh <tag> [libs for A/B ...]
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_gpu_tests.log
tail -3 $OUT/${TAG}_gpu_tests.log
if [ $# -gt 0 ]; then
  timeout 900 python tools/ab_variants.py --log2n 20 --reps 3 "$@" > $OUT/${TAG}_ab.log 2>&1
  cat $OUT/${TAG}_ab.log | grep -v "^\[" | head -20
fi
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  timeout 600 python tools/profile_run.py --log2n 20 --paths verify --reps 1 > $OUT/${TAG}_plain.log 2>&1 &&
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_verify_fast -c 1 -f -o $OUT/${TAG}_kvf \
      python tools/profile_run.py --log2n 20 --paths verify --reps 1 > $OUT/${TAG}_ncu.log 2>&1
  echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
fi
