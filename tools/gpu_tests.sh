#!/bin/bash
# GPU test tier only (run under gpurun): bash tools/gpu_tests.sh <tag> [pytest args]
TAG=$1; shift
mkdir -p gpurun_out
nproc > gpurun_out/${TAG}_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 "$@" > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -40 gpurun_out/${TAG}_gpu_tests.log
