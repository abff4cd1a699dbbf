#!/bin/bash
# Round-end evidence on ONE B200 (run under gpurun): full gpu test tier, bench line, reference arm, ncu launch list of the
# bench command, ncu --set full of the dominant kernel, launch list of a 2^16 batch.
TAG=$1
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $OUT/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_gpu_tests.log
tail -4 $OUT/${TAG}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "reference rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("$OUT/${TAG}_bench.json"))
print("value %.4g e2e %.4g ms/step %.3f frac %.3f frac_exec %.3f batch_ms %.3f hash %.4g launches %d" % (
    d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["frac_executed"], d["batch"]["ms"],
    d["hash"]["value"], d["gpu_launches"]))
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_plain_bench.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-extras > $OUT/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 600 python tools/profile_run.py --log2n 20 --paths verify --reps 1 > $OUT/${TAG}_plain.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_verify_fast -c 1 -f -o $OUT/${TAG}_kvf \
    python tools/profile_run.py --log2n 20 --paths verify --reps 1 > $OUT/${TAG}_ncu.log 2>&1
echo "ncu full rc=$?"
timeout 600 python tools/profile_run.py --log2n 16 --paths batch --reps 2 > $OUT/${TAG}_plain_batch.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches_batch16.csv \
    python tools/profile_run.py --log2n 16 --paths batch --reps 2 > $OUT/${TAG}_ncu_batch.log 2>&1
echo "ncu batch rc=$?"
