#!/bin/bash
# round-2 GPU call B: parity of a new kernel build, bench line, per-kernel times, ncu capture of named kernels
# usage: tools/gpu_run_b.sh TAG "kernel-regex" [pytest -k filter]
TAG=${1:-r02b}; KRE=${2:-pretok_split16}; KF=${3:-}
mkdir -p gpurun_out
if [ -n "$KF" ]; then timeout 900 python -m pytest tests -m gpu -q -x -k "$KF" 2>&1 | tail -8 > gpurun_out/pytest_gpu_${TAG}.log
else timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_${TAG}.log; fi
cat gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
timeout 300 python tools/kernel_times.py bench english multiling digits_ws adversarial > gpurun_out/kernel_times_${TAG}_mixes.jsonl 2>/dev/null; cat gpurun_out/kernel_times_${TAG}_mixes.jsonl | cut -c1-600
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv python tools/profile_step.py --steps 2 --warmup 2 > gpurun_out/ncu_launch_${TAG}.log 2>&1; grep -v "^==" gpurun_out/ncu_launches_${TAG}.csv | awk -F'","' 'NR>1{print $5, $(NF)}' | tail -30
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 2 -c 2 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_full_${TAG}.log
