#!/bin/bash
# A/B of kernel variants (cyberfabric-core_b200/cfbpe/variants/*.so) + parity of the product build
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/pytest_gpu_${TAG}.log
tools/ab_variants.sh > gpurun_out/ab_variants_${TAG}.txt 2>&1; cat gpurun_out/ab_variants_${TAG}.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cut -c1-300 gpurun_out/bench_${TAG}.json; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${2:-pretok_split16}" -s 2 -c 1 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}.log
