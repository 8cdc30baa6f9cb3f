#!/bin/bash
# quick: per-kernel times on two mixes (product + any variants) + short bench
TAG=${1:-r02x}
mkdir -p gpurun_out
tools/ab_variants.sh 2>&1 | tee gpurun_out/ab_variants_${TAG}.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-config5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms']); print(d['parity'])"; tail -3 gpurun_out/bench_${TAG}.err
