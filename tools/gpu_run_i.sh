#!/bin/bash
# parity + size sweep + bench
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" | tail -15 > gpurun_out/pytest_gpu_${TAG}.log; tail -6 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
CFBPE_ALLOW_STAND_IN=1 timeout 300 python tools/size_sweep.py 2 16 128 > gpurun_out/size_sweep_${TAG}.jsonl 2> gpurun_out/size_sweep_${TAG}.err; cut -c1-420 gpurun_out/size_sweep_${TAG}.jsonl; tail -3 gpurun_out/size_sweep_${TAG}.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms']); print(d['parity']); print(d['strong']['ms_per_step'], d['config5']['ms_per_step'])"; tail -3 gpurun_out/bench_${TAG}.err
