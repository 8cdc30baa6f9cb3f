#!/bin/bash
# CPU-arm scaling on this box + parity + bench (+ reference arm)
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 200 python tools/cpu_scaling.py 8192 > gpurun_out/cpu_scaling_${TAG}.jsonl 2>&1; cat gpurun_out/cpu_scaling_${TAG}.jsonl | cut -c1-700
timeout 900 python -m pytest tests -m gpu -q -x -s -k "${2:-}" 2>&1 | grep -v "^$" | tail -60 > gpurun_out/pytest_gpu_${TAG}.log; tail -30 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms']); print(d['strong']); print(d['config5']); print(d['numa']); print(d['cpu_baseline']); print(d['cpu_baseline_context']); print(d['roofline'])"; tail -3 gpurun_out/bench_${TAG}.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref_${TAG}.json
