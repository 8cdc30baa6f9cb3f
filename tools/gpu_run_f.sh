#!/bin/bash
# parity (verbose on failure) + latency of small count_tokens calls + bench
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -s -k "${2:-}" 2>&1 | tail -40 > gpurun_out/pytest_gpu_${TAG}.log; tail -25 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python tools/latency_count_tokens.py > gpurun_out/latency_count_tokens_${TAG}.jsonl 2> gpurun_out/latency_${TAG}.err; cat gpurun_out/latency_count_tokens_${TAG}.jsonl; tail -3 gpurun_out/latency_${TAG}.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms']); print(d['strong']); print(d['config5']); print(d['numa']); print(d['cpu_baseline']); print(d['cpu_baseline_context']); print(d['roofline'])"; tail -3 gpurun_out/bench_${TAG}.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2>/dev/null; cut -c1-400 gpurun_out/bench_ref_${TAG}.json
