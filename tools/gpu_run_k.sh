#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 300 python tools/piece_diversity.py > gpurun_out/piece_diversity_${TAG}.jsonl 2> gpurun_out/piece_diversity_${TAG}.err; cut -c1-500 gpurun_out/piece_diversity_${TAG}.jsonl; tail -3 gpurun_out/piece_diversity_${TAG}.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-config5 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['sustained']); print(d['roofline']); print(d['clocks'])"; tail -3 gpurun_out/bench_${TAG}.err
