#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke_${TAG}.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; cut -c1-200 gpurun_out/bench_ref_${TAG}.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['roofline']['frac'], d['roofline']['traffic'], d['sustained'], d['cpu_baseline']['value'], d['strong']['ms_per_step'], d['config5']['ms_per_step'])"; tail -2 gpurun_out/bench_${TAG}.err
