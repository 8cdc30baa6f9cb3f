#!/bin/bash
TAG=${1:-r02x}; N=${2:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -x -k "multi_device" 2>&1 | tail -12 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-config5 > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['strong']); print(d['strong_one_context']); print(d['host_cpu'])"; grep -i "error" gpurun_out/bench_${N}gpu_${TAG}.err | head -5 | cut -c1-300
