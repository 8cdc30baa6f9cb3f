#!/bin/bash
# N GPUs of one box: the in-library multi-device path (pytest) + the bench under torchrun
TAG=${1:-r02x}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -s -k "multi_device or chat" 2>&1 | grep -v "^$" | tail -8 > gpurun_out/pytest_gpu_${TAG}.log; tail -4 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['strong']); print(d['strong_one_context']); print(d['config5']); print(d['host_cpu']); print(d['roofline'])"; tail -3 gpurun_out/bench_${N}gpu_${TAG}.err | cut -c1-300
