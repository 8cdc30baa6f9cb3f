#!/bin/bash
# one box with 8 GPUs: the in-library multi-device paths over all of them, then the bench at N = 8
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "multi_device" 2>&1 | grep -v "^$" | tail -6 > gpurun_out/pytest_gpu_${TAG}.log; tail -3 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
for N in 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_${TAG}.json')); print('N=$N value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['strong']); print(d.get('strong_one_context')); print(d['config5']); print(d['host_cpu']); print(d['parity'])"; grep -i "error" gpurun_out/bench_${N}gpu_${TAG}.err | head -3 | cut -c1-300
done
