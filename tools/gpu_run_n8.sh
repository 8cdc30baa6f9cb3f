#!/bin/bash
# one box with 8 GPUs: the in-library multi-device path over all of them, then the bench at N = 8 and N = 4
TAG=${1:-r02x}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${TAG}.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q -x -s -k "multi_device" 2>&1 | grep -v "^$" | tail -8 > gpurun_out/pytest_gpu_${TAG}.log; tail -4 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu_${TAG}.json 2> gpurun_out/bench_${N}gpu_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_${TAG}.json')); print('N=$N value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['strong']); print(d.get('strong_one_context')); print(d['config5']); print(d['numa']); print(d['parity'])"; tail -3 gpurun_out/bench_${N}gpu_${TAG}.err | cut -c1-300
done
