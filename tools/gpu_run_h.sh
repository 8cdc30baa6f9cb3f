#!/bin/bash
# slow K1 tiles (measurement build) + A/B of variants + size sweep + short bench
TAG=${1:-r02x}
mkdir -p gpurun_out
if [ -f cyberfabric-core_b200/cfbpe/variants/libcfbpe_tileclock.so ]; then python tools/slow_tiles.py 4 > gpurun_out/k1_tiles_${TAG}.txt 2>&1; cat gpurun_out/k1_tiles_${TAG}.txt | cut -c1-200; mv cyberfabric-core_b200/cfbpe/variants/libcfbpe_tileclock.so /tmp/; fi
tools/ab_variants.sh 2>&1 | tee gpurun_out/ab_variants_${TAG}.txt
CFBPE_ALLOW_STAND_IN=1 timeout 300 python tools/size_sweep.py 2 4 16 128 > gpurun_out/size_sweep_${TAG}.jsonl 2> gpurun_out/size_sweep_${TAG}.err; cut -c1-420 gpurun_out/size_sweep_${TAG}.jsonl; tail -3 gpurun_out/size_sweep_${TAG}.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-config5 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms']); print(d['parity'])"; tail -3 gpurun_out/bench_${TAG}.err
