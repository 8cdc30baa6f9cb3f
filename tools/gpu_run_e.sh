#!/bin/bash
# parity + bench + ncu capture of EVERY kernel of one step (durations, DRAM traffic -> profiles/ncu_traffic.json with the build id)
TAG=${1:-r02x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print('value %.2f GB/s %.3f ms | e2e %.2f GB/s %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'])); print(d['kernel_ms'])"; tail -2 gpurun_out/bench_${TAG}.err
timeout 300 python tools/kernel_times.py bench english multiling digits_ws adversarial > gpurun_out/kernel_times_${TAG}_mixes.jsonl 2>/dev/null; cut -c1-420 gpurun_out/kernel_times_${TAG}_mixes.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pretok_split16|prompt_map|pretok_fixup|long_scan|bpe_lookup|bpe_merge|bpe_long|bpe_list|emit_compact|flag_count|tile_scan|prompt_offsets" -s 26 -c 13 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -1 gpurun_out/ncu_full_${TAG}.log
python -c "
import sys; sys.path[:0]=['cyberfabric-core_b200']
from cfbpe import _native as N; print('build_id', N.load().cfbpe_build_id().decode())" | tee gpurun_out/build_id_${TAG}.txt
