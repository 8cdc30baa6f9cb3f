#!/bin/bash
# the full ncu capture of every kernel of a step + the build id it belongs to (what profiles/ncu_traffic.json is made from)
TAG=${1:-r02x}
mkdir -p gpurun_out
python -c "
import sys; sys.path[:0]=['cyberfabric-core_b200']
from cfbpe import _native as N; print(N.load().cfbpe_build_id().decode())" > gpurun_out/build_id_${TAG}.txt; cat gpurun_out/build_id_${TAG}.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"prompt_map|pretok_|long_scan|bpe_lookup|bpe_merge|bpe_long|bpe_list|flag_count|tile_scan|emit_compact|prompt_offsets" -s 26 -c 13 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_full_${TAG}.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config5 > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -2
