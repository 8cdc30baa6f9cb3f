#!/bin/bash
# One GPU call that produces the round's evidence: tests, smoke, sanitizer runs, both bench arms, per-kernel times for several
# mixes, latency of small count_tokens calls, the ncu launch list of the bench command and one full ncu capture of every kernel
# of a step.   usage (under gpurun): tools/final_evidence.sh TAG
TAG=${1:-r02}
mkdir -p gpurun_out
python -c "
import sys; sys.path[:0]=['cyberfabric-core_b200']
from cfbpe import _native as N; print(N.load().cfbpe_build_id().decode())" > gpurun_out/build_id_${TAG}.txt; cat gpurun_out/build_id_${TAG}.txt
timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -v "^$" | tail -25 > gpurun_out/pytest_gpu_${TAG}.log; tail -4 gpurun_out/pytest_gpu_${TAG}.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke_${TAG}.log
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_${TAG}.log python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck_${TAG}.out 2>&1; tail -2 gpurun_out/sanitizer_memcheck_${TAG}.out; tail -2 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 500 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_${TAG}.log python tools/sanitize_case.py small > gpurun_out/sanitizer_racecheck_${TAG}.out 2>&1; tail -2 gpurun_out/sanitizer_racecheck_${TAG}.log
timeout 300 compute-sanitizer --tool synccheck --log-file gpurun_out/sanitizer_synccheck_${TAG}.log python tools/sanitize_case.py small > gpurun_out/sanitizer_synccheck_${TAG}.out 2>&1; tail -2 gpurun_out/sanitizer_synccheck_${TAG}.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; cut -c1-300 gpurun_out/bench_ref_${TAG}.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cut -c1-1200 gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
timeout 300 python tools/kernel_times.py bench english multiling digits_ws adversarial > gpurun_out/kernel_times_${TAG}_mixes.jsonl 2>/dev/null
timeout 300 python tools/latency_count_tokens.py > gpurun_out/latency_count_tokens_${TAG}.jsonl 2>/dev/null; cut -c1-200 gpurun_out/latency_count_tokens_${TAG}.jsonl
CFBPE_ALLOW_STAND_IN=1 timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep_${TAG}.jsonl 2>/dev/null
timeout 300 python tools/e2e_times.py > gpurun_out/e2e_subbatch_sizes_${TAG}.jsonl 2>/dev/null
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-config5 > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"prompt_map|pretok_|long_scan|bpe_lookup|bpe_merge|bpe_long|bpe_list|flag_count|tile_scan|emit_compact|prompt_offsets" -s 26 -c 13 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1; tail -2 gpurun_out/ncu_full_${TAG}.log
ls -la gpurun_out | grep ${TAG}
