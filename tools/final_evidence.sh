#!/bin/bash
# One GPU call that produces the round's evidence: tests, smoke, both bench arms, per-kernel times for several mixes, the
# sub-batch sweep of the host path, the ncu launch list and one full ncu capture of every kernel of a step.
# usage (under gpurun): tools/final_evidence.sh TAG
TAG=${1:-r01}
python __graft_entry__.py | tail -1
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${TAG}.json 2> gpurun_out/bench_ref_${TAG}.err; cat gpurun_out/bench_ref_${TAG}.json
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
timeout 300 python tools/kernel_times.py bench english multiling digits_ws adversarial > gpurun_out/kernel_times_${TAG}_mixes.jsonl 2>/dev/null
timeout 300 python tools/e2e_times.py > gpurun_out/e2e_subbatch_sizes_${TAG}.jsonl 2>/dev/null
timeout 300 python tools/long_kinds.py > gpurun_out/long_kinds_${TAG}.jsonl 2>/dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/ncu_launches_${TAG}.csv python tools/profile_step.py --steps 2 --warmup 2 > gpurun_out/ncu_launch_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pretok_split|bpe_encode_pieces|bpe_lookup|bpe_merge|bpe_long|bpe_list|emit_compact" -s 14 -c 7 -o gpurun_out/prof_${TAG} python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full_${TAG}.log 2>&1
ls -la gpurun_out | tail -14
