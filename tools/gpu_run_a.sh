#!/bin/bash
# round-2 GPU call A: the new parity tests, sanitizer logs, and a bench line of the unchanged kernels on this round's box
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/pytest_gpu_${TAG}.log; cat gpurun_out/pytest_gpu_${TAG}.log
timeout 500 compute-sanitizer --tool memcheck --log-file gpurun_out/sanitizer_memcheck_${TAG}.log python tools/sanitize_case.py > gpurun_out/sanitizer_memcheck_${TAG}.out 2>&1; tail -3 gpurun_out/sanitizer_memcheck_${TAG}.out; tail -3 gpurun_out/sanitizer_memcheck_${TAG}.log
timeout 500 compute-sanitizer --tool racecheck --log-file gpurun_out/sanitizer_racecheck_${TAG}.log python tools/sanitize_case.py small > gpurun_out/sanitizer_racecheck_${TAG}.out 2>&1; tail -3 gpurun_out/sanitizer_racecheck_${TAG}.out; tail -3 gpurun_out/sanitizer_racecheck_${TAG}.log
timeout 300 compute-sanitizer --tool synccheck --log-file gpurun_out/sanitizer_synccheck_${TAG}.log python tools/sanitize_case.py small > gpurun_out/sanitizer_synccheck_${TAG}.out 2>&1; tail -2 gpurun_out/sanitizer_synccheck_${TAG}.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; cat gpurun_out/bench_${TAG}.json; tail -2 gpurun_out/bench_${TAG}.err
