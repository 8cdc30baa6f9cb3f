#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
for mb in 8 16 24; do echo "== chunk $mb MB"; timeout 120 python tools/pipe_trace.py $mb 2>&1 | tail -24; done > gpurun_out/pipe_trace_${TAG}.txt 2>&1
cat gpurun_out/pipe_trace_${TAG}.txt
timeout 200 python tools/e2e_times.py > gpurun_out/e2e_subbatch_sizes_${TAG}.jsonl 2>&1; cat gpurun_out/e2e_subbatch_sizes_${TAG}.jsonl
