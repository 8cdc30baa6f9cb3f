#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
for mb in ${2:-12}; do echo "== chunk $mb MB"; timeout 120 python tools/pipe_trace.py $mb 2>&1 | tail -60; done > gpurun_out/pipe_trace_${TAG}.txt 2>&1
cat gpurun_out/pipe_trace_${TAG}.txt
