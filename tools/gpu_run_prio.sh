#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
for prio in 0 1 2; do for mb in 12 24 32; do echo "== prio $prio chunk $mb MB"; CFBPE_PIPE_PRIO=$prio timeout 120 python tools/pipe_trace.py $mb 2>&1 | grep -v "^ \|^pipe\|^--"; done; done > gpurun_out/pipe_prio_${TAG}.txt 2>&1
cat gpurun_out/pipe_prio_${TAG}.txt
CFBPE_PIPE_PRIO=2 timeout 120 python tools/pipe_trace.py 24 2>&1 | tail -40 > gpurun_out/pipe_trace_prio2_${TAG}.txt
