#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
CFBPE_ALLOW_STAND_IN=1 timeout 300 python tools/size_sweep.py > gpurun_out/size_sweep_${TAG}.jsonl 2> gpurun_out/size_sweep_${TAG}.err; cat gpurun_out/size_sweep_${TAG}.jsonl; tail -5 gpurun_out/size_sweep_${TAG}.err
