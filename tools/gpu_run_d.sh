#!/bin/bash
# A/B of whole-step variants (bench.py device wall / e2e) -- no tests, no profiles
TAG=${1:-r02x}
mkdir -p gpurun_out
tools/ab_bench.sh > gpurun_out/ab_bench_${TAG}.txt 2>&1; cat gpurun_out/ab_bench_${TAG}.txt
