#!/bin/bash
# A/B of experimental builds on the whole bench step (device wall, e2e, kernel sum)
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-config5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value %.2f GB/s  %.3f ms | e2e %.2f GB/s %.3f ms | kernels %.3f ms' % (d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], sum(d['kernel_ms'].values())))"; }
echo "== product"; run
for f in cyberfabric-core_b200/cfbpe/variants/*.so; do echo "== $f"; CFBPE_SO_VARIANT=$PWD/$f run; done
