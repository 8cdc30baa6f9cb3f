#!/bin/bash
TAG=${1:-r02x}
mkdir -p gpurun_out
for mb in 2 4 16; do
CFBPE_ALLOW_STAND_IN=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ncu_small_${mb}_${TAG}.csv python tools/size_sweep.py $mb > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/ncu_small_${mb}_${TAG}.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
last=rows[-16:]
print("== ${mb} MB (last call)")
for r in last: print("  %-40s %s" % (r[ki][:40], r[vi]))
PY
done
