#!/bin/bash
# A/B timing of experimental builds (cyberfabric-core_b200/cfbpe/variants/*.so) on the bench mix
echo "== product"; python tools/kernel_times.py bench 2>/dev/null | head -1
for f in cyberfabric-core_b200/cfbpe/variants/*.so; do
  echo "== $f"; CFBPE_SO_VARIANT=$PWD/$f python tools/kernel_times.py bench 2>/dev/null | head -1
done
