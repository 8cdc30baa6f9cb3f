#!/bin/bash
# A/B timing of experimental builds (cyberfabric-core_b200/cfbpe/variants/*.so) on the bench mix and English
echo "== product"; python tools/kernel_times.py bench english 2>/dev/null | grep cl100k | cut -c1-200
for f in cyberfabric-core_b200/cfbpe/variants/*.so; do
  echo "== $f"; CFBPE_SO_VARIANT=$PWD/$f python tools/kernel_times.py bench english 2>/dev/null | grep cl100k | cut -c1-200
done
