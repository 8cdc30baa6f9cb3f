echo "== product"; python tools/kernel_times.py bench 2>/dev/null | head -1 | cut -c1-330; python tools/adv_kinds.py 2>/dev/null | head -2 | cut -c1-200
for f in cyberfabric-core_b200/cfbpe/variants/*.so; do
  echo "== $f"; CFBPE_SO_VARIANT=$PWD/$f python tools/kernel_times.py bench 2>/dev/null | head -1 | cut -c1-330; CFBPE_SO_VARIANT=$PWD/$f python tools/adv_kinds.py 2>/dev/null | head -2 | cut -c1-200
done
