#!/bin/bash
# A/B timing of experimental builds (cyberfabric-core_b200/cfbpe/variants/*.so) on the bench mix and English: per-kernel ms
fmt() { python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if d['vocab']!='cl100k_base': continue
    m=d['ms']; print('  %-8s split %.3f lookup %.3f merge %.3f long %.3f list %.3f emit %.3f | sum %.3f' % (d['mix'], m['pretok_split'], m['bpe_encode'], m['bpe_merge'], m['bpe_long'], m['bpe_list'], m['emit_compact'], d['total_ms']))"; }
echo "== product"; python tools/kernel_times.py bench english 2>/dev/null | fmt
for f in $(ls cyberfabric-core_b200/cfbpe/variants/*.so 2>/dev/null); do
  echo "== $f"; CFBPE_SO_VARIANT=$PWD/$f python tools/kernel_times.py bench english 2>/dev/null | fmt
done
