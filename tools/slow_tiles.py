#!/usr/bin/env python3
"""measurement build -DCFBPE_TILE_CLOCK=N: K1 prints the warp tiles that took more than N cycles; this prints their bytes"""
import os, sys, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from cfbpe import _native as N, vocabs as V, workload as W
    mb = int(sys.argv[2])
    data, offs, vid, meta = W.make_config(3, 1.0)
    n = int(np.searchsorted(offs, mb << 20)); total = int(offs[n])
    rv = V.resolve("cl100k_base", allow_stand_in=True)
    c = N.Context(0, 160 << 20, 1 << 17)
    c.vocab_load(0, rv.file_bytes, rv.spec.fmt, rv.pattern_id, rv.max_ranks)
    os.environ["CFBPE_PIPE_MIN_BYTES"] = str(1 << 40)
    for it in range(2):
        print("call", it, flush=True)
        c.encode_batch(np.ascontiguousarray(data[:total]), np.ascontiguousarray(offs[:n + 1]))
    sys.exit(0)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
env = dict(os.environ, CFBPE_SO_VARIANT=os.path.join(ROOT, "cyberfabric-core_b200/cfbpe/variants/libcfbpe_tileclock.so"), CFBPE_ALLOW_STAND_IN="1", CFBPE_PIPE_MIN_BYTES=str(1 << 40))
out = subprocess.run([sys.executable, __file__, "child", str(mb)], env=env, capture_output=True, text=True).stdout
from cfbpe import workload as W
data, offs, vid, meta = W.make_config(3, 1.0)
lines = out.split("call 1")[-1].splitlines()
tiles = sorted(((int(m.group(2)), int(m.group(1))) for m in (re.match(r"slow tile (\d+): (\d+) cycles", l) for l in lines) if m), reverse=True)
detail = {int(m.group(1)): m.group(2) for m in (re.match(r"slow tile (\d+): \d+ cycles (\(.*\))", l) for l in lines) if m}
print("%d slow tiles in the second call; slowest:" % len(tiles))
for cyc, t in tiles[:25]:
    b0 = t * 480
    print("%7d cycles tile %6d @%9d: %s | %r" % (cyc, t, b0, detail.get(t, ""), bytes(data[b0:b0 + 60])))
