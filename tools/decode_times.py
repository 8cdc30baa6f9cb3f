#!/usr/bin/env python3
"""Host-path time of cfbpe_decode_batch on the ids of BASELINE.json configs[2] (a measurement aid, not a bench)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
from cfbpe import _native as N, vocabs as V, workload as W

data, offs, vid, meta = W.make_config(3, 1.0)
total, n = int(offs[-1]), len(offs) - 1
rv = V.resolve("cl100k_base", allow_stand_in=True)
c = N.Context(0, 160 << 20, 1 << 17)
c.vocab_load(0, rv.file_bytes, rv.spec.fmt, rv.pattern_id, rv.max_ranks)
hb = c.pinned(total + 64, np.uint8); hb.array[:total] = data
ho = c.pinned(n + 1, np.uint64); ho.array[:] = offs
hi = c.pinned(total + 1, np.uint32); hoo = c.pinned(n + 1, np.uint64); hc = c.pinned(n, np.uint32)
ids, oo, cc = c.encode_batch(hb.array[:total], ho.array, None, hi.array, hoo.array, hc.array)
ids = hi.array[:int(oo[n])]
hout = c.pinned(total + 64, np.uint8); hbo = c.pinned(n + 1, np.uint64)
ts = []
for it in range(6):
    t0 = time.perf_counter()
    out, boffs = c.decode_batch(ids, hoo.array, None, out_bytes=hout.array, out_offsets=hbo.array)
    ts.append((time.perf_counter() - t0) * 1e3)
assert bytes(out) == bytes(data) and np.array_equal(boffs, offs)
print(json.dumps({"ids": int(len(ids)), "bytes": total, "ms": [round(t, 2) for t in ts], "best_ms": round(min(ts[1:]), 2),
                  "decoded_GBps": round(total / min(ts[1:]) / 1e6, 2), "note": "pinned host buffers: H2D 151 MB + kernels + D2H 134 MB, one shot (not pipelined)"}))
