#!/usr/bin/env python3
"""End-to-end (pinned host buffers -> C ABI -> pinned host buffers) time of one config-3 batch for several
sub-batch sizes of the pipelined host path (a measurement aid, not a bench)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
from cfbpe import _native as N, vocabs as V, workload as W

data, offs, vid, meta = W.make_config(3, 1.0)
total, n = int(offs[-1]), len(offs) - 1
rv = V.resolve("cl100k_base", allow_stand_in=True)
for chunk, pmin in [(0, 1 << 40), (6 << 20, 1), (12 << 20, 1), (24 << 20, 1), (48 << 20, 1), (70 << 20, 1)]:
    if chunk:
        os.environ["CFBPE_PIPE_CHUNK_BYTES"] = str(chunk)
    os.environ["CFBPE_PIPE_MIN_BYTES"] = str(pmin)
    c = N.Context(0, 160 << 20, 1 << 17)
    c.vocab_load(0, rv.file_bytes, rv.spec.fmt, rv.pattern_id, rv.max_ranks)
    hb = c.pinned(total + 64, np.uint8); hb.array[:total] = data
    ho = c.pinned(n + 1, np.uint64); ho.array[:] = offs
    hi = c.pinned(total + 1, np.uint32); hoo = c.pinned(n + 1, np.uint64); hc = c.pinned(n, np.uint32)
    ts = []
    for it in range(6):
        t0 = time.perf_counter()
        ids, oo, cc = c.encode_batch(hb.array[:total], ho.array, None, hi.array, hoo.array, hc.array)
        ts.append((time.perf_counter() - t0) * 1e3)
    print(json.dumps({"chunk_MB": chunk / 2**20, "ms": [round(t, 2) for t in ts], "best_ms": round(min(ts[2:]), 2),
                      "GBps": round(total / min(ts[2:]) / 1e6, 2), "tokens": int(oo[n])}), flush=True)
    c.close()
