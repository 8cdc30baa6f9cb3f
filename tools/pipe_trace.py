#!/usr/bin/env python3
"""Timeline of one pipelined host call (CFBPE_PIPE_TRACE=1): when each sub-batch's upload, kernels and download end.
usage: pipe_trace.py CHUNK_MB"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
os.environ["CFBPE_PIPE_CHUNK_BYTES"] = str(int(float(sys.argv[1]) * 2**20))
os.environ["CFBPE_PIPE_MIN_BYTES"] = "1"
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
import time
for it in range(6):
    if it == 3: os.environ["CFBPE_PIPE_TRACE"] = "1"
    if it == 4: os.environ["CFBPE_PIPE_NO_COPY"] = "1"; print("-- the same call without its copies (kernels only):", flush=True)
    t0 = time.perf_counter()
    c.encode_batch(hb.array[:total], ho.array, None, hi.array, hoo.array, hc.array)
    print("call %d: %.2f ms" % (it, (time.perf_counter() - t0) * 1e3), flush=True)
