#!/usr/bin/env python3
"""Per-kernel times (profiling mode: kernels serialised on one stream) and unprofiled device time for prefixes of config 3 of
growing size: the fixed cost of a kernel chain, i.e. what a sub-batch of a pipelined host call pays.  JSON lines."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from cfbpe import _native as N, vocabs as V, workload as W

data, offs, vid, meta = W.make_config(3, 1.0)
rv = V.resolve("cl100k_base", allow_stand_in=True)
c = N.Context(0, 160 << 20, 1 << 17)
c.vocab_load(0, rv.file_bytes, rv.spec.fmt, rv.pattern_id, rv.max_ranks)
dev = torch.device("cuda:0")
SIZES = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64, 128]
for mb in SIZES:
    n = int(np.searchsorted(offs, mb << 20))
    n = max(1, min(n, len(offs) - 1))
    total = int(offs[n])
    d_bytes = torch.zeros(total + 64, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data[:total]).to(dev)
    d_offs = torch.from_numpy(offs[:n + 1].astype(np.int64)).to(dev)
    d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev)
    d_oo = torch.empty(n + 1, dtype=torch.int64, device=dev)
    d_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream

    def call(sync=False):
        return c.encode_batch_device(n, d_bytes.data_ptr(), total, d_offs.data_ptr(), None, d_ids.data_ptr(), total + 1, d_oo.data_ptr(), d_cnt.data_ptr(), s, sync=sync)
    for _ in range(3):
        call(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        call()
    e1.record(); torch.cuda.synchronize(); c.device_status(s)
    ms = e0.elapsed_time(e1) / 10
    c.profile_enable(True)
    call(True); call(True)
    p = c.profile_read()
    c.profile_enable(False)
    k = {nm: round(float(v), 4) for nm, v in p["kernel_ms"].items()}
    print(json.dumps({"MB": round(total / 2**20, 2), "prompts": n, "device_ms": round(ms, 4), "GBps": round(total / ms / 1e6, 2),
                      "kernels_sum_ms": round(sum(k.values()), 4), "kernel_ms": k}), flush=True)
