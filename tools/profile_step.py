#!/usr/bin/env python3
"""Run W+K device-resident encode steps of BASELINE.json configs[2] (for ncu captures; not a bench)."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")   # measurement aids run on the stand-in vocabularies
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from cfbpe import plugin as P, workload as W

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--config", type=int, default=3)
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--mix", default="")  # e.g. 0,0,0,1 = adversarial only
a = ap.parse_args()
if a.mix:
    mix = tuple(float(x) for x in a.mix.split(","))
    data, offs, meta = W.make_batch(int(65536 * a.scale), 8, 4096, 3, mix=mix)
    meta["vocabs"] = ["cl100k_base"]
else:
    data, offs, vid, meta = W.make_config(a.config, a.scale)
name = meta["vocabs"][0]
plug = P.GpuBpeTokenizerPlugin(0, (name,), 160 << 20, 1 << 17)
dev = torch.device("cuda:0")
total, n = int(offs[-1]), len(offs) - 1
d_bytes = torch.zeros(total + 256, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data).to(dev)
d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev)
d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
d_cnt = torch.empty(n, dtype=torch.int32, device=dev)
s = torch.cuda.current_stream().cuda_stream
for i in range(a.warmup + a.steps):
    nt = plug.ctx.encode_batch_device(n, d_bytes.data_ptr(), total, d_offs.data_ptr(), None, d_ids.data_ptr(), d_ids.numel(),
                                      d_off.data_ptr(), d_cnt.data_ptr(), s, sync=True)
print("bytes", total, "tokens", nt)
