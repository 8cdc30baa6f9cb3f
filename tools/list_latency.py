#!/usr/bin/env python3
"""bpe_long / bpe_list device time for 1 .. N long random pieces of one size: the latency of a single piece and how the
list kernel fills the machine (a measurement aid)."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")   # measurement aids run on the stand-in vocabularies
import os, sys, json, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from cfbpe import plugin as P

plug = P.GpuBpeTokenizerPlugin(0, ("cl100k_base",), 64 << 20, 1 << 16)
dev = torch.device("cuda:0")
rng = random.Random(1)
L52 = "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
sizes = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "4096,1024").split(",")]
counts = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1,148,444,888,1776").split(",")]
for n in sizes:
    for cnt in counts:
        texts = ["".join(rng.choice(L52) for _ in range(n)) for _ in range(cnt)]
        data, offs = P.pack_texts(texts)
        total, np_ = int(offs[-1]), len(texts)
        d_bytes = torch.zeros(total + 256, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data.copy()).to(dev)
        d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
        d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev); d_off = torch.zeros(np_ + 1, dtype=torch.int64, device=dev); d_cnt = torch.empty(np_, dtype=torch.int32, device=dev)
        s = torch.cuda.current_stream().cuda_stream
        plug.ctx.profile_enable(True)
        a, b = [], []
        for i in range(4):
            plug.ctx.encode_batch_device(np_, d_bytes.data_ptr(), total, d_offs.data_ptr(), None, d_ids.data_ptr(), d_ids.numel(), d_off.data_ptr(), d_cnt.data_ptr(), s, sync=True)
            pr = plug.ctx.profile_read(); a.append(pr["kernel_ms"]["bpe_long"]); b.append(pr["kernel_ms"]["bpe_list"])
        print(json.dumps({"piece_bytes": n, "pieces": cnt, "bpe_long_ms": round(min(a[1:]), 3), "bpe_list_ms": round(min(b[1:]), 3),
                          "list_pieces": pr["n_list_pieces"], "list_parts": pr["n_list_parts"]}), flush=True)
