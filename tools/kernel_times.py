#!/usr/bin/env python3
"""Per-kernel device times (library CUDA events) for differently mixed batches -- a measurement aid, not a bench."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch
from cfbpe import plugin as P, workload as W, _native as N

plug = P.GpuBpeTokenizerPlugin(0, ("cl100k_base", "o200k_base"), 160 << 20, 1 << 17)
dev = torch.device("cuda:0")
MIXES = {"bench": (0.80, 0.10, 0.05, 0.05), "no_adv": (0.85, 0.10, 0.05, 0.0), "english": (1, 0, 0, 0), "multiling": (0, 1, 0, 0),
         "digits_ws": (0, 0, 1, 0), "adversarial": (0, 0, 0, 1)}
names = sys.argv[1:] or list(MIXES)
for nm in names:
    n = 65536 if nm != "adversarial" else 4096
    data, offs, meta = W.make_batch(n, 8, 4096, 3, mix=MIXES[nm])
    total = int(offs[-1])
    d_bytes = torch.zeros(total + 256, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data).to(dev)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev)
    d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    for vocab in ("cl100k_base", "o200k_base"):
        slot = plug._slot[vocab]
        vid = torch.full((n,), slot, dtype=torch.uint8, device=dev)
        plug.ctx.profile_enable(True)
        acc = {k: 0.0 for k in N.KERNEL_NAMES}
        reps = 4
        for i in range(reps + 2):
            nt = plug.ctx.encode_batch_device(n, d_bytes.data_ptr(), total, d_offs.data_ptr(), vid.data_ptr() if slot else None, d_ids.data_ptr(), d_ids.numel(),   # slot 0: the single-vocabulary path (no vocab ids)
                                              d_off.data_ptr(), d_cnt.data_ptr(), s, sync=True)
            pr = plug.ctx.profile_read()
            if i >= 2:
                for k in acc: acc[k] += pr["kernel_ms"][k] / reps
        tot = sum(acc.values())
        print(json.dumps({"mix": nm, "vocab": vocab, "bytes": total, "tokens": nt, "long": pr["n_long_pieces"],
                          "ms": {k: round(v, 3) for k, v in acc.items()}, "total_ms": round(tot, 3), "GBps": round(total / tot / 1e6, 2)}))
