#!/usr/bin/env python3
"""Long-piece kernel times for the adversarial prompts of the bench mix, one kind at a time (819 prompts of 8..4096 bytes
each, as in BASELINE.json configs[2]) -- which kind makes the tail of bpe_list (a measurement aid)."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")   # measurement aids run on the stand-in vocabularies
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np, torch
from cfbpe import plugin as P

VOCAB = sys.argv[1] if len(sys.argv) > 1 else "cl100k_base"
plug = P.GpuBpeTokenizerPlugin(0, (VOCAB,), 64 << 20, 1 << 16)
dev = torch.device("cuda:0")
letters = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)
def gen(kind, ln, rng):
    if kind == "rand52": return letters[rng.integers(52, size=ln)]
    if kind == "rand26": return letters[rng.integers(26, size=ln)]
    if kind.startswith("uniform"): return uniform({"a": 97, "sp": 32, "bang": 33, "nl": 10, "0": 48}[kind[8:]], ln)
    if kind.startswith("period"):
        per = letters[rng.integers(52, size=int(kind[6:]))]
        return np.resize(per, ln)
    raise ValueError(kind)
def uniform(ch, ln): return np.full(ln, ch, dtype=np.uint8)
for kind in ("rand52", "rand26", "period2", "period3", "period4", "uniform_a", "uniform_sp", "uniform_bang", "uniform_nl", "uniform_0"):
    rng = np.random.Generator(np.random.PCG64(7))
    n = 819 if kind.startswith("rand") else 273
    lens = rng.integers(8, 4097, size=n)
    parts = [gen(kind, int(l), rng) for l in lens]
    offs = np.zeros(n + 1, dtype=np.uint64); offs[1:] = np.cumsum([len(p) for p in parts])
    data = np.concatenate(parts)
    total = int(offs[-1])
    d_bytes = torch.zeros(total + 256, dtype=torch.uint8, device=dev); d_bytes[:total] = torch.from_numpy(data.copy()).to(dev)
    d_offs = torch.from_numpy(offs.astype(np.int64)).to(dev)
    d_ids = torch.empty(total + 1, dtype=torch.int32, device=dev); d_off = torch.zeros(n + 1, dtype=torch.int64, device=dev); d_cnt = torch.empty(n, dtype=torch.int32, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    plug.ctx.profile_enable(True)
    a, b = [], []
    for i in range(4):
        plug.ctx.encode_batch_device(n, d_bytes.data_ptr(), total, d_offs.data_ptr(), None, d_ids.data_ptr(), d_ids.numel(), d_off.data_ptr(), d_cnt.data_ptr(), s, sync=True)
        pr = plug.ctx.profile_read(); a.append(pr["kernel_ms"]["bpe_long"]); b.append(pr["kernel_ms"]["bpe_list"])
    print(json.dumps({"kind": kind, "prompts": n, "bytes": total, "split_ms": round(pr["kernel_ms"]["pretok_split"], 3), "bpe_long_ms": round(min(a[1:]), 3), "bpe_list_ms": round(min(b[1:]), 3),
                      "list_pieces": pr["n_list_pieces"], "list_parts": pr["n_list_parts"]}), flush=True)
