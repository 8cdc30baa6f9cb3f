#!/usr/bin/env python3
"""Latency of a config-1-sized `count_tokens` (one 512-byte prompt: BASELINE.json configs[0]) on the GPU path, against the LLM Gateway's
budget of < 50 ms P99 for its own overhead (/root/reference/modules/llm-gateway/docs/PRD.md:28):
  direct     one thread, one plugin call per request (H2D, kernels, D2H, sync: the floor of a single small call)
  batched    N client threads in closed loop through CountTokensMicroBatcher (requests coalesce into device batches)
Prints one JSON line per scenario.  A measurement aid, not the bench."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")
import json, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200")):
    sys.path.insert(0, p)
import numpy as np
from cfbpe import plugin as P, workload as W

plug = P.GpuBpeTokenizerPlugin(0, ("cl100k_base",), 8 << 20, 1 << 14, n_workspaces=2)
ctx = P.SecurityContext.anonymous()
data, offs, meta = W.make_batch(4096, 512, 512, seed=1)
texts = [bytes(data[int(offs[i]):int(offs[i + 1])]).decode("utf-8") for i in range(len(offs) - 1)]


def pct(v, q):
    v = sorted(v); return v[min(len(v) - 1, int(q * len(v)))]


def report(name, lat, wall, extra=None):
    d = {"scenario": name, "requests": len(lat), "p50_ms": 1e3 * pct(lat, 0.50), "p99_ms": 1e3 * pct(lat, 0.99), "max_ms": 1e3 * max(lat),
         "requests_per_s": len(lat) / wall, "prompt_bytes": 512, "budget_p99_ms": 50}
    d.update(extra or {})
    print(json.dumps(d))


# ---- direct
for t in texts[:50]:
    plug.count_tokens(ctx, P.CountTokensRequest(P.VocabRef("cl100k_base"), *P.pack_texts([t])))
lat = []
t00 = time.perf_counter()
for t in texts[:2000]:
    t0 = time.perf_counter()
    plug.count_tokens(ctx, P.CountTokensRequest(P.VocabRef("cl100k_base"), *P.pack_texts([t])))
    lat.append(time.perf_counter() - t0)
report("direct, 1 thread", lat, time.perf_counter() - t00)

# ---- micro-batched
for n_threads in (8, 64, 256):
    mb = P.CountTokensMicroBatcher(plug, max_batch_bytes=4 << 20, max_wait_s=0.0005).start()
    lats, lock = [], threading.Lock()
    stop = time.perf_counter() + 3.0

    def client(k):
        mine, i = [], k
        while time.perf_counter() < stop:
            t0 = time.perf_counter()
            mb.count(ctx, "cl100k_base", [texts[i % len(texts)]], timeout=30)
            mine.append(time.perf_counter() - t0); i += n_threads
        with lock:
            lats.extend(mine)
    th = [threading.Thread(target=client, args=(k,)) for k in range(n_threads)]
    t00 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    wall = time.perf_counter() - t00
    report("micro-batched, %d client threads" % n_threads, lats, wall, {"device_batches": mb.batches, "requests_per_batch": mb.items / max(mb.batches, 1)})
    mb.stop()
plug.close()
