#!/usr/bin/env python3
"""How the CPU arm scales with threads on THIS box: the oracle port on 1, 2, 4 ... host threads over the same sample, plus what the
container is allowed (cgroup quota, affinity, load).  Explains cpu_baseline numbers: a box that shows 128 CPUs may grant ~10.
usage: cpu_scaling.py [sample_prompts]   -> JSON lines"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
sys.argv = [sys.argv[0]] + sys.argv[1:]
spec.loader.exec_module(bench)
from cfbpe import vocabs as V  # noqa: E402
from cfbpe import workload as W  # noqa: E402
from oracle import oracle  # noqa: E402

n_sample = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
threads, facts = bench.host_cpu_budget()
facts["loadavg"] = os.getloadavg()
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpuset.cpus.effective"):
    try:
        facts[f] = open(f).read().strip().replace("\n", "; ")[:300]
    except Exception as e:   # noqa: BLE001
        facts[f] = type(e).__name__
print(json.dumps({"host": facts, "threads_budget": threads}))
data, offs, _, meta = W.make_config(3, 1.0)
rv = V.resolve("cl100k_base", allow_stand_in=True)
ov = oracle.OracleVocab(rv.file_bytes, rv.max_ranks)
sub = offs[:n_sample + 1]
nb = int(sub[-1])
oracle.encode_batch([ov], [rv.pattern_id], data[:nb], sub, nthreads=1, want_ids=True)
k = 1
while True:
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        oracle.encode_batch([ov], [rv.pattern_id], data[:nb], sub, nthreads=k, want_ids=True)
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"threads": k, "MB_per_s": round(nb / best / 1e6, 1), "per_thread": round(nb / best / 1e6 / k, 2), "sample_bytes": nb}))
    sys.stdout.flush()
    if k >= (os.cpu_count() or 1):
        break
    k = min(2 * k, os.cpu_count() or 1)
