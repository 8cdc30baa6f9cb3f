#!/usr/bin/env python3
"""Multi-GPU smoke (run under torchrun, one process per GPU): rank 0 parses the rank file, the packed tables travel by
NCCL broadcast, every rank encodes its byte-balanced shard of ONE batch, per-shard token totals and per-prompt counts
are all_gathered; rank 0 checks the reassembled result against the oracle."""
import os; os.environ.setdefault("CFBPE_ALLOW_STAND_IN", "1")   # measurement aids run on the stand-in vocabularies
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cyberfabric-core_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist
from cfbpe import dist as D, plugin as P, workload as W

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
plug = D.load_vocab_everywhere(lambda blobs: P.GpuBpeTokenizerPlugin(local, ("cl100k_base",), 64 << 20, 1 << 16, import_blobs=blobs),
                               ["cl100k_base"], 0, dev)
data, offs, vid, meta = W.make_config(3, 0.125)          # same batch on every rank (seeded)
sh_bytes, sh_offs, _, (lo, hi) = D.shard_batch(data, offs, None, rank, world)
r = plug.encode_batch(P.SecurityContext.anonymous(), P.EncodeBatchRequest(P.VocabRef("cl100k_base"), np.ascontiguousarray(sh_bytes), sh_offs))
totals = D.gather_totals(len(r.ids), dev)
counts = D.gather_counts(r.counts, dev)
ok = True
if rank == 0:
    from oracle import oracle
    rv = plug.resolved["cl100k_base"]
    ov = oracle.OracleVocab(rv.file_bytes, rv.max_ranks)
    want_ids, want_off, want_counts = oracle.encode_batch([ov], [rv.pattern_id], data, offs, nthreads=os.cpu_count())
    ok = np.array_equal(counts, want_counts) and int(totals.sum()) == len(want_ids) and np.array_equal(r.ids, want_ids[:len(r.ids)])
    print("nccl smoke: world=%d shards=%s totals=%s ok=%s" % (world, [(lo, hi)], totals.tolist(), ok))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
