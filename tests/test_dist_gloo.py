"""The N>1 host logic on CPU: world_size 2 over gloo (127.0.0.1).  Rank 0 "parses" a table blob, both
ranks end up with identical bytes (what cfbpe_vocab_import would install), shards are disjoint and the
gathered per-shard counts reassemble the global count vector."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, os.path.join(ROOT, "cyberfabric-core_b200"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "simt"))
    import torch.distributed as dist
    from cfbpe import dist as D
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # init: only rank 0 builds the packed tables (host builder via the emulator harness), then broadcast
        blob = None
        if rank == 0:
            import simlib
            data = open(os.path.join(ROOT, "vocabs", "tekken_240911.tiktoken"), "rb").read()
            lines = b"\n".join(data.splitlines()[:2000])
            sv = simlib.SimVocab(lines, 0, 0, 0)
            blob = np.arange(sv.table_bytes % 100000 + 1000, dtype=np.uint32).view(np.uint8).copy()   # stand-in payload of realistic size
        got = D.broadcast_blob(blob, 0)
        np.save(os.path.join(tmpdir, "blob_%d.npy" % rank), got)
        # per batch: shard by bytes, "encode" = one token per 3 bytes, gather totals and per-prompt counts
        rng = np.random.default_rng(7)
        lens = rng.integers(0, 300, size=501)
        offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        payload = rng.integers(0, 255, size=int(offs[-1]), dtype=np.uint8)
        sh_bytes, sh_offs, _, (lo, hi) = D.shard_batch(payload, offs, None, rank, world)
        counts = (np.diff(sh_offs.astype(np.int64)) // 3).astype(np.uint32)
        totals = D.gather_totals(int(counts.sum()))
        allc = D.gather_counts(counts)
        np.save(os.path.join(tmpdir, "totals_%d.npy" % rank), totals)
        np.save(os.path.join(tmpdir, "counts_%d.npy" % rank), allc)
        np.save(os.path.join(tmpdir, "range_%d.npy" % rank), np.array([lo, hi]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    b0, b1 = np.load(tmp_path / "blob_0.npy"), np.load(tmp_path / "blob_1.npy")
    assert b0.size > 1000 and np.array_equal(b0, b1)
    r0, r1 = np.load(tmp_path / "range_0.npy"), np.load(tmp_path / "range_1.npy")
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 501
    rng = np.random.default_rng(7)
    lens = rng.integers(0, 300, size=501)
    want = (lens // 3).astype(np.uint32)
    for r in (0, 1):
        assert np.array_equal(np.load(tmp_path / ("counts_%d.npy" % r)), want)
        t = np.load(tmp_path / ("totals_%d.npy" % r))
        assert t.tolist() == [int(want[:r0[1]].sum()), int(want[r0[1]:].sum())]
