"""One process per GPU: batch sharding, vocab-table broadcast at init, per-shard count gather.

The prompt batch shards trivially (prompts are independent), so the data path has NO collective
(SURVEY.md section 8(e)).  torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is
used for exactly two things:
  * init:      rank `src` parses the rank file and builds the packed tables once; the blob
               (cfbpe_vocab_export) is broadcast and installed on every other rank
               (cfbpe_vocab_import) -- only one rank parses;
  * per batch: all_gather of each shard's token total (8 bytes per rank), and optionally of the
               per-prompt counts when a dense global out_offsets is wanted.
Token-id streams stay on the GPU that produced them.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def shard_by_bytes(offsets: np.ndarray, world: int) -> List[Tuple[int, int]]:
    """Contiguous prompt ranges [lo, hi) per rank, balanced by BYTES: rank r starts at the first prompt
    whose start offset is >= r * total / world (boundaries snapped to prompt starts)."""
    n = len(offsets) - 1
    total = int(offsets[n])
    cuts = [0]
    for r in range(1, world):
        target = (total * r) // world
        i = int(np.searchsorted(offsets[:n + 1], target, side="left"))
        cuts.append(min(max(i, cuts[-1]), n))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_batch(data: np.ndarray, offsets: np.ndarray, vocab_ids: Optional[np.ndarray], rank: int, world: int):
    """this rank's (bytes, offsets, vocab_ids, (lo, hi)) view of a packed batch"""
    lo, hi = shard_by_bytes(offsets, world)[rank]
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    offs = (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)
    vid = None if vocab_ids is None else np.ascontiguousarray(vocab_ids[lo:hi])
    return data[b0:b1], offs, vid, (lo, hi)


def broadcast_blob(blob: Optional[np.ndarray], src: int, device=None) -> np.ndarray:
    """Broadcast a packed table blob from rank `src` (size first, then payload) over the default process
    group.  `device`: torch device the backend needs the tensor on (cuda for NCCL, None/cpu for gloo)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    dev = torch.device("cpu") if device is None else device
    size = torch.tensor([0 if blob is None else blob.size], dtype=torch.int64, device=dev)
    dist.broadcast(size, src=src)
    n = int(size.item())
    if rank == src:
        t = torch.from_numpy(np.ascontiguousarray(blob)).to(dev)
    else:
        t = torch.empty(n, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()


def load_vocab_everywhere(plugin_factory, names, src: int = 0, device=None):
    """Build the plugin on every rank with only rank `src` parsing rank files: src loads normally and
    exports; the others import the broadcast blobs.  `plugin_factory(import_blobs)` -> plugin."""
    import torch.distributed as dist
    rank = dist.get_rank()
    if rank == src:
        plug = plugin_factory(None)
        for nm in names:
            broadcast_blob(plug.export_vocab(nm), src, device)
        return plug
    blobs = {nm: broadcast_blob(None, src, device) for nm in names}
    return plugin_factory(blobs)


def gather_totals(local_total, device=None) -> np.ndarray:
    """all_gather of one int64 per rank (token totals of the shards)"""
    import torch
    import torch.distributed as dist
    dev = torch.device("cpu") if device is None else device
    t = local_total if isinstance(local_total, torch.Tensor) else torch.tensor([int(local_total)], dtype=torch.int64, device=dev)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out).cpu().numpy()


def gather_counts(local_counts: np.ndarray, device=None, sizes=None) -> np.ndarray:
    """all_gather of the per-prompt counts of every shard (variable length), in rank order: a dense global
    counts vector from which out_offsets follows by one exclusive scan.  `sizes`: the shards' prompt counts when every rank
    already knows them (they all cut the batch with shard_by_bytes) -- saves the first collective and its synchronisation."""
    import torch
    import torch.distributed as dist
    dev = torch.device("cpu") if device is None else device
    world = dist.get_world_size()
    if sizes is None:
        sizes = gather_totals(len(local_counts), device)
    m = int(max(sizes)) if world else 0
    pad = torch.zeros(max(m, 1), dtype=torch.int32, device=dev)
    pad[:len(local_counts)] = torch.from_numpy(local_counts.view(np.int32) if local_counts.dtype == np.uint32 else local_counts.astype(np.int32)).to(dev, non_blocking=True)
    out = torch.empty(world * max(m, 1), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(out, pad)
    host = out.cpu().numpy().view(np.uint32).reshape(world, max(m, 1))
    return np.concatenate([host[r, :int(s)] for r, s in enumerate(sizes)])
