"""Multi-GPU sharding of the parse path: screenshots are independent (ref:util/utils.py:417-496 is a pure function of
its inputs), so rank r of R takes screenshots i = r (mod R), weights are replicated, and the only exchange is ONE
gather of fixed-size padded result records to rank 0 per batch.  NCCL over NVLink on the GPUs, gloo in
the CPU tests.  The payload (~10 KB per screenshot) is latency-bound; no fused compute+collective kernel is warranted.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

REC_BOXES, REC_TOK = 300, 128


def shard_indices(n: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n, world))


def record_width(max_new_tokens: int) -> int:
    return 2 + REC_BOXES * 4 + REC_TOK * (max_new_tokens + 1)


def pack_records(results: Sequence[Tuple[list, torch.Tensor]], max_new_tokens: int) -> torch.Tensor:
    """[(filtered_boxes_elem, caption ids)] -> float32 [B, record_width]: n_elem, n_cap, bboxes[300x4], ids[128x(T+1)]
    (ids padded with the pad token 1; token ids < 2^24 are exact in float32)."""
    T1 = max_new_tokens + 1
    host = np.zeros((len(results), record_width(max_new_tokens)), np.float32)
    for b, (elems, ids) in enumerate(results):
        nb, nt = min(len(elems), REC_BOXES), min(int(ids.shape[0]), REC_TOK)
        host[b, 0], host[b, 1] = nb, nt
        if nb:
            host[b, 2:2 + nb * 4] = np.asarray([e["bbox"] for e in elems[:nb]], np.float32).ravel()
        if nt:
            tk = np.ones((nt, T1), np.float32)
            tk[:, :ids.shape[1]] = ids[:nt].numpy()
            host[b, 2 + REC_BOXES * 4:2 + REC_BOXES * 4 + tk.size] = tk.ravel()
    return torch.from_numpy(host)


def unpack_records(rec: torch.Tensor, max_new_tokens: int):
    T1 = max_new_tokens + 1
    out = []
    r = rec.cpu().numpy()
    for row in r:
        nb, nt = int(row[0]), int(row[1])
        boxes = row[2:2 + nb * 4].reshape(nb, 4)
        ids = row[2 + REC_BOXES * 4:2 + REC_BOXES * 4 + nt * T1].reshape(nt, T1).astype(np.int64)
        out.append((boxes, ids))
    return out


def gather_records(rec: torch.Tensor, rank: int, world: int, bufs=None):
    """One collective per batch.  Returns the list of per-rank record tensors on rank 0, None elsewhere."""
    import torch.distributed as dist
    if world == 1:
        return [rec]
    if rank == 0 and bufs is None:
        bufs = [torch.zeros_like(rec) for _ in range(world)]
    dist.gather(rec, bufs if rank == 0 else None, dst=0)
    return bufs if rank == 0 else None


def interleave(per_rank: Sequence[Sequence], n: int, world: int) -> list:
    """Inverse of shard_indices: per-rank result lists back to global screenshot order."""
    out = [None] * n
    for r, items in enumerate(per_rank):
        for k, i in enumerate(shard_indices(n, r, world)):
            out[i] = items[k]
    return out
