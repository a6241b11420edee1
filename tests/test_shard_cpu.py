"""CPU (gloo, world_size 2): the N>1 plumbing of the parse path -- sharding, record packing, the single gather."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from omniparser_b200 import shard


def _fake_results(rank, n, T):
    rng = np.random.default_rng(100 + rank)
    out = []
    for k in range(n):
        ne, nc = int(rng.integers(1, 80)), int(rng.integers(0, 60))
        elems = [{"bbox": [float(np.float32(v)) for v in rng.uniform(0, 1, 4)]} for _ in range(ne)]
        ids = torch.from_numpy(rng.integers(0, 51289, size=(nc, T + 1)))
        out.append((elems, ids))
    return out


def _worker(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_global = 7
    mine = shard.shard_indices(n_global, rank, world)
    res = _fake_results(rank, len(mine), T)
    rec = shard.pack_records(res, T)
    # ragged shards: pad to the largest shard so the gather has one fixed shape
    width = (n_global + world - 1) // world
    pad = torch.zeros((width, rec.shape[1]))
    pad[:rec.shape[0]] = rec
    got = shard.gather_records(pad, rank, world)
    if rank == 0:
        per_rank = [shard.unpack_records(g[:len(shard.shard_indices(n_global, r, world))], T) for r, g in enumerate(got)]
        q.put(shard.interleave(per_rank, n_global, world))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_roundtrip():
    T, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert len(merged) == 7 and all(m is not None for m in merged)
    # rank r's k-th result must land at global index r + k*world, bit-exact
    for r in range(world):
        mine = shard.shard_indices(7, r, world)
        ref = _fake_results(r, len(mine), T)
        for k, i in enumerate(mine):
            boxes, ids = merged[i]
            assert np.array_equal(boxes, np.asarray([e["bbox"] for e in ref[k][0]], np.float32))
            assert np.array_equal(ids, ref[k][1].numpy())


def _pipe_worker(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    steps, B = 5, 3
    pipe = shard.GatherPipe(rank, world, "cpu", B, T)
    for s in range(steps):
        pipe.submit(_fake_results(10 * rank + s, B, T))       # returns at once; the worker thread runs the collective
    pipe.drain()
    if rank == 0:
        q.put([[shard.unpack_records(g, T) for g in step] for step in pipe.received])
    pipe.close()
    dist.barrier()
    dist.destroy_process_group()


def test_async_gather_pipe_two_ranks():
    """the bench's gather path: submit() never blocks on the peer, drain() completes every step's gather, in order"""
    T, world = 8, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 500)
    procs = [ctx.Process(target=_pipe_worker, args=(r, world, port, T, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert len(got) == 5
    for s, step in enumerate(got):
        for r in range(world):
            ref = _fake_results(10 * r + s, 3, T)
            for (boxes, ids), (elems, rids) in zip(step[r], ref):
                assert np.array_equal(boxes, np.asarray([e["bbox"] for e in elems], np.float32))
                assert np.array_equal(ids, rids.numpy())


def test_dense_screenshot_fits_the_record_and_overflow_raises():
    """300 icons (= max_det, all captioned, the reference captions every box in chunks of 128) + 200 OCR boxes round-trip;
    anything beyond the capacity raises instead of being clamped."""
    import pytest
    T = 8
    rng = np.random.default_rng(1)
    elems = [{"bbox": [float(np.float32(v)) for v in rng.uniform(0, 1, 4)]} for _ in range(500)]
    ids = torch.from_numpy(rng.integers(0, 51289, size=(300, T + 1)))
    rec = shard.pack_records([(elems, ids)], T)
    (boxes, rids), = shard.unpack_records(rec, T)
    assert boxes.shape == (500, 4) and np.array_equal(rids, ids.numpy())
    with pytest.raises(ValueError, match="capacity"):
        shard.pack_records([(elems, torch.zeros((301, T + 1), dtype=torch.long))], T)
    with pytest.raises(ValueError, match="capacity"):
        shard.pack_records([(elems * 3, ids)], T)


def test_shard_indices_cover():
    for n in (0, 1, 7, 64):
        for w in (1, 2, 8):
            allidx = sorted(i for r in range(w) for i in shard.shard_indices(n, r, w))
            assert allidx == list(range(n))
