"""world_size-2 gloo test of the multi-GPU host logic: index sharding, the all-gather(v) of result
rows and the global top-B selection (vboc_b200/distributed.py)."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from vboc_b200 import distributed as vd


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 11
    lo, hi = vd.shard_range(total, rank, world)
    rows = np.arange(lo, hi, dtype=np.float64)[:, None] * np.ones((1, 6))
    allr = vd.all_gather_rows(rows)
    vals = np.arange(lo, hi, dtype=np.float64) % 4 + 0.01 * np.arange(lo, hi)
    idx, top = vd.global_topk(vals, 3, index_offset=lo)
    q.put((rank, lo, hi, allr, idx, top))
    dist.destroy_process_group()


def test_shard_gather_topk():
    world, port = 2, 29541
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    ref_vals = np.arange(11, dtype=np.float64) % 4 + 0.01 * np.arange(11)
    ref_idx = np.argsort(-ref_vals)[:3]
    cover = []
    for rank, lo, hi, allr, idx, top in res:
        cover += list(range(lo, hi))
        assert allr.shape == (11, 6) and np.array_equal(allr[:, 0], np.arange(11))
        assert sorted(idx.tolist()) == sorted(ref_idx.tolist())
        assert np.allclose(np.sort(top), np.sort(ref_vals[ref_idx]))
    assert sorted(cover) == list(range(11))


def test_shard_range_partitions():
    for total in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            parts = [vd.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
