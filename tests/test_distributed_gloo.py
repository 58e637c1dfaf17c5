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


def _al_worker(rank, world, port, q):
    """Sharded active-learning loop on 2 ranks with a synthetic labeller: unique selections, the expected pool
    shrinkage and identical windows on both ranks (the replicated-pool path returned duplicates)."""
    import torch
    from vboc_b200 import al_loop
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, N, nx = 1, 3, 2
    full = np.random.default_rng(0).uniform(-1, 1, size=(101, nx))
    lo, hi = vd.shard_range(len(full), rank, world)
    pool = full[lo:hi]
    labelled = []

    def label_fn(X, xg):
        labelled.append(X.copy())
        return (np.abs(X[:, 1]) < 0.5).astype(np.int64), np.repeat(X[:, None, :], N + 1, axis=1)

    def query_fn(model, pool, B):
        etp = 1.0 - np.abs(np.abs(pool[:, 1]) - 0.5)
        idx, emax = vd.sharded_topk(etp, B)
        return idx, etp, emax

    noop = lambda m, X: None
    hist = []
    Xi, Xt, rest = al_loop.active_learning(n, pool, 20, 16, None, None, 0.0, 1.0, noop, noop, etp_stop=0.0, max_rounds=2,
                                           N=N, label_fn=label_fn, query_fn=query_fn, history=hist, sharded=True)
    q.put((rank, Xi, Xt, rest, np.concatenate(labelled), [h["etpmax"] for h in hist]))
    dist.destroy_process_group()


def test_sharded_active_learning_two_ranks():
    world, port = 2, 29547
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_al_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=120) for _ in ps], key=lambda r: r[0])
    for p in ps:
        p.join(timeout=60)
    (_, Xi0, Xt0, rest0, lab0, e0), (_, Xi1, Xt1, rest1, lab1, e1) = res
    assert np.array_equal(Xi0, Xi1) and np.array_equal(Xt0, Xt1) and e0 == e1   # identical windows on all ranks
    assert Xi0.shape == (20, 4)                              # 20 initial rows, 16 out / 16 in per round
    labelled = np.concatenate([lab0, lab1])
    assert len(labelled) == 20 + 2 * 16                       # B states per round in total, not per rank
    assert len(np.unique(labelled, axis=0)) == len(labelled)  # nothing labelled twice
    assert len(rest0) + len(rest1) == 101 - 20 - 32
    # the global selection: the labelled query states are the 32 most uncertain of the states left after the start
    full = np.random.default_rng(0).uniform(-1, 1, size=(101, 2))
    parts = [full[slice(*vd.shard_range(101, r, 2))] for r in range(2)]
    init = np.concatenate([parts[r][:10] for r in range(2)])
    left = np.concatenate([parts[r][10:] for r in range(2)])
    etp = 1.0 - np.abs(np.abs(left[:, 1]) - 0.5)
    want = left[np.argsort(-etp)[:32]]
    got = np.concatenate([lab0[10:], lab1[10:]])
    assert set(map(tuple, want)) == set(map(tuple, got))
