"""Multi-GPU plumbing of the hot path.

Every OCP is independent, so the problem index range is sharded over the ranks with NO traffic
during the solves (the single-host analogue is `Pool.map(data_generation, range(num_prob))`,
VBOC/triplependulum_vboc.py:399-405).  The only exchange steps are
  * one all-gather(v) of the result rows before the NN fit (the reference flattens `traj`, :404-405);
  * AL: a global top-B by entropy = local top-B per rank, all-gather of the candidates, final selection
    (replaces `np.argpartition`, AL/triplependulum_al.py:267-270).
Works with any torch.distributed backend (nccl over NVLink on the GPU box, gloo in the CPU tests).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` problems for `rank`; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dev():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def all_gather_rows(rows):
    """all-gather-v of a (n_i, d) float64 array: every rank gets the concatenation in rank order."""
    rows = np.ascontiguousarray(rows, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rows
    world, dev = dist.get_world_size(), _dev()
    d = rows.shape[1]
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([rows.shape[0]], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    m = max(counts)
    pad = torch.zeros((m, d), dtype=torch.float64, device=dev)
    pad[: rows.shape[0]] = torch.from_numpy(rows).to(dev)
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, counts)], axis=0)


def global_topk(values, k, index_offset=0):
    """Indices (global) and values of the k largest entries over all ranks' `values`."""
    values = np.asarray(values, dtype=np.float64)
    kk = min(k, values.shape[0])
    loc = np.argpartition(-values, kk - 1)[:kk] if kk > 0 else np.empty(0, dtype=np.int64)
    cand = np.stack([values[loc], (loc + index_offset).astype(np.float64)], axis=1) if kk else np.empty((0, 2))
    allc = all_gather_rows(cand)
    order = np.argsort(-allc[:, 0], kind="stable")[:k]
    return allc[order, 1].astype(np.int64), allc[order, 0]


def sharded_topk(values, k):
    """Global top-k over a pool that is SHARDED across the ranks (every rank holds different rows).  Returns the
    LOCAL indices (into this rank's `values`, largest index first, as AL/triplependulum_al.py:267-270 sorts them)
    of the entries of the global top-k that live on this rank, and the global maximum of the selected values.
    The ranks' selections are disjoint and their sizes add up to min(k, total pool size), so each rank labels and
    removes its own part.  Without an initialised process group this is the plain single-process selection."""
    values = np.asarray(values, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        kk = min(k, values.shape[0])
        loc = np.argpartition(-values, kk - 1)[:kk] if kk > 0 else np.empty(0, dtype=np.int64)
        return np.sort(loc)[::-1].astype(np.int64), (float(values[loc].max()) if kk else 0.0)
    world, rank, dev = dist.get_world_size(), dist.get_rank(), _dev()
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([values.shape[0]], dtype=torch.int64, device=dev))
    counts = [int(c.item()) for c in counts]
    lo = sum(counts[:rank])
    gidx, gval = global_topk(values, k, index_offset=lo)
    mine = gidx[(gidx >= lo) & (gidx < lo + counts[rank])] - lo
    return np.sort(mine)[::-1].astype(np.int64), (float(gval.max()) if gval.size else 0.0)
