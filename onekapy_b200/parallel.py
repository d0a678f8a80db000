"""Multi-GPU plumbing: one process per GPU, realizations sharded, ONE data-path collective.

Realizations are i.i.d. given their parameter rows and the only shared state of the
reference's loop is the additive grid and total_weight (oneka/probabilityfield.py:357-358),
so the path shards by realization index with no exchange until the end:

  * shard_range        contiguous slice of [0, R) per rank
  * allreduce_counts   the single allreduce(sum) of the integer count grid (torch.distributed form: gloo on CPU in the tests;
                       on GPUs Engine.allreduce_counts issues the same sum through the C ABI, oneka_allreduce_counts)
  * gather_rows        one all-gather of a short packed vector (bounding box, counts) between kernel phases
  * reduce_bbox        4 doubles min/max, so that every rank works on the same lattice

Counts are integers, so the reduced grid is order-independent and bit-reproducible for any
number of GPUs.  The functions take a torch.distributed process group; with group=None they
are the identity (single GPU).  They work on CPU tensors with the gloo backend too, which is
how tests/test_parallel_gloo.py covers the N>1 logic without GPUs.
"""
import numpy as np


def shard_range(R, rank, world):
    """Contiguous, balanced: the first R % world ranks get one extra realization."""
    base, extra = divmod(int(R), int(world))
    r0 = rank * base + min(rank, extra)
    return r0, r0 + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


def reduce_bbox(bbox, group=None, device=None):
    """(min x, max x, min y, max y) over all ranks.  Empty shards contribute +-inf."""
    if group is None:
        return tuple(float(v) for v in bbox)
    import torch
    dist = _dist()
    # one MIN allreduce on (x0, -x1, y0, -y1)
    t = torch.tensor([bbox[0], -bbox[1], bbox[2], -bbox[3]], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    v = t.cpu().numpy()
    return (float(v[0]), float(-v[1]), float(v[2]), float(-v[3]))


def any_rank(flag, group=None, device=None):
    if group is None:
        return bool(flag)
    import torch
    dist = _dist()
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return bool(int(t.item()))


def sum_int(value, group=None, device=None):
    if group is None:
        return int(value)
    import torch
    dist = _dist()
    t = torch.tensor([int(value)], dtype=torch.int64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def gather_rows(values, group=None, device=None):
    """All ranks' copies of a short float64 vector -> ndarray [world, n] (ONE collective, one host read).

    Engine.run packs everything the ranks have to agree on between two kernel phases (bounding box, realization count,
    number of flagged realizations) into one such vector instead of issuing a blocking scalar collective per item."""
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    if group is None:
        return v[None, :].copy()
    import torch
    dist = _dist()
    t = torch.as_tensor(v, dtype=torch.float64).to(device or "cpu")
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return torch.stack(out).cpu().numpy()


def union_bbox(rows):
    """(min x, max x, min y, max y) over the finite rows of [n, 4]; all-infinite when there is none."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 4)
    ok = rows[np.isfinite(rows).all(axis=1)]
    if len(ok) == 0:
        return (np.inf, -np.inf, np.inf, -np.inf)
    return (float(ok[:, 0].min()), float(ok[:, 1].max()), float(ok[:, 2].min()), float(ok[:, 3].max()))


def allreduce_counts(counts, group=None):
    """In-place sum of the per-rank count grids (torch int32 tensor holding uint32 bit patterns:
    two's-complement addition is the same bits, and counts never exceed the realization count)."""
    if group is None:
        return counts
    dist = _dist()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def init_from_env(backend=None):
    """Join the default process group from torchrun's environment (RANK/WORLD_SIZE/MASTER_*).
    Returns (rank, world, group or None)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world == 1:
        return 0, 1, None
    import torch
    dist = _dist()
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, **kw)
    return rank, world, dist.group.WORLD
