"""Multi-GPU parity ON HARDWARE (pytest -m gpu; skipped on a box with fewer than 2 GPUs).

Two ranks, one process per GPU (torch.multiprocessing.spawn, NCCL over 127.0.0.1), realizations sharded by
parallel.shard_range, the count grids summed by THE collective of the path -- oneka_allreduce_counts through the
library's own communicator (Engine.init_comm) -- and, for comparison, through torch.distributed:

    sharded Engine.run grid        == single-GPU grid == fixed/auto grids of the executed reference (tests/golden)
    sharded Engine.run_exact grid  == single-GPU grid == the executed reference's auto-expanding grid, cell for cell
                                      (two-pass and one-pass schemes; rank 1's paths come after rank 0's)

The reference's grid is additive over realizations (oneka/probabilityfield.py:357-358), so integer sums make the result
independent of the number of GPUs.  The CPU-side logic of the same code runs under gloo in tests/test_parallel_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CAPTURES = ["sto_basic.npz", "sto_perham.npz", "unc_basic.npz"]


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, own_comm):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    from helpers import geom
    from test_gpu_parity import spec_of
    from onekapy_b200 import parallel
    from onekapy_b200.engine import Engine
    import bench
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    group = dist.group.WORLD
    eng = Engine(rank)
    if own_comm:
        eng.init_comm(group)
        assert eng._comm == (world, rank)
    try:
        # ---- the collective itself: every rank contributes rank + 1 to every cell ----
        t = torch.full((257, 63), rank + 1, dtype=torch.int32, device=eng.device)
        eng.allreduce_counts(t, group)
        eng.synchronize()
        assert int(t.min()) == int(t.max()) == world * (world + 1) // 2

        # ---- executed-reference fixtures ----
        for name in CAPTURES:
            g = np.load(os.path.join(HERE, "golden", name))
            s, spec, par = spec_of(g)
            r0, r1 = parallel.shard_range(len(par), rank, world)
            ref = geom(g, "auto_")
            want_auto = g["auto_counts"].astype(np.uint32)
            res = eng.run(spec, par.slice(r0, r1), group=group, reuse_lattice=False)
            gm = res["geom"]
            assert (gm.xmin, gm.xmax, gm.ymin, gm.ymax, gm.nrows, gm.ncols) == (ref["xmin"], ref["xmax"], ref["ymin"], ref["ymax"], ref["nrows"], ref["ncols"])
            assert res["total_weight"] == len(par)
            if rank == 0:
                whole = eng.run(spec, par, reuse_lattice=False)
                assert np.array_equal(whole["counts"], res["counts"]), name
                assert np.all(res["counts"] >= want_auto)
            for kw in (dict(two_pass_below=10**9), dict(two_pass_below=0, reuse_lattice=False)):
                ex = eng.run_exact(spec, par.slice(r0, r1), group=group, **kw)
                nd = int(np.count_nonzero(ex["counts"] != want_auto))
                if rank == 0:
                    print("[%d GPUs%s] %s run_exact(%s): differing cells vs the executed reference %d of %d"
                          % (world, ", own communicator" if own_comm else "", name, kw, nd, np.count_nonzero(want_auto)), flush=True)
                assert nd == 0 and ex["total_weight"] == len(par), (name, kw, nd)

        # ---- a larger sharded run (perham, 96 realizations x 64 paths): N-GPU grid == 1-GPU grid, both flows ----
        spec, par, _ = bench.make_workload("c3", 96, 64, 11)
        r0, r1 = parallel.shard_range(len(par), rank, world)
        a = eng.run(spec, par.slice(r0, r1), group=group, reuse_lattice=False)
        b = eng.run_exact(spec, par.slice(r0, r1), group=group, reuse_lattice=False)
        assert a["total_weight"] == b["total_weight"] == 96
        if rank == 0:
            a1 = eng.run(spec, par, reuse_lattice=False)
            b1 = eng.run_exact(spec, par, reuse_lattice=False)
            assert a1["geom"] == a["geom"] and np.array_equal(a1["counts"], a["counts"])
            assert b1["geom"] == b["geom"] and np.array_equal(b1["counts"], b["counts"])
            assert np.all(a["counts"] >= b["counts"]) and 0 < b["stats"]["affected_realizations"] < 96
            print("[%d GPUs] perham 96 x 64: run == 1-GPU run, run_exact == 1-GPU run_exact (%d cells set, %d fewer with the exact clip; "
                  "%d affected realizations on rank 0)" % (world, np.count_nonzero(a["counts"]),
                                                           int((a["counts"].astype(np.int64) - b["counts"]).sum()), b["stats"]["affected_realizations"]), flush=True)
        dist.barrier()
    finally:
        eng.close()
        dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("own_comm", [True, False], ids=["oneka_allreduce_counts", "torch_distributed"])
def test_two_gpus_equal_one_gpu_equal_reference(own_comm):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), own_comm), nprocs=2, join=True)
