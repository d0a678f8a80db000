"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 \
        tools/multi_gpu_check.py

Every rank runs Engine.run on ITS shard of the same seeded realization rows with the NCCL group;
rank 0 also runs all rows alone and the two grids must be identical (integer sums are order-free).
Also compares against the oracle on the golden sto_basic fixture, sharded over the ranks.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import bench
    from onekapy_b200 import parallel
    from onekapy_b200.engine import Engine, FlowSpec, RealizationParams
    from helpers import scal

    rank, world, group = parallel.init_from_env()
    eng = Engine(int(os.environ.get("LOCAL_RANK", "0")))

    # (1) seeded perham rows, sharded vs whole
    spec, par, _ = bench.make_workload("c3", 512, 500, seed=77)         # same seed on every rank -> same rows
    r0, r1 = parallel.shard_range(len(par), rank, world)
    res = eng.run(spec, par.slice(r0, r1), group=group)
    if rank == 0:
        whole = eng.run(spec, par)
        assert res["geom"] == whole["geom"], (res["geom"], whole["geom"])
        assert res["total_weight"] == whole["total_weight"] == len(par)
        assert np.array_equal(res["counts"], whole["counts"])
        print("[multi-gpu] %d ranks: sharded grid == single-GPU grid (%d x %d, %d nonzero, max %d)"
              % (world, res["geom"].nrows, res["geom"].ncols, np.count_nonzero(res["counts"]), res["counts"].max()), flush=True)

    # (2) golden fixture (executed reference, auto-expanding), 6 realizations over the ranks
    g = np.load(os.path.join(ROOT, "tests", "golden", "sto_basic.npz"))
    s = scal(g)
    spec = FlowSpec(well_xy=g["wells_xyr"][:, :2].copy(), xtarget=s["xt"], ytarget=s["yt"], rtarget=s["rt"], npaths=s["P"],
                    duration=s["duration"], base=s["base"], spacing=s["spacing"], umbra=s["umbra"], confined=s["confined"],
                    tol=s["tol"], maxstep=s["maxstep"])
    r0, r1 = parallel.shard_range(len(g["k"]), rank, world)
    par = RealizationParams(q=g["q"][r0:r1], cond=g["k"][r0:r1], poro=g["n"][r0:r1], thick=g["H"][r0:r1], coef=g["coef"][r0:r1])
    res = eng.run(spec, par, group=group)
    if rank == 0:
        ref = g["auto_geom"]
        gm = res["geom"]
        assert [gm.xmin, gm.xmax, gm.ymin, gm.ymax, gm.nrows, gm.ncols] == list(ref[[0, 1, 2, 3, 6, 7]])
        want = g["auto_counts"].astype(np.uint32)
        assert res["total_weight"] == 6.0 and np.all(res["counts"] >= want)
        print("[multi-gpu] golden sto_basic over %d ranks: geometry identical, differing cells %d of %d"
              % (world, np.count_nonzero(res["counts"] != want), np.count_nonzero(want)), flush=True)
    # (3) exact emulation of the auto-expanding grid across ranks: rank k's paths come after rank k-1's
    res = eng.run_exact(spec, par, group=group)
    if rank == 0:
        gm = res["geom"]
        assert [gm.xmin, gm.xmax, gm.ymin, gm.ymax, gm.nrows, gm.ncols] == list(ref[[0, 1, 2, 3, 6, 7]])
        nd = np.count_nonzero(res["counts"] != want)
        print("[multi-gpu] run_exact over %d ranks: differing cells %d of %d" % (world, nd, np.count_nonzero(want)), flush=True)
        assert nd == 0 and res["total_weight"] == 6.0
    # (4) fewer realizations than ranks: empty shards must take part in every collective
    r0, r1 = parallel.shard_range(1, rank, world)
    one = RealizationParams(q=g["q"][r0:r1], cond=g["k"][r0:r1], poro=g["n"][r0:r1], thick=g["H"][r0:r1], coef=g["coef"][r0:r1])
    a = eng.run(spec, one, group=group)
    b = eng.run_exact(spec, one, group=group)
    if rank == 0:
        assert a["total_weight"] == b["total_weight"] == 1.0 and a["counts"].max() == 1 and a["geom"] == b["geom"]
        print("[multi-gpu] 1 realization over %d ranks (empty shards): ok, %d cells" % (world, np.count_nonzero(b["counts"])), flush=True)
    if group is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    if rank == 0:
        print("[multi-gpu] ok", flush=True)


if __name__ == "__main__":
    main()
