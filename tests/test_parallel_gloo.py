"""N>1 host logic on CPU: world_size-2 gloo process group.  Realizations are sharded by index,
each rank produces its shard's count grid (here with the ORACLE as the stand-in for the kernels,
which is allowed in tests), and the grids are summed with the same allreduce_counts / reduce_bbox
/ sum_int calls Engine.run makes under NCCL.  The reduced grid must equal the single-process grid."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    from onekapy_b200 import parallel
    from onekapy_b200.lattice import LatticeGeom, final_geometry
    from oracle import oracle as O
    from helpers import scal

    r, w, group = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world) and group is not None
    g = np.load(os.path.join(ROOT, "tests", "golden", "sto_basic.npz"))
    s = scal(g)
    R = len(g["k"])
    r0, r1 = parallel.shard_range(R, rank, world)
    start = O.start_ring(s["xt"], s["yt"], s["rt"], s["P"])
    # every rank must work on the same lattice: bbox of the local shard -> min/max allreduce
    pf = O.Field(s["spacing"], s["spacing"], s["xt"], s["yt"])
    res = O.capture(pf, 0, g["wells_xyr"][:, :2], s["base"], s["xt"], s["yt"], s["confined"], g["q"][r0:r1],
                    g["k"][r0:r1], g["n"][r0:r1], g["H"][r0:r1], g["coef"][r0:r1], start, s["duration"], s["umbra"],
                    s["tol"], s["maxstep"])
    local_bbox = (pf.xmin + pf.deltax, pf.xmax - pf.deltax, pf.ymin + pf.deltay, pf.ymax - pf.deltay) if r1 > r0 \
        else (np.inf, -np.inf, np.inf, -np.inf)
    bbox = parallel.reduce_bbox(local_bbox, group)
    # the packed form Engine.run / run_exact use: ONE all-gather of (bbox, R, flag count) per agreement
    rows = parallel.gather_rows(list(local_bbox) + [r1 - r0, rank], group)
    assert rows.shape == (world, 6) and rows[:, 5].tolist() == list(range(world)) and int(rows[:, 4].sum()) == R
    assert parallel.union_bbox(rows[:, :4]) == bbox
    assert parallel.union_bbox(rows[:0, :4]) == (np.inf, -np.inf, np.inf, -np.inf)
    geom = LatticeGeom.anchored(s["spacing"], s["spacing"], s["xt"], s["yt"]).expanded(*g["lattice"])
    assert geom.strictly_contains(bbox)
    pf = O.Field(s["spacing"], s["spacing"], s["xt"], s["yt"])
    pf.expand(*g["lattice"])
    O.capture(pf, 1, g["wells_xyr"][:, :2], s["base"], s["xt"], s["yt"], s["confined"], g["q"][r0:r1], g["k"][r0:r1],
              g["n"][r0:r1], g["H"][r0:r1], g["coef"][r0:r1], start, s["duration"], s["umbra"], s["tol"], s["maxstep"])
    counts = torch.from_numpy(pf.pgrid.astype(np.int32))
    parallel.allreduce_counts(counts, group)
    total = parallel.sum_int(r1 - r0, group)
    assert parallel.any_rank(rank == 1, group) is True and parallel.any_rank(False, group) is False
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), counts=counts.numpy(), total=total, bbox=np.array(bbox),
             shard=np.array([r0, r1]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from onekapy_b200.parallel import shard_range
    for R in (0, 1, 5, 8, 1000, 10007):
        for world in (1, 2, 3, 8):
            parts = [shard_range(R, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == R
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_identity_without_group():
    from onekapy_b200 import parallel
    assert parallel.reduce_bbox((1.0, 2.0, 3.0, 4.0)) == (1.0, 2.0, 3.0, 4.0)
    assert parallel.sum_int(7) == 7 and parallel.any_rank(True) is True
    assert parallel.gather_rows([1.0, 2.0]).tolist() == [[1.0, 2.0]]
    assert parallel.union_bbox([[0, 1, 2, 3], [np.nan, 0, 0, 0], [-1, 0.5, 2.5, 9]]) == (-1.0, 1.0, 2.0, 9.0)
    assert parallel.init_from_env() == (0, 1, None) or os.environ.get("WORLD_SIZE", "1") != "1"


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce(tmp_path, golden):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    g = golden("sto_basic.npz")
    assert np.array_equal(a["counts"], b["counts"])
    assert np.array_equal(a["counts"], g["fixed_counts"].astype(np.int32))       # == single-process reference grid
    assert int(a["total"]) == int(b["total"]) == len(g["k"])
    assert list(a["shard"]) == [0, 3] and list(b["shard"]) == [3, 6]
    v = g["verts"]
    # the reduced bbox brackets every vertex (it is the outermost interior lattice line, not the exact min/max)
    assert a["bbox"][0] <= v[:, 0].min() + 10 and a["bbox"][1] >= v[:, 0].max() - 10
    assert np.array_equal(a["bbox"], b["bbox"])
