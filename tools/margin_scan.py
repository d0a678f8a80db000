"""How many realizations leave the estimated lattice, and what Engine.run costs, as a function of the pilot margin."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from onekapy_b200.engine import Engine
eng = Engine(0)
for wl in sys.argv[1:] or ["c3", "c5", "c4"]:
    spec, par, _ = bench.make_workload(wl, 0, 0, 1)
    for margin in (0.05, 0.15, 0.3, 0.5):
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            res = eng.run(spec, par, margin=margin)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        print("%s margin %.2f: %.1f ms, rerun %d of %d" % (wl, margin, 1e3 * (t1 - t0), res["stats"]["rerun_realizations"], len(par)), flush=True)
