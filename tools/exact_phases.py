"""Where Engine.run_exact spends its time (ONEKA_PHASES=1 synchronises at every phase boundary).  Run on the GPU box:
    ONEKA_PHASES=1 python tools/exact_phases.py [c3|c4|c1|c5] [R]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ONEKA_PHASES"] = "1"
import numpy as np
import bench
from onekapy_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "c3"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 0
spec, par, label = bench.make_workload(wl, R, 0, 1)
eng = Engine(0)
for k in range(3):
    t0 = time.perf_counter(); a = eng.run(spec, par, reuse_lattice=False); t1 = time.perf_counter()
    b = eng.run_exact(spec, par, reuse_lattice=False); t2 = time.perf_counter()
    print("%s R=%d: run %.1f ms, run_exact %.1f ms, affected %s flagged %s, phases %s" % (
        wl, len(par), 1e3 * (t1 - t0), 1e3 * (t2 - t1), b["stats"]["affected_realizations"], b["stats"]["rerun_realizations"],
        {k: round(v, 1) for k, v in b["stats"]["phases_ms"].items()}), flush=True)
print("cells set: run %d, run_exact %d" % (np.count_nonzero(a["counts"]), np.count_nonzero(b["counts"])))
