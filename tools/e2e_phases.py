"""Wall-clock phases of Engine.run at C3 (developer tool)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from onekapy_b200.engine import Engine, start_ring
from onekapy_b200.lattice import LatticeGeom, final_geometry

eng = Engine(0)
spec, par, _ = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "c3", 0, 0, 1)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); res = eng.run(spec, par); t1 = T()
    print("run total %.1f ms  attempts %.3g  work lattice %dx%d final %dx%d  rerun %d of %d" % (1e3*(t1-t0), res["stats"]["attempts"], res["work_geom"].nrows, res["work_geom"].ncols, res["geom"].nrows, res["geom"].ncols, res["stats"]["rerun_realizations"], len(par)))
R = len(par)
t0 = T(); start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths); t1 = T(); print("start_ring %.1f ms" % (1e3*(t1-t0)))
t0 = T(); dp = eng.upload(spec, par, start); t1 = T(); print("upload %.1f ms" % (1e3*(t1-t0)))
t0 = T(); eng.reset_stats(); sub = eng.upload(spec, par.slice(0, R, max(1, R//256)), start[::max(1, spec.npaths//128)]); eng.capture(spec, sub); bb = eng.read_stats()["bbox"]; t1 = T(); print("pilot %.1f ms" % (1e3*(t1-t0)), bb)
w, h = bb[1]-bb[0], bb[3]-bb[2]
geom = LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget).expanded(bb[0]-.5*w, bb[1]+.5*w, bb[2]-.5*h, bb[3]+.5*h)
t0 = T(); counts = eng.new_counts(geom); t1 = T(); print("new_counts %.1f ms" % (1e3*(t1-t0)))
t0 = T(); eng.reset_stats(); eng.capture(spec, dp, geom, counts); st = eng.read_stats(); t1 = T(); print("capture %.1f ms" % (1e3*(t1-t0)), st["bbox"], geom.strictly_contains(st["bbox"]))
sys.exit(0)
t0 = T(); final = final_geometry(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget, st["bbox"]); i0, j0 = geom.offset_of(final); out = counts[i0:i0+final.nrows, j0:j0+final.ncols].contiguous().cpu().numpy(); t1 = T(); print("crop+D2H %.1f ms" % (1e3*(t1-t0)))
