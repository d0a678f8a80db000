"""Wall-clock phases of Engine.run (developer tool):  python tools/e2e_phases.py [c3|c4|c5] [R]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from onekapy_b200.engine import Engine, start_ring, RealizationParams
from onekapy_b200.lattice import LatticeGeom, final_geometry

eng = Engine(0)
name = sys.argv[1] if len(sys.argv) > 1 else "c3"
spec, par, _ = bench.make_workload(name, int(sys.argv[2]) if len(sys.argv) > 2 else 0, 0, 1)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
par = RealizationParams(q=pin(par.q), cond=pin(par.cond), poro=pin(par.poro), thick=pin(par.thick), coef=pin(par.coef))
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T(); res = eng.run(spec, par, reuse_lattice=False); t1 = T()
    print("run total %.1f ms  attempts %.3g  work lattice %dx%d final %dx%d  rerun %d of %d" % (1e3*(t1-t0), res["stats"]["attempts"], res["work_geom"].nrows, res["work_geom"].ncols, res["geom"].nrows, res["geom"].ncols, res["stats"]["rerun_realizations"], len(par)))
R = len(par)
t0 = T(); start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths); t1 = T(); print("start_ring %.2f ms" % (1e3*(t1-t0)))
t0 = T(); dp = eng.upload(spec, par, start); t1 = T(); print("upload %.2f ms" % (1e3*(t1-t0)))
t0 = T(); eng.reset_stats(); sub = eng.upload(spec, par.slice(0, R, max(1, R//256)), start[::max(1, spec.npaths//128)]); eng.capture(spec, sub); bb = eng.read_stats()["bbox"]; t1 = T(); print("pilot %.2f ms" % (1e3*(t1-t0)))
w, h = bb[1]-bb[0], bb[3]-bb[2]
geom = res["work_geom"]
t0 = T(); counts = eng.new_counts(geom); flags = torch.zeros(R, dtype=torch.int32, device="cuda"); t1 = T(); print("new_counts %.2f ms (%.1f MB)" % (1e3*(t1-t0), counts.numel()*4/1e6))
eng.set_profiling(True) if hasattr(eng, "set_profiling") else None
t0 = T(); eng.reset_stats(); eng.capture(spec, dp, geom, counts, flags=flags); st = eng.read_stats(); t1 = T(); print("guarded capture %.2f ms" % (1e3*(t1-t0)), getattr(eng, "kernel_ms", lambda: "")())
t0 = T(); nflag = int(flags.sum().item()); t1 = T(); print("flags.sum %.2f ms -> %d" % (1e3*(t1-t0), nflag))
final = final_geometry(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget, st["bbox"])
if nflag:
    from onekapy_b200.engine import _copy_overlap
    t0 = T(); counts2 = eng.new_counts(final); _copy_overlap(counts, geom, counts2, final); t1 = T(); print("carry over %.2f ms" % (1e3*(t1-t0)))
    t0 = T(); sel = dp.select(flags.nonzero().reshape(-1)); eng.capture(spec, sel, final, counts2); t1 = T(); print("rerun of %d flagged %.2f ms" % (nflag, 1e3*(t1-t0)))
    counts, geom = counts2, final
i0, j0 = geom.offset_of(final)
t0 = T(); out = counts[i0:i0+final.nrows, j0:j0+final.ncols].contiguous(); t1 = T(); print("crop %.2f ms" % (1e3*(t1-t0)))
t0 = T(); o = out.cpu().numpy(); t1 = T(); print("D2H %.2f ms (%.1f MB)" % (1e3*(t1-t0), o.nbytes/1e6))
