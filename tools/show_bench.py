"""Summarise a bench.py JSON line (file argument): headline, roofline, parity, breakdown, e2e legs, extra configs."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
ff = d['config']['farfield'] or {}
print("N=%d value %.4g att/s  %.2f ms/step  %.0f real/s  lattice %s  ff tiles %sx%s order %s eta %s near %.2f" % (
    d['n_gpus'], d['value'], d['ms_per_step'], d['realizations_per_s'], d['config']['lattice'], ff.get('ntx'), ff.get('nty'), ff.get('order'), ff.get('eta'), ff.get('mean_near', 0)))
r = d['roofline']
print("roofline: %.2f TF/s of %.2f probe = %.3f (nominal %.3f), executed est %.3f, kernel %.2f ms, share %s, flush %.3f ms" % (
    r['achieved'], r['peak'], r['frac'], r['frac_of_nominal'], r['frac_executed_estimate'], r['kernel_ms_per_launch'], r['kernel_share_of_step'], r['flush_kernel_ms_per_step']))
p = d.get('parity') or {}
print("parity:", {k: p.get(k) for k in ('endpoint_max_rel_err', 'step_counts_equal', 'attempts_equal', 'differing_cells', 'cells_nonzero')})
b = d['breakdown']
print("breakdown: step", [round(v, 2) for v in b['per_rank_step_ms']], "capture", [round(v, 2) for v in b['per_rank_capture_ms']], "kernel", [round(v, 2) for v in b['per_rank_track_kernel_ms']],
      "allreduce", [round(v, 3) for v in b['per_rank_allreduce_ms']], "skew %.3f" % b['skew_ms'], b['allreduce_via'])
print("grid_check:", d.get('grid_check'))
c = d.get('cpu_baseline')
if c: print("cpu: %.4g att/s on %d cores (%s)" % (c['value'], c['cores'], c['sample']))
for k in ('e2e', 'e2e_exact', 'e2e_dropin'):
    e = d.get(k)
    if e: print("%s: %.4g att/s  %.1f ms  %s" % (k, e['value'], e.get('ms_per_step', e.get('ms_per_call', 0)), {a: e[a] for a in ('affected_realizations', 'host_sampling_ms_per_call') if a in e}))
ra = d['raster']
if 'roofline' in ra: print("raster: %.1f G word-ops/s = %.3f of RED peak %.0f; segments/s %.4g" % (ra['roofline']['achieved'], ra['roofline']['frac'], ra['roofline']['peak'], ra['segments_per_s']))
print("launches", d['gpu_launches'], "clocks", d['clocks'])
for k, v in (d.get('configs') or {}).items():
    if 'error' in v:
        print("==", k, v); continue
    f = v['config']['farfield'] or {}
    print("== %-14s %.4g att/s %8.2f ms  %7.0f real/s  lattice %s frac %.3f  ff %sx%s o%s near %.2f" % (k, v['value'], v['ms_per_step'], v['realizations_per_s'], v['config']['lattice'], v['roofline']['frac'], f.get('ntx'), f.get('nty'), f.get('order'), f.get('mean_near', 0)))
    pp = v.get('parity') or {}
    if pp: print("     parity", {a: pp.get(a) for a in ('endpoint_max_rel_err', 'step_counts_equal', 'attempts_equal', 'differing_cells', 'cells_nonzero')})
    for e in ('e2e', 'e2e_exact'):
        if v.get(e): print("     %s %.4g att/s %.2f ms affected %s" % (e, v[e]['value'], v[e]['ms_per_step'], v[e].get('affected_realizations')))
    bb = v['breakdown']
    if len(bb['per_rank_step_ms']) > 1:
        print("     step", [round(x, 1) for x in bb['per_rank_step_ms']], "kernel", [round(x, 1) for x in bb['per_rank_track_kernel_ms']], "allreduce", [round(x, 2) for x in bb['per_rank_allreduce_ms']], "skew %.2f" % bb['skew_ms'])
    if 'roofline' in v['raster']: print("     raster %.1f G word-ops/s = %.3f of RED peak" % (v['raster']['roofline']['achieved'], v['raster']['roofline']['frac']))
    if v.get('grid_check'): print("     grid_check", v['grid_check'])
