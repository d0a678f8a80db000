#!/usr/bin/env python
"""bench.py -- throughput of the capture-zone hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c1|c5|c4full] [--impl ours|reference] [--legs ...]

A "step" is one pass of the hot path (track + rasterise + register, then -- for N > 1 -- the
allreduce of the count grid, oneka_allreduce_counts) over one batch of pre-sampled realization rows:

  c3 (default)  data/perham.py field case, 10 000 realizations x 1000 paths PER GPU per step
                (BASELINE.json configs[2], the largest single-GPU configuration)
  c4            synthetic 200-well field, --realizations per GPU per step (default 2048) x 1000 paths
  c4full        BASELINE.json configs[3] itself: 1 000 000 realizations x 1000 paths of the 200-well field, sharded over
                the N GPUs (125 000 per GPU at N = 8), one allreduce
  c1            data/basic.py, 100 x 100 (the reference's CPU-runnable case)
  c5            data/basic.py field on a fine lattice (spacing 4, umbra 20 -> ~15x15-node windows, a
                4096 x 4096-class grid), 2048 realizations x 1000 paths: rasterisation stress

metric = particle-steps/s = DOPRI5 attempts per second summed over all particles and GPUs;
realizations/s is reported beside it.  The headline leg scales WEAKLY (per-GPU work fixed as N grows).

The ONE JSON line carries, besides the contract's keys for the headline leg:
  parity        the oracle (CPU restatement, pinned to the executed reference) on a seeded subsample of the TIMED rows, same
                lattice: endpoint max relative error, step counts equal, differing cells (BASELINE.md section 3 promises these
                with every number); repeated inside every extra leg;
  configs       short extra legs, outside the headline's timed region: c4 (far field on) and c4_direct (off), c5, c1,
                c3_unconfined, c4_strong (a FIXED total of 125 000 realizations sharded over the N GPUs: strong scaling) and, at
                N = 8, c4full -- each with value, e2e, roofline, parity, breakdown;
  breakdown     per rank: capture ms, track-kernel ms, flush ms, allreduce ms (includes waiting for the slowest rank) and
                the skew between ranks, so that the limiter of a scaling curve can be named from the record;
  grid_check    (N > 1) the allreduced grid == the sum of the per-rank grids gathered on rank 0, and rank 0 -- alone, on
                its own GPU -- recomputes other ranks' shards from their seeds and finds the same grids;
  e2e           Engine.run() with HOST buffers every step (H2D from pinned memory, pilot, guarded capture, allreduce, crop,
                D2H); e2e_exact = Engine.run_exact (the drop-in default: the reference's order-dependent clip reproduced);
                e2e_dropin = oneka.stochastic.create_stochastic_capturezone(...) itself, wall clock incl. host sampling + fit;
  raster        the rasteriser against ITS roofline: bit-set word operations per second over the measured RED.OR / shared
                atomicOr peaks (oneka_red_probe), cell tests per second.

Timing: W >= 3 warm-up steps, then K steps between two barrier + synchronize brackets, timed on
the device with CUDA events on the launching stream, max over ranks.  L2 is flushed (256 MiB
write) before every timed step; the flush is inside the bracket (40 us against >100 ms steps).
`roofline` is the fused tracking+raster kernel against the FP64 pipe: algorithmic flops per
attempt = 257 + 90*Nw over the kernel's CUDA-event time, divided by an FP64
DFMA probe measured in the same process (MEASURED_PEAKS.json carries no FP64 figure); the fraction of the
nominal peak (148 SM x 64 lanes x 2 x max clock) is printed beside it.
`cpu_baseline` / `--impl reference`: the C restatement of the reference (oracle/, OpenMP over
realizations, all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C4FULL_TOTAL = 1000000         # BASELINE.json configs[3]
C4_STRONG_TOTAL = 125000       # one GPU's share of it: the fixed total of the strong-scaling leg


# ------------------------------------------------------------------------------------------------
def make_workload(name, realizations, npaths, seed, unconfined=False):
    from onekapy_b200 import problems, synthetic
    from onekapy_b200.engine import FlowSpec
    if name == "c3":
        pb = problems.load("perham")
        R = realizations or 10000
        P = npaths or 1000
        label = "C3 perham field case (29 wells, 102 obs), %d realizations x %d paths per GPU per step" % (R, P)
    elif name in ("c4", "c4full"):
        pb = synthetic.well_field(200)
        R = realizations or 2048
        P = npaths or 1000
        label = "C4 synthetic 200-well field (seed 2020), %d realizations x %d paths per GPU per step" % (R, P)
    elif name == "c1":
        pb = problems.load("basic")
        R = realizations or 100
        P = npaths or 100
        label = "C1 basic (2 wells), %d realizations x %d paths per GPU per step" % (R, P)
    elif name == "c5":
        pb = problems.load("basic")
        pb["spacing"], pb["umbra"] = 4.0, 20.0          # ~13 x 17 km of capture zones -> a 4096 x 4096-class lattice
        R = realizations or 2048
        P = npaths or 1000
        label = "C5 basic field on a fine lattice (spacing 4, umbra 20: ~15x15-node windows, 4096^2-class grid), %d realizations x %d paths per GPU per step" % (R, P)
    else:
        raise SystemExit("unknown workload %r" % name)
    if make_workload.umbra:
        pb["umbra"] = float(make_workload.umbra)
        label += ", umbra=%g (scan)" % pb["umbra"]
    if make_workload.spacing:
        pb["spacing"] = float(make_workload.spacing)
        label += ", spacing=%g (scan)" % pb["spacing"]
    if unconfined:
        pb["confined"] = False
        label += ", confined=False"
    params = synthetic.sample_rows_fast(pb, R, seed)
    make_workload.problem = pb
    xt, yt, rt = pb["wells"][pb["target"]][0:3]
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in pb["wells"]], dtype=float), xtarget=float(xt),
                    ytarget=float(yt), rtarget=float(rt), npaths=P, duration=float(pb["duration"]), base=float(pb["base"]),
                    spacing=float(pb["spacing"]), umbra=float(pb["umbra"]), confined=bool(pb["confined"]),
                    tol=float(pb["tol"]), maxstep=float(pb["maxstep"]))
    return spec, params, label


make_workload.umbra = make_workload.spacing = None      # --umbra / --spacing: lattice scans (tools/ab_run.sh), never the default line


def host_sampling_rate(pb, R):
    """Host share of the drop-in call (steps 1-4 of oneka/stochastic.py:186-199: variates, fit, A..F draw) for R
    realizations of this problem, one core: the rate the GPU has to be fed at.  Not part of any timed region."""
    from onekapy_b200.host.stochastic import sample_realizations
    from onekapy_b200.host.utilities import filter_obs
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    xt, yt = pb["wells"][pb["target"]][0:2]
    state = np.random.get_state()
    t0 = time.perf_counter()
    sample_realizations(R, pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"], obs, xt, yt,
                        rng=np.random.default_rng(0), fit_method="qr", log_rows=False)
    dt = time.perf_counter() - t0
    np.random.set_state(state)
    return {"realizations_per_s": R / dt, "cores": 1, "seconds": dt,
            "what": "host.stochastic.sample_realizations(fit_method='qr'): reference-order variates, shared-QR fit, A..F draw"}


def flops_per_attempt(nw):
    """6 evaluations x (20 + 15 Nw) + 137 of Runge-Kutta algebra."""
    return 257 + 90 * nw


class ClockSampler(threading.Thread):
    """nvidia-smi style clock / throttle-reason samples during the timed region (pynvml)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (test infrastructure; executed here only as the checker / the CPU baseline)
def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_run(spec, params, geom, nthreads=0, want_paths=False):
    """The CPU restatement on a fixed lattice; returns (seconds, result dict, field)."""
    from oracle import oracle as O
    from onekapy_b200.engine import start_ring
    pf = O.Field(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget)
    pf.expand(geom.xmin + 0.5 * geom.deltax, geom.xmax - 0.5 * geom.deltax, geom.ymin + 0.5 * geom.deltay,
              geom.ymax - 0.5 * geom.deltay)
    assert (pf.nrows, pf.ncols, pf.xmin, pf.ymin) == (geom.nrows, geom.ncols, geom.xmin, geom.ymin)
    start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths)
    t0 = time.perf_counter()
    res = O.capture(pf, 1, spec.well_xy, spec.base, spec.xtarget, spec.ytarget, spec.confined, params.q, params.cond,
                    params.poro, params.thick, params.coef, start, spec.duration, spec.umbra, spec.tol, spec.maxstep,
                    nthreads=nthreads, want_paths=want_paths)
    return time.perf_counter() - t0, res, pf


def cpu_geom(spec, params):
    """A lattice for the CPU arm without a GPU: bounding box from the oracle's own auto-expanding pass
    over a few realizations, generously padded."""
    from oracle import oracle as O
    from onekapy_b200.engine import start_ring
    from onekapy_b200.lattice import LatticeGeom
    sub = params.slice(0, min(len(params), 2))
    pf = O.Field(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget)
    start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, min(spec.npaths, 64))
    O.capture(pf, 0, spec.well_xy, spec.base, spec.xtarget, spec.ytarget, spec.confined, sub.q, sub.cond, sub.poro,
              sub.thick, sub.coef, start, spec.duration, spec.umbra, spec.tol, spec.maxstep, want_paths=False)
    w, h = pf.xmax - pf.xmin, pf.ymax - pf.ymin
    return LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget).expanded(
        pf.xmin - w, pf.xmax + w, pf.ymin - h, pf.ymax + h)


def cpu_sample(spec, params, geom, target_s=12.0):
    """Bounded CPU sample: calibrate on one realization per thread, then size the sample for ~target_s."""
    cores = _cores()
    n0 = min(len(params), cores)
    t, res, _ = oracle_run(spec, params.slice(0, n0), geom, nthreads=cores)
    per = t / max(1, n0) * cores                       # seconds of one thread per realization
    n = int(min(len(params), max(n0, cores * max(1, int(target_s / max(per, 1e-9))))))
    if n > n0:
        t, res, _ = oracle_run(spec, params.slice(0, n), geom, nthreads=cores)
    else:
        n = n0
    return dict(seconds=t, realizations=n, attempts=res["attempts"], steps=res["steps"], cores=cores)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (C restatement, OpenMP, all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    wl = "c4" if args.workload == "c4full" else args.workload
    spec, params, label = make_workload(wl, args.realizations, args.npaths, args.seed, args.unconfined)
    geom = cpu_geom(spec, params)
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the
    # reference arm on one core at N > 1
    cores = _cores()
    # calibrate the per-step sample so that warmup + steps finish in ~2 minutes
    t1, _, _ = oracle_run(spec, params.slice(0, min(len(params), cores)), geom, nthreads=cores)
    per_real = t1 / min(len(params), cores)
    budget = 100.0 / max(1, args.steps + args.warmup)
    n = int(min(len(params), max(cores, int(budget / max(per_real, 1e-9)))))
    sub = params.slice(0, n)
    for _ in range(args.warmup):
        oracle_run(spec, sub, geom, nthreads=cores)
    tot_t, tot_a = 0.0, 0
    for _ in range(args.steps):
        t, res, _ = oracle_run(spec, sub, geom, nthreads=cores)
        tot_t += t
        tot_a += res["attempts"]
    value = tot_a / tot_t
    rps = n * args.steps / tot_t
    sample = "%d of %d realizations x %d paths per step, fixed lattice %dx%d" % (n, len(params), spec.npaths, geom.nrows, geom.ncols)
    line = {"impl": "reference", "metric": "particle-steps/s", "value": value, "unit": "DOPRI5 attempts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "realizations_per_s": rps,
            "config": {"workload": label, "wells": int(len(spec.well_xy))},
            "cpu_baseline": {"value": value, "unit": "DOPRI5 attempts/s", "cores": cores, "kind": "port", "sample": sample,
                             "realizations_per_s": rps},
            "e2e": {"value": value, "unit": "DOPRI5 attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
class Ctx:
    """What every leg needs: the engine of this rank, the process group, a 256 MiB L2-flush buffer."""

    def __init__(self, eng, rank, world, group, torch, dist):
        self.eng, self.rank, self.world, self.group, self.torch, self.dist = eng, rank, world, group, torch, dist
        self.dev = eng.device
        self.flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.probe_tf = None
        self.red = None
        self.max_mhz = None

    def barrier(self):
        if self.group is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def gather(self, values):
        """[world, n] float64 of a short per-rank vector (one all-gather)."""
        from onekapy_b200 import parallel
        return parallel.gather_rows(values, self.group, self.dev if self.group is not None else None)


def pilot_lattice(cx, spec, params, dp, start, margin=0.25, full=False):
    """The fixed lattice of a leg: bounding box of a tracking-only pass (all rows, or a strided pilot for the big legs)
    min/max-reduced over the ranks, grown by `margin`.  -> (geom, ff_box)"""
    from onekapy_b200 import parallel
    from onekapy_b200.lattice import LatticeGeom
    eng = cx.eng
    R = len(params)
    eng.reset_stats()
    if full or R <= 256:
        try:                                            # (large well fields: tile grid from a strided pilot, so that this full pass is fast too)
            eng._farfield_from_pilot(spec, params, dp)
        except Exception as exc:                        # set-up convenience only: without it the pass below runs on direct sums
            print("bench: far-field pilot skipped (%r)" % (exc,), file=sys.stderr)
        eng.reset_stats()
        eng.capture(spec, dp)
    else:
        rstep, pstep = max(1, R // 256), max(1, spec.npaths // 128)
        eng.capture(spec, eng.upload(spec, params.slice(0, R, rstep), start[::pstep]))
    bbox = parallel.union_bbox(cx.gather(eng.read_stats()["bbox"]))
    w, h = bbox[1] - bbox[0], bbox[3] - bbox[2]
    m = 0.0 if full else margin
    base = LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget)
    geom = base.expanded(bbox[0] - m * w, bbox[1] + m * w, bbox[2] - m * h, bbox[3] + m * h)
    ff_box = None if full else (bbox[0] - 0.1 * w, bbox[1] + 0.1 * w, bbox[2] - 0.1 * h, bbox[3] + 0.1 * h)
    return geom, ff_box


def parity_block(cx, spec, params, geom, n, seed=5):
    """the oracle on a seeded subsample of the timed rows, same lattice (rank 0's rows)."""
    from onekapy_b200.engine import RealizationParams
    eng = cx.eng
    R = len(params)
    rows = np.sort(np.random.default_rng(seed).choice(R, size=min(n, R), replace=False))
    sub = RealizationParams(q=params.q[rows], cond=params.cond[rows], poro=params.poro[rows], thick=params.thick[rows],
                            coef=params.coef[rows])
    dp = eng.upload(spec, sub)
    counts = eng.new_counts(geom)
    eng.reset_stats()
    pp = eng.capture(spec, dp, geom, counts, per_path=True)
    st = eng.read_stats()
    got = counts.cpu().numpy().view(np.uint32)
    t, res, pf = oracle_run(spec, sub, geom, nthreads=_cores(), want_paths=True)
    want = pf.pgrid.astype(np.uint32)
    end = pp["end_xy"].cpu().numpy()
    scale = np.maximum(np.abs(res["end_xy"]).max(axis=2), 1.0)
    rel = float((np.abs(end - res["end_xy"]).max(axis=2) / scale).max())
    ndiff = int(np.count_nonzero(got != want))
    nz = int(np.count_nonzero(want))
    return {"rows": [int(r) for r in rows], "paths": int(len(rows) * spec.npaths), "endpoint_max_rel_err": rel, "endpoint_tolerance": 1e-6,
            "step_counts_equal": bool(np.array_equal(pp["nverts"].cpu().numpy(), res["nverts"])),
            "attempts_equal": bool(st["attempts"] == res["attempts"]), "attempts": int(st["attempts"]),
            "differing_cells": ndiff, "cells_nonzero": nz, "differing_cell_fraction": ndiff / max(1, nz),
            "oracle_seconds": t, "against": "oracle/ (C restatement pinned bit for bit to the executed reference, tests/golden) on the same rows and lattice"}


def postprocess_block(cx, counts, total_weight, sigma=2.0, reps=3):
    """SURVEY N3 on the leg's own count grid, device-resident, CUDA events: the exceedance histogram behind create_impact_plot
    (oneka/visualize.py:382-386 sorts the whole grid) and the separable FP64 Gaussian of create_probability_plot (:228-233)."""
    import ctypes as C
    from onekapy_b200 import _cabi
    from onekapy_b200.host.postprocess import gaussian_taps
    torch, eng = cx.torch, cx.eng
    nrows, ncols = int(counts.shape[0]), int(counts.shape[1])
    ncell = nrows * ncols
    nbins = int(total_weight) + 1
    hist = torch.zeros(max(nbins, 2), dtype=torch.int64, device=cx.dev)
    tmp = torch.empty((nrows, ncols), dtype=torch.float64, device=cx.dev)
    out = torch.empty_like(tmp)
    w, lw = gaussian_taps(sigma)
    s = torch.cuda.current_stream(cx.dev)

    def timed(fn):
        fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(s)
        for _ in range(reps):
            fn()
        b.record(s)
        b.synchronize()
        return a.elapsed_time(b) / reps
    h_ms = timed(lambda: _cabi.check(eng._L.oneka_count_histogram(eng._h, counts.data_ptr(), ncell, max(nbins, 2), hist.data_ptr())))
    g_ms = timed(lambda: _cabi.check(eng._L.oneka_gaussian_smooth(eng._h, counts.data_ptr(), nrows, ncols, float(total_weight), w.ctypes.data,
                                                                   int(lw), tmp.data_ptr(), out.data_ptr())))
    captured = int(ncell - int(hist[0].item()))
    return {"cells": ncell, "count_histogram_ms": h_ms, "count_histogram_GBps": 4.0 * ncell / (h_ms * 1e-3) / 1e9,
            "gaussian_smooth_ms": g_ms, "gaussian_smooth_GBps": 28.0 * ncell / (g_ms * 1e-3) / 1e9, "sigma_nodes": sigma, "taps": 2 * lw + 1,
            "cells_captured_at_least_once": captured,
            "note": "algorithmic bytes: histogram 4 B per cell; smooth 4 B read + 8 B written (rows), 8 B read + 8 B written (columns) = 28 B per cell"}


def run_leg(cx, name, wl, R, P, steps, warmup, seed, farfield="auto", unconfined=False, full_pilot=False, e2e_steps=0,
            parity_n=0, sampler=None, scaling="weak", total=None, exact_e2e=False, note=None, postprocess=False):
    """One timed leg: K steps on rows resident in HBM (device-timed, max over ranks) + optional e2e / parity."""
    torch, eng = cx.torch, cx.eng
    from onekapy_b200.engine import RealizationParams, start_ring
    eng.farfield = farfield
    spec, params, label = make_workload(wl, R, P, seed + cx.rank, unconfined)
    problem = make_workload.problem
    R, P, nw = len(params), spec.npaths, len(spec.well_xy)
    start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, P)
    dp = eng.upload(spec, params, start)
    geom, ff_box = pilot_lattice(cx, spec, params, dp, start, full=full_pilot)
    counts = eng.new_counts(geom)

    def step(ev=None):
        cx.flush_buf.fill_(1)                                # L2 flush (126 MB L2)
        s = torch.cuda.current_stream(cx.dev)
        if ev:
            ev[0].record(s)
        eng.capture(spec, dp, geom, counts, ff_box=ff_box)
        if ev:
            ev[1].record(s)
        if cx.group is not None:
            eng.allreduce_counts(counts, cx.group)
        if ev:
            ev[2].record(s)

    # warm-up; a strided pilot may have missed an outlier: the lattice must hold every vertex before anything is timed
    for attempt in range(3):
        eng.reset_stats()
        for _ in range(max(1 if attempt else 3, warmup if not attempt else 1)):
            step()
        st = eng.read_stats()
        from onekapy_b200 import parallel
        bb = parallel.union_bbox(cx.gather(list(st["bbox"])))
        if geom.strictly_contains(bb):                       # (windows reaching an umbra beyond the outermost vertices are clipped by
            break                                            #  the lattice edge exactly as insert() clips them, probabilityfield.py:298-301)
        w, h = bb[1] - bb[0], bb[3] - bb[2]
        geom = geom.expanded(bb[0] - 0.1 * w, bb[1] + 0.1 * w, bb[2] - 0.1 * h, bb[3] + 0.1 * h)
        counts = eng.new_counts(geom)
    else:
        raise SystemExit("bench: lattice still clips after two enlargements")
    cx.barrier()

    # ---- value: K steps, inputs resident in HBM, device-timed ----
    eng.reset_stats()
    eng.set_profiling(True)
    eng.kernel_ms(reset=True)
    launches0 = eng.launch_count()
    if sampler is not None:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.barrier()
    ev0.record(torch.cuda.current_stream(cx.dev))
    for k in range(steps):
        step(evs[k])
    ev1.record(torch.cuda.current_stream(cx.dev))
    cx.barrier()
    if sampler is not None:
        sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    ff_info = eng.farfield_info()           # tiled far-field expansion of the well sum, or None = direct sums
    launches = eng.launch_count() - launches0
    stats = eng.read_stats()
    kms = eng.kernel_ms(reset=True)
    eng.set_profiling(False)
    cap_ms = sum(e[0].elapsed_time(e[1]) for e in evs)
    ar_ms = sum(e[1].elapsed_time(e[2]) for e in evs)
    per_rank = cx.gather([ms, cap_ms, ar_ms, kms["track_ms"], kms["flush_ms"], stats["attempts"], stats["steps"], stats["n_not_ok"],
                          R * steps, launches])
    ms_max = float(per_rank[:, 0].max())
    attempts, steps_acc, n_not_ok, reals = (float(per_rank[:, k].sum()) for k in (5, 6, 7, 8))
    value = attempts / (ms_max * 1e-3)
    rps = reals / (ms_max * 1e-3)
    breakdown = {"per_rank_step_ms": [float(v) / steps for v in per_rank[:, 0]],
                 "per_rank_capture_ms": [float(v) / steps for v in per_rank[:, 1]],
                 "per_rank_track_kernel_ms": [float(v) / steps for v in per_rank[:, 3]],
                 "per_rank_flush_kernel_ms": [float(v) / steps for v in per_rank[:, 4]],
                 "per_rank_allreduce_ms": [float(v) / steps for v in per_rank[:, 2]],
                 "per_rank_attempts_per_step": [float(v) / steps for v in per_rank[:, 5]],
                 "skew_ms": float(per_rank[:, 1].max() - per_rank[:, 1].min()) / steps,
                 "allreduce_bytes": int(counts.numel() * 4) if cx.group is not None else 0,
                 "allreduce_via": ("oneka_allreduce_counts (library-owned NCCL communicator)" if eng._comm else "torch.distributed") if cx.group is not None else None,
                 "note": "per step, CUDA events on the launching stream; allreduce_ms includes waiting for the slowest rank, so "
                         "min over ranks ~ the collective itself and (max - min) ~ the skew of the capture times"}

    # ---- rasteriser throughput against ITS roofline: one more, untimed, step into a fresh grid ----
    fresh = eng.new_counts(geom)
    eng.reset_stats()
    eng.capture(spec, dp, geom, fresh, ff_box=ff_box)
    rst = eng.read_stats()
    ragg = cx.gather([float(fresh.sum(dtype=torch.int64).item()), float(rst["steps"]), float(rst["exact_tests"])]).sum(axis=0)
    cells_step, segs_step, exact_step = (float(v) for v in ragg)
    post = None
    if postprocess:                                          # N3 on this rank's grid of one step (counts <= R)
        try:
            post = postprocess_block(cx, fresh, R)
        except Exception as exc:
            post = {"error": repr(exc)}
    del fresh
    step_s = ms_max * 1e-3 / steps
    # window of insert() (probabilityfield.py:298-301) for a segment of the mean accepted length: rows x columns tested by the
    # reference; the kernel issues one RED.OR (two when the row's bits straddle a word) per row that meets the capsule
    mean_len = 0.9 * spec.maxstep                       # accepted steps sit at the space cap almost everywhere (controller, capturezone.py:247)
    rows_seg = (mean_len * 0.64 + 2 * spec.umbra) / spec.spacing + 1          # E|dy| = 2/pi x length for an isotropic direction
    # bit-sets per window row: the plain flavour's spans straddle a 32-bit word in a quarter of the rows; the heavy flavour's
    # 64-bit bit-sets straddle a pair in an eighth, and it skips the rows behind a chained segment's start (0.36 umbra / spacing
    # of them for an isotropic direction).  Checked against ncu's RED counts at C5: 18.0 (plain) and 14.1 (heavy) per segment.
    flavour = eng.raster_flavour(spec.umbra, spec.spacing, eng.farfield_info() is not None)
    per_row = 1.25 if flavour == "plain" else 1.125 * (1.0 - 0.36 * (spec.umbra / spec.spacing) / rows_seg)
    raster = {"segments_per_s": segs_step / step_s, "cells_registered_per_s": cells_step / step_s,
              "cells_registered_per_segment": cells_step / max(1.0, segs_step),
              "exact_fp64_retests_per_segment": exact_step / max(1.0, segs_step),
              "window_rows_per_segment_estimate": rows_seg, "cell_tests_per_s_estimate": segs_step / step_s * rows_seg * rows_seg,
              "bitset_word_ops_per_s_estimate": segs_step / step_s * rows_seg * per_row, "flavour": flavour, "bitsets_per_window_row": per_row,
              "note": "cells registered = bits set in the per-realization bitmaps = sum of the count grid; word ops = one RED.OR per window row "
                      "(plain flavour: x1.25 for rows straddling a word; heavy: 64-bit, x1.125, minus the skipped rows behind a chained segment's start)"}
    if cx.red:
        peak = cx.red["l2_sector_per_lane"] * cx.world
        raster["roofline"] = {"bound": "atomic (bit-set RED.OR to L2, one 32-byte sector per lane: every lane rasterises another particle)",
                              "achieved": raster["bitset_word_ops_per_s_estimate"] / 1e9, "peak": peak, "unit": "1e9 word ops/s (all GPUs)",
                              "frac": raster["bitset_word_ops_per_s_estimate"] / 1e9 / peak,
                              "frac_of_coalesced_peak": raster["bitset_word_ops_per_s_estimate"] / 1e9 / (cx.red["l2_lane_private"] * cx.world),
                              "peaks": cx.red,
                              "reading": "peak = oneka_red_probe mode 4 (the rasteriser's own pattern: a sector per lane, a bitmap row further per operation); "
                                         "the coalesced lane-private figure (8 lanes per sector) is not reachable by particles that are cells apart. "
                                         "ncu at C5: plain flavour 151 G sector-REDs/s, L2 throughput 64 %, issue slots 65 %, long-scoreboard stalls on top "
                                         "(atomic-bound); heavy flavour 22 % fewer bit-sets, issue slots 77 %, L2 57 % (profiles/r02_track_kernel_c5_*_raw.csv)"}

    # ---- roofline of the fused tracking + raster kernel (this rank) ----
    if cx.probe_tf is None:
        cx.probe_tf, _ = eng.fp64_probe(1 << 16)
    probe_tf = cx.probe_tf
    track_ms = kms["track_ms"] / max(1, kms["track_launches"])
    att_per_launch = stats["attempts"] / max(1, kms["track_launches"])
    achieved = att_per_launch * flops_per_attempt(nw) / (track_ms * 1e-3) / 1e12
    nominal = 148 * 64 * 2 * (cx.max_mhz or 1965) * 1e6 / 1e12
    # DRAM traffic of the fused kernel from the ncu --set full capture in profiles/ (bytes per 1000 realizations of the
    # perham field, dram__bytes_read.sum + dram__bytes_write.sum), scaled to this launch: the path is not HBM-bound
    traffic = (137.0e6 if ff_info else 27.9e6) * (R / 1000.0) if wl == "c3" and not unconfined else None
    # with the far-field compression the kernel EXECUTES fewer flops than the reference's formulation needs: per evaluation
    # 20 (regional) + 16 per near well + 8 per polynomial term + ~16 of tile lookup, FMA = 2 (estimate from the mean near count)
    if ff_info:
        exec_flops = 6 * (20 + 16 * (ff_info["mean_near"] + 1.0) + 8 * ff_info["order"] + 16) + 137
    else:
        exec_flops = flops_per_attempt(nw)
    roofline = {"bound": "fp64", "kernel": "track_kernel<%s, raster%s>" % ("unconfined" if unconfined else "confined", ", far field" if ff_info else ""),
                "achieved": achieved, "peak": probe_tf,
                "unit": "TFLOP/s", "frac": achieved / probe_tf, "frac_of_nominal": achieved / nominal, "peak_nominal": nominal, "traffic": traffic,
                "traffic_note": "bytes per launch scaled from %s (ncu --set full at R=1000: dram__bytes_read.sum + dram__bytes_write.sum); algorithmic input is (Nw+9)*8 B per realization; "
                                "the rest is the realization's far-field coefficient table (97 KB, written once by the coefficient GEMM, read once by its four CTAs) and the registration bitmaps; "
                                "HBM is idle (< 0.2 %% of peak), the bound is the issue port / FP64 pipe" % ("profiles/r02_track_kernel_c3_raw.csv" if ff_info else "profiles/r01_track_kernel_raw.csv"),
                "peak_source": "in-run DFMA probe (oneka_fp64_probe); MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 lanes x 2 x max clock = %.1f" % nominal,
                "flops_per_attempt": flops_per_attempt(nw), "attempts_per_launch": att_per_launch, "kernel_ms_per_launch": track_ms,
                "flops_executed_per_attempt_estimate": exec_flops, "frac_executed_estimate": achieved * exec_flops / flops_per_attempt(nw) / probe_tf,
                "frac_note": "achieved/frac count the ALGORITHMIC flops of the reference's formulation (257 + 90 Nw per attempt, SURVEY 8d); "
                             "with the far-field compression active the kernel executes fewer (frac_executed_estimate), so frac may exceed 1",
                "flush_kernel_ms_per_step": kms["flush_ms"] / steps,
                "kernel_share_of_step": kms["track_ms"] / ms if cx.world == 1 else None}

    out = {"leg": name, "value": value, "unit": "DOPRI5 attempts/s", "ms_per_step": ms_max / steps, "steps": steps,
           "realizations_per_s": rps, "accepted_steps_per_s": steps_acc / (ms_max * 1e-3), "scaling": scaling,
           "config": {"workload": label if total is None else label.replace("per GPU per step", "on this rank, a %d-realization total sharded over %d GPU(s)" % (total, cx.world)),
                      "wells": nw, "lattice": [geom.nrows, geom.ncols], "paths_not_ok": n_not_ok, "farfield": ff_info,
                      "l2": "flushed before every step (256 MiB write)"},
           "roofline": roofline, "raster": raster, "breakdown": breakdown, "gpu_launches": int(per_rank[:, 9].sum())}
    if note:
        out["note"] = note
    if post is not None:
        out["postprocess"] = post

    # ---- e2e: the public calls with host buffers, every step ----
    if e2e_steps > 0:
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        hp = RealizationParams(q=pin(params.q), cond=pin(params.cond), poro=pin(params.poro), thick=pin(params.thick), coef=pin(params.coef))
        h2d = int(params.q.nbytes + params.cond.nbytes + params.poro.nbytes + params.thick.nbytes + params.coef.nbytes
                  + spec.well_xy.nbytes + start.nbytes)

        def timed(fn, api):
            for _ in range(2):                               # warm: the work lattice differs from the resident-input leg's,
                res = fn()                                   # so the bitmap workspace is re-allocated on the first call
            cx.barrier()
            t0 = time.perf_counter()
            att = 0
            for _ in range(e2e_steps):
                res = fn()
                att += res["stats"]["attempts"]
            cx.barrier()
            g = cx.gather([time.perf_counter() - t0, att])
            secs, tot = float(g[:, 0].max()), float(g[:, 1].sum())
            return {"value": tot / secs, "unit": "DOPRI5 attempts/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(res["counts"].nbytes + 10 * 8), "realizations_per_s": cx.world * R * e2e_steps / secs,
                    "ms_per_step": 1e3 * secs / e2e_steps, "api": api, "steps": e2e_steps,
                    "affected_realizations": res["stats"].get("affected_realizations")}

        out["e2e"] = timed(lambda: eng.run(spec, hp, group=cx.group, reuse_lattice=False),
                           "Engine.run(spec, params_host, reuse_lattice=False) = H2D + pilot pass + guarded capture (+ allreduce) + crop + D2H, all inside every timed call")
        if exact_e2e:
            out["e2e_exact"] = timed(lambda: eng.run_exact(spec, hp, group=cx.group, reuse_lattice=False),
                                     "Engine.run_exact(spec, params_host, reuse_lattice=False): the drop-in default -- the reference's order-dependent clip reproduced "
                                     "cell for cell: H2D + pilot + ONE fused pass with per-path boxes + fix-up of the affected realizations (+ allreduce) + D2H")

    # ---- parity gate on a seeded subsample of the timed rows (rank 0) ----
    if parity_n > 0 and cx.rank == 0:
        out["parity"] = parity_block(cx, spec, params, geom, parity_n)
    cx.barrier()
    out["_state"] = (spec, params, geom, problem, dp, ff_box, start)
    return out


def grid_check(cx, leg_state, wl, R, P, seed, unconfined, recompute):
    """N > 1: (i) the allreduced grid == the sum of the per-rank grids (gathered on rank 0); (ii) rank 0, ALONE on its own GPU,
    recomputes the shards of the ranks in `recompute` from their seeds and finds the grids those ranks reported."""
    torch, eng, dist = cx.torch, cx.eng, cx.dist
    spec, params, geom, _, dp, ff_box, start = leg_state
    local = eng.new_counts(geom)
    eng.capture(spec, dp, geom, local, ff_box=ff_box)
    summed = local.clone()
    eng.allreduce_counts(summed, cx.group)
    parts = [torch.empty_like(local) for _ in range(cx.world)] if cx.rank == 0 else None
    dist.gather(local, parts, dst=0, group=cx.group)
    out = None
    if cx.rank == 0:
        total = torch.zeros_like(local, dtype=torch.int64)
        for p in parts:
            total += p
        ok_sum = bool(torch.equal(total, summed.to(torch.int64)))
        ok_shards, redone = True, []
        for k in recompute:
            sk, pk, _ = make_workload(wl, R, P, seed + k, unconfined)
            mine = eng.new_counts(geom)
            eng.capture(sk, eng.upload(sk, pk, start), geom, mine, ff_box=ff_box)
            same = bool(torch.equal(mine, parts[k]))
            ok_shards &= same
            redone.append(int(k))
        out = {"allreduced_equals_sum_of_rank_grids": ok_sum, "ranks_recomputed_on_rank0": redone, "recomputed_shards_equal": ok_shards,
               "cells_nonzero": int((summed != 0).sum().item()), "sum_of_counts": int(summed.sum(dtype=torch.int64).item()),
               "ok": ok_sum and ok_shards}
    cx.barrier()
    return out


def dropin_leg(cx, R, P, steps):
    """The REAL drop-in: oneka.stochastic.create_stochastic_capturezone with the reference's positional signature
    (oneka/stochastic.py:76-81) on the perham problem: host sampling + fit + Engine.run_exact + D2H + ProbabilityField, wall clock."""
    from onekapy_b200 import problems
    from onekapy_b200.host.utilities import filter_obs
    from oneka.stochastic import create_stochastic_capturezone
    pb = problems.load("perham")
    obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
    eng = cx.eng
    eng.farfield = "auto"

    def call():
        return create_stochastic_capturezone(pb["target"], P, pb["duration"], R, pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"],
                                             pb["wells"], obs, pb["spacing"], pb["umbra"], pb["confined"], pb["tol"], pb["maxstep"],
                                             rng=np.random.default_rng(1), engine=eng)
    np.random.seed(12345)
    # the objects this process has piled up by now (problem tables, parity traces, JSON pieces) go to the permanent generation: a full
    # collection that rescans them in the middle of a 146 ms call showed up as one 150-210 ms call among five (tools/dropin_calls.py
    # in a fresh process: 145.5-147 ms every call, collector on or off)
    import gc
    gc.collect()
    gc.freeze()
    for _ in range(2):
        call()
    cx.torch.cuda.synchronize(cx.dev)
    t0 = time.perf_counter()
    att = 0
    each = []
    for _ in range(steps):
        tc = time.perf_counter()
        pf = call()
        each.append(round(1e3 * (time.perf_counter() - tc), 2))
        att += eng.last_stats["attempts"]
    secs = time.perf_counter() - t0
    gc.unfreeze()
    # where a call's time goes: the three pieces of create_stochastic_capturezone timed one by one on the same problem
    from onekapy_b200.host.stochastic import sample_realizations
    from onekapy_b200.host.probabilityfield import ProbabilityField
    from onekapy_b200.engine import FlowSpec
    xt, yt, rt = pb["wells"][pb["target"]][0:3]
    t1 = time.perf_counter()
    params, _, _ = sample_realizations(R, pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"], obs, xt, yt, rng=np.random.default_rng(1), log_rows=False)
    host_s = time.perf_counter() - t1
    spec = FlowSpec(well_xy=np.array([[w[0], w[1]] for w in pb["wells"]], dtype=float), xtarget=float(xt), ytarget=float(yt), rtarget=float(rt),
                    npaths=P, duration=float(pb["duration"]), base=float(pb["base"]), spacing=float(pb["spacing"]), umbra=float(pb["umbra"]),
                    confined=bool(pb["confined"]), tol=float(pb["tol"]), maxstep=float(pb["maxstep"]))
    eng.run_exact(spec, params)
    t2 = time.perf_counter()
    res = eng.run_exact(spec, params)
    eng_s = time.perf_counter() - t2
    t3 = time.perf_counter()
    ProbabilityField.from_counts(res["geom"], res["counts"], res["total_weight"])
    fld_s = time.perf_counter() - t3
    return {"value": att / secs, "unit": "DOPRI5 attempts/s", "realizations_per_s": R * steps / secs, "ms_per_call": 1e3 * secs / steps, "ms_each_call": each, "ms_per_call_median": float(np.median(each)), "steps": steps,
            "host_sampling_ms_per_call": 1e3 * host_s, "engine_run_exact_ms_per_call": 1e3 * eng_s, "probabilityfield_from_counts_ms": 1e3 * fld_s,
            "grid": [int(pf.nrows), int(pf.ncols)], "total_weight": float(pf.total_weight),
            "affected_realizations": eng.last_stats.get("affected_realizations"),
            "api": "oneka.stochastic.create_stochastic_capturezone(target, npaths, duration, nrealizations, base, c_dist, p_dist, t_dist, stochastic_wells, "
                   "observations, spacing, umbra, confined, tol, maxstep) -> ProbabilityField: sample_realizations (host, one core) + Engine.run_exact + from_counts, wall clock"}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from onekapy_b200 import parallel
    from onekapy_b200.engine import Engine

    rank, world, group = parallel.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if group is not None:
        dist.barrier()
    eng = Engine(local)
    if group is not None and not args.torch_allreduce:
        eng.init_comm(group)                       # the count-grid allreduce goes through the C ABI (oneka_allreduce_counts)
    cx = Ctx(eng, rank, world, group, torch, dist)
    sampler = ClockSampler(local)
    cx.max_mhz = sampler.max_mhz
    ff = "off" if args.farfield == "off" else "auto"
    try:                                           # the rasteriser's roofline denominators (rank-local, ~50 ms)
        cx.red = {"l2_sector_per_lane": eng.red_probe(4)[0], "l2_lane_private": eng.red_probe(0)[0], "l2_warp_contended": eng.red_probe(1)[0],
                  "shared_lane_private": eng.red_probe(2)[0], "shared_warp_contended": eng.red_probe(3)[0],
                  "unit": "1e9 atomic word operations/s, oneka_red_probe"}
    except Exception as exc:
        print("bench: red probe failed (%r)" % (exc,), file=sys.stderr)

    # ---- the headline leg ----
    wl = args.workload
    if wl == "c4full":
        r0, r1 = parallel.shard_range(C4FULL_TOTAL, rank, world)
        main = run_leg(cx, "c4full", "c4", args.realizations or (r1 - r0), args.npaths, args.steps, args.warmup, args.seed, farfield=ff,
                       e2e_steps=0 if args.no_e2e else 1, parity_n=2, sampler=sampler, scaling="strong", total=C4FULL_TOTAL,
                       note="BASELINE.json configs[3]: 1M realizations x 1000 paths sharded over the GPUs, one allreduce per step")
    else:
        main = run_leg(cx, wl, wl, args.realizations, args.npaths, args.steps, args.warmup, args.seed, farfield=ff, unconfined=args.unconfined,
                       full_pilot=(wl != "c4" or (args.realizations or 2048) <= 4096), e2e_steps=0 if args.no_e2e else max(1, args.steps),
                       parity_n={"c3": 6, "c4": 2, "c5": 1, "c1": 4}.get(wl, 2), sampler=sampler, exact_e2e=not args.no_e2e)
    spec, params, geom, problem, dp, ff_box, start = main["_state"]
    R, P = len(params), spec.npaths

    check = None
    if group is not None:
        per_rank_R = args.realizations if wl != "c4full" else R
        check = grid_check(cx, main["_state"], "c4" if wl == "c4full" else wl, per_rank_R or R, args.npaths or P, args.seed, args.unconfined,
                           recompute=list(range(world)) if wl != "c4full" else [world - 1])

    # ---- the drop-in call itself (N = 1); before the CPU baseline: its OpenMP workers keep spinning on the host cores for a while ----
    dropin = None
    if world == 1 and wl == "c3" and not args.no_e2e and not args.unconfined:
        try:
            dropin = dropin_leg(cx, R, P, max(1, min(5, args.steps)))
        except Exception as exc:
            dropin = {"error": repr(exc)}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import oracle as O
        O.build()
        c = cpu_sample(spec, params, geom)
        cpu = {"value": c["attempts"] / c["seconds"], "unit": "DOPRI5 attempts/s", "cores": c["cores"], "kind": "port",
               "sample": "%d of %d realizations x %d paths, same rows and lattice, OpenMP over realizations, %.1f s"
                         % (c["realizations"], R, P, c["seconds"]),
               "realizations_per_s": c["realizations"] / c["seconds"],
               "python_reference_note": "the reference itself is pure Python and cannot run on this box (it is not in the repo); executed in the "
                                        "survey container it made 2.5 k attempts/s per core on perham and 5.6 k on basic (BASELINE.md section 2), "
                                        "~600x slower per core than this C port, which reproduces its traces bit for bit"}

    # ---- the other configurations of BASELINE.json, short legs outside the headline's timed region ----
    legs = {}
    want = [] if args.legs == "none" else (["c4", "c4_direct", "c5", "c1", "c3_unconfined", "c4_strong"] + (["c4full"] if world == 8 else [])
                                           if args.legs == "auto" else [s for s in args.legs.split(",") if s])
    if args.legs == "auto" and (wl != "c3" or args.unconfined or args.realizations or args.npaths):
        want = []
    del dp
    main.pop("_state")
    for leg in want:
        try:
            if leg == "c4":
                r = run_leg(cx, leg, "c4", 2048, 0, 3, 3, args.seed, farfield="auto", full_pilot=True, e2e_steps=2, parity_n=2, exact_e2e=True)
            elif leg == "c4_direct":
                r = run_leg(cx, leg, "c4", 1024, 0, 2, 3, args.seed, farfield="off", full_pilot=True, parity_n=1,
                            note="the same 200-well field with the far-field compression off: the reference's own formulation (direct sum over all wells)")
            elif leg == "c5":
                r = run_leg(cx, leg, "c5", 2048, 0, 3, 3, args.seed, full_pilot=True, e2e_steps=2, parity_n=1, postprocess=True)
            elif leg == "c1":
                r = run_leg(cx, leg, "c1", 100, 100, 5, 3, args.seed, full_pilot=True, e2e_steps=3, parity_n=8, exact_e2e=True,
                            note="the reference's CPU-sized configuration: 10 000 particles on 148 SMs, launch-latency bound (~1 % of the FP64 peak by construction)")
            elif leg == "c3_unconfined":
                r = run_leg(cx, leg, "c3", 4000, 0, 3, 3, args.seed, unconfined=True, full_pilot=True, parity_n=2)
            elif leg == "c4_strong":
                r0, r1 = parallel.shard_range(C4_STRONG_TOTAL, rank, world)
                r = run_leg(cx, leg, "c4", r1 - r0, 0, 1, 3, args.seed, scaling="strong", total=C4_STRONG_TOTAL,
                            note="STRONG scaling: a fixed total of %d realizations x 1000 paths of the 200-well field (one GPU's share of BASELINE configs[3]) sharded over the N GPUs" % C4_STRONG_TOTAL)
            elif leg == "c4full":
                r0, r1 = parallel.shard_range(C4FULL_TOTAL, rank, world)
                r = run_leg(cx, leg, "c4", r1 - r0, 0, 1, 3, args.seed, scaling="strong", total=C4FULL_TOTAL, parity_n=2,
                            note="BASELINE.json configs[3] itself: 1M realizations x 1000 paths sharded over the GPUs, far field on, one allreduce per step")
                if group is not None:
                    r["grid_check"] = grid_check(cx, r["_state"], "c4", r1 - r0, 0, args.seed, False, recompute=[world - 1])
                r["wall_s_for_the_whole_configuration"] = r["ms_per_step"] * 1e-3
            else:
                raise ValueError("unknown leg %r" % leg)
            r.pop("_state", None)
            legs[leg] = r
        except Exception as exc:                                 # an extra leg must never cost the headline line
            import traceback
            traceback.print_exc()
            legs[leg] = {"error": repr(exc)}
            if group is not None:
                raise

    if rank == 0:
        try:
            host_rows = host_sampling_rate(problem, min(R, 20000))
        except Exception as exc:                             # informational only
            host_rows = {"error": repr(exc)}
        line = {"metric": "particle-steps/s", "value": main["value"], "unit": "DOPRI5 attempts/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": main["scaling"],
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "realizations_per_s": main["realizations_per_s"], "accepted_steps_per_s": main["accepted_steps_per_s"],
                "config": dict(main["config"], parallelism="realizations sharded over %d GPU(s), one allreduce of the count grid per step (%s)"
                               % (world, main["breakdown"]["allreduce_via"] or "single GPU: none")),
                "roofline": main["roofline"], "raster": main["raster"], "parity": main.get("parity"), "breakdown": main["breakdown"],
                "grid_check": check, "cpu_baseline": cpu, "e2e": main.get("e2e"), "e2e_exact": main.get("e2e_exact"), "e2e_dropin": dropin,
                "host_sampling": host_rows, "gpu_launches": main["gpu_launches"], "clocks": sampler.summary(), "configs": legs}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.barrier()
    eng.close()
    if group is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c3", "c4", "c5", "c4full"])
    ap.add_argument("--realizations", type=int, default=0, help="realizations per GPU per step (0 = workload default)")
    ap.add_argument("--npaths", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20200725)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end legs (kernel A/B runs while tuning)")
    ap.add_argument("--legs", default="auto", help="extra legs under `configs`: auto (all of them with the default c3 invocation, none otherwise), "
                                                   "none, or a comma list of c4,c4_direct,c5,c1,c3_unconfined,c4_strong,c4full")
    ap.add_argument("--torch-allreduce", action="store_true", help="sum the count grids with torch.distributed instead of oneka_allreduce_counts")
    ap.add_argument("--farfield", default="auto", choices=["auto", "off"],
                    help="auto: tiled far-field expansion of the well sum where it pays (default); off: direct sums only")
    ap.add_argument("--unconfined", action="store_true", help="confined=False: the head-dependent velocity of model.py:353-389")
    ap.add_argument("--umbra", type=float, default=0.0, help="lattice scans only: override the workload's umbra (turns the extra legs off)")
    ap.add_argument("--spacing", type=float, default=0.0, help="lattice scans only: override the workload's grid spacing (turns the extra legs off)")
    args = ap.parse_args()
    make_workload.umbra, make_workload.spacing = args.umbra or None, args.spacing or None
    if args.umbra or args.spacing:
        args.legs = "none"
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
