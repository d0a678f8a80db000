#!/usr/bin/env python
"""bench.py -- throughput of the capture-zone hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c1|c5] [--impl ours|reference]

A "step" is one pass of the hot path (track + rasterise + register, then -- for N > 1 -- the
NCCL allreduce of the count grid) over one batch of pre-sampled realization rows:

  c3 (default)  data/perham.py field case, 10 000 realizations x 1000 paths PER GPU per step
                (BASELINE.json configs[2], the largest single-GPU configuration)
  c4            synthetic 200-well field, --realizations per GPU per step (default 2048) x 1000 paths
                (BASELINE.json configs[3] is 1M realizations over 8 GPUs = 61 such steps per GPU)
  c1            data/basic.py, 100 x 100 (the reference's CPU-runnable case)
  c5            data/basic.py field on a fine lattice (spacing 4, umbra 20 -> ~15x15-node windows, a
                4096 x 4096-class grid), 2048 realizations x 1000 paths: rasterisation stress

metric = particle-steps/s = DOPRI5 attempts per second summed over all particles and GPUs;
realizations/s is reported beside it.  Weak scaling: per-GPU work is fixed as N grows.

Timing: W >= 3 warm-up steps, then K steps between two barrier + synchronize brackets, timed on
the device with CUDA events on the launching stream, max over ranks.  L2 is flushed (256 MiB
write) before every timed step; the flush is inside the bracket (40 us against >100 ms steps).
`e2e` times the public call Engine.run() with HOST buffers: H2D of the parameter rows from pinned
memory, guarded capture on the estimated lattice, (allreduce,) crop and D2H of the count grid,
EVERY step, including the pilot pass that estimates the lattice (reuse_lattice=False: nothing is carried
over from one timed call to the next; the pilot's attempts are not counted as work, its time is).
`roofline` is the fused tracking+raster kernel against the FP64 pipe: algorithmic flops per
attempt = 257 + 90*Nw over the kernel's CUDA-event time, divided by an FP64
DFMA probe measured in the same process (MEASURED_PEAKS.json carries no FP64 figure).
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


# ------------------------------------------------------------------------------------------------
def make_workload(name, realizations, npaths, seed, unconfined=False):
    from onekapy_b200 import problems, synthetic
    from onekapy_b200.engine import FlowSpec
    if name == "c3":
        pb = problems.load("perham")
        R = realizations or 10000
        P = npaths or 1000
        label = "C3 perham field case (29 wells, 102 obs), %d realizations x %d paths per GPU per step" % (R, P)
    elif name == "c4":
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
def oracle_run(spec, params, geom, nthreads=0):
    """The CPU restatement on a fixed lattice; returns (seconds, attempts, steps)."""
    from oracle import oracle as O
    from onekapy_b200.engine import start_ring
    pf = O.Field(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget)
    pf.expand(geom.xmin + 0.5 * geom.deltax, geom.xmax - 0.5 * geom.deltax, geom.ymin + 0.5 * geom.deltay,
              geom.ymax - 0.5 * geom.deltay)
    start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, spec.npaths)
    t0 = time.perf_counter()
    res = O.capture(pf, 1, spec.well_xy, spec.base, spec.xtarget, spec.ytarget, spec.confined, params.q, params.cond,
                    params.poro, params.thick, params.coef, start, spec.duration, spec.umbra, spec.tol, spec.maxstep,
                    nthreads=nthreads, want_paths=False)
    return time.perf_counter() - t0, res["attempts"], res["steps"]


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
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    n0 = min(len(params), cores)
    t, att, stp = oracle_run(spec, params.slice(0, n0), geom, nthreads=cores)
    per = t / max(1, n0) * cores                       # seconds of one thread per realization
    n = int(min(len(params), max(n0, cores * max(1, int(target_s / max(per, 1e-9))))))
    if n > n0:
        t, att, stp = oracle_run(spec, params.slice(0, n), geom, nthreads=cores)
    else:
        n = n0
    return dict(seconds=t, realizations=n, attempts=att, steps=stp, cores=cores)


def run_reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (C restatement, OpenMP, all cores)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    spec, params, label = make_workload(args.workload, args.realizations, args.npaths, args.seed, args.unconfined)
    geom = cpu_geom(spec, params)
    # all the host threads the box has: torchrun exports OMP_NUM_THREADS=1 to its workers, which would leave the
    # reference arm on one core at N > 1
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    # calibrate the per-step sample so that warmup + steps finish in ~2 minutes
    t1, a1, _ = oracle_run(spec, params.slice(0, min(len(params), cores)), geom, nthreads=cores)
    per_real = t1 / min(len(params), cores)
    budget = 100.0 / max(1, args.steps + args.warmup)
    n = int(min(len(params), max(cores, int(budget / max(per_real, 1e-9)))))
    sub = params.slice(0, n)
    for _ in range(args.warmup):
        oracle_run(spec, sub, geom, nthreads=cores)
    tot_t, tot_a = 0.0, 0
    for _ in range(args.steps):
        t, a, _ = oracle_run(spec, sub, geom, nthreads=cores)
        tot_t += t
        tot_a += a
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
def run_ours(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    from onekapy_b200 import parallel
    from onekapy_b200.engine import Engine, start_ring
    from onekapy_b200.lattice import LatticeGeom

    rank, world, group = parallel.init_from_env()
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if rank == 0:
        ge.build()
    if group is not None:
        dist.barrier()
    eng = Engine(local)
    if args.farfield == "off":
        eng.farfield = "off"
    dev = eng.device
    spec, params, label = make_workload(args.workload, args.realizations, args.npaths, args.seed + rank, args.unconfined)
    R, P, nw = len(params), spec.npaths, len(spec.well_xy)

    def barrier():
        if group is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- set-up (untimed): rows to HBM, lattice from a pilot pass over all realizations ----
    start = start_ring(spec.xtarget, spec.ytarget, spec.rtarget, P)
    dp = eng.upload(spec, params, start)
    try:                                            # (large well fields: tile grid from a strided pilot, so that this full pass is fast too)
        eng._farfield_from_pilot(spec, params, dp)
    except Exception as exc:                        # set-up convenience only: without it the pass below runs on direct sums
        print("bench: far-field pilot skipped (%r)" % (exc,), file=sys.stderr)
    eng.reset_stats()
    eng.capture(spec, dp)
    bbox = parallel.reduce_bbox(eng.read_stats()["bbox"], group, dev if group is not None else None)
    geom = LatticeGeom.anchored(spec.spacing, spec.spacing, spec.xtarget, spec.ytarget).expanded(*bbox)
    counts = eng.new_counts(geom)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        flush_buf.fill_(1)                                   # L2 flush (126 MB L2)
        eng.capture(spec, dp, geom, counts)
        if group is not None:
            parallel.allreduce_counts(counts, group)

    for _ in range(max(3, args.warmup)):
        step()
    barrier()

    # ---- value: K steps, inputs resident in HBM, device-timed ----
    eng.reset_stats()
    eng.set_profiling(True)
    eng.kernel_ms(reset=True)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(torch.cuda.current_stream(dev))
    for _ in range(args.steps):
        step()
    ev1.record(torch.cuda.current_stream(dev))
    barrier()
    sampler.stop_flag = True
    ms = ev0.elapsed_time(ev1)
    ff_info = eng.farfield_info()           # tiled far-field expansion of the well sum, or None = direct sums
    launches = eng.launch_count() - launches0
    stats = eng.read_stats()
    kms = eng.kernel_ms(reset=True)
    eng.set_profiling(False)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    agg = torch.tensor([stats["attempts"], stats["steps"], stats["n_not_ok"], R * args.steps], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    attempts, steps_acc, n_not_ok, reals = [float(v) for v in agg.tolist()]
    value = attempts / (ms * 1e-3)
    rps = reals / (ms * 1e-3)

    # ---- rasteriser throughput (SURVEY.md 8d asks for it at C5): one more, untimed, step into a fresh grid ----
    fresh = eng.new_counts(geom)
    eng.reset_stats()
    eng.capture(spec, dp, geom, fresh)
    rst = eng.read_stats()
    ragg = torch.tensor([float(fresh.sum(dtype=torch.int64).item()), float(rst["steps"]), float(rst["exact_tests"])], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(ragg, op=dist.ReduceOp.SUM)
    cells_step, segs_step, exact_step = [float(v) for v in ragg.tolist()]
    step_s = ms * 1e-3 / args.steps
    raster = {"segments_per_s": segs_step / step_s, "cells_registered_per_s": cells_step / step_s,
              "cells_registered_per_segment": cells_step / max(1.0, segs_step),
              "exact_fp64_retests_per_segment": exact_step / max(1.0, segs_step),
              "note": "cells registered = bits set in the per-realization bitmaps = sum of the count grid; every one is also one count increment of the flush"}
    del fresh

    # ---- roofline of the fused tracking + raster kernel (this rank) ----
    probe_tf, _ = eng.fp64_probe(1 << 16)
    track_ms = kms["track_ms"] / max(1, kms["track_launches"])
    att_per_launch = stats["attempts"] / max(1, kms["track_launches"])
    achieved = att_per_launch * flops_per_attempt(nw) / (track_ms * 1e-3) / 1e12
    nominal = 148 * 64 * 2 * (sampler.max_mhz or 1965) * 1e6 / 1e12
    # DRAM traffic of the fused kernel from the ncu --set full capture in profiles/ (27.9 MB per 1000 realizations of
    # the perham field, dram__bytes_read.sum + dram__bytes_write.sum), scaled to this launch: the path is not HBM-bound
    # (far field on: 75.1 MB per 1000 realizations, profiles/r01_track_kernel_farfield_raw.csv -- the 28 KB coefficient
    # table of each realization and more bitmap write-back; still 0.05 % of the HBM peak)
    traffic = (75.1e6 if ff_info else 27.9e6) * (R / 1000.0) if args.workload == "c3" else None
    # with the far-field compression the kernel EXECUTES fewer flops than the reference's formulation needs: per evaluation
    # 20 (regional) + 16 per near well + 8 per polynomial term + ~16 of tile lookup, FMA = 2 (estimate from the mean near count)
    if ff_info:
        exec_flops = 6 * (20 + 16 * (ff_info["mean_near"] + 1.0) + 8 * ff_info["order"] + 16) + 137
    else:
        exec_flops = flops_per_attempt(nw)
    roofline = {"bound": "fp64", "kernel": "track_kernel<confined, raster%s>" % (", far field" if ff_info else ""),
                "achieved": achieved, "peak": probe_tf,
                "unit": "TFLOP/s", "frac": achieved / probe_tf, "traffic": traffic,
                "traffic_note": "bytes per launch scaled from profiles/r01_track_kernel%s_raw.csv (ncu --set full at R=1000); HBM is idle (< 0.1 %% of peak), the bound is the FP64 pipe / issue port" % ("_farfield" if ff_info else ""),
                "peak_source": "in-run DFMA probe (oneka_fp64_probe); MEASURED_PEAKS.json has no FP64 figure; nominal 148 SM x 64 lanes x 2 x max clock = %.1f" % nominal,
                "flops_per_attempt": flops_per_attempt(nw), "attempts_per_launch": att_per_launch, "kernel_ms_per_launch": track_ms,
                "flops_executed_per_attempt_estimate": exec_flops, "frac_executed_estimate": achieved * exec_flops / flops_per_attempt(nw) / probe_tf,
                "frac_note": "achieved/frac count the ALGORITHMIC flops of the reference's formulation (257 + 90 Nw per attempt, SURVEY 8d); "
                             "with the far-field compression active the kernel executes fewer (frac_executed_estimate), so frac may exceed 1",
                "flush_kernel_ms_per_step": kms["flush_ms"] / args.steps,
                "kernel_share_of_step": kms["track_ms"] / ms if world == 1 else None}

    # ---- e2e: the public call with host buffers, every step ----
    from onekapy_b200.engine import RealizationParams
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
    hp = RealizationParams(q=pin(params.q), cond=pin(params.cond), poro=pin(params.poro), thick=pin(params.thick), coef=pin(params.coef))
    if args.no_e2e:
        hp = None
    for _ in range(2 if hp is not None else 0):              # warm: the work lattice differs from the resident-input leg's,
        res = eng.run(spec, hp, group=group, reuse_lattice=False)   # so the bitmap workspace is re-allocated on the first call
    e2e_steps = max(1, args.steps)
    barrier()
    t0 = time.perf_counter()
    e_att = 0
    for _ in range(e2e_steps if hp is not None else 0):
        res = eng.run(spec, hp, group=group, reuse_lattice=False)
        e_att += res["stats"]["attempts"]
    barrier()
    e_s = time.perf_counter() - t0
    et = torch.tensor([e_s], dtype=torch.float64, device=dev)
    ea = torch.tensor([float(e_att)], dtype=torch.float64, device=dev)
    if group is not None:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        dist.all_reduce(ea, op=dist.ReduceOp.SUM)
    h2d = int(params.q.nbytes + params.cond.nbytes + params.poro.nbytes + params.thick.nbytes + params.coef.nbytes
              + spec.well_xy.nbytes + start.nbytes)
    d2h = int(res["counts"].nbytes + 10 * 8) if hp is not None else 0
    e2e = None if hp is None else {"value": float(ea.item()) / float(et.item()), "unit": "DOPRI5 attempts/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "realizations_per_s": world * R * e2e_steps / float(et.item()),
           "api": "Engine.run(spec, params_host, reuse_lattice=False) = H2D + pilot pass + guarded capture (+ allreduce) + crop + D2H, all inside every timed call", "steps": e2e_steps}

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

    if rank == 0:
        try:
            host_rows = host_sampling_rate(make_workload.problem, R)
        except Exception as exc:                             # informational only
            host_rows = {"error": repr(exc)}
        line = {"metric": "particle-steps/s", "value": value, "unit": "DOPRI5 attempts/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "realizations_per_s": rps, "accepted_steps_per_s": steps_acc / (ms * 1e-3),
                "config": {"workload": label, "wells": nw, "lattice": [geom.nrows, geom.ncols], "l2": "flushed before every step (256 MiB write)",
                           "paths_not_ok": n_not_ok, "farfield": ff_info, "parallelism": "realizations sharded over %d GPU(s), one NCCL allreduce of the count grid per step" % world},
                "roofline": roofline, "raster": raster, "cpu_baseline": cpu, "e2e": e2e, "host_sampling": host_rows,
                "gpu_launches": int(launches), "clocks": sampler.summary()}
        print(json.dumps(line), flush=True)
    if group is not None:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=["c1", "c3", "c4", "c5"])
    ap.add_argument("--realizations", type=int, default=0, help="realizations per GPU per step (0 = workload default)")
    ap.add_argument("--npaths", type=int, default=0)
    ap.add_argument("--seed", type=int, default=20200725)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (kernel A/B runs while tuning)")
    ap.add_argument("--farfield", default="auto", choices=["auto", "off"],
                    help="auto: tiled far-field expansion of the well sum where it pays (default); off: direct sums only")
    ap.add_argument("--unconfined", action="store_true", help="confined=False: the head-dependent velocity of model.py:353-389")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
