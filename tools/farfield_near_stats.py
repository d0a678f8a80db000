#!/usr/bin/env python
"""Evaluation-weighted near-list lengths of the far-field tile grid on real traces (CPU only).

    python tools/farfield_near_stats.py [c3|c4] [realizations] [paths]

Traces come from the host emulation of the device code (tests/emu); for each tile count / (order, eta) the script prints the
mean number of near wells per VERTEX (what an evaluation pays for, as opposed to the per-tile mean `oneka_set_farfield`
reports) and a rough cycle estimate per evaluation (26 per padded near well + 9.3 per term + 40), to pick settings worth
timing on the GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench                                        # noqa: E402
from emu import emu                                 # noqa: E402
from onekapy_b200 import _cabi                      # noqa: E402
from onekapy_b200.engine import farfield_grid, start_ring   # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c3"
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    P = int(sys.argv[3]) if len(sys.argv) > 3 else 48
    spec, par, _ = bench.make_workload(name, R, P, 11)
    out = emu.capture(spec, par, start_ring(spec.xtarget, spec.ytarget, spec.rtarget, P), 2, max_verts=1500)
    mask = np.arange(out["verts"].shape[2])[None, None, :] < out["nverts"][:, :, None]
    pts = np.ascontiguousarray(out["verts"][mask])
    box = (pts[:, 0].min(), pts[:, 0].max(), pts[:, 1].min(), pts[:, 1].max())
    wxy = np.ascontiguousarray(spec.well_xy)
    w = np.ones(len(wxy))
    L = _cabi.load()
    print("%s: %d wells, %d vertices, box %.0f x %.0f m" % (name, len(wxy), len(pts), box[1] - box[0], box[3] - box[2]))
    for tiles in (64, 100, 144, 256):
        for order, eta in ((28, 0.3), (24, 0.25), (22, 0.2), (32, 0.35)):
            g = farfield_grid(box, tiles)
            res = np.zeros((len(pts), 2))
            near = np.zeros(len(pts), dtype=np.int32)
            _cabi.check(L.oneka_farfield_eval_host(len(wxy), wxy.ctypes.data, w.ctypes.data, spec.xtarget, spec.ytarget, g["x0"], g["y0"],
                                                   g["tile"], g["ntx"], g["nty"], order, eta, 0, len(pts), pts.ctypes.data,
                                                   res.ctypes.data, near.ctypes.data))
            nn = near[near >= 0]
            pad = (nn + 1) // 2 * 2
            print("  %3d tiles of %4.0f m, order %d, eta %.2f: near wells per vertex %.2f (padded %.2f, max %d); ~%.0f cycles per evaluation"
                  % (g["ntx"] * g["nty"], g["tile"], order, eta, nn.mean(), pad.mean(), nn.max(), 26 * pad.mean() + 9.3 * order + 40))


if __name__ == "__main__":
    main()
