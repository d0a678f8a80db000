"""Export the reference's problem-definition modules (data/*.py) as JSON inputs.

The reference ships its field cases as Python modules of upper-case constants
(`/root/reference/data/basic.py:1-53`, `data/perham.py:1-158`, ...).  They are
*inputs* (well coordinates, observations, distributions), not code.  This script
reads them where they lie and writes one JSON document per problem under
`onekapy_b200/problems/`, so that `bench.py`, `smoke()` and the GPU tests can
load the same field cases on the GPU box, where `/root/reference` does not exist.

Run here (container) only:   python tools/export_problems.py
"""
import importlib
import json
import os
import sys

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "onekapy_b200", "problems")

NAMES = ["basic", "basic_deterministic", "perham", "carlos", "barnesville",
         "long_prairie", "paynesville"]


def dist(v):
    """Distribution spec -> JSON (scalar stays scalar, tuple -> list)."""
    if isinstance(v, tuple):
        return [float(t) for t in v]
    return float(v)


def main():
    sys.path.insert(0, REF)
    os.makedirs(OUT, exist_ok=True)
    for name in NAMES:
        m = importlib.import_module("data." + name)
        doc = {
            "projectname": m.PROJECTNAME,
            "source": "data/%s.py" % name,
            "target": int(m.TARGET),
            "npaths": int(m.NPATHS),
            "duration": float(m.DURATION),
            "nrealizations": int(m.NREALIZATIONS),
            "base": float(m.BASE),
            "c_dist": dist(m.C_DIST),
            "p_dist": dist(m.P_DIST),
            "t_dist": dist(m.T_DIST),
            "buffer": float(m.BUFFER),
            "spacing": float(m.SPACING),
            "umbra": float(m.UMBRA),
            "smooth": float(m.SMOOTH),
            "confined": bool(m.CONFINED),
            "tol": float(m.TOL),
            "maxstep": float(m.MAXSTEP),
            "wells": [[float(w[0]), float(w[1]), float(w[2]), dist(w[3])] for w in m.WELLS],
            "observations": [[float(t) for t in ob] for ob in m.OBSERVATIONS],
        }
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(doc, f, indent=0)
        print(name, len(doc["wells"]), "wells", len(doc["observations"]), "obs")


if __name__ == "__main__":
    main()
