"""Problem definitions exported from the reference's data/*.py by tools/export_problems.py."""
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))


def _dist(v):
    # distributions are tuples in the reference (stochastic.py:297 tests `type(arg) is not tuple`)
    return tuple(v) if isinstance(v, list) else v


def load(name):
    """-> dict with the reference's argument names (lower case), tuples restored."""
    with open(os.path.join(_HERE, name + ".json")) as f:
        d = json.load(f)
    for k in ("c_dist", "p_dist", "t_dist"):
        d[k] = _dist(d[k])
    d["wells"] = [(w[0], w[1], w[2], _dist(w[3])) for w in d["wells"]]
    d["observations"] = [tuple(o) for o in d["observations"]]
    return d


def names():
    return sorted(f[:-5] for f in os.listdir(_HERE) if f.endswith(".json"))
