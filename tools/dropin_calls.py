"""Per-call wall clock of the drop-in call (10 000 x 1000 perham), 12 calls, with the garbage collector on / off (GPU box)."""
import gc, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from onekapy_b200 import problems
from onekapy_b200.host.utilities import filter_obs
from onekapy_b200.engine import Engine
from oneka.stochastic import create_stochastic_capturezone
pb = problems.load("perham")
obs = filter_obs(pb["observations"], pb["wells"], pb["buffer"])
eng = Engine(0)
R, P = 10000, 1000
def call():
    return create_stochastic_capturezone(pb["target"], P, pb["duration"], R, pb["base"], pb["c_dist"], pb["p_dist"], pb["t_dist"], pb["wells"], obs,
                                         pb["spacing"], pb["umbra"], pb["confined"], pb["tol"], pb["maxstep"], rng=np.random.default_rng(1), engine=eng)
for label in ("gc on", "gc off", "gc on, ONEKA_PHASES"):
    if label == "gc off":
        gc.collect(); gc.disable()
    else:
        gc.enable()
    if "PHASES" in label:
        os.environ["ONEKA_PHASES"] = "1"
    ts, ph = [], []
    for _ in range(12):
        t0 = time.perf_counter(); call(); ts.append(round(1e3 * (time.perf_counter() - t0), 1))
        ph.append((eng.last_stats or {}).get("phases_ms"))
    print(label, ts, "gc counts", gc.get_count())
    if "PHASES" in label:
        for t, p in zip(ts, ph):
            print("   ", t, {k: round(v, 1) for k, v in (p or {}).items()})
