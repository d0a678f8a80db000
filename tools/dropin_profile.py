"""cProfile of the drop-in call oneka.stochastic.create_stochastic_capturezone on the perham problem (GPU box)."""
import cProfile, os, pstats, sys, time
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
for _ in range(2):
    call()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    call()
pr.disable()
print("3 calls: %.1f ms each" % (1e3 * (time.perf_counter() - t0) / 3))
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
