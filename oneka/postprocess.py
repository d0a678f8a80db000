"""`oneka.postprocess`: the numbers behind the reference's plots (oneka/visualize.py), computed on the device;
the implementation lives in onekapy_b200.host.postprocess."""
from onekapy_b200.host.postprocess import *  # noqa: F401,F403
