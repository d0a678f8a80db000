"""Post-processing of a ProbabilityField on the device: the numbers behind the reference's
plots, without the plots.

  smooth_probability   Z of create_probability_plot   oneka/visualize.py:228-233
  impact_curve         (pr, area) of create_impact_plot  oneka/visualize.py:382-386
  decile_table         the table oneka() logs          oneka/oneka.py:278-288
  deterministic_area   area of create_deterministic_plot  oneka/visualize.py:312-331

The capture-zone grids of this package are integer valued (every realization registers with weight 1.0,
oneka/stochastic.py:265), so the "sort the whole grid" of the impact curve is a histogram of the counts.
"""
import ctypes as C

import numpy as np

from .. import _cabi
from ..engine import default_engine, OnekaError

ACRE = 4046.86      # m^2 per acre, oneka/oneka.py:286


class CaptureZoneError(Exception):
    """Empty capture zone (oneka/visualize.py:57-63)."""


def _counts_of(pf):
    if pf.total_weight <= 0:
        raise CaptureZoneError('Empty capture zone')                    # visualize.py:224-226
    c = np.asarray(pf.pgrid)
    ci = np.rint(c).astype(np.uint32)
    if not np.array_equal(ci, c):
        raise OnekaError("post-processing expects an integer-valued pgrid (weight-1 realizations)")
    return np.ascontiguousarray(ci)


def gaussian_taps(sigma, truncate=4.0):
    """The normalised 1-D kernel scipy.ndimage.gaussian_filter1d builds (order 0)."""
    sd = float(sigma)
    lw = int(truncate * sd + 0.5)
    x = np.arange(-lw, lw + 1)
    phi = np.exp(-0.5 / (sd * sd) * x ** 2)
    return phi / phi.sum(), lw


def smooth_probability(pf, smooth, engine=None):
    """Z = pgrid/total_weight, Gaussian-smoothed with sigma = smooth nodes (mode='constant', cval=0)."""
    counts = _counts_of(pf)
    if not smooth > 0:
        return counts / pf.total_weight                                  # visualize.py:230-233
    eng = engine or default_engine()
    torch = eng.torch
    w, lw = gaussian_taps(smooth)
    d_counts = torch.as_tensor(counts.view(np.int32)).to(eng.device)
    tmp = torch.empty(counts.shape, dtype=torch.float64, device=eng.device)
    out = torch.empty_like(tmp)
    _cabi.check(eng._L.oneka_gaussian_smooth(eng._h, d_counts.data_ptr(), counts.shape[0], counts.shape[1],
                                             float(pf.total_weight), w.ctypes.data, int(lw), tmp.data_ptr(), out.data_ptr()))
    return out.cpu().numpy()


def count_histogram(pf, engine=None):
    """hist[c] = number of nodes captured by exactly c realizations, c = 0 .. total_weight."""
    counts = _counts_of(pf)
    eng = engine or default_engine()
    torch = eng.torch
    nbins = int(round(pf.total_weight)) + 1
    d_counts = torch.as_tensor(counts.view(np.int32)).to(eng.device)
    hist = torch.zeros(max(nbins, 2), dtype=torch.int64, device=eng.device)
    _cabi.check(eng._L.oneka_count_histogram(eng._h, d_counts.data_ptr(), counts.size, max(nbins, 2), hist.data_ptr()))
    eng.synchronize()
    return hist.cpu().numpy()[:nbins]


def impact_curve(spacing, pf, engine=None):
    """(pr, area): probability of capture exceeds pr[i] over area[i]  (visualize.py:382-386)."""
    hist = count_histogram(pf, engine)
    values = np.arange(len(hist) - 1, -1, -1)
    pr = np.repeat(values / pf.total_weight, hist[::-1])
    area = (np.arange(pr.shape[0]) + 1) * spacing ** 2
    return pr, area


def decile_table(pr, area):
    """Rows (pr, area m^2, area acres) for p = 0.05 .. 0.95  (oneka/oneka.py:283-286)."""
    rows = []
    for p in np.linspace(0.05, 0.95, 19):
        i = int(np.argmax(pr <= p))
        rows.append((float(pr[i]), float(area[i]), float(area[i]) / ACRE))
    return rows


def deterministic_area(pf, engine=None):
    """Area [m^2] of the nodes with pgrid > 0, each counted as one cell of the node grid (visualize.py:312-331)."""
    hist = count_histogram(pf, engine)
    X = np.linspace(pf.xmin, pf.xmax, pf.ncols)
    Y = np.linspace(pf.ymin, pf.ymax, pf.nrows)
    cell_area = (X[1] - X[0]) * (Y[1] - Y[0])
    return float((pf.nrows * pf.ncols - hist[0]) * cell_area)
