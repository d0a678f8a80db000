"""Post-processing on the device against the reference's own formulas:
scipy.ndimage.gaussian_filter (oneka/visualize.py:233), flip(sort(.)) (visualize.py:382-386),
the decile scan (oneka/oneka.py:283-286) and the cell count (visualize.py:316-331)."""
import numpy as np
import pytest

from helpers import geom

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def field(golden):
    from onekapy_b200.host.probabilityfield import ProbabilityField
    from onekapy_b200.lattice import LatticeGeom
    g = golden("sto_basic.npz")
    r = geom(g, "auto_")
    gm = LatticeGeom(r["deltax"], r["deltay"], r["xmin"], r["xmax"], r["ymin"], r["ymax"], r["nrows"], r["ncols"])
    return ProbabilityField.from_counts(gm, g["auto_counts"], r["total_weight"])


@pytest.mark.parametrize("smooth", [0.0, 1.0, 2.0, 4.0])
def test_smooth_probability_matches_scipy(field, smooth):
    import scipy.ndimage
    from onekapy_b200.host.postprocess import smooth_probability
    Z = field.pgrid / field.total_weight
    want = scipy.ndimage.gaussian_filter(Z, smooth, mode='constant', cval=0.0) if smooth > 0 else Z
    got = smooth_probability(field, smooth)
    assert got.shape == want.shape
    assert np.abs(got - want).max() < 1e-13


def test_impact_curve_and_deciles(field):
    from onekapy_b200.host.postprocess import impact_curve, decile_table, deterministic_area, count_histogram
    spacing = field.deltax
    pr_ref = np.flip(np.sort(field.pgrid / field.total_weight, axis=None))          # visualize.py:382
    area_ref = (np.arange(pr_ref.shape[0]) + 1) * spacing ** 2
    pr, area = impact_curve(spacing, field)
    assert np.array_equal(pr, pr_ref) and np.array_equal(area, area_ref)
    rows = decile_table(pr, area)
    for p, row in zip(np.linspace(0.05, 0.95, 19), rows):                            # oneka.py:283-286
        i = np.argmax(pr_ref <= p)
        assert row == (pr_ref[i], area_ref[i], area_ref[i] / 4046.86)
    hist = count_histogram(field)
    assert hist.sum() == field.pgrid.size and hist[0] == np.count_nonzero(field.pgrid == 0)
    X = np.linspace(field.xmin, field.xmax, field.ncols)
    Y = np.linspace(field.ymin, field.ymax, field.nrows)
    assert deterministic_area(field) == np.count_nonzero(field.pgrid > 0) * (X[1] - X[0]) * (Y[1] - Y[0])


def test_empty_field_raises():
    from onekapy_b200.host.postprocess import impact_curve, CaptureZoneError
    from onekapy_b200.host.probabilityfield import ProbabilityField
    with pytest.raises(CaptureZoneError):
        impact_curve(1.0, ProbabilityField(1.0, 1.0, 0.0, 0.0))
