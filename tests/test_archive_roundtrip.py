"""SURVEY N4, second half: archive compatibility (oneka/archive.py:46-89).

The reference archives a run as a bz2-compressed pickle of a dict whose 'pfield' entry is the returned
ProbabilityField.  Pickle stores the class by module path, so the field this package returns must pickle as
`oneka.probabilityfield.ProbabilityField` with the reference's attribute names -- then archives written here load in the
reference (and its load_oneka consumers keep working), and archives written by the reference load here.
"""
import bz2
import os
import pickle
import pickletools
import subprocess
import sys

import numpy as np
import pytest

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ATTRS = ["deltax", "deltay", "nrows", "ncols", "xmin", "xmax", "ymin", "ymax", "total_weight", "pgrid", "rgrid"]


def _field():
    from onekapy_b200.lattice import LatticeGeom
    from oneka.probabilityfield import ProbabilityField
    geom = LatticeGeom.anchored(4.0, 4.0, 100.0, 200.0).expanded(60.0, 171.0, 150.0, 260.0)
    rng = np.random.default_rng(5)
    counts = rng.integers(0, 40, size=(geom.nrows, geom.ncols)).astype(np.uint32)
    return ProbabilityField.from_counts(geom, counts, 40.0), counts


def _dump(path, pfield):
    """The body of dump_oneka (oneka/archive.py:59-85) for a given file name."""
    oneka_dict = {'projectname': 'roundtrip', 'runtime': 1.5, 'target': 0, 'npaths': 10, 'duration': 3652.5, 'nrealizations': 40,
                  'base': 0.0, 'c_dist': (10, 20), 'p_dist': 0.25, 't_dist': (10, 15, 20), 'stochastic_wells': [(1.0, 2.0, 0.25, 100.0)],
                  'observations': [(0.0, 0.0, 10.0, 1.0)], 'buffer': 100, 'spacing': 4.0, 'umbra': 8.0, 'smooth': 2,
                  'confined': True, 'tol': 1.0, 'maxstep': 10.0, 'pfield': pfield}
    with bz2.BZ2File(path, "w") as fp:
        pickle.dump(oneka_dict, fp)


def test_pickles_under_the_reference_class_path(tmp_path):
    pf, counts = _field()
    assert type(pf).__module__ == "oneka.probabilityfield" and type(pf).__qualname__ == "ProbabilityField"
    path = str(tmp_path / "Oneka.bz2")
    _dump(path, pf)
    raw = bz2.BZ2File(path, "r").read()
    ops = [(op.name, arg) for op, arg, _ in pickletools.genops(raw)]
    globs = [arg for name, arg in ops if name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE", "BINUNICODE")]
    assert "oneka.probabilityfield" in globs and "ProbabilityField" in globs
    assert not any(isinstance(a, str) and a.startswith("onekapy_b200") for a in globs)       # nothing of this package's layout leaks in
    with bz2.BZ2File(path, "r") as fp:                                                        # load_oneka (archive.py:87-91)
        d = pickle.load(fp)
    back = d["pfield"]
    assert type(back) is type(pf) and sorted(vars(back)) == sorted(ATTRS)
    for a in ATTRS:
        assert np.array_equal(getattr(back, a), getattr(pf, a))
    assert back.pgrid.dtype == np.float64 and back.rgrid.dtype == np.bool_ and np.array_equal(back.pgrid, counts.astype(float))


_REF_SIDE = r"""
import bz2, pickle, sys
import numpy as np
np.float = float; np.bool = bool                      # the aliases the reference needs on NumPy >= 1.24
sys.path.insert(0, %r)
import oneka.probabilityfield as rp
assert rp.__file__.startswith(%r), rp.__file__
mode, path_in, path_out = sys.argv[1:4]
if mode == "load":                                    # an archive written by onekapy_b200 -> the reference's class
    with bz2.BZ2File(path_in, "r") as fp:
        d = pickle.load(fp)
    pf = d["pfield"]
    assert type(pf) is rp.ProbabilityField
    pf.expand(pf.xmin - 1.0, pf.xmax + 1.0, pf.ymin + 1.0, pf.ymax - 1.0)       # a method of the reference works on the loaded state
    np.savez(path_out, pgrid=pf.pgrid, rgrid=pf.rgrid, geom=[pf.deltax, pf.deltay, pf.nrows, pf.ncols, pf.xmin, pf.xmax, pf.ymin, pf.ymax, pf.total_weight])
else:                                                 # an archive written by the reference
    pf = rp.ProbabilityField(4.0, 4.0, 100.0, 200.0)
    pf.rasterize([100.0, 130.0, 150.0], [200.0, 215.0, 190.0], 8.0)
    pf.register(1.0)
    pf.rasterize([100.0, 90.0], [200.0, 240.0], 8.0)
    pf.register(1.0)
    with bz2.BZ2File(path_out, "w") as fp:
        pickle.dump({"pfield": pf, "spacing": 4.0}, fp)
""" % (REF, REF)


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree exists in the build container only")
def test_archive_crosses_to_the_reference_and_back(tmp_path):
    """Written here -> loaded by the EXECUTED reference (its own class, its own expand()); written by the reference -> loaded here."""
    pf, counts = _field()
    ours = str(tmp_path / "ours.bz2")
    _dump(ours, pf)
    script = str(tmp_path / "ref_side.py")
    open(script, "w").write(_REF_SIDE)
    env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
    out = str(tmp_path / "seen_by_reference.npz")
    subprocess.run([sys.executable, script, "load", ours, out], check=True, cwd=str(tmp_path), env=env)
    seen = np.load(out)
    assert seen["geom"][2] == pf.nrows and seen["geom"][3] == pf.ncols + 2          # expanded by one column each side
    assert np.array_equal(seen["pgrid"][:, 1:-1], pf.pgrid) and seen["geom"][8] == pf.total_weight
    theirs = str(tmp_path / "theirs.bz2")
    subprocess.run([sys.executable, script, "dump", "-", theirs], check=True, cwd=str(tmp_path), env=env)
    with bz2.BZ2File(theirs, "r") as fp:
        d = pickle.load(fp)
    from onekapy_b200.host.probabilityfield import ProbabilityField
    got = d["pfield"]
    assert type(got) is ProbabilityField and got.total_weight == 2.0 and got.pgrid.max() == 2.0
    assert sorted(vars(got)) == sorted(ATTRS)
    got.expand(got.xmin - 1.0, got.xmax - 1.0, got.ymin + 1.0, got.ymax - 1.0)                       # and this package's methods work on it
    assert got.pgrid.shape == (got.nrows, got.ncols)


def test_golden_archive_written_by_the_reference_loads_here(golden):
    """tests/golden/archive_ref.bz2: dump_oneka's layout written by the executed reference (make_golden.py)."""
    path = os.path.join(HERE, "golden", "archive_ref.bz2")
    with bz2.BZ2File(path, "r") as fp:
        d = pickle.load(fp)
    from oneka.probabilityfield import ProbabilityField
    pf = d["pfield"]
    assert type(pf) is ProbabilityField and sorted(vars(pf)) == sorted(ATTRS)
    g = golden("insert.npz")
    assert set(d) >= {"projectname", "target", "npaths", "duration", "nrealizations", "spacing", "umbra", "pfield"}
    assert pf.pgrid.dtype == np.float64 and pf.pgrid.shape == (pf.nrows, pf.ncols) and pf.total_weight == d["nrealizations"]
    assert np.array_equal(pf.pgrid, np.load(os.path.join(HERE, "golden", "archive_ref_pgrid.npz"))["pgrid"].astype(float))
    assert len(g.files) > 0
