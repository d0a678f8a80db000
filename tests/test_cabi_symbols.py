"""The C-ABI library loads without a GPU and exports exactly what include/oneka_b200.h declares;
the product never touches oracle/ ; without a device the product fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.build()
    from onekapy_b200 import _cabi
    return _cabi.load()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "oneka_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oneka_[a-z0-9_]+)\s*\(", src)))


def test_header_library_binding_agree(lib):
    from onekapy_b200 import _cabi
    hdr = header_symbols()
    assert hdr == sorted(_cabi.SYMBOLS)
    out = subprocess.check_output(["nm", "-D", "--defined-only", _cabi.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (oneka_[a-z0-9_]+)", out)))
    assert exported == hdr
    for name in hdr:
        assert getattr(lib, name) is not None
    assert lib.oneka_abi_version() == 1


def test_struct_layouts_match_header():
    from onekapy_b200 import _cabi
    # oneka_model_desc: 2 x int32, 6 x double, int64 ; oneka_lattice: 4 double, 2 int32, double ; oneka_stats: 6 u64 + 4 double
    assert ctypes.sizeof(_cabi.ModelDesc) == 8 + 6 * 8 + 8
    assert ctypes.sizeof(_cabi.Lattice) == 4 * 8 + 8 + 8
    assert ctypes.sizeof(_cabi.Stats) == 6 * 8 + 4 * 8


def test_built_for_sm100a_with_native_sass():
    from onekapy_b200 import _cabi
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.check_output([cuobjdump, "-lelf", _cabi.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    from onekapy_b200 import _cabi
    from onekapy_b200.engine import Engine, OnekaError
    assert not lib.oneka_create(0)
    assert b"no CUDA device" in lib.oneka_last_error()
    with pytest.raises(_cabi.OnekaError):
        _cabi.create(0)
    with pytest.raises(OnekaError):
        Engine(0)
    from onekapy_b200.host.model import Model
    with pytest.raises(OnekaError):
        Model(0.0, 1.0, 0.2, 10.0, [(0.0, 0.0, 0.1, 1.0)]).compute_discharge(1.0, 1.0)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/; the host emulation of the device code
    (tests/emu) is test infrastructure too and must stay invisible to the package (ONEKA_EMU only appears as the
    header's own #ifdef)."""
    bad = []
    for top in ("onekapy_b200", "oneka", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    if re.search(r"^\s*(from|import)\s+oracle\b|oneka_oracle|liboneka_oracle", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
                    if re.search(r"^\s*(from|import)\s+(tests\.)?emu\b|oneka_emu|define\s+ONEKA_EMU", txt, flags=re.M):
                        bad.append(os.path.join(dirpath, f))
    assert not bad, bad
