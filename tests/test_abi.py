"""CPU tests of the C-ABI boundary: the built library loads and exports every symbol include/orbx.h declares."""
import ctypes
import os
import re

import numpy as np
import pytest

import __graft_entry__ as ge
from multi_orbslam3_b200 import orbx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    ge.build()
    return ctypes.CDLL(orbx.LIB_PATH)


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "orbx.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(orbx_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_wrapper_agree():
    assert declared_symbols() == sorted(orbx.EXPORTS)


def test_library_exports_every_declared_symbol(built):
    for name in declared_symbols():
        assert hasattr(built, name), name


def test_structs_match_header():
    assert ctypes.sizeof(orbx.Params) == 40
    assert ctypes.sizeof(orbx.MatcherParams) == 16
    assert orbx.KP_DTYPE.itemsize == 28 and orbx.PROJQ_DTYPE.itemsize == 32


def test_every_entry_point_cites_reference():
    txt = open(os.path.join(ROOT, "include", "orbx.h")).read()
    assert txt.count("R/src/") + txt.count("R/include/") >= 20


def test_no_device_fails_loudly(built):
    """Without a GPU the product path must raise, never fall back to a CPU implementation."""
    if orbx.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(orbx.OrbxError):
        orbx.ORBextractor(1000, 1.2, 8, 20, 7)
    with pytest.raises(orbx.OrbxError):
        orbx.ORBmatcher(0.9, True)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "multi_orbslam3_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#\s*include.*oracle", r"liborb_oracle", r"orb_oracle", r"\borc_[a-z]"):
                    assert not re.search(pat, src, flags=re.M), (os.path.join(dirpath, f), pat)


def test_fast_segment_plan_host_logic():
    """Host-side geometry of the FAST kernel (no device needed): for every level width the cell grid follows the
    reference's formulas and the segments are whole cells, cover every cell and fit the 288-byte tile."""
    import ctypes as C
    L = orbx.lib()
    for w in list(range(40, 700)) + [752, 1241, 1920, 2048, 3840, 4128]:
        v = [C.c_int() for _ in range(5)]
        assert L.orbx_fast_segment_plan(w, *[C.byref(x) for x in v]) == 0
        ncols, wcell, segcells, nseg, wmax = [x.value for x in v]
        fw = w - 32
        assert ncols == max(int(np.float32(fw) / np.float32(30)), 0) if fw > 0 else ncols == 0
        if ncols == 0:
            continue
        assert wcell == int(np.ceil(np.float32(fw) / np.float32(ncols)))
        assert 1 <= segcells <= 16 and nseg == -(-ncols // segcells) and nseg * segcells >= ncols
        assert 0 < wmax <= 288 and wmax % 16 == 0
    assert L.orbx_fast_segment_plan(4, *[C.byref(C.c_int()) for _ in range(5)]) != 0
