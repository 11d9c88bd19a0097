"""The C-ABI shared library loads and exports every symbol that include/vbsky_b200.h declares, and the
Python binding table covers the same set.  No compute calls (CPU only)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "vbsky_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vbsky_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vbsky_b200 import _lib

    syms = _header_symbols()
    assert len(syms) >= 15
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported by the library"


def test_binding_table_matches_header():
    from vbsky_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _header_symbols()


def test_error_reporting_never_throws():
    from vbsky_b200 import _lib

    rc = _lib.lib.vbsky_hky(-1.0, None)
    assert rc == -1 and "NULL" in _lib.last_error() or "kappa" in _lib.last_error()
    assert _lib.lib.vbsky_version() >= 100


def test_argument_checks_run_before_any_device_work():
    """Bad shapes and NULL pointers are refused with VBSKY_EINVAL and a message, also on a box without a GPU."""
    import ctypes as C

    import numpy as np

    from vbsky_b200 import _lib

    L = _lib.lib
    dd = np.zeros(4)
    assert L.vbsky_hky_dkappa(0.0, _lib.np_ptr(dd, C.c_double)) == -1 and "kappa" in _lib.last_error()
    assert L.vbsky_hky_dkappa(2.7, None) == -1
    assert L.vbsky_hky_dkappa(2.7, _lib.np_ptr(dd, C.c_double)) == 0
    s = 1.0 / (4.7 * 4.7)
    assert np.allclose(dd, [-2 * s, -2 * s, 4 * s, 0.0], rtol=1e-15)
    assert L.vbsky_adagrad_update(5, None, None, None, None, 1.0, 0.9, None) == -1
    assert L.vbsky_adagrad_update(0, None, None, None, None, 1.0, 0.9, None) == 0
    assert L.vbsky_upgma_device(1, 1, None, None, None, None) == -1
    assert L.vbsky_upgma_device(1, 5000, C.c_void_p(8), C.c_void_p(8), C.c_void_p(8), None) == -1 and "4096" in _lib.last_error()
    assert L.vbsky_supgma_distances_device(0, 4, None, None, 0, None, None, None) == -1
    assert L.vbsky_snp_dists_device(None, 1, 1, None, 1, 2, None, None) == -1
    assert L.vbsky_snp_planes_bytes(3, 33) == 3 * 4 * 2 * 4
    assert L.vbsky_prune_spectral_grad(None, None, 1, None, None, None, None, None, None, None, 0, None) == -1


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "vbsky_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "oracle/" not in txt or f.endswith(".md"), f"{f} references oracle/"
