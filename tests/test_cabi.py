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


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "vbsky_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".cuh", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle"
                assert "oracle/" not in txt or f.endswith(".md"), f"{f} references oracle/"
