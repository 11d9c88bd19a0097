"""ctypes wrapper of the C/OpenMP oracle port (oracle/prune_ref.c) -- test infrastructure."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvbsky_oracle.so")


def _load():
    if not os.path.exists(_SO):
        subprocess.check_call(["make", "-C", _HERE])
    lib = C.CDLL(_SO)
    lib.vbsky_oracle_threads.restype = C.c_int
    lib.vbsky_oracle_prune.restype = C.c_int
    return lib


_lib = None


def set_threads(n: int) -> None:
    """Explicit OpenMP thread count (launchers such as torchrun export OMP_NUM_THREADS=1)."""
    global _lib
    _lib = _lib or _load()
    _lib.vbsky_oracle_set_threads(int(n))


def threads() -> int:
    global _lib
    _lib = _lib or _load()
    return _lib.vbsky_oracle_threads()


def prune(td, bl, Q, patterns, counts, want_grad=True, want_site_ll=False):
    """td: TreeData-like (postorder, child_parent, parent_child, siblings); patterns uint8 [S,n] masks.
    Returns (ll, grad [2n-2] or None, site_ll [S] or None)."""
    global _lib
    _lib = _lib or _load()
    n = len(td.sample_times)
    p = lambda a, t: np.ascontiguousarray(a, dtype=t)
    pc, cp, po, sib = p(td.parent_child, np.int32), p(td.child_parent, np.int32), p(td.postorder, np.int32), p(td.siblings, np.int32)
    bl, pi, P, d, Pinv = (p(x, np.float64) for x in (bl, Q.pi, Q.P, Q.d, Q.Pinv))
    patterns, counts = p(patterns, np.uint8), p(counts, np.float64)
    S = patterns.shape[0]
    site = np.empty(S) if want_site_ll else None
    grad = np.empty(2 * n - 2) if want_grad else None
    ll = C.c_double(0.0)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.vbsky_oracle_prune(C.c_int(n), vp(pc), vp(cp), vp(po), vp(sib), vp(bl), vp(pi), vp(P), vp(d), vp(Pinv),
                            C.c_int64(S), vp(patterns), vp(counts), vp(site), C.byref(ll), vp(grad))
    return ll.value, grad, site
