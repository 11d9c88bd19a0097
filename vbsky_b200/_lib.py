"""ctypes binding of libvbsky_b200.so (the C ABI declared in include/vbsky_b200.h).

The shared library is built in-tree by ``vbsky_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is no CPU fallback: if the library is missing, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VBSKY_B200_LIB") or os.path.join(_HERE, "libvbsky_b200.so")  # env override: A/B builds


class VbskyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"[vbsky_b200 {code}] {msg}")
        self.code = code


class VbskyKeyError(VbskyError, KeyError):
    """Unknown sequence character (the reference raises KeyError, substitution.py:48)."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `make -C vbsky_b200/csrc` "
        "(or `python -c 'import __graft_entry__ as g; g.build()'`). vbsky_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)


class SubstModel(C.Structure):
    _fields_ = [("pi", C.c_double * 4), ("P", C.c_double * 16), ("d", C.c_double * 4), ("Pinv", C.c_double * 16)]


class Params(C.Structure):
    """vbsky_params: device pointers, one row per unit."""

    _fields_ = [(k, C.c_void_p) for k in ("proportions", "root_proportion", "origin", "origin_start", "clock_rate", "R", "delta", "s", "rho", "grid")]


class ParamGrads(C.Structure):
    """vbsky_param_grads."""

    _fields_ = [(k, C.c_void_p) for k in ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho", "grid")]


_vp = C.c_void_p
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)
_u8p = C.POINTER(C.c_uint8)

# name -> (restype, argtypes); every symbol of include/vbsky_b200.h must appear here
# (tests/test_cabi.py checks the header against this table and against the .so).
SIGNATURES = {
    "vbsky_last_error": (C.c_char_p, []),
    "vbsky_version": (C.c_int, []),
    "vbsky_device_count": (C.c_int, []),
    "vbsky_hky": (C.c_int, [C.c_double, C.POINTER(SubstModel)]),
    "vbsky_subst_rate_matrix": (C.c_int, [C.POINTER(SubstModel), _f64p]),
    "vbsky_hky_dkappa": (C.c_int, [C.c_double, _f64p]),
    "vbsky_prune_spectral_grad": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), _f64p, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vbsky_prune_param_grad": (C.c_int, [_vp, _vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), _f64p, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vbsky_encode_tips": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, _u8p]),
    "vbsky_compress_patterns": (C.c_int, [_u8p, C.c_int64, C.c_int64, _i64p, _u8p, _i64p, _i64p]),
    "vbsky_encode_tips_device": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, _i64p, _vp]),
    "vbsky_compress_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int64]),
    "vbsky_compress_patterns_device": (C.c_int, [C.c_int32, C.c_int64, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vbsky_snp_planes_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "vbsky_snp_pack_device": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, _vp]),
    "vbsky_snp_dists_device": (C.c_int, [_vp, C.c_int64, C.c_int64, _vp, C.c_int32, C.c_int32, _vp, _vp]),
    "vbsky_supgma_distances_device": (C.c_int, [C.c_int32, C.c_int32, _vp, _vp, C.c_int32, _vp, _vp, _vp]),
    "vbsky_upgma_device": (C.c_int, [C.c_int32, C.c_int32, _vp, _vp, _vp, _vp]),
    "vbsky_layout_create": (C.c_int, [C.c_int32, C.c_int32, _i32p, _i32p, _i32p, _f64p, C.POINTER(_vp)]),
    "vbsky_layout_destroy": (None, [_vp]),
    "vbsky_layout_info": (C.c_int, [_vp, _i32p]),
    "vbsky_layout_get": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _i32p, _f64p, _i32p]),
    "vbsky_layout_get_chunks": (C.c_int, [_vp, C.c_int32, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "vbsky_tips_create": (C.c_int, [_vp, C.c_int64, _u8p, _f64p, C.POINTER(_vp)]),
    "vbsky_tips_destroy": (None, [_vp]),
    "vbsky_tips_set_counts": (C.c_int, [_vp, _f64p]),
    "vbsky_prune_workspace_bytes": (C.c_size_t, [_vp, _vp, C.c_int32, C.c_int32]),
    "vbsky_prune_value_and_grad": (
        C.c_int,
        [_vp, _vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), _vp, _vp, _vp, _vp, C.c_size_t, _vp],
    ),
    "vbsky_prune_workspace_bytes_f32": (C.c_size_t, [_vp, _vp, C.c_int32, C.c_int32]),
    "vbsky_prune_value_and_grad_f32": (
        C.c_int,
        [_vp, _vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), _vp, _vp, _vp, _vp, C.c_size_t, _vp],
    ),
    "vbsky_prune_last_stats": (C.c_int, [_i64p]),
    "vbsky_prune_timing": (C.c_int, [C.c_int32, _f64p, _i64p]),
    "vbsky_local_flow_sample": (C.c_int, [C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vbsky_local_flow_backward": (C.c_int, [C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vbsky_adagrad_update": (C.c_int, [C.c_int64, _vp, _vp, _vp, _vp, C.c_double, C.c_double, _vp]),
    "vbsky_heights_forward": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "vbsky_heights_backward": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                         C.POINTER(ParamGrads), _vp, C.c_size_t, _vp]),
    "vbsky_bdsky_workspace_bytes": (C.c_size_t, [C.c_int32, C.c_int32]),
    "vbsky_bdsky_value_and_grad": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, _vp, _vp, C.c_int32,
                                             _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_size_t, _vp]),
    "vbsky_loglik_workspace_bytes": (C.c_size_t, [_vp, _vp, C.c_int32, C.c_int32]),
    "vbsky_loglik_value_and_grad": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), C.POINTER(SubstModel),
                                              C.c_int32, _vp, _vp, _vp, C.POINTER(ParamGrads), _vp, C.c_size_t, _vp]),
    "vbsky_tips_tiles": (C.c_int32, [_vp]),
    "vbsky_prune_value_and_grad_tiles": (
        C.c_int,
        [_vp, _vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), C.c_int32, C.c_int32, _vp, _vp, _vp, _vp, C.c_size_t, _vp],
    ),
    "vbsky_loglik_forward_partial": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), C.POINTER(SubstModel), C.c_int32,
                                               C.c_int32, C.c_int32, C.c_int32, C.POINTER(ParamGrads), _vp, C.c_size_t, _vp]),
    "vbsky_loglik_exchange_buffer": (_vp, [_vp, _vp, C.c_int32, C.c_int32, _vp, _i64p]),
    "vbsky_loglik_backward_complete": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), C.POINTER(SubstModel), C.c_int32,
                                                 _vp, _vp, _vp, C.POINTER(ParamGrads), _vp, C.c_size_t, _vp]),
    "vbsky_reduce_elbo_count": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "vbsky_reduce_elbo": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _vp, _vp, C.POINTER(ParamGrads), C.c_double, _vp, _vp]),
    "vbsky_session_create": (C.c_int, [_vp, _vp, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "vbsky_session_destroy": (None, [_vp]),
    "vbsky_prune_value_and_grad_host": (C.c_int, [_vp, C.c_int32, _vp, _vp, C.POINTER(SubstModel), _vp, _vp, _vp]),
    "vbsky_loglik_value_and_grad_host": (C.c_int, [_vp, C.c_int32, _vp, C.c_int32, C.POINTER(Params), C.POINTER(SubstModel),
                                                   C.c_int32, _vp, _vp, _vp, C.POINTER(ParamGrads), _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.restype = _res
    _fn.argtypes = _args


def last_error() -> str:
    return lib.vbsky_last_error().decode("utf-8", "replace")


def check(rc: int) -> int:
    if rc < 0:
        msg = last_error()
        if rc == -4:
            raise VbskyKeyError(rc, msg)
        raise VbskyError(rc, msg)
    return rc


def np_ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))
