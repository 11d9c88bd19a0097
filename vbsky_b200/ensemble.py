"""Device-resident ensemble of trees + tip data and the batched calls into the CUDA path.

An ``Ensemble`` is what ``SeqData.prep_data`` (fasta.py:333-360) leaves behind in the reference --
``tds`` and padded ``tip_data_cs`` -- uploaded once; ``prune`` then evaluates
``vmap(prune_loglik) * counts`` (bdsky.py:331-343) and its gradient for any batch of
(tree, sample) units in one call.  PyTorch is used for device memory and streams only.
"""
import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .substitution import SubstitutionModel
from .tips import PackedTips
from .tree_data import TreeData


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class Ensemble:
    def __init__(self, tds: Sequence[TreeData], tips: Sequence[PackedTips], Q: SubstitutionModel,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise RuntimeError("vbsky_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.T = len(tds)
        if self.T < 1 or len(tips) != self.T:
            raise ValueError("need one tip-data entry per tree")
        self.n = int(tds[0].n)
        self.N = 2 * self.n - 1
        if any(td.n != self.n for td in tds):
            raise ValueError("all trees of an ensemble must have the same number of tips")
        self.S = int(tips[0].patterns.shape[0])
        if any(t.patterns.shape != (self.S, self.n) for t in tips):
            raise ValueError("tip patterns must be [S, n] with one S for the whole ensemble (use pad_tips)")
        self.Q = Q
        self._model = Q.c_struct()
        self.tds = list(tds)

        pc = np.ascontiguousarray(np.stack([td.parent_child for td in tds]), dtype=np.int32)
        cp = np.ascontiguousarray(np.stack([td.child_parent for td in tds]), dtype=np.int32)
        po = np.ascontiguousarray(np.stack([td.postorder for td in tds]), dtype=np.int32)
        st = np.ascontiguousarray(np.stack([td.sample_times for td in tds]), dtype=np.float64)
        patterns = np.ascontiguousarray(np.stack([t.patterns for t in tips]), dtype=np.uint8)
        counts = np.ascontiguousarray(np.stack([t.counts for t in tips]), dtype=np.float64)
        self._layout = C.c_void_p()
        self._tips = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.vbsky_layout_create(
                self.T, self.n, _lib.np_ptr(pc, C.c_int32), _lib.np_ptr(cp, C.c_int32),
                _lib.np_ptr(po, C.c_int32), _lib.np_ptr(st, C.c_double), C.byref(self._layout)))
            _lib.check(_lib.lib.vbsky_tips_create(
                self._layout, self.S, _lib.np_ptr(patterns, C.c_uint8), _lib.np_ptr(counts, C.c_double),
                C.byref(self._tips)))
        self._ws = None
        info = (C.c_int32 * 8)()
        _lib.check(_lib.lib.vbsky_layout_info(self._layout, info))
        self.depth_up, self.depth_down, self.max_levels = info[2], info[3], info[4]
        self.mean_stored_messages = info[6]  # per tree: non-root internal nodes that are not cherries (their messages travel through HBM)

    def close(self):
        if getattr(self, "_session", None):
            _lib.lib.vbsky_session_destroy(self._session)
            self._session = None
        if getattr(self, "_layout", None):
            _lib.lib.vbsky_layout_destroy(self._layout)
            self._layout = None
        if getattr(self, "_tips", None):
            _lib.lib.vbsky_tips_destroy(self._tips)
            self._tips = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def schedule(self, t: int):
        """Host copies of tree t's kernel schedules (tests / inspection)."""
        n = self.n
        up = np.zeros((n - 1, 4), dtype=np.int32)
        down = np.zeros((n - 1, 4), dtype=np.int32)
        dn = np.zeros(2 * n - 2, dtype=np.int32)
        lst = np.zeros(2 * n - 1)
        sib = np.zeros(2 * n - 1, dtype=np.int32)
        _lib.check(_lib.lib.vbsky_layout_get(self._layout, t, _lib.np_ptr(up, C.c_int32), _lib.np_ptr(down, C.c_int32),
                                              _lib.np_ptr(dn, C.c_int32), _lib.np_ptr(lst, C.c_double),
                                              _lib.np_ptr(sib, C.c_int32)))
        return dict(up=up, down=down, down_nodes=dn, lower_sampling_times=lst, siblings=sib)

    def set_counts(self, counts: np.ndarray):
        """Replace the pattern weights [T, S] on the device (synchronous)."""
        counts = np.ascontiguousarray(counts, dtype=np.float64).reshape(self.T, self.S)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib.vbsky_tips_set_counts(self._tips, _lib.np_ptr(counts, C.c_double)))

    def _workspace(self, nbytes: int) -> torch.Tensor:
        """One cached scratch buffer per Ensemble, grown on demand.  It is shared by every call made through this object,
        so all of them must be issued on ONE stream (or be ordered by the caller): two streams evaluating through the same
        Ensemble concurrently would overwrite each other's step tables.  Use one Ensemble per stream / host thread (the
        handles underneath are immutable and may be shared; tests/test_multidevice_gpu.py does exactly that)."""
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        return self._ws

    def prune(self, bl: torch.Tensor, unit_tree: torch.Tensor, want_grad: bool = True, want_site_ll: bool = False,
              precision: str = "f64", tile_range=None):
        """bl [U, 2n-2] fp64 (already multiplied by the clock rate), unit_tree [U] int32, both on the
        device.  Returns (ll [U], dll_dbl [U, 2n-2] or None, site_ll [U, S] or None).
        ``precision="f32"`` runs the fp32-storage kernels (inputs/outputs stay fp64; 1e-5 tolerance)."""
        if precision not in ("f64", "f32"):
            raise ValueError("precision must be 'f64' or 'f32'")
        f_ws = _lib.lib.vbsky_prune_workspace_bytes if precision == "f64" else _lib.lib.vbsky_prune_workspace_bytes_f32
        f_run = _lib.lib.vbsky_prune_value_and_grad if precision == "f64" else _lib.lib.vbsky_prune_value_and_grad_f32
        U = bl.shape[0]
        if bl.dtype != torch.float64 or bl.device != self.device or bl.shape != (U, 2 * self.n - 2):
            raise ValueError(f"bl must be a float64 [{U}, {2 * self.n - 2}] tensor on {self.device}")
        if unit_tree.dtype != torch.int32 or unit_tree.device != self.device or unit_tree.shape != (U,):
            raise ValueError("unit_tree must be an int32 [U] tensor on the ensemble's device")
        bl = bl.contiguous()
        with torch.cuda.device(self.device):
            need = f_ws(self._layout, self._tips, U, int(want_grad))
            if need == 0:
                raise _lib.VbskyError(-1, _lib.last_error())
            ws = self._workspace(need)
            ll = torch.empty(U, dtype=torch.float64, device=self.device)
            dll = torch.empty(U, 2 * self.n - 2, dtype=torch.float64, device=self.device) if want_grad else None
            site = torch.empty(U, self.S, dtype=torch.float64, device=self.device) if want_site_ll else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if tile_range is not None:
                if precision != "f64":
                    raise ValueError("tile ranges are an fp64 feature")
                if site is not None:
                    site.zero_()  # only the range's patterns are written
                _lib.check(_lib.lib.vbsky_prune_value_and_grad_tiles(
                    self._layout, self._tips, U, _ptr(unit_tree), _ptr(bl), C.byref(self._model), int(tile_range[0]), int(tile_range[1]),
                    _ptr(site), _ptr(ll), _ptr(dll), _ptr(ws), ws.numel(), C.c_void_p(stream)))
            else:
                _lib.check(f_run(
                    self._layout, self._tips, U, _ptr(unit_tree), _ptr(bl), C.byref(self._model),
                    _ptr(site), _ptr(ll), _ptr(dll), _ptr(ws), ws.numel(), C.c_void_p(stream)))
        return ll, dll, site

    def prune_spectral_grad(self, bl: torch.Tensor, unit_tree: torch.Tensor, g):
        """``vbsky_prune_spectral_grad``: returns (ll [U], out [U, 2n-2]) where ``out`` is eq. (9) with
        G = P diag(g) Pinv in place of Q.  For a parameter that only moves the spectrum, d ll/d theta =
        (bl * out).sum(1) with g = d d/d theta (``substitution.HKY_dkappa`` for kappa)."""
        U = bl.shape[0]
        if bl.dtype != torch.float64 or bl.device != self.device or bl.shape != (U, 2 * self.n - 2):
            raise ValueError(f"bl must be a float64 [{U}, {2 * self.n - 2}] tensor on {self.device}")
        if unit_tree.dtype != torch.int32 or unit_tree.device != self.device or unit_tree.shape != (U,):
            raise ValueError("unit_tree must be an int32 [U] tensor on the ensemble's device")
        g = np.ascontiguousarray(np.asarray(g, dtype=np.float64).reshape(4))
        bl = bl.contiguous()
        with torch.cuda.device(self.device):
            need = _lib.lib.vbsky_prune_workspace_bytes(self._layout, self._tips, U, 1)
            if need == 0:
                raise _lib.VbskyError(-1, _lib.last_error())
            ws = self._workspace(need)
            ll = torch.empty(U, dtype=torch.float64, device=self.device)
            out = torch.empty(U, 2 * self.n - 2, dtype=torch.float64, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib.vbsky_prune_spectral_grad(
                self._layout, self._tips, U, _ptr(unit_tree), _ptr(bl), C.byref(self._model),
                _lib.np_ptr(g, C.c_double), _ptr(ll), _ptr(out), _ptr(ws), ws.numel(), C.c_void_p(stream)))
        return ll, out

    def prune_param_grad(self, bl: torch.Tensor, unit_tree: torch.Tensor, dQ, dpi_dtheta=None):
        """``vbsky_prune_param_grad``: d ll[u] / d theta for an arbitrary direction dQ/dtheta [4, 4] of the rate matrix
        (need not commute with Q) plus, if the parameter also moves the root frequencies, ``dpi_dtheta`` [4].
        Returns (ll [U], dll_dtheta [U], per-branch contributions [U, 2n-2], d ll / d pi [U, 4])."""
        U = bl.shape[0]
        if bl.dtype != torch.float64 or bl.device != self.device or bl.shape != (U, 2 * self.n - 2):
            raise ValueError(f"bl must be a float64 [{U}, {2 * self.n - 2}] tensor on {self.device}")
        if unit_tree.dtype != torch.int32 or unit_tree.device != self.device or unit_tree.shape != (U,):
            raise ValueError("unit_tree must be an int32 [U] tensor on the ensemble's device")
        dQ = np.ascontiguousarray(np.asarray(dQ, dtype=np.float64).reshape(16))
        bl = bl.contiguous()
        with torch.cuda.device(self.device):
            need = _lib.lib.vbsky_prune_workspace_bytes(self._layout, self._tips, U, 1)
            if need == 0:
                raise _lib.VbskyError(-1, _lib.last_error())
            ws = self._workspace(need)
            ll = torch.empty(U, dtype=torch.float64, device=self.device)
            out = torch.empty(U, 2 * self.n - 2, dtype=torch.float64, device=self.device)
            dpi = torch.empty(U, 4, dtype=torch.float64, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib.vbsky_prune_param_grad(
                self._layout, self._tips, U, _ptr(unit_tree), _ptr(bl), C.byref(self._model), _lib.np_ptr(dQ, C.c_double),
                _ptr(ll), _ptr(out), _ptr(dpi), _ptr(ws), ws.numel(), C.c_void_p(stream)))
        total = out.sum(1)
        if dpi_dtheta is not None:
            total = total + dpi @ torch.as_tensor(np.asarray(dpi_dtheta, dtype=np.float64), device=self.device)
        return ll, total, out, dpi

    def dll_dkappa(self, bl: torch.Tensor, unit_tree: torch.Tensor, kappa: float) -> torch.Tensor:
        """d ll[u] / d kappa for an ensemble built with ``HKY(kappa)`` (one extra evaluation of the pruning kernel)."""
        from .substitution import HKY_dkappa

        _, out = self.prune_spectral_grad(bl, unit_tree, HKY_dkappa(kappa))
        return (bl * out).sum(1)

    def prune_host(self, bl: np.ndarray, unit_tree: np.ndarray, ll_out: Optional[np.ndarray] = None,
                   dll_out: Optional[np.ndarray] = None, want_grad: bool = True):
        """Host-buffer call (vbsky_prune_value_and_grad_host): bl [U, 2n-2] float64 and unit_tree [U]
        int32 NumPy arrays (pinned memory recommended) -> (ll [U], dll_dbl [U, 2n-2]) NumPy arrays.
        Copies in, computes, copies out and synchronises."""
        U = bl.shape[0]
        assert bl.dtype == np.float64 and bl.flags.c_contiguous and bl.shape == (U, 2 * self.n - 2)
        assert unit_tree.dtype == np.int32 and unit_tree.shape == (U,)
        with torch.cuda.device(self.device):
            self._ensure_session(U, 0)
            ll = np.empty(U) if ll_out is None else ll_out
            dll = None
            if want_grad:
                dll = np.empty((U, 2 * self.n - 2)) if dll_out is None else dll_out
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib.vbsky_prune_value_and_grad_host(
                self._session, U, unit_tree.ctypes.data_as(C.c_void_p), bl.ctypes.data_as(C.c_void_p),
                C.byref(self._model), ll.ctypes.data_as(C.c_void_p),
                None if dll is None else dll.ctypes.data_as(C.c_void_p), C.c_void_p(stream)))
        return ll, dll

    def _ensure_session(self, U: int, m: int):
        cap = getattr(self, "_session_cap", (0, -1))
        if getattr(self, "_session", None) is None or cap[0] < U or cap[1] < m:
            if getattr(self, "_session", None) is not None:
                _lib.lib.vbsky_session_destroy(self._session)
            self._session = C.c_void_p()
            cap = (max(U, cap[0]), max(m, cap[1]))
            _lib.check(_lib.lib.vbsky_session_create(self._layout, self._tips, cap[0], cap[1], C.byref(self._session)))
            self._session_cap = cap

    GRAD_KEYS = ("proportions", "root_proportion", "origin", "clock_rate", "R", "delta", "s", "rho")
    PARAM_KEYS = ("proportions", "root_proportion", "origin", "origin_start", "clock_rate", "R", "delta", "s", "rho")
    OPTIONAL_KEYS = ("grid",)  # prespecified skyline grid [U, m+1] (bdsky.py:317-319); absent = equidistant intervals

    @property
    def tiles(self) -> int:
        """Number of 256-pattern tiles per tree (the granularity of site-pattern-block sharding)."""
        return int(_lib.lib.vbsky_tips_tiles(self._tips))

    def loglik_device(self, params: dict, unit_tree: torch.Tensor, m: int, flags: int = 7, grads: Optional[dict] = None,
                      out: Optional[dict] = None, tile_range=None, exchange=None):
        """Raw device call of vbsky_loglik_value_and_grad (no autograd): ``params`` / ``grads`` map the
        vbsky_params / vbsky_param_grads field names to contiguous CUDA fp64 tensors ([U, ...]); ``out``
        may provide preallocated 'll', 'tree', 'data' [U] tensors.  Returns ``out``.

        Site-pattern-block sharding: ``tile_range=(begin, end)`` evaluates the data term over those 256-pattern tiles only
        (``vbsky_loglik_forward_partial``), then ``exchange(buf)`` is called with the contiguous fp64 tensor
        [ll_data (U) | d/d bl (U x (2n-2))] -- sum it over the ranks that share these units, in place (one all-reduce) --
        and ``vbsky_loglik_backward_complete`` finishes the evaluation."""
        U = unit_tree.shape[0]
        cin = _lib.Params()
        for k in self.PARAM_KEYS + self.OPTIONAL_KEYS:
            v = params.get(k)
            setattr(cin, k, 0 if v is None or v.numel() == 0 else v.data_ptr())
        cg = None
        if grads is not None:
            cg = _lib.ParamGrads()
            for k in self.GRAD_KEYS + self.OPTIONAL_KEYS:
                v = grads.get(k)
                setattr(cg, k, 0 if v is None or v.numel() == 0 else v.data_ptr())
        out = out if out is not None else {}
        for k in ("ll", "tree", "data"):
            if k not in out:
                out[k] = torch.empty(U, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            need = _lib.lib.vbsky_loglik_workspace_bytes(self._layout, self._tips, U, m)
            if need == 0:
                raise _lib.VbskyError(-1, _lib.last_error())
            ws = self._workspace(need)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if tile_range is None:
                _lib.check(_lib.lib.vbsky_loglik_value_and_grad(
                    self._layout, self._tips, U, _ptr(unit_tree), m, C.byref(cin), C.byref(self._model), flags,
                    _ptr(out["ll"]), _ptr(out["tree"]), _ptr(out["data"]), None if cg is None else C.byref(cg),
                    _ptr(ws), ws.numel(), C.c_void_p(stream)))
                return out
            _lib.check(_lib.lib.vbsky_loglik_forward_partial(
                self._layout, self._tips, U, _ptr(unit_tree), m, C.byref(cin), C.byref(self._model), flags,
                int(tile_range[0]), int(tile_range[1]), int(cg is not None), None if cg is None else C.byref(cg),
                _ptr(ws), ws.numel(), C.c_void_p(stream)))
            if exchange is not None:
                cnt = C.c_int64(0)
                ptr = _lib.lib.vbsky_loglik_exchange_buffer(self._layout, self._tips, U, m, _ptr(ws), C.byref(cnt))
                if not ptr:
                    raise _lib.VbskyError(-1, _lib.last_error())
                off = ptr - ws.data_ptr()
                exchange(ws[off: off + 8 * cnt.value].view(torch.float64))
            _lib.check(_lib.lib.vbsky_loglik_backward_complete(
                self._layout, self._tips, U, _ptr(unit_tree), m, C.byref(cin), C.byref(self._model), flags,
                _ptr(out["ll"]), _ptr(out["tree"]), _ptr(out["data"]), None if cg is None else C.byref(cg),
                _ptr(ws), ws.numel(), C.c_void_p(stream)))
        return out

    def loglik_host(self, params: dict, unit_tree: np.ndarray, m: int, flags: int = 7, grads: Optional[dict] = None,
                    out: Optional[dict] = None):
        """vbsky_loglik_value_and_grad_host: same as ``loglik_device`` with NumPy (ideally pinned) buffers;
        copies in, computes, copies out, synchronises."""
        U = unit_tree.shape[0]
        hp = lambda a: 0 if a is None or a.size == 0 else a.ctypes.data
        cin = _lib.Params()
        for k in self.PARAM_KEYS + self.OPTIONAL_KEYS:
            setattr(cin, k, hp(params.get(k)))
        cg = None
        if grads is not None:
            cg = _lib.ParamGrads()
            for k in self.GRAD_KEYS + self.OPTIONAL_KEYS:
                setattr(cg, k, hp(grads.get(k)))
        out = out if out is not None else {}
        for k in ("ll", "tree", "data"):
            if k not in out:
                out[k] = np.empty(U)
        with torch.cuda.device(self.device):
            self._ensure_session(U, m)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib.vbsky_loglik_value_and_grad_host(
                self._session, U, unit_tree.ctypes.data_as(C.c_void_p), m, C.byref(cin), C.byref(self._model), flags,
                out["ll"].ctypes.data_as(C.c_void_p), out["tree"].ctypes.data_as(C.c_void_p),
                out["data"].ctypes.data_as(C.c_void_p), None if cg is None else C.byref(cg), C.c_void_p(stream)))
        return out

    @staticmethod
    def kernel_timing(enable: bool):
        """Enable/disable CUDA-event timing of the fused pruning kernel; returns (ms_sum, launches)
        accumulated since the previous call."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        _lib.check(_lib.lib.vbsky_prune_timing(int(enable), C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    @staticmethod
    def last_stats():
        out = (C.c_int64 * 8)()
        _lib.check(_lib.lib.vbsky_prune_last_stats(out))
        return dict(launches=out[0], ctas=out[1], tile=out[2], smem=out[3], algorithmic_bytes=out[4])
