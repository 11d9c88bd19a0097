"""TEST INFRASTRUCTURE ONLY -- a torch-fp64-backed stand-in for the subset of JAX that the reference's hot-path
files use, so that the UNMODIFIED sources under /root/reference/vbsky (substitution, prune, util, tree_data, bdsky,
optim, prob/*, fasta.loop) can be imported and executed in this image, where JAX itself is absent.

Every primitive is evaluated eagerly on ``torch.float64`` CPU tensors:
  * ``lax.scan`` / ``fori_loop`` / ``cond`` are Python loops / branches (concrete values, no tracing);
  * ``vmap`` is a Python loop over the mapped axis followed by a stack of the outputs;
  * ``grad`` / ``value_and_grad`` use ``torch.autograd`` (the derivative rules of the same primitive ops);
  * ``custom_jvp`` runs the user's JVP rule and transposes it with autograd (the rule is linear in the tangents),
    which is what JAX's reverse mode does with a custom JVP.
Semantics that differ between NumPy and JAX follow JAX (written from its documented behaviour; the package is not
available to diff against): ``take`` / gather indices clamp, scatter (``.at[].set``) wraps negative indices,
sorts are stable, ``linspace`` is ``start*(1-i/div) + stop*(i/div)`` with the endpoint appended exactly,
``logit(x) = log(x/(1-x))``, ``clip = minimum(maximum(x, lo), hi)``.

``jax.random`` is NOT bit-compatible (threefry is not restated): keys are integer tuples and draws come from a NumPy
generator seeded by the key.  Parity tests therefore record the draws and feed the same numbers to the other side.
"""
import math
from typing import Any

import numpy as np
import torch

F64 = torch.float64
I64 = torch.int64


def _is_scalar(x):
    return isinstance(x, (int, float, bool, np.integer, np.floating, np.bool_))


def T(x) -> torch.Tensor:
    """Anything array-like -> torch tensor (float64 / int64 / bool), sharing autograd history if it has one."""
    if isinstance(x, Array):
        return x._t
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (bool, np.bool_)):
        return torch.tensor(bool(x))
    if isinstance(x, (int, np.integer)):
        return torch.tensor(int(x), dtype=I64)
    if isinstance(x, (float, np.floating)):
        return torch.tensor(float(x), dtype=F64)
    if isinstance(x, (list, tuple)) and any(isinstance(e, (Array, torch.Tensor)) for e in _flat_list(x)):
        return torch.stack([_promote_pair(T(e), F64_HINT(x)) for e in x])
    a = np.asarray(x)
    if a.dtype == object:
        raise TypeError(f"cannot convert {type(x)} to an array")
    if a.dtype.kind == "f":
        a = a.astype(np.float64)
    elif a.dtype.kind in "iu":
        a = a.astype(np.int64)
    return torch.from_numpy(np.ascontiguousarray(a))


def _flat_list(x):
    for e in x:
        if isinstance(e, (list, tuple)):
            yield from _flat_list(e)
        else:
            yield e


def F64_HINT(x):
    """Result dtype of a Python list of mixed scalars/arrays: float64 if any element is floating."""
    for e in _flat_list(x):
        if isinstance(e, (float, np.floating)):
            return F64
        if isinstance(e, (Array, torch.Tensor)) and T(e).dtype.is_floating_point:
            return F64
        if isinstance(e, np.ndarray) and e.dtype.kind == "f":
            return F64
    return None


def _promote_pair(t, dtype):
    return t if dtype is None or t.dtype == dtype else t.to(dtype)


def _fl(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype.is_floating_point else t.to(F64)


def _binary(a, b):
    a, b = T(a), T(b)
    if a.dtype != b.dtype:
        if a.dtype.is_floating_point or b.dtype.is_floating_point:
            a, b = _fl(a), _fl(b)
        elif a.dtype == torch.bool:
            a = a.to(b.dtype)
        elif b.dtype == torch.bool:
            b = b.to(a.dtype)
    return a, b


def _index(idx, dim_sizes=None):
    """Translate an index expression; integer-array indices clamp to the valid range like a JAX gather."""
    if isinstance(idx, tuple):
        return tuple(_index_one(i) for i in idx)
    return _index_one(idx)


def _index_one(i):
    if isinstance(i, Array):
        i = i._t
    if isinstance(i, np.ndarray):
        i = torch.from_numpy(np.ascontiguousarray(i if i.dtype.kind == "b" else i.astype(np.int64)))
    if isinstance(i, (np.integer,)):
        i = int(i)
    if isinstance(i, torch.Tensor) and i.dtype not in (torch.bool, I64):
        i = i.to(I64)
    return i


def _clamp_gather_index(t: torch.Tensor, idx):
    """JAX gathers clamp out-of-range integer indices (after wrapping negatives)."""
    items = idx if isinstance(idx, tuple) else (idx,)
    out, dim = [], 0
    for it in items:
        if it is None:
            out.append(it)
            continue
        if it is Ellipsis:
            out.append(it)
            dim = t.dim() - (len([j for j in items[items.index(it) + 1:] if j is not None]))
            continue
        if isinstance(it, torch.Tensor) and it.dtype == I64:
            size = t.shape[dim]
            it = torch.where(it < 0, it + size, it).clamp(0, max(size - 1, 0))
        elif isinstance(it, torch.Tensor) and it.dtype == torch.bool:
            dim += it.dim() - 1
        out.append(it)
        dim += 1
    return tuple(out) if isinstance(idx, tuple) else out[0]


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, _index(idx)

    def _apply(self, v, op):
        base = self.arr._t
        v = T(v)
        if base.dtype != v.dtype:
            if base.dtype.is_floating_point or v.dtype.is_floating_point:
                base, v = _fl(base), _fl(v)
            else:
                v = v.to(base.dtype)
        out = base.clone()
        if op == "set":
            out[self.idx] = v
        elif op == "add":
            out[self.idx] = out[self.idx] + v
        else:
            raise NotImplementedError(op)
        return Array(out)

    def set(self, v):
        return self._apply(v, "set")

    def add(self, v):
        return self._apply(v, "add")


class Array:
    """Minimal ``jnp.ndarray`` look-alike over a torch tensor."""

    __array_priority__ = 1000
    __slots__ = ("_t",)

    def __init__(self, t):
        self._t = t if isinstance(t, torch.Tensor) else T(t)

    # ---- numpy interop -------------------------------------------------------------------------------------------
    def __array__(self, dtype=None, copy=None):
        a = self._t.detach().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __repr__(self):
        return f"ShimArray({self._t.detach().numpy()!r})"

    # ---- shape protocol ------------------------------------------------------------------------------------------
    shape = property(lambda s: tuple(s._t.shape))
    ndim = property(lambda s: s._t.dim())
    size = property(lambda s: s._t.numel())
    dtype = property(lambda s: np.dtype({F64: "float64", I64: "int64", torch.bool: "bool", torch.float32: "float32",
                                         torch.int32: "int32"}[s._t.dtype]))
    T = property(lambda s: Array(s._t.T if s._t.dim() == 2 else s._t))
    at = property(lambda s: _At(s))
    real = property(lambda s: s)

    def __len__(self):
        if self._t.dim() == 0:
            raise TypeError("len() of unsized object")
        return self._t.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield Array(self._t[i])

    def __getitem__(self, idx):
        idx = _clamp_gather_index(self._t, _index(idx))
        return Array(self._t[idx])

    def __float__(self):
        return float(self._t.detach())

    def __int__(self):
        return int(self._t.detach())

    def __index__(self):
        if self._t.dtype.is_floating_point:
            raise TypeError("float array used as an index")
        return int(self._t)

    def __bool__(self):
        return bool(self._t.detach())

    def __hash__(self):
        return id(self)

    def item(self):
        return self._t.detach().item()

    def tolist(self):
        return self._t.detach().tolist()

    def astype(self, dt):
        dt = np.dtype(dt)
        return Array(self._t.to({"f": F64, "i": I64, "u": I64, "b": torch.bool}[dt.kind]))

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self._t.reshape(shape))

    def ravel(self):
        return Array(self._t.reshape(-1))

    def flatten(self):
        return self.ravel()

    def diagonal(self):
        return Array(torch.diagonal(self._t))

    def copy(self):
        return Array(self._t.clone())

    # ---- reductions ----------------------------------------------------------------------------------------------
    def sum(self, axis=None, keepdims=False):
        t = self._t.to(I64) if self._t.dtype == torch.bool else self._t
        return Array(t.sum() if axis is None else t.sum(axis, keepdim=keepdims))

    def mean(self, axis=None):
        t = _fl(self._t)
        return Array(t.mean() if axis is None else t.mean(axis))

    def max(self, axis=None):
        return Array(self._t.max() if axis is None else self._t.max(axis).values)

    def min(self, axis=None):
        return Array(self._t.min() if axis is None else self._t.min(axis).values)

    def all(self, axis=None):
        return Array(self._t.bool().all() if axis is None else self._t.bool().all(axis))

    def any(self, axis=None):
        return Array(self._t.bool().any() if axis is None else self._t.bool().any(axis))

    def argsort(self, axis=-1):
        return Array(torch.argsort(self._t, dim=axis, stable=True))

    def cumsum(self, axis=0):
        return Array(torch.cumsum(self._t, axis))

    def dot(self, o):
        return dot(self, o)

    # ---- arithmetic ----------------------------------------------------------------------------------------------
    def _bin(self, o, f, r=False):
        if isinstance(o, (dict, str)) or o is None:
            return NotImplemented
        a, b = _binary(self, o)
        return Array(f(b, a) if r else f(a, b))

    __add__ = lambda s, o: s._bin(o, torch.add)
    __radd__ = lambda s, o: s._bin(o, torch.add, True)
    __sub__ = lambda s, o: s._bin(o, torch.sub)
    __rsub__ = lambda s, o: s._bin(o, torch.sub, True)
    __mul__ = lambda s, o: s._bin(o, torch.mul)
    __rmul__ = lambda s, o: s._bin(o, torch.mul, True)
    __truediv__ = lambda s, o: s._bin(o, lambda a, b: torch.true_divide(_fl(a), _fl(b)))
    __rtruediv__ = lambda s, o: s._bin(o, lambda a, b: torch.true_divide(_fl(a), _fl(b)), True)
    __floordiv__ = lambda s, o: s._bin(o, torch.floor_divide)
    __mod__ = lambda s, o: s._bin(o, torch.remainder)
    __pow__ = lambda s, o: s._bin(o, torch.pow)
    __rpow__ = lambda s, o: s._bin(o, torch.pow, True)
    __matmul__ = lambda s, o: s._bin(o, torch.matmul)
    __rmatmul__ = lambda s, o: s._bin(o, torch.matmul, True)
    __and__ = lambda s, o: s._bin(o, torch.bitwise_and)
    __rand__ = lambda s, o: s._bin(o, torch.bitwise_and, True)
    __or__ = lambda s, o: s._bin(o, torch.bitwise_or)
    __ror__ = lambda s, o: s._bin(o, torch.bitwise_or, True)
    __lt__ = lambda s, o: s._bin(o, torch.lt)
    __le__ = lambda s, o: s._bin(o, torch.le)
    __gt__ = lambda s, o: s._bin(o, torch.gt)
    __ge__ = lambda s, o: s._bin(o, torch.ge)
    __eq__ = lambda s, o: s._bin(o, torch.eq)
    __ne__ = lambda s, o: s._bin(o, torch.ne)
    __neg__ = lambda s: Array(-s._t)
    __pos__ = lambda s: s
    __abs__ = lambda s: Array(s._t.abs())
    __invert__ = lambda s: Array(~s._t)


def A(x) -> Array:
    return x if isinstance(x, Array) else Array(T(x))


# ---------------------------------------------------------------------------------------------------------------------
# jax.numpy
# ---------------------------------------------------------------------------------------------------------------------
pi = math.pi
inf = math.inf
nan = math.nan
newaxis = None
ndarray = Array
float64 = np.float64
int64 = np.int64


def _dtype(dt, default=F64):
    if dt is None:
        return default
    if dt is int:
        return I64
    if dt is float:
        return F64
    if dt is bool:
        return torch.bool
    return {"f": F64, "i": I64, "u": I64, "b": torch.bool}[np.dtype(dt).kind]


def _shape(s):
    if isinstance(s, (int, np.integer)):
        return (int(s),)
    return tuple(int(i) for i in s)


def array(x, dtype=None):
    t = T(x)
    return Array(t if dtype is None else t.to(_dtype(dtype)))


asarray = array


def zeros(shape, dtype=None):
    return Array(torch.zeros(_shape(shape), dtype=_dtype(dtype)))


def ones(shape, dtype=None):
    return Array(torch.ones(_shape(shape), dtype=_dtype(dtype)))


def full(shape, fill_value, dtype=None):
    v = T(fill_value)
    return Array(torch.broadcast_to(v, _shape(shape)).clone() if dtype is None
                 else torch.broadcast_to(v.to(_dtype(dtype)), _shape(shape)).clone())


def zeros_like(x):
    return Array(torch.zeros_like(T(x)))


def ones_like(x):
    return Array(torch.ones_like(T(x)))


def eye(n):
    return Array(torch.eye(int(n), dtype=F64))


def arange(*args, dtype=None):
    if any(isinstance(a, (float, np.floating)) for a in args):
        return Array(torch.arange(*[float(a) for a in args], dtype=F64))
    return Array(torch.arange(*[int(a) for a in args], dtype=_dtype(dtype, I64)))


def linspace(start, stop, num=50, endpoint=True):
    """JAX's formulation: ``start*(1-step) + stop*step`` with ``step = iota/div`` and the endpoint appended exactly."""
    start, stop = _fl(T(start)), _fl(T(stop))
    num = int(num)
    div = (num - 1) if endpoint else num
    if num > 1:
        step = torch.arange(div, dtype=F64) / div
        out = start * (1 - step) + stop * step
        if endpoint:
            out = torch.cat([out, stop.reshape(1)])
        return Array(out)
    return Array(start.reshape(1)) if num == 1 else Array(torch.zeros(0, dtype=F64))


def _un(f):
    def g(x):
        return Array(f(_fl(T(x))))
    return g


exp = _un(torch.exp)
log = _un(torch.log)
log1p = _un(torch.log1p)
expm1 = _un(torch.expm1)
sqrt = _un(torch.sqrt)
square = _un(torch.square)
tanh = _un(torch.tanh)
sign = _un(torch.sign)


def abs_(x):
    return Array(T(x).abs())


absolute = abs_


def isfinite(x):
    return Array(torch.isfinite(_fl(T(x))))


def isnan(x):
    return Array(torch.isnan(_fl(T(x))))


def isclose(a, b, rtol=1e-05, atol=1e-08, equal_nan=False):
    a, b = _binary(a, b)
    a, b = _fl(a), _fl(b)
    return Array(torch.isclose(a, b, rtol=rtol, atol=atol, equal_nan=equal_nan))


def allclose(a, b, rtol=1e-05, atol=1e-08):
    return Array(isclose(a, b, rtol, atol)._t.all())


def minimum(a, b):
    return Array(torch.minimum(*_binary(a, b)))


def maximum(a, b):
    return Array(torch.maximum(*_binary(a, b)))


def clip(x, a_min=None, a_max=None):
    out = A(x)
    if a_min is not None:
        out = maximum(out, a_min)
    if a_max is not None:
        out = minimum(out, a_max)
    return out


def where(c, a, b):
    a, b = _binary(a, b)
    return Array(torch.where(T(c).bool(), a, b))


def add(a, b):
    return A(a) + b


def subtract(a, b):
    return A(a) - b


def multiply(a, b):
    return A(a) * b


def sum_(x, axis=None):
    return A(x).sum(axis)


def mean(x, axis=None):
    return A(x).mean(axis)


def all_(x, axis=None):
    return A(x).all(axis)


def any_(x, axis=None):
    return A(x).any(axis)


def max_(x, axis=None):
    return A(x).max(axis)


def min_(x, axis=None):
    return A(x).min(axis)


def cumsum(x, axis=0):
    return Array(torch.cumsum(T(x), axis))


def diff(x, n=1, axis=-1):
    return Array(torch.diff(T(x), n=n, dim=axis))


def concatenate(xs, axis=0):
    ts = [T(x) for x in xs]
    if any(t.dtype.is_floating_point for t in ts):
        ts = [_fl(t) for t in ts]
    return Array(torch.cat(ts, dim=axis))


def append(x, v):
    return concatenate([A(x).ravel(), atleast_1d(v).ravel()])


def stack(xs, axis=0):
    ts = [T(x) for x in xs]
    if any(t.dtype.is_floating_point for t in ts):
        ts = [_fl(t) for t in ts]
    return Array(torch.stack(ts, dim=axis))


def atleast_1d(x):
    t = T(x)
    return Array(t.reshape(1) if t.dim() == 0 else t)


def split(x, indices_or_sections, axis=0):
    t = T(x)
    if isinstance(indices_or_sections, (int, np.integer)):
        return [Array(p) for p in torch.tensor_split(t, int(indices_or_sections), dim=axis)]
    idx = [int(i) for i in np.asarray(indices_or_sections).tolist()]
    return [Array(p) for p in torch.tensor_split(t, idx, dim=axis)]


def repeat(x, repeats, axis=None):
    t = T(x)
    if axis is None:
        return Array(t.reshape(-1).repeat_interleave(int(repeats)))
    return Array(t.repeat_interleave(int(repeats), dim=axis))


def tile(x, reps):
    return Array(T(x).repeat(*_shape(reps)))


def take(a, indices, axis=None, mode=None):
    """Gather with clamped indices (the behaviour the reference relies on: bdsky.py:130,138-139,195 read index m+1 for
    the most recent tip and its own example stays finite)."""
    t, i = T(a), T(indices).to(I64)
    if axis is None:
        t, axis = t.reshape(-1), 0
    size = t.shape[axis]
    i = torch.where(i < 0, i + size, i).clamp(0, size - 1)
    if i.dim() == 0:
        return Array(t.select(axis, int(i)))
    out = torch.index_select(t, axis, i.reshape(-1))
    return Array(out.reshape(t.shape[:axis] + tuple(i.shape) + t.shape[axis + 1:]))


def searchsorted(a, v, side="left"):
    a, v = _binary(a, v)
    return Array(torch.searchsorted(a.contiguous(), v.contiguous(), right=(side == "right")))


def argsort(x, axis=-1):
    return A(x).argsort(axis)


def sort(x, axis=-1):
    return Array(torch.sort(T(x), dim=axis, stable=True).values)


def dot(a, b):
    a, b = _binary(a, b)
    if a.dim() == 0 or b.dim() == 0:
        return Array(a * b)
    if a.dim() == 1 and b.dim() == 1:
        return Array((a * b).sum())
    return Array(torch.matmul(a, b))


def matmul(a, b):
    return Array(torch.matmul(*_binary(a, b)))


def einsum(spec, *ops):
    ts = [_fl(T(o)) for o in ops]
    return Array(torch.einsum(spec, *ts))


def diag(x):
    return Array(torch.diag(T(x)))


def outer(a, b):
    return Array(torch.outer(*_binary(a, b)))


def tril_indices(n, k=0, m=None):
    r, c = np.tril_indices(int(n), k, m)
    return Array(T(r)), Array(T(c))


def tril_indices_from(x, k=0):
    return tril_indices(T(x).shape[0], k)


def nan_to_num(x, nan=0.0, posinf=None, neginf=None):
    if _is_scalar(x):
        return float(np.nan_to_num(x))
    t = T(x)
    if not t.dtype.is_floating_point:
        return Array(t)
    return Array(torch.nan_to_num(t, nan=nan, posinf=posinf, neginf=neginf))


def logaddexp(a, b):
    a, b = _binary(a, b)
    return Array(torch.logaddexp(_fl(a), _fl(b)))


def expand_dims(x, axis):
    return Array(T(x).unsqueeze(axis))


def squeeze(x, axis=None):
    return Array(T(x).squeeze() if axis is None else T(x).squeeze(axis))


def reshape(x, shape):
    return A(x).reshape(shape)


def transpose(x, axes=None):
    return Array(T(x).permute(*axes) if axes is not None else T(x).T)


def broadcast_to(x, shape):
    return Array(torch.broadcast_to(T(x), _shape(shape)))


def ndim(x):
    return T(x).dim()


def shape(x):
    return tuple(T(x).shape)


class linalg:
    @staticmethod
    def norm(x, ord=None):
        return Array(torch.linalg.vector_norm(_fl(T(x)).reshape(-1)))


def vectorize(f):
    """``jnp.vectorize`` with the default (scalar) signature: broadcast the arguments, call ``f`` once per element."""
    def g(*args):
        ts = torch.broadcast_tensors(*[T(a) for a in args])
        shp = ts[0].shape
        if len(shp) == 0:
            return A(f(*[Array(t) for t in ts]))
        flat = [t.reshape(-1) for t in ts]
        outs = [T(f(*[Array(t[i]) for t in flat])) for i in range(flat[0].shape[0])]
        return Array(torch.stack(outs).reshape(shp) if outs else torch.zeros(shp, dtype=F64))
    return g


# ---------------------------------------------------------------------------------------------------------------------
# pytrees
# ---------------------------------------------------------------------------------------------------------------------
def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_map(f, tree, *rest):
    if isinstance(tree, dict):
        return type(tree)((k, tree_map(f, v, *[r[k] for r in rest])) for k, v in tree.items()) if type(tree) is not dict \
            else {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if _is_namedtuple(tree):
        return type(tree)(*[tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree)])
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    if tree is None:
        return None
    return f(tree, *rest)


def tree_leaves(tree):
    out = []
    tree_map(lambda x: out.append(x), tree)
    return out


def tree_flatten(tree):
    leaves = tree_leaves(tree)

    def unflatten(new):
        it = iter(new)
        return tree_map(lambda _: next(it), tree)

    return leaves, unflatten


def ravel_pytree(tree):
    leaves = [A(l) for l in tree_leaves(tree)]
    shapes = [l.shape for l in leaves]
    flat = concatenate([l.ravel() for l in leaves]) if leaves else zeros(0)

    def unravel(v):
        v, out, off = A(v), [], 0
        for s in shapes:
            cnt = int(np.prod(s, dtype=np.int64))
            out.append(v[off: off + cnt].reshape(s))
            off += cnt
        it = iter(out)
        return tree_map(lambda _: next(it), tree)

    return flat, unravel


# ---------------------------------------------------------------------------------------------------------------------
# lax
# ---------------------------------------------------------------------------------------------------------------------
def _tree_index(xs, i):
    return tree_map(lambda x: A(x)[i], xs)


def _tree_len(xs):
    return len(A(tree_leaves(xs)[0]))


def _tree_stack(items):
    first = items[0]
    if first is None:
        return None
    if isinstance(first, dict):
        return {k: _tree_stack([it[k] for it in items]) for k in first}
    if isinstance(first, (list, tuple)) and not _is_namedtuple(first):
        return type(first)(_tree_stack([it[i] for it in items]) for i in range(len(first)))
    if _is_namedtuple(first):
        return type(first)(*[_tree_stack([it[i] for it in items]) for i in range(len(first))])
    return stack([A(it) for it in items])


def scan(f, init, xs, length=None, reverse=False):
    n = length if xs is None else _tree_len(xs)
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, y = f(carry, None if xs is None else _tree_index(xs, i))
        ys[i] = y
    return carry, (_tree_stack(ys) if n else None)


def fori_loop(lower, upper, body, init):
    val = init
    for i in range(int(lower), int(upper)):
        val = body(i, val)
    return val


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def stop_gradient(x):
    return tree_map(lambda l: Array(T(l).detach()), x)


# ---------------------------------------------------------------------------------------------------------------------
# transformations
# ---------------------------------------------------------------------------------------------------------------------
def jit(f=None, static_argnums=None, static_argnames=None, **kw):
    if f is None:
        return lambda g: g
    return f


def vmap(f, in_axes=0, out_axes=0):
    def g(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), "vmap: in_axes must match the positional arguments"
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                for leaf in tree_leaves(a):
                    n = T(leaf).shape[ax]
                    break
                if n is not None:
                    break
        outs = []
        for i in range(n):
            sl = [a if ax is None else tree_map(lambda l, ax=ax: Array(T(l).select(ax, i)), a)
                  for a, ax in zip(args, axes)]
            outs.append(f(*sl))
        return _tree_stack(outs)
    return g


def _leafify(tree):
    """Fresh autograd leaves for every array leaf of a pytree."""
    return tree_map(lambda l: Array(_fl(T(l)).detach().clone().requires_grad_(True)), tree)


def value_and_grad(f, argnums=0, has_aux=False):
    def g(*args, **kw):
        args = list(args)
        p = _leafify(args[argnums])
        args[argnums] = p
        with torch.enable_grad():
            out = f(*args, **kw)
            val = T(out)
            leaves = [l._t for l in tree_leaves(p)]
            grads = torch.autograd.grad(val, leaves, allow_unused=True)
        it = iter(grads)
        gtree = tree_map(lambda l: Array(torch.zeros_like(l._t) if (gr := next(it)) is None else gr), p)
        return Array(val.detach()), gtree
    return g


def grad(f, argnums=0):
    vg = value_and_grad(f, argnums)
    return lambda *a, **k: vg(*a, **k)[1]


class custom_jvp:
    """``jax.custom_jvp``: the primal function is used when nothing upstream needs a gradient; otherwise the JVP rule
    is evaluated at detached primals with autograd-tracked zero tangents and its (linear) tangent output is
    transposed by ``torch.autograd.grad`` -- the VJP JAX's reverse mode derives from the same rule."""

    def __init__(self, fun, nondiff_argnums=()):
        self.fun, self.nondiff = fun, tuple(nondiff_argnums)
        self.jvp = None
        self.__name__ = getattr(fun, "__name__", "custom_jvp")
        self.__doc__ = fun.__doc__

    def defjvp(self, jvp):
        self.jvp = jvp
        return jvp

    def __call__(self, *args):
        diff_idx = [i for i in range(len(args)) if i not in self.nondiff]
        tracked = [(i, l) for i in diff_idx for l in tree_leaves(args[i])
                   if isinstance(l, (Array, torch.Tensor)) and T(l).requires_grad]
        if not tracked or not torch.is_grad_enabled():
            return self.fun(*args)
        nondiff = [args[i] for i in self.nondiff]
        primals = tuple(tree_map(lambda l: Array(T(l).detach()) if isinstance(l, (Array, torch.Tensor)) else l, args[i])
                        for i in diff_idx)
        with torch.enable_grad():
            tangents = tuple(tree_map(lambda l: Array(torch.zeros_like(_fl(T(l))).requires_grad_(True)), p) for p in primals)
            out, out_dot = self.jvp(*nondiff, primals, tangents)
            tleaves = [l._t for t in tangents for l in tree_leaves(t)]
            cot = torch.autograd.grad(T(out_dot), tleaves, allow_unused=True)
        # linearisation: out + <cot, x - stop_gradient(x)> carries exactly the rule's VJP into the outer graph
        res = T(out).detach()
        k = 0
        for i in diff_idx:
            for l in tree_leaves(args[i]):
                c = cot[k]
                k += 1
                if c is None or not isinstance(l, (Array, torch.Tensor)) or not T(l).requires_grad:
                    continue
                x = T(l)
                res = res + (c.detach() * (x - x.detach())).sum()
        return Array(res)


# ---------------------------------------------------------------------------------------------------------------------
# jax.random (NOT bit-compatible with threefry; see module docstring)
# ---------------------------------------------------------------------------------------------------------------------
class _Random:
    recorder = None  # tests may set a list here; every normal() draw is appended as (key, ndarray)
    injected = None  # or a callable (key, shape) -> ndarray overriding the draw

    @staticmethod
    def PRNGKey(seed):
        return (int(seed),)

    @staticmethod
    def split(key, num=2):
        return [tuple(key) + (i,) for i in range(int(num))]

    @staticmethod
    def _gen(key):
        return np.random.default_rng(list(key) + [len(key)])

    @classmethod
    def normal(cls, key, shape=(), dtype=None):
        shape = _shape(shape)
        x = cls.injected(key, shape) if cls.injected is not None else cls._gen(key).standard_normal(shape)
        if cls.recorder is not None:
            cls.recorder.append((tuple(key), np.array(x)))
        return Array(T(np.asarray(x, dtype=np.float64)))

    @classmethod
    def uniform(cls, key, shape=(), dtype=None, minval=0.0, maxval=1.0):
        return Array(T(cls._gen(key).uniform(minval, maxval, _shape(shape))))

    @classmethod
    def beta(cls, key, a, b, shape=None, dtype=None):
        return Array(T(cls._gen(key).beta(float(a), float(b), _shape(shape))))


random = _Random


# ---------------------------------------------------------------------------------------------------------------------
# jax.scipy / jax.nn
# ---------------------------------------------------------------------------------------------------------------------
def xlogy(x, y):
    x, y = _binary(x, y)
    x, y = _fl(x), _fl(y)
    ok = x != 0
    sx, sy = torch.where(ok, x, torch.ones_like(x)), torch.where(ok, y, torch.ones_like(y))
    return Array(torch.where(ok, sx * torch.log(sy), torch.zeros_like(sx * sy)))


def xlog1py(x, y):
    x, y = _binary(x, y)
    x, y = _fl(x), _fl(y)
    ok = x != 0
    sx, sy = torch.where(ok, x, torch.ones_like(x)), torch.where(ok, y, torch.ones_like(y))
    return Array(torch.where(ok, sx * torch.log1p(sy), torch.zeros_like(sx * sy)))


def logit(x):
    x = _fl(T(x))
    return Array(torch.log(x / (1 - x)))


def expit(x):
    return Array(torch.sigmoid(_fl(T(x))))


def gammaln(x):
    return Array(torch.lgamma(_fl(T(x))))


def softplus(x):
    x = _fl(T(x))
    return Array(torch.logaddexp(x, torch.zeros_like(x)))


def norm_logpdf(x, loc=0, scale=1):
    x, loc, scale = _fl(T(x)), _fl(T(loc)), _fl(T(scale))
    scale_sqrd = scale * scale
    return Array((torch.log(2 * math.pi * scale_sqrd) + (x - loc) ** 2 / scale_sqrd) / -2)


def gamma_logpdf(x, a, loc=0, scale=1):
    x, a, loc, scale = _fl(T(x)), _fl(T(a)), _fl(T(loc)), _fl(T(scale))
    y = (x - loc) / scale
    lp = xlogy(a - 1, y)._t - y - torch.lgamma(a) - torch.log(scale)
    return Array(torch.where(x < loc, torch.full_like(lp, -math.inf), lp))


def beta_logpdf(x, a, b, loc=0, scale=1):
    x, a, b, loc, scale = _fl(T(x)), _fl(T(a)), _fl(T(b)), _fl(T(loc)), _fl(T(scale))
    y = (x - loc) / scale
    lbeta = torch.lgamma(a) + torch.lgamma(b) - torch.lgamma(a + b)
    lp = xlogy(a - 1, y)._t + xlog1py(b - 1, -y)._t - lbeta - torch.log(scale)
    return Array(torch.where((y < 0) | (y > 1), torch.full_like(lp, -math.inf), lp))


def solve_triangular(a, b, lower=False, trans=0):
    a, b = _fl(T(a)), _fl(T(b))
    vec = b.dim() == 1
    out = torch.linalg.solve_triangular(a, b.reshape(-1, 1) if vec else b, upper=not lower)
    return Array(out.reshape(-1) if vec else out)


def id_print(x, **kw):
    return x


# ---------------------------------------------------------------------------------------------------------------------
# jax.example_libraries.optimizers.adagrad (third-party; restated from its published source, as in oracle/optim.py)
# ---------------------------------------------------------------------------------------------------------------------
def adagrad(step_size, momentum=0.9):
    ss = step_size if callable(step_size) else (lambda i: step_size)

    def init(x0):
        leaves, unflatten = tree_flatten(x0)
        return [(A(x), zeros_like(x), zeros_like(x)) for x in leaves], unflatten

    def update(i, g, state):
        triples, unflatten = state
        new = []
        for (x, g_sq, m), gl in zip(triples, tree_leaves(g)):
            gl = A(gl)
            g_sq = g_sq + square(gl)
            g_sq_inv_sqrt = where(g_sq > 0, 1.0 / sqrt(g_sq), 0.0)
            m = (1.0 - momentum) * (gl * g_sq_inv_sqrt) + momentum * m
            x = x - ss(i) * m
            new.append((x, g_sq, m))
        return new, unflatten

    def get_params(state):
        triples, unflatten = state
        return unflatten([t[0] for t in triples])

    return init, update, get_params
