"""TEST INFRASTRUCTURE ONLY -- loads the UNMODIFIED reference sources (/root/reference/vbsky/*.py) on top of the
torch-fp64 JAX stand-in in ``core.py`` so that the reference's own arithmetic can be executed in this image and used
to pin the oracle restatement and to generate the golden vectors under tests/golden/.

    ref = oracle.refshim.load_reference()        # None if /root/reference is absent (e.g. on the GPU box)
    ref.prune.prune_loglik(...), ref.bdsky.loglik(...), ref.optim.loss(...), ref.fasta.SeqData.loop ...

What is the reference and what is not: every line of ``vbsky/*.py`` that runs is the reference's.  What runs
*underneath* it -- ``jax`` (core.py), ``msprime.HKY`` (restated from msprime's published construction; call site
substitution.py:75-78), ``jax.example_libraries.optimizers.adagrad`` -- is a stand-in written here, because those
packages are not installable offline.  ``tskit``, ``Bio``, ``ete3``, ``sh``, ``sklearn``, ``matplotlib`` are stubbed as
empty modules: the code paths that need them (newick / UPGMA / plotting) are not on the hot path and are not executed.

The fake modules are registered in ``sys.modules`` only while the reference is being imported and removed afterwards,
so that nothing else in the process ever sees a ``jax`` that is not JAX.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = "/root/reference"
_cached = None


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "vbsky"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


class _Stub:
    """Attribute sink for third-party names the executed code paths never touch."""

    def __init__(self, name):
        self._name = name

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Stub(f"{self._name}.{k}")

    def __call__(self, *a, **k):
        raise RuntimeError(f"{self._name} is a stub: this reference code path is not executable in the shim")

    def __mro_entries__(self, bases):
        return (object,)


def _stub_module(name, *names):
    m = types.ModuleType(name)
    for n in names:
        setattr(m, n, _Stub(f"{name}.{n}"))
    m.__getattr__ = lambda k: _Stub(f"{name}.{k}")
    return m


def _build_modules():
    import numpy as np

    from . import core as c

    jnp = _mod("jax.numpy")
    for k, v in vars(c).items():
        if not k.startswith("_") and k not in ("np", "torch", "math", "Any", "T", "A", "F64", "I64", "random"):
            setattr(jnp, k, v)
    for nm in ("abs", "sum", "all", "any", "max", "min"):
        setattr(jnp, nm, getattr(c, nm + "_"))
    jnp.ndarray = c.Array

    lax = _mod("jax.lax", scan=c.scan, fori_loop=c.fori_loop, cond=c.cond, stop_gradient=c.stop_gradient)
    rnd = _mod("jax.random", PRNGKey=c.random.PRNGKey, split=c.random.split, normal=c.random.normal,
               uniform=c.random.uniform, beta=c.random.beta)
    special = _mod("jax.scipy.special", xlogy=c.xlogy, xlog1py=c.xlog1py, logit=c.logit, expit=c.expit, gammaln=c.gammaln)
    norm = _mod("jax.scipy.stats.norm", logpdf=c.norm_logpdf)
    gamma = _mod("jax.scipy.stats.gamma", logpdf=c.gamma_logpdf)
    beta = _mod("jax.scipy.stats.beta", logpdf=c.beta_logpdf)
    stats = _mod("jax.scipy.stats", norm=norm, gamma=gamma, beta=beta)
    linalg = _mod("jax.scipy.linalg", solve_triangular=c.solve_triangular)
    jscipy = _mod("jax.scipy", special=special, stats=stats, linalg=linalg)
    nn = _mod("jax.nn", softplus=c.softplus, sigmoid=c.expit)
    tree_util = _mod("jax.tree_util", tree_map=c.tree_map, tree_multimap=c.tree_map, tree_leaves=c.tree_leaves,
                     tree_flatten=c.tree_flatten, register_pytree_node_class=lambda cls: cls)
    flatten_util = _mod("jax.flatten_util", ravel_pytree=c.ravel_pytree)
    host_callback = _mod("jax.experimental.host_callback", id_print=c.id_print)
    experimental = _mod("jax.experimental", host_callback=host_callback)
    optimizers = _mod("jax.example_libraries.optimizers", adagrad=c.adagrad)
    stax = _stub_module("jax.example_libraries.stax", "Relu", "Dense", "serial")
    example_libraries = _mod("jax.example_libraries", optimizers=optimizers, stax=stax)
    src_special = _mod("jax._src.scipy.special", expit=c.expit, logit=c.logit)
    src_scipy = _mod("jax._src.scipy", special=src_special)
    src = _mod("jax._src", scipy=src_scipy)
    config = types.SimpleNamespace(update=lambda *a, **k: None)
    jax = _mod("jax", numpy=jnp, lax=lax, random=rnd, scipy=jscipy, nn=nn, tree_util=tree_util, flatten_util=flatten_util,
               experimental=experimental, example_libraries=example_libraries, _src=src, config=config,
               jit=c.jit, vmap=c.vmap, grad=c.grad, value_and_grad=c.value_and_grad, custom_jvp=c.custom_jvp,
               tree_map=c.tree_map)
    jax.__path__ = []  # mark as package
    for m in (jscipy, stats, experimental, example_libraries, src, src_scipy):
        m.__path__ = []

    class _HKY:
        """msprime.HKY(kappa) with uniform equilibrium frequencies (restated; see oracle/substitution.py)."""

        def __init__(self, kappa, equilibrium_frequencies=None):
            pi = np.full(4, 0.25) if equilibrium_frequencies is None else np.asarray(equilibrium_frequencies, float)
            M = np.full((4, 4), 1.0)
            for i, j in ((0, 2), (1, 3), (2, 0), (3, 1)):
                M[i, j] = kappa
            M = M * pi[None, :]
            np.fill_diagonal(M, 0.0)
            M = M / M.sum(1).max()
            np.fill_diagonal(M, 1.0 - M.sum(1))
            self.transition_matrix = M
            self.root_distribution = pi

    msprime = _mod("msprime", HKY=_HKY)
    misc_util = _mod("numpy.distutils.misc_util", is_sequence=lambda x: isinstance(x, (list, tuple)))
    distutils = _mod("numpy.distutils", misc_util=misc_util)
    distutils.__path__ = []

    mods = {m.__name__: m for m in (jax, jnp, lax, rnd, jscipy, special, stats, norm, gamma, beta, linalg, nn, tree_util,
                                    flatten_util, experimental, host_callback, example_libraries, optimizers, stax, src,
                                    src_scipy, src_special, msprime, distutils, misc_util)}
    if importlib.util.find_spec("Bio") is None:
        from . import phylo

        bio = _stub_module("Bio")
        bio.__path__ = []
        bio.Phylo = _mod("Bio.Phylo", read=phylo.read, to_networkx=phylo.to_networkx, write=_Stub("Bio.Phylo.write"))
        bio.Phylo.__path__ = []
        mods["Bio"], mods["Bio.Phylo"] = bio, bio.Phylo
    for name in ("tskit", "sh", "ete3", "Bio", "Bio.SeqIO", "Bio.AlignIO", "Bio.Phylo", "Bio.Align", "Bio.Seq", "Bio.SeqRecord",
                 "Bio.Phylo.TreeConstruction", "Bio.Phylo.Consensus", "sklearn", "sklearn.linear_model", "matplotlib",
                 "matplotlib.pyplot"):
        if name not in mods and importlib.util.find_spec(name.split(".")[0]) is None:
            m = _stub_module(name)
            m.__path__ = []
            mods[name] = m
    return mods


class _Active:
    """Context manager that makes the stand-in modules importable again (the reference has a few function-level
    imports, e.g. ``from Bio import Phylo`` in ``TreeData.from_newick``)."""

    def __init__(self, mods):
        self.mods = mods

    def __enter__(self):
        self.saved = {k: sys.modules.get(k) for k in self.mods}
        sys.modules.update(self.mods)
        return self

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        return False


def load_reference(root: str = REFERENCE_ROOT, fresh: bool = False):
    """Import the unmodified reference package and return a namespace of its modules (cached unless ``fresh``)."""
    global _cached
    if _cached is not None and not fresh:
        return _cached
    if not os.path.isdir(os.path.join(root, "vbsky")):
        return None
    mods = _build_modules()
    saved_vbsky = {k: v for k, v in sys.modules.items() if k == "vbsky" or k.startswith("vbsky.")}
    for k in saved_vbsky:
        del sys.modules[k]
    sys.path.insert(0, root)
    try:
        with _Active(mods):
            ns = types.SimpleNamespace()
            for name in ("substitution", "tree_data", "util", "prune", "bdsky", "optim", "prob", "prob.transform",
                         "prob.distribution"):
                setattr(ns, name.replace(".", "_"), importlib.import_module("vbsky." + name))
            try:
                ns.fasta = importlib.import_module("vbsky.fasta")
            except Exception as e:  # fasta pulls in the whole preprocessing stack; optional
                ns.fasta, ns.fasta_error = None, e
        ns.jax = mods["jax"]
        ns.jnp = mods["jax.numpy"]
        from . import core

        ns.core = core
        ns.active = lambda: _Active(mods)
    finally:
        sys.path.remove(root)
        for k in [k for k in sys.modules if k == "vbsky" or k.startswith("vbsky.")]:
            del sys.modules[k]
        sys.modules.update(saved_vbsky)
    if not fresh:
        _cached = ns
    return ns
