"""Pins the oracle to the reference's OWN SOURCE, executed here.

``oracle/refshim`` imports the unmodified files under /root/reference/vbsky on top of a torch-fp64 stand-in for the JAX
subset they use (JAX itself is not installable in this image).  Every test below runs the reference's functions --
``SubstitutionModel.expm_action``, ``prune.prune_loglik`` and its custom JVP, ``bdsky._tree_loglik`` / ``loglik``,
``TreeData.*``, ``util.order_events``, ``fasta._process_tips`` / ``pad_tips``, ``optim.loss`` under ``value_and_grad`` -- and
asserts that the oracle restatement (``oracle/*.py``, the C port) returns the same numbers to ~1e-13.  Change one line of
the reference arithmetic (or of the oracle) and these fail; ``test_pin_is_sensitive_to_a_one_line_change`` demonstrates it.

/root/reference does not exist on the GPU box: there these tests are skipped and the committed fixtures under
tests/golden/ (generated from the same reference calls by tests/golden/make_golden.py) carry the pin.
"""
import math
import os
import shutil

import networkx as nx
import numpy as np
import pytest
import torch

import oracle.bdsky as obdsky
import oracle.cref as cref
import oracle.optim as oopt
import oracle.prune as oprune
import oracle.substitution as osub
import oracle.tips as otips
import oracle.tree_data as otd
from oracle import refshim
from oracle.util import TipData, order_events

pytestmark = pytest.mark.skipif(not refshim.available(), reason="/root/reference is not mounted (GPU box): goldens carry the pin")

RTOL = 1e-13


@pytest.fixture(scope="module")
def ref():
    return refshim.load_reference()


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-300))


def _random_graph(n, rng, caterpillar=False):
    """Random rooted binary tree as a DiGraph with random edge weights, inserted root-first."""
    live = [f"t{i}" for i in range(n)]
    kids, k = {}, 0
    while len(live) > 1:
        i, j = (0, 1) if caterpillar else sorted(rng.choice(len(live), size=2, replace=False))
        a, b = live[i], live[j]
        node = f"i{k}"
        k += 1
        kids[node] = (a, b)
        live = [x for idx, x in enumerate(live) if idx not in (i, j)]
        live.insert(0, node) if caterpillar else live.append(node)
    G = nx.DiGraph()
    stack = [live[0]]
    while stack:
        u = stack.pop()
        for v in kids.get(u, ()):
            G.add_edge(u, v, weight=float(rng.uniform(0.1, 1.0)))
            stack.append(v)
    return G


def _trees(ref, n, rng, caterpillar=False):
    G = _random_graph(n, rng, caterpillar)
    rtd, rmap = ref.tree_data.TreeData.from_nx(G, "weight")
    otd_, omap = otd.TreeData.from_nx(G, "weight")
    return G, rtd, rmap, otd_, omap


def _partials(n, S, rng, ambiguous=0.1):
    alphabet = list("ACGT")
    amb = list("RYMWSKBDHVN-?")
    cols = ["".join(rng.choice(amb) if rng.random() < ambiguous else rng.choice(alphabet) for _ in range(n)) for _ in range(S)]
    return np.array([osub.encode_partials(c) for c in cols])  # [S, n, 4]


# ---------------------------------------------------------------------------------------------------------------------
def test_substitution_models_and_expm_action(ref):
    rng = np.random.default_rng(0)
    for kappa in (1.0, 2.7, 0.3, 11.0):
        Qr, Qo = ref.substitution.HKY(kappa), osub.HKY(kappa)
        for f in ("pi", "P", "d", "Pinv"):
            assert np.array_equal(np.asarray(getattr(Qr, f)), np.asarray(getattr(Qo, f))), f
        assert np.array_equal(np.asarray(Qr.Q), np.asarray(Qo.Q))
        for t in (0.0, 1e-9, 1e-4, 0.03, 1.0, 40.0, 2000.0):
            for _ in range(4):
                p = rng.dirichlet(np.ones(4)) if rng.random() < 0.7 else rng.integers(0, 2, 4).astype(float)
                for right in (True, False):
                    a = np.asarray(Qr.expm_action(ref.jnp.array(t), ref.jnp.array(p), right))
                    b = Qo.expm_action(t, p, right)
                    np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-16)  # summation order of the 4x4 products differs (torch vs NumPy)
            np.testing.assert_allclose(np.asarray(Qr.expm(t)), Qo.expm(t), rtol=1e-13, atol=1e-16)
    jr, jo = ref.substitution.JC69(), osub.JC69()
    assert np.array_equal(np.asarray(jr.d), jo.d) and np.array_equal(np.asarray(jr.P), jo.P)
    # encode_partials: every BEAST code
    for ch in "ACGTURYMWSKBDHVNX-?":
        assert np.array_equal(ref.substitution.encode_partials(ch * 3), osub.encode_partials(ch * 3))


@pytest.mark.parametrize("n,seed,cat", [(2, 1, False), (3, 2, False), (7, 3, False), (16, 4, False), (24, 5, False), (12, 6, True)])
def test_tree_arrays_and_height_transform(ref, n, seed, cat):
    """from_nx numbering (bit-exact), derived arrays, height transform and its inverse."""
    from vbsky_b200.tree_data import TreeData as PTD

    rng = np.random.default_rng(seed)
    G, rtd, rmap, o, omap = _trees(ref, n, rng, cat)
    ptd, pmap = PTD.from_nx(G, "weight")
    assert rmap == omap == pmap
    for f in ("postorder", "child_parent", "parent_child"):
        assert np.array_equal(np.asarray(getattr(rtd, f)), getattr(o, f)), f
        assert np.array_equal(np.asarray(getattr(rtd, f)), getattr(ptd, f)), f
    assert np.array_equal(np.asarray(rtd.sample_times), o.sample_times)
    assert np.array_equal(np.asarray(rtd.sample_times), ptd.sample_times)
    assert (rtd.n, rtd.N) == (o.n, o.N)
    assert np.array_equal(np.asarray(rtd.preorder), o.preorder)
    assert np.array_equal(np.asarray(rtd.lower_sampling_times), o.lower_sampling_times)
    assert np.array_equal(np.asarray(rtd.lower_sampling_times), ptd.lower_sampling_times)
    if n > 2:
        # siblings: the root's entry is a scatter leftover (-1) in the reference and is never read
        assert np.array_equal(np.asarray(rtd.siblings)[:-1], np.asarray(o.siblings)[:-1])
    props = rng.uniform(0.05, 0.95, n - 2)
    rh = float(rng.uniform(0.5, 3.0))
    hr = np.asarray(rtd.height_transform(rh, ref.jnp.array(props)))
    ho = o.height_transform(rh, props)
    assert np.array_equal(hr, ho)
    assert np.array_equal(hr, ptd.height_transform(rh, props))
    if n > 2:
        rh2, p2 = rtd.inverse_height_transform(ref.jnp.array(hr))
        rh3, p3 = o.inverse_height_transform(ho)
        assert float(rh2) == float(rh3)
        np.testing.assert_allclose(np.asarray(p2), p3, rtol=1e-15, atol=0)


def test_from_newick_and_doctest_values(ref):
    """The reference's from_newick (tree_data.py:149-254) on a restated Bio.Phylo reader vs the oracle's."""
    TD = ref.tree_data.TreeData
    with ref.active():
        for nwk in ("((human:1,chimp:2):1,gorilla:2.5);", "((human,chimp),gorilla)", "(A,B,C,D)", "((a:1,b:2):3,c:4)",
                    "(((a:0.3,b:0.1):0.2,(c:0.5,d:0.25):0.1):0.4,(e:1.0,(f:0.2,g:0.3):0.6):0.05);"):
            rtd, rmap = TD.from_newick(nwk)
            o, omap = otd.TreeData.from_newick(nwk)
            assert rmap == omap
            for f in ("postorder", "child_parent", "parent_child", "sample_times"):
                assert np.array_equal(np.asarray(getattr(rtd, f)), getattr(o, f)), (nwk, f)
            assert rtd.to_newick() == o.to_newick()
        td, nm = TD.from_newick("((human:1,chimp:2):1,gorilla:2.5);")
        assert (nm["human"], nm["chimp"], nm["gorilla"]) == (0, 1, 2)  # tree_data.py:166-171
        assert np.asarray(td.sample_times).tolist() == [1.0, 0.0, 0.5]  # :172-173
        assert TD.from_newick("(a,b)")[0].lower_sampling_times.tolist() == [0.0, 0.0, 0.0]  # :88-89
        assert TD.from_newick("(a:1,b:2)")[0].lower_sampling_times.tolist() == [1.0, 0.0, 1.0]  # :93-94
        assert TD.from_newick("((a:1,b:2):3,c:4)")[0].lower_sampling_times.tolist() == [1.0, 0.0, 1.0, 1.0, 1.0]  # :97-98


def test_order_events(ref):
    rng = np.random.default_rng(2)
    ev = ref.util.order_events([0, 0.5, 1.0], [0.0, 0.0, 0.1, 0.7, 0.7, 0.9], [True, True, False, True, True, False]).tolist()
    assert len(ev) == 9 and ev[3] == [0.1, 1.0] and ev[7] == [0.9, 2.0]  # util.py:41-62
    for n in (3, 9, 30):
        heights = np.concatenate([rng.uniform(0, 1, n), rng.uniform(0, 2, n - 1)])
        heights[rng.integers(n)] = 0.0
        heights[: n // 3] = 0.0  # ties
        is_sample = np.arange(2 * n - 1) < n
        times = np.linspace(0, 2.5, 6)
        a = np.asarray(ref.util.order_events(times, heights, is_sample))
        b = np.asarray(order_events(times, heights, is_sample))
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n,S,kappa,seed,cat", [(2, 12, 2.7, 1, False), (5, 30, 1.0, 2, False), (13, 40, 2.7, 3, False),
                                                (24, 24, 5.0, 4, False), (10, 24, 0.5, 5, True)])
def test_prune_loglik_and_custom_jvp(ref, n, S, kappa, seed, cat):
    """Per-site log-likelihoods (prune.py:99-131) and the eq. (9) gradient (prune.py:134-150) -- the latter obtained the
    way the reference's callers obtain it: reverse-mode through the custom JVP, vmapped over sites (bdsky.py:331-343)."""
    jnp, jax = ref.jnp, ref.jax
    rng = np.random.default_rng(seed)
    G, rtd, _, o, _ = _trees(ref, n, rng, cat)
    part = _partials(n, S, rng)
    counts = rng.integers(0, 5, S).astype(float)
    bl = rng.uniform(1e-4, 0.4, 2 * n - 2)
    bl[rng.integers(2 * n - 2)] = 3.0
    Qr, Qo = ref.substitution.HKY(kappa), osub.HKY(kappa)
    vm = jax.vmap(ref.prune.prune_loglik, (None, None, 0, None, None, None))
    site_ref = np.asarray(vm(jnp.array(bl), Qr, part, rtd, True, False))
    site_or = oprune.prune_loglik(bl, Qo, part, o)
    np.testing.assert_allclose(site_or, site_ref, rtol=RTOL, atol=0)

    def data_term(b):  # bdsky.py:331-343
        return (vm(b, Qr, part, rtd, True, False) * counts).sum()

    v, g = jax.value_and_grad(data_term)(jnp.array(bl))
    tot, go, _ = oprune.data_loglik_and_grad(bl, Qo, part, counts, o)
    assert abs(tot - float(v)) <= RTOL * abs(float(v))
    np.testing.assert_allclose(go, np.asarray(g), rtol=1e-12, atol=1e-13 * np.abs(np.asarray(g)).max())
    # the C/OpenMP port (the timed CPU baseline) against the reference too
    masks = (part.astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(-1).astype(np.uint8)
    ll_c, g_c, site_c = cref.prune(o, bl, Qo, masks, counts, want_grad=True, want_site_ll=True)
    np.testing.assert_allclose(site_c, site_ref, rtol=1e-12, atol=0)
    np.testing.assert_allclose(g_c, np.asarray(g), rtol=1e-10, atol=1e-12 * np.abs(np.asarray(g)).max())
    # rescale=False path (prune.py:32-35)
    if n <= 13:
        a = np.asarray(jax.vmap(ref.prune.prune_loglik, (None, None, 0, None, None, None))(jnp.array(bl), Qr, part, rtd, False, False))
        np.testing.assert_allclose(oprune.prune_loglik(bl, Qo, part, o, rescale=False), a, rtol=RTOL, atol=0)


def test_prune_partials_arrays(ref):
    """The intermediate arrays of both scans (prune.py:15-96)."""
    rng = np.random.default_rng(11)
    n = 9
    G, rtd, _, o, _ = _trees(ref, n, rng)
    part = _partials(n, 1, rng)[0]
    bl = rng.uniform(1e-3, 0.5, 2 * n - 2)
    Qr, Qo = ref.substitution.HKY(2.7), osub.HKY(2.7)
    post_r, lc_r = ref.prune._compute_postorder_partials(ref.jnp.array(bl), Qr, ref.jnp.array(part), rtd, True)
    post_o, lc_o = oprune.compute_postorder_partials(bl, Qo, part, o)
    np.testing.assert_allclose(post_o, np.asarray(post_r), rtol=RTOL, atol=0)
    np.testing.assert_allclose(lc_o, np.asarray(lc_r), rtol=RTOL, atol=1e-15)
    pre_r, lp_r = ref.prune._compute_preorder_partials(ref.jnp.array(bl), Qr, post_r, lc_r, rtd, True)
    pre_o, lp_o = oprune.compute_preorder_partials(bl, Qo, post_o, lc_o, o)
    np.testing.assert_allclose(pre_o, np.asarray(pre_r), rtol=RTOL, atol=0)
    np.testing.assert_allclose(lp_o, np.asarray(lp_r), rtol=RTOL, atol=1e-15)


def _bdsky_inputs(n, m, rng, rho_sampling):
    lam = rng.lognormal(0.5, 0.4, m)
    psi = rng.uniform(0.05, 0.6, m)
    mu = rng.uniform(0.1, 1.0, m)
    rho = np.zeros(m)
    tm = 4.0
    times = np.linspace(0, tm, m + 1)
    heights = np.concatenate([rng.uniform(0, 1.5, n), rng.uniform(1.5, 3.5, n - 1)])
    heights[0] = 0.0
    if rho_sampling:
        rho[-1] = 0.3
        rho[m // 2] = 0.2
        heights[1] = 0.0  # a second tip exactly at t_m (rho-sampled)
        heights[2] = tm - times[m // 2 + 1]  # and one on an interior rho-sampling time
    xs = tm - heights
    return lam, psi, mu, rho, xs, times


@pytest.mark.parametrize("n,m,seed,rho_sampling,cond", [(4, 1, 1, False, True), (10, 5, 2, False, True), (10, 5, 3, True, True),
                                                        (25, 12, 4, False, False), (25, 12, 5, True, False)])
def test_tree_loglik_value_and_autodiff(ref, n, m, seed, rho_sampling, cond):
    """bdsky._tree_loglik (bdsky.py:18-216) and its gradient w.r.t. every argument: the reference differentiated by
    torch.autograd through the shim's primitives, the oracle by torch.autograd through its own restatement."""
    jnp, jax = ref.jnp, ref.jax
    rng = np.random.default_rng(seed)
    G, rtd, _, o, _ = _trees(ref, n, rng)
    args = _bdsky_inputs(n, m, rng, rho_sampling)

    def f(a):
        return ref.bdsky._tree_loglik(a[0], a[1], a[2], a[3], a[4], a[5], rtd, False, cond)

    v, g = jax.value_and_grad(f)([jnp.array(a) for a in args])
    ot = [torch.tensor(a, requires_grad=True) for a in args]
    vo = obdsky.tree_loglik(*ot, n, condition_on_survival=cond)
    vo.backward()
    assert np.isfinite(float(v))
    assert abs(vo.item() - float(v)) <= RTOL * abs(float(v))
    for k, (a, b) in enumerate(zip(g, ot)):
        want = np.asarray(a)
        np.testing.assert_allclose(b.grad.numpy(), want, rtol=1e-11, atol=1e-12 * max(np.abs(want).max(), 1e-300), err_msg=str(k))


def test_tree_loglik_extinction_fallback_branch(ref):
    """psi = rho = 0 on old intervals with lam < mu and long intervals drives p_i to 1 in floating point, which takes the
    logit fallback of bdsky.py:62-79."""
    jnp = ref.jnp
    rng = np.random.default_rng(7)
    n, m = 6, 4
    G, rtd, _, o, _ = _trees(ref, n, rng)
    lam = np.array([0.5, 0.4, 2.0, 2.5])
    mu = np.array([3.0, 2.5, 0.5, 0.4])
    psi = np.array([0.0, 0.0, 0.5, 0.6])
    rho = np.zeros(m)
    tm = 400.0
    times = np.linspace(0, tm, m + 1)
    heights = np.concatenate([rng.uniform(0, 50, n), rng.uniform(50, 150, n - 1)])
    heights[0] = 0.0
    xs = tm - heights
    v = ref.bdsky._tree_loglik(*[jnp.array(a) for a in (lam, psi, mu, rho, xs, times)], rtd, False, True)
    vo = obdsky.tree_loglik(lam, psi, mu, rho, xs, times, n, condition_on_survival=True)
    assert np.isfinite(float(v))
    assert abs(vo.item() - float(v)) <= RTOL * abs(float(v))


@pytest.mark.parametrize("c", [(True, True, True), (False, True, False), (False, False, True), (True, False, False)])
@pytest.mark.parametrize("equidistant", [True, False])
def test_bdsky_loglik_value_and_grad(ref, c, equidistant):
    """bdsky.loglik (bdsky.py:268-345): every term switch, both grid modes, gradients w.r.t. every parameter."""
    jnp, jax = ref.jnp, ref.jax
    rng = np.random.default_rng(21)
    n, S, m = 11, 25, 6
    G, rtd, _, o, _ = _trees(ref, n, rng)
    part = _partials(n, S, rng)
    counts = rng.integers(1, 4, S).astype(float)
    origin, origin_start = 0.9, float(np.asarray(rtd.sample_times).max()) + 0.2
    params = {"proportions": rng.uniform(0.1, 0.9, n - 2), "root_proportion": np.array([0.6]), "origin": np.array([origin]),
              "origin_start": np.array([origin_start]), "clock_rate": np.array([0.03]), "R": rng.lognormal(0, 0.3, m),
              "delta": np.full(m, 2.5), "s": np.full(m, 0.15), "rho": np.zeros(m)}
    if not equidistant:
        tm = origin + origin_start
        params["grid"] = np.concatenate([[tm], np.sort(rng.uniform(0.05 * tm, 0.95 * tm, m - 1))[::-1], [0.0]])
    Qr, Qo = ref.substitution.HKY(2.7), osub.HKY(2.7)

    def rprior(p):
        return ref.bdsky._lognorm_logpdf(jnp.log(p["R"]), 0.0, 1.0).sum() + ref.bdsky._lognorm_logpdf(jnp.log(p["origin"]), 0.3, 0.5).sum()

    def oprior(p):
        ln = lambda lx, mu_, sig: (math.log(2 * math.pi) + ((lx - mu_) / sig) ** 2) / -2 - lx - math.log(sig)
        return ln(torch.log(p["R"]), 0.0, 1.0).sum() + ln(torch.log(p["origin"]), 0.3, 0.5).sum()

    def f(p):
        return ref.bdsky.loglik(p, rtd, ref.util.TipData(part, counts), Qr, c, False, True, equidistant, rprior)

    v, g = jax.value_and_grad(f)({k: jnp.array(x) for k, x in params.items()})
    op = {k: torch.tensor(x, requires_grad=True) for k, x in params.items()}
    vo = obdsky.loglik(op, o, TipData(part, counts), Qo, c, True, equidistant, oprior)
    vo.backward()
    assert abs(vo.item() - float(v)) <= RTOL * abs(float(v))
    for k in params:
        want = np.asarray(g[k])
        got = np.zeros_like(want) if op[k].grad is None else op[k].grad.numpy()
        np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12 * max(np.abs(want).max(), 1e-300), err_msg=k)


def test_process_tips_and_pad_tips(ref):
    """fasta._process_tips (fasta.py:539-548) and SeqData.pad_tips (fasta.py:325-331), with the product's packed
    equivalents checked bit-exact against the reference's dense output."""
    from types import SimpleNamespace

    from vbsky_b200.substitution import codes_to_partials
    from vbsky_b200.tips import pad_tips as ppad
    from vbsky_b200.tips import process_tips as pproc

    rng = np.random.default_rng(3)
    outs_r, outs_o, outs_p = [], [], []
    for n, L in ((6, 80), (6, 200), (6, 30)):
        names = [f"s{i}" for i in range(n)]
        seqs = ["".join(rng.choice(list("ACGTN-RY"), L, p=[.3, .3, .15, .15, .04, .02, .02, .02])) for _ in range(n)]
        perm = rng.permutation(n)
        node_mapping = {nm: int(perm[i]) for i, nm in enumerate(names)}
        msa = [SimpleNamespace(description=nm, seq=s) for nm, s in zip(names, seqs)]
        tr, Sr = ref.fasta._process_tips(msa, node_mapping, {nm: nm for nm in names})
        to, So = otips.process_tips(names, seqs, node_mapping)
        assert Sr == So and np.array_equal(tr.partials, to.partials) and np.array_equal(tr.counts, to.counts)
        tp = pproc(names, seqs, node_mapping)
        assert np.array_equal(codes_to_partials(tp.patterns), tr.partials) and np.array_equal(tp.counts, tr.counts)
        outs_r.append(tr), outs_o.append(to), outs_p.append(tp)
    holder = SimpleNamespace(tip_data_cs=list(outs_r), max_partial_count=max(t.partials.shape[0] for t in outs_r))
    ref.fasta.SeqData.pad_tips(holder)
    for a, b, p in zip(holder.tip_data_cs, otips.pad_tips(outs_o), ppad(outs_p)):
        assert np.array_equal(a.partials, b.partials) and np.array_equal(a.counts, b.counts)
        assert np.array_equal(codes_to_partials(p.patterns), a.partials) and np.array_equal(p.counts, a.counts)


def _default_flows(ref, n, m, earliest, clock_rate):
    """The flows of SeqData.setup_flows (fasta.py:362-385) for a tree with n tips."""
    F, Tr, Const = ref.fasta, ref.prob_transform.Transform, ref.prob_distribution.Constant
    gf = ref.prob.VF(origin=Tr(1, F.pos), origin_start=Const(earliest), delta=Const(np.repeat(36.5, m)), R=Tr(m, F.pos),
                     rho_m=Const(0), s=Const(np.repeat(0.02, m)), precision=Tr(1, F.pos), clock_rate=Const(clock_rate))
    lf = {"proportions": Tr(n - 2, F.z1), "root_proportion": Tr(1, F.z1)}
    return ref.prob.VF(gf | lf)


def _oracle_flows(ref, flows):
    Const = ref.prob_distribution.Constant
    out = {}
    for k, v in flows.items():
        if isinstance(v, Const):
            out[k] = ("const", np.atleast_1d(v.c))
        else:
            out[k] = ("pos" if k in ("origin", "R", "precision") else "z1", v.dim)
    return out


def _ref_prior(ref):
    jnp, jax = ref.jnp, ref.jax

    def prior(p):  # shaped like paper/covid.py:149-175 (GMRF on log R with a gamma hyper-prior on the precision)
        ll = jax.scipy.stats.gamma.logpdf(p["precision"][0], a=0.001, scale=1 / 0.001)
        ll += ref.bdsky._lognorm_logpdf(jnp.log(p["R"]), 0.0, 1.0).sum()
        ll += ref.bdsky._lognorm_logpdf(jnp.log(p["origin"]), 1.0, 0.5).sum()
        ll -= (p["precision"][0] / 2) * (jnp.diff(jnp.log(p["R"])) ** 2).sum()
        return ll

    return prior


def _oracle_prior(p):
    ln = lambda lx, mu_, sig: (math.log(2 * math.pi) + ((lx - mu_) / sig) ** 2) / -2 - lx - math.log(sig)
    tau, a, scale = p["precision"][0], 0.001, 1 / 0.001
    ll = (a - 1) * torch.log(tau / scale) - tau / scale - math.lgamma(a) - math.log(scale)
    ll = ll + ln(torch.log(p["R"]), 0.0, 1.0).sum() + ln(torch.log(p["origin"]), 1.0, 0.5).sum()
    return ll - (tau / 2) * (torch.diff(torch.log(p["R"])) ** 2).sum()


@pytest.mark.parametrize("c", [((True, True), (True, True, True)), ((False, True), (False, True, True)), ((True, False), (True, True, True))])
def test_optim_loss_value_and_grad(ref, c):
    """optim.loss (optim.py:35-78) under value_and_grad with the reference's default flow families, against the oracle's
    restatement fed with the same base draws: the ELBO and d/d(every variational parameter)."""
    jax = ref.jax
    rng = np.random.default_rng(31)
    n, S, m = 9, 20, 5
    G, rtd, _, o, _ = _trees(ref, n, rng)
    part = _partials(n, S, rng)
    counts = rng.integers(1, 4, S).astype(float)
    flows = _default_flows(ref, n, m, float(np.asarray(rtd.sample_times).max()) + 0.3, 0.02)
    params = jax.tree_util.tree_map(lambda x: rng.normal(size=np.shape(x)) * 0.1, {k: v.params for k, v in flows.items()})
    rec = []
    ref.core.random.recorder = rec
    try:
        f, g = jax.value_and_grad(ref.optim.loss)(params, flows, rtd, ref.util.TipData(part, counts), jax.random.PRNGKey(3),
                                                  ref.substitution.HKY(2.7), c, False, True, _ref_prior(ref))
    finally:
        ref.core.random.recorder = None
    names = [k for k, v in flows.items() if not isinstance(v, ref.prob_distribution.Constant)]
    assert len(rec) == len(names)
    eps = {k: r[1].reshape(-1) for k, r in zip(names, rec)}
    op = {k: {kk: torch.tensor(np.asarray(params[k]["transform"]["t1"][kk]), requires_grad=True) for kk in ("mu", "log_sigma")}
          for k in names}
    fo = oopt.loss(op, _oracle_flows(ref, flows), o, TipData(part, counts), eps, osub.HKY(2.7), c, True, _oracle_prior)
    assert abs(fo.item() - float(f)) <= RTOL * abs(float(f))
    fo.backward()
    for k in names:
        for kk in ("mu", "log_sigma"):
            want = np.asarray(g[k]["transform"]["t1"][kk])
            got = np.zeros_like(want) if op[k][kk].grad is None else op[k][kk].grad.numpy()
            np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-12 * max(np.abs(want).max(), 1e-300), err_msg=f"{k}.{kk}")


def test_pin_is_sensitive_to_a_one_line_change(ref, tmp_path):
    """A copy of the reference with ONE arithmetic line of prune.py altered (rescaling constant dropped from the
    log-constant accumulation) no longer matches the oracle: the pin really exercises the reference's lines."""
    dst = tmp_path / "mutated"
    shutil.copytree(os.path.join(refshim.REFERENCE_ROOT, "vbsky"), dst / "vbsky")
    src = (dst / "vbsky" / "prune.py").read_text()
    needle = "const.at[parent].set(jnp.log(c) + const[children].sum())"
    assert src.count(needle) == 1
    (dst / "vbsky" / "prune.py").write_text(src.replace(needle, "const.at[parent].set(jnp.log(c) + const[children].max())"))
    mut = refshim.load_reference(str(dst), fresh=True)
    rng = np.random.default_rng(1)
    n, S = 8, 6
    G, rtd, _, o, _ = _trees(ref, n, rng)
    part = _partials(n, S, rng, ambiguous=0.0)
    bl = rng.uniform(0.01, 0.3, 2 * n - 2)
    good = np.asarray(ref.jax.vmap(ref.prune.prune_loglik, (None, None, 0, None, None, None))(
        ref.jnp.array(bl), ref.substitution.HKY(2.7), part, rtd, True, False))
    mtd = mut.tree_data.TreeData(*[np.asarray(x) for x in rtd])
    bad = np.asarray(mut.jax.vmap(mut.prune.prune_loglik, (None, None, 0, None, None, None))(
        mut.jnp.array(bl), mut.substitution.HKY(2.7), part, mtd, True, False))
    want = oprune.prune_loglik(bl, osub.HKY(2.7), part, o)
    np.testing.assert_allclose(want, good, rtol=RTOL)
    assert _rel(bad, want) > 1e-3
