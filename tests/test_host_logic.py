"""Host-side logic of the product vs the oracle (CPU only): node numbering, derived arrays, HKY
spectrum, tip encoding, pattern compression/padding (bit-exact), and the kernel's evaluation
schedules, which are interpreted here in NumPy with the oracle's arithmetic."""
import ctypes as C

import networkx as nx
import numpy as np
import pytest

import oracle.prune as oprune
import oracle.substitution as osub
import oracle.tips as otips
import oracle.tree_data as otd
from vbsky_b200 import _lib
from vbsky_b200.substitution import HKY, JC69, codes_to_partials, encode_codes, encode_partials, partials_to_codes
from vbsky_b200.synth import evolve_codes, random_tree, sample_heights
from vbsky_b200.tips import PackedTips, compress_patterns, pad_tips, process_tips
from vbsky_b200.tree_data import TreeData


def _o(td):
    return otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)


def test_hky_matches_oracle():
    for k in (1.0, 2.7, 0.3, 11.0):
        a, b = HKY(k), osub.HKY(k)
        np.testing.assert_allclose(np.sort(a.d), np.sort(b.d), atol=1e-15)
        np.testing.assert_allclose(a.Q, b.Q, atol=1e-15)
        np.testing.assert_array_equal(a.pi, b.pi)
        for t in (0.0, 1e-6, 1e-3, 0.3, 5.0):
            np.testing.assert_allclose(a.expm(t), b.expm(t), atol=2e-15)
    np.testing.assert_allclose(JC69().Q, osub.JC69().Q, atol=1e-15)


def test_from_nx_numbering_bit_exact():
    rng = np.random.default_rng(0)
    for _ in range(40):
        n = int(rng.integers(2, 40))
        td = random_tree(n, rng)
        edges = [(int(td.child_parent[c]), c) for c in range(td.N - 1)]
        G = nx.DiGraph()
        for i in rng.permutation(len(edges)):
            G.add_edge(f"x{edges[i][0]}", f"x{edges[i][1]}", branch_length=float(rng.uniform(0.1, 2)))
        a, ma = TreeData.from_nx(G)
        b, mb = otd.TreeData.from_nx(G)
        assert ma == mb
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
        assert np.array_equal(a.siblings, b.siblings)
        assert np.array_equal(a.lower_sampling_times, b.lower_sampling_times)
        assert (a.child_parent[:-1] > np.arange(a.N - 1)).all() and a.child_parent[-1] == -1
        if n > 2:
            pr = rng.uniform(0.05, 0.95, n - 2)
            h = a.height_transform(1.3, pr)
            assert np.array_equal(h, b.height_transform(1.3, pr))
            rh, pr2 = a.inverse_height_transform(h)
            rh_o, pr_o = b.inverse_height_transform(h)
            np.testing.assert_array_equal(pr2, pr_o)
            assert rh == rh_o


def test_encode_and_ambiguity_table():
    chars = "ACGTURYMWSKBDHVNX-?"
    for ch in chars:
        assert np.array_equal(encode_partials(ch), osub.encode_partials(ch))
    with pytest.raises(KeyError):
        encode_codes(["ACgT"])
    with pytest.raises(KeyError):
        osub.encode_partials("ACgT")
    codes = encode_codes(["ACGT", "NNRY"])
    assert codes.shape == (4, 2) and codes.dtype == np.uint8
    assert np.array_equal(partials_to_codes(codes_to_partials(codes)), codes)


def test_pattern_compression_bit_exact():
    rng = np.random.default_rng(1)
    for _ in range(25):
        L, n = int(rng.integers(1, 500)), int(rng.integers(1, 9))
        seqs = ["".join(rng.choice(list("ACGTN-RYK"), L, p=[.3, .3, .15, .15, .02, .02, .02, .02, .02])) for _ in range(n)]
        names = [f"s{i}" for i in range(n)]
        mapping = {nm: int(i) for nm, i in zip(names, rng.permutation(n))}
        got = process_tips(names, seqs, mapping)
        want, S = otips.process_tips(names, seqs, mapping)
        assert got.patterns.shape[0] == S
        assert np.array_equal(codes_to_partials(got.patterns), want.partials)
        assert np.array_equal(got.counts, want.counts)


def test_pad_tips_matches_reference_rule():
    rng = np.random.default_rng(2)
    packed, oracle_in = [], []
    for _ in range(4):
        L, n = int(rng.integers(5, 60)), 5
        codes = rng.choice(np.array([1, 2, 4, 8, 15], dtype=np.uint8), (L, n))
        pat, cnt = compress_patterns(codes)
        packed.append(PackedTips(pat, cnt.astype(np.float64)))
        oracle_in.append(otips.TipData(codes_to_partials(pat), cnt.astype(np.float64)))
    got, want = pad_tips(packed), otips.pad_tips(oracle_in)
    for g, w in zip(got, want):
        assert np.array_equal(codes_to_partials(g.patterns), w.partials) and np.array_equal(g.counts, w.counts)


def test_hcv_example_patterns():
    """example/hcv.nexus (63 taxa x 411 sites, ambiguity codes): 246 unique patterns -- fixture generated
    from the reference's example data by tests/golden/make_golden.py."""
    import json
    import os

    gold = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "hcv_patterns.json")))
    seqs = gold["seqs"]
    codes = encode_codes(seqs)
    pat, cnt = compress_patterns(codes)
    assert pat.shape == (gold["n_patterns"], len(seqs)) == (246, 63)
    assert cnt.tolist() == gold["counts"]
    assert [int(x) for x in pat[:, 0]] == gold["first_column_codes"]


# ------------------------------------------------------------------------------- schedules
def _layout(tds):
    T, n = len(tds), tds[0].n
    pc = np.ascontiguousarray(np.stack([td.parent_child for td in tds]), dtype=np.int32)
    cp = np.ascontiguousarray(np.stack([td.child_parent for td in tds]), dtype=np.int32)
    po = np.ascontiguousarray(np.stack([td.postorder for td in tds]), dtype=np.int32)
    st = np.ascontiguousarray(np.stack([td.sample_times for td in tds]), dtype=np.float64)
    h = C.c_void_p()
    _lib.check(_lib.lib.vbsky_layout_create(T, n, _lib.np_ptr(pc, C.c_int32), _lib.np_ptr(cp, C.c_int32),
                                            _lib.np_ptr(po, C.c_int32), _lib.np_ptr(st, C.c_double), C.byref(h)))
    return h


def _schedule(h, t, n):
    up = np.zeros((n - 1, 4), dtype=np.int32)
    down = np.zeros((n - 1, 4), dtype=np.int32)
    dn = np.zeros(2 * n - 2, dtype=np.int32)
    lst = np.zeros(2 * n - 1)
    sib = np.zeros(2 * n - 1, dtype=np.int32)
    _lib.check(_lib.lib.vbsky_layout_get(h, t, _lib.np_ptr(up, C.c_int32), _lib.np_ptr(down, C.c_int32),
                                          _lib.np_ptr(dn, C.c_int32), _lib.np_ptr(lst, C.c_double), _lib.np_ptr(sib, C.c_int32)))
    return up, down, dn, lst, sib


def _interpret(up, down, dn, n, bl, Q, partials, depth):
    """Walks the two schedules exactly as prune_kernel does (on-chip slots, scratch rows, scale-free
    eq. 9) with the oracle's expm_action; returns per-site ll and d ll / d bl."""
    S = partials.shape[0]
    root = 2 * n - 2
    stack = np.full((depth, S, 4), np.nan)
    scratch = np.full((n - 1, S, 4), np.nan)
    logscale = np.zeros(S)
    ll = None
    carry = np.full((S, 4), np.nan)
    for step, (x, par, z, w) in enumerate(up):
        a, b = x & 0xFFFF, (x >> 16) & 0xFFFF
        assert bool(w & 1) == (a < n) and bool(w & 2) == (b < n) and bool(w & 4) == (par == root)
        ma = Q.expm_action(bl[a], partials[:, a]) if w & 1 else (carry.copy() if w & 8 else stack[z & 0xFF].copy())
        assert not (w & 16), "only child a is ever the carried result"
        mb = Q.expm_action(bl[b], partials[:, b]) if w & 2 else stack[(z >> 8) & 0xFF].copy()
        assert not np.isnan(ma).any() and not np.isnan(mb).any()
        pr = ma * mb
        if w & 4:
            ll = np.log(pr @ Q.pi) + logscale
        else:
            c = pr.sum(1)
            logscale += np.log(c)
            m = Q.expm_action(bl[par], pr / c[:, None])
            carry = m
            if w & 32:
                stack[(z >> 16) & 0xFF] = m
            is_cherry = a < n and b < n
            assert bool(w & 0x40) == is_cherry
            if not is_cherry:
                scratch[step] = m  # a cherry's message never travels through scratch: the down pass rebuilds it
    up_row = {int(par): i for i, (_, par, _, _) in enumerate(up)}
    g = np.zeros((S, 2 * n - 2))
    Qm = Q.Q
    carry = np.tile(Q.pi, (S, 1))  # pre[root] = pi seeds the carry; the root step is flagged "carried"
    for step, (x, par, z, w) in enumerate(down):
        a, b = x & 0xFFFF, (x >> 16) & 0xFFFF
        assert dn[2 * step] == a and dn[2 * step + 1] == b
        sp = z & 0xFF
        assert (w & 8) or sp != 0xFF
        assert (w & 3) != 1  # canonical order: a lone internal child is always child a (the kernel relies on it)
        pre = carry.copy() if w & 8 else stack[sp].copy()
        assert not np.isnan(pre).any()
        carry = np.full((S, 4), np.nan)
        def rebuilt(child, rec):
            # the cherry's own down-record follows: both children tips; message = expm_action(bl[c], normalised product)
            (x2, par2, _, w2) = down[rec]
            assert par2 == child and (w2 & 3) == 3
            a2, b2 = x2 & 0xFFFF, (x2 >> 16) & 0xFFFF
            pr = Q.expm_action(bl[a2], partials[:, a2]) * Q.expm_action(bl[b2], partials[:, b2])
            return Q.expm_action(bl[child], pr / pr.sum(1, keepdims=True))

        if w & 0x80:
            assert w & 0x40, "b is rebuilt only when a is"
        ma = Q.expm_action(bl[a], partials[:, a]) if w & 1 else (rebuilt(a, step + 1) if w & 0x40 else scratch[up_row[a]])
        mb = Q.expm_action(bl[b], partials[:, b]) if w & 2 else (rebuilt(b, step + 2) if w & 0x80 else scratch[up_row[b]])
        assert not np.isnan(ma).any() and not np.isnan(mb).any()
        # Q m_c: for tips the kernel uses (Q P(t)) e_mask with the unclipped message
        qa = (Q.expm(bl[a]) @ partials[:, a].T).T @ Qm.T if w & 1 else ma @ Qm.T
        qb = (Q.expm(bl[b]) @ partials[:, b].T).T @ Qm.T if w & 2 else mb @ Qm.T
        wa, wb = pre * mb, pre * ma
        den = (wa * ma).sum(1)
        g[:, a] = (wa * qa).sum(1) / den
        g[:, b] = (wb * qb).sum(1) / den
        if not (w & 1):
            B = Q.expm_action(bl[a], wa, right=False)
            carry = B / B.sum(1, keepdims=True)  # an internal a is expanded by the next step
            assert step + 1 < len(down) and down[step + 1][1] == a
        if not (w & 2):
            B = Q.expm_action(bl[b], wb, right=False)
            stack[(z >> 16) & 0xFF] = B / B.sum(1, keepdims=True)  # an internal b is parked
    return ll, g


def _check_chunks(h, t, n, up, down):
    """Ring-stage chunking of the down pass: chunks of <= 3 consecutive steps covering 0..n-2 that never split a step
    from the cherry records it rebuilds from, and the message list = scratch rows of the internal non-cherry children."""
    info = (C.c_int32 * 8)()
    _lib.check(_lib.lib.vbsky_layout_info(h, info))
    cap = info[5]
    nch, nmsg = C.c_int32(), C.c_int32()
    cstart = np.zeros(cap + 1, dtype=np.int32)
    mstart = np.zeros(cap + 1, dtype=np.int32)
    mrows = np.zeros(max(n - 2, 1), dtype=np.int32)
    _lib.check(_lib.lib.vbsky_layout_get_chunks(h, t, C.byref(nch), _lib.np_ptr(cstart, C.c_int32), _lib.np_ptr(mstart, C.c_int32),
                                                 C.byref(nmsg), _lib.np_ptr(mrows, C.c_int32)))
    nch, nmsg = nch.value, nmsg.value
    assert cstart[0] == 0 and cstart[nch] == n - 1 and mstart[0] == 0 and mstart[nch] == nmsg
    sizes = np.diff(cstart[: nch + 1])
    assert (sizes >= 1).all() and (sizes <= 3).all()
    chunk_of = np.repeat(np.arange(nch), sizes)
    up_row = {int(par): i for i, (_, par, _, _) in enumerate(up)}
    want_rows = []
    n_cherries = 0
    for step, (x, par, z, w) in enumerate(down):
        a, b = x & 0xFFFF, (x >> 16) & 0xFFFF
        if w & 0x40:
            assert chunk_of[step + 1] == chunk_of[step]
            n_cherries += 1
        elif not (w & 1):
            want_rows.append(up_row[a])
        if w & 0x80:
            assert chunk_of[step + 2] == chunk_of[step]
            n_cherries += 1
        elif not (w & 2):
            want_rows.append(up_row[b])
    assert mrows[:nmsg].tolist() == want_rows
    assert nmsg == (n - 2) - n_cherries
    # per-chunk message counts match the prefix array
    per_chunk = np.zeros(nch, dtype=int)
    for step, (x, par, z, w) in enumerate(down):
        per_chunk[chunk_of[step]] += int(not (w & 1) and not (w & 0x40)) + int(not (w & 2) and not (w & 0x80))
    assert np.array_equal(np.diff(mstart[: nch + 1]), per_chunk)


@pytest.mark.parametrize("n,caterpillar", [(2, False), (3, False), (7, False), (33, False), (64, True), (120, False)])
def test_schedules_reproduce_oracle(n, caterpillar):
    rng = np.random.default_rng(n)
    tds = [random_tree(n, rng, caterpillar=caterpillar) for _ in range(2)]
    h = _layout(tds)
    info = (C.c_int32 * 8)()
    _lib.check(_lib.lib.vbsky_layout_info(h, info))
    depth = max(info[2], info[3])
    assert depth <= int(np.ceil(np.log2(n))) + 1
    Q = osub.HKY(2.7)
    for t, td in enumerate(tds):
        up, down, dn, lst, sib = _schedule(h, t, n)
        assert np.array_equal(lst, td.lower_sampling_times)
        assert np.array_equal(sib[:-1], td.siblings[:-1])
        assert sorted(dn.tolist()) == list(range(2 * n - 2))  # every branch gets exactly one gradient slot
        sh = sample_heights(td, 1, rng, origin=2.0)
        bl = sh.branch_lengths[0] * 0.2
        codes = evolve_codes(td, bl, Q, 40, rng, missing=0.1)
        partials = codes_to_partials(codes)
        ll, g = _interpret(up, down, dn, n, bl, Q, partials, depth)
        _check_chunks(h, t, n, up, down)
        ll_o, g_o = oprune.prune_loglik_and_grad(bl, Q, partials, _o(td))
        np.testing.assert_allclose(ll, ll_o, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(g, g_o, rtol=1e-9, atol=1e-9 * np.abs(g_o).max())
    _lib.lib.vbsky_layout_destroy(h)


def test_layout_rejects_malformed_trees():
    td = random_tree(6, np.random.default_rng(0))
    bad_cp = td.child_parent.copy()
    bad_cp[-1] = 3
    with pytest.raises(_lib.VbskyError):
        _layout([TreeData(td.postorder, bad_cp, td.parent_child, td.sample_times)])
    bad_pc = td.parent_child.copy()
    bad_pc[td.n] = [td.N - 1, 0]  # child index larger than the parent
    with pytest.raises(_lib.VbskyError):
        _layout([TreeData(td.postorder, td.child_parent, bad_pc, td.sample_times)])
    with pytest.raises(_lib.VbskyError):
        _layout([TreeData(td.postorder[::-1].copy(), td.child_parent, td.parent_child, td.sample_times)])


def test_lognorm_logpdf_and_transform_match_scipy():
    import torch
    from scipy.stats import norm

    from vbsky_b200.bdsky import bdsky_transform, lognorm_logpdf

    x = np.array([-1.3, 0.0, 0.4, 2.2])
    got = lognorm_logpdf(torch.as_tensor(x), 0.3, 0.5).numpy()
    want = norm.logpdf((x - 0.3) / 0.5) - x - np.log(0.5)  # bdsky.py:230-231
    assert np.allclose(got, want, rtol=1e-14, atol=1e-14)
    p = {k: torch.as_tensor(v) for k, v in dict(R=[1.5, 2.0], delta=[3.0, 4.0], s=[0.1, 0.2], rho=[0.0, 0.3]).items()}
    lam, psi, mu, rho = bdsky_transform(p)
    assert np.allclose(lam, [4.5, 8.0]) and np.allclose(psi, [0.3, 0.8]) and np.allclose(mu, [2.7, 3.2]) and np.allclose(rho, [0.0, 0.3])


def test_merges_to_tree_numbering_matches_newick_route():
    """UPGMA merge list -> TreeData must number nodes exactly as the reference's write-Newick / from_newick round trip."""
    import oracle.treebuild as otb
    from vbsky_b200.treebuild import merges_to_newick, merges_to_tree

    rng = np.random.default_rng(4)
    n = 15
    x = rng.random((n, 2))
    D = np.linalg.norm(x[:, None] - x[None], axis=-1)
    merges, heights, nested = otb.upgma(D, [f"t{i}" for i in range(n)])
    names = [f"t{i}" for i in range(n)]
    td, nm = merges_to_tree(merges, names, {s: float(i) for i, s in enumerate(names)})
    want, want_map = otd.TreeData.from_newick(otb.nested_to_newick(nested))
    assert np.array_equal(td.child_parent, want.child_parent) and np.array_equal(td.parent_child, want.parent_child)
    assert np.array_equal(td.postorder, want.postorder)
    assert all(nm[s] == want_map[s] for s in names)
    assert all(td.sample_times[nm[s]] == float(i) for i, s in enumerate(names))
    td2, map2 = TreeData.from_newick(merges_to_newick(merges, heights, names))
    assert np.array_equal(td2.child_parent, td.child_parent) and all(map2[s] == nm[s] for s in names)


def test_reference_arm_never_maps_the_product_library():
    """bench.py --impl reference (the CPU arm) must run without importing vbsky_b200: its JSON line lists the in-tree shared
    objects mapped into the process, and only the oracle's C port may be there."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "tiny", "--steps", "1",
                          "--warmup", "1", "--ref-units-per-step", "2"], capture_output=True, text=True, timeout=600,
                         env={**os.environ, "OMP_NUM_THREADS": "1"})  # as under torchrun: the arm must set its own thread count
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["impl"] == "reference" and line["native_so_loaded"] == ["oracle/libvbsky_oracle.so"]
    assert line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)  # not the launcher's OMP_NUM_THREADS=1
    assert line["config"]["workload"].startswith("tiny:") and line["e2e"]["h2d_bytes_per_step"] == 0
