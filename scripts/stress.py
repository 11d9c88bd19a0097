"""Randomised parity sweep of the pruning kernel (developer tool): python scripts/stress.py [n_cases] [seed] [mode]
mode "gpu" (default): CUDA kernel vs the NumPy restatement of prune.py (oracle/prune.py), on the GPU box;
mode "cpu": C/OpenMP port vs the NumPy restatement (no GPU needed).  Every case is seeded by (seed, case) alone.
Random tree sizes, pattern counts, branch-length scales (down to exact zeros -> literal-clip variant), ambiguity
fractions, kappa, compressed and raw patterns, caterpillars."""
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle.cref as cref  # noqa: E402
import oracle.prune as oprune  # noqa: E402
import oracle.substitution as osub  # noqa: E402
import oracle.tree_data as otd  # noqa: E402
from vbsky_b200 import synth  # noqa: E402
from vbsky_b200.substitution import HKY, codes_to_partials  # noqa: E402
from vbsky_b200.tips import PackedTips, compress_patterns, pad_tips  # noqa: E402

AMBIG = np.array([3, 5, 6, 7, 9, 10, 11, 12, 13, 14, 15], dtype=np.uint8)


def gen_case(seed, case):
    rng = np.random.default_rng([seed, case])
    T = int(rng.integers(1, 4))
    n = int(rng.choice([2, 3, 4, 5, 8, 17, 33, 64, 100, 200, 301]))
    S = int(rng.choice([1, 2, 63, 64, 65, 255, 256, 257, 700, 1500]))
    K = int(rng.integers(1, 4))
    kappa = float(rng.choice([1.0, 2.7, 0.3, 11.0]))
    scale = float(10.0 ** rng.uniform(-7, 0.5))
    p_amb = float(rng.choice([0.0, 0.01, 0.3]))
    p_zero = float(rng.choice([0.0, 0.0, 0.05]))
    cat = bool(rng.random() < 0.15)
    comp = bool(rng.random() < 0.3)
    Q = HKY(kappa)
    tds, tips, bls = [], [], []
    for t in range(T):
        td = synth.random_tree(n, rng, caterpillar=cat)
        sh = synth.sample_heights(td, K, rng, origin=0.3)
        bl = sh.branch_lengths * scale
        bl[rng.random(bl.shape) < p_zero] = 0.0
        codes = synth.evolve_codes(td, np.maximum(bl[0], 1e-3), Q, S, rng, missing=0.0)
        amb = rng.random(codes.shape) < p_amb
        codes[amb] = rng.choice(AMBIG, int(amb.sum()))
        if comp:
            pat, cnt = compress_patterns(codes)
            tp = PackedTips(pat, cnt.astype(np.float64))
        else:
            tp = PackedTips(codes, rng.integers(0, 4, S).astype(np.float64))
        tds.append(td), tips.append(tp), bls.append(bl)
    desc = f"T={T} n={n} S={S} K={K} kappa={kappa} scale={scale:.3g} amb={p_amb} zero={p_zero} cat={cat} comp={comp}"
    return tds, pad_tips(tips), np.concatenate(bls), np.repeat(np.arange(T, dtype=np.int32), K), kappa, desc


def errors(ll, dll, site, want, g, sll):
    e_ll = np.max(np.abs(site - sll) / (np.abs(sll) + 1e-4))
    e_tot = abs(ll - want) / max(abs(want), 1e-4)
    # floor: a gradient that is exactly zero in exact arithmetic (all counts zero, or two fully ambiguous tips: Q 1 = 0)
    # comes out as 1e-20 ... 1e-15 rounding residue on either side; errors are judged against at least 1e-8
    e_g = np.max(np.abs(dll - g)) / max(np.abs(g).max(), 1e-8)
    return max(e_ll, e_tot), e_g


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    mode = sys.argv[3] if len(sys.argv) > 3 else "gpu"
    only = int(sys.argv[4]) if len(sys.argv) > 4 else None
    worst = {"A": [0.0, 0.0], "Z": [0.0, 0.0], "X": [0.0, 0.0]}
    bad = {"A": 0, "Z": 0, "X": 0}
    units = {"A": 0, "Z": 0, "X": 0}
    for case in range(cases) if only is None else [only]:
        tds, tips, bl, ut, kappa, desc = gen_case(seed, case)
        oQ = osub.HKY(kappa)
        if mode == "gpu":
            import torch

            from vbsky_b200.ensemble import Ensemble

            dev = torch.device("cuda", 0)
            ens = Ensemble(tds, tips, HKY(kappa), device=dev)
            ll, dll, site = ens.prune(torch.as_tensor(bl, device=dev), torch.as_tensor(ut, device=dev), want_grad=True, want_site_ll=True)
            ll, dll, site = ll.cpu().numpy(), dll.cpu().numpy(), site.cpu().numpy()
            ens.close()
        for u in range(len(ut)):
            t = ut[u]
            td = otd.TreeData(tds[t].postorder, tds[t].child_parent, tds[t].parent_child, tds[t].sample_times)
            want, g, sll = oprune.data_loglik_and_grad(bl[u], oQ, codes_to_partials(tips[t].patterns), tips[t].counts, td)
            if mode == "gpu":
                got = (ll[u], dll[u], site[u])
            else:
                got = cref.prune(td, bl[u], oQ, tips[t].patterns, tips[t].counts, want_grad=True, want_site_ll=True)
            e_v, e_g = errors(*got, want, g, sll)
            # A: every branch >= 1e-7 (the reference's P @ (exp(td) * (Pinv @ p)) is accurate to ~1e-16 / t there);
            # Z: exact zeros (literal clips bind) next to otherwise well-conditioned branches; X: tiny or negative
            # branch lengths, where the reference's own result is rounding noise and parity is not defined.
            pos = bl[u][bl[u] != 0]
            cls = "X" if (pos.size and pos.min() < 1e-7) else ("Z" if (bl[u] == 0).any() else "A")
            units[cls] += 1
            worst[cls] = [max(worst[cls][0], e_v), max(worst[cls][1], e_g)]
            if not (e_v < 1e-9 and e_g < 1e-6):
                bad[cls] += 1
                if cls != "X":
                    print(f"MISMATCH[{cls}] case {case} unit {u}: {desc}: value {e_v:.3g} grad {e_g:.3g}")
    for cls in "AZX":
        print(f"class {cls}: {units[cls]} units, {bad[cls]} out of tolerance; worst relative error value {worst[cls][0]:.3g}, grad {worst[cls][1]:.3g}")


if __name__ == "__main__":
    main()
