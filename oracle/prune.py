"""Oracle restatement of /root/reference/vbsky/prune.py (test infrastructure).

The reference evaluates one site pattern per call and is vmapped over patterns
(bdsky.py:333).  Here the node loops are Python loops and the site axis is a leading NumPy
batch axis, so ``tip_partials`` is [S, n, 4] (or [n, 4] for a single site).  Arithmetic per
node follows prune.py line by line, including the order of operations and the clips inside
``expm_action``.
"""
import numpy as np

from .substitution import SubstitutionModel
from .tree_data import TreeData


def compute_postorder_partials(branch_lengths, Q: SubstitutionModel, tip_partials, td: TreeData, rescale=True):
    """prune.py:15-49.  Returns post [..., 2n-1, 4], log-consts [..., 2n-1]."""
    tip_partials = np.asarray(tip_partials, dtype=np.float64)
    assert tip_partials.shape[-2] == td.n
    lead = tip_partials.shape[:-2]
    A = tip_partials.shape[-1]
    post = np.concatenate([tip_partials, np.zeros(lead + (td.n - 1, A))], axis=-2)
    const = np.zeros(lead + (2 * td.n - 1,))
    for parent in range(td.n, 2 * td.n - 1):
        children = td.parent_child[parent]
        t = branch_lengths[children]
        Pprod = Q.expm_action(t[0], post[..., children[0], :]) * Q.expm_action(t[1], post[..., children[1], :])
        if rescale:
            c = Pprod.sum(-1)
        else:
            c = np.ones(lead)
        Pprod = Pprod / c[..., None]
        post[..., parent, :] = Pprod
        const[..., parent] = np.log(c) + const[..., children].sum(-1)
    return post, const


def compute_preorder_partials(branch_lengths, Q: SubstitutionModel, postorder_partials, log_post_const, td: TreeData, rescale=True):
    """prune.py:52-96.  Returns pre [..., 2n-1, 4], log-consts [..., 2n-1]."""
    assert postorder_partials.shape[-2] == 2 * td.n - 1
    lead = postorder_partials.shape[:-2]
    M, A = postorder_partials.shape[-2:]
    pre = np.zeros(lead + (M, A))
    pre[..., -1, :] = Q.pi
    const = np.zeros(lead + (M,))
    sibs = td.siblings
    for child in td.postorder[:-1][::-1]:
        parent = td.child_parent[child]
        sib = sibs[child]
        A0 = Q.expm_action(branch_lengths[sib], postorder_partials[..., sib, :])
        Am = pre[..., parent, :] * A0
        B = Q.expm_action(branch_lengths[child], Am, right=False)
        if rescale:
            c = B.sum(-1)
        else:
            c = np.ones(lead)
        B = B / c[..., None]
        pre[..., child, :] = B
        const[..., child] = np.log(c) + const[..., parent] + log_post_const[..., sib]
    return pre, const


def prune_loglik(branch_lengths, Q: SubstitutionModel, tip_partials, td: TreeData, rescale=True):
    """prune.py:99-131.  Per-site log-likelihood(s)."""
    tip_partials = np.asarray(tip_partials, dtype=np.float64)
    branch_lengths = np.asarray(branch_lengths, dtype=np.float64)
    assert Q.P.ndim == 2 and Q.P.shape[0] == Q.P.shape[1]
    assert tip_partials.shape[-2:] == (td.n, Q.P.shape[0])
    assert branch_lengths.ndim == 1 and len(branch_lengths) == 2 * (td.n - 1)
    pop, log_consts = compute_postorder_partials(branch_lengths, Q, tip_partials, td, rescale)
    return np.log(pop[..., -1, :] @ Q.pi) + log_consts[..., -1]


def prune_loglik_and_grad(branch_lengths, Q: SubstitutionModel, tip_partials, td: TreeData, rescale=True):
    """prune.py:134-150 (the custom JVP): returns (log_P [...], grad [..., 2n-2]) with
    grad[i] = d log_P / d branch_lengths[i], eq. (9)."""
    tip_partials = np.asarray(tip_partials, dtype=np.float64)
    branch_lengths = np.asarray(branch_lengths, dtype=np.float64)
    post, log_post_const = compute_postorder_partials(branch_lengths, Q, tip_partials, td, rescale)
    P = post[..., -1, :] @ Q.pi
    log_P = np.log(P) + log_post_const[..., -1]
    pre, log_pre_const = compute_preorder_partials(branch_lengths, Q, post, log_post_const, td, rescale)
    grad = np.einsum("...nu,uv,...nv->...n", pre[..., :-1, :], Q.Q, post[..., :-1, :]) * np.exp(
        log_pre_const[..., :-1] + log_post_const[..., :-1] - log_P[..., None]
    )
    return log_P, grad


def data_loglik_and_grad(branch_lengths, Q, partials, counts, td, chunk=4096):
    """bdsky.py:331-343 data term and its branch-length gradient:
    sum_s counts[s] * prune_loglik(bl, Q, partials[s], td), d/d bl (the chain through
    ``clock_rate`` is the caller's).  Chunked over sites to bound memory."""
    S = partials.shape[0]
    tot = 0.0
    g = np.zeros(2 * td.n - 2)
    site_ll = np.empty(S)
    for s0 in range(0, S, chunk):
        ll, gr = prune_loglik_and_grad(branch_lengths, Q, partials[s0 : s0 + chunk], td)
        site_ll[s0 : s0 + chunk] = ll
        w = counts[s0 : s0 + chunk]
        tot += (ll * w).sum()
        g += (gr * w[:, None]).sum(0)
    return tot, g, site_ll


def brute_force_loglik(branch_lengths, Q: SubstitutionModel, tip_partials, td: TreeData):
    """Independent check (not in the reference): sum over all internal-state assignments of
    pi[root] * prod_edges expm(t)[parent,child] * prod_tips partial.  Exponential; tiny trees only."""
    n, N = td.n, td.N
    Pm = [Q.expm(t) for t in branch_lengths]
    tot = 0.0
    import itertools

    for states in itertools.product(range(4), repeat=n - 1):
        st = {n + i: s for i, s in enumerate(states)}
        pr = Q.pi[st[N - 1]]
        for c in range(N - 1):
            pa = td.child_parent[c]
            if c < n:
                pr *= (Pm[c][st[pa]] * tip_partials[c]).sum()
            else:
                pr *= Pm[c][st[pa], st[c]]
        tot += pr
    return np.log(tot)
