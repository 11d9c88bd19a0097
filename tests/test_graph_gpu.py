"""The fused loglik entry point only enqueues work on the caller's stream (no allocation, no synchronisation), so a
whole evaluation can be captured in a CUDA graph and replayed on updated parameters -- the launch-bound regime of
small ensembles (BASELINE config 1)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_loglik_is_cuda_graph_capturable(cuda):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    T, n, S, K, m = 4, 10, 134, 3, 10
    data = make_ensemble(T, n, S, K, seed=3, Q=HKY(2.7), clock_rate=0.05, origin=1.0)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
    U = T * K
    g = torch.Generator(device="cpu").manual_seed(0)
    rnd = lambda *shape, lo=0.1, hi=0.9: (lo + (hi - lo) * torch.rand(*shape, generator=g, dtype=torch.float64)).to(cuda)
    params = dict(proportions=rnd(U, n - 2), root_proportion=rnd(U), origin=rnd(U, lo=1.0, hi=2.0),
                  origin_start=torch.ones(U, dtype=torch.float64, device=cuda), clock_rate=rnd(U, lo=0.01, hi=0.1),
                  R=rnd(U, m, lo=0.8, hi=2.0), delta=rnd(U, m, lo=1.0, hi=3.0), s=rnd(U, m, lo=0.05, hi=0.5),
                  rho=torch.zeros(U, m, dtype=torch.float64, device=cuda))
    unit_tree = torch.arange(T, dtype=torch.int32, device=cuda).repeat_interleave(K)
    grads = {k: torch.zeros_like(params[k]) for k in Ensemble.GRAD_KEYS}
    out = {k: torch.zeros(U, dtype=torch.float64, device=cuda) for k in ("ll", "tree", "data")}
    side = torch.cuda.Stream(device=cuda)
    side.wait_stream(torch.cuda.current_stream(cuda))
    with torch.cuda.stream(side):  # warm-up on a side stream (workspace allocation, occupancy cache)
        for _ in range(2):
            ens.loglik_device(params, unit_tree, m, 7, grads, out)
    torch.cuda.current_stream(cuda).wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ens.loglik_device(params, unit_tree, m, 7, grads, out)
    # new parameter values in place, replay, compare with a direct call
    params["proportions"].copy_(rnd(U, n - 2))
    params["R"].copy_(rnd(U, m, lo=0.8, hi=2.0))
    graph.replay()
    torch.cuda.synchronize()
    got = {k: v.clone() for k, v in out.items()}
    got_g = {k: v.clone() for k, v in grads.items()}
    ens.loglik_device(params, unit_tree, m, 7, grads, out)
    torch.cuda.synchronize()
    for k in out:
        assert torch.equal(got[k], out[k]) and torch.isfinite(out[k]).all()
    for k in grads:
        assert torch.equal(got_g[k], grads[k])
    # replay is cheaper than eight separate launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    for _ in range(50):
        ens.loglik_device(params, unit_tree, m, 7, grads, out)
    ev[1].record()
    for _ in range(50):
        graph.replay()
    ev[2].record()
    torch.cuda.synchronize()
    print("direct %.1f us/eval, graph replay %.1f us/eval" % (ev[0].elapsed_time(ev[1]) * 20, ev[1].elapsed_time(ev[2]) * 20))
    ens.close()
