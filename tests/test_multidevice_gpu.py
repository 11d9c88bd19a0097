"""Handles live on the device that was current at creation: results are identical on every device, and a compute
entry point called with another device current is refused with VBSKY_EINVAL instead of touching foreign memory."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(dev, data):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY

    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=dev)
    bl = torch.as_tensor(np.stack([data.heights[t].branch_lengths[k] for t in range(2) for k in range(2)]) * 0.05, device=dev)
    ut = torch.as_tensor(np.repeat(np.arange(2, dtype=np.int32), 2), device=dev)
    ll, dll, _ = ens.prune(bl, ut)
    return ens, bl, ut, ll, dll


def test_wrong_current_device_is_refused_and_devices_agree(cuda):
    from vbsky_b200 import _lib
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    data = make_ensemble(2, 12, 300, 2, seed=1, Q=HKY(2.7), clock_rate=0.05)
    ens0, _, _, ll0, dll0 = _run(torch.device("cuda", 0), data)
    ens1, bl, ut, ll1, dll1 = _run(torch.device("cuda", 1), data)  # the current device stays 0 outside the calls
    assert torch.equal(ll0.cpu(), ll1.cpu()) and torch.equal(dll0.cpu(), dll1.cpu())
    ws = torch.empty(_lib.lib.vbsky_prune_workspace_bytes(ens1._layout, ens1._tips, 4, 1), dtype=torch.uint8, device=bl.device)
    rc = _lib.lib.vbsky_prune_value_and_grad(
        ens1._layout, ens1._tips, 4, C.c_void_p(ut.data_ptr()), C.c_void_p(bl.data_ptr()), C.byref(ens1._model), None,
        C.c_void_p(ll1.data_ptr()), C.c_void_p(dll1.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), None)
    assert rc == -1 and "device" in _lib.last_error()
    ens0.close(), ens1.close()


def test_concurrent_host_threads_and_streams(cuda):
    """Handles are immutable and the library keeps no shared mutable state: two host threads, each with its own
    stream, workspace and ensemble, get the results of serial execution."""
    import threading

    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    datas = [make_ensemble(2, 40, 3000, 2, seed=s, Q=HKY(2.7), clock_rate=0.02) for s in (1, 2)]
    inputs, serial = [], []
    for data in datas:
        bl = torch.as_tensor(np.stack([data.heights[t].branch_lengths[k] for t in range(2) for k in range(2)]) * 0.02, device=cuda)
        ut = torch.as_tensor(np.repeat(np.arange(2, dtype=np.int32), 2), device=cuda)
        ens = Ensemble(data.tds, data.tips, HKY(2.7), device=cuda)
        ll, dll, _ = ens.prune(bl, ut)
        inputs.append((ens, bl, ut))
        serial.append((ll.clone(), dll.clone()))
    torch.cuda.synchronize()
    results = [None, None]

    def work(i):
        ens, bl, ut = inputs[i]
        st = torch.cuda.Stream(device=cuda)
        with torch.cuda.stream(st):
            for _ in range(20):
                ll, dll, _ = ens.prune(bl, ut)
            st.synchronize()
        results[i] = (ll, dll)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    for i in range(2):
        assert torch.equal(results[i][0], serial[i][0]) and torch.equal(results[i][1], serial[i][1])
        inputs[i][0].close()


def test_reduce_elbo_kernel_matches_reference_packing(cuda):
    """vbsky_reduce_elbo (the all-reduce payload built by one kernel) vs the plain-torch packing of
    tests/test_parallel_gloo.py, for a full unit list and for round-robin shards of it."""
    from test_parallel_gloo import pack_reference
    from vbsky_b200.parallel import payload_count, reduce_elbo, shard_units

    gen = torch.Generator(device=cuda)
    gen.manual_seed(0)
    for (T, K, n, m) in ((7, 3, 6, 5), (100, 10, 200, 50), (5, 1, 2, 1), (3, 4, 9, 0)):
        for world in (1, 2, 8):
            for rank in sorted({0, world - 1}):
                units = shard_units(T, K, rank, world)
                if not units:
                    continue
                U = len(units)
                rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=cuda, generator=gen)
                ut = torch.tensor([t for t, _ in units], dtype=torch.int32, device=cuda)
                ll = rnd(U)
                grads = {"proportions": rnd(U, n - 2), "root_proportion": rnd(U), "origin": rnd(U), "clock_rate": rnd(U),
                         "R": rnd(U, m), "delta": rnd(U, m), "s": rnd(U, m), "rho": rnd(U, m)}
                got = reduce_elbo(ut, ll, grads, T, n, m, K)
                want = pack_reference(ut, ll, grads, T, n, m, K)
                assert got.shape == (payload_count(T, n, m),)
                torch.testing.assert_close(got, want, rtol=1e-12, atol=1e-12)
                again = reduce_elbo(ut, ll, grads, T, n, m, K)
                assert torch.equal(got, again)  # fixed summation order


def _fit_worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    hist, theta = _small_fit(dev)
    q.put((rank, hist, theta.cpu().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def _small_fit(dev):
    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.fit import EnsembleELBO
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    data = make_ensemble(3, 10, 300, 1, seed=4, Q=HKY(2.7), clock_rate=0.05, origin=1.0, compress=True)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=dev)
    model = EnsembleELBO(ens, K=5, m=4, delta=2.0, s=0.2, clock_rate=0.05, seed=7)
    hist = model.fit(n_iter=4, step_size=0.1)
    theta = model.theta.detach().clone()
    ens.close()
    return hist, theta


def test_two_gpu_fit_equals_single_gpu_fit(cuda):
    """Data-parallel fit: units dealt round-robin over two NCCL ranks, the WHOLE gradient (global and per-tree local
    parameters) all-reduced, identical Adagrad updates on both ranks == the single-process fit."""
    import socket

    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    hist1, theta1 = _small_fit(cuda)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fit_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, hist, theta in res:
        np.testing.assert_allclose(hist, hist1, rtol=1e-11)
        np.testing.assert_allclose(theta, theta1.cpu().numpy(), rtol=1e-9, atol=1e-11)
    assert np.array_equal(res[0][2], res[1][2])  # replicas stay bit-identical
