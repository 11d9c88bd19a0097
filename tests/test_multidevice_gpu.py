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
