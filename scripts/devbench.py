"""Developer micro-benchmark of the pruning call (not the contract bench): python scripts/devbench.py T n S K [reps]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from vbsky_b200.ensemble import Ensemble
from vbsky_b200.substitution import HKY
from vbsky_b200.synth import make_ensemble

T, n, S, K = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
Ttips = min(T, 4)  # reuse a few alignments to keep setup short
t0 = time.time()
import os
from vbsky_b200 import synth
if os.environ.get('CATERPILLAR'):
    _rt = synth.random_tree
    synth.random_tree = lambda n, rng, caterpillar=False: _rt(n, rng, caterpillar=True)
data = make_ensemble(Ttips, n, S, K, seed=3, Q=HKY(2.7))
dev = torch.device("cuda", 0)
tds = [data.tds[t % Ttips] for t in range(T)]
tips = [data.tips[t % Ttips] for t in range(T)]
ens = Ensemble(tds, tips, HKY(2.7), device=dev)
print(f"setup {time.time()-t0:.1f}s depth_up={ens.depth_up} depth_down={ens.depth_down}")
bl = np.stack([data.heights[t % Ttips].branch_lengths[k] for t in range(T) for k in range(K)]) * data.clock_rate
if os.environ.get('BL_UNIFORM'):
    bl = np.random.default_rng(0).uniform(1e-5, 1e-3, bl.shape)
bl = torch.as_tensor(bl, device=dev)
ut = torch.as_tensor(np.repeat(np.arange(T, dtype=np.int32), K), device=dev)
PREC = os.environ.get('PREC', 'f64')
for want_grad in (True, False):
    for _ in range(2):
        ens.prune(bl, ut, want_grad=want_grad, precision=PREC)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ll, dll, _ = ens.prune(bl, ut, want_grad=want_grad, precision=PREC)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = ens.last_stats()
    print(f"grad={want_grad} T={T} n={n} S={S} K={K}: {ms:.3f} ms/eval  {1e3/ms:.2f} eval/s  alg {st['algorithmic_bytes']/1e9:.2f} GB -> {st['algorithmic_bytes']/ms/1e6:.0f} GB/s  ctas={st['ctas']} smem={st['smem']}  ll0={ll[0].item():.6f}")
