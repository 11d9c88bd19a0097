"""BASELINE.json config 5 (sweep over tips n, patterns S, samples K) for the pruning call, one tree:
python scripts/sweep.py > gpurun_out/sweep.jsonl.  One JSON line per point: ms per value+gradient evaluation,
algorithmic GB/s of prune_kernel's traffic and its fraction of the measured HBM peak (MEASURED_PEAKS.json)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from vbsky_b200.ensemble import Ensemble  # noqa: E402
from vbsky_b200.substitution import HKY  # noqa: E402
from vbsky_b200.synth import make_ensemble  # noqa: E402

dev = torch.device("cuda", 0)
peak = 6450.9
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
grid = [(n, S, 16) for n in (64, 200, 1024, 4096) for S in (1000, 10000, 100000)]
grid += [(200, 29903, K) for K in (1, 4, 64)]
for n, S, K in grid:
    t0 = time.time()
    data = make_ensemble(1, n, S, K, seed=n + S, Q=HKY(2.7), device=dev)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=dev)
    bl = torch.as_tensor(data.heights[0].branch_lengths * data.clock_rate, device=dev)
    ut = torch.zeros(K, dtype=torch.int32, device=dev)
    for _ in range(2):
        ens.prune(bl, ut)
    torch.cuda.synchronize()
    reps = 5 if n * S * K < 5e8 else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ll, dll, _ = ens.prune(bl, ut)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    st = ens.last_stats()
    gbs = st["algorithmic_bytes"] / ms / 1e6
    print(json.dumps({"n": n, "S": S, "K": K, "ms_per_eval": round(ms, 4), "evals_per_s": round(1e3 / ms, 2),
                      "work_items": K * ((S + 255) // 256), "ctas": st["ctas"], "smem": st["smem"],
                      "algorithmic_gb": round(st["algorithmic_bytes"] / 1e9, 3), "gbs": round(gbs, 1), "frac_of_hbm_peak": round(gbs / peak, 3),
                      "finite": bool(torch.isfinite(ll).all() and torch.isfinite(dll).all()), "setup_s": round(time.time() - t0, 1)}), flush=True)
    ens.close()
    del data, ens, bl
    torch.cuda.empty_cache()
