"""BASELINE.json config 5 (sweep over tips n, patterns S, samples K, one tree) for the pruning call, at 1..8 GPUs:
    python scripts/sweep.py > gpurun_out/sweep.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep.py > ...
One JSON line per point: ms per value+gradient evaluation (max over ranks), evaluations/s, algorithmic GB/s of
prune_kernel's traffic summed over ranks and its fraction of N x the measured HBM peak (MEASURED_PEAKS.json).
With N > 1 the K units are dealt to the ranks when there are enough of them, otherwise every rank takes all units over
its own range of 256-pattern tiles and the partial [ll | d/d bl] are all-reduced (parallel.shard_plan)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from vbsky_b200.ensemble import Ensemble  # noqa: E402
from vbsky_b200.parallel import shard_plan, tile_range  # noqa: E402
from vbsky_b200.substitution import HKY  # noqa: E402
from vbsky_b200.synth import make_ensemble  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peak = 6450.9
if os.path.exists("MEASURED_PEAKS.json"):
    peak = json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", peak)
grid = [(n, S, 16) for n in (64, 200, 1024, 4096) for S in (1000, 10000, 100000)]
grid += [(200, 29903, K) for K in (1, 4, 64)]
for n, S, K in grid:
    t0 = time.time()
    data = make_ensemble(1, n, S, K, seed=n + S, Q=HKY(2.7), device=dev)
    ens = Ensemble(data.tds, data.tips, HKY(2.7), device=dev)
    ru, rt = shard_plan(K, ens.tiles, world, min_units_per_rank=8)
    mine = list(range(K)) if rt > 1 else list(range(rank, K, world))
    tr = tile_range(ens.tiles, rank, world) if rt > 1 else None
    bl = torch.as_tensor(data.heights[0].branch_lengths[mine] * data.clock_rate, device=dev)
    ut = torch.zeros(len(mine), dtype=torch.int32, device=dev)

    def evaluate():
        ll, dll, _ = ens.prune(bl, ut, tile_range=tr)
        if tr is not None:
            buf = torch.cat([ll, dll.reshape(-1)])
            dist.all_reduce(buf)
            ll = buf[: ll.numel()]
        return ll, dll

    for _ in range(2):
        evaluate()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    reps = 5 if n * S * K < 5e8 * world else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ll, dll = evaluate()
    e1.record()
    torch.cuda.synchronize()
    st = ens.last_stats()
    v = torch.tensor([e0.elapsed_time(e1) / reps, float(st["algorithmic_bytes"])], dtype=torch.float64, device=dev)
    if world > 1:
        vmax = v.clone()
        dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(v)
        ms, alg = vmax[0].item(), v[1].item()
    else:
        ms, alg = v[0].item(), v[1].item()
    if rank == 0:
        gbs = alg / ms / 1e6
        print(json.dumps({"n_gpus": world, "n": n, "S": S, "K": K, "ms_per_eval": round(ms, 4), "evals_per_s": round(1e3 / ms, 2),
                          "sharding": "site-blocks" if rt > 1 else "units", "work_items_rank0": len(mine) * (ens.tiles if tr is None else tr[1] - tr[0]),
                          "ctas": st["ctas"], "smem": st["smem"], "algorithmic_gb": round(alg / 1e9, 3), "gbs": round(gbs, 1),
                          "frac_of_hbm_peak": round(gbs / (peak * world), 3),
                          "finite": bool(torch.isfinite(ll).all() and torch.isfinite(dll).all()), "setup_s": round(time.time() - t0, 1)}), flush=True)
    ens.close()
    del data, ens, bl
    torch.cuda.empty_cache()
if world > 1:
    dist.destroy_process_group()
