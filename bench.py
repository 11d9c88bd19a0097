#!/usr/bin/env python
"""Benchmark of the VBSKY hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W                 our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   CPU restatement of the reference

Metric (BASELINE.json): ELBO+gradient evaluations per second on the SARS-CoV-2-shaped ensemble
(config 3: 100 trees x 200 tips x 29,903 site patterns, K=10 Monte-Carlo samples of node heights per
tree, HKY(2.7), fp64).  One step = one evaluation = value and gradient of the data log-likelihood
for all T*K (tree, sample) units; the units are sharded round-robin over the N ranks (strong
scaling, one small all-reduce of the ELBO value per step).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (T, n, S, K, clock_rate)
    "config3": (100, 200, 29903, 10, 1.12e-3),
    "config2": (1, 500, 10000, 16, 1.0),
    "config4": (1000, 50, 2000, 10, 1.12e-3),
    "config1": (10, 10, 134, 10, 1.12e-3),  # HCV-example shape (example.ipynb: 10 trees x 10 tips, ~134 patterns)
    "tiny": (4, 16, 300, 2, 1.12e-3),
}
KAPPA = 2.7
M_INTERVALS = 50  # skyline intervals (paper/covid.py m=50)
NCU_DRAM_BYTES_CONFIG3 = 204.845845e9 + 191.205230e9  # dram read + write of prune_kernel<1,0,1>, profiles/r1_prune_kernel_bench_ncu_summary.txt
UNIT = "ELBO+grad evals/s"
METRIC = "elbo_grad_evals_per_sec"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=0, help="override T (development only; marks the line as reduced)")
    ap.add_argument("--cpu-sample-units", type=int, default=400)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--compress", action="store_true",
                    help="merge identical site columns per tree first (np.unique semantics, on the device) as prep_data does; "
                         "not the BASELINE configuration, which uses every site as a pattern")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(args, device=None, trees_needed=None):
    """Seeded synthetic ensemble of the named shape.  Returns (data, T, n, S, K, clock)."""
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    T, n, S, K, clock = WORKLOADS[args.workload]
    if args.trees:
        T = args.trees
    seed = sorted(WORKLOADS).index(args.workload)
    data = make_ensemble(T if trees_needed is None else trees_needed, n, S, K, seed=seed, Q=HKY(KAPPA), clock_rate=clock, device=device)
    return data, T, n, S, K, clock


def cpu_reference_sample(data, units, K, clock):
    """Times the C/OpenMP oracle port on the given units (list of (tree, sample)); returns seconds."""
    import oracle.cref as cref
    import oracle.substitution as osub
    import oracle.tree_data as otd

    Q = osub.HKY(KAPPA)
    t0 = time.perf_counter()
    acc = 0.0
    for (t, k) in units:
        td = data.tds[t]
        otdd = otd.TreeData(td.postorder, td.child_parent, td.parent_child, td.sample_times)
        ll, g, _ = cref.prune(otdd, data.heights[t].branch_lengths[k] * clock, Q, data.tips[t].patterns, data.tips[t].counts, True, False)
        acc += ll + g.sum()
    return time.perf_counter() - t0, cref.threads(), acc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, n, S, K, clock = WORKLOADS[args.workload]
    if args.trees:
        T = args.trees
    per_step = 40  # units timed per step (a bounded sample of the T*K units of one evaluation)
    ntrees = min(T, per_step)
    data, T, n, S, K, clock = build_workload(args, device=None, trees_needed=ntrees)
    units = [(i % ntrees, (i // ntrees) % K) for i in range(per_step)]
    for _ in range(args.warmup):
        cpu_reference_sample(data, units[:1], K, clock)
    t0 = time.perf_counter()
    threads = 1
    for _ in range(args.steps):
        _, threads, _ = cpu_reference_sample(data, units, K, clock)
    dt = (time.perf_counter() - t0) / args.steps
    sec_per_eval = dt * (T * K) / per_step
    value = 1.0 / sec_per_eval
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_per_eval * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {T} trees x {n} tips x {S} site patterns, K={K}, HKY({KAPPA}), value+grad of the data term"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{per_step} of {T * K} (tree,sample) units per step, C/OpenMP restatement of prune.py (JAX reference not installable: no jax/msprime/tskit wheels, no network); scaled by {T * K}/{per_step}"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from vbsky_b200.ensemble import Ensemble
    from vbsky_b200.parallel import allreduce_elbo, pack_global, shard_units
    from vbsky_b200.substitution import HKY

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("multi-GPU runs are launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    data, T, n, S, K, clock = build_workload(args, device=dev)
    sites = S
    if args.compress:
        from vbsky_b200.tips import PackedTips, compress_patterns_device, pad_tips

        packed = compress_patterns_device(torch.as_tensor(np.stack([tp.patterns for tp in data.tips]), device=dev))
        data = data._replace(tips=pad_tips([PackedTips(p.cpu().numpy(), c.cpu().numpy().astype(np.float64)) for p, c in packed]))
        S = int(data.tips[0].patterns.shape[0])
    Q = HKY(KAPPA)
    ens = Ensemble(data.tds, data.tips, Q, device=dev)
    # (tree, sample) units, round-robin over ranks
    units = [(t, k) for t in range(T) for k in range(K)]
    mine = shard_units(T, K, rank, world)
    U = len(mine)
    m = M_INTERVALS
    rng = np.random.default_rng(12345)
    Rk = rng.lognormal(0.0, 0.5, (K, m))  # one draw of the global skyline parameters per MC sample
    host = {
        "proportions": np.stack([data.heights[t].proportions[k] for (t, k) in mine]),
        "root_proportion": np.array([data.heights[t].root_proportion[k] for (t, k) in mine]),
        "origin": np.full(U, data.origin),
        "origin_start": np.array([data.tds[t].sample_times.max() for (t, _) in mine]),
        "clock_rate": np.full(U, clock),
        "R": np.stack([Rk[k] for (_, k) in mine]),
        "delta": np.full((U, m), 36.5), "s": np.full((U, m), 0.02), "rho": np.zeros((U, m)),
    }
    host = {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in host.items()}
    ut_host = np.array([t for (t, _) in mine], dtype=np.int32)
    params = {k: torch.as_tensor(v, device=dev) for k, v in host.items()}
    ut = torch.as_tensor(ut_host, device=dev)
    grads = {k: torch.empty_like(params[k]) for k in Ensemble.GRAD_KEYS}
    out = {k: torch.empty(U, dtype=torch.float64, device=dev) for k in ("ll", "tree", "data")}
    red = torch.zeros(1 + 2 + 4 * m, dtype=torch.float64, device=dev)

    def step():
        # ELBO likelihood part: mean over the K samples of loglik, summed over trees (optim.py:72-75, fasta.py:456-487
        # batched over the ensemble), and its gradient; gradients of the per-tree (local) parameters stay on their owner.
        ens.loglik_device(params, ut, m, flags=7, grads=grads, out=out)
        allreduce_elbo(pack_global(out["ll"], grads, K, out=red))

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ens.kernel_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` capture exactly the timed steps
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    kms, klaunches = ens.kernel_timing(False)
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = tms.item() / args.steps
    stats = ens.last_stats()
    ll_dev = out["ll"].cpu().numpy()

    # ---- end-to-end through the host-buffer C ABI: pinned params -> H2D -> 8 kernels -> D2H of ll + every gradient
    pin = lambda a: torch.as_tensor(a).pin_memory().numpy()
    hp = {k: pin(v) for k, v in host.items()}
    hut = pin(ut_host)
    hg = {k: pin(np.empty_like(host[k])) for k in Ensemble.GRAD_KEYS}
    ho = {k: pin(np.empty(U)) for k in ("ll", "tree", "data")}
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step():
        ens.loglik_host(hp, hut, m, flags=7, grads=hg, out=ho)
        if world > 1:
            v = torch.tensor([float(ho["ll"].sum()) / K], dtype=torch.float64, device=dev)
            dist.all_reduce(v)

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([(time.perf_counter() - t0) * 1e3 / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    assert np.allclose(ho["ll"], ll_dev, rtol=1e-12), "host-buffer path disagrees with the device path"
    assert np.isfinite(ll_dev).all() and all(np.isfinite(v).all() for v in hg.values()), "non-finite ELBO / gradient"
    h2d = sum(v.nbytes for v in host.values()) + ut_host.nbytes
    d2h = 3 * U * 8 + sum(v.nbytes for v in hg.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    kern_ms = kms / max(klaunches, 1)
    achieved = stats["algorithmic_bytes"] / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    line = {
        "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {T} trees x {n} tips x {S} site patterns, K={K} height samples/tree, HKY({KAPPA}), "
                        f"m={M_INTERVALS} skyline intervals; one step = value+grad of bdsky.loglik (BDSKY tree prior + data log-likelihood) "
                        f"w.r.t. all parameters for all {T * K} (tree,sample) units",
            "sharding": f"{T * K} units round-robin over {world} rank(s); one all-reduce of [ELBO, d/d global params] ({1 + 2 + 4 * M_INTERVALS} doubles) per step",
            "l2": "working set per step (message scratch + step tables, > 500 MB) exceeds the 126 MB L2; no explicit flush",
            "reduced": bool(args.trees),
            **({"patterns": f"{sites} sites per tree compressed on the device to at most {S} unique site patterns (padded to the ensemble maximum)"}
               if args.compress else {}),
        },
        "e2e": {"value": 1e3 / e2e_ms.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": "vbsky_loglik_value_and_grad_host: pinned host parameters in, ll + all gradients out, per step; tip codes resident (uploaded once, as prep_data is one-off in the reference)"},
        "gpu_launches": int(8 * args.steps),  # chain_forward, bdsky, tables, prune x2 (clip-free + literal-clip units), finalize, combine, chain_backward
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "kernel": "prune_kernel<true>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src,
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from profiles/r1_prune_kernel_bench_ncu_summary.txt
            # (one ncu --set full capture of this workload on one GPU); null for any other workload
            "traffic": NCU_DRAM_BYTES_CONFIG3 if (args.workload == "config3" and world == 1 and not args.trees and not args.compress) else None,
            "algorithmic_bytes_per_launch": stats["algorithmic_bytes"], "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_step,
        },
    }
    if not args.no_cpu_baseline and world == 1:
        nu = max(1, args.cpu_sample_units)
        sample_units = units[:: max(1, len(units) // nu)][:nu]
        sec, threads, _ = cpu_reference_sample(data, sample_units, K, clock)
        v = 1.0 / (sec * len(units) / len(sample_units))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"{len(sample_units)} of {len(units)} (tree,sample) units, C/OpenMP restatement of prune.py (data term + d/dbl; the O(n+m) tree prior is not included), {sec:.1f} s, scaled"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
