#!/usr/bin/env python
"""Benchmark of the VBSKY hot path on B200 (contract: see the task brief / DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W                    our CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W   CPU arm: the reference's algorithm on the host cores

Metric (BASELINE.json): ELBO+gradient evaluations per second on the SARS-CoV-2-shaped ensemble (config 3: 100 trees x
200 tips x 29,903 site patterns, K = 10 Monte-Carlo samples of node heights per tree, HKY(2.7), m = 50 skyline
intervals, fp64).  One step = one evaluation = value and gradient of ``bdsky.loglik`` (BDSKY tree prior + data
log-likelihood) w.r.t. every parameter for all T*K (tree, sample) units, reduced to [ELBO, d/d global parameters,
d/d per-tree parameters]; the units are dealt round-robin to the N ranks (strong scaling; one all-reduce per step).
Prints ONE JSON line on rank 0.
"""
import argparse
import importlib.util
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
    # name: (T, n, S, K, clock_rate, m)
    "config3": (100, 200, 29903, 10, 1.12e-3, 50),
    "config2": (1, 500, 10000, 16, 1.0, 10),
    "config4": (1000, 50, 2000, 10, 1.12e-3, 50),
    "config1": (10, 10, 134, 10, 0.79e-3, 10),  # HCV-example shape (example.ipynb: 10 trees x 10 tips, ~134 patterns, m = 10)
    "tiny": (4, 16, 300, 2, 1.12e-3, 5),
}
KAPPA = 2.7
NCU_DRAM_BYTES_CONFIG3 = None  # filled from profiles/ (dram read + write of the headline kernel, one ncu --set full capture)
UNIT = "ELBO+grad evals/s"
METRIC = "elbo_grad_evals_per_sec"
PARAM_SEED = 12345


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=0, help="override T (development only; marks the line as reduced)")
    ap.add_argument("--cpu-sample-units", type=int, default=200)
    ap.add_argument("--ref-units-per-step", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--compress", action="store_true",
                    help="merge identical site columns per tree first (np.unique semantics, on the device) as prep_data does; "
                         "not the BASELINE configuration, which uses every site as a pattern")
    return ap.parse_args()


def workload_config(name, T, n, S, K, m, reduced):
    """The `config` object: identical in both arms."""
    return {
        "workload": f"{name}: {T} trees x {n} tips x {S} site patterns, K={K} height samples/tree, HKY({KAPPA}), m={m} skyline "
                    f"intervals; one step = value+grad of bdsky.loglik (BDSKY tree prior + data log-likelihood) w.r.t. all "
                    f"parameters for all {T * K} (tree,sample) units",
        "units": T * K,
        "l2": "working set per step (message scratch + step tables, > 500 MB at config 3) exceeds the 126 MB L2; no explicit flush",
        "reduced": bool(reduced),
    }


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


def native_so_loaded():
    """In-tree shared objects mapped into this process (from /proc/self/maps): evidence of which native code ran."""
    out = set()
    try:
        for line in open("/proc/self/maps"):
            path = line.split(None, 5)[-1].strip() if line.count(" ") >= 5 else ""
            if path.startswith(ROOT) and ".so" in path:
                out.add(os.path.relpath(path, ROOT))
    except OSError:
        pass
    return sorted(out)


def global_draws(K, m):
    """One draw of the global skyline parameters per MC sample (shared by both arms)."""
    return np.random.default_rng(PARAM_SEED).lognormal(0.0, 0.5, (K, m))


def unit_params(data, units, K, m, clock, origin_start):
    """Per-unit parameter rows of ``bdsky.loglik`` for the given (tree, sample) units.  ``origin_start`` is ONE
    dataset-level constant (SeqData.earliest, fasta.py:164,172,368)."""
    Rk = global_draws(K, m)
    U = len(units)
    host = {
        "proportions": np.stack([data.heights[t].proportions[k] for (t, k) in units]),
        "root_proportion": np.array([data.heights[t].root_proportion[k] for (t, k) in units]),
        "origin": np.full(U, data.origin), "origin_start": np.full(U, origin_start), "clock_rate": np.full(U, clock),
        "R": np.stack([Rk[k] for (_, k) in units]),
        "delta": np.full((U, m), 36.5), "s": np.full((U, m), 0.02), "rho": np.zeros((U, m)),
    }
    return {k: np.ascontiguousarray(v, dtype=np.float64) for k, v in host.items()}


# =====================================================================================================================
# CPU arm: the reference's algorithm on the host cores.  Nothing of vbsky_b200 is imported (the product library is
# never mapped): inputs come from vbsky_b200/synth.py loaded BY PATH (NumPy only) with the oracle's tree class.
# =====================================================================================================================
def _load_synth_neutral():
    spec = importlib.util.spec_from_file_location("vbsky_synth_neutral", os.path.join(ROOT, "vbsky_b200", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class CpuArm:
    """oracle/ executed as the CPU implementation: the data term + d/d(branch lengths) by the C/OpenMP port of prune.py's
    two scans on all host threads (oracle/prune_ref.c), the height chain, BDSKY tree prior and their gradients by the
    torch-fp64 restatement of bdsky.py / tree_data.py (oracle/bdsky.py) -- both pinned to the reference's own source by
    tests/test_reference_pin.py."""

    def __init__(self, workload, T_override=0, max_trees=None):
        import networkx as nx

        import oracle.cref as cref
        import oracle.substitution as osub
        import oracle.tree_data as otd

        self.cref = cref
        self.threads = os.cpu_count() or 1
        cref.set_threads(self.threads)  # explicit: launchers such as torchrun export OMP_NUM_THREADS=1
        T, n, S, K, clock, m = WORKLOADS[workload]
        if T_override:
            T = T_override
        self.T, self.n, self.S, self.K, self.clock, self.m = T, n, S, K, clock, m
        self.Q = osub.HKY(KAPPA)
        synth = _load_synth_neutral()

        def tree_from_edges(edges):
            G = nx.DiGraph()
            G.add_edges_from(edges)
            return otd.TreeData.from_nx(G)[0]

        self.ntrees = T if max_trees is None else min(T, max_trees)
        seed = sorted(WORKLOADS).index(workload)
        # the first `ntrees` trees / heights / sequences of the workload (synth.make_ensemble is prefix-stable on the NumPy path)
        full_tds = synth.make_ensemble(T, n, 1, K, seed=seed, Q=self.Q, clock_rate=clock, tree_from_edges=tree_from_edges, missing=0.0)
        self.origin_start = max(float(td.sample_times.max()) for td in full_tds.tds)
        self.data = synth.make_ensemble(self.ntrees, n, S, K, seed=seed, Q=self.Q, clock_rate=clock, tree_from_edges=tree_from_edges)
        assert all(np.array_equal(a.child_parent, b.child_parent) for a, b in zip(self.data.tds, full_tds.tds))

    def units(self, count):
        return [(i % self.ntrees, (i // self.ntrees) % self.K) for i in range(count)]

    def evaluate(self, units):
        """value + gradient of bdsky.loglik for the given units; returns (seconds, checksum)."""
        import torch

        import oracle.bdsky as obdsky
        from oracle.util import TipData

        host = unit_params(self.data, units, self.K, self.m, self.clock, self.origin_start)
        t0 = time.perf_counter()
        acc = 0.0
        for i, (t, k) in enumerate(units):
            td, tip = self.data.tds[t], self.data.tips[t]
            p = {key: torch.tensor(np.atleast_1d(v[i]), requires_grad=(key != "origin_start")) for key, v in host.items()}
            lt, parts = obdsky.loglik(p, td, TipData(None, tip.counts), self.Q, c=(False, True, False), return_parts=True)
            bls = parts["branch_lengths"] * p["clock_rate"][0]
            ld, g, _ = self.cref.prune(td, bls.detach().numpy(), self.Q, tip.patterns, tip.counts, True, False)
            gt = torch.as_tensor(g)
            total = lt + ld + ((bls - bls.detach()) * gt).sum()  # carries the eq. (9) gradient into the height chain
            total.backward()
            acc += total.item() + sum(float(v.grad.sum()) for v in p.values() if v.grad is not None)
        return time.perf_counter() - t0, acc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T0, _, _, K0, _, _ = WORKLOADS[args.workload]
    per_step = max(1, min(args.ref_units_per_step, (args.trees or T0) * K0))
    arm = CpuArm(args.workload, args.trees, max_trees=min(8, per_step))
    T, n, S, K, m = arm.T, arm.n, arm.S, arm.K, arm.m
    units = arm.units(per_step)
    for _ in range(max(args.warmup, 1)):
        arm.evaluate(units[: max(1, per_step // 8)])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.evaluate(units)
    sec_step = (time.perf_counter() - t0) / args.steps
    frac = per_step / (T * K)
    value = frac / sec_step  # evaluations of the WHOLE ensemble per second, extrapolated from the sampled units
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec_step * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, T, n, S, K, m, args.trees),
        "sample": {"units_per_step": per_step, "units_per_evaluation": T * K, "fraction": frac,
                   "note": "ms_per_step is the measured time of ONE SAMPLED step (units_per_step units); value = fraction / seconds_per_step "
                           "extrapolates it to whole-ensemble evaluations per second (units are independent and equally sized)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.threads, "kind": "port",
                         "sample": f"{per_step} of {T * K} (tree,sample) units per step; C/OpenMP port of prune.py (data term + d/dbl, "
                                   f"{arm.threads} threads set explicitly) + torch-fp64 restatement of bdsky.py/tree_data.py (height chain, "
                                   f"BDSKY prior, gradients); both pinned to the reference's source (tests/test_reference_pin.py). The JAX "
                                   f"reference itself is not installable offline (no jax/msprime/tskit wheels)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "native_so_loaded": native_so_loaded(),  # only oracle/libvbsky_oracle.so: vbsky_b200 is never imported by this arm
    }
    assert "vbsky_b200" not in sys.modules and not any("libvbsky_b200" in p for p in line["native_so_loaded"])
    print(json.dumps(line), flush=True)


# =====================================================================================================================
# our arm
# =====================================================================================================================
def build_workload(name, trees_override, device):
    from vbsky_b200.substitution import HKY
    from vbsky_b200.synth import make_ensemble

    T, n, S, K, clock, m = WORKLOADS[name]
    if trees_override:
        T = trees_override
    seed = sorted(WORKLOADS).index(name)
    data = make_ensemble(T, n, S, K, seed=seed, Q=HKY(KAPPA), clock_rate=clock, device=device)
    return data, T, n, S, K, clock, m


def payload_checksum(vec):
    """Order-independent summary of the reduced [ELBO, gradient] vector: must agree across N = 1/2/4/8 to ~1e-12."""
    w = np.cos(np.arange(vec.size) * 0.37) + 1.5
    return float(np.dot(vec, w) / np.abs(vec).sum())


class Evaluator:
    """One rank's share of a workload: resident ensemble, per-unit parameters, the fused step and its reduction."""

    def __init__(self, name, trees_override, dev, rank, world, compress=False):
        import torch

        from vbsky_b200.ensemble import Ensemble
        from vbsky_b200.parallel import payload_count, shard_plan, shard_units, tile_range
        from vbsky_b200.substitution import HKY

        self.torch, self.dev, self.rank, self.world = torch, dev, rank, world
        data, T, n, S, K, clock, m = build_workload(name, trees_override, dev)
        self.sites = S
        if compress:
            from vbsky_b200.tips import PackedTips, compress_patterns_device, pad_tips

            packed = compress_patterns_device(torch.as_tensor(np.stack([tp.patterns for tp in data.tips]), device=dev))
            data = data._replace(tips=pad_tips([PackedTips(p.cpu().numpy(), c.cpu().numpy().astype(np.float64)) for p, c in packed]))
            S = int(data.tips[0].patterns.shape[0])
        self.data, self.T, self.n, self.S, self.K, self.clock, self.m = data, T, n, S, K, clock, m
        self.ens = Ensemble(data.tds, data.tips, HKY(KAPPA), device=dev)
        self.all_units = [(t, k) for t in range(T) for k in range(K)]
        # units shard without any exchange; with too few of them (one big tree, few samples) every rank takes ALL units
        # over its own range of 256-pattern tiles instead (site-pattern-block sharding)
        self.ranks_units, self.ranks_tiles = shard_plan(T * K, self.ens.tiles, world)
        if self.ranks_tiles > 1:
            self.mine = list(self.all_units)
            self.tile_range = tile_range(self.ens.tiles, rank, world)
        else:
            self.mine = shard_units(T, K, rank, world)
            self.tile_range = None
        self.U = len(self.mine)
        origin_start = max(float(td.sample_times.max()) for td in data.tds)
        self.host = unit_params(data, self.mine, K, m, clock, origin_start)
        self.ut_host = np.array([t for (t, _) in self.mine], dtype=np.int32)
        self.params = {k: torch.as_tensor(v, device=dev) for k, v in self.host.items()}
        self.ut = torch.as_tensor(self.ut_host, device=dev)
        self.grads = {k: torch.empty_like(self.params[k]) for k in Ensemble.GRAD_KEYS}
        self.out = {k: torch.empty(self.U, dtype=torch.float64, device=dev) for k in ("ll", "tree", "data")}
        self.payload = torch.zeros(payload_count(T, n, m), dtype=torch.float64, device=dev)
        self.launches_per_step = 9  # chain_forward, bdsky, tables, prune (clip-free), prune (literal-clip), finalize, combine, chain_backward, reduce_elbo

    def step(self):
        from vbsky_b200.parallel import allreduce_elbo, reduce_elbo

        # ELBO likelihood part: mean over the K samples of loglik, summed over trees (optim.py:72-75 batched over the
        # ensemble) and its gradient w.r.t. every parameter; then ONE all-reduce of [ELBO, d/d global, d/d per-tree local]
        if self.tile_range is None:
            self.ens.loglik_device(self.params, self.ut, self.m, flags=7, grads=self.grads, out=self.out)
            reduce_elbo(self.ut, self.out["ll"], self.grads, self.T, self.n, self.m, self.K, out=self.payload)
            allreduce_elbo(self.payload)
        else:
            # site-block mode: the exchange is the sum of the per-unit partial [ll_data | d/d bl] inside the step; afterwards
            # every rank holds the complete per-unit results, so the payload needs no second exchange
            self.ens.loglik_device(self.params, self.ut, self.m, flags=7, grads=self.grads, out=self.out, tile_range=self.tile_range,
                                   exchange=allreduce_elbo)
            reduce_elbo(self.ut, self.out["ll"], self.grads, self.T, self.n, self.m, self.K, out=self.payload)

    def time_steps(self, steps, warmup, dist=None, sampler=None):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            self.step()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        if sampler is not None:
            sampler.start()
        self.ens.kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()  # lets `ncu --profile-from-start off` capture exactly the timed steps
        e0.record()
        for _ in range(steps):
            self.step()
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        if dist is not None:
            dist.barrier()
        ms_total = e0.elapsed_time(e1)
        kms, klaunches = self.ens.kernel_timing(False)
        clocks = sampler.stop() if sampler is not None else None
        tms = torch.tensor([ms_total], dtype=torch.float64, device=self.dev)
        if dist is not None:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return tms.item() / steps, kms / max(klaunches, 1), clocks

    def time_graph_replay(self, steps):
        """The same step captured once into a CUDA graph and replayed (single rank only): what a fitting loop does with a
        launch-bound step.  Returns ms per step, or None if capture is unavailable."""
        torch = self.torch
        if self.world > 1:
            return None
        try:
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self.step()
            for _ in range(3):
                graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        except Exception:
            return None

    def close(self):
        self.ens.close()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram read + write bytes of the headline kernel at config 3 / 1 GPU from the committed ncu summary (profiles/)."""
    path = os.path.join(ROOT, "profiles", "headline_kernel_dram_bytes.json")
    if os.path.exists(path):
        return json.load(open(path)).get("dram_bytes_per_launch")
    return NCU_DRAM_BYTES_CONFIG3


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from vbsky_b200.ensemble import Ensemble

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
    D = dist if world > 1 else None

    ev = Evaluator(args.workload, args.trees, dev, rank, world, args.compress)
    T, n, S, K, m, U = ev.T, ev.n, ev.S, ev.K, ev.m, ev.U
    ms_step, kern_ms, clocks = ev.time_steps(args.steps, args.warmup, D, ClockSampler(local_rank) if rank == 0 else None)
    stats = ev.ens.last_stats()
    payload = ev.payload.cpu().numpy()  # all-reduced: identical on every rank
    ll_dev = ev.out["ll"].cpu().numpy()

    # ---- end-to-end through the host-buffer C ABI: pinned params -> H2D -> kernels -> D2H of ll + every per-unit gradient;
    #      at N > 1 the same payload as the device leg is then formed from the host buffers and all-reduced
    pin = lambda a: torch.as_tensor(a).pin_memory().numpy()
    hp = {k: pin(v) for k, v in ev.host.items()}
    hut = pin(ev.ut_host)
    hg = {k: pin(np.empty_like(ev.host[k])) for k in Ensemble.GRAD_KEYS}
    ho = {k: pin(np.empty(U)) for k in ("ll", "tree", "data")}
    e2e_steps = max(3, min(args.steps, 10))
    pay_dev = torch.empty_like(ev.payload)
    pay_pin = torch.empty(ev.payload.numel(), dtype=torch.float64).pin_memory()
    first = np.r_[0, np.flatnonzero(np.diff(ev.ut_host)) + 1]  # segment starts of the (sorted) unit -> tree map
    trees_here = ev.ut_host[first]

    hp_t = {k: torch.from_numpy(v) for k, v in hp.items()}  # pinned views

    def e2e_step():
        if ev.tile_range is not None:
            # site-block mode: pinned host parameters -> device, the sharded step (with its in-step exchange), payload -> host
            for k, v in hp_t.items():
                ev.params[k].copy_(v, non_blocking=True)
            ev.step()
            pay_pin.copy_(ev.payload, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return
        ev.ens.loglik_host(hp, hut, m, flags=7, grads=hg, out=ho)
        if world > 1:
            v = pay_pin.numpy()
            v[:] = 0.0
            v[0] = ho["ll"].sum()
            v[1], v[2] = hg["origin"].sum(), hg["clock_rate"].sum()
            for i, k in enumerate(("R", "delta", "s", "rho")):
                v[3 + i * m: 3 + (i + 1) * m] = hg[k].sum(0)
            loc = v[3 + 4 * m:].reshape(T, n - 1)
            loc[trees_here, : n - 2] = np.add.reduceat(hg["proportions"], first, axis=0)
            loc[trees_here, n - 2] = np.add.reduceat(hg["root_proportion"], first)
            v /= K
            pay_dev.copy_(pay_pin, non_blocking=True)
            dist.all_reduce(pay_dev)
            pay_pin.copy_(pay_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()

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
        assert np.allclose(pay_pin.numpy(), payload, rtol=1e-10, atol=1e-10 * np.abs(payload).max()), "e2e payload differs from the device leg"
    assert np.isfinite(payload).all(), "non-finite ELBO / gradient"
    h2d = sum(v.nbytes for v in ev.host.values()) + ev.ut_host.nbytes
    if ev.tile_range is None:
        assert np.allclose(ho["ll"], ll_dev, rtol=1e-12), "host-buffer path disagrees with the device path"
        assert all(np.isfinite(v).all() for v in hg.values()), "non-finite gradient"
        d2h = 3 * U * 8 + sum(v.nbytes for v in hg.values())
        if world > 1:
            h2d += payload.nbytes
            d2h += payload.nbytes
    else:
        h2d -= ev.ut_host.nbytes
        d2h = payload.nbytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    achieved = stats["algorithmic_bytes"] / (kern_ms * 1e-3) / 1e9 if kern_ms > 0 else 0.0
    # SURVEY 8(d)'s B_min counts one stored message per non-root internal node; this kernel rebuilds the cherries' messages
    # on chip instead, so it moves fewer bytes than B_min.  Both figures are reported; `frac` uses the bytes actually moved.
    bmin_bytes = stats["algorithmic_bytes"] + U * float(ev.S) * max(n - 2 - ev.ens.mean_stored_messages, 0) * 64.0
    headline = args.workload == "config3" and world == 1 and not args.trees and not args.compress
    npay = payload.size
    line = {
        "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, T, n, ev.sites, K, m, args.trees),
        "sharding": (f"{T * K} units round-robin over {world} rank(s); one all-reduce of [ELBO, d/d global params, d/d per-tree local params] "
                     f"({npay} doubles) per step, payload built by one kernel (vbsky_reduce_elbo)") if ev.tile_range is None else
                    (f"site-pattern blocks: every rank evaluates all {T * K} units over its own range of the {ev.ens.tiles} 256-pattern tiles "
                     f"(rank 0: tiles {ev.tile_range}); one all-reduce of the per-unit partial [ll_data | d/d bl] "
                     f"({U * (2 * n - 1)} doubles) inside the step"),
        # the reduced result of the last step: identical (to rounding of the summation order) for every N
        "elbo": float(payload[0]), "grad_checksum": payload_checksum(payload[1:]), "grad_l1": float(np.abs(payload[1:]).sum()),
        "e2e": {"value": 1e3 / e2e_ms.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "note": ("vbsky_loglik_value_and_grad_host: pinned host parameters in, ll + all per-unit gradients out, per step; tip codes "
                         "resident (uploaded once, as prep_data is one-off in the reference)"
                         + ("; then the same payload as the device leg is formed from the host buffers and all-reduced (NCCL)" if world > 1 else ""))
                        if ev.tile_range is None else
                        "site-block mode: pinned host parameters -> device, the sharded step with its in-step all-reduce, reduced payload -> pinned host"},
        "gpu_launches": int(ev.launches_per_step * args.steps),
        "native_so_loaded": native_so_loaded(),
        "clocks": clocks,
        "roofline": {
            "bound": "hbm", "kernel": "prune_kernel<grad, clip-free>", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "peak_source": peak_src,
            "traffic": ncu_traffic() if headline else None,
            "algorithmic_bytes_per_launch": stats["algorithmic_bytes"], "kernel_ms": kern_ms, "kernel_share_of_step": kern_ms / ms_step,
            "note": "algorithmic bytes = what this algorithm must move: messages of the non-root internal nodes that are not cherries "
                    "(written once, read once), tip codes, counts, gradient partials; cherries' messages are rebuilt on chip. The kernel is "
                    "no longer HBM-bound (ncu: LSU data pipe 68 %, issue slots 48 %, FP64 pipe 41 %, DRAM 48 %): per-step dependency latency "
                    "at 12 consumer warps/SM limits it",
            "survey_bmin_bytes_per_launch": bmin_bytes,
            "frac_at_survey_bmin": (bmin_bytes / (kern_ms * 1e-3) / 1e9 / peak) if kern_ms > 0 else None,
        },
    }
    if args.compress:
        line["patterns"] = f"{ev.sites} sites per tree compressed on the device to at most {S} unique site patterns (padded to the ensemble maximum)"
    ev.close()
    del ev
    torch.cuda.empty_cache()

    if world == 1 and not args.no_other_configs and headline:
        # the other BASELINE shapes (parity-test cases, not bench lines), measured after the timed region
        others = {}
        for name in ("config1", "config2", "config4"):
            try:
                o = Evaluator(name, 0, dev, 0, 1)
                oms, okern, _ = o.time_steps(10, 3)
                ost = o.ens.last_stats()
                gms = o.time_graph_replay(20)
                others[name] = {"ms_per_step": oms, "evals_per_s": 1e3 / oms, "ms_per_step_graph_replay": gms, "units": o.T * o.K,
                                "shape": f"{o.T} trees x {o.n} tips x {o.S} patterns, K={o.K}, m={o.m}",
                                "prune_kernel_ms": okern, "roofline_frac": ost["algorithmic_bytes"] / (okern * 1e-3) / 1e9 / peak if okern > 0 else None}
                o.close()
                del o
                torch.cuda.empty_cache()
            except Exception as e:  # never lose the headline line to a side measurement
                others[name] = {"error": repr(e)[:200]}
        line["other_configs"] = others

    if not args.no_cpu_baseline and world == 1:
        try:
            arm = CpuArm(args.workload, args.trees, max_trees=8)
            nu = max(1, args.cpu_sample_units)
            sec, _ = arm.evaluate(arm.units(nu))
            v = (nu / (T * K)) / sec
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": arm.threads, "kind": "port",
                                    "sample": f"{nu} of {T * K} (tree,sample) units ({sec:.1f} s): C/OpenMP port of prune.py on {arm.threads} threads + "
                                              f"torch-fp64 restatement of the height chain / BDSKY prior (oracle/), scaled"}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
