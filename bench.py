#!/usr/bin/env python
"""bench.py - edges/sec of the Equiformer graph-attention hot path (fwd+bwd training step) on B200.

Contract (one JSON line on stdout from rank 0):
    python bench.py --gpus N --steps K --warmup W            # sm_100a kernels behind the drop-in modules
    python bench.py --impl reference --gpus N --steps K --warmup W   # the oracle restatement on the host cores

Workload (BASELINE.json configs[1]): synthetic QM9-like batch, 128 molecules x ~18 atoms, radius 5 A, model
``graph_attention_transformer_nonlinear_l2`` (Lmax=2, 6 blocks), fp32, one training step = forward + L1 loss +
backward (+ gradient all-reduce for N>1) + AdamW update.  Metric: edges processed per second, whole job.

* ``value``    - inputs already resident in HBM when the timed region starts.
* ``e2e``      - same step through the public module API starting from pinned HOST tensors: H2D copy of
                 (pos, batch, node_atom, target) and D2H read of the loss inside the timed region, every step.
* ``roofline`` - the dominant hand-written kernel by summed device time (CUDA events around every launch of our
                 kernels, on the launching stream, inside the timed region): algorithmic bytes / time vs the measured
                 HBM copy bandwidth in MEASURED_PEAKS.json.
* ``cpu_baseline`` - the oracle (reference-style op chain, torch CPU) timed on a bounded sample of the same batch
                     (N = 1 only; null in multi-GPU runs).

Multi-GPU: one process per GPU (torchrun), independent molecule batch per rank (weak scaling), NCCL all-reduce of
one flat gradient bucket per step; time = max over ranks, measured with CUDA events between barriers.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL_NAME = "graph_attention_transformer_nonlinear_l2"
N_GRAPHS = 128
METRIC = "edges/sec fwd+bwd, QM9 Lmax=2 batch"
WORKLOAD = "QM9 synthetic batch: 128 molecules x ~18 atoms, radius 5 A, Lmax=2 (graph_attention_transformer_nonlinear_l2), fwd+bwd+AdamW"


def synthetic_batch(seed: int):
    from equiformer_b200.synthetic import qm9_like_batch
    pos, batch, z = qm9_like_batch(N_GRAPHS, seed=seed)
    g = torch.Generator().manual_seed(seed + 1000)
    target = torch.randn(N_GRAPHS, 1, generator=g)
    return pos, batch, z, target


def count_edges(pos, batch, radius=5.0) -> int:
    from equiformer_b200.graph import radius_graph
    return int(radius_graph(pos, radius, batch, max_num_neighbors=1000).shape[1])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            try:
                sm.append(float(s[0])); mx.append(float(s[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(kernel: str):
    """dram bytes per launch of the dominant kernel from the committed ncu summary (profiles/), if present."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get(kernel)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------- reference arm


def oracle_step(params, cfg, pos, batch, z, target, n_graphs):
    """One fwd+bwd of the oracle op chain (CPU, fp32) - what `--impl reference` and cpu_baseline time."""
    from oracle import equiformer_ref as R
    for v in params.values():
        if v.is_floating_point():
            v.grad = None
    out = R.model_forward(params, cfg, pos, batch, z, n_graphs)
    loss = (out - target).abs().mean()
    loss.backward()
    return float(loss.detach())


def cpu_sample(n_sample_graphs: int, seed: int = 0):
    from equiformer_b200.nets import model_entrypoint
    from oracle import equiformer_ref as R
    pos, batch, z, target = synthetic_batch(seed)
    keep = batch < n_sample_graphs
    pos, batch, z, target = pos[keep], batch[keep], z[keep], target[:n_sample_graphs]
    torch.manual_seed(0)
    model = model_entrypoint(MODEL_NAME)(irreps_in="5x0e", radius=5.0, num_basis=128)
    params = {k: v.requires_grad_(v.is_floating_point() and v.numel() > 0)
              for k, v in R.cast_params(model.state_dict(), torch.float32).items()}
    edges = count_edges(pos, batch)
    return params, R.Config(), pos, batch, z, target, n_sample_graphs, edges


def host_threads() -> int:
    """Threads for the CPU arm: one per physical core (torchrun pins OMP_NUM_THREADS=1, and one thread per logical
    core oversubscribes the oracle's small matmuls by two orders of magnitude -- measured 6 vs 620 edges/s)."""
    try:
        import psutil
        n = psutil.cpu_count(logical=False)
    except Exception:
        n = None
    if not n:
        n = max(1, (os.cpu_count() or 2) // 2)
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    return max(1, n)


def run_reference(args):
    """The reference's CPU path stand-in: oracle op chain on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(host_threads())
    threads = torch.get_num_threads()
    n_sample = args.ref_graphs
    params, cfg, pos, batch, z, target, n_graphs, edges = cpu_sample(n_sample)
    for _ in range(args.warmup):
        oracle_step(params, cfg, pos, batch, z, target, n_graphs)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(params, cfg, pos, batch, z, target, n_graphs)
    dt = (time.perf_counter() - t0) / args.steps
    value = edges / dt
    sample = f"{n_graphs} of the {N_GRAPHS} molecules ({edges} edges) per step, oracle op chain fwd+bwd, torch CPU fp32"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": value, "unit": "edges/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------- our arm


def run_ours(args):
    import torch.distributed as dist
    from equiformer_b200 import _lib, ops
    from equiformer_b200.nets import model_entrypoint
    from equiformer_b200.parallel import FlatAdamW, FlatGradAllReduce, broadcast_parameters, init_distributed

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the edge path has no CPU implementation "
                           "(use --impl reference for the CPU oracle)")
    if not _lib.LIB_PATH.exists():
        _lib.build()
    rank, local, world = init_distributed("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.backends.cuda.matmul.allow_tf32 = False  # fp32 parity: the reference trains with --no-amp
    torch.backends.cudnn.allow_tf32 = False

    torch.manual_seed(0)
    model = model_entrypoint(MODEL_NAME)(irreps_in="5x0e", radius=5.0, num_basis=128).to(dev)
    model.train()
    for m in model.modules():  # alpha_drop is the only stochastic piece; parity runs and this bench use p=0
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    broadcast_parameters(model)
    bucket = FlatGradAllReduce(model.parameters())
    opt = FlatAdamW(model.named_parameters(), bucket, lr=5e-4, weight_decay=5e-3, no_decay=model.no_weight_decay())

    pos, batch, z, target = synthetic_batch(seed=rank)  # independent molecules per rank (weak scaling)
    edges_local = count_edges(pos, batch)
    host = [t.pin_memory() for t in (pos, batch, z, target)]
    dev_in = [t.to(dev) for t in host]
    h2d_bytes = sum(t.numel() * t.element_size() for t in host)

    def l1_loss(out, tgt):
        return (out - tgt).abs().mean()

    def step_eager(inputs):
        p, b, zz, tgt = inputs
        bucket.zero_grad()
        out = model(f_in=None, pos=p, batch=b, node_atom=zz, n_graphs=tgt.shape[0])
        loss = l1_loss(out, tgt)
        loss.backward()
        bucket.reduce()
        opt.step()
        return loss

    graphed = None
    if args.graph:
        from equiformer_b200.graphs import GraphedForwardBackward
        graphed = GraphedForwardBackward(model, l1_loss, bucket, max_radius=5.0)

    def step_graph(inputs):
        loss = graphed(*inputs)        # neighbour search (eager) + CUDA-graph replay of forward, loss, backward
        bucket.reduce()
        opt.step()
        return loss

    step = step_graph if args.graph else step_eager

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n_steps, from_host, profile, fn=None):
        fn = fn or step
        barrier()
        ops.PROFILE = profile
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        last = None
        for _ in range(n_steps):
            if from_host:
                inputs = [t.to(dev, non_blocking=True) for t in host]
                last = fn(inputs).item()            # D2H read of the loss every step
            else:
                last = fn(dev_in)
        end.record()
        barrier()
        ops.PROFILE = None
        ms = torch.tensor([start.elapsed_time(end)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / n_steps, last

    for _ in range(max(args.warmup, 3)):
        step_eager(dev_in)
    torch.cuda.synchronize()
    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    if args.graph:
        for _ in range(max(args.warmup, 3)):
            step(dev_in)                            # includes the one-off capture for this (atoms, edges) signature
        torch.cuda.synchronize()

    # per-kernel CUDA-event timing needs eager launches (events cannot be read back from inside a graph replay):
    # an instrumented eager pass of the same step gives the roofline numbers, the headline is timed on `step`.
    profile = ops.KernelProfile(time_events=True, presleep_cycles=200_000)   # ~0.1 ms GPU-side head start per timed launch
    ms_eager, _ = timed(args.steps, from_host=False, profile=profile, fn=step_eager)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    # `ncu --profile-from-start off` (tools/rounds/gpu_round10.sh) lists exactly the launches of the headline region
    mark = os.environ.get("EQF_BENCH_CUDA_PROFILER") == "1"
    if mark:
        torch.cuda.cudart().cudaProfilerStart()
    if args.graph:
        ms_step, _ = timed(args.steps, from_host=False, profile=None)
    else:
        ms_step = ms_eager
    if mark:
        torch.cuda.cudart().cudaProfilerStop()
    clocks = sampler.stop() if sampler else None
    ms_e2e, last_loss = timed(args.steps, from_host=True, profile=None)

    edges_total = torch.tensor([edges_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(edges_total)
    edges_total = edges_total.item()

    if rank == 0:
        peak, peak_src = measured_peaks()
        summ = profile.summary()
        # hand-written kernels only (everything ops.py times except the CUTLASS-template fallback GEMM)
        own = {k: v for k, v in summ.items() if not k.startswith("gemm_fast_f32")}
        dominant = max(own, key=lambda k: own[k]["ms"]) if own else None
        roof = None
        kernels = {}
        for name, d in summ.items():
            kernels[name] = {"launches_per_step": d["launches"] / args.steps, "ms_per_step": d["ms"] / args.steps,
                             "gb_s": d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else None}
        if dominant:
            d = summ[dominant]
            achieved = d["bytes"] / (d["ms"] * 1e-3) / 1e9
            # the GEMM family is bound by its operand streams (tall-skinny fp32 A / G, emulated-fp32 math keeps the
            # tensor pipe far from its peak): like the streaming kernels it is held against the HBM roofline
            roof = {"kernel": dominant, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": ncu_traffic(dominant), "peak_source": peak_src,
                    "bytes_per_launch": d["bytes"] / d["launches"], "us_per_launch": d["ms"] * 1e3 / d["launches"],
                    "share_of_step": d["ms"] / (ms_step * args.steps),
                    "launches_per_step": d["launches"] / args.steps,
                    "runner_up": {k: {"gb_s": own[k]["bytes"] / (own[k]["ms"] * 1e-3) / 1e9, "frac": own[k]["bytes"] / (own[k]["ms"] * 1e-3) / 1e9 / peak,
                                      "ms_per_step": own[k]["ms"] / args.steps}
                                  for k in sorted(own, key=lambda k: -own[k]["ms"])[1:4]},
                    "timed_in": "instrumented eager pass of the same step: CUDA events around each launch of our kernels, "
                                "a GPU-side delay queued before each pair keeps host launch gaps out of the interval"}
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # the CPU leg is an N = 1 figure (rank 0's host cores)
            torch.set_num_threads(host_threads())
            params, cfg, cpos, cbatch, cz, ctgt, cgraphs, cedges = cpu_sample(args.ref_graphs)
            oracle_step(params, cfg, cpos, cbatch, cz, ctgt, cgraphs)
            t0 = time.perf_counter()
            reps = 0
            while reps < 2 or time.perf_counter() - t0 < 10.0:
                oracle_step(params, cfg, cpos, cbatch, cz, ctgt, cgraphs)
                reps += 1
                if time.perf_counter() - t0 > 30.0:
                    break
            dt = (time.perf_counter() - t0) / reps
            cpu = {"value": cedges / dt, "unit": "edges/s", "cores": torch.get_num_threads(), "kind": "port",
                   "sample": f"{cgraphs} of the {N_GRAPHS} molecules ({cedges} edges), oracle op chain fwd+bwd x{reps}, torch CPU fp32"}
        line = {
            "metric": METRIC, "value": edges_total / (ms_step * 1e-3), "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "edges_per_step": edges_total, "atoms_per_rank": int(pos.shape[0]),
                       "parallelism": f"dp{world}", "l2": f"inputs larger than L2: {mem_gb:.2f} GB of activations per step",
                       "dropout": "attention-weight dropout p = 0 (configuration default 0.2, an [E, 4] mask): the only "
                                  "stochastic op of the step, off so that the run is comparable with the parity tests",
                       "gemm": ("tcgen05, fp32-accurate: hand-written 3xTF32 kernels (edge-level forward / dgrad / wgrad, "
                                "node-level wgrad; A through TMEM for outputs of <= 128 / <= 64 columns) + cuBLAS for the "
                                "node-level forward and data-gradient products"
                                if ops.gemm_backend() == "cutlass" else "cuBLAS SGEMM fp32 (allow_tf32=False)"),
                       "launch": ("CUDA-graph replay of forward+loss+backward per (atoms, edges) signature; neighbour "
                                  "search, all-reduce and AdamW eager") if args.graph else "eager",
                       "eager_ms_per_step": ms_eager},
            "e2e": {"value": edges_total / (ms_e2e * 1e-3), "unit": "edges/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4},
            "gpu_launches": profile.launches,   # our kernels launched in the instrumented pass (same count per replay)
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "kernels": kernels,
            "loss": last_loss, "grad_bucket_bytes": bucket.nbytes,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-graphs", type=int, default=8, help="molecules in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", dest="graph", action="store_true", default=True)
    ap.add_argument("--no-graph", dest="graph", action="store_false")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
