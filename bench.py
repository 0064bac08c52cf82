#!/usr/bin/env python
"""bench.py -- MP atoms/sec on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c1]

Workload (default ``c3`` = BASELINE.json configs[2], the configuration the
metric is quoted on): per GPU a batch of 1024 synthetic signals of 2^15
samples, a 4096-atom x 2048-sample dictionary, 512 greedy iterations.  One
"step" is one complete pursuit of the batch (first full correlation pass + 512
iterations) = 524 288 atoms per GPU.  Batches shard across ranks with no
data-path collective (weak scaling: every rank codes its own batch).

Prints ONE JSON line (rank 0).  ``value`` is device-resident throughput,
``e2e`` the same metric through the host-buffer C-ABI entry
(mpb200_sparse_code_host: pinned host signals in, events + residual out, copies
inside the timed region).  ``roofline`` is for the dominant kernel (the window
re-correlation), timed live with CUDA events on its stream.  ``cpu_baseline``
is the CPU oracle (port of the reference path) on a bounded sample.
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
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (batch per GPU, n_samples, n_atoms, atom_size, n_steps, description)
    "c3": (1024, 2 ** 15, 4096, 2048, 512,
           "BASELINE configs[2]: batch 1024 x 2^15 samples, 4096-atom x 2048 dictionary, 512 iterations"),
    "c2": (64, 2 ** 15, 512, 1024, 256,
           "BASELINE configs[1]: batch 64 x 2^15 samples, 512-atom x 1024 dictionary, 256 iterations"),
    "c1": (1, 2 ** 15, 512, 512, 32,
           "BASELINE configs[0]: 1 x 2^15 samples, 512 atoms x 512 samples, 32 iterations"),
    # multi-band codec (6 independent per-band pursuits around an FFT band split): handled by run_multiband()
    "c4": (16, 2 ** 16, 1024, 128, 64,
           "BASELINE configs[3]: multi-band dictionary (modules/multibanddict.py), 6 bands (2048..65536 samples) x "
           "1024 atoms x 128 samples, per-band MP on 2^16-sample signals; batch 16 and 64 iterations per band "
           "(experiments/e_2024_4_24/experiment.py:28-40)"),
    # atom-sharded (not batch-sharded): handled by run_atom_sharded()
    "c5": (1, 2 ** 20, 16384, 2048, 2048,
           "BASELINE configs[4]: single 2^20-sample signal, 16384-atom x 2048 dictionary, 2048 iterations, "
           "atom-sharded with a per-step NCCL all-gather of 16-byte records"),
}
METRIC = "MP atoms/sec (4096x2048 dict, 2^15 sig)"
METRIC_BY_WORKLOAD = {"c2": "MP atoms/sec (512x1024 dict, 2^15 sig)", "c1": "MP atoms/sec (512x512 dict, 2^15 sig)"}
UNIT = "atoms/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch (marks the run non-standard)")
    ap.add_argument("--iterations", type=int, default=0, help="override the iteration count (non-standard)")
    ap.add_argument("--mode", default="auto")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="c5 only: winner exchange fused into the pursuit over peer memory, or per-step NCCL all-gather")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--max-seconds", type=float, default=700.0,
                    help="wall-clock budget of the whole command: the W warm-up and K timed steps are always run as "
                         "asked; the legs after them (end-to-end steps, CPU sample, atom-sharded sub-record, strong-"
                         "scaling leg) shrink to what is left")
    ap.add_argument("--e2e-steps", type=int, default=2, help="steps of the end-to-end leg (at most --steps)")
    ap.add_argument("--no-atom-sharded", action="store_true", help="skip the configs[4] sub-record")
    return ap.parse_args()


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,clocks.mem")

    def __init__(self, index: int):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 6:
                self.samples.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, watts, mem = [], None, set(), [], []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for p in self.samples:
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except ValueError:
                continue
            for name, flag in zip(names, p[2:6]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
            try:                                       # board power and memory clock beside the SM clock: the headline
                watts.append(float(p[6]))              # kernel runs at the board's power cap
                mem.append(float(p[7]))
            except (ValueError, IndexError):
                pass
        sm.sort()
        watts.sort()
        mem.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": watts[len(watts) // 2] if watts else None,
                "mem_mhz": mem[len(mem) // 2] if mem else None}


# --------------------------------------------------------------------------
# synthetic inputs
# --------------------------------------------------------------------------
def make_inputs(torch, mpb, dev, batch, n, k, a, n_events, seed):
    """Planted-atom signals (SURVEY.md 8d family P), built on the device with the
    library's own decode kernel: sum of n_events unit atoms at uniform positions,
    amplitudes U(0.5,1), plus N(0, 0.01^2) noise, max-normed per signal."""
    g = torch.Generator(device="cpu").manual_seed(0)
    d = torch.zeros(k, a).uniform_(-1, 1, generator=g)
    d_dev = mpb.unit_norm(d.to(dev))
    g = torch.Generator(device="cpu").manual_seed(seed)
    ev = batch * n_events
    atom = torch.randint(0, k, (ev,), generator=g)
    pos = torch.randint(0, max(1, n - a + 1), (ev,), generator=g)
    amp = torch.zeros(ev).uniform_(0.5, 1.0, generator=g)
    rows = torch.arange(batch).repeat_interleave(n_events)
    sig = torch.zeros(batch, n, device=dev)
    mpb.scatter_add(sig, d_dev, atom.to(dev), rows.to(dev), pos.to(dev), amp.to(dev))
    gd = torch.Generator(device=dev).manual_seed(seed)
    sig += 0.01 * torch.randn(batch, n, device=dev, generator=gd)
    sig /= sig.abs().amax(dim=-1, keepdim=True) + 1e-8
    return d, sig


# --------------------------------------------------------------------------
# CPU baseline (oracle port of the reference path)
# --------------------------------------------------------------------------
def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.lower().startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_atoms_per_second(torch, n, k, a, budget_s, seed=1):
    """Times the CPU oracle (a port of modules/matchingpursuit.py::sparse_code,
    same torch kernels as the reference) on ONE signal of the workload, with
    all host threads, for both correlation forms; returns the faster."""
    from oracle import mp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = O.make_dictionary(k, a, seed=0)
    sig = O.make_planted_signals(d, 1, n, 16, seed=seed)
    best, detail = 0.0, {}
    with torch.no_grad():
        for name, kw in (("conv1d", {}), ("fft", {"approx": n})):
            t0 = time.perf_counter()
            O.greedy_pursuit(sig, d, 1, **kw)                      # warm-up step, also sizes the sample
            per_step = time.perf_counter() - t0
            steps = int(max(2, min(64, (budget_s / 2) / max(per_step, 1e-3))))
            t0 = time.perf_counter()
            O.greedy_pursuit(sig, d, steps, **kw)
            dt = time.perf_counter() - t0
            rate = steps / dt
            detail[name] = {"steps": steps, "seconds": round(dt, 3), "atoms_per_s": round(rate, 4)}
            best = max(best, rate)
    sample = (f"1 signal x {n} samples, {k}x{a} dictionary; " +
              ", ".join(f"{nm}: {v['steps']} iterations in {v['seconds']} s" for nm, v in detail.items()) +
              "; faster form quoted")
    return best, cores, sample, detail


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle
    port -- the reference is pure Python and cannot travel to the GPU box; the
    port is pinned to it by tests/golden) on the host cores, same metric."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch, n, k, a, s, desc = WORKLOADS[args.workload]
    from oracle import mp_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = O.make_dictionary(k, a, seed=0)
    sig = O.make_planted_signals(d, 1, n, 16, seed=1)
    kw = {"approx": n} if a >= 1024 else {}
    iters = 2 if k * a >= 2 ** 22 else 8       # bounded sample per step
    with torch.no_grad():
        for _ in range(args.warmup):
            O.greedy_pursuit(sig, d, 1, **kw)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            O.greedy_pursuit(sig, d, iters, **kw)
        dt = time.perf_counter() - t0
    value = args.steps * iters / dt
    sample = (f"each step = {iters} greedy iterations on 1 signal x {n} samples with the {k}x{a} dictionary "
              f"({'FFT' if kw else 'conv1d'} correlation form, {cores} threads)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC_BY_WORKLOAD.get(args.workload, METRIC), "value": value, "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (planted atoms + noise, seeded; dictionary U(-1,1) unit-normed)",
        # same keys as the repo arm's config; the CPU cannot run the whole workload (weeks), so every step is the
        # bounded sample named in cpu_baseline.sample -- of the same dictionary, signal length and metric
        "config": {"workload": desc, "batch_per_gpu": batch, "n_samples": n, "n_atoms": k, "atom_size": a,
                   "iterations": s, "sampled": {"signals": 1, "iterations_per_step": iters}},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "cpu_model": cpu_model(), "torch_threads": torch.get_num_threads()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


# --------------------------------------------------------------------------
# multi-band workload (configs[3])
# --------------------------------------------------------------------------
def run_multiband(args):
    """6 band dictionaries, FFT band split on the device, one pursuit per band.  `value` times the split and
    the six array-level pursuits with device-resident inputs; `e2e` times MultibandDictionaryLearning.encode --
    the call a user of modules/multibanddict.py makes (:399-404) -- with HOST signals in and the reference's
    per-event tuples out.  Batches shard across ranks without communication."""
    import torch
    import torch.distributed as dist
    import matching_pursuit_b200 as mpb
    from matching_pursuit_b200 import decompose as mdec
    from oracle import mp_oracle as O

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    batch, n, k, a, s, desc = WORKLOADS["c4"]
    if args.batch:
        batch = args.batch
    if args.iterations:
        s = args.iterations
    sizes = [2048 * 2 ** i for i in range(6)]
    specs = [mpb.BandSpec(sz, k, a, device=dev, signal_samples=n, is_lowest_band=(i == 0)) for i, sz in enumerate(sizes)]
    for i, spec in enumerate(specs):
        spec.d = O.make_dictionary(k, a, seed=10 + i).to(dev)
    model = mpb.MultibandDictionaryLearning(specs, n_samples=n)
    # family-P signals: planted atoms of a length-128 dictionary plus noise
    x_host = O.make_planted_signals(O.make_dictionary(k, a, seed=0), batch, n, 4 * s, seed=1 + rank).pin_memory()
    x = x_host.to(dev)

    def device_pass():
        split = mdec.fft_frequency_decompose(x, sizes[0])
        return [mpb.sparse_code_arrays(split[sz], spec.d, s) for sz, spec in zip(sizes, specs)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        device_pass()
    barrier()
    launches0 = mpb.lib().mpb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = device_pass()
    e1.record()
    barrier()
    launches = mpb.lib().mpb200_launch_count() - launches0
    model.encode(x_host, s)
    barrier()
    step_ms = []
    for _ in range(args.steps):                         # host work is part of this figure: median of the per-step
        t0 = time.perf_counter()                        # wall times, so one scheduler hiccup does not decide it
        enc = model.encode(x_host, s)
        torch.cuda.synchronize()
        step_ms.append(1e3 * (time.perf_counter() - t0))
    wall_ms = sorted(step_ms)[len(step_ms) // 2] * args.steps
    t = torch.tensor([e0.elapsed_time(e1), wall_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    atoms = batch * s * len(sizes) * world
    if rank == 0:
        modes = sorted({mpb.matchingpursuit.get_plan(k, a, sz, batch, dev, "auto").mode for sz in sizes})
        line = {
            "metric": "MP atoms/sec (6 bands x 1024x128 dict, 2^16 sig)", "value": atoms * args.steps / (float(t[0]) / 1e3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": float(t[0]) / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (planted atoms + noise, seeded)",
            "config": {"workload": desc, "batch_per_gpu": batch, "iterations_per_band": s, "band_sizes": sizes,
                       "modes": modes, "parallelism": f"batch-sharded x{world}, no collective"},
            "gpu_launches": int(launches),
            "e2e": {"value": atoms * args.steps / (float(t[1]) / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": batch * n * 4 * world,
                    "d2h_bytes_per_step": int(sum(len(enc[sz][0]) for sz in sizes)) * (a * 4 + 24) * world,
                    "api": "MultibandDictionaryLearning.encode (host signals in, reference event tuples out)"},
        }
        if not args.no_cpu_baseline and world == 1:
            torch.set_num_threads(os.cpu_count() or 1)
            ob = O.MultibandOracle([O.BandOracle(sz, O.make_dictionary(k, a, seed=10 + i), is_lowest_band=(i == 0))
                                    for i, sz in enumerate(sizes)], n)
            xs, cs = x_host[:1], max(2, s // 16)
            t0 = time.perf_counter()
            with torch.no_grad():
                ob.encode(xs, cs)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": cs * len(sizes) / dt, "unit": UNIT, "cores": os.cpu_count() or 1,
                                    "kind": "port", "sample": f"1 signal, {cs} iterations per band, all 6 bands, "
                                                              f"{round(dt, 2)} s"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------
# atom-sharded workload (configs[4])
# --------------------------------------------------------------------------
def measure_atom_sharded(torch, dist, mpb, dev, world, rank, iterations, exchange, mode, steps=1, warmup=1,
                         latency_iterations=512):
    """configs[4] on the `world` GPUs of this job: residual and dictionary replicated, rank g owns K/world atoms, one
    winner exchange per iteration.  Timed: the whole pursuit (first pass + `iterations` iterations) with the exchange
    fused into the pursuit kernels over peer memory (`exchange="p2p"`, the default) or as a host-driven per-step NCCL
    all-gather (`"nccl"`).  The per-step collective latency is ALWAYS measured on the NCCL form (CUDA events around
    every all_gather_into_tensor of one 16-byte record per rank).  Returns the record (same on every rank)."""
    from matching_pursuit_b200.distributed import AtomShardedPursuit
    batch, n, k, a, _, desc = WORKLOADS["c5"]
    s = iterations
    d, sig = make_inputs(torch, mpb, dev, batch, n, k, a, n_events=min(s, 1024), seed=1)   # same on every rank
    pursuit = AtomShardedPursuit(k, a, n, batch, device=dev, mode=mode, exchange=exchange).set_dictionary(d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        pursuit.run(sig, min(s, 64))
    barrier()
    launches0 = mpb.lib().mpb200_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        atom, pos, val, res = pursuit.run(sig, s)
    e1.record()
    barrier()
    launches = mpb.lib().mpb200_launch_count() - launches0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    timed_out = pursuit.engine.plan.exchange_timed_out() if (exchange == "p2p" and world > 1) else False
    plan_mode = pursuit.engine.plan.mode
    # separate, un-timed pass with events around every exchange: per-step collective latency of the NCCL form
    if exchange == "nccl" or world == 1:
        nccl_ref = pursuit
    else:
        nccl_ref = AtomShardedPursuit(k, a, n, batch, device=dev, mode="recorrelate", exchange="nccl").set_dictionary(d)
    nccl_ref.run(sig, min(s, latency_iterations), time_exchange=True)
    ex = sorted(nccl_ref.exchange_ms)
    if nccl_ref is not pursuit:
        nccl_ref.close()
    pursuit.close()
    # all ranks must agree on the sequence
    if world > 1:
        chk = torch.stack([atom.double().sum(), pos.double().sum(), val.double().sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        agree = bool(torch.equal(lo, hi))
    else:
        agree = True
    fused = exchange == "p2p" and world > 1
    return {
        "workload": desc, "n_gpus": world, "iterations": s, "steps": steps,
        "atoms_per_s": s * steps / (ms_total / 1e3), "ms_per_step": ms_total / steps,
        "us_per_iteration": 1e3 * ms_total / steps / s,
        "atoms_per_rank": (k + world - 1) // world, "mode": plan_mode,
        "exchange": ("fused into k_apply: 8-byte-atomic stores into peer mailboxes over NVLink" if fused else
                     "per-step NCCL all_gather_into_tensor between local_best and apply"),
        "exchange_timed_out": timed_out, "ranks_agree": agree, "gpu_launches": int(launches),
        "exchange_latency_us": {"mean": 1e3 * sum(ex) / max(len(ex), 1), "p50": 1e3 * ex[len(ex) // 2] if ex else None,
                                "p99": 1e3 * ex[min(len(ex) - 1, int(0.99 * len(ex)))] if ex else None,
                                "samples": len(ex),
                                "what": "CUDA-event time around the per-step NCCL all_gather_into_tensor of one "
                                        "16-byte record per rank (NCCL form of the exchange; at 1 GPU there is no "
                                        "collective and this is the empty event pair)"},
    }


def run_atom_sharded(args):
    import torch
    import torch.distributed as dist
    import matching_pursuit_b200 as mpb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    s = args.iterations or WORKLOADS["c5"][4]
    r = measure_atom_sharded(torch, dist, mpb, dev, world, rank, s, args.exchange, args.mode, steps=args.steps,
                             warmup=args.warmup)
    if rank == 0:
        print(json.dumps({
            "metric": "MP atoms/sec (16384x2048 dict, 2^20 sig, atom-sharded)", "value": r["atoms_per_s"],
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (planted atoms + noise, seeded)",
            "config": {"workload": r["workload"], "iterations": s, "atoms_per_rank": r["atoms_per_rank"],
                       "mode": r["mode"], "parallelism": f"atom-sharded x{world}", "exchange": r["exchange"]},
            "exchange_timed_out": r["exchange_timed_out"], "gpu_launches": r["gpu_launches"],
            "ranks_agree": r["ranks_agree"], "exchange_latency_us": r["exchange_latency_us"],
            "us_per_iteration": r["us_per_iteration"],
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measured_traffic(kernel, workload, signals_per_launch):
    """dram read+write bytes per launch of `kernel` from the committed `ncu --set full` capture
    (profiles/traffic.json: bytes per signal of one launch, measured on this workload's shapes), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t[kernel][workload]["dram_bytes_per_signal"] * signals_per_launch
    except Exception:
        return None


# --------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload == "c5":
        run_atom_sharded(args)
        return
    if args.workload == "c4":
        run_multiband(args)
        return
    import torch
    import torch.distributed as dist
    import matching_pursuit_b200 as mpb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the engine has no CPU path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    batch, n, k, a, s, desc = WORKLOADS[args.workload]
    standard = True
    if args.batch:
        batch, standard = args.batch, False
    if args.iterations:
        s, standard = args.iterations, False
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else \
        (6650.0, "fallback (B200_PROFILING.md)")

    t_start = time.perf_counter()

    def left():          # seconds of the wall-clock budget that remain
        return args.max_seconds - (time.perf_counter() - t_start)

    t_setup = time.perf_counter()
    d, sig = make_inputs(torch, mpb, dev, batch, n, k, a, n_events=min(s, 256), seed=1 + rank)
    plan = mpb.Plan(k, a, n, batch, mode=args.mode, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    plan.set_dictionary(d)
    torch.cuda.synchronize()
    dict_ms = 1e3 * (time.perf_counter() - t0)
    setup_s = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput -------------------------------------
    t_w = time.perf_counter()
    for _ in range(args.warmup):
        plan.sparse_code(sig, s, want_residual=True)
    barrier()
    step_s = (time.perf_counter() - t_w) / max(args.warmup, 1)      # host estimate of one step, for the budget only
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # Per-kernel CUDA events (two per iteration, on the plan's stream) ride inside the timed region when a step is
    # long (headline: one event per 9 ms launch); for the latency-bound workloads they would serialise the
    # programmatically dependent launches they sit between, so there the kernel times come from one extra,
    # untimed step after the region.
    events_in_region = step_s >= 0.05
    plan.timing(events_in_region)
    launches0 = mpb.lib().mpb200_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = plan.sparse_code(sig, s, want_residual=True)
    ev1.record()
    barrier()
    launches = mpb.lib().mpb200_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if not events_in_region:
        plan.timing(True)
        plan.sparse_code(sig, s, want_residual=True)
        barrier()
    kernel_times = plan.timing_read()
    ev_steps = args.steps if events_in_region else 1        # steps the per-kernel events cover
    plan.timing(False)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    atoms_per_step = batch * s * world
    value = atoms_per_step * args.steps / (ms_total / 1e3)

    # ---- end to end through the host-buffer C-ABI entry ------------------
    # Same workload, same step; host signals in, events + residual out, copies inside the timed region.  The kernels
    # are warm and the staging buffers were sized by the first call, so the leg is `e2e_steps` timed steps (at most
    # --e2e-steps, at most K, and as many as the wall-clock budget still holds: a step is a whole pursuit of the
    # batch, tens of seconds at the headline shape).  All ranks agree on the count.
    def agree_min(x):
        t_ = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_, op=dist.ReduceOp.MIN)
        return float(t_.item())

    e2e = None
    if not args.no_e2e:
        sig_host = sig.cpu().pin_memory()
        outs = (torch.empty(batch, s, dtype=torch.int32).pin_memory(),
                torch.empty(batch, s, dtype=torch.int32).pin_memory(),
                torch.empty(batch, s, dtype=torch.float32).pin_memory(),
                torch.empty(batch, n, dtype=torch.float32).pin_memory())
        reserve = 45.0                                          # CPU sample, sub-records, teardown
        fit = int(agree_min((left() - reserve) / max(step_s, 1e-3)))
        e2e_steps = max(1, min(args.e2e_steps, args.steps, fit))
        if step_s < 1.0:                                        # cheap steps: also warm the host path once
            plan.sparse_code_host(sig_host, s, out=outs)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            plan.sparse_code_host(sig_host, s, out=outs)
        e1.record()
        barrier()
        te = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        h2d = sig_host.numel() * 4
        d2h = sum(o.numel() * 4 for o in outs)
        e2e = {"value": atoms_per_step * e2e_steps / (float(te.item()) / 1e3), "unit": UNIT,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world, "steps": e2e_steps,
               "ms_per_step": float(te.item()) / e2e_steps,
               "api": "mpb200_sparse_code_host (Plan.sparse_code_host), pinned host buffers; "
                      f"{e2e_steps} timed step(s) after the {args.warmup}+{args.steps} device-resident ones"}
        del sig_host, outs
    clocks = sampler.stop() if rank == 0 else None

    # ---- strong-scaling form of the same configuration --------------------
    # BASELINE configs[2] as written: ONE batch of 1024 signals split over the N GPUs (the headline `value` above is
    # the weak form: 1024 signals per GPU).  At N = 1 the two coincide.
    strong = None
    if args.workload == "c3" and standard:
        if world == 1:
            strong = {"global_batch": batch, "batch_per_gpu": batch, "value": value, "unit": UNIT,
                      "ms_per_step": ms_total / args.steps, "steps": args.steps, "note": "identical to the weak form at 1 GPU"}
        elif agree_min(left()) > 60.0 + step_s / world:
            lo, hi = mpb.distributed.shard_batch(batch, world, rank)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.sparse_code(sig[lo:hi], s, want_residual=True)
            e1.record()
            barrier()
            ts = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            strong = {"global_batch": batch, "batch_per_gpu": hi - lo, "value": batch * s / (float(ts.item()) / 1e3),
                      "unit": UNIT, "ms_per_step": float(ts.item()), "steps": 1,
                      "note": "one batch of 1024 signals split over the ranks, no collective; max over ranks"}

    # ---- roofline inputs are read before the plan is released -------------
    info = plan.info
    plan_mode, plan_fft2, plan_resident, plan_bytes = plan.mode, plan.fft_size2, plan.resident_batch, int(plan.device_bytes)
    plan.close()
    del plan, sig, out
    torch.cuda.empty_cache()

    # ---- configs[4] sub-record: atom sharding with a per-step winner exchange ---------
    atom_sharded = None
    if args.workload == "c3" and standard and not args.no_atom_sharded and agree_min(left()) > 40.0:
        try:
            atom_sharded = measure_atom_sharded(torch, dist, mpb, dev, world, rank, 512, "p2p", "auto", steps=1, warmup=1,
                                                latency_iterations=256)
        except Exception as exc:                                 # the sub-record must never cost the headline line
            atom_sharded = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------
    m_fft, blk, nb = info.fft_size, info.block, info.n_blocks
    corr_ms, corr_n = kernel_times["recorrelate"]
    apply_ms, apply_n = kernel_times["apply"]
    first_ms, first_n = kernel_times["first_pass"]
    nvb = (2 * a - 2) // blk + 2                        # blocks refreshed per winner (worst case)
    # algorithmic bytes of ONE re-correlation launch (one iteration of the whole batch), SURVEY.md 8(d):
    #   pair spectra streamed once (shared by all signals)      8 * (K/2) * M
    #   window spectra of the batch                              8 * B * M
    #   refreshed block maxima (value + position)                8 * B * K * nvb
    #   row re-reduction over the block-max table                4 * B * K * NB  (+ 8 * B * K written)
    alg_bytes = 8 * ((k + 1) // 2) * m_fft + 8 * batch * m_fft + 8 * batch * k * nvb + 4 * batch * k * nb \
        + 8 * batch * k
    roofline = None
    gram_ms, gram_n = kernel_times["gram_update"]
    if plan_mode == "gram" and gram_n:
        # dominant kernel in GRAM mode: k_gram_update.  Algorithmic bytes per launch:
        # per signal 4*K*W (Gram row read) + 8*K*W (map window read-modify-write), W = 2A-1.
        w = 2 * a - 1
        gram_bytes = 12 * k * w * batch
        per_launch_s = gram_ms / gram_n / 1e3
        achieved = gram_bytes / per_launch_s / 1e9
        roofline = {"bound": "hbm", "kernel": "k_gram_update (map window -= v * Gram row, fused block/row maxima)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src, "traffic": None, "alg_bytes_per_launch": gram_bytes,
                    "ms_per_launch": gram_ms / gram_n, "share_of_step": gram_ms / ev_steps / (ms_total / args.steps)}
    elif plan_mode == "sgram" and gram_n:
        # dominant kernel in SGRAM mode: k_delta.  The Gram rows are synthesised from L2-resident spectra, so the
        # algorithmic HBM bytes per atom-step are the map window read-modify-write alone: 8*K*W, W = 2A-1
        # (SURVEY.md 8d "Gram incremental update" minus its 4*K*W table read).
        w = 2 * a - 1
        n_sub = -(-batch // plan_resident)
        per_launch_signals = batch / n_sub
        delta_bytes = 8 * k * w * per_launch_signals
        per_launch_s = gram_ms / gram_n / 1e3
        achieved = delta_bytes / per_launch_s / 1e9
        flops = per_launch_signals * ((k + 1) // 2) * 5.0 * plan_fft2 * (plan_fft2.bit_length() - 1)
        roofline = {"bound": "hbm", "kernel": "k_delta (Gram rows synthesised by inverse FFT of cached spectra; TMA-staged "
                                              "map window -= v * row; fused block/row maxima)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src, "traffic": measured_traffic("k_delta", args.workload, per_launch_signals),
                    "traffic_source": "NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one "
                                      "k_delta launch from the committed `ncu --set full` capture "
                                      "(profiles/traffic.json, bytes per signal) x signals per launch",
                    "alg_bytes_per_launch": delta_bytes, "signals_per_launch": per_launch_signals,
                    "ms_per_launch": gram_ms / gram_n, "share_of_step": gram_ms / ev_steps / (ms_total / args.steps),
                    "fp32": {"note": "secondary bound: nominal 5*M2*log2(M2) FLOP per inverse transform against the "
                                     "nominal (unmeasured) FP32 FMA peak",
                             "achieved_tflops": flops / per_launch_s / 1e12,
                             "nominal_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}}
    elif corr_n and apply_n == 0 and s > 1:
        # the whole loop ran as ONE cooperative launch per resident batch (k_pursue_fused): latency bound by design --
        # one grid barrier and a handful of L2 round trips per iteration; the figure that matters is the time per iteration
        per_launch_s = corr_ms / corr_n / 1e3
        iters = s - 1                                            # iterations that re-correlate (the last one only subtracts)
        achieved = alg_bytes * iters / per_launch_s / 1e9
        roofline = {"bound": "hbm", "kernel": "k_pursue_fused (one cooperative launch per pursuit: winner selection, residual "
                                              "window transform, spectrum product + inverse FFT, block/row maxima, one grid "
                                              "barrier per iteration)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src, "traffic": None, "alg_bytes_per_launch": alg_bytes * iters,
                    "ms_per_launch": corr_ms / corr_n, "us_per_iteration": 1e6 * per_launch_s / iters,
                    "share_of_step": corr_ms / ev_steps / (ms_total / args.steps),
                    "note": "latency bound, not bandwidth bound: the fraction is reported for completeness"}
    elif corr_n:
        per_launch_s = corr_ms / corr_n / 1e3
        achieved = alg_bytes / per_launch_s / 1e9
        flops = batch * ((k + 1) // 2) * 5.0 * m_fft * (m_fft.bit_length() - 1)   # 5 M log2 M per complex IFFT
        roofline = {"bound": "hbm", "kernel": "k_corr (window re-correlation: fused spectrum product + inverse FFT "
                                              "+ block/row maxima)",
                    "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                    "peak_source": peak_src, "traffic": None,
                    "alg_bytes_per_launch": alg_bytes, "ms_per_launch": corr_ms / corr_n,
                    "share_of_step": corr_ms / ev_steps / (ms_total / args.steps),
                    "fp32": {"note": "the kernel is FP32-pipe bound, not HBM bound: nominal 5*M*log2(M) FLOP per "
                                     "inverse transform against the nominal (unmeasured) FP32 FMA peak",
                             "achieved_tflops": flops / per_launch_s / 1e12,
                             "nominal_peak_tflops": 148 * 128 * 2 * 1.965e9 / 1e12}}
    line = {
        "metric": METRIC_BY_WORKLOAD.get(args.workload, METRIC), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (planted atoms + noise, seeded; dictionary U(-1,1) unit-normed)",
        "config": {"workload": desc if standard else f"NON-STANDARD batch={batch} iterations={s} of: {desc}",
                   "batch_per_gpu": batch, "n_samples": n, "n_atoms": k, "atom_size": a, "iterations": s,
                   "mode": plan_mode, "fft_size": m_fft, "fft_size2": plan_fft2, "block": blk,
                   "resident_batch": plan_resident,
                   "parallelism": f"batch-sharded x{world}, no collective",
                   "l2": f"inputs larger than L2: plan working set {plan_bytes >> 20} MiB + signals "
                         f"{batch * n * 4 >> 20} MiB"},
        "gpu_launches": int(launches), "clocks": clocks,
        "kernel_ms": {"measured": "CUDA events inside the timed region" if events_in_region else
                                  "CUDA events in one extra step after the timed region",
                      "first_pass": first_ms / max(first_n, 1), "apply_per_iteration": apply_ms / max(apply_n, 1),
                      "recorrelate_per_iteration": corr_ms / max(corr_n, 1),
                      "gram_update_per_iteration": gram_ms / max(gram_n, 1)},
        "setup": {"dictionary_tables_ms": dict_ms, "inputs_and_plan_s": setup_s,
                  "plan_device_bytes": plan_bytes},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if strong is not None:
        line["strong_scaling"] = strong
    if atom_sharded is not None:
        line["atom_sharded"] = atom_sharded
    if roofline is not None:
        line["roofline"] = roofline
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, detail = cpu_atoms_per_second(torch, n, k, a, max(4.0, min(args.cpu_seconds, left() - 10.0)))
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "cpu_model": cpu_model(), "torch_threads": torch.get_num_threads()}
    line["wall_s"] = round(time.perf_counter() - t_start, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
