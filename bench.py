#!/usr/bin/env python
"""bench.py -- coefficient-sets/s of the hot path on N B200s.

    python bench.py --gpus N --steps K --warmup W            # the CUDA path
    python bench.py --impl reference --gpus N ...            # the reference's CPU algorithm

A "step" is one pass of the hot path (normalisation + Symphony + Heyvaerts, all
eight coefficients) over one batch of synthetic pitchy power-law points, the
crank-out-pitchypl workload BASELINE.json's metric is quoted on.  The batch per
GPU is fixed (weak scaling): the full 10 M-point configuration is --points
10000000.  One process per GPU; points are independent, so ranks exchange
nothing on the data path (torch.distributed only carries the barrier and the
max-over-ranks of the timing).

Printed JSON (rank 0): value = whole-job sets/s with inputs resident in HBM;
e2e = the same through the public C ABI with host buffers (H2D and D2H copies
inside the timed region); roofline = FP64 work of the dominant kernel against the
measured FP64 FMA peak; cpu_baseline = the oracle on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "8-coefficient sets/sec"
UNIT = "coefficient-sets/s"
SEED = 20260

# FP64 floating-point operations executed per Gauss-Kronrod application (one pass of the 32 lanes
# over 31 nodes, or over two 15-node panels in the Symphony gamma integral; FMA = 2): the
# predicated-on thread counts of DFMA, DMUL and DADD on the SASS page of one ncu --set full launch
# of each product kernel, divided by the applications that launch counted
# (profiles/r01_fast_kernels_details.txt and r01_fast_kernels_summary.md: 4096 seeded pitchy
# power-law points, 20.1 / 93.3 ms).
FLOP_PER_APPLICATION = {"symphony": 39.49e3, "heyvaerts": 28.55e3}
# DRAM bytes (read + write) per point of the same captures (register spills to local memory; the
# algorithmic traffic is ~110 B per point): the path does not touch HBM.
DRAM_BYTES_PER_POINT = {"symphony": (1.29e6 + 45.68e6) / 4096, "heyvaerts": (2.44e6 + 40.55e6) / 4096}

def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--points", type=int, default=262144,
                    help="points per GPU per step (the full BASELINE configs[2] batch is 10000000)")
    ap.add_argument("--config", default="pitchy_pl", choices=["pitchy_pl", "powerlaw", "pitchy_kappa"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="points of the CPU sample (0 = auto, ~20 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(config, points):
    return (f"{config} (crank-out-pitchypl shape: BASELINE configs[2] point distribution, seeded), {points} points per "
            "GPU per step (a slice of the 10 M-point batch; --points 10000000 runs all of it), all 8 coefficients, "
            "mode=fast")


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def max_over_ranks(values, world, device):
    """Timing of a multi-GPU step = the slowest rank (one all-reduce MAX of a tiny tensor)."""
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def whole_job_value(points_per_rank, world, steps, ms):
    """Units all ranks processed / max-over-ranks time (weak scaling: per-GPU batch fixed)."""
    return points_per_rank * world * steps / (ms * 1e-3)


def host_threads():
    """Every host core this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers;
    the CPU arm must not inherit that, it is asked for all the host threads it can use."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(config, n_sample, kind_label):
    """The oracle (the reference's algorithm; its own bessel.c) on the host cores."""
    from oracle import oracle as O
    from rimphony_b200.sampler import synthetic_batch

    threads = host_threads()
    if n_sample <= 0:
        n_sample = max(16, 12 * threads)  # ~1 s per point per core => ~15-25 s
    kind, s, theta, params = synthetic_batch(config, n_sample, seed=SEED)
    t0 = time.perf_counter()
    O.batch(kind, s, theta, params, n_threads=threads)
    dt = time.perf_counter() - t0
    return {"value": n_sample / dt, "unit": UNIT, "cores": threads, "kind": kind_label,
            "sample": f"first {n_sample} points of the seeded {config} batch, one point per OpenMP thread, "
                      f"{dt:.1f} s wall"}, dt


def run_reference(args):
    """--impl reference: the reference's CPU algorithm.  rimphony itself cannot be built in
    this image (no Rust toolchain, no GSL), so this is the oracle port: the restated
    symphony.rs / heyvaerts.rs control flow driving the reference's own bessel.c compiled
    in place (oracle/_ref).  Rank 0 alone runs; other ranks exit 0."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle as O
    from rimphony_b200.sampler import synthetic_batch

    threads = host_threads()
    n_sample = args.cpu_sample if args.cpu_sample > 0 else max(8, 4 * threads)
    kind, s, theta, params = synthetic_batch(args.config, n_sample, seed=SEED)
    for _ in range(min(args.warmup, 1)):
        O.batch(kind, s[: max(threads, 1)], theta[: max(threads, 1)],
                [np.asarray(p)[: max(threads, 1)] if np.ndim(p) else p for p in params], n_threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.batch(kind, s, theta, params, n_threads=threads)
    dt = time.perf_counter() - t0
    value = args.steps * n_sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.points),
                   "note": "CPU reference arm: each step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{n_sample} points per step x {args.steps} steps of the seeded {args.config} batch"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import rimphony_b200 as R
    from rimphony_b200.sampler import synthetic_batch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.points
    kind, s, theta, params = synthetic_batch(args.config, n, seed=SEED, shard=rank)
    n_params = len(params)

    # --- device-resident leg -------------------------------------------------
    d_s = torch.from_numpy(s).to(dev)
    d_theta = torch.from_numpy(theta).to(dev)
    d_params = [torch.from_numpy(np.atleast_1d(np.asarray(p, dtype=np.float64))).to(dev) for p in params]
    d_out = torch.empty(8 * n, dtype=torch.float64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # a real (non-default) stream: the library launches on the stream it is handed, and the
    # CUDA events below must be recorded on that same stream to see the kernels
    stream = torch.cuda.Stream(device=dev)

    def step_device():
        with torch.cuda.stream(stream):
            flush.fill_(1)
            R.compute_all_dimensionless_device(kind, d_s, d_theta, d_params, d_out, d_status, stream=stream,
                                               synchronize=False)

    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step_device()
    barrier()
    launches0 = R.kernel_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = R.kernel_launch_count() - launches0
    ms_dev = e0.elapsed_time(e1)

    # per-kernel device time and work counters of one more (untimed) step, for the roofline
    res = R.compute_all_dimensionless_batch(kind, s, theta, params, device=local_rank, extras=True)
    k_ms = res.kernel_ms
    apps_sym = float(res.counters[0].astype(np.float64).sum())
    apps_hey = float(res.counters[1].astype(np.float64).sum())
    nan_rate = float(np.isnan(res.values).any(axis=0).mean())

    # --- end-to-end leg: host buffers through the public C ABI --------------
    h_s = torch.from_numpy(s).pin_memory().numpy()
    h_theta = torch.from_numpy(theta).pin_memory().numpy()
    h_params = [torch.from_numpy(np.ascontiguousarray(p)).pin_memory().numpy() if np.ndim(p) else p for p in params]

    def step_host():
        flush.fill_(1)
        return R.compute_all_dimensionless_batch(kind, h_s, h_theta, h_params, device=local_rank)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step_host()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    h2d = 8 * n * (2 + sum(1 for p in params if np.ndim(p))) + 8 * sum(1 for p in params if not np.ndim(p))
    d2h = 8 * 8 * n + 4 * n
    del out

    ms_dev, ms_e2e = max_over_ranks([ms_dev, t_e2e * 1e3], world, dev)

    if rank == 0:
        value = whole_job_value(n, world, args.steps, ms_dev)
        e2e_value = whole_job_value(n, world, args.steps, ms_e2e)
        peak = R.fp64_peak_tflops(local_rank)
        # The two product kernels run concurrently on two streams and share the SMs, so the roofline
        # is taken over both: FP64 flops executed by the step's Symphony + Heyvaerts launches (rule
        # applications counted by the kernels x flops per application from ncu) over the CUDA-event
        # span of the step; `dominant` names the larger contributor.
        flops_sym = apps_sym * FLOP_PER_APPLICATION["symphony"]
        flops_hey = apps_hey * FLOP_PER_APPLICATION["heyvaerts"]
        dominant = "heyvaerts" if flops_hey >= flops_sym else "symphony"
        achieved = (flops_sym + flops_hey) / (k_ms[3] * 1e-3) * 1e-12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, n), "seed": SEED, "l2": "flushed between steps (256 MB write)",
                       "nan_rate": nan_rate, "sharding": f"{world} independent shards, no data-path collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "fp64", "kernel": "k_symphony_fast + k_heyvaerts_fast (concurrent; larger share: " + dominant + ")",
                         "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None,
                         "traffic": (DRAM_BYTES_PER_POINT["symphony"] + DRAM_BYTES_PER_POINT["heyvaerts"]) * n,
                         "traffic_note": "bytes per launch, scaled per point from the ncu --set full capture in profiles/",
                         "peak_source": "measured in this run: register-resident DFMA kernel "
                                        "(MEASURED_PEAKS.json has no FP64 figure)",
                         "kernel_ms": {"normalize": k_ms[0], "symphony": k_ms[1], "heyvaerts": k_ms[2], "span": k_ms[3]},
                         "gk31_applications_per_point": {"symphony": apps_sym / n, "heyvaerts": apps_hey / n}},
        }
        if not args.no_cpu_baseline and world == 1:
            cb, _ = cpu_baseline(args.config, args.cpu_sample, "port")
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
