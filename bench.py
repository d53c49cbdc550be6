#!/usr/bin/env python
"""bench.py — headline benchmark of the BFGS hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batched solve: 2^16 independent n=60 Rosenbrock problems from
N(0,1) starts (BASELINE.json configs[1]) per GPU, default BackTracking (order 2),
tol 1e-8, fp64.  With N GPUs every rank solves its own 2^16 problems (the batch
shards with no communication; weak scaling) and `value` is the whole job's
converged solves/s, timed on the device (CUDA events, max over ranks).

Prints ONE JSON line (rank 0).  See DESIGN.md §7 for every field.
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

N_DIM = 60
BATCH = 1 << 16
SEED = 20261018
METRIC = "converged BFGS solves/sec (n=60 Rosenbrock, 2^16 batch per GPU)"
# DRAM bytes of ONE launch of the headline kernel at this workload (dram__bytes_read.sum +
# dram__bytes_write.sum, ncu --set full on this command): profiles/r1_tile60_tmem_bench_ncu.txt for
# the kernel with the tiles in tensor memory (CTAs of 128 threads), profiles/r1_tile60_bench_ncu.txt
# for the all-register kernel (CTAs of 64 threads)
TILE_KERNEL_DRAM_BYTES_PER_LAUNCH = {128: 32798464 + 14649600, 64: 33576192 + 32338944}
UNIT = "solves/s"


def make_starts(rank, world, batch=BATCH, n=N_DIM):
    """Θ for this rank: column j of the rank's matrix is problem rank + j*world of the
    global Philox(key=20261018) N(0,1) stream (interleaved sharding)."""
    rng = np.random.Generator(np.random.Philox(key=SEED + rank))
    return rng.standard_normal((batch, n)).T.copy(order="F")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = str(gpu_index)
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", self.gpu, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_run(steps, warmup, sample):
    """The reference's CPU path.  Julia cannot run here (no runtime in the image), so
    this times the C restatement (oracle/, kind "port") with a Threads.@threads-style
    pool on all host cores over a bounded sample of the same workload."""
    from oracle import binding as ob
    ob.build()
    Theta = make_starts(0, 1, batch=sample)
    model = ob.rosenbrock()
    cores = ob.num_cores()
    times, conv, iters = [], 0, 0
    for s in range(warmup + steps):
        t = time.perf_counter()
        r = ob.optimize_batch(model, Theta, nthreads=cores, want_xg=True)
        dt = time.perf_counter() - t
        if s >= warmup:
            times.append(dt)
            conv = r["converged"]
            iters = int(r["iters"].sum())
    dt = float(np.mean(times))
    # single-problem latency, BenchmarkTools-style, beside README.md:80 (330.825 µs)
    lat = []
    for _ in range(200):
        t = time.perf_counter()
        ob.optimize(model, Theta[:, 0])
        lat.append(time.perf_counter() - t)
    return dict(solves_per_s=conv / dt, iters_per_s=iters / dt, ms_per_step=dt * 1e3, cores=cores,
                sample=sample, single_problem_us_median=float(np.median(lat) * 1e6),
                mean_iters=iters / max(conv, 1))


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    sample = args.cpu_sample
    r = cpu_reference_run(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["solves_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "batched Rosenbrock n=60, N(0,1) starts, BackTracking order 2, tol 1e-8 "
                               "(BASELINE.json configs[1]); each step = %d-problem sample" % sample},
        "cpu_baseline": {"value": r["solves_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": "%d of 65536 problems per step, all host cores (pthread pool over problems)" % sample,
                         "single_problem_us_median": r["single_problem_us_median"],
                         "iters_per_s": r["iters_per_s"],
                         "note": "C restatement of the reference (oracle/qn_oracle.c); the Julia reference cannot run in this image"},
        "e2e": {"value": r["solves_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import torch
    import qn_b200 as qn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    n, batch = N_DIM, args.batch
    eng = qn.Engine(local_rank)
    obj = qn.Rosenbrock()
    ls = qn.BackTracking()
    flags = qn._lib.FLAG_FORCE_SMEM if args.path == "smem" else 0

    Theta_h = make_starts(rank, world, batch=batch)
    # pinned host buffers for the end-to-end leg (host API, copies inside the timed region)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    th_pin = pin(Theta_h.T.copy())           # (batch, n) C-order == n×batch column-major
    x_pin = torch.empty((batch, n), dtype=torch.float64).pin_memory()
    g_pin = torch.empty((batch, n), dtype=torch.float64).pin_memory()
    ell_pin = torch.empty(batch, dtype=torch.float64).pin_memory()
    it_pin = torch.empty(batch, dtype=torch.int32).pin_memory()
    fe_pin = torch.empty(batch, dtype=torch.int32).pin_memory()
    st_pin = torch.empty(batch, dtype=torch.int32).pin_memory()

    # device-resident buffers for the kernel-only leg
    th_d = th_pin.to(dev)
    x_d = torch.empty_like(th_d)
    g_d = torch.empty_like(th_d)
    ell_d = torch.empty(batch, dtype=torch.float64, device=dev)
    it_d = torch.empty(batch, dtype=torch.int32, device=dev)
    fe_d = torch.empty(batch, dtype=torch.int32, device=dev)
    st_d = torch.empty(batch, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    stream = torch.cuda.current_stream()

    def step_device():
        eng.optimize_batched_device(obj, n, batch, th_d.data_ptr(), ls=ls, x_ptr=x_d.data_ptr(),
                                    g_ptr=g_d.data_ptr(), ell_ptr=ell_d.data_ptr(), iters_ptr=it_d.data_ptr(),
                                    fevals_ptr=fe_d.data_ptr(), status_ptr=st_d.data_ptr(), flags=flags,
                                    stream=stream.cuda_stream)

    import ctypes as C
    lsc = ls._c()
    handle = eng.model_handle(obj)

    def step_host():
        qn._lib.check(eng._lib.qn_bfgs_optimize_batched_f64(
            eng._ctx, handle, n, batch, th_pin.data_ptr(), C.byref(lsc), 1e-8, 10000, None, 0, 0,
            x_pin.data_ptr(), g_pin.data_ptr(), ell_pin.data_ptr(), it_pin.data_ptr(), fe_pin.data_ptr(),
            st_pin.data_ptr(), 0, None, flags))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- roofline denominators, measured live on this GPU (rank 0 only prints them)
    mb = eng.microbench() if rank == 0 else None

    # ---- kernel-only leg: inputs resident in HBM, CUDA events on the launch stream
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xff)          # L2 flush between timed steps (outside the events)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_dev = max_over_ranks(float(np.sum(ms_steps)))
    conv_local = int((st_d == 0).sum().item())
    iters_local = int(it_d.sum().item())
    fevals_local = int(fe_d.sum().item())
    conv_total = sum_over_ranks(conv_local)
    iters_total = sum_over_ranks(iters_local)
    launch = eng.last_launch()

    # ---- end-to-end leg: the public host API, pinned host buffers, copies inside
    for _ in range(max(1, args.warmup // 2)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    t_e2e_local = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e_local)
    conv_e2e = sum_over_ranks(int((st_pin == 0).sum().item()))
    # identical results through both entry points
    same = bool(torch.equal(x_pin, x_d.cpu()) and torch.equal(it_pin, it_d.cpu()))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(1, 0, args.cpu_sample)

    if dist is not None:
        dist.barrier()
    if rank == 0:
        K = args.steps
        sec = ms_dev * 1e-3
        value = conv_total * K / sec
        its_per_s = iters_total * K / sec
        flops_alg = 8.0 * n * n            # 8 n² flops per BFGS iteration
        bytes_alg = 16.0 * n * n + 96.0 * n  # 16 n² + 96 n bytes per iteration
        per_gpu_its = its_per_s / world
        ach_tf = per_gpu_its * flops_alg / 1e12
        ach_gbs = per_gpu_its * bytes_alg / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "batched Rosenbrock n=60, 2^16 N(0,1) starts per GPU, BackTracking order 2, "
                                   "tol 1e-8 (BASELINE.json configs[1])",
                       "n": n, "per_gpu_batch": batch, "global_batch": batch * world,
                       "path": ["register-resident", "smem-resident", "stream"][launch["path"]],
                       "tile_home": ("tensor memory (tcgen05.ld/st), 8 problems per SM" if launch["block"] == 128
                                     else "register file, 6 problems per SM") if launch["path"] == 0 else None,
                       "grid": launch["grid"], "block": launch["block"], "smem_bytes": launch["smem_bytes"],
                       "l2": "256 MB buffer written between timed steps (L2 flush, outside the events)",
                       "parallelism": "independent problems sharded over %d GPU(s), no collective" % world},
            "iterations_per_s": its_per_s,
            "mean_iterations_per_solve": iters_total / max(conv_total, 1),
            "fevals_per_iteration": fevals_local / max(iters_local, 1),
            "converged": int(conv_total), "problems": batch * world,
            "wall_ms_total": t_wall * 1e3,
            "e2e": {"value": conv_e2e * K / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(th_pin.numel() * 8),
                    "d2h_bytes_per_step": int((x_pin.numel() + g_pin.numel() + ell_pin.numel()) * 8 + 3 * 4 * batch),
                    "ms_per_step": t_e2e / K * 1e3, "same_results_as_device_leg": same,
                    "api": "qn_bfgs_optimize_batched_f64 (host pointers, pinned)"},
            "gpu_launches": K * launch["kernel_launches"] * world,
            "roofline": {
                "bound": "fp64", "achieved": ach_tf, "peak": mb["dfma"], "unit": "TFLOP/s",
                "frac": ach_tf / mb["dfma"],
                "traffic": TILE_KERNEL_DRAM_BYTES_PER_LAUNCH.get(launch["block"]) if (batch == BATCH and launch["path"] == 0) else None,
                "traffic_note": "DRAM bytes per launch from profiles/r1_tile60_tmem_bench_ncu.txt (r1_tile60_bench_ncu.txt for the "
                                "all-register kernel); the algorithmic I/O of a launch is 31.5 MB of starts in, 64.2 MB of results "
                                "out (part of which is still in L2 when the kernel ends) — B^-1 never touches HBM",
                "pipes": "ncu on this launch: FP64 pipe 51 % busy, LSU data pipe 56 %, issue slots 61 %, 4 warps per scheduler "
                         "(profiles/r1_tile60_tmem_bench_ncu.txt) — no single pipe is the bound; see DESIGN.md 5.1",
                "peak_source": "qn_microbench_dfma, measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry)",
                "algorithmic": "8 n^2 = %d flops per BFGS iteration x iterations, per GPU" % int(flops_alg),
                "smem_equiv": {"achieved": ach_gbs, "peak": mb["lds"], "unit": "GB/s", "frac": ach_gbs / mb["lds"],
                               "algorithmic": "16 n^2 + 96 n = %d bytes per iteration" % int(bytes_alg),
                               "peak_source": "qn_microbench_lds (conflict-free LDS.128), measured live"},
                "hbm_measured_gbs": mb["hbm"],
            },
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = {
                "value": cpu["solves_per_s"], "unit": UNIT, "cores": cpu["cores"], "kind": "port",
                "sample": "%d of 65536 problems, all host cores (pthread pool over problems)" % cpu["sample"],
                "single_problem_us_median": cpu["single_problem_us_median"], "iters_per_s": cpu["iters_per_s"],
                "note": "C restatement of the reference (oracle/qn_oracle.c); the Julia reference cannot run in this image"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--path", default="auto", choices=["auto", "smem"])
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        return run_reference_arm(args, rank)
    if world == 1 and args.gpus > 1:
        # the driver launches N>1 through torchrun; a bare `--gpus N` re-launches itself that way
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        return subprocess.call(cmd)
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
