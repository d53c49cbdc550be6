#!/usr/bin/env python
"""bench.py — headline benchmark of the BFGS hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one batched solve of BASELINE.json configs[1]: ONE global batch of 2^16 independent
n = 60 Rosenbrock problems from N(0,1) starts (Philox key 20261018), default BackTracking (order 2),
tol 1e-8, fp64.  With N GPUs the SAME 2^16 problems are sharded, problem j -> rank j mod N, with no
data-path collective (strong scaling: total work is fixed); `value` is the whole job's converged
solves/s, timed on the device (CUDA events, max over ranks).

Legs of one run (all inside this one command):
  value        kernel-only: each rank's slice resident in HBM, qn_bfgs_optimize_batched_device_f64
  e2e          host buffers, copies inside the timed region.  N = 1: qn_bfgs_optimize_batched_f64.
               N > 1: rank 0 alone calls the library's own multi-GPU entry
               qn_bfgs_optimize_batched_multi_f64 over all N GPUs (host scatter + gather inside the
               timed region, the other ranks idle); `e2e_per_rank` is the one-process-per-GPU form
               (every rank the host API on its slice, max over ranks)
  tail         one extra step with exit stamps: share of slot time after the work queue ran dry
  weak         (N > 1) every rank its own 2^16 problems, a few steps — the round-1 number, kept as a key
  cpu_baseline (N = 1) the CPU ports on a bounded sample, all host cores

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
METRIC = "converged BFGS solves/sec (n=60 Rosenbrock, 2^16 batch) at 1/2/4/8 B200"  # BASELINE.json, verbatim
UNIT = "solves/s"
WORKLOAD = ("batched Rosenbrock n=60, one global batch of 2^16 N(0,1) starts (Philox key 20261018), "
            "BackTracking order 2, tol 1e-8 (BASELINE.json configs[1])")
# DRAM bytes of ONE launch of the headline kernel on the full 2^16 batch (dram__bytes_read.sum +
# dram__bytes_write.sum, ncu --set full on this command): profiles/r2_tile60_final_bench_ncu.txt for the
# kernel with the tiles in tensor memory (CTAs of 128 threads), profiles/r1_tile60_bench_ncu.txt for
# the all-register kernel (CTAs of 64 threads)
TILE_KERNEL_DRAM_BYTES_PER_LAUNCH = {128: 31864320 + 11734016, 64: 33576192 + 32338944}


def global_starts(batch=BATCH, n=N_DIM, key=SEED):
    """The global n×batch start matrix (column j = problem j), column-major."""
    rng = np.random.Generator(np.random.Philox(key=key))
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
PORT_NOTE = ("C restatements of the reference (oracle/): the Julia reference cannot run in this image. "
             "'fast' = free summation order + contraction, what @turbo/@fastmath license (oracle/qn_fastport.c); "
             "'strict' = the canonical-order parity oracle (oracle/qn_oracle.c)")


def cpu_reference_run(steps, warmup, sample):
    """The reference's CPU path.  Julia cannot run here (no runtime in the image), so this times the C
    ports with a Threads.@threads-style pool over problems on all host cores: the FAST port (the
    number every ratio should divide by) for `warmup + steps` passes over `sample` problems, and the
    strict parity oracle once on a bounded sample beside it."""
    from oracle import binding as ob
    ob.build()
    Theta = global_starts(batch=sample)
    cores = ob.num_cores()
    times, conv, iters = [], 0, 0
    for s in range(warmup + steps):
        t = time.perf_counter()
        r = ob.fast_optimize_batch(Theta, nthreads=cores, want_xg=True)
        dt = time.perf_counter() - t
        if s >= warmup:
            times.append(dt)
            conv = r["converged"]
            iters = int(r["iters"].sum())
    dt = float(np.mean(times))
    # the strict oracle on at most 8192 problems, once
    ns = min(sample, 8192)
    t = time.perf_counter()
    rs = ob.optimize_batch(ob.rosenbrock(), Theta[:, :ns], nthreads=cores, want_xg=True)
    dts = time.perf_counter() - t
    # single-problem latency, BenchmarkTools-style, beside README.md:80 (330.825 µs, unstated CPU)
    lat_fast = ob.fast_time_single(Theta[:, 0], reps=300)
    model = ob.rosenbrock()
    lat = []
    for _ in range(100):
        t = time.perf_counter()
        ob.optimize(model, Theta[:, 0])
        lat.append(time.perf_counter() - t)
    return dict(solves_per_s=conv / dt, iters_per_s=iters / dt, ms_per_step=dt * 1e3, cores=cores,
                sample=sample, converged=conv, mean_iters=iters / max(conv, 1),
                single_problem_us_median=float(np.median(lat_fast) * 1e6),
                strict_solves_per_s=rs["converged"] / dts, strict_sample=ns,
                strict_single_problem_us_median=float(np.median(lat) * 1e6))


def cpu_baseline_block(r):
    return {"value": r["solves_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": "%d of 65536 problems per step, all host cores (pthread pool over problems, fast port)" % r["sample"],
            "single_problem_us_median": r["single_problem_us_median"],
            "readme_single_problem_us_median": 330.825,
            "iters_per_s": r["iters_per_s"],
            "strict_oracle": {"value": r["strict_solves_per_s"], "unit": UNIT,
                              "sample": "%d problems, once" % r["strict_sample"],
                              "single_problem_us_median": r["strict_single_problem_us_median"]},
            "note": PORT_NOTE}


def run_reference_arm(args, rank):
    if rank != 0:
        return 0
    sample = args.cpu_sample if args.cpu_sample > 0 else BATCH  # default: the FULL 2^16 batch per step
    r = cpu_reference_run(args.steps, args.warmup, sample)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["solves_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n": N_DIM, "global_batch": BATCH,
                   "problems_per_step": sample},
        "cpu_baseline": cpu_baseline_block(r),
        "e2e": {"value": r["solves_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args, rank, world, local_rank):
    import ctypes as C
    import torch
    import qn_b200 as qn

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    cpu_group = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that put no kernel on any GPU

    n, gbatch = N_DIM, args.batch
    eng = qn.Engine(local_rank)
    obj = qn.Rosenbrock()
    ls = qn.BackTracking()
    lsc = ls._c()
    handle = eng.model_handle(obj)
    flags = qn._lib.FLAG_FORCE_SMEM if args.path == "smem" else 0
    pin = lambda a: torch.from_numpy(a).pin_memory()

    Theta_g = global_starts(batch=gbatch)           # the global batch, identical on every rank
    mine = np.arange(rank, gbatch, world)           # problem j -> rank j mod N
    batch = len(mine)

    class Buffers:
        """Pinned host and device buffers of one slice of `b` problems."""
        def __init__(self, Theta):
            b = Theta.shape[1]
            self.b = b
            self.th_pin = pin(np.ascontiguousarray(Theta.T))  # (b, n) C-order == n×b column-major
            self.x_pin = torch.empty((b, n), dtype=torch.float64).pin_memory()
            self.g_pin = torch.empty((b, n), dtype=torch.float64).pin_memory()
            self.ell_pin = torch.empty(b, dtype=torch.float64).pin_memory()
            self.it_pin = torch.empty(b, dtype=torch.int32).pin_memory()
            self.fe_pin = torch.empty(b, dtype=torch.int32).pin_memory()
            self.st_pin = torch.empty(b, dtype=torch.int32).pin_memory()

        def to_device(self):
            self.th_d = self.th_pin.to(dev)
            self.x_d = torch.empty_like(self.th_d)
            self.g_d = torch.empty_like(self.th_d)
            self.ell_d = torch.empty(self.b, dtype=torch.float64, device=dev)
            self.it_d = torch.empty(self.b, dtype=torch.int32, device=dev)
            self.fe_d = torch.empty(self.b, dtype=torch.int32, device=dev)
            self.st_d = torch.empty(self.b, dtype=torch.int32, device=dev)
            return self

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device(B):
        eng.optimize_batched_device(obj, n, B.b, B.th_d.data_ptr(), ls=ls, x_ptr=B.x_d.data_ptr(),
                                    g_ptr=B.g_d.data_ptr(), ell_ptr=B.ell_d.data_ptr(), iters_ptr=B.it_d.data_ptr(),
                                    fevals_ptr=B.fe_d.data_ptr(), status_ptr=B.st_d.data_ptr(), flags=flags,
                                    stream=stream.cuda_stream)

    def step_host(B):
        qn._lib.check(eng._lib.qn_bfgs_optimize_batched_f64(
            eng._ctx, handle, n, B.b, B.th_pin.data_ptr(), C.byref(lsc), 1e-8, 10000, None, 0, 0,
            B.x_pin.data_ptr(), B.g_pin.data_ptr(), B.ell_pin.data_ptr(), B.it_pin.data_ptr(), B.fe_pin.data_ptr(),
            B.st_pin.data_ptr(), 0, None, flags))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_(x, op):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    max_over_ranks = lambda x: reduce_(x, dist.ReduceOp.MAX) if dist else x
    sum_over_ranks = lambda x: reduce_(x, dist.ReduceOp.SUM) if dist else x

    def timed_device_leg(B, steps, warmup):
        for _ in range(warmup):
            step_device(B)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            flush.fill_(k & 0xff)          # L2 flush between timed steps (outside the events)
            ev[k][0].record(stream)
            step_device(B)
            ev[k][1].record(stream)
        barrier()
        ms = float(np.sum([a.elapsed_time(b) for a, b in ev]))
        return max_over_ranks(ms), ms

    # ---- roofline denominators, measured live on this GPU (rank 0 only prints them)
    mb = eng.microbench() if rank == 0 else None

    # ---- kernel-only leg: this rank's slice of the global batch resident in HBM
    S = Buffers(Theta_g[:, mine]).to_device()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    ms_dev, ms_local = timed_device_leg(S, args.steps, args.warmup)
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    conv_local = int((S.st_d == 0).sum().item())
    iters_local = int(S.it_d.sum().item())
    fevals_local = int(S.fe_d.sum().item())
    conv_total = sum_over_ranks(conv_local)
    iters_total = sum_over_ranks(iters_local)
    iters_max = max_over_ranks(iters_local)
    launch = eng.last_launch()

    # ---- tail: one more step with exit stamps (not timed): idle share of the problem slots
    eng.set_option("record_exit", 1)
    step_device(S)
    torch.cuda.synchronize()
    tail = eng.tail_profile()
    eng.set_option("record_exit", 0)
    tail_idle_max = max_over_ranks(tail["idle_fraction"])
    tail_idle_mean = sum_over_ranks(tail["idle_fraction"]) / world

    # ---- end-to-end, one process per GPU: the public host API on this rank's slice, pinned host buffers
    for _ in range(max(1, args.warmup // 2)):
        step_host(S)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host(S)
    torch.cuda.synchronize()
    t_e2e_local = time.perf_counter() - t0
    barrier()
    t_e2e = max_over_ranks(t_e2e_local)
    conv_e2e = sum_over_ranks(int((S.st_pin == 0).sum().item()))
    same = bool(torch.equal(S.x_pin, S.x_d.cpu()) and torch.equal(S.it_pin, S.it_d.cpu()))
    h2d = int(S.th_pin.numel() * 8)
    d2h = int((S.x_pin.numel() + S.g_pin.numel() + S.ell_pin.numel()) * 8 + 3 * 4 * S.b)
    e2e_per_rank = {"value": conv_e2e * args.steps / t_e2e, "unit": UNIT,
                    "h2d_bytes_per_step": int(sum_over_ranks(h2d)), "d2h_bytes_per_step": int(sum_over_ranks(d2h)),
                    "ms_per_step": t_e2e / args.steps * 1e3, "same_results_as_device_leg": same,
                    "api": "qn_bfgs_optimize_batched_f64 (host pointers, pinned), every rank its slice, max over ranks"}

    # ---- end-to-end through the library's own multi-GPU entry: rank 0 alone drives all N GPUs
    e2e_multi = None
    weak = None
    if world > 1:
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
        if rank == 0:
            engines = [eng] + [qn.Engine(d) for d in range(1, world)]
            G = Buffers(Theta_g)
            ctxs = (C.c_void_p * world)(*[e._ctx for e in engines])
            models = (C.c_void_p * world)(*[e.model_handle(obj) for e in engines])

            def step_multi():
                qn._lib.check(eng._lib.qn_bfgs_optimize_batched_multi_f64(
                    ctxs, world, models, n, gbatch, G.th_pin.data_ptr(), C.byref(lsc), 1e-8, 10000,
                    G.x_pin.data_ptr(), G.g_pin.data_ptr(), G.ell_pin.data_ptr(), G.it_pin.data_ptr(),
                    G.fe_pin.data_ptr(), G.st_pin.data_ptr(), flags))

            for _ in range(2):
                step_multi()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step_multi()
            t_multi = time.perf_counter() - t0
            conv_multi = int((G.st_pin == 0).sum().item())
            # problem by problem the same bits as this rank's own slice
            same_multi = bool(torch.equal(G.x_pin[rank::world], S.x_pin) and torch.equal(G.it_pin[rank::world], S.it_pin))
            e2e_multi = {"value": conv_multi * args.steps / t_multi, "unit": UNIT,
                         "h2d_bytes_per_step": int(G.th_pin.numel() * 8),
                         "d2h_bytes_per_step": int((G.x_pin.numel() + G.g_pin.numel() + G.ell_pin.numel()) * 8 + 3 * 4 * gbatch),
                         "ms_per_step": t_multi / args.steps * 1e3, "same_results_as_device_leg": same_multi,
                         "api": "qn_bfgs_optimize_batched_multi_f64 from ONE process (rank 0) over all %d GPUs: host scatter "
                                "(strided H2D) and gather (D2H + de-interleave) inside the timed region, other ranks idle" % world}
            for e in engines[1:]:
                e.close()
        if cpu_group is not None:
            dist.barrier(group=cpu_group)
        # ---- weak scaling (round 1's number): every rank its own 2^16 problems
        W = Buffers(global_starts(batch=gbatch, key=SEED + 1 + rank)).to_device()
        wk = max(2, min(3, args.steps))
        ms_weak, _ = timed_device_leg(W, wk, 1)
        conv_weak = sum_over_ranks(int((W.st_d == 0).sum().item()))
        weak = {"value": conv_weak * wk / (ms_weak * 1e-3), "unit": UNIT, "steps": wk,
                "per_gpu_batch": gbatch, "scaling": "weak"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_reference_run(1, 0, args.cpu_sample if args.cpu_sample > 0 else 8192)

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
        tmem = launch["block"] == 128 and launch["path"] == 0
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": n, "global_batch": gbatch, "per_gpu_batch": batch,
                       "sharding": "problem j -> rank j mod %d, no data-path collective" % world,
                       "path": ["register-resident", "smem-resident", "stream"][launch["path"]],
                       "tile_home": ("tensor memory (tcgen05.ld/st), 8 problems per SM" if tmem
                                     else "register file, 6 problems per SM") if launch["path"] == 0 else None,
                       "grid": launch["grid"], "block": launch["block"], "smem_bytes": launch["smem_bytes"],
                       "l2": "256 MB buffer written between timed steps (L2 flush, outside the events)",
                       "parallelism": "independent problems sharded over %d GPU(s), no collective" % world},
            "iterations_per_s": its_per_s,
            "mean_iterations_per_solve": iters_total / max(conv_total, 1),
            "fevals_per_iteration": fevals_local / max(iters_local, 1),
            "converged": int(conv_total), "problems": gbatch,
            "wall_ms_total": t_wall * 1e3,
            "tail": {"idle_fraction_max_over_ranks": tail_idle_max, "idle_fraction_mean": tail_idle_mean,
                     "slots_per_gpu": tail["slots"], "problems_per_slot": batch / max(tail["slots"], 1),
                     "iteration_imbalance_max_over_mean": iters_max * world / max(iters_total, 1),
                     "note": "share of the launch's problem-slot time spent after the work queue ran dry (exit stamps, "
                             "one extra untimed step): the limiter of strong scaling — problems are indivisible and "
                             "take 270-680 iterations each"},
            "e2e": e2e_multi if e2e_multi is not None else e2e_per_rank,
            "gpu_launches": K * launch["kernel_launches"] * world,
            "roofline": {
                "bound": "fp64", "achieved": ach_tf, "peak": mb["dfma"], "unit": "TFLOP/s",
                "frac": ach_tf / mb["dfma"],
                "traffic": TILE_KERNEL_DRAM_BYTES_PER_LAUNCH.get(launch["block"]) if (batch == BATCH and launch["path"] == 0) else None,
                "traffic_note": "DRAM bytes per launch on the full 2^16 batch from the ncu capture under profiles/ (tensor-memory "
                                "kernel / all-register kernel); the algorithmic I/O of a launch is 31.5 MB of starts in, 64.2 MB "
                                "of results out (part of which is still in L2 when the kernel ends) — B^-1 never touches HBM",
                "peak_source": "qn_microbench_dfma, measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry)",
                "algorithmic": "8 n^2 = %d flops per BFGS iteration x iterations, per GPU" % int(flops_alg),
                "smem_bytes_accounting": {"achieved": ach_gbs, "peak": mb["lds"], "unit": "GB/s",
                                          "note": "SURVEY's 16 n^2 + 96 n bytes per iteration against measured LDS bandwidth, for "
                                                  "reference only: B^-1 lives in tensor memory, these bytes never move through "
                                                  "shared memory, so this is NOT a shared-memory roofline fraction"},
                "hbm_measured_gbs": mb["hbm"],
            },
            "clocks": clocks,
        }
        if e2e_multi is not None:
            line["e2e_per_rank"] = e2e_per_rank
        if weak is not None:
            line["weak"] = weak
        if cpu is not None:
            line["cpu_baseline"] = cpu_baseline_block(cpu)
        emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


_JSON_FD = None


def claim_stdout():
    """stdout carries ONE JSON line and nothing else: file descriptor 1 is pointed at stderr for the
    rest of the process (NCCL prints its version banner to stdout, libraries may chat), the line
    itself goes to a private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="global batch (default 2^16)")
    ap.add_argument("--path", default="auto", choices=["auto", "smem"])
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="problems per CPU step (0: reference arm = the full 2^16 batch, cpu_baseline leg = 8192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        claim_stdout()
        return run_reference_arm(args, rank)
    if world == 1 and args.gpus > 1:
        # the driver launches N>1 through torchrun; a bare `--gpus N` re-launches itself that way
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000)] + sys.argv
        return subprocess.call(cmd)
    claim_stdout()
    return run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
