#!/usr/bin/env python
"""bench.py -- headline benchmark: Pallas MSM points/s over resident generators (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--log-n L] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

A "step" is one MSM  sum_i s_i * G_i  over n = 2^L points per GPU (default 2^24, the size the metric is quoted on;
1 GiB of affine bases + 512 MiB of scalars per GPU, so every step streams inputs far larger than the 126 MB L2).
With N > 1 the MSM is sharded by point slice (weak scaling: every rank holds its own 2^L-point slice of an
N * 2^L-point MSM) and the partial results are combined with one all-gather per step.

`value`     : points/s, inputs resident in HBM, CUDA events on the library's stream, max over ranks; two steps in flight
              (submit / collect over device-resident scalars), the strictly sequential figure is `sequential_calls`.
`e2e`       : same metric through the C ABI call halo_msm_gens with HOST (pinned) scalars: H2D of the step's scalars
              and D2H of the window sums inside the timed region.
`roofline`  : integer pipe (IMAD) for the dominant phase, the bucket accumulation (pair-tree passes k_pair_fwd / k_pair_bwd
              and the XYZZ tail k_accumulate); peak measured in-run by the library's IMAD microbenchmark
              (MEASURED_PEAKS.json carries no integer figure).
`cpu_baseline` / `--impl reference`: the reference's arkworks algorithm restated in C (oracle/), timed on the host
              cores on a bounded sample of the same workload.  The reference itself is Rust and cannot run here.
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
sys.path.insert(0, ROOT)

CURVE = os.environ.get("HALO_B200_CURVE", "pallas")  # "vesta": the other half of the cycle, same kernels
METRIC = f"{CURVE}_msm_points_per_s"
UNIT = "points/s"
IMAD_PER_MODMUL = 136  # 2 N^2 + N for N = 8 limbs
CANON_W = 16           # canonical c = 16 -> 16 windows x 10 modmul per point in the bucket accumulation


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def host_threads():
    """All host cores this process may use.  torchrun exports OMP_NUM_THREADS=1; the CPU arm must not inherit that."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)  # read by libgomp when the oracle library is loaded
    return n


def cpu_msm_baseline(log_n_sample, threads, seed=4):
    """The reference's CPU path (arkworks-shaped Pippenger restated in oracle/halo_oracle.c) on a bounded sample."""
    from oracle import oracle as O

    n = 1 << log_n_sample
    bases = O.derive_points(2, n)
    scalars = O.random_scalars(n, seed)
    t = time.perf_counter()
    O.msm_affine(bases, scalars, threads=threads)
    dt = time.perf_counter() - t
    reps = max(1, min(16, int(round(10.0 / dt))))  # ~10 s of wall clock in total, fresh scalars per repetition
    total = dt
    for r in range(1, reps):
        scalars = O.random_scalars(n, seed + r)
        t = time.perf_counter()
        O.msm_affine(bases, scalars, threads=threads)
        total += time.perf_counter() - t
    return n * reps / total, total, reps


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    threads = host_threads()
    from oracle import oracle as O

    n = 1 << args.cpu_log_n
    bases = O.derive_points(2, n)
    scalars = O.random_scalars(n, 4)
    for _ in range(min(args.warmup, 1)):
        O.msm_affine(bases, scalars, threads=threads)
    t = time.perf_counter()
    for _ in range(args.steps):
        O.msm_affine(bases, scalars, threads=threads)
    dt = (time.perf_counter() - t) / args.steps
    val = n / dt
    sample = f"MSM of 2^{args.cpu_log_n} points per step (bounded sample of the 2^{args.log_n}-point workload), windows spread over {threads} host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 (255-bit Montgomery)", "data": "synthetic",
        "config": {"workload": f"{CURVE}_msm_2^{args.log_n}_per_gpu", "sample": sample},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Rust/arkworks (no toolchain here): timed arm is its algorithm restated in C (oracle/), kind=port",
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist

    import halo_accumulation_b200 as H
    from halo_accumulation_b200 import parallel

    rank, world, local = dist_env()
    if not os.path.exists(os.path.join(ROOT, "halo-accumulation_b200", "lib", "libhalo_b200.so")) and rank == 0:
        H.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("HALO_NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << args.log_n
    ctx = H.Context(local, n)
    first = rank * n
    t0 = time.perf_counter()
    ctx.derive_generators_range(first, n)  # this rank's point slice of the N * n point MSM
    derive_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    if not args.no_precompute:
        ctx.precompute_generators(0)       # FIXED-base tables: multiples 2^(off_w) G_i of the resident generators (setup)
    precompute_s = time.perf_counter() - t0
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    d_scalars = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device=dev, generator=g)
    d_scalars[:, 3] &= (1 << 62) - 1  # any 256-bit value < 2^254 < r is a valid Montgomery residue
    h_scalars = torch.empty((n, 4), dtype=torch.int64, pin_memory=True)
    h_scalars.copy_(d_scalars)
    torch.cuda.synchronize()
    h_np = h_scalars.numpy().view(np.uint64)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        part = ctx.msm_gens_resident(d_scalars.data_ptr(), n)
        return parallel.combine(part, None, dev) if world > 1 else part

    def step_e2e():
        part = ctx.msm_gens(h_np)
        return parallel.combine(part, None, dev) if world > 1 else part

    # ---- resident (`value`) ----
    # K independent MSM steps over scalars resident in HBM, two steps in flight (halo_msm_gens_submit_resident /
    # _collect): the counting sort of step k+1 (L2-atomic bound) runs beside the bucket accumulation of step k
    # (integer-pipe bound).  The same K steps as strictly sequential blocking calls are reported as `sequential_calls`.
    def run_resident_pipelined(k_steps):
        t = ctx.msm_gens_submit_resident(d_scalars.data_ptr(), n)
        out = None
        for k in range(k_steps):
            nxt = ctx.msm_gens_submit_resident(d_scalars.data_ptr(), n) if k + 1 < k_steps else None
            part = ctx.msm_gens_collect(t)
            out = parallel.combine(part, None, dev) if world > 1 else part
            t = nxt
        return out

    for _ in range(args.warmup):
        res_seq = step_resident()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        res_seq = step_resident()
    seq_ms = max(ctx.timer_stop(), 0.0) / args.steps
    run_resident_pipelined(max(args.warmup, 2))
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.kernel_launches()
    ctx.timer_start()
    w0 = time.perf_counter()
    res = run_resident_pipelined(args.steps)
    ev_ms = ctx.timer_stop()
    barrier()
    wall_ms = (time.perf_counter() - w0) * 1e3
    launches = ctx.kernel_launches() - l0
    clocks = sampler.summary()
    ms_step = max(ev_ms, 0.0) / args.steps
    # ---- e2e (host buffers through the C ABI) ----
    # (a) blocking call halo_msm_gens: H2D, kernels, D2H strictly in sequence
    for _ in range(max(args.warmup, 1)):
        step_e2e()
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        res_e = step_e2e()
    barrier()
    e2e_sync_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    # (b) pipelined calls halo_msm_gens_submit / _collect, two steps in flight: the H2D copy of step k+1 (its own 512 MiB,
    #     copied inside the timed region like every other step's) overlaps the kernels of step k
    def collect(t):
        part = ctx.msm_gens_collect(t)
        return parallel.combine(part, None, dev) if world > 1 else part

    # warm-up: `warmup` steps through the same two-slot pattern (both slots allocate their device buffers on first use)
    t = ctx.msm_gens_submit(h_np)
    for k in range(max(args.warmup, 2)):
        nxt = ctx.msm_gens_submit(h_np) if k + 1 < max(args.warmup, 2) else None
        collect(t)
        t = nxt
    barrier()
    w0 = time.perf_counter()
    t = ctx.msm_gens_submit(h_np)
    for k in range(args.steps):
        nxt = ctx.msm_gens_submit(h_np) if k + 1 < args.steps else None
        res_p = collect(t)
        t = nxt
    barrier()
    e2e_ms = (time.perf_counter() - w0) * 1e3 / args.steps
    # ---- per-phase profile (dominant phase: bucket accumulation), CUDA events on the library's stream ----
    ctx.set_profiling(True)
    acc_ms = []
    for _ in range(max(3, min(args.steps, 5))):
        ctx.msm_gens_resident(d_scalars.data_ptr(), n)
        acc_ms.append(ctx.last_msm_timings())
    ctx.set_profiling(False)
    phases = {k: float(np.mean([t[k] for t in acc_ms])) for k in acc_ms[0]}
    # max over ranks
    if world > 1:
        t = torch.tensor([ms_step, e2e_ms, wall_ms / args.steps, phases["accumulate"], e2e_sync_ms, seq_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step, e2e_ms, wall_step, phases["accumulate"], e2e_sync_ms, seq_ms = [float(x) for x in t.tolist()]
    else:
        wall_step = wall_ms / args.steps
    same = (bool(np.array_equal(res, res_e)) or _points_equal(res, res_e)) and _points_equal(res, res_p) and _points_equal(res, res_seq)

    out = None
    if rank == 0:
        total_points = n * world
        # integer-pipe peak, measured here: independent IMAD chains on every SM
        sms = torch.cuda.get_device_properties(local).multi_processor_count
        blocks, threads, iters = sms * 8, 256, 4096
        imad_ms = min(ctx.test_imad_throughput(0, blocks, threads, iters) for _ in range(3))
        imad_peak = blocks * threads * iters * 16 / imad_ms / 1e9  # T IMAD32/s: dependent-multiplicand mad.lo.u32 chains
        alg_imad = n * CANON_W * 10 * IMAD_PER_MODMUL  # algorithmic IMAD32 of one k_accumulate launch (canonical c = 16)
        achieved = alg_imad / (phases["accumulate"] * 1e-3) / 1e12
        traffic = None  # dram__bytes_read + dram__bytes_write of one k_accumulate launch, from the committed ncu capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))["accumulate_phase_pair_tree"]
            if tj["workload"] == f"{CURVE}_msm_2^{args.log_n}_per_gpu" and tj["fixed_base_tables"] == (not args.no_precompute):
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads_cpu = host_threads()
            v, dt, reps = cpu_msm_baseline(args.cpu_log_n, threads_cpu)
            cpu = {"value": v, "unit": UNIT, "cores": threads_cpu, "kind": "port",
                   "sample": f"{reps} MSMs of 2^{args.cpu_log_n} points ({dt:.1f} s in total), arkworks-shaped Pippenger restated in C, windows over {threads_cpu} threads"}
        secondary = None
        if world == 1 and not args.no_secondary:
            secondary = secondary_metrics(ctx, args)
        out = {
            "metric": METRIC, "value": total_points / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 (255-bit Montgomery)", "data": "synthetic",
            "config": {"workload": f"{CURVE}_msm_2^{args.log_n}_per_gpu", "points_per_gpu": n, "total_points": total_points,
                       "bases": "derived generators G_i (main.rs:18-45 rule), resident", "scalars": "uniform 254-bit, seeded",
                       "parallelism": f"point-slice x{world}, one all-gather of {world} x 96 B per step" if world > 1 else "single GPU",
                       "l2": "inputs_exceed_l2 (>= 1.5 GiB streamed per step)", "window_c": "auto",
                       "steps_in_flight": "2 (halo_msm_gens_submit_resident / _collect: the counting sort of step k+1 overlaps the accumulation of step k)",
                       "fixed_base_tables": (not args.no_precompute)},
            "wall_ms_per_step": wall_step,
            "sequential_calls": {"api": "halo_msm_gens_resident (one blocking call per step, nothing overlaps)", "ms_per_step": seq_ms,
                                 "value": total_points / (seq_ms * 1e-3)},
            "e2e": {"value": total_points / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "api": "halo_msm_gens_submit / halo_msm_gens_collect (pinned host scalars, two steps in flight)",
                    "h2d_bytes_per_step": n * 32, "d2h_bytes_per_step": 3 * 128, "result_matches_resident": same,
                    "blocking_call": {"api": "halo_msm_gens (one call; internally two point slices, 5/16 and 11/16, through the pipeline slots for n >= 2^23)", "value": total_points / (e2e_sync_ms * 1e-3), "ms_per_step": e2e_sync_ms}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "imad", "bound_note": "integer pipe (IMAD32 issue rate): the path is big-integer arithmetic, neither HBM nor tensor bound; BASELINE.json's north_star asks for the fraction of the integer-pipe roofline", "kernel": "bucket accumulation phase: k_pair_fwd / k_pair_bwd x 4 tree passes (affine, batched inversion) + k_accumulate (XYZZ tail)", "achieved": achieved, "peak": imad_peak, "unit": "TIMAD32/s",
                         "frac": achieved / imad_peak, "traffic": traffic, "traffic_unit": "DRAM bytes of the phase per MSM (ncu, profiles/r01_traffic.json); pass 0 of the tree (19 of 33 ms) is HBM bound on 128-byte random accesses at 3.6-4.0 TB/s, the rest integer-pipe bound",
                         "hbm_gbs_phase": (traffic / (phases["accumulate"] * 1e-3) / 1e9) if traffic else None,
                         "peak_source": "measured in this run (libhalo_b200 mad.lo.u32 microbenchmark, 16 independent chains per thread, all SMs); MEASURED_PEAKS.json has no integer-pipe figure. IMAD.WIDE / IMAD.HI issue at half this rate (profiles/r01_imad_pipe_rates.jsonl)",
                         "algorithmic": f"{n} pts x {CANON_W} windows x 10 modmul x {IMAD_PER_MODMUL} IMAD32 per launch",
                         "launch_ms": phases["accumulate"], "phases_ms": phases,
                         "whole_msm_frac": (n * 160 + 14.68e6) * IMAD_PER_MODMUL / (ms_step * 1e-3) / 1e12 / imad_peak},
            "cpu_baseline": cpu,
            "secondary": secondary,
            "derive_generators_s": derive_s, "precompute_tables_s": precompute_s,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return out


def _points_equal(a, b):
    import halo_accumulation_b200 as H

    return H.points_equal(a, b)


def secondary_metrics(ctx, args):
    """ASDL decider and PCDL open at n = 2^20 (BASELINE.json metric, second half; configs 3-4), device timings."""
    from halo_accumulation_b200 import acc, pcdl

    lg = args.secondary_log_n
    n, d = 1 << lg, (1 << lg) - 1
    ctx.derive_generators_range(0, n)
    if not args.no_precompute:
        ctx.precompute_generators(0)
    rng = np.random.Generator(np.random.PCG64(5))

    def rs(k):
        a = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 62) - 1)
        return a

    p, z = rs(n), rs(1)[0]
    w, wb, q = rs(1)[0], rs(1)[0], rs(n - 1)
    def timed(f, reps=2):  # one warm-up call (first use allocates the context's device buffers), then the best of `reps`
        out = f()
        best = 1e18
        for _ in range(reps):
            t = time.perf_counter()
            out = f()
            best = min(best, (time.perf_counter() - t) * 1e3)
        return out, best

    Cm, commit_ms = timed(lambda: pcdl.commit(ctx, p, d, w))
    pi, open_ms = timed(lambda: pcdl.open(ctx, p, Cm, d, z, w, q, wb))
    _, open_plain_ms = timed(lambda: pcdl.open(ctx, p, pcdl.commit(ctx, p, d), d, z))
    from halo_accumulation_b200 import group

    v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, n))
    pcdl.check(ctx, Cm, d, z, v, pi)
    best = 1e9
    for _ in range(3):
        t = time.perf_counter()
        pcdl.check(ctx, Cm, d, z, v, pi)
        best = min(best, (time.perf_counter() - t) * 1e3)
    # one accumulation step + decider
    q0 = acc.new_instance(Cm, d, z, v, pi)
    h0r, wr, qr, wbr = rs(2), rs(1)[0], rs(n - 1), rs(1)[0]
    a, prover_ms = timed(lambda: acc.prover(ctx, d, [q0], h0r, wr, qr, wbr))
    acc.verifier(ctx, d, [q0], a)
    dec = 1e9
    for _ in range(3):
        t = time.perf_counter()
        acc.decider(ctx, a)
        dec = min(dec, (time.perf_counter() - t) * 1e3)
    return {f"asdl_decider_ms_2^{lg}": dec, f"pcdl_check_ms_2^{lg}": best, f"pcdl_open_hiding_ms_2^{lg}": open_ms, f"pcdl_open_ms_2^{lg}": open_plain_ms,
            f"pcdl_commit_ms_2^{lg}": commit_ms, f"asdl_prover_ms_2^{lg}": prover_ms, "timing": "host wall clock around the synchronous call, best of 2 after one warm-up (open_ms includes a commit)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--log-n", type=int, default=24, help="log2 of the points per GPU")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-log-n", type=int, default=20, help="log2 of the bounded CPU sample")
    ap.add_argument("--secondary-log-n", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-precompute", action="store_true", help="variable-base path only (no tables of precomputed multiples)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
