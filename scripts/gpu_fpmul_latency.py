"""Latency-bound regime of the field multiplication: one warp per SM sub-partition (148 CTAs x 128 threads), 1 or 2
independent chains per thread.  ns per dependent multiplication = ms * 1e6 / iters."""
import sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
iters = 20000
for threads in (32, 128, 256, 512):
    for il in (1, 2):
        for var in (0, 1):
            ms = ctx.test_fp_mul_throughput(148, threads, iters, (var + 1) * 10 + il)
            print(f"threads/SM={threads} ilp={il} variant={var}: {ms*1e6/iters/il:.0f} ns per multiplication per chain-slot ({ms*1e6/iters:.0f} ns per iteration of {il})")
