"""Throughput of the generated Montgomery multiplication variants vs the portable C++ version (K1)."""
import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
SMS = 148
out = open("gpurun_out/fpmul_probe.jsonl", "w")
for var in (-1, 0, 1, 2, 3):
    for il in (1, 2):
        for threads, bps in ((128, 4), (128, 8), (256, 2), (256, 4)):
            blocks, iters = SMS * bps, 2000
            ms = min(ctx.test_fp_mul_throughput(blocks, threads, iters, (var + 1) * 10 + il) for _ in range(2))
            n = blocks * threads * iters * il
            row = dict(bench="fp_mul", variant=var, ilp=il, threads=threads, blocks_per_sm=bps, ms=ms, gmodmul_s=n / ms / 1e6)
            print(json.dumps(row)); out.write(json.dumps(row) + "\n")
