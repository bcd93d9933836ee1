"""First-contact probe on the B200: integer-pipe microbenchmarks + MSM phase timings.  Writes JSON lines."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import halo_accumulation_b200 as H
from oracle import oracle as O

out = open("gpurun_out/probe.jsonl", "a")


def emit(**kw):
    print(json.dumps(kw))
    out.write(json.dumps(kw) + "\n")
    out.flush()


ctx = H.Context(0, 1 << 24)
SMS = 148
for kind, name in [(0, "mad.lo.u32"), (1, "mad.wide.u32"), (2, "mad.hi.u32")]:
    for threads in (256, 512, 1024):
        blocks, iters = SMS * (2048 // threads), 4096
        ms = ctx.test_imad_throughput(kind, blocks, threads, iters)
        ops = blocks * threads * iters * 16
        emit(bench="imad", kind=name, threads=threads, ms=ms, tops=ops / ms / 1e9)
for ilp in (1, 2, 4):
    for threads, bps in ((128, 4), (256, 2), (256, 4), (512, 2)):
        blocks, iters = SMS * bps, 2000
        ms = ctx.test_fp_mul_throughput(blocks, threads, iters, ilp)
        n = blocks * threads * iters * ilp
        emit(bench="fp_mul", ilp=ilp, threads=threads, blocks_per_sm=bps, ms=ms, gmodmul_s=n / ms / 1e6)

ctx.set_profiling(True)
for lg in (10, 12, 14, 16, 18, 20, 22, 24):
    n = 1 << lg
    t = time.time()
    ctx.derive_generators(n)
    t_der = time.time() - t
    sc = O.random_scalars(n, lg)
    for c in (0, 6, 8, 10) if lg < 16 else (0, 11, 12, 13, 14, 15, 16, 17):
        ctx.set_msm_window(c)
        ctx.msm_gens(sc)
        t = time.time()
        ctx.msm_gens(sc)
        wall = time.time() - t
        emit(bench="msm", lg=lg, c=c, wall_ms=wall * 1e3, derive_s=t_der, **ctx.last_msm_timings())
    ctx.set_msm_window(0)
