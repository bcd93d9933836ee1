"""MSM phase timings: variable-base vs FIXED-base (precomputed multiples) across sizes and window widths."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
import torch
out = open("gpurun_out/msm_probe.jsonl", "a")
def emit(**kw):
    print(json.dumps(kw)); out.write(json.dumps(kw) + "\n"); out.flush()
lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 18, 20, 22, 24]
cs_fixed = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
ctx = H.Context(0, 1 << max(lgs))
import os
for kv in os.environ.get('TUNE', '').split(','):
    if kv:
        k_, v_ = kv.split('='); ctx.set_tuning(k_, int(v_))
ctx.set_profiling(True)
for lg in lgs:
    n = 1 << lg
    ctx.derive_generators(n)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    d[:, 3] &= (1 << 62) - 1
    torch.cuda.synchronize()
    ctx.set_fixed_base(False)
    ref = ctx.msm_gens_resident(d.data_ptr(), n)
    for cv in ([0] + [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]):
        ctx.set_msm_window(cv)
        ctx.msm_gens_resident(d.data_ptr(), n)
        t = time.perf_counter(); ctx.msm_gens_resident(d.data_ptr(), n); wall = (time.perf_counter() - t) * 1e3
        emit(mode="variable", lg=lg, c=cv, wall_ms=wall, **ctx.last_msm_timings())
    ctx.set_msm_window(0)
    ctx.set_fixed_base(True)
    for c in cs_fixed:
        t = time.perf_counter(); ctx.precompute_generators(c); pre = time.perf_counter() - t
        r = ctx.msm_gens_resident(d.data_ptr(), n)
        t = time.perf_counter(); r = ctx.msm_gens_resident(d.data_ptr(), n); wall = (time.perf_counter() - t) * 1e3
        emit(mode="fixed", lg=lg, c=c, ok=H.points_equal(r, ref), precompute_s=pre, wall_ms=wall, **ctx.last_msm_timings())
