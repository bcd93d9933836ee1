"""Pair-tree passes (msm_pairs.cu) vs XYZZ-only accumulation: phase timings of the FIXED-base MSM per size / window / P."""
import json, sys, time
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
import torch
out = open("gpurun_out/pairs_probe.jsonl", "a")
def emit(**kw):
    print(json.dumps(kw)); out.write(json.dumps(kw) + "\n"); out.flush()
lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [24]
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
Ps = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 4]
ctx = H.Context(0, 1 << max(lgs))
ctx.set_profiling(True)
import os
for kv in os.environ.get("TUNE", "").split(","):
    if kv:
        k_, v_ = kv.split("="); ctx.set_tuning(k_, int(v_))
for lg in lgs:
    n = 1 << lg
    ctx.derive_generators(n)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    d[:, 3] &= (1 << 62) - 1
    torch.cuda.synchronize()
    ref = None
    for c in cs:
        ctx.precompute_generators(c)
        for P in Ps:
            ctx.set_tuning("pair_passes", P)
            r = ctx.msm_gens_resident(d.data_ptr(), n)
            if ref is None: ref = r
            best = None
            for _ in range(3):
                t = time.perf_counter(); r = ctx.msm_gens_resident(d.data_ptr(), n); wall = (time.perf_counter() - t) * 1e3
                tm = ctx.last_msm_timings()
                if best is None or tm["total"] < best[1]["total"]: best = (wall, tm)
            emit(lg=lg, c=c, P=P, tune=os.environ.get("TUNE", ""), ok=bool(H.points_equal(r, ref)), wall_ms=best[0], **best[1])
