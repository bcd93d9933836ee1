"""BASELINE config 5 on one GPU: Pallas MSM over resident generators, n = 2^16 .. 2^24, FIXED-base (tables) and
variable-base, device-resident scalars; best of 3 by CUDA events (library stream) + points/s."""
import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
import torch
out = open("gpurun_out/msm_sweep.jsonl", "w")
lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 18, 20, 22, 24]
ctx = H.Context(0, 1 << max(lgs))
for lg in lgs:
    n = 1 << lg
    ctx.derive_generators(n)
    g = torch.Generator(device="cuda"); g.manual_seed(lg)
    d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
    d[:, 3] &= (1 << 62) - 1
    torch.cuda.synchronize()
    ref = None
    for mode in ("variable", "fixed"):
        ctx.set_fixed_base(mode == "fixed")
        if mode == "fixed":
            ctx.precompute_generators(0)
        r = ctx.msm_gens_resident(d.data_ptr(), n)
        if ref is None: ref = r
        best = 1e9
        for _ in range(3):
            ctx.timer_start(); r = ctx.msm_gens_resident(d.data_ptr(), n); best = min(best, ctx.timer_stop())
        row = dict(lg=lg, mode=mode, ms=best, points_per_s=n / best * 1e3, ok=bool(H.points_equal(r, ref)))
        print(json.dumps(row)); out.write(json.dumps(row) + "\n"); out.flush()
