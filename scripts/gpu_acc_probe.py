import json, sys, time
sys.path.insert(0, ".")
import halo_accumulation_b200 as H, torch
lg = int(sys.argv[1]); n = 1 << lg
ctx = H.Context(0, n); ctx.set_profiling(True); ctx.derive_generators(n)
g = torch.Generator(device="cuda"); g.manual_seed(lg)
d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); d[:, 3] &= (1 << 62) - 1
torch.cuda.synchronize()
for fixed, cs in ((False, [16]), (True, [20, 22])):
    ctx.set_fixed_base(fixed)
    for c in cs:
        if fixed: ctx.precompute_generators(c)
        else: ctx.set_msm_window(c)
        for static, bps in ((1, 0), (2, 4), (0, 0)):
            ctx.set_tuning("acc_static", static); ctx.set_tuning("acc_blocks_per_sm", bps)
            ctx.msm_gens_resident(d.data_ptr(), n); ctx.msm_gens_resident(d.data_ptr(), n)
            t = ctx.last_msm_timings()
            print("fixed" if fixed else "var", "c=%d" % c, "static" if static else "dyn bps=%d" % bps, "acc=%.3f total=%.3f" % (t["accumulate"], t["total"]))
