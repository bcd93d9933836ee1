"""Pipelined MSM steps (two in flight, device-resident scalars) per setting of the sort-ahead throttle ("sort_ahead" = CTAs
per SM of the counting sort that runs beside the previous MSM's accumulation; 0 = no overlap)."""
import json, sys, time
sys.path.insert(0, ".")
import torch
import halo_accumulation_b200 as H

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
vals = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 1, 2, 3, 4, 6, 8]
n = 1 << lg
ctx = H.Context(0, n)
ctx.derive_generators(n); ctx.precompute_generators(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g)
d[:, 3] &= (1 << 62) - 1
torch.cuda.synchronize()
def run(k):
    t = ctx.msm_gens_submit_resident(d.data_ptr(), n); out = None
    for i in range(k):
        nxt = ctx.msm_gens_submit_resident(d.data_ptr(), n) if i + 1 < k else None
        out = ctx.msm_gens_collect(t); t = nxt
    return out
ref = ctx.msm_gens_resident(d.data_ptr(), n)
for v in vals:
    ctx.set_tuning("sort_ahead", v)
    run(3)
    ctx.timer_start(); out = run(10); ms = ctx.timer_stop() / 10
    print(json.dumps(dict(lg=lg, sort_ahead=v, ms_per_msm=ms, ok=bool(H.points_equal(out, ref)))), flush=True)
