"""Pipelined submit/collect (pinned host scalars) vs resident MSM at 2^24, with and without the pair tree."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import halo_accumulation_b200 as H
n = 1 << 24
ctx = H.Context(0, n); ctx.derive_generators(n); ctx.precompute_generators(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); d[:, 3] &= (1 << 62) - 1
h = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); h.copy_(d); torch.cuda.synchronize()
hn = h.numpy().view(np.uint64)
steps = 8
import os
for P in (-1, -1):
    ctx.set_tuning("pair_passes", P)
    ctx.set_tuning("sort_ahead", int(os.environ.get("AHEAD", "1")))
    for _ in range(2): ctx.msm_gens_resident(d.data_ptr(), n)
    t = time.perf_counter()
    for _ in range(steps): ctx.msm_gens_resident(d.data_ptr(), n)
    res = (time.perf_counter() - t) / steps * 1e3
    ctx.msm_gens_collect(ctx.msm_gens_submit(hn))
    t = time.perf_counter()
    tk = ctx.msm_gens_submit(hn)
    for k in range(steps):
        nx = ctx.msm_gens_submit(hn) if k + 1 < steps else None
        ctx.msm_gens_collect(tk); tk = nx
    pip = (time.perf_counter() - t) / steps * 1e3
    ctx.msm_gens_collect(ctx.msm_gens_submit_resident(d.data_ptr(), n))
    t = time.perf_counter()
    tk = ctx.msm_gens_submit_resident(d.data_ptr(), n)
    for k in range(steps):
        nx = ctx.msm_gens_submit_resident(d.data_ptr(), n) if k + 1 < steps else None
        ctx.msm_gens_collect(tk); tk = nx
    rpip = (time.perf_counter() - t) / steps * 1e3
    t = time.perf_counter()
    for _ in range(steps): ctx.msm_gens(hn)
    blk = (time.perf_counter() - t) / steps * 1e3
    # pure H2D
    t = time.perf_counter()
    for _ in range(steps): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); h2d = (time.perf_counter() - t) / steps * 1e3
    print(f"pair_passes={P}: resident {res:.2f} ms, resident pipelined {rpip:.2f} ms, pipelined {pip:.2f} ms, blocking {blk:.2f} ms, H2D alone {h2d:.2f} ms")
