"""Blocking halo_msm_gens from pinned host scalars at 2^24: size of the first of the two point slices (its H2D copy is the
one that nothing overlaps)."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import halo_accumulation_b200 as H
n = 1 << 24
ctx = H.Context(0, n); ctx.derive_generators(n); ctx.precompute_generators(0)
g = torch.Generator(device="cuda"); g.manual_seed(1)
d = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device="cuda", generator=g); d[:, 3] &= (1 << 62) - 1
h = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); h.copy_(d); torch.cuda.synchronize()
hn = h.numpy().view(np.uint64)
ref = ctx.msm_gens_resident(d.data_ptr(), n)
for first in (8, 6, 5, 4, 3, 2, 8):
    ctx.set_tuning("split_first_16ths", first)
    r = ctx.msm_gens(hn); r = ctx.msm_gens(hn)
    t = time.perf_counter()
    for _ in range(6): r = ctx.msm_gens(hn)
    ms = (time.perf_counter() - t) / 6 * 1e3
    print(json.dumps({"first_slice_16ths": first, "blocking_ms": ms, "ok": bool(H.points_equal(r, ref))}), flush=True)
