"""Blocking halo_msm_gens at 2^24 from PAGEABLE host scalars (what a Rust Vec<Fr> is) and from pinned ones: number of staging
threads that copy pageable chunks into the pinned ring x the cuts of the two or three point slices (sixteenths of n)."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import halo_accumulation_b200 as H
n = 1 << 24
ctx = H.Context(0, n); ctx.derive_generators(n); ctx.precompute_generators(0)
rng = np.random.Generator(np.random.PCG64(7))
hp = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); hp[:, 3] &= np.uint64((1 << 62) - 1)   # pageable
d = torch.from_numpy(hp.view(np.int64)).cuda(); torch.cuda.synchronize()
hpin = torch.empty((n, 4), dtype=torch.int64, pin_memory=True); hpin.copy_(d); torch.cuda.synchronize()
hpn = hpin.numpy().view(np.uint64)
ref = ctx.msm_gens_resident(d.data_ptr(), n)
cuts = [(5, 0), (4, 0), (2, 4), (2, 5), (3, 5), (2, 6), (3, 6), (1, 4), (5, 0)]
for kind, buf, threads_list in (("pinned", hpn, (4,)), ("pageable", hp, (8, 4))):
    for threads in threads_list:
        for a, b in cuts:
            ctx.set_tuning("stage_threads", threads); ctx.set_tuning("split_first_16ths", a); ctx.set_tuning("split_second_16ths", b)
            r = ctx.msm_gens(buf); r = ctx.msm_gens(buf)
            t = time.perf_counter()
            for _ in range(6): r = ctx.msm_gens(buf)
            ms = (time.perf_counter() - t) / 6 * 1e3
            print(json.dumps({"host": kind, "stage_threads": threads, "cuts_16ths": [a, b], "blocking_ms": round(ms, 2), "ok": bool(H.points_equal(r, ref))}), flush=True)
