"""A/B of the bucket reduction: one-lane k_reduce_slabs ("reduce_quad" = 0) against the quad-cooperative k_reduce_slabs_quad
(1), per MSM size and mode, with the per-phase CUDA-event timings of the library; results checked against the oracle."""
import json
import sys
import time

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import halo_accumulation_b200 as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [10, 13, 16, 18, 20, 22]
ctx = H.Context(0, 1 << max(lgs))
for lg in lgs:
    n = 1 << lg
    ctx.derive_generators(n)
    sc = O.random_scalars(n, 7)
    exp = O.msm_derived_by_dlog(0, sc, threads=8)
    for mode in ("variable", "fixed"):
        if mode == "fixed":
            if lg < 17:
                continue
            ctx.precompute_generators(0)
        for quad in (0, 1):
            ctx.set_tuning("reduce_quad", quad)
            ok = bool(O.pt_eq(ctx.msm_gens(sc), exp))
            ctx.set_profiling(True)
            ts = []
            for _ in range(5):
                ctx.msm_gens(sc)
                ts.append(ctx.last_msm_timings())
            ctx.set_profiling(False)
            best = 1e9
            for _ in range(10):
                t = time.perf_counter()
                ctx.msm_gens(sc)
                best = min(best, (time.perf_counter() - t) * 1e3)
            row = dict(lg=lg, mode=mode, reduce_quad=quad, ok=ok, wall_ms_host_scalars=best,
                       **{k: float(np.median([t[k] for t in ts])) for k in ts[0]})
            print(json.dumps(row), flush=True)
ctx.close()
