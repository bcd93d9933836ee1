"""A/B of one tuning knob on the MSM: python scripts/gpu_tune_ab.py <key> <v0,v1,..> [lg,lg,..] [fixed|variable|both]
Per size / mode / value: per-phase CUDA-event timings of the library (median of 5) and an oracle check (discrete-log
property of the derived generators)."""
import json
import sys

sys.path.insert(0, ".")
import numpy as np  # noqa: E402

import halo_accumulation_b200 as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

key = sys.argv[1]
vals = [int(x) for x in sys.argv[2].split(",")]
lgs = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [20, 22, 24]
modes = sys.argv[4] if len(sys.argv) > 4 else "fixed"
ctx = H.Context(0, 1 << max(lgs))
for lg in lgs:
    n = 1 << lg
    ctx.derive_generators(n)
    sc = O.random_scalars(n, 7)
    exp = O.msm_derived_by_dlog(0, sc, threads=16)
    for mode in (("variable", "fixed") if modes == "both" else (modes,)):
        if mode == "fixed":
            ctx.precompute_generators(0)
        ctx.set_fixed_base(mode == "fixed")
        for v in vals:
            ctx.set_tuning(key, v)
            ctx.set_tuning("split_blocking", 0)  # one MSM per call, so that the per-phase timings describe it
            ok = bool(O.pt_eq(ctx.msm_gens(sc), exp))
            ctx.set_profiling(True)
            ts = []
            for _ in range(5):
                ctx.msm_gens(sc)
                ts.append(ctx.last_msm_timings())
            ctx.set_profiling(False)
            print(json.dumps(dict(lg=lg, mode=mode, key=key, value=v, ok=ok, **{k: round(float(np.median([t[k] for t in ts])), 4) for k in ts[0]})), flush=True)
ctx.close()
