"""Small variable-base MSMs for ncu: python scripts/gpu_small_msm_ncu_case.py <lg> <c> [TUNE via env]"""
import os
import sys

sys.path.insert(0, ".")
import halo_accumulation_b200 as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

lg, c = int(sys.argv[1]), int(sys.argv[2])
ctx = H.Context(0, 1 << lg)
for kv in os.environ.get("TUNE", "").split(","):
    if kv:
        k_, v_ = kv.split("=")
        ctx.set_tuning(k_, int(v_))
ctx.derive_generators(1 << lg)
ctx.set_msm_window(c)
sc = O.random_scalars(1 << lg, 7)
for _ in range(2):
    ctx.msm_gens(sc)
ctx.close()
