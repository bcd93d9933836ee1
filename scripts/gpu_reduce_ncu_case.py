"""Small workload for an ncu launch list of the bucket-reduction kernels: one 2^13-point variable-base MSM and one 2^20-point
FIXED-base MSM (python scripts/gpu_reduce_ncu_case.py [lg_var] [lg_fixed])."""
import sys

sys.path.insert(0, ".")
import halo_accumulation_b200 as H  # noqa: E402
from oracle import oracle as O  # noqa: E402

lv = int(sys.argv[1]) if len(sys.argv) > 1 else 13
lf = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx = H.Context(0, 1 << max(lv, lf))
ctx.derive_generators(1 << lv)
sc = O.random_scalars(1 << lv, 7)
for _ in range(2):
    ctx.msm_gens(sc)
ctx.derive_generators(1 << lf)
ctx.precompute_generators(0)
sc = O.random_scalars(1 << lf, 7)
for _ in range(2):
    ctx.msm_gens(sc)
ctx.close()
