"""Small end-to-end case for compute-sanitizer memcheck: MSM (variable, fixed, skewed), open, check, prover, decider."""
import sys
import numpy as np
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
from halo_accumulation_b200 import acc, group, pcdl
ctx = H.Context(0, 1 << 13)
n = 1 << 13
ctx.derive_generators(n)
rng = np.random.Generator(np.random.PCG64(1))
def rs(k):
    a = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
a = ctx.msm_gens(rs(n))
ctx.precompute_generators(12); ctx.set_fixed_base(True)
b = ctx.msm_gens(rs(n))
ctx.msm_gens(np.tile(rs(1), (n, 1)))            # oversized buckets
ctx.msm(ctx.get_generators(0, 100), rs(100), np.array([1] + [0] * 99, dtype=np.uint8))
t = ctx.msm_gens_submit(rs(n)); ctx.msm_gens_collect(t)
for P in (1, 3, 4):                               # pair-tree passes (K2b), ragged sizes, skewed scalars, both base modes
    ctx.set_tuning('pair_passes', P)
    for fixed in (True, False):
        ctx.set_fixed_base(fixed)
        ctx.msm_gens(rs(n)); ctx.msm_gens(rs(n - 77), off=5); ctx.msm_gens(np.tile(rs(1), (n, 1)))
ctx.set_tuning('pair_passes', -1); ctx.set_fixed_base(True)
for D in (1, 3):                                  # deferred head rounds of the opening
    ctx.set_tuning('ipa_defer_rounds', D)
    pp = rs(1000); Cq = pcdl.commit(ctx, pp, 1023); pcdl.open(ctx, pp, Cq, 1023, rs(1)[0])
ctx.set_tuning('ipa_defer_rounds', -1)
for m, hide in ((256, True), (1 << 13, False)):
    d = m - 1
    p, z = rs(m - 3), rs(1)[0]
    w, wb, q = (rs(1)[0], rs(1)[0], rs(m - 4)) if hide else (None, None, None)
    Cm = pcdl.commit(ctx, p, d, w)
    pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
    v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, m - 3))
    pcdl.check(ctx, Cm, d, z, v, pi)
    inst = acc.new_instance(Cm, d, z, v, pi)
    ac = acc.prover(ctx, d, [inst], rs(2), rs(1)[0], rs(m - 1), rs(1)[0])
    acc.verifier(ctx, d, [inst], ac); acc.decider(ctx, ac)
print('canaries (overwritten, live):', H.check_canaries())
ctx.close()
print("sanitize case ok")
