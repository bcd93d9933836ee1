import ctypes as C, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
from halo_accumulation_b200 import pcdl, group
from halo_accumulation_b200._capi import p64
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, d = 1 << lg, (1 << lg) - 1
ctx = H.Context(0, n)
import os
for kv in os.environ.get('TUNE', '').split(','):
    if kv:
        k_, v_ = kv.split('='); ctx.set_tuning(k_, int(v_))
ctx.derive_generators(n); ctx.precompute_generators(0)
rng = np.random.Generator(np.random.PCG64(5))
def rs(k):
    a = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
p, z = rs(n), rs(1)[0]
Cm = pcdl.commit(ctx, p, d); pi = pcdl.open(ctx, p, Cm, d, z)
v = group.scalar_dot(ctx, p, group.construct_powers(ctx, z, n))
def best(f, reps=5):
    f(); b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); f(); b = min(b, (time.perf_counter() - t) * 1e3)
    return b
h, U = pcdl.succinct_check(ctx, Cm, d, z, v, pi)
out = np.zeros(12, dtype=np.uint64)
ctx.set_profiling(bool(int(__import__("os").environ.get("PROF", "0"))))
res = dict(lg=lg, check_ms=best(lambda: pcdl.check(ctx, Cm, d, z, v, pi)), succinct_ms=best(lambda: pcdl.succinct_check(ctx, Cm, d, z, v, pi)),
           h_msm_ms=best(lambda: ctx._chk(ctx._lib.halo_h_msm(ctx._h, p64(h.xis), lg, p64(out)))))
res["h_msm_phases"] = ctx.last_msm_timings()
print(json.dumps(res))
