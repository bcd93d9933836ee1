"""Where random_instance (benches/acc.rs:15-29) spends its time at n = 2^lg: commit, evaluation, hiding open -- per piece."""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
from halo_accumulation_b200 import acc, group, pcdl

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n, d = 1 << lg, (1 << lg) - 1
ctx = H.Context(0, n)
ctx.derive_generators(n); ctx.precompute_generators(0)
rng = np.random.Generator(np.random.PCG64(3))
def rs(m):
    a = rng.integers(0, 1 << 64, size=(m, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
rows = []
for rep in range(6):
    dp = int(rng.integers(d // 2, d))
    p, w, z, wb, q = rs(dp + 1), rs(1)[0], rs(1)[0], rs(1)[0], rs(dp)
    t0 = time.perf_counter(); Cm = pcdl.commit(ctx, p, d, w)
    t1 = time.perf_counter(); pw = group.construct_powers(ctx, z, dp + 1)
    t2 = time.perf_counter(); v = group.scalar_dot(ctx, p, pw)
    t3 = time.perf_counter(); pi = pcdl.open(ctx, p, Cm, d, z, w, q, wb)
    t4 = time.perf_counter()
    rows.append(dict(dp=dp, commit=(t1 - t0) * 1e3, powers=(t2 - t1) * 1e3, dot=(t3 - t2) * 1e3, open_hiding=(t4 - t3) * 1e3))
for r in rows: print(json.dumps(r))
