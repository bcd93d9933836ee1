"""Per-round timing of the IPA open through the C ABI (halo_ipa_*), n = 2^lg."""
import ctypes as C, json, sys, time
import numpy as np
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
from halo_accumulation_b200._capi import p64
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
ctx = H.Context(0, n)
ctx.derive_generators(n); ctx.precompute_generators(0)
lib = ctx._lib
import os
for kv in os.environ.get('TUNE', '').split(','):
    if kv:
        k_, v_ = kv.split('='); ctx.set_tuning(k_, int(v_))
rng = np.random.Generator(np.random.PCG64(5))
def rs(k):
    a = rng.integers(0, 1 << 64, size=(k, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
p, z = rs(n), rs(1)[0]
S, Hh = ctx.get_SH()
for rep in range(2):
    st = C.c_void_p(); v = np.zeros(4, dtype=np.uint64)
    t0 = time.perf_counter()
    assert lib.halo_ipa_begin(ctx._h, p64(p), C.c_uint64(n), C.c_uint64(n), p64(z), C.byref(st), p64(v)) == 0
    t_begin = time.perf_counter() - t0
    assert lib.halo_ipa_set_hprime(st, p64(Hh)) == 0
    rows = []
    L, R = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    for r in range(lg):
        xi, xinv = rs(1)[0], rs(1)[0]  # a fresh challenge per round, as the transcript produces
        t1 = time.perf_counter(); assert lib.halo_ipa_round_lr(st, p64(L), p64(R)) == 0
        t2 = time.perf_counter(); assert lib.halo_ipa_round_fold(st, p64(xi), p64(xinv)) == 0
        # fold is asynchronous: force completion for timing
        ctx.timer_start(); ctx.timer_stop()
        t3 = time.perf_counter()
        rows.append((r, n >> (r + 1), (t2 - t1) * 1e3, (t3 - t2) * 1e3))
    U, c = np.zeros(12, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    assert lib.halo_ipa_finish(st, p64(U), p64(c)) == 0
    lib.halo_ipa_destroy(st)
    total = time.perf_counter() - t0
print(json.dumps(dict(tune=os.environ.get('TUNE',''), lg=lg, begin_ms=t_begin * 1e3, total_ms=total * 1e3, lr_ms=sum(r[2] for r in rows), fold_ms=sum(r[3] for r in rows))))
for r in rows: print("round %2d m=%8d lr=%8.3f ms fold=%8.3f ms" % r)
