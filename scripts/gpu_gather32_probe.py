"""Would an x-only copy of the FIXED-base table help the forward kernel of the first pair-tree pass?  Random 32-byte
gathers by lane pairs against random 64-byte gathers by lane quads, same number of points."""
import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
for gib, nb in ((13, -64), (6.5, -32), (13, -32), (1, -64), (0.5, -32)):
    for cps in (4, 8):
        blocks, threads, iters = 148 * cps, 256, 2000
        ms = ctx.test_gather_throughput(int(gib * 2**30), blocks, threads, iters, nb)
        g = blocks * threads * iters // (4 if nb == -64 else 2)
        print(json.dumps({"table_gib": gib, "mode": nb, "ctas_per_sm": cps, "ms": ms, "gathers_per_s": g / ms * 1e3}))
