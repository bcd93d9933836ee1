import sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
for gib in (0.125, 0.5, 1, 2, 4, 8, 13, 26, 52):
    table = int(gib * 2**30)
    for nb in (64, -64):
        blocks, threads, iters = 148 * 8, 256, 2000
        ms = ctx.test_gather_throughput(table, blocks, threads, iters, nb)
        g = blocks * threads * iters // (4 if nb < 0 else 1)
        print(f"table {gib:6.3f} GiB  mode {nb:4d}: {g/ms*1e3/1e9:6.2f} G gathers/s")
