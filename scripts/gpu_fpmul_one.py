import sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
for var in (0, 1, 4):
    ms = ctx.test_fp_mul_throughput(148 * 4, 256, 2000, (var + 1) * 10 + 1)
    print(var, ms, 148 * 4 * 256 * 2000 / ms / 1e6)
