"""Random-gather ceiling of HBM3e on this B200: useful GB/s of 64-byte (and 32 / 16-byte) random reads from a 13 GiB table."""
import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
out = open("gpurun_out/gather_probe.jsonl", "w")
table = 13 << 30
for nbytes in (64, -64, 32, 16):
    for bps in (2, 4, 8):
        blocks, threads, iters = 148 * bps, 256, 2000
        ms = ctx.test_gather_throughput(table, blocks, threads, iters, nbytes)
        g = blocks * threads * iters // (4 if nbytes < 0 else 1)
        row = dict(table_gib=table / 2**30, bytes_per_gather=nbytes, ctas_per_sm=bps, ms=ms, gathers_per_s=g / ms * 1e3, useful_gb_s=g * abs(nbytes) / ms / 1e6)
        print(json.dumps(row)); out.write(json.dumps(row) + "\n")
