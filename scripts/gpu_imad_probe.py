import json, sys
sys.path.insert(0, ".")
import halo_accumulation_b200 as H
ctx = H.Context(0, 1 << 10)
SMS = 148
per_iter = {0: 16, 1: 16, 2: 16, 3: 8, 4: 8, 5: 8, 6: 8, 7: 16}
names = {0: "mad.lo.u32 (IMAD)", 1: "mad.wide.u32 split by ptxas (IMAD.WIDE+IADD3 pair)", 2: "mad.hi.u32 (IMAD.HI)", 3: "wide+carry-out+addc", 4: "IMAD.WIDE.X chains (carry in+out)", 5: "fused lo.cc/hi pair, no carry", 6: "IMAD.WIDE.X carry-in only (+1 IADD3 each)", 7: "IADD3/IADD3.X pairs (ALU)"}
for kind in (0, 2, 1, 5, 3, 6, 4, 7):
    for threads, bps in ((256, 4), (256, 8)):
        blocks, iters = SMS * bps, 4096
        ms = min(ctx.test_imad_throughput(kind, blocks, threads, iters) for _ in range(2))
        ops = blocks * threads * iters * per_iter[kind]
        print(json.dumps(dict(kind=names[kind], threads=threads, bps=bps, ms=ms, tops=ops / ms / 1e9)))
