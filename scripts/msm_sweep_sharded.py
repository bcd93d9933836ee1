"""BASELINE config 5, sharded: ONE Pallas MSM of n = 2^16 .. 2^24 points split across the g GPUs of a node (strong scaling:
rank r owns the point slice [r n / g, (r + 1) n / g) of generators and scalars), FIXED-base tables per slice,
device-resident scalars, one all-gather of g x 96 B per MSM.  Launch: torchrun --nproc-per-node g scripts/msm_sweep_sharded.py
[lg,lg,...].  Time per MSM = max over ranks of the CUDA-event time of the local partial (library stream) + the all-gather
and ordered sum, measured between barriers on the device; result checked against the sum of the ranks' partials computed
by the variable-base path (a different kernel path; both paths are checked against the oracle in tests/test_gpu_core.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import halo_accumulation_b200 as H  # noqa: E402
from halo_accumulation_b200 import parallel  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lgs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [16, 18, 20, 22, 24]
ctx = H.Context(local, (1 << max(lgs)) // world + 1)
dev = torch.device("cuda", local)
rows = []
for lg in lgs:
    n = 1 << lg
    sh = parallel.ShardedMSM(ctx, n)  # derives this rank's generator slice
    ctx.set_fixed_base(True)
    ctx.precompute_generators(0)
    g = torch.Generator(device="cuda")
    g.manual_seed(1000 * lg + sh.first % 997)
    d = torch.randint(-(1 << 63), (1 << 63) - 1, (sh.count, 4), dtype=torch.int64, device="cuda", generator=g)
    d[:, 3] &= (1 << 62) - 1
    torch.cuda.synchronize()
    full = sh.msm_resident(d.data_ptr())  # warm-up, and the result
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best_local, best_wall = 1e9, 1e9
    for _ in range(4):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.timer_start()
        part = sh.partial_resident(d.data_ptr())
        ms_local = ctx.timer_stop()
        out = parallel.combine(part, None, dev)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([ms_local, wall], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best_local, best_wall = min(best_local, t[0].item()), min(best_wall, t[1].item())
    # cross-check through the variable-base path
    ctx.set_fixed_base(False)
    full_var = sh.msm_resident(d.data_ptr())
    ok = bool(H.points_equal(full, full_var)) and bool(H.points_equal(out, full))
    row = dict(lg=lg, gpus=world, points_per_gpu=sh.count, partial_ms_max_over_ranks=best_local, msm_ms_with_allgather=best_wall,
               points_per_s=n / best_wall * 1e3, ok=ok)
    if rank == 0:
        print(json.dumps(row), flush=True)
        rows.append(row)
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    with open(f"gpurun_out/msm_sweep_sharded_g{world}.jsonl", "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
dist.barrier()
ctx.close()
dist.destroy_process_group()
