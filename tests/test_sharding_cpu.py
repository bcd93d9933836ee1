"""CPU (gloo, world_size 2) test of the sharded-MSM plumbing: slicing, the single all-gather and the ordered
sum.  The per-rank partial MSM is computed by the oracle here (no GPU in this container); the GPU path is the
same code with Context.msm_gens as the partial."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from halo_accumulation_b200 import parallel
    from oracle import oracle as O

    first, count = parallel.slice_bounds(n, rank, world)
    bases = O.derive_points(2 + first, count)          # this rank's generator slice
    scalars = O.random_scalars(n, 4)[first:first + count]
    partial = O.msm_affine(bases, scalars)
    total = parallel.combine(partial)
    q.put((rank, first, count, total.tolist()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 1025])
def test_sharded_msm_gloo_world2(oracle, n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # slices tile [0, n) and both ranks hold the same, correct total
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == n
    exp = oracle.msm_affine(oracle.derive_points(2, n), oracle.random_scalars(n, 4), threads=4)
    for r in res:
        assert oracle.pt_eq(np.array(r[3], dtype=np.uint64), exp)


def test_slice_bounds_cover():
    from halo_accumulation_b200 import parallel

    for n in (1, 7, 8, 1 << 20, (1 << 24) + 3):
        for g in (1, 2, 4, 8):
            spans = [parallel.slice_bounds(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
