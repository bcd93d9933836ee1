"""CPU (gloo, world_size 2) test of the host-side plumbing of the sharded MSM: the slice rule (Python mirror == the C ABI's
halo_comm_slice), the broadcast of the 128-byte communicator id through torch.distributed, and the ordered sum of per-rank
partials (halo_points_sum).  The data path itself (local Pippenger + ncclAllGather inside libhalo_b200.so) needs GPUs and is
covered by tests/test_gpu_multi.py; here the per-rank partial MSM is computed by the oracle."""
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
    import halo_accumulation_b200 as H
    from halo_accumulation_b200 import parallel
    from oracle import oracle as O

    # the id travels from rank 0 to everyone (here a recognisable fake: making a real one is NCCL's business)
    uid = parallel.broadcast_unique_id(lambda: bytes((7 * i + 1) % 256 for i in range(128)))
    first, count = parallel.slice_bounds(n, rank, world)
    assert (first, count) == H.comm_slice(n, rank, world)
    bases = O.derive_points(2 + first, count)          # this rank's generator slice
    scalars = O.random_scalars(n, 4)[first:first + count]
    partial = O.msm_affine(bases, scalars)
    gathered = [None] * world
    dist.all_gather_object(gathered, partial.tolist())  # test plumbing; the product gathers with NCCL inside the library
    total = H.points_sum(np.array(gathered, dtype=np.uint64))
    q.put((rank, first, count, total.tolist(), uid))
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
    # slices tile [0, n), both ranks hold the same id and the same, correct total
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1] and res[1][1] + res[1][2] == n
    assert res[0][4] == res[1][4] == bytes((7 * i + 1) % 256 for i in range(128))
    exp = oracle.msm_affine(oracle.derive_points(2, n), oracle.random_scalars(n, 4), threads=4)
    for r in res:
        assert oracle.pt_eq(np.array(r[3], dtype=np.uint64), exp)


def test_slice_bounds_cover():
    import halo_accumulation_b200 as H
    from halo_accumulation_b200 import parallel

    for n in (1, 7, 8, 1 << 20, (1 << 24) + 3):
        for g in (1, 2, 4, 8):
            spans = [parallel.slice_bounds(n, r, g) for r in range(g)]
            assert spans == [H.comm_slice(n, r, g) for r in range(g)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (f0, c0), (f1, _) in zip(spans, spans[1:]):
                assert f0 + c0 == f1


def test_comm_api_refuses_without_a_device():
    """No GPU here: the communicator id can be made (NCCL loads at run time), but nothing computes without a context."""
    import halo_accumulation_b200 as H

    lib = H._capi.load()
    assert lib.halo_nccl_version() > 20000
    assert len(H.Comm.unique_id()) == 128
    import ctypes as C

    out = C.c_void_p()
    assert lib.halo_comm_init_rank(None, (C.c_uint8 * 128)(), 1, 0, C.byref(out)) == -1
