"""Multi-GPU parity tests of the sharded MSM behind the C ABI (include/halo_b200.h "multi-GPU", csrc/comm.cu): the local
Pippenger per point slice, ONE ncclAllGather of the ranks' reduction partials on the library stream, ordered sum, one finish.
Need >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); on a one-GPU box they skip.
Every result is compared with the oracle's Pippenger over the whole point set and with the discrete-log property of the
derived generators."""
import os
import subprocess
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        import torch

        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs")


@needs2
@pytest.mark.parametrize("n_total,window", [(1 << 18, 0), (1 << 18, -1), ((1 << 17) + 4097, -1), (3001, -1)])
def test_node_form_matches_oracle(halo, oracle, n_total, window):
    """halo_mgpu_*: one caller thread, all GPUs of the box; window >= 0 builds FIXED-base tables per slice."""
    O = oracle
    g = min(_ngpus(), 8)
    m = halo.MultiGpu(list(range(g)), n_total, window)
    try:
        gs = O.derive_points_fast(2, n_total)
        for n, seed in ((n_total, 1), (n_total - n_total // 3, 2), (max(1, n_total // (2 * g)), 3), (0, 4)):
            sc = O.random_scalars(n, seed)
            got = m.msm_gens(sc)
            if n == 0:
                assert O.pt_to_affine(got)[1]
                continue
            assert O.pt_eq(got, O.msm_affine(gs[:n], sc, threads=8)), (n_total, n)
            assert O.pt_eq(got, O.msm_derived_by_dlog(0, sc, threads=8))
    finally:
        m.close()


@needs2
def test_rank_form_two_threads_matches_oracle(halo, oracle):
    """halo_comm_init_rank from one thread per GPU (the shape of one process per GPU): uneven slices, FIXED and variable
    base, host and device-resident scalars, and the point-wise all-gather."""
    import torch

    O = oracle
    g = 2
    n_total = (1 << 18) + 1
    uid = halo.Comm.unique_id()
    sc = O.random_scalars(n_total, 9)
    exp = O.msm_affine(O.derive_points_fast(2, n_total), sc, threads=8)
    res, errs = {}, []

    def work(r):
        try:
            ctx = halo.Context(r, 1 << 19)
            comm = halo.Comm(ctx, uid, g, r)
            try:
                first, count = halo.comm_slice(n_total, r, g)
                comm.derive_generators(n_total)
                loc = sc[first:first + count]
                res[(r, "var")] = comm.msm_gens_sharded(loc, n_total)
                comm.precompute_generators(0)
                res[(r, "fix")] = comm.msm_gens_sharded(loc, n_total)
                d = torch.from_numpy(loc.view(np.int64).copy()).to(torch.device("cuda", r))
                torch.cuda.synchronize(r)
                res[(r, "res")] = comm.msm_gens_sharded_resident(d.data_ptr(), count, n_total)
                res[(r, "gather")] = comm.allgather_sum(ctx.msm_gens(loc))
                # the host-scalar call as three (default) or two pipelined point slices (automatic from 2^23 points per rank;
                # forced here), FIXED and variable base
                ctx.set_tuning("split_blocking", 12)
                res[(r, "split_fix")] = comm.msm_gens_sharded(loc, n_total)
                ctx.set_fixed_base(False)
                res[(r, "split_var")] = comm.msm_gens_sharded(loc, n_total)
                ctx.set_tuning("split_second_16ths", 0)
                res[(r, "split2_var")] = comm.msm_gens_sharded(loc, n_total)
                ctx.set_fixed_base(True)
                res[(r, "split2_fix")] = comm.msm_gens_sharded(loc, n_total)
                ctx.set_tuning("split_second_16ths", 5)
                ctx.set_tuning("split_blocking", 23)
                # a short MSM that lives entirely on rank 0's slice: rank 1 contributes nothing
                k = 1000
                res[(r, "short")] = comm.msm_gens_sharded(sc[:k] if r == 0 else sc[:0], k)
            finally:
                comm.close()
                ctx.close()
        except Exception as e:
            errs.append((r, e))

    th = [threading.Thread(target=work, args=(r,)) for r in range(g)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for r in range(g):
        for k in ("var", "fix", "res", "gather", "split_fix", "split_var", "split2_var", "split2_fix"):
            assert O.pt_eq(res[(r, k)], exp), (r, k)
        assert O.pt_eq(res[(r, "short")], O.msm_derived_by_dlog(0, sc[:1000], threads=4)), r


@needs2
def test_c_program_drives_two_gpus(halo):
    """tests/cdriver/mgpu_driver.c: a C11 consumer of halo_mgpu_* (the binding a Rust caller of group.rs:24-26 would use
    for more than one GPU), compared with the oracle inside the same process."""
    halo.build()
    lib = os.path.join(ROOT, "halo-accumulation_b200", "lib")
    orc = os.path.join(ROOT, "oracle")
    exe = os.path.join(ROOT, "tests", "cdriver", "mgpu_driver.bin")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "cdriver", "mgpu_driver.c"), f"-L{lib}", "-lhalo_b200", f"-L{orc}", "-l:liboracle.so",
                           f"-Wl,-rpath,{lib}", f"-Wl,-rpath,{orc}"])
    r = subprocess.run([exe, str(min(_ngpus(), 8))], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "mgpu_driver ok" in r.stdout
