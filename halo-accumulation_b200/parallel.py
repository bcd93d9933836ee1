"""Sharded MSM across the GPUs of one node: the only part of the hot path that shards.

Rank r of g owns the contiguous point slice [r*n/g, (r+1)*n/g) of the generators (derived on its own GPU, K6) and the
matching slice of the scalars.  The whole data path is behind the C ABI (include/halo_b200.h, csrc/comm.cu): every rank
runs the single-GPU Pippenger on its slice and the per-rank partials meet in ONE ncclAllGather enqueued by the library on
its own stream; the ranks' partials are added in rank order and finished once.  Folds and the h-expansion stay on one GPU.

One process per GPU.  torch.distributed is plumbing only: it carries the 128-byte NCCL unique id from rank 0 to the other
ranks (any backend: nccl under torchrun on the GPU box, gloo in the CPU tests) and the barriers of the benchmark."""
import numpy as np
import torch
import torch.distributed as dist

from . import Comm, comm_slice


def slice_bounds(n_total, rank, world):
    """Contiguous point slice of rank `rank`: [first, first + count).  Python mirror of halo_comm_slice (same rule; the CPU
    test compares the two)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def broadcast_unique_id(make_id, group=None, device=None):
    """Rank 0 calls make_id() (-> 128 bytes); every rank returns those bytes.  One broadcast through torch.distributed."""
    rank = dist.get_rank(group)
    if rank == 0:
        uid = make_id()
        assert len(uid) == Comm.ID_BYTES
        t = torch.tensor(list(uid), dtype=torch.uint8)
    else:
        t = torch.zeros(Comm.ID_BYTES, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return bytes(t.cpu().tolist())


def make_comm(ctx, group=None, device=None):
    """A library communicator (halo_comm) spanning the ranks of `group`, bound to this rank's context."""
    uid = broadcast_unique_id(Comm.unique_id, group, device)
    return Comm(ctx, uid, dist.get_world_size(group), dist.get_rank(group))


class ShardedMSM:
    """One MSM over n_total derived generators, sharded by point slice over the ranks of `group`."""

    def __init__(self, ctx, n_total, group=None, comm=None, precompute=False):
        self.ctx, self.group, self.n_total = ctx, group, n_total
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.first, self.count = comm_slice(n_total, self.rank, self.world)
        assert (self.first, self.count) == slice_bounds(n_total, self.rank, self.world)
        self.comm = comm if comm is not None else make_comm(ctx, group, torch.device("cuda", ctx.device) if dist.get_backend(group) == "nccl" else None)
        self.comm.derive_generators(n_total)
        if precompute:
            self.comm.precompute_generators(0)

    def local_slice(self, scalars):
        """The rows of a full [n_total, 4] scalar array this rank owns."""
        return scalars[self.first:self.first + self.count]

    def msm(self, local_scalars):
        """Host scalars of this rank's slice -> the full MSM result (same point on every rank)."""
        return self.comm.msm_gens_sharded(local_scalars, self.n_total)

    def msm_resident(self, d_ptr):
        return self.comm.msm_gens_sharded_resident(d_ptr, self.count, self.n_total)
