"""Sharded MSM across the GPUs of one node: the only part of the hot path that shards.

Rank r of g owns the contiguous point slice [r*n/g, (r+1)*n/g) of the generators (derived on its own GPU, K6)
and the matching slice of the scalars.  Each rank runs the single-GPU Pippenger on its slice; the g partial
results (one 96-byte Jacobian point each) are exchanged with ONE all-gather (NCCL over NVLink/NVSwitch; gloo
in the CPU tests) and every rank adds them in rank order.  Folds and the h-expansion stay on one GPU.
One process per GPU; torch.distributed is plumbing only."""
import numpy as np
import torch
import torch.distributed as dist

from . import points_sum


def slice_bounds(n_total, rank, world):
    """Contiguous point slice of rank `rank`: [first, first + count)."""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def allgather_points(partial_jac, group=None, device=None):
    """One all-gather of the per-rank partial points -> [world, 12] uint64 (rank order)."""
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(partial_jac, dtype=np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t, group=group)
    return np.stack([o.cpu().numpy().view(np.uint64) for o in out])


def combine(partial_jac, group=None, device=None):
    """All-gather + ordered sum: the full MSM result on every rank."""
    return points_sum(allgather_points(partial_jac, group, device))


class ShardedMSM:
    def __init__(self, ctx, n_total, group=None):
        self.ctx, self.group, self.n_total = ctx, group, n_total
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.first, self.count = slice_bounds(n_total, self.rank, self.world)
        ctx.derive_generators_range(self.first, self.count)
        self.device = torch.device("cuda", ctx.device)

    def local_slice(self, scalars):
        """The rows of a full [n_total, 4] scalar array this rank owns."""
        return scalars[self.first:self.first + self.count]

    def partial(self, local_scalars):
        return self.ctx.msm_gens(local_scalars)

    def partial_resident(self, d_ptr):
        return self.ctx.msm_gens_resident(d_ptr, self.count)

    def msm(self, local_scalars):
        return combine(self.partial(local_scalars), self.group, self.device)

    def msm_resident(self, d_ptr):
        return combine(self.partial_resident(d_ptr), self.group, self.device)
