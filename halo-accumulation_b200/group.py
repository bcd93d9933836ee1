"""Mirror of code/src/group.rs on the device: scalar_dot, point_dot(_affine), construct_powers."""
import numpy as np

from ._capi import arr, p64
import ctypes as C


def scalar_dot(ctx, xs, ys):
    """group.rs:13-15"""
    xs, ys = arr(xs).reshape(-1, 4), arr(ys).reshape(-1, 4)
    n = min(xs.shape[0], ys.shape[0])
    out = np.zeros(4, dtype=np.uint64)
    ctx._chk(ctx._lib.halo_scalar_dot(ctx._h, p64(xs), p64(ys), C.c_uint64(n), p64(out)))
    return out


def point_dot(ctx, xs, Gs_jac):
    """group.rs:18-21 (Projective bases)"""
    return ctx.msm_jac(Gs_jac, xs)


def point_dot_affine(ctx, xs, Gs_affine):
    """group.rs:24-26"""
    return ctx.msm(Gs_affine, xs)


def construct_powers(ctx, z, n):
    """group.rs:29-37"""
    z = arr(z, (4,))
    out = np.zeros((n, 4), dtype=np.uint64)
    ctx._chk(ctx._lib.halo_construct_powers(ctx._h, p64(z), C.c_uint64(n), p64(out)))
    return out
