"""Mirror of code/src/pcdl.rs: commit / open / succinct_check / check and HPoly."""
import ctypes as C

import numpy as np

from . import _host
from ._capi import arr, p64
from ._host import EvalProof, Rejected  # noqa: F401


class HPoly:
    """pcdl.rs:44-92"""

    def __init__(self, xis):
        self.xis = arr(xis).reshape(-1, 4)

    def get_poly(self, ctx):
        lg_n = self.xis.shape[0] - 1
        out = np.zeros((1 << lg_n, 4), dtype=np.uint64)
        ctx._chk(ctx._lib.halo_h_expand(ctx._h, p64(self.xis), lg_n, p64(out)))
        return out

    def eval(self, z):
        z = arr(z, (4,))
        out = np.zeros(4, dtype=np.uint64)
        _host.chk(_host.lib().halo_h_eval(p64(self.xis), self.xis.shape[0] - 1, p64(z), p64(out)))
        return out


def commit(ctx, p, d, w=None):
    """pcdl.rs:99-110"""
    p = arr(p).reshape(-1, 4)
    out = np.zeros(12, dtype=np.uint64)
    wk, wp = _host.opt(w)
    _host.chk(_host.lib().halo_pcdl_commit(ctx._h, p64(p), C.c_uint64(p.shape[0]), C.c_uint64(d), wp, p64(out)))
    return out


def open(ctx, p, Cm, d, z, w=None, q=None, w_bar=None):
    """pcdl.rs:120-242.  The reference's rng draws are explicit: q (deg p coefficients) and w_bar when hiding."""
    p, Cm, z = arr(p).reshape(-1, 4), arr(Cm, (12,)), arr(z, (4,))
    pi = EvalProof()
    wk, wp = _host.opt(w)
    wbk, wbp = _host.opt(w_bar)
    qa = arr(q).reshape(-1, 4) if q is not None else None
    _host.chk(_host.lib().halo_pcdl_open(ctx._h, p64(p), C.c_uint64(p.shape[0]), p64(Cm), C.c_uint64(d), p64(z), wp,
                                         p64(qa) if qa is not None else None, C.c_uint64(qa.shape[0] if qa is not None else 0),
                                         wbp, C.byref(pi)))
    return pi


def succinct_check(ctx, Cm, d, z, v, pi):
    """pcdl.rs:252-314 -> (HPoly, U); raises Rejected"""
    Cm, z, v = arr(Cm, (12,)), arr(z, (4,)), arr(v, (4,))
    lg = max(int(d + 1).bit_length() - 1, 0)
    xis = np.zeros((lg + 1, 4), dtype=np.uint64)
    U = np.zeros(12, dtype=np.uint64)
    _host.chk(_host.lib().halo_pcdl_succinct_check(ctx._h, p64(Cm), C.c_uint64(d), p64(z), p64(v), C.byref(pi), p64(xis), p64(U)))
    return HPoly(xis), U


def check(ctx, Cm, d, z, v, pi):
    """pcdl.rs:323-342; raises Rejected"""
    Cm, z, v = arr(Cm, (12,)), arr(z, (4,)), arr(v, (4,))
    _host.chk(_host.lib().halo_pcdl_check(ctx._h, p64(Cm), C.c_uint64(d), p64(z), p64(v), C.byref(pi)))
