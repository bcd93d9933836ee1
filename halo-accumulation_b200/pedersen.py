"""Mirror of code/src/pedersen.rs."""
import ctypes as C

import numpy as np

from . import _host
from ._capi import arr, p64


def commit(ctx, w, Gs, ms):
    """pedersen.rs:6-20 commit(w, Gs, ms).  Gs: [n,8] affine points, or an int n meaning GS[0..n) of the context."""
    ms = arr(ms).reshape(-1, 4)
    out = np.zeros(12, dtype=np.uint64)
    wk, wp = _host.opt(w)
    if isinstance(Gs, (int, np.integer)):
        gp, n_gs = None, int(Gs)
    else:
        Gs = arr(Gs).reshape(-1, 8)
        gp, n_gs = p64(Gs), Gs.shape[0]
    _host.chk(_host.lib().halo_pedersen_commit(ctx._h, wp, gp, C.c_uint64(n_gs), p64(ms), C.c_uint64(ms.shape[0]), p64(out)))
    return out
