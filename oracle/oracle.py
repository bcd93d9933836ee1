"""oracle/oracle.py -- TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# HALO_B200_CURVE=vesta selects the Vesta build of the restatement (same sources, the two Pasta moduli in swapped roles)
CURVE = os.environ.get("HALO_B200_CURVE", "pallas")
assert CURVE in ("pallas", "vesta"), CURVE
_LIB_NAME = "liboracle.so" if CURVE == "pallas" else "liboracle_vesta.so"
_LIB_PATH = os.path.join(_HERE, _LIB_NAME)
MAX_LG = 32

u64p = C.POINTER(C.c_uint64)
u8p = C.POINTER(C.c_uint8)


class EvalProof(C.Structure):
    _fields_ = [
        ("lg_n", C.c_uint32),
        ("hiding", C.c_uint32),
        ("Ls", (C.c_uint64 * 12) * MAX_LG),
        ("Rs", (C.c_uint64 * 12) * MAX_LG),
        ("U", C.c_uint64 * 12),
        ("c", C.c_uint64 * 4),
        ("C_bar", C.c_uint64 * 12),
        ("w_prime", C.c_uint64 * 4),
    ]


class Instance(C.Structure):
    _fields_ = [
        ("C", C.c_uint64 * 12),
        ("d", C.c_uint64),
        ("z", C.c_uint64 * 4),
        ("v", C.c_uint64 * 4),
        ("pi", EvalProof),
    ]


class Accumulator(C.Structure):
    _fields_ = [
        ("C_bar", C.c_uint64 * 12),
        ("d", C.c_uint64),
        ("z", C.c_uint64 * 4),
        ("v", C.c_uint64 * 4),
        ("pi", EvalProof),
        ("h0", (C.c_uint64 * 4) * 2),
        ("U0", C.c_uint64 * 12),
        ("w", C.c_uint64 * 4),
    ]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only; no reference sources involved)."""
    srcs = [os.path.join(_HERE, f) for f in ("halo_oracle.c", "halo_oracle.h", "fp_tmpl.h")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", _LIB_NAME], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_init()
        _lib.orc_params_gs.restype = u64p
        _lib.orc_params_n.restype = C.c_uint64
    return _lib


def _p(a):
    """uint64 numpy array -> pointer (array must stay alive in the caller)."""
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def _arr(x, shape=None):
    a = np.ascontiguousarray(x, dtype=np.uint64)
    if shape is not None:
        a = a.reshape(shape)
    return a


# ---------------- field ----------------
FQ, FR = 0, 1
P_MOD = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
R_MOD = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001
if CURVE == "vesta":  # coordinate field (FQ) and scalar field (FR) of the build's curve
    P_MOD, R_MOD = R_MOD, P_MOD
_MODS = {FQ: P_MOD, FR: R_MOD}
_MONT = 1 << 256
_MASK = (1 << 64) - 1


def int_to_limbs(v):
    return [(v >> (64 * i)) & _MASK for i in range(4)]


def limbs_to_int(l):
    return sum(int(x) << (64 * i) for i, x in enumerate(l))


def to_mont(vals, which=FR):
    """ints (canonical) -> [n,4] uint64 Montgomery limbs."""
    m = _MODS[which]
    return np.array([int_to_limbs(int(v) % m * _MONT % m) for v in vals], dtype=np.uint64).reshape(-1, 4)


def from_mont(arr, which=FR):
    m = _MODS[which]
    rinv = pow(_MONT, -1, m)
    a = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [limbs_to_int(row) * rinv % m for row in a]


def random_scalars(n, seed):
    """Uniform-looking scalars as Montgomery limbs: any 4-limb value < r is the Montgomery form of
    some scalar, so masking the top two bits of random limbs (=> < 2^254 < r) gives valid inputs."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 62) - 1)
    return a


def fp_mul(a, b, which):
    a, b = _arr(a), _arr(b)
    r = np.zeros(4, dtype=np.uint64)
    lib().orc_fp_mul(which, _p(a), _p(b), _p(r))
    return r


def fp_mul_cross(a, b, which):
    """Mismatches between the assembly multiplication / addition / subtraction and their C definitions over pairs of residues."""
    a, b = np.ascontiguousarray(a, dtype=np.uint64), np.ascontiguousarray(b, dtype=np.uint64)
    f = lib().orc_fp_mul_cross
    f.restype = C.c_uint64
    return int(f(which, _p(a), _p(b), C.c_uint64(a.shape[0])))


def fp_inv(a, which):
    a = _arr(a)
    r = np.zeros(4, dtype=np.uint64)
    lib().orc_fp_inv(which, _p(a), _p(r))
    return r


def sha3_256(msg: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    buf = (C.c_uint8 * max(1, len(msg))).from_buffer_copy(msg or b"\0")
    lib().orc_sha3_256(buf, C.c_uint64(len(msg)), out)
    return bytes(out)


# ---------------- curve ----------------
def pt_to_affine(p):
    """Jacobian[12] -> (affine[8], inf)."""
    p = _arr(p, (12,))
    a = np.zeros(8, dtype=np.uint64)
    inf = lib().orc_pt_to_affine(_p(p), _p(a))
    return a, bool(inf)


def pt_to_affine_ints(p):
    a, inf = pt_to_affine(p)
    if inf:
        return None
    x, y = from_mont(a.reshape(2, 4), FQ)
    return (x, y)


def pt_from_affine_ints(pt):
    out = np.zeros(12, dtype=np.uint64)
    if pt is None:
        lib().orc_pt_from_affine(_p(np.zeros(8, dtype=np.uint64)), 1, _p(out))
        return out
    aff = to_mont([pt[0], pt[1]], FQ).reshape(8)
    lib().orc_pt_from_affine(_p(aff), 0, _p(out))
    return out


def pt_eq(a, b):
    a, b = _arr(a, (12,)), _arr(b, (12,))
    return bool(lib().orc_pt_eq(_p(a), _p(b)))


def pt_add(a, b):
    a, b = _arr(a, (12,)), _arr(b, (12,))
    r = np.zeros(12, dtype=np.uint64)
    lib().orc_pt_add(_p(a), _p(b), _p(r))
    return r


def pt_mul(p, k):
    p, k = _arr(p, (12,)), _arr(k, (4,))
    r = np.zeros(12, dtype=np.uint64)
    lib().orc_pt_mul(_p(p), _p(k), _p(r))
    return r


def pt_serialize_compressed(p) -> bytes:
    p = _arr(p, (12,))
    out = (C.c_uint8 * 33)()
    lib().orc_pt_serialize_compressed(_p(p), out)
    return bytes(out)


def affine_to_jac(aff):
    """[n,8] affine -> [n,12] Jacobian with z = 1 (Montgomery)."""
    aff = _arr(aff).reshape(-1, 8)
    one = to_mont([1], FQ)[0]
    out = np.zeros((aff.shape[0], 12), dtype=np.uint64)
    out[:, :8] = aff
    out[:, 8:] = one
    return out


# ---------------- parameters ----------------
def derive_points(start, count):
    out = np.zeros((count, 8), dtype=np.uint64)
    lib().orc_derive_points(C.c_uint64(start), C.c_uint64(count), _p(out))
    return out


def derive_points_fast(start, count):
    """orc_derive_points through a fixed-base table (setup of the benchmark's CPU arm)."""
    out = np.zeros((count, 8), dtype=np.uint64)
    lib().orc_derive_points_fast(C.c_uint64(start), C.c_uint64(count), _p(out))
    return out


def msm_derived_by_dlog(first, scalars, threads=1):
    """sum_i scalars[i] * G_{first+i} over the derived generators through their known discrete logs (property check)."""
    scalars = _arr(scalars).reshape(-1, 4)
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_msm_derived_by_dlog(C.c_uint64(first), _p(scalars), C.c_uint64(scalars.shape[0]), int(threads), _p(out))
    return out


def derive_params(n):
    lib().orc_derive_params(C.c_uint64(n))


def set_params(S, H, gs):
    S, H, gs = _arr(S, (12,)), _arr(H, (12,)), _arr(gs).reshape(-1, 8)
    lib().orc_set_params(_p(S), _p(H), _p(gs), C.c_uint64(gs.shape[0]))


def params():
    n = lib().orc_params_n()
    S = np.zeros(12, dtype=np.uint64)
    H = np.zeros(12, dtype=np.uint64)
    lib().orc_params_SH(_p(S), _p(H))
    gs = np.ctypeslib.as_array(lib().orc_params_gs(), shape=(n, 8)).copy()
    return S, H, gs


# ---------------- group.rs ----------------
def msm_affine(bases, scalars, threads=1, inf=None):
    bases, scalars = _arr(bases).reshape(-1, 8), _arr(scalars).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    infp = None
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
        infp = inf.ctypes.data_as(u8p)
    lib().orc_msm_affine(_p(bases), infp, _p(scalars), C.c_uint64(n), threads, _p(out))
    return out


def msm_naive(bases, scalars, inf=None):
    bases, scalars = _arr(bases).reshape(-1, 8), _arr(scalars).reshape(-1, 4)
    n = min(bases.shape[0], scalars.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    infp = None
    if inf is not None:
        inf = np.ascontiguousarray(inf, dtype=np.uint8)
        infp = inf.ctypes.data_as(u8p)
    lib().orc_msm_naive(_p(bases), infp, _p(scalars), C.c_uint64(n), _p(out))
    return out


def point_dot(scalars, points_jac, threads=1):
    scalars, pts = _arr(scalars).reshape(-1, 4), _arr(points_jac).reshape(-1, 12)
    n = min(scalars.shape[0], pts.shape[0])
    out = np.zeros(12, dtype=np.uint64)
    lib().orc_point_dot(_p(scalars), _p(pts), C.c_uint64(n), threads, _p(out))
    return out


def scalar_dot(xs, ys):
    xs, ys = _arr(xs).reshape(-1, 4), _arr(ys).reshape(-1, 4)
    n = min(xs.shape[0], ys.shape[0])
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_scalar_dot(_p(xs), _p(ys), C.c_uint64(n), _p(out))
    return out


def construct_powers(z, n):
    z = _arr(z, (4,))
    out = np.zeros((n, 4), dtype=np.uint64)
    lib().orc_construct_powers(_p(z), C.c_uint64(n), _p(out))
    return out


# ---------------- pedersen / pcdl ----------------
def _opt(a):
    if a is None:
        return None, None
    a = _arr(a, (4,))
    return a, _p(a)


def pedersen_commit(w, gs, ms, threads=1):
    gs, ms = _arr(gs).reshape(-1, 8), _arr(ms).reshape(-1, 4)
    out = np.zeros(12, dtype=np.uint64)
    wk, wp = _opt(w)
    rc = lib().orc_pedersen_commit(wp, _p(gs), C.c_uint64(gs.shape[0]), _p(ms), C.c_uint64(ms.shape[0]), threads, _p(out))
    if rc:
        raise ValueError(f"pedersen commit failed rc={rc}")
    return out


def h_get_poly(xis):
    xis = _arr(xis).reshape(-1, 4)
    lg_n = xis.shape[0] - 1
    out = np.zeros((1 << lg_n, 4), dtype=np.uint64)
    lib().orc_h_get_poly(_p(xis), lg_n, _p(out))
    return out


def h_eval(xis, z):
    xis, z = _arr(xis).reshape(-1, 4), _arr(z, (4,))
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_h_eval(_p(xis), xis.shape[0] - 1, _p(z), _p(out))
    return out


def pcdl_commit(coeffs, d, w=None, threads=1):
    coeffs = _arr(coeffs).reshape(-1, 4)
    out = np.zeros(12, dtype=np.uint64)
    wk, wp = _opt(w)
    rc = lib().orc_pcdl_commit(_p(coeffs), C.c_uint64(coeffs.shape[0]), C.c_uint64(d), wp, threads, _p(out))
    if rc:
        raise ValueError(f"pcdl commit failed rc={rc}")
    return out


def pcdl_open(coeffs, Cm, d, z, w=None, q=None, w_bar=None, threads=1):
    coeffs, Cm, z = _arr(coeffs).reshape(-1, 4), _arr(Cm, (12,)), _arr(z, (4,))
    pi = EvalProof()
    wk, wp = _opt(w)
    wbk, wbp = _opt(w_bar)
    qa = _arr(q).reshape(-1, 4) if q is not None else np.zeros((1, 4), dtype=np.uint64)
    nq = qa.shape[0] if q is not None else 0
    rc = lib().orc_pcdl_open(_p(coeffs), C.c_uint64(coeffs.shape[0]), _p(Cm), C.c_uint64(d), _p(z), wp, _p(qa),
                             C.c_uint64(nq), wbp, threads, C.byref(pi))
    if rc:
        raise ValueError(f"pcdl open failed rc={rc}")
    return pi


def pcdl_succinct_check(Cm, d, z, v, pi):
    Cm, z, v = _arr(Cm, (12,)), _arr(z, (4,)), _arr(v, (4,))
    lg = max(1, int(d + 1).bit_length() - 1)
    xis = np.zeros((lg + 1, 4), dtype=np.uint64)
    U = np.zeros(12, dtype=np.uint64)
    rc = lib().orc_pcdl_succinct_check(_p(Cm), C.c_uint64(d), _p(z), _p(v), C.byref(pi), _p(xis), _p(U))
    return rc, xis, U


def pcdl_check(Cm, d, z, v, pi, threads=1):
    Cm, z, v = _arr(Cm, (12,)), _arr(z, (4,)), _arr(v, (4,))
    return lib().orc_pcdl_check(_p(Cm), C.c_uint64(d), _p(z), _p(v), C.byref(pi), threads)


# ---------------- acc ----------------
def acc_prover(d, qs, h0, w, q, w_bar, threads=1):
    arr = (Instance * max(1, len(qs)))(*qs)
    h0, w, w_bar = _arr(h0, (2, 4)), _arr(w, (4,)), _arr(w_bar, (4,))
    qa = _arr(q).reshape(-1, 4)
    acc = Accumulator()
    rc = lib().orc_acc_prover(C.c_uint64(d), arr, C.c_uint64(len(qs)), _p(h0), _p(w), _p(qa), C.c_uint64(qa.shape[0]),
                              _p(w_bar), threads, C.byref(acc))
    if rc:
        raise ValueError(f"acc prover failed rc={rc}")
    return acc


def acc_verifier(d, qs, acc, threads=1):
    arr = (Instance * max(1, len(qs)))(*qs)
    return lib().orc_acc_verifier(C.c_uint64(d), arr, C.c_uint64(len(qs)), C.byref(acc), threads)


def acc_decider(acc, threads=1):
    return lib().orc_acc_decider(C.byref(acc), threads)


def acc_to_instance(acc):
    q = Instance()
    lib().orc_acc_to_instance(C.byref(acc), C.byref(q))
    return q


def make_instance(Cm, d, z, v, pi):
    q = Instance()
    C.memmove(q.C, _arr(Cm, (12,)).ctypes.data, 96)
    q.d = d
    C.memmove(q.z, _arr(z, (4,)).ctypes.data, 32)
    C.memmove(q.v, _arr(v, (4,)).ctypes.data, 32)
    q.pi = pi
    return q
