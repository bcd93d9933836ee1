"""oracle/pyref.py -- TEST INFRASTRUCTURE ONLY.

Pure-Python (big-int, affine arithmetic) restatement of the same algorithms as halo_oracle.c, used
to cross-check the C oracle on small cases with an implementation that shares no code with it.
Reference: code/src/{main.rs:18-45, group.rs, pcdl.rs:56-91}; arkworks conventions restated from
the published 0.5.0 crates (not in the reference tree).
"""
import hashlib
import os

P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001  # Pallas base field Fq
R = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001  # Pallas scalar field Fr
if os.environ.get("HALO_B200_CURVE", "pallas") == "vesta":  # Vesta: y^2 = x^3 + 5 over Pallas' Fr, scalars in Pallas' Fq
    P, R = R, P
MONT = 1 << 256
B = 5
GEN = (P - 1, 2)
GENESIS = b"To understand recursion, one must first understand recursion"


def inv(a, m):
    return pow(a, -1, m)


def pt_add(a, b):
    """Affine addition on y^2 = x^3 + 5; None is the point at infinity."""
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = 3 * x1 * x1 * inv(2 * y1, P) % P
    else:
        lam = (y2 - y1) * inv(x2 - x1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def pt_neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def pt_mul(a, k):
    k %= R
    acc = None
    while k:
        if k & 1:
            acc = pt_add(acc, a)
        a = pt_add(a, a)
        k >>= 1
    return acc


def msm(points, scalars):
    acc = None
    for pt, k in zip(points, scalars):
        acc = pt_add(acc, pt_mul(pt, k))
    return acc


def generator_scalar(k):
    """main.rs:22-28: SHA3-256(genesis || k as usize LE) -> from_le_bytes_mod_order."""
    dg = hashlib.sha3_256(GENESIS + k.to_bytes(8, "little")).digest()
    return int.from_bytes(dg, "little") % R


def generator(k):
    """main.rs:18-32; S = generator(0), H = generator(1), GS[i] = generator(i + 2) (main.rs:35-45)."""
    return pt_mul(GEN, generator_scalar(k))


def serialize_compressed_point(pt):
    """ark-ec SW compressed form for Pallas: 32-byte LE x then one flag byte (see halo_oracle.c)."""
    if pt is None:
        return bytes(32) + b"\x40"
    x, y = pt
    flag = 0x80 if y > (P - y) else 0x00
    return x.to_bytes(32, "little") + bytes([flag])


def serialize_scalar(s):
    return (s % R).to_bytes(32, "little")


def rho(tag, *args):
    """group.rs:41-89: args are ints (scalars) or points (tuple/None)."""
    data = b""
    for a in args:
        data += serialize_scalar(a) if isinstance(a, int) else serialize_compressed_point(a)
    dg = hashlib.sha3_256(data + tag.to_bytes(4, "little")).digest()
    return int.from_bytes(dg, "little") % R


def h_coeffs(xis):
    """pcdl.rs:56-77 closed form: coeff[j] = prod_{b: bit b of j set} xi_{lg n - b} (pinned by pcdl.rs:496-508)."""
    lg_n = len(xis) - 1
    out = []
    for j in range(1 << lg_n):
        v = 1
        for b in range(lg_n):
            if (j >> b) & 1:
                v = v * xis[lg_n - b] % R
        out.append(v)
    return out


def h_eval(xis, z):
    """pcdl.rs:79-91."""
    lg_n = len(xis) - 1
    v = (1 + xis[lg_n] * z) % R
    zi = z
    for i in range(1, lg_n):
        zi = zi * zi % R
        v = v * (1 + xis[lg_n - i] * zi) % R
    return v


# ---- limb helpers (Montgomery 4 x u64 little-endian, arkworks in-memory layout) ----
def to_mont_limbs(v, mod):
    m = v * MONT % mod
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def from_mont_limbs(limbs, mod):
    m = sum(int(l) << (64 * i) for i, l in enumerate(limbs))
    return m * inv(MONT, mod) % mod
