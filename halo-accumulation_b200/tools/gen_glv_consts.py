"""Derives the GLV constants of csrc/glv.cuh (namespace glv, HALO_GLV_BETA_MONT) for a Pasta curve.

Both curves are y^2 = x^3 + 5 with generator (-1, 2) and j-invariant 0: the base field holds a primitive cube root of
unity beta, the scalar field the matching lambda with lambda * (x, y) = (beta x, y).  The lattice
{(a, b): a + b lambda = 0 (mod r)} has the reduced basis (a1, b1), (a2, b2) with b2 = a1; glv_split needs
a1, -b1, a2 and the fixed-point reciprocals g1 = floor(b2 2^384 / r), g2 = floor(-b1 2^384 / r).

Run: python gen_glv_consts.py [pallas|vesta].  With `pallas` the output must equal the constants that the Pallas build
has carried since the fold kernel was written (asserted below), which pins the derivation; `vesta` prints the block that
sits under HALO_CURVE_VESTA.
"""
import sys

PALLAS_P = 0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001
PALLAS_R = 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001


def pt_add(a, b, p):
    if a is None:
        return b
    if b is None:
        return a
    (x1, y1), (x2, y2) = a, b
    if x1 == x2:
        if (y1 + y2) % p == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, p) % p
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, p) % p
    x3 = (lam * lam - x1 - x2) % p
    return (x3, (lam * (x1 - x3) - y1) % p)


def pt_mul(a, k, p):
    acc = None
    while k:
        if k & 1:
            acc = pt_add(acc, a, p)
        a = pt_add(a, a, p)
        k >>= 1
    return acc


def cube_roots(m):
    """The two primitive cube roots of unity mod the prime m (m = 1 mod 3)."""
    g = 2
    while True:
        w = pow(g, (m - 1) // 3, m)
        if w != 1:
            return w, w * w % m
        g += 1


def reduced_basis(lam, r):
    """Extended Euclid on (r, lambda) stopped at sqrt(r) (Gallant-Lambert-Vanstone), then the shorter second vector."""
    s0, t0, r0 = 1, 0, r
    s1, t1, r1 = 0, 1, lam
    rows = [(r0, t0), (r1, t1)]
    while r1 != 0:
        q = r0 // r1
        r0, r1 = r1, r0 - q * r1
        s0, s1 = s1, s0 - q * s1
        t0, t1 = t1, t0 - q * t1
        rows.append((r1, t1))
    import math
    lim = math.isqrt(r)
    l = max(i for i, (rem, _) in enumerate(rows) if rem >= lim)
    v1 = (rows[l + 1][0], -rows[l + 1][1])
    c0 = (rows[l][0], -rows[l][1])
    c2 = (rows[l + 2][0], -rows[l + 2][1])
    v2 = c0 if c0[0] ** 2 + c0[1] ** 2 <= c2[0] ** 2 + c2[1] ** 2 else c2
    return v1, v2


def limbs(v, n):
    assert 0 <= v < 1 << (64 * n)
    return "{" + ", ".join(f"0x{(v >> (64 * i)) & (2**64 - 1):016x}ull" for i in range(n)) + "}"


def derive(curve):
    p, r = (PALLAS_P, PALLAS_R) if curve == "pallas" else (PALLAS_R, PALLAS_P)
    gen = (p - 1, 2)
    assert (gen[1] ** 2 - gen[0] ** 3 - 5) % p == 0
    found = []
    for beta in cube_roots(p):
        for lam in cube_roots(r):
            if pt_mul(gen, lam, p) == (beta * gen[0] % p, gen[1]):
                found.append((beta, lam))
    return p, r, gen, found


def constants(p, r, beta, lam):
    """The layout glv_split expects: a1 = b2 > 0, b1 < 0, a2 > 0 (three limbs); None if this (beta, lambda) pair does not
    give that sign pattern (the other pair does)."""
    v1, v2 = reduced_basis(lam, r)
    for (a1, b1), (a2, b2) in ((v1, v2), (v2, v1)):
        for s1 in (1, -1):
            for s2 in (1, -1):
                A1, B1, A2, B2 = s1 * a1, s1 * b1, s2 * a2, s2 * b2
                if A1 > 0 and B1 < 0 and A2 > 0 and B2 == A1 and A1 < 1 << 128 and -B1 < 1 << 128:
                    assert (A1 + B1 * lam) % r == 0 and (A2 + B2 * lam) % r == 0
                    g1 = (B2 << 384) // r
                    g2 = (-B1 << 384) // r
                    return dict(A1=A1, B1N=-B1, A2=A2, G1=g1, G2=g2, beta_mont=beta * (1 << 256) % p)
    return None


def emit(c):
    bm = c["beta_mont"]
    print("c_beta_mont = {" + ", ".join(f"0x{(bm >> (32 * i)) & 0xffffffff:08x}u" for i in range(8)) + "}")
    print("A1  =", limbs(c["A1"], 2))
    print("B1N =", limbs(c["B1N"], 2))
    print("A2  =", limbs(c["A2"], 3))
    print("G1  =", limbs(c["G1"], 5))
    print("G2  =", limbs(c["G2"], 5))


def main():
    curve = sys.argv[1] if len(sys.argv) > 1 else "pallas"
    p, r, gen, found = derive(curve)
    for beta, lam in found:
        c = constants(p, r, beta, lam)
        print(f"# {curve}: beta = 0x{beta:064x}\n#        lambda = 0x{lam:064x}  usable = {c is not None}")
        if c:
            emit(c)
            if curve == "pallas" and beta == 0x2D33357CB532458ED3552A23A8554E5005270D29D19FC7D27B7FD22F0201B547:
                assert c["A1"] == 0x49E69D1640A899538CB1279300000000
                assert c["B1N"] == 0x49E69D1640F049157FCAE1C700000001
                assert c["A2"] == 0x93CD3A2C8198E2690C7C095A00000001
                assert limbs(c["G1"], 5).startswith("{0x4a95a2d972171db4ull, 0x61afdea68480fa55ull")
                assert limbs(c["G2"], 5).startswith("{0xc689c5879f98a4deull, 0x61afdea683e7688aull")
                print("# matches csrc/glv.cuh (Pallas)")


if __name__ == "__main__":
    main()
