// ec.cuh -- Pallas group law for the device hot path (y^2 = x^3 + 5 over Fq, a = 0).
//
// Replaces arkworks' short-Weierstrass arithmetic as used by the reference at group.rs:18-26 (MSM),
// pcdl.rs:216-218 (generator fold) and main.rs:31 (generator derivation).  Group elements are
// representation independent once normalised, so the device is free to use extended Jacobian
// "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): mixed add 8M+2S, full add 12M+2S,
// doubling 6M+3S, no inversion.  Results cross the C ABI as Jacobian (X*ZZ, Y*ZZZ, ZZ), the
// reference's `Projective` layout (consts.rs:14-21).
#pragma once
#include "fp.cuh"

namespace halo {

struct alignas(16) affine_t {  // 64 B; infinity is encoded as (0, 0), which is not on the curve (b = 5)
    fq_t x, y;
};
struct alignas(16) xyzz_t {  // 128 B; infinity <=> zz == 0
    fq_t x, y, zz, zzz;
};
struct alignas(16) jac_t {  // 96 B; infinity <=> z == 0 (arkworks Projective)
    fq_t x, y, z;
};

HALO_HD bool affine_is_inf(const affine_t& p) { return fp_is_zero(p.x) && fp_is_zero(p.y); }
HALO_HD void affine_set_inf(affine_t& p) {
    fp_zero(p.x);
    fp_zero(p.y);
}
HALO_HD bool xyzz_is_inf(const xyzz_t& p) { return fp_is_zero(p.zz); }
HALO_HD void xyzz_set_inf(xyzz_t& p) {
    fp_zero(p.x);
    fp_zero(p.y);
    fp_zero(p.zz);
    fp_zero(p.zzz);
}
HALO_HD void xyzz_from_affine(xyzz_t& r, const affine_t& p) {
    if (affine_is_inf(p)) {
        xyzz_set_inf(r);
        return;
    }
    r.x = p.x;
    r.y = p.y;
    fp_one(r.zz);
    fp_one(r.zzz);
}

// mdbl-2008-s-1 (a = 0): r = 2 * p for affine p
HALO_HD void xyzz_dbl_affine(xyzz_t& r, const affine_t& p) {
    if (affine_is_inf(p)) {
        xyzz_set_inf(r);
        return;
    }
    fq_t U, V, W, S, M, t;
    fp_dbl(U, p.y);
    fp_sqr(V, U);
    fp_mul(W, U, V);
    fp_mul(S, p.x, V);
    fp_sqr(t, p.x);
    fp_dbl(M, t);
    fp_add(M, M, t);
    fp_sqr(r.x, M);
    fp_sub(r.x, r.x, S);
    fp_sub(r.x, r.x, S);
    fp_sub(t, S, r.x);
    fp_mul(t, M, t);
    fp_mul(S, W, p.y);
    fp_sub(r.y, t, S);
    r.zz = V;
    r.zzz = W;
}

// dbl-2008-s-1 (a = 0): 6M + 3S.  r may alias p.
HALO_HD void xyzz_dbl(xyzz_t& r, const xyzz_t& p) {
    if (xyzz_is_inf(p)) {
        r = p;
        return;
    }
    fq_t U, V, W, S, M, t, x3;
    fp_dbl(U, p.y);
    fp_sqr(V, U);
    fp_mul(W, U, V);
    fp_mul(S, p.x, V);
    fp_sqr(t, p.x);
    fp_dbl(M, t);
    fp_add(M, M, t);
    fp_sqr(x3, M);
    fp_sub(x3, x3, S);
    fp_sub(x3, x3, S);
    fp_sub(t, S, x3);
    fp_mul(t, M, t);
    fp_mul(S, W, p.y);
    fp_sub(r.y, t, S);
    r.x = x3;
    fp_mul(r.zz, V, p.zz);
    fp_mul(r.zzz, W, p.zzz);
}

// madd-2008-s: acc += (neg ? -q : q) for affine q.  8M + 2S on the common path; handles
// acc = inf, q = inf, acc = q (doubling) and acc = -q (infinity) exactly -- duplicate bases and
// cancelling terms are legal MSM inputs (msm_unchecked accepts them, group.rs:20,25).
HALO_HD void xyzz_madd(xyzz_t& acc, const affine_t& q, bool neg) {
    if (affine_is_inf(q)) return;
    fq_t qy;
    fp_cneg(qy, q.y, neg);
    if (xyzz_is_inf(acc)) {
        acc.x = q.x;
        acc.y = qy;
        fp_one(acc.zz);
        fp_one(acc.zzz);
        return;
    }
    fq_t U2, S2, Pp, R, PP, PPP, Q, t;
    fp_mul(U2, q.x, acc.zz);
    fp_mul(S2, qy, acc.zzz);
    fp_sub(Pp, U2, acc.x);
    fp_sub(R, S2, acc.y);
    if (fp_is_zero(Pp)) {
        if (fp_is_zero(R)) {
            affine_t d;
            d.x = q.x;
            d.y = qy;
            xyzz_dbl_affine(acc, d);
        } else {
            xyzz_set_inf(acc);
        }
        return;
    }
    fp_sqr(PP, Pp);
    fp_mul(PPP, Pp, PP);
    fp_mul(Q, acc.x, PP);
    fp_sqr(t, R);
    fp_sub(t, t, PPP);
    fp_sub(t, t, Q);
    fp_sub(acc.x, t, Q);
    fp_sub(t, Q, acc.x);
    fp_mul(t, R, t);
    fp_mul(Q, acc.y, PPP);
    fp_sub(acc.y, t, Q);
    fp_mul(acc.zz, acc.zz, PP);
    fp_mul(acc.zzz, acc.zzz, PPP);
}

// add-2008-s: acc += q for XYZZ q.  12M + 2S.
HALO_HD void xyzz_add(xyzz_t& acc, const xyzz_t& q) {
    if (xyzz_is_inf(q)) return;
    if (xyzz_is_inf(acc)) {
        acc = q;
        return;
    }
    fq_t U1, U2, S1, S2, Pp, R, PP, PPP, Q, t;
    fp_mul(U1, acc.x, q.zz);
    fp_mul(U2, q.x, acc.zz);
    fp_mul(S1, acc.y, q.zzz);
    fp_mul(S2, q.y, acc.zzz);
    fp_sub(Pp, U2, U1);
    fp_sub(R, S2, S1);
    if (fp_is_zero(Pp)) {
        if (fp_is_zero(R)) {
            xyzz_dbl(acc, acc);
        } else {
            xyzz_set_inf(acc);
        }
        return;
    }
    fp_sqr(PP, Pp);
    fp_mul(PPP, Pp, PP);
    fp_mul(Q, U1, PP);
    fp_sqr(t, R);
    fp_sub(t, t, PPP);
    fp_sub(t, t, Q);
    fp_sub(acc.x, t, Q);
    fp_sub(t, Q, acc.x);
    fp_mul(t, R, t);
    fp_mul(Q, S1, PPP);
    fp_sub(acc.y, t, Q);
    fp_mul(t, acc.zz, q.zz);
    fp_mul(acc.zz, t, PP);
    fp_mul(t, acc.zzz, q.zzz);
    fp_mul(acc.zzz, t, PPP);
}

HALO_HD void xyzz_neg(xyzz_t& p) { fp_neg(p.y, p.y); }

// XYZZ -> Jacobian with Z = ZZ: x = X/ZZ = (X*ZZ)/ZZ^2, y = Y/ZZZ = (Y*ZZZ)/ZZZ^2 = (Y*ZZZ)/ZZ^3.
HALO_HD void xyzz_to_jac(jac_t& r, const xyzz_t& p) {
    if (xyzz_is_inf(p)) {
        fp_one(r.x);
        fp_one(r.y);
        fp_zero(r.z);
        return;
    }
    fp_mul(r.x, p.x, p.zz);
    fp_mul(r.y, p.y, p.zzz);
    r.z = p.zz;
}
// Jacobian -> XYZZ: ZZ = Z^2, ZZZ = Z^3
HALO_HD void jac_to_xyzz(xyzz_t& r, const jac_t& p) {
    if (fp_is_zero(p.z)) {
        xyzz_set_inf(r);
        return;
    }
    r.x = p.x;
    r.y = p.y;
    fp_sqr(r.zz, p.z);
    fp_mul(r.zzz, r.zz, p.z);
}
// Normalise to affine (one inversion).  Infinity -> (0, 0).
HALO_HD void xyzz_to_affine(affine_t& r, const xyzz_t& p) {
    if (xyzz_is_inf(p)) {
        affine_set_inf(r);
        return;
    }
    // 1/ZZZ, then 1/ZZ = ZZ^2 / ZZZ^2 ... cheaper: i = 1/(ZZ*ZZZ); 1/ZZ = i*ZZZ; 1/ZZZ = i*ZZ
    fq_t i, t;
    fp_mul(t, p.zz, p.zzz);
    fp_inv(i, t);
    fp_mul(t, i, p.zzz);
    fp_mul(r.x, p.x, t);
    fp_mul(t, i, p.zz);
    fp_mul(r.y, p.y, t);
}

// k * p by MSB-first double-and-add over a canonical 256-bit scalar (8 x u32).  Host glue and setup only.
HALO_HD void xyzz_mul_canon(xyzz_t& r, const xyzz_t& p, const uint32_t k[8]) {
    xyzz_t acc;
    xyzz_set_inf(acc);
    for (int i = 255; i >= 0; i--) {
        xyzz_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1u) xyzz_add(acc, p);
    }
    r = acc;
}

}  // namespace halo
