// glv.cuh -- host side of the GLV endomorphism (shared by ipa.cu and the host layer above the C ABI).
//
// Both Pasta curves have j-invariant 0: the coordinate field holds a primitive cube root of unity beta, the scalar field
// the matching lambda with lambda * (x, y) = (beta x, y) = phi(P).  A scalar k is split as k = k1 + k2 lambda (mod r) with
// |k1|, |k2| < 2^130 by Babai rounding against the reduced basis (a1, b1), (a2, b2) of {(a, b): a + b lambda = 0 mod r};
// the pair is then written in joint sparse form.  Used by the generator fold (K4: digits broadcast to the kernel through
// constant memory) and by `Projective * Fr` in the host glue (xyzz_mul_glv: 129 doublings + ~65 additions instead of
// 255 + ~128).  Constants of both curves come from tools/gen_glv_consts.py (which re-derives the Pallas set as a check).
// Plain C++: nothing here runs on the device.
#pragma once
#include "ec.cuh"

namespace halo {

#if defined(HALO_CURVE_VESTA)
#define HALO_GLV_BETA_MONT {0x7feeeee3u, 0x410e7d20u, 0xd8fa2279u, 0x6afdf14fu, 0xeca4d4d7u, 0xfd3d8a04u, 0x77dba4efu, 0x2de2d607u}
#else
#define HALO_GLV_BETA_MONT {0x9e65eac8u, 0xfbdfd7aau, 0xe50025fbu, 0x0cd4d654u, 0x3785b99au, 0xd59892a3u, 0x585e8789u, 0x2a27fb62u}
#endif

struct GlvDigits {
    int8_t d1[136];  // joint-sparse-form digits of k1 (sign folded in), LSB first
    int8_t d2[136];  // ... and of k2
    int16_t top;     // highest index with a non-zero digit in either (-1 if xi == 0)
};

namespace glv {
typedef unsigned __int128 u128;
// out[0 .. na+nb) = a * b (little-endian u64 limbs)
inline void mul(uint64_t* out, const uint64_t* a, int na, const uint64_t* b, int nb) {
    for (int i = 0; i < na + nb; i++) out[i] = 0;
    for (int i = 0; i < na; i++) {
        u128 c = 0;
        for (int j = 0; j < nb; j++) {
            c += (u128)a[i] * b[j] + out[i + j];
            out[i + j] = (uint64_t)c;
            c >>= 64;
        }
        out[i + nb] = (uint64_t)c;
    }
}
// 5-limb two's complement helpers
inline void add5(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    u128 c = 0;
    for (int i = 0; i < 5; i++) {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
}
inline void neg5(uint64_t* r, const uint64_t* a) {
    u128 c = 1;
    for (int i = 0; i < 5; i++) {
        c += (uint64_t)~a[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
}
inline void sub5(uint64_t* r, const uint64_t* a, const uint64_t* b) {
    uint64_t nb[5];
    neg5(nb, b);
    add5(r, a, nb);
}
// basis and fixed-point reciprocals g_i = floor(|.| 2^384 / r) (tools: see DESIGN.md; checked by the open parity tests)
#if defined(HALO_CURVE_VESTA)
static const uint64_t A1[2] = {0x8cb1279300000001ull, 0x49e69d1640a89953ull};            // a1 = b2
static const uint64_t B1N[2] = {0x7fcae1c700000000ull, 0x49e69d1640f04915ull};           // -b1
static const uint64_t A2[3] = {0x0c7c095a00000001ull, 0x93cd3a2c8198e269ull, 0x0ull};    // a2
static const uint64_t G1[5] = {0x841414c24bf99a82ull, 0x61afdea685cc1578ull, 0x32c49e4c00000003ull, 0x279a745902a2654eull, 0x1ull};
static const uint64_t G2[5] = {0x0009789fdd747ae0ull, 0x61afdea6853283aeull, 0xff2b871bffffffffull, 0x279a745903c12455ull, 0x1ull};
#else
static const uint64_t A1[2] = {0x8cb1279300000000ull, 0x49e69d1640a89953ull};            // a1 = b2
static const uint64_t B1N[2] = {0x7fcae1c700000001ull, 0x49e69d1640f04915ull};           // -b1
static const uint64_t A2[3] = {0x0c7c095a00000001ull, 0x93cd3a2c8198e269ull, 0x0ull};    // a2
static const uint64_t G1[5] = {0x4a95a2d972171db4ull, 0x61afdea68480fa55ull, 0x32c49e4bffffffffull, 0x279a745902a2654eull, 0x1ull};
static const uint64_t G2[5] = {0xc689c5879f98a4deull, 0x61afdea683e7688aull, 0xff2b871c00000003ull, 0x279a745903c12455ull, 0x1ull};
#endif

}  // namespace glv

// xi = k1 + k2 lambda: magnitudes (< 2^130, three limbs) and signs
inline void glv_split(const fr_t& xi, uint64_t k1m[3], bool& n1, uint64_t k2m[3], bool& n2) {
    using namespace glv;
    uint32_t kc[8];
    fp_to_canon(kc, xi);
    uint64_t k[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) k[i] = (uint64_t)kc[2 * i] | ((uint64_t)kc[2 * i + 1] << 32);
    // c1 = (k * g1) >> 384, c2 = (k * g2) >> 384   (< 2^130)
    uint64_t prod[9], c1[3], c2[3];
    mul(prod, k, 4, G1, 5);
    c1[0] = prod[6]; c1[1] = prod[7]; c1[2] = prod[8];
    mul(prod, k, 4, G2, 5);
    c2[0] = prod[6]; c2[1] = prod[7]; c2[2] = prod[8];
    // k1 = k - c1 a1 - c2 a2 ; k2 = c1 (-b1) - c2 b2   (b2 = a1), as 5-limb two's complement
    uint64_t t1[6], t2[6], k1[5], k2[5], s[5];
    mul(t1, c1, 3, A1, 2);   // 5 limbs
    mul(t2, c2, 3, A2, 3);   // 6 limbs, top is zero
    add5(s, t1, t2);
    sub5(k1, k, s);
    mul(t1, c1, 3, B1N, 2);
    mul(t2, c2, 3, A1, 2);
    sub5(k2, t1, t2);
    n1 = (k1[4] >> 63) != 0;
    n2 = (k2[4] >> 63) != 0;
    if (n1) neg5(k1, k1);
    if (n2) neg5(k2, k2);
    for (int i = 0; i < 3; i++) {
        k1m[i] = k1[i];
        k2m[i] = k2[i];
    }
}

// Joint sparse form (Solinas) of (|k1|, |k2|), signs folded in: at most one of any two consecutive positions is
// non-zero in both rows on average half of the positions carry an addition, against two thirds for two separate NAFs.
inline void make_glv_jsf(const fr_t& xi, GlvDigits& dg) {
    uint64_t k[2][3];
    bool neg[2];
    glv_split(xi, k[0], neg[0], k[1], neg[1]);
    for (int i = 0; i < 136; i++) dg.d1[i] = dg.d2[i] = 0;
    int d[2] = {0, 0}, top = -1;
    auto nz = [&](int r) { return (k[r][0] | k[r][1] | k[r][2]) != 0 || d[r] != 0; };
    for (int pos = 0; pos < 136 && (nz(0) || nz(1)); pos++) {
        int l[2], u[2];
        for (int r = 0; r < 2; r++) l[r] = (int)((k[r][0] & 7u) + (unsigned)d[r]) & 7;
        for (int r = 0; r < 2; r++) {
            if ((l[r] & 1) == 0) {
                u[r] = 0;
            } else {
                u[r] = (l[r] & 3) == 1 ? 1 : -1;
                if ((l[r] == 3 || l[r] == 5) && (l[1 - r] & 3) == 2) u[r] = -u[r];
            }
        }
        for (int r = 0; r < 2; r++) {
            if (2 * d[r] == 1 + u[r]) d[r] = 1 - d[r];
            k[r][0] = (k[r][0] >> 1) | (k[r][1] << 63);
            k[r][1] = (k[r][1] >> 1) | (k[r][2] << 63);
            k[r][2] >>= 1;
        }
        dg.d1[pos] = (int8_t)(neg[0] ? -u[0] : u[0]);
        dg.d2[pos] = (int8_t)(neg[1] ? -u[1] : u[1]);
        if (u[0] || u[1]) top = pos;
    }
    dg.top = (int16_t)top;
}

// `Projective * Fr` on the host: k P = k1 P + k2 phi(P), joint double-and-add over the joint sparse form.  The combined
// points cost one addition: P + phi(P) = -phi^2(P) = (beta^2 X, -Y) is free in XYZZ, P - phi(P) is one full addition.
inline void xyzz_mul_glv(xyzz_t& r, const xyzz_t& p, const fr_t& k) {
    if (xyzz_is_inf(p)) {
        r = p;
        return;
    }
    GlvDigits dg;
    make_glv_jsf(k, dg);
    const uint32_t beta_limbs[8] = HALO_GLV_BETA_MONT;
    fq_t beta;
    for (int i = 0; i < 8; i++) beta.v[i] = beta_limbs[i];
    // tab[a + 1][b + 1] = a P + b phi(P) for a, b in {-1, 0, 1}
    xyzz_t tab[3][3];
    xyzz_t phi = p;
    fp_mul(phi.x, p.x, beta);
    xyzz_t sum = phi;  // P + phi(P) = (beta^2 X, -Y) = (beta * (beta X), -Y)
    fp_mul(sum.x, phi.x, beta);
    fp_neg(sum.y, p.y);
    xyzz_t nphi = phi;
    xyzz_neg(nphi);
    xyzz_t diff = p;  // P - phi(P)
    xyzz_add(diff, nphi);
    auto neg = [](const xyzz_t& q) {
        xyzz_t t = q;
        if (!xyzz_is_inf(t)) xyzz_neg(t);
        return t;
    };
    xyzz_set_inf(tab[1][1]);
    tab[2][1] = p;
    tab[0][1] = neg(p);
    tab[1][2] = phi;
    tab[1][0] = nphi;
    tab[2][2] = sum;
    tab[0][0] = neg(sum);
    tab[2][0] = diff;
    tab[0][2] = neg(diff);
    xyzz_t acc;
    xyzz_set_inf(acc);
    for (int i = dg.top; i >= 0; i--) {
        xyzz_dbl(acc, acc);
        if (dg.d1[i] | dg.d2[i]) xyzz_add(acc, tab[dg.d1[i] + 1][dg.d2[i] + 1]);
    }
    r = acc;
}

}  // namespace halo
