/*
 * oracle/halo_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle + timed CPU baseline).
 * See halo_oracle.h for scope and the parity-pinning statement.
 *
 * Every function cites the reference file:line it restates (paths relative to the reference
 * tree, code/src/...).  Third-party arithmetic (arkworks 0.5.0, sha3 0.10.8; Cargo.lock:51-167,
 * :773-775) is not in the reference tree; its published algorithms are restated from scratch.
 */
#include "halo_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef uint64_t u64;
typedef unsigned __int128 u128;
typedef unsigned char u8;

/* The two Pasta moduli.  Pallas (default): coordinates over p, scalars over r.  -DHALO_CURVE_VESTA builds the same
 * restatement for Vesta, y^2 = x^3 + 5 over r with scalar field p: "fq" always names the coordinate
 * field of the build's curve and "fr" its scalar field.
 *   p = 0x40000000000000000000000000000000224698fc094cf91b992d30ed00000001   (Pallas base field)
 *   r = 0x40000000000000000000000000000000224698fc0994a8dd8c46eb2100000001   (Pallas scalar field) */
#define PASTA_P0 0x992d30ed00000001ULL
#define PASTA_P1 0x224698fc094cf91bULL
#define PASTA_R0 0x8c46eb2100000001ULL
#define PASTA_R1 0x224698fc0994a8ddULL
#ifdef HALO_CURVE_VESTA
#define BASE0 PASTA_R0
#define BASE1 PASTA_R1
#define SCAL0 PASTA_P0
#define SCAL1 PASTA_P1
#else
#define BASE0 PASTA_P0
#define BASE1 PASTA_P1
#define SCAL0 PASTA_R0
#define SCAL1 PASTA_R1
#endif

#define FP_NAME fq
#define FP_MOD0 BASE0
#define FP_MOD1 BASE1
#define FP_MOD2 0x0000000000000000ULL
#define FP_MOD3 0x4000000000000000ULL
#include "fp_tmpl.h"

#define FP_NAME fr
#define FP_MOD0 SCAL0
#define FP_MOD1 SCAL1
#define FP_MOD2 0x0000000000000000ULL
#define FP_MOD3 0x4000000000000000ULL
#include "fp_tmpl.h"

/* ------------------------------------------------------------------------------------------ */
/* SHA3-256 (FIPS 202), the `sha3` crate's Sha3_256 used at group.rs:52-55 and main.rs:22-25.  */
/* ------------------------------------------------------------------------------------------ */
static const u64 KECCAK_RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int KECCAK_ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline u64 rotl64(u64 x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

static void keccak_f1600(u64 st[25]) {
    for (int round = 0; round < 24; round++) {
        u64 C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
        /* rho + pi: B[y, 2x+3y] = rot(A[x, y]) */
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], KECCAK_ROT[x + 5 * y]);
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        st[0] ^= KECCAK_RC[round];
    }
}

void orc_sha3_256(const u8 *msg, u64 len, u8 out[32]) {
    enum { RATE = 136 };
    u64 st[25];
    u8 block[RATE];
    memset(st, 0, sizeof st);
    while (len >= RATE) {
        for (int i = 0; i < RATE / 8; i++) {
            u64 v = 0;
            for (int j = 7; j >= 0; j--) v = (v << 8) | msg[8 * i + j];
            st[i] ^= v;
        }
        keccak_f1600(st);
        msg += RATE;
        len -= RATE;
    }
    memset(block, 0, RATE);
    memcpy(block, msg, len);
    block[len] ^= 0x06;
    block[RATE - 1] ^= 0x80;
    for (int i = 0; i < RATE / 8; i++) {
        u64 v = 0;
        for (int j = 7; j >= 0; j--) v = (v << 8) | block[8 * i + j];
        st[i] ^= v;
    }
    keccak_f1600(st);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) out[8 * i + j] = (u8)(st[i] >> (8 * j));
}

/* ------------------------------------------------------------------------------------------ */
/* init                                                                                        */
/* ------------------------------------------------------------------------------------------ */
static int g_inited = 0;
static u64 FQ_FIVE[4]; /* curve constant b = 5 */
static u64 GEN_X[4], GEN_Y[4]; /* generator (-1, 2), ark-pallas */

void orc_init(void) {
    if (g_inited) return;
    fq_init();
    fr_init();
    fq_from_u64(FQ_FIVE, 5);
    u64 one[4];
    fq_one(one);
    fq_neg(GEN_X, one);
    fq_from_u64(GEN_Y, 2);
    g_inited = 1;
}
__attribute__((constructor)) static void orc_ctor(void) { orc_init(); }

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* exported field helpers                                                                      */
/* ------------------------------------------------------------------------------------------ */
void orc_fp_mul(int w, const u64 a[4], const u64 b[4], u64 r[4]) { if (w) fr_mul(r, a, b); else fq_mul(r, a, b); }
/* count of pairs on which the assembly multiplication / addition / subtraction and their C definitions (fp_tmpl.h) disagree,
 * or a result is not reduced */
u64 orc_fp_mul_cross(int w, const u64 *a, const u64 *b, u64 n) {
    u64 bad = 0;
    for (u64 i = 0; i < n; i++) {
        u64 x[4], y[4];
        if (w) {
            fr_mul(x, a + 4 * i, b + 4 * i); fr_mul_c(y, a + 4 * i, b + 4 * i); bad += !fr_eq(x, y) || fr_geq_p(x);
            fr_add(x, a + 4 * i, b + 4 * i); fr_add_c(y, a + 4 * i, b + 4 * i); bad += !fr_eq(x, y) || fr_geq_p(x);
            fr_sub(x, a + 4 * i, b + 4 * i); fr_sub_c(y, a + 4 * i, b + 4 * i); bad += !fr_eq(x, y) || fr_geq_p(x);
        } else {
            fq_mul(x, a + 4 * i, b + 4 * i); fq_mul_c(y, a + 4 * i, b + 4 * i); bad += !fq_eq(x, y) || fq_geq_p(x);
            fq_add(x, a + 4 * i, b + 4 * i); fq_add_c(y, a + 4 * i, b + 4 * i); bad += !fq_eq(x, y) || fq_geq_p(x);
            fq_sub(x, a + 4 * i, b + 4 * i); fq_sub_c(y, a + 4 * i, b + 4 * i); bad += !fq_eq(x, y) || fq_geq_p(x);
        }
    }
    return bad;
}
void orc_fp_add(int w, const u64 a[4], const u64 b[4], u64 r[4]) { if (w) fr_add(r, a, b); else fq_add(r, a, b); }
void orc_fp_sub(int w, const u64 a[4], const u64 b[4], u64 r[4]) { if (w) fr_sub(r, a, b); else fq_sub(r, a, b); }
void orc_fp_inv(int w, const u64 a[4], u64 r[4]) { if (w) fr_inv(r, a); else fq_inv(r, a); }
void orc_fp_to_canon(int w, const u64 a[4], u64 r[4]) { if (w) fr_to_canon(r, a); else fq_to_canon(r, a); }
void orc_fp_from_canon(int w, const u64 a[4], u64 r[4]) { if (w) fr_from_canon(r, a); else fq_from_canon(r, a); }
void orc_fp_from_le_bytes_mod_order(int w, const u8 b[32], u64 r[4]) {
    if (w) fr_from_le_bytes_mod_order(r, b); else fq_from_le_bytes_mod_order(r, b);
}

/* ------------------------------------------------------------------------------------------ */
/* Pallas: y^2 = x^3 + 5 over Fq, Jacobian coordinates (X/Z^2, Y/Z^3), a = 0.                  */
/* ark-ec short_weierstrass::Projective; infinity <=> Z == 0.                                  */
/* ------------------------------------------------------------------------------------------ */
typedef struct { u64 x[4], y[4], z[4]; } jac_t;
typedef struct { u64 x[4], y[4]; } aff_t;

static inline void jac_set_inf(jac_t *p) { fq_one(p->x); fq_one(p->y); fq_zero(p->z); }
static inline int jac_is_inf(const jac_t *p) { return fq_is_zero(p->z); }

/* dbl-2009-l (a = 0): 2M + 5S */
static void jac_double(jac_t *r, const jac_t *p) {
    if (jac_is_inf(p)) { *r = *p; return; }
    u64 A[4], B[4], C[4], D[4], E[4], Fv[4], t[4], z3[4];
    fq_sqr(A, p->x);
    fq_sqr(B, p->y);
    fq_sqr(C, B);
    fq_add(t, p->x, B);
    fq_sqr(t, t);
    fq_sub(t, t, A);
    fq_sub(t, t, C);
    fq_dbl(D, t);
    fq_dbl(E, A);
    fq_add(E, E, A);
    fq_sqr(Fv, E);
    fq_mul(z3, p->y, p->z);
    fq_dbl(z3, z3);
    fq_sub(t, Fv, D);
    fq_sub(r->x, t, D);
    fq_sub(t, D, r->x);
    fq_mul(t, E, t);
    fq_dbl(C, C);
    fq_dbl(C, C);
    fq_dbl(C, C);
    fq_sub(r->y, t, C);
    fq_copy(r->z, z3);
}

/* add-2007-bl: 11M + 5S */
static void jac_add(jac_t *r, const jac_t *p, const jac_t *q) {
    if (jac_is_inf(p)) { *r = *q; return; }
    if (jac_is_inf(q)) { *r = *p; return; }
    u64 z1z1[4], z2z2[4], u1[4], u2[4], s1[4], s2[4], h[4], i[4], j[4], rr[4], v[4], t[4];
    fq_sqr(z1z1, p->z);
    fq_sqr(z2z2, q->z);
    fq_mul(u1, p->x, z2z2);
    fq_mul(u2, q->x, z1z1);
    fq_mul(s1, p->y, q->z);
    fq_mul(s1, s1, z2z2);
    fq_mul(s2, q->y, p->z);
    fq_mul(s2, s2, z1z1);
    if (fq_eq(u1, u2)) {
        if (fq_eq(s1, s2)) { jac_double(r, p); return; }
        jac_set_inf(r);
        return;
    }
    fq_sub(h, u2, u1);
    fq_dbl(i, h);
    fq_sqr(i, i);
    fq_mul(j, h, i);
    fq_sub(rr, s2, s1);
    fq_dbl(rr, rr);
    fq_mul(v, u1, i);
    jac_t o;
    fq_sqr(o.x, rr);
    fq_sub(o.x, o.x, j);
    fq_sub(o.x, o.x, v);
    fq_sub(o.x, o.x, v);
    fq_sub(t, v, o.x);
    fq_mul(t, rr, t);
    fq_mul(s1, s1, j);
    fq_dbl(s1, s1);
    fq_sub(o.y, t, s1);
    fq_add(t, p->z, q->z);
    fq_sqr(t, t);
    fq_sub(t, t, z1z1);
    fq_sub(t, t, z2z2);
    fq_mul(o.z, t, h);
    *r = o;
}

/* madd-2007-bl: 7M + 4S */
static void jac_add_affine(jac_t *r, const jac_t *p, const aff_t *q) {
    if (jac_is_inf(p)) {
        fq_copy(r->x, q->x);
        fq_copy(r->y, q->y);
        fq_one(r->z);
        return;
    }
    u64 z1z1[4], u2[4], s2[4], h[4], hh[4], i[4], j[4], rr[4], v[4], t[4];
    fq_sqr(z1z1, p->z);
    fq_mul(u2, q->x, z1z1);
    fq_mul(s2, q->y, p->z);
    fq_mul(s2, s2, z1z1);
    if (fq_eq(p->x, u2)) {
        if (fq_eq(p->y, s2)) { jac_double(r, p); return; }
        jac_set_inf(r);
        return;
    }
    fq_sub(h, u2, p->x);
    fq_sqr(hh, h);
    fq_dbl(i, hh);
    fq_dbl(i, i);
    fq_mul(j, h, i);
    fq_sub(rr, s2, p->y);
    fq_dbl(rr, rr);
    fq_mul(v, p->x, i);
    jac_t o;
    fq_sqr(o.x, rr);
    fq_sub(o.x, o.x, j);
    fq_sub(o.x, o.x, v);
    fq_sub(o.x, o.x, v);
    fq_sub(t, v, o.x);
    fq_mul(t, rr, t);
    fq_mul(j, p->y, j);
    fq_dbl(j, j);
    fq_sub(o.y, t, j);
    fq_add(t, p->z, h);
    fq_sqr(t, t);
    fq_sub(t, t, z1z1);
    fq_sub(o.z, t, hh);
    *r = o;
}

static void jac_neg(jac_t *r, const jac_t *p) {
    *r = *p;
    fq_neg(r->y, p->y);
}

/* `Projective * Fr` = mul_bigint: MSB-first double-and-add over the canonical scalar bits
 * (the per-element operation of the fold at pcdl.rs:218 and of main.rs:31). */
static void jac_mul(jac_t *r, const jac_t *p, const u64 k_mont[4]) {
    u64 k[4];
    fr_to_canon(k, k_mont);
    jac_t acc;
    jac_set_inf(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        int bit = (k[i >> 6] >> (i & 63)) & 1;
        if (started) jac_double(&acc, &acc);
        if (bit) {
            jac_add(&acc, &acc, p);
            started = 1;
        }
    }
    *r = acc;
}

/* into_affine: one field inversion (group.rs:19). Returns the infinity flag. */
static int jac_to_affine(aff_t *a, const jac_t *p) {
    if (jac_is_inf(p)) {
        fq_zero(a->x);
        fq_zero(a->y);
        return 1;
    }
    u64 zi[4], zi2[4], zi3[4];
    fq_inv(zi, p->z);
    fq_sqr(zi2, zi);
    fq_mul(zi3, zi2, zi);
    fq_mul(a->x, p->x, zi2);
    fq_mul(a->y, p->y, zi3);
    return 0;
}
static void jac_from_affine(jac_t *p, const aff_t *a, int inf) {
    if (inf) { jac_set_inf(p); return; }
    fq_copy(p->x, a->x);
    fq_copy(p->y, a->y);
    fq_one(p->z);
}
/* Projective equality: cross-multiplied comparison (representation independent). */
static int jac_eq(const jac_t *a, const jac_t *b) {
    int ia = jac_is_inf(a), ib = jac_is_inf(b);
    if (ia || ib) return ia && ib;
    u64 z1z1[4], z2z2[4], l[4], r[4];
    fq_sqr(z1z1, a->z);
    fq_sqr(z2z2, b->z);
    fq_mul(l, a->x, z2z2);
    fq_mul(r, b->x, z1z1);
    if (!fq_eq(l, r)) return 0;
    fq_mul(l, a->y, z2z2);
    fq_mul(l, l, b->z);
    fq_mul(r, b->y, z1z1);
    fq_mul(r, r, a->z);
    return fq_eq(l, r);
}

void orc_pt_add(const u64 a[12], const u64 b[12], u64 r[12]) { jac_t o; jac_add(&o, (const jac_t *)a, (const jac_t *)b); memcpy(r, &o, 96); }
void orc_pt_add_affine(const u64 a[12], const u64 b[8], int b_inf, u64 r[12]) {
    jac_t o;
    if (b_inf) o = *(const jac_t *)a; else jac_add_affine(&o, (const jac_t *)a, (const aff_t *)b);
    memcpy(r, &o, 96);
}
void orc_pt_double(const u64 a[12], u64 r[12]) { jac_t o; jac_double(&o, (const jac_t *)a); memcpy(r, &o, 96); }
void orc_pt_mul(const u64 p[12], const u64 k[4], u64 r[12]) { jac_t o; jac_mul(&o, (const jac_t *)p, k); memcpy(r, &o, 96); }
int orc_pt_eq(const u64 a[12], const u64 b[12]) { return jac_eq((const jac_t *)a, (const jac_t *)b); }
int orc_pt_to_affine(const u64 p[12], u64 aff[8]) { return jac_to_affine((aff_t *)aff, (const jac_t *)p); }
void orc_pt_from_affine(const u64 aff[8], int inf, u64 p[12]) { jac_from_affine((jac_t *)p, (const aff_t *)aff, inf); }
int orc_pt_on_curve_affine(const u64 aff[8]) {
    u64 l[4], r[4];
    fq_sqr(l, aff + 4);
    fq_sqr(r, aff);
    fq_mul(r, r, aff);
    fq_add(r, r, FQ_FIVE);
    return fq_eq(l, r);
}

/* ark-serialize `serialize_compressed` of a short-Weierstrass point (ark-ec 0.5 models/short_weierstrass):
 * normalise to affine; write x via Fp::serialize_with_flags with SWFlags (2 flag bits).  The output length
 * is ceil((MODULUS_BIT_SIZE + 2) / 8) = ceil(257 / 8) = 33 bytes for Pallas: 32 bytes little-endian
 * canonical x then one byte that carries only the flags: bit7 = "y is negative" (y > -y as canonical
 * integers), bit6 = infinity (x written as 0).  [arkworks source not in the reference tree -- restated
 * from the published crate; parity unpinned, see header.] */
static void jac_serialize_compressed(const jac_t *p, u8 out[33]) {
    aff_t a;
    int inf = jac_to_affine(&a, p);
    memset(out, 0, 33);
    if (inf) { out[32] = 0x40; return; }
    fq_to_le_bytes(out, a.x);
    u64 y[4], ny[4], nym[4];
    fq_to_canon(y, a.y);
    fq_neg(nym, a.y);
    fq_to_canon(ny, nym);
    int y_gt = 0;
    for (int i = 3; i >= 0; i--) {
        if (y[i] > ny[i]) { y_gt = 1; break; }
        if (y[i] < ny[i]) break;
    }
    if (y_gt) out[32] = 0x80;
}
void orc_pt_serialize_compressed(const u64 p[12], u8 out[33]) { jac_serialize_compressed((const jac_t *)p, out); }

/* ------------------------------------------------------------------------------------------ */
/* Fiat-Shamir: group.rs:41-64 (rho_0!, tag 0) and :66-89 (rho_1!, tag 1).                     */
/* data = concat(serialize_compressed(arg)); digest = SHA3-256(data || u32_le(tag));           */
/* challenge = from_le_bytes_mod_order(digest) in Fr.                                          */
/* ------------------------------------------------------------------------------------------ */
typedef struct { u8 *buf; u64 len, cap; } tr_t;
static void tr_init(tr_t *t) { t->cap = 256; t->len = 0; t->buf = (u8 *)malloc(t->cap); }
static void tr_bytes(tr_t *t, const u8 *b, u64 n) {
    if (t->len + n > t->cap) {
        while (t->len + n > t->cap) t->cap *= 2;
        t->buf = (u8 *)realloc(t->buf, t->cap);
    }
    memcpy(t->buf + t->len, b, n);
    t->len += n;
}
static void tr_point(tr_t *t, const jac_t *p) { u8 b[33]; jac_serialize_compressed(p, b); tr_bytes(t, b, 33); }
static void tr_scalar(tr_t *t, const u64 s[4]) { u8 b[32]; fr_to_le_bytes(b, s); tr_bytes(t, b, 32); }
static void tr_u64(tr_t *t, u64 v) { u8 b[8]; for (int i = 0; i < 8; i++) b[i] = (u8)(v >> (8 * i)); tr_bytes(t, b, 8); }
static void tr_u8(tr_t *t, u8 v) { tr_bytes(t, &v, 1); }
static void tr_finish(tr_t *t, uint32_t tag, u64 out[4]) {
    u8 tb[4] = {(u8)tag, (u8)(tag >> 8), (u8)(tag >> 16), (u8)(tag >> 24)};
    tr_bytes(t, tb, 4);
    u8 dg[32];
    orc_sha3_256(t->buf, t->len, dg);
    fr_from_le_bytes_mod_order(out, dg);
    free(t->buf);
}

/* ------------------------------------------------------------------------------------------ */
/* public parameters: main.rs:18-45                                                            */
/* ------------------------------------------------------------------------------------------ */
static const char GENESIS[] = "To understand recursion, one must first understand recursion";

/* main.rs:18-32 get_generator_hash */
static void generator_hash(u64 k, jac_t *out) {
    u8 msg[sizeof(GENESIS) - 1 + 8];
    memcpy(msg, GENESIS, sizeof(GENESIS) - 1);
    for (int i = 0; i < 8; i++) msg[sizeof(GENESIS) - 1 + i] = (u8)(k >> (8 * i)); /* usize::to_le_bytes */
    u8 dg[32];
    orc_sha3_256(msg, sizeof msg, dg);
    u64 s[4];
    fr_from_le_bytes_mod_order(s, dg);
    jac_t g;
    fq_copy(g.x, GEN_X);
    fq_copy(g.y, GEN_Y);
    fq_one(g.z);
    jac_mul(out, &g, s);
}

void orc_derive_points(u64 start, u64 count, u64 *out_affine) {
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)count; i++) {
        jac_t p;
        generator_hash(start + (u64)i, &p);
        aff_t a;
        jac_to_affine(&a, &p);
        memcpy(out_affine + 8 * i, &a, 64);
    }
}

/* Batched normalisation (Montgomery's trick): one inversion per block instead of one per point. */
static void jac_batch_to_affine(aff_t *out, const jac_t *in, int cnt) {
    u64 pre[256][4], run[4], inv[4], zi[4], zi2[4], zi3[4];
    fq_one(run);
    for (int i = 0; i < cnt; i++) {
        fq_copy(pre[i], run);
        if (!jac_is_inf(&in[i])) fq_mul(run, run, in[i].z);
    }
    fq_inv(inv, run);
    for (int i = cnt - 1; i >= 0; i--) {
        if (jac_is_inf(&in[i])) {
            fq_zero(out[i].x);
            fq_zero(out[i].y);
            continue;
        }
        fq_mul(zi, inv, pre[i]);
        fq_mul(inv, inv, in[i].z);
        fq_sqr(zi2, zi);
        fq_mul(zi3, zi2, zi);
        fq_mul(out[i].x, in[i].x, zi2);
        fq_mul(out[i].y, in[i].y, zi3);
    }
}

/* The same points by a fixed-base table of the generator (16 windows of 16 bits: e * 65536^w * (-1, 2), 64 MiB), 16 mixed
 * additions per point instead of a 255-step double-and-add, normalised in blocks: the CPU arm of the benchmark needs 2^24
 * bases and must not spend minutes deriving them.  tests/test_oracle_golden.py checks it against orc_derive_points and
 * consts.rs. */
void orc_derive_points_fast(u64 start, u64 count, u64 *out_affine) {
    enum { TW = 16, TE = 65535, BLK = 256 };
    aff_t *table = (aff_t *)malloc((size_t)TW * TE * sizeof(aff_t));
    jac_t *col = (jac_t *)malloc((size_t)TE * sizeof(jac_t));
    jac_t base;
    fq_copy(base.x, GEN_X);
    fq_copy(base.y, GEN_Y);
    fq_one(base.z);
    for (int w = 0; w < TW; w++) {
        jac_t acc = base;
        for (int e = 0; e < TE; e++) { /* (e + 1) * 65536^w * G */
            col[e] = acc;
            jac_add(&acc, &acc, &base);
        }
        base = acc; /* 65536 * previous base */
#pragma omp parallel for schedule(static)
        for (int e0 = 0; e0 < TE; e0 += BLK) jac_batch_to_affine(&table[(size_t)w * TE + e0], &col[e0], TE - e0 < BLK ? TE - e0 : BLK);
    }
    free(col);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long i0 = 0; i0 < (long long)count; i0 += BLK) {
        jac_t pts[BLK];
        int cnt = (long long)count - i0 < BLK ? (int)((long long)count - i0) : BLK;
        for (int j = 0; j < cnt; j++) {
            u64 k = start + (u64)i0 + (u64)j;
            u8 msg[sizeof(GENESIS) - 1 + 8];
            memcpy(msg, GENESIS, sizeof(GENESIS) - 1);
            for (int b = 0; b < 8; b++) msg[sizeof(GENESIS) - 1 + b] = (u8)(k >> (8 * b));
            u8 dg[32];
            orc_sha3_256(msg, sizeof msg, dg);
            u64 sm[4], sc[4];
            fr_from_le_bytes_mod_order(sm, dg);
            fr_to_canon(sc, sm);
            jac_set_inf(&pts[j]);
            for (int w = 0; w < TW; w++) {
                unsigned e = (unsigned)(sc[w >> 2] >> (16 * (w & 3))) & 0xffffu;
                if (e) jac_add_affine(&pts[j], &pts[j], &table[(size_t)w * TE + (e - 1)]);
            }
        }
        jac_batch_to_affine((aff_t *)(out_affine + 8 * (u64)i0), pts, cnt);
    }
    free(table);
}

/* Size-independent property of an MSM over DERIVED generators: G_i = s_{i+2} * (-1, 2) (main.rs:18-32), hence
 *   sum_i a_i G_{first+i} = (sum_i a_i s_{first+i+2} mod r) * (-1, 2)
 * -- one SHA3 and one Fr multiplication per point instead of a Pippenger.  Used to check MSM results at sizes and shard
 * layouts where running the full CPU MSM on every rank would take minutes. */
void orc_msm_derived_by_dlog(u64 first, const u64 *scalars, u64 n, int threads, u64 out[12]) {
    if (threads < 1) threads = 1;
    u64 *partial = (u64 *)calloc((size_t)threads * 4, sizeof(u64));
#pragma omp parallel num_threads(threads)
    {
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
        u64 acc[4], prod[4];
        fr_zero(acc);
        for (u64 i = (u64)t; i < n; i += (u64)nt) {
            u64 k = first + i + 2;
            u8 msg[sizeof(GENESIS) - 1 + 8];
            memcpy(msg, GENESIS, sizeof(GENESIS) - 1);
            for (int b = 0; b < 8; b++) msg[sizeof(GENESIS) - 1 + b] = (u8)(k >> (8 * b));
            u8 dg[32];
            orc_sha3_256(msg, sizeof msg, dg);
            u64 s[4];
            fr_from_le_bytes_mod_order(s, dg);
            fr_mul(prod, s, scalars + 4 * i);
            fr_add(acc, acc, prod);
        }
        fr_copy(partial + 4 * t, acc);
    }
    u64 tot[4];
    fr_zero(tot);
    for (int t = 0; t < threads; t++) fr_add(tot, tot, partial + 4 * t);
    free(partial);
    jac_t g, r;
    fq_copy(g.x, GEN_X);
    fq_copy(g.y, GEN_Y);
    fq_one(g.z);
    jac_mul(&r, &g, tot);
    memcpy(out, &r, 96);
}

static jac_t PP_S, PP_H;
static aff_t *PP_GS = NULL;
static u64 PP_N = 0;

void orc_set_params(const u64 S[12], const u64 H[12], const u64 *gs_affine, u64 n) {
    memcpy(&PP_S, S, 96);
    memcpy(&PP_H, H, 96);
    free(PP_GS);
    PP_GS = (aff_t *)malloc(n * sizeof(aff_t));
    memcpy(PP_GS, gs_affine, n * sizeof(aff_t));
    PP_N = n;
}
/* main.rs:35-45 get_pp */
void orc_derive_params(u64 n) {
    u64 *buf = (u64 *)malloc((n + 2) * 64);
    orc_derive_points(0, n + 2, buf);
    jac_t S, H;
    jac_from_affine(&S, (aff_t *)buf, 0);
    jac_from_affine(&H, (aff_t *)(buf + 8), 0);
    orc_set_params((u64 *)&S, (u64 *)&H, buf + 16, n);
    free(buf);
}
const u64 *orc_params_gs(void) { return (const u64 *)PP_GS; }
void orc_params_SH(u64 S[12], u64 H[12]) { memcpy(S, &PP_S, 96); memcpy(H, &PP_H, 96); }
u64 orc_params_n(void) { return PP_N; }

/* ------------------------------------------------------------------------------------------ */
/* group.rs:24-26 -> ark-ec VariableBaseMSM::msm_unchecked (msm_bigint_wnaf shape):            */
/* c = 3 if n < 32 else floor(ceil(log2 n) * 69 / 100) + 2; signed radix-2^c digits; per       */
/* window 2^(c-1) Jacobian buckets filled by mixed add/sub, running-sum sweep from the top     */
/* bucket, windows combined high -> low with c doublings each.  Serial unless threads > 1.     */
/* ------------------------------------------------------------------------------------------ */
static int ceil_log2(u64 n) { int l = 0; while (((u64)1 << l) < n) l++; return l; }

/* One window of arkworks' msm_bigint_wnaf: bucket accumulation over the points [lo, hi), then (reduce != 0) the running-sum
 * sweep from the top bucket.  reduce == 0 leaves the raw buckets in `buckets` for a merge (see orc_msm_affine). */
static void msm_window_fill(const aff_t *bases, const u8 *inf, const int32_t *digits, u64 lo, u64 hi, int W, int w, jac_t *buckets, u64 nb) {
    for (u64 b = 0; b < nb; b++) jac_set_inf(&buckets[b]);
    for (u64 i = lo; i < hi; i++) {
        int32_t d = digits[i * W + w];
        if (d == 0 || (inf && inf[i])) continue;
        if (d > 0) {
            jac_add_affine(&buckets[d - 1], &buckets[d - 1], &bases[i]);
        } else {
            aff_t nq = bases[i];
            fq_neg(nq.y, bases[i].y);
            jac_add_affine(&buckets[-d - 1], &buckets[-d - 1], &nq);
        }
    }
}
static void msm_window_sweep(const jac_t *buckets, u64 nb, jac_t *out) {
    jac_t running, res;
    jac_set_inf(&running);
    jac_set_inf(&res);
    for (u64 b = nb; b-- > 0;) {
        jac_add(&running, &running, &buckets[b]);
        jac_add(&res, &res, &running);
    }
    *out = res;
}
static void msm_window(const aff_t *bases, const u8 *inf, const int32_t *digits, u64 n, int W, int w, int c, jac_t *out) {
    u64 nb = (u64)1 << (c - 1);
    jac_t *buckets = (jac_t *)malloc(nb * sizeof(jac_t));
    msm_window_fill(bases, inf, digits, 0, n, W, w, buckets, nb);
    msm_window_sweep(buckets, nb, out);
    free(buckets);
}

void orc_msm_affine(const u64 *bases_affine, const u8 *inf, const u64 *scalars, u64 n, int threads, u64 out[12]) {
    jac_t total;
    jac_set_inf(&total);
    if (n == 0) { memcpy(out, &total, 96); return; }
    int c = n < 32 ? 3 : (ceil_log2(n) * 69 / 100) + 2;
    int W = 255 / c + 1; /* c*W >= 256: the top window always absorbs the last carry */
    int32_t *digits = (int32_t *)malloc(n * (u64)W * sizeof(int32_t));
#pragma omp parallel for if (threads > 1) num_threads(threads > 1 ? threads : 1)
    for (long long i = 0; i < (long long)n; i++) {
        u64 k[5];
        fr_to_canon(k, scalars + 4 * i); /* into_bigint */
        k[4] = 0;
        int carry = 0;
        for (int w = 0; w < W; w++) {
            int bit = w * c;
            u64 v = k[bit >> 6] >> (bit & 63);
            if ((bit & 63) + c > 64) v |= k[(bit >> 6) + 1] << (64 - (bit & 63));
            int64_t d = (int64_t)(v & (((u64)1 << c) - 1)) + carry;
            carry = 0;
            if (d > ((int64_t)1 << (c - 1))) { d -= (int64_t)1 << c; carry = 1; }
            digits[i * W + w] = (int32_t)d;
        }
    }
    jac_t *wsum = (jac_t *)malloc(W * sizeof(jac_t));
    /* arkworks' `parallel` feature runs one window per thread (the reference leaves it off: threads <= 1).  With more
     * threads than windows (15 windows at n = 2^24, 16-32 host threads), or fewer windows than a multiple of the thread
     * count, whole cores would idle; so each window is also cut into S point chunks with their own bucket arrays, merged
     * bucket by bucket before the window's running-sum sweep.  Same group element; only the CPU baseline's utilisation
     * changes. */
    int S = 1;
    if (threads > 1 && n >= ((u64)1 << 16)) {
        S = (threads + W - 1) / W;
        if (W % threads != 0 && S < 2) S = 2;
        if (S > 8) S = 8;
    }
    if (S == 1) {
#pragma omp parallel for schedule(dynamic, 1) if (threads > 1) num_threads(threads > 1 ? threads : 1)
        for (int w = 0; w < W; w++) msm_window((const aff_t *)bases_affine, inf, digits, n, W, w, c, &wsum[w]);
    } else {
        u64 nb = (u64)1 << (c - 1);
        jac_t *bk = (jac_t *)malloc((u64)W * S * nb * sizeof(jac_t));
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int t = 0; t < W * S; t++) {
            int w = t / S, sidx = t % S;
            u64 lo = n * (u64)sidx / S, hi = n * (u64)(sidx + 1) / S;
            msm_window_fill((const aff_t *)bases_affine, inf, digits, lo, hi, W, w, bk + (u64)t * nb, nb);
        }
#pragma omp parallel for schedule(static) num_threads(threads)
        for (long long j = 0; j < (long long)((u64)W * nb); j++) {
            u64 w = (u64)j / nb, b = (u64)j % nb;
            jac_t *dst = bk + (w * S) * nb + b;
            for (int sidx = 1; sidx < S; sidx++) jac_add(dst, dst, bk + (w * S + sidx) * nb + b);
        }
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
        for (int w = 0; w < W; w++) msm_window_sweep(bk + ((u64)w * S) * nb, nb, &wsum[w]);
        free(bk);
    }
    for (int w = W - 1; w >= 1; w--) {
        jac_add(&total, &total, &wsum[w]);
        for (int k = 0; k < c; k++) jac_double(&total, &total);
    }
    jac_add(&total, &total, &wsum[0]);
    free(wsum);
    free(digits);
    memcpy(out, &total, 96);
}

void orc_msm_naive(const u64 *bases_affine, const u8 *inf, const u64 *scalars, u64 n, u64 out[12]) {
    jac_t total;
    jac_set_inf(&total);
    for (u64 i = 0; i < n; i++) {
        if (inf && inf[i]) continue;
        jac_t b, t;
        jac_from_affine(&b, (const aff_t *)(bases_affine + 8 * i), 0);
        jac_mul(&t, &b, scalars + 4 * i);
        jac_add(&total, &total, &t);
    }
    memcpy(out, &total, 96);
}

/* group.rs:18-21 point_dot: one into_affine (one inversion) per element, then the MSM */
void orc_point_dot(const u64 *scalars, const u64 *points_jac, u64 n, int threads, u64 out[12]) {
    aff_t *aff = (aff_t *)malloc((n ? n : 1) * sizeof(aff_t));
    u8 *inf = (u8 *)malloc(n ? n : 1);
#pragma omp parallel for if (threads > 1) num_threads(threads > 1 ? threads : 1)
    for (long long i = 0; i < (long long)n; i++) inf[i] = (u8)jac_to_affine(&aff[i], (const jac_t *)(points_jac + 12 * i));
    orc_msm_affine((const u64 *)aff, inf, scalars, n, threads, out);
    free(aff);
    free(inf);
}

/* group.rs:13-15 */
void orc_scalar_dot(const u64 *xs, const u64 *ys, u64 n, u64 out[4]) {
    u64 acc[4], t[4];
    fr_zero(acc);
    for (u64 i = 0; i < n; i++) {
        fr_mul(t, xs + 4 * i, ys + 4 * i);
        fr_add(acc, acc, t);
    }
    fr_copy(out, acc);
}
/* group.rs:29-37 */
void orc_construct_powers(const u64 z[4], u64 n, u64 *out) {
    u64 cur[4];
    fr_one(cur);
    for (u64 i = 0; i < n; i++) {
        fr_copy(out + 4 * i, cur);
        fr_mul(cur, cur, z);
    }
}

/* pedersen.rs:6-20 */
int orc_pedersen_commit(const u64 *w, const u64 *gs_affine, u64 n_gs, const u64 *ms, u64 n_ms, int threads, u64 out[12]) {
    if (n_gs != n_ms) return ORC_ELEN; /* pedersen.rs:7-12 assert */
    jac_t acc;
    orc_msm_affine(gs_affine, NULL, ms, n_gs, threads, (u64 *)&acc);
    if (w) {
        jac_t sw;
        jac_mul(&sw, &PP_S, w);
        jac_add(&acc, &sw, &acc);
    }
    memcpy(out, &acc, 96);
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------ */
/* pcdl.rs                                                                                     */
/* ------------------------------------------------------------------------------------------ */
static int is_pow2(u64 n) { return n && !(n & (n - 1)); }

/* pcdl.rs:56-77. Multiplying h (degree < 2^i) by (1 + xi X^{2^i}) is h + xi * X^{2^i} * h with no
 * overlap, so the dense product the reference computes equals this in-place doubling. */
void orc_h_get_poly(const u64 *xis, uint32_t lg_n, u64 *out) {
    fr_one(out);
    for (uint32_t i = 0; i < lg_n; i++) {
        u64 power = (u64)1 << i;
        const u64 *xi = xis + 4 * (lg_n - i);
        for (u64 j = 0; j < power; j++) fr_mul(out + 4 * (j + power), out + 4 * j, xi);
    }
}
/* pcdl.rs:79-91 */
void orc_h_eval(const u64 *xis, uint32_t lg_n, const u64 z[4], u64 out[4]) {
    u64 one[4], v[4], zi[4], t[4];
    fr_one(one);
    fr_mul(t, xis + 4 * lg_n, z);
    fr_add(v, one, t);
    fr_copy(zi, z);
    for (uint32_t i = 1; i < lg_n; i++) {
        fr_sqr(zi, zi);
        fr_mul(t, xis + 4 * (lg_n - i), zi);
        fr_add(t, one, t);
        fr_mul(v, v, t);
    }
    fr_copy(out, v);
}

static u64 poly_degree(const u64 *coeffs, u64 n) { /* DensePolynomial::degree after trimming */
    while (n > 0 && fr_is_zero(coeffs + 4 * (n - 1))) n--;
    return n ? n - 1 : 0;
}

/* pcdl.rs:99-110 */
int orc_pcdl_commit(const u64 *coeffs, u64 n_coeffs, u64 d, const u64 *w, int threads, u64 out[12]) {
    u64 n = d + 1;
    if (!is_pow2(n) || n > PP_N) return ORC_EINVAL;
    if (n_coeffs && poly_degree(coeffs, n_coeffs) > d) return ORC_EINVAL;
    u64 *padded = (u64 *)calloc(n * 4, sizeof(u64));
    u64 ncopy = n_coeffs < n ? n_coeffs : n;
    memcpy(padded, coeffs, ncopy * 32);
    int rc = orc_pedersen_commit(w, (const u64 *)PP_GS, n, padded, n, threads, out);
    free(padded);
    return rc;
}

/* pcdl.rs:120-242 */
int orc_pcdl_open(const u64 *p_coeffs, u64 n_coeffs, const u64 C[12], u64 d, const u64 z[4], const u64 *w,
                  const u64 *q_coeffs, u64 n_q, const u64 *w_bar, int threads, orc_eval_proof *pi) {
    u64 n = d + 1;
    if (!is_pow2(n) || n > PP_N) return ORC_EINVAL;
    u64 deg = poly_degree(p_coeffs, n_coeffs);
    if (deg > d) return ORC_EINVAL;
    uint32_t lg_n = (uint32_t)ceil_log2(n);
    memset(pi, 0, sizeof *pi);
    pi->lg_n = lg_n;

    u64 *cs = (u64 *)calloc(n * 4, sizeof(u64));
    memcpy(cs, p_coeffs, (n_coeffs < n ? n_coeffs : n) * 32);
    u64 *zs = (u64 *)malloc(n * 32);
    orc_construct_powers(z, n, zs); /* :186 */

    /* 1. v = p(z)  (:135) */
    u64 v[4];
    orc_scalar_dot(cs, zs, n, v);

    jac_t C_prime = *(const jac_t *)C;
    if (w) { /* :137-164 hiding */
        if (n_q != deg || deg == 0) { free(cs); free(zs); return ORC_ELEN; }
        /* p_bar = q * (X - z) (:140-142) */
        u64 *pbar = (u64 *)calloc(n * 4, sizeof(u64));
        u64 t[4];
        for (u64 i = 0; i <= n_q; i++) {
            u64 acc[4];
            fr_zero(acc);
            if (i >= 1) fr_copy(acc, q_coeffs + 4 * (i - 1));
            if (i < n_q) {
                fr_mul(t, z, q_coeffs + 4 * i);
                fr_sub(acc, acc, t);
            }
            fr_copy(pbar + 4 * i, acc);
        }
        /* C_bar = commit(p_bar, d, w_bar) (:149) */
        jac_t C_bar;
        orc_pcdl_commit(pbar, n, d, w_bar, threads, (u64 *)&C_bar);
        /* alpha = rho_0(C, z, v, C_bar) (:153) */
        u64 a[4];
        tr_t tr;
        tr_init(&tr);
        tr_point(&tr, (const jac_t *)C);
        tr_scalar(&tr, z);
        tr_scalar(&tr, v);
        tr_point(&tr, &C_bar);
        tr_finish(&tr, 0, a);
        /* p' = p + alpha * p_bar (:156) */
        for (u64 i = 0; i < n; i++) {
            fr_mul(t, pbar + 4 * i, a);
            fr_add(cs + 4 * i, cs + 4 * i, t);
        }
        free(pbar);
        /* w' = w_bar * alpha + w (:159) */
        u64 wp[4];
        fr_mul(wp, w_bar, a);
        fr_add(wp, wp, w);
        /* C' = C + C_bar * alpha - S * w' (:162) */
        jac_t t1, t2;
        jac_mul(&t1, &C_bar, a);
        jac_add(&C_prime, (const jac_t *)C, &t1);
        jac_mul(&t2, &PP_S, wp);
        jac_neg(&t2, &t2);
        jac_add(&C_prime, &C_prime, &t2);
        pi->hiding = 1;
        memcpy(pi->C_bar, &C_bar, 96);
        fr_copy(pi->w_prime, wp);
    }

    /* xi_0 = rho_0(C', z, v); H' = H * xi_0 (:180-181) */
    u64 xi[4];
    {
        tr_t tr;
        tr_init(&tr);
        tr_point(&tr, &C_prime);
        tr_scalar(&tr, z);
        tr_scalar(&tr, v);
        tr_finish(&tr, 0, xi);
    }
    jac_t H_prime;
    jac_mul(&H_prime, &PP_H, xi);

    jac_t *gs = (jac_t *)malloc(n * sizeof(jac_t));
    for (u64 i = 0; i < n; i++) jac_from_affine(&gs[i], &PP_GS[i], 0); /* :185 */

    u64 m = n / 2;
    for (uint32_t round = 0; round < lg_n; round++) { /* :195-227 */
        u64 dot[4];
        jac_t L, R, t;
        /* L = <c_r, g_l> + H' * <c_r, z_l> (:203-204) */
        orc_scalar_dot(cs + 4 * m, zs, m, dot);
        orc_point_dot(cs + 4 * m, (const u64 *)gs, m, threads, (u64 *)&L);
        jac_mul(&t, &H_prime, dot);
        jac_add(&L, &L, &t);
        /* R = <c_l, g_r> + H' * <c_l, z_r> (:207-208) */
        orc_scalar_dot(cs, zs + 4 * m, m, dot);
        orc_point_dot(cs, (const u64 *)(gs + m), m, threads, (u64 *)&R);
        jac_mul(&t, &H_prime, dot);
        jac_add(&R, &R, &t);
        memcpy(pi->Ls[round], &L, 96);
        memcpy(pi->Rs[round], &R, 96);
        /* xi_{i+1} = rho_0(xi_i, L, R) (:212) */
        u64 xin[4], xin_inv[4];
        tr_t tr;
        tr_init(&tr);
        tr_scalar(&tr, xi);
        tr_point(&tr, &L);
        tr_point(&tr, &R);
        tr_finish(&tr, 0, xin);
        fr_inv(xin_inv, xin);
        fr_copy(xi, xin);
        /* fold (:216-224) */
#pragma omp parallel for if (threads > 1) num_threads(threads > 1 ? threads : 1)
        for (long long j = 0; j < (long long)m; j++) {
            jac_t gp;
            u64 tt[4];
            jac_mul(&gp, &gs[j + m], xin);
            jac_add(&gs[j], &gs[j], &gp);
            fr_mul(tt, cs + 4 * (j + m), xin_inv);
            fr_add(cs + 4 * j, cs + 4 * j, tt);
            fr_mul(tt, zs + 4 * (j + m), xin);
            fr_add(zs + 4 * j, zs + 4 * j, tt);
        }
        m /= 2;
    }
    memcpy(pi->U, &gs[0], 96); /* :230 */
    fr_copy(pi->c, cs);        /* :231 */
    free(gs);
    free(cs);
    free(zs);
    return ORC_OK;
}

/* pcdl.rs:252-314 */
int orc_pcdl_succinct_check(const u64 C[12], u64 d, const u64 z[4], const u64 v[4], const orc_eval_proof *pi,
                            u64 *xis_out, u64 U_out[12]) {
    u64 n = d + 1;
    if (!is_pow2(n) || n > PP_N) return ORC_EINVAL; /* :261-262 */
    uint32_t lg_n = (uint32_t)ceil_log2(n);
    if (pi->lg_n != lg_n) return ORC_EINVAL; /* Ls[i] would index out of bounds in the reference */
    jac_t C_prime = *(const jac_t *)C;
    if (pi->hiding) { /* :272-279 */
        u64 a[4];
        tr_t tr;
        tr_init(&tr);
        tr_point(&tr, (const jac_t *)C);
        tr_scalar(&tr, z);
        tr_scalar(&tr, v);
        tr_point(&tr, (const jac_t *)pi->C_bar);
        tr_finish(&tr, 0, a);
        jac_t t1, t2;
        jac_mul(&t1, (const jac_t *)pi->C_bar, a);
        jac_add(&C_prime, &C_prime, &t1);
        jac_mul(&t2, &PP_S, pi->w_prime);
        jac_neg(&t2, &t2);
        jac_add(&C_prime, &C_prime, &t2);
    }
    /* :282-285 */
    u64 *xis = (u64 *)malloc((lg_n + 1) * 32);
    {
        tr_t tr;
        tr_init(&tr);
        tr_point(&tr, &C_prime);
        tr_scalar(&tr, z);
        tr_scalar(&tr, v);
        tr_finish(&tr, 0, xis);
    }
    jac_t H_prime, C_i, t;
    jac_mul(&H_prime, &PP_H, xis);
    /* :288 */
    jac_mul(&t, &H_prime, v);
    jac_add(&C_i, &C_prime, &t);
    /* :291-298 */
    for (uint32_t i = 0; i < lg_n; i++) {
        u64 *xn = xis + 4 * (i + 1), xinv[4];
        tr_t tr;
        tr_init(&tr);
        tr_scalar(&tr, xis + 4 * i);
        tr_point(&tr, (const jac_t *)pi->Ls[i]);
        tr_point(&tr, (const jac_t *)pi->Rs[i]);
        tr_finish(&tr, 0, xn);
        fr_inv(xinv, xn);
        jac_t a, b;
        jac_mul(&a, (const jac_t *)pi->Ls[i], xinv);
        jac_mul(&b, (const jac_t *)pi->Rs[i], xn);
        jac_add(&a, &a, &b);
        jac_add(&C_i, &C_i, &a);
    }
    /* :301-304 */
    u64 hz[4], vp[4];
    orc_h_eval(xis, lg_n, z, hz);
    fr_mul(vp, pi->c, hz);
    /* :307-310 */
    jac_t rhs, t2;
    jac_mul(&rhs, (const jac_t *)pi->U, pi->c);
    jac_mul(&t2, &H_prime, vp);
    jac_add(&rhs, &rhs, &t2);
    int ok = jac_eq(&C_i, &rhs);
    if (ok) {
        if (xis_out) memcpy(xis_out, xis, (lg_n + 1) * 32);
        if (U_out) memcpy(U_out, pi->U, 96);
    }
    free(xis);
    return ok ? ORC_OK : ORC_REJECT_SUCCINCT;
}

/* pcdl.rs:323-342 */
int orc_pcdl_check(const u64 C[12], u64 d, const u64 z[4], const u64 v[4], const orc_eval_proof *pi, int threads) {
    u64 n = d + 1;
    if (!is_pow2(n) || n > PP_N) return ORC_EINVAL;
    uint32_t lg_n = (uint32_t)ceil_log2(n);
    u64 *xis = (u64 *)malloc((lg_n + 1) * 32);
    jac_t U;
    int rc = orc_pcdl_succinct_check(C, d, z, v, pi, xis, (u64 *)&U);
    if (rc) { free(xis); return rc; }
    u64 *h = (u64 *)malloc(n * 32);
    orc_h_get_poly(xis, lg_n, h);
    jac_t comm;
    orc_pedersen_commit(NULL, (const u64 *)PP_GS, n, h, n, threads, (u64 *)&comm); /* :338 */
    free(h);
    free(xis);
    return jac_eq(&U, &comm) ? ORC_OK : ORC_REJECT_U; /* :339 */
}

/* ------------------------------------------------------------------------------------------ */
/* acc.rs                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    u64 C_bar[12];
    u64 z[4];
    u64 alpha[4];
    u64 *alphas; /* [m+1][4] */
    u64 *xis;    /* [m][lg_n+1][4] */
    uint32_t lg_n;
    u64 m;
} common_out;

/* acc.rs:135-188 */
static int common_subroutine(u64 d, const orc_instance *qs, u64 m, const u64 h0[2][4], const u64 U0[12],
                             const u64 w[4], int threads, common_out *out) {
    u64 n = d + 1;
    if (!is_pow2(n) || n > PP_N) return ORC_EINVAL;
    uint32_t lg_n = (uint32_t)ceil_log2(n);
    /* (3) U_0 == commit(h_0, d, None) (:152-155) */
    jac_t chk;
    int rc = orc_pcdl_commit((const u64 *)h0, 2, d, NULL, threads, (u64 *)&chk);
    if (rc) return rc;
    if (!jac_eq((const jac_t *)U0, &chk)) return ORC_REJECT_U0;
    jac_t *Us = (jac_t *)malloc((m + 1) * sizeof(jac_t));
    u64 *xis = (u64 *)malloc((m ? m : 1) * (lg_n + 1) * 32);
    Us[0] = *(const jac_t *)U0;
    for (u64 i = 0; i < m; i++) { /* :158-170 */
        rc = orc_pcdl_succinct_check(qs[i].C, qs[i].d, qs[i].z, qs[i].v, &qs[i].pi, xis + i * (lg_n + 1) * 4 * 1, (u64 *)&Us[i + 1]);
        if (rc == ORC_OK && qs[i].d != d) rc = ORC_REJECT_D;
        if (rc) { free(Us); free(xis); return rc; }
    }
    /* alpha = rho_1(hs) (:173): AccumulatedHPolys{h_0: Some(poly), hs: Vec<HPoly{xis}>, alpha: None, alphas: []} */
    tr_t tr;
    tr_init(&tr);
    tr_u8(&tr, 1); /* Option::Some */
    u64 h0len = 2;
    while (h0len > 0 && fr_is_zero(h0[h0len - 1])) h0len--; /* DensePolynomial trims trailing zeros */
    tr_u64(&tr, h0len);
    for (u64 i = 0; i < h0len; i++) tr_scalar(&tr, h0[i]);
    tr_u64(&tr, m);
    for (u64 i = 0; i < m; i++) {
        tr_u64(&tr, lg_n + 1);
        for (uint32_t k = 0; k <= lg_n; k++) tr_scalar(&tr, xis + (i * (lg_n + 1) + k) * 4);
    }
    tr_u8(&tr, 0);  /* alpha: None */
    tr_u64(&tr, 0); /* alphas: empty */
    tr_finish(&tr, 1, out->alpha);
    /* alphas = powers (:79-82), C = point_dot(alphas, Us) (:178) */
    out->alphas = (u64 *)malloc((m + 1) * 32);
    orc_construct_powers(out->alpha, m + 1, out->alphas);
    jac_t Cacc;
    orc_point_dot(out->alphas, (const u64 *)Us, m + 1, threads, (u64 *)&Cacc);
    /* z = rho_1(C, alpha) (:181) */
    tr_init(&tr);
    tr_point(&tr, &Cacc);
    tr_scalar(&tr, out->alpha);
    tr_finish(&tr, 1, out->z);
    /* C_bar = C + S * w (:184) */
    jac_t sw;
    jac_mul(&sw, &PP_S, w);
    jac_add(&Cacc, &Cacc, &sw);
    memcpy(out->C_bar, &Cacc, 96);
    out->xis = xis;
    out->lg_n = lg_n;
    out->m = m;
    free(Us);
    return ORC_OK;
}
static void common_free(common_out *c) { free(c->alphas); free(c->xis); }

/* AccumulatedHPolys::eval acc.rs:97-106 */
static void acc_h_eval(const common_out *c, const u64 h0[2][4], const u64 z[4], u64 out[4]) {
    u64 v[4], t[4];
    fr_mul(t, h0[1], z); /* h_0.evaluate(z), degree 1 */
    fr_add(v, h0[0], t);
    for (u64 i = 0; i < c->m; i++) {
        orc_h_eval(c->xis + i * (c->lg_n + 1) * 4, c->lg_n, z, t);
        fr_mul(t, t, c->alphas + 4 * (i + 1));
        fr_add(v, v, t);
    }
    fr_copy(out, v);
}

/* acc.rs:190-220 */
int orc_acc_prover(u64 d, const orc_instance *qs, u64 m, const u64 h0[2][4], const u64 w[4], const u64 *q_coeffs,
                   u64 n_q, const u64 w_bar[4], int threads, orc_accumulator *acc) {
    u64 n = d + 1;
    memset(acc, 0, sizeof *acc);
    /* U_0 = commit(h_0, d, None) (:195) */
    int rc = orc_pcdl_commit((const u64 *)h0, 2, d, NULL, threads, acc->U0);
    if (rc) return rc;
    memcpy(acc->h0, h0, 64);
    fr_copy(acc->w, w);
    common_out c;
    rc = common_subroutine(d, qs, m, h0, acc->U0, w, threads, &c);
    if (rc) return rc;
    memcpy(acc->C_bar, c.C_bar, 96);
    acc->d = d;
    fr_copy(acc->z, c.z);
    acc_h_eval(&c, h0, c.z, acc->v); /* :205 */
    /* h.get_poly() (:85-94): h_0 + sum alpha^{i+1} * h_i */
    u64 *h = (u64 *)calloc(n * 4, sizeof(u64));
    u64 *hi = (u64 *)malloc(n * 32);
    fr_copy(h, h0[0]);
    if (n > 1) fr_copy(h + 4, h0[1]);
    for (u64 i = 0; i < m; i++) {
        orc_h_get_poly(c.xis + i * (c.lg_n + 1) * 4, c.lg_n, hi);
        u64 t[4];
        for (u64 j = 0; j < n; j++) {
            fr_mul(t, hi + 4 * j, c.alphas + 4 * (i + 1));
            fr_add(h + 4 * j, h + 4 * j, t);
        }
    }
    free(hi);
    rc = orc_pcdl_open(h, n, acc->C_bar, d, acc->z, w, q_coeffs, n_q, w_bar, threads, &acc->pi); /* :209 */
    free(h);
    common_free(&c);
    return rc;
}

/* acc.rs:223-243 */
int orc_acc_verifier(u64 d, const orc_instance *qs, u64 m, const orc_accumulator *acc, int threads) {
    common_out c;
    int rc = common_subroutine(d, qs, m, acc->h0, acc->U0, acc->w, threads, &c);
    if (rc) return rc;
    u64 hv[4];
    acc_h_eval(&c, acc->h0, acc->z, hv);
    if (!jac_eq((const jac_t *)c.C_bar, (const jac_t *)acc->C_bar)) rc = ORC_REJECT_CBAR;
    else if (!fr_eq(c.z, acc->z)) rc = ORC_REJECT_Z;
    else if (d != acc->d) rc = ORC_REJECT_D;
    else if (!fr_eq(hv, acc->v)) rc = ORC_REJECT_V;
    common_free(&c);
    return rc;
}

/* acc.rs:245-255 */
int orc_acc_decider(const orc_accumulator *acc, int threads) {
    return orc_pcdl_check(acc->C_bar, acc->d, acc->z, acc->v, &acc->pi, threads);
}

/* acc.rs:121-131 */
void orc_acc_to_instance(const orc_accumulator *acc, orc_instance *q) {
    memcpy(q->C, acc->C_bar, 96);
    q->d = acc->d;
    fr_copy(q->z, acc->z);
    fr_copy(q->v, acc->v);
    q->pi = acc->pi;
}
