/*
 * oracle/fp_tmpl.h -- TEST INFRASTRUCTURE ONLY (CPU oracle). Not part of the product path.
 *
 * 255-bit prime field, 4 x u64 little-endian limbs, Montgomery form with R = 2^256.
 * This is the in-memory representation arkworks' `Fp<MontBackend<_,4>>` uses and the one the
 * reference prints into code/src/consts.rs (main.rs:47-53, consts.rs:4-21).
 *
 * Include with FP_NAME (prefix) and FP_MOD0..3 (modulus limbs) defined; every constant
 * (R, R^2, -p^-1 mod 2^64) is derived from the modulus at init time by F(init)().
 */
#define FP_CAT_(a, b) a##_##b
#define FP_CAT(a, b) FP_CAT_(a, b)
#define F(name) FP_CAT(FP_NAME, name)

static const u64 F(P)[4] = {FP_MOD0, FP_MOD1, FP_MOD2, FP_MOD3};
static u64 F(INV);    /* -p^-1 mod 2^64 */
static u64 F(R)[4];   /* 2^256 mod p  == Montgomery form of 1 */
static u64 F(R2)[4];  /* 2^512 mod p */

static inline int F(geq_p)(const u64 a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > F(P)[i]) return 1;
        if (a[i] < F(P)[i]) return 0;
    }
    return 1;
}
static inline void F(sub_p)(u64 a[4]) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a[i] - F(P)[i] - br;
        a[i] = (u64)t;
        br = (t >> 64) & 1;
    }
}
static inline void F(copy)(u64 r[4], const u64 a[4]) { r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = a[3]; }
static inline int F(is_zero)(const u64 a[4]) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static inline int F(eq)(const u64 a[4], const u64 b[4]) {
    return ((a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3])) == 0;
}
static inline void F(zero)(u64 r[4]) { r[0] = r[1] = r[2] = r[3] = 0; }
static inline void F(one)(u64 r[4]) { F(copy)(r, F(R)); }

/* p < 2^255 so a + b never carries out of 256 bits. */
static inline void F(add)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u128 c = 0;
    u64 t[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)a[i] + b[i];
        t[i] = (u64)c;
        c >>= 64;
    }
    if (F(geq_p)(t)) F(sub_p)(t);
    F(copy)(r, t);
}
static inline void F(sub)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u128 br = 0;
    u64 t[4];
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - b[i] - br;
        t[i] = (u64)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)t[i] + F(P)[i];
            t[i] = (u64)c;
            c >>= 64;
        }
    }
    F(copy)(r, t);
}
static inline void F(neg)(u64 r[4], const u64 a[4]) {
    if (F(is_zero)(a)) { F(zero)(r); return; }
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)F(P)[i] - a[i] - br;
        r[i] = (u64)d;
        br = (d >> 64) & 1;
    }
}
static inline void F(dbl)(u64 r[4], const u64 a[4]) { F(add)(r, a, a); }

/* CIOS Montgomery multiplication: r = a*b*R^-1 mod p. */
static inline void F(mul)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a[j] * b[i] + t[j];
            t[j] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (u64)c;
        t[5] = (u64)(c >> 64);
        u64 m = t[0] * F(INV);
        c = (u128)m * F(P)[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F(P)[j] + t[j];
            t[j - 1] = (u64)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (u64)c;
        t[4] = t[5] + (u64)(c >> 64);
    }
    u64 o[4] = {t[0], t[1], t[2], t[3]};
    if (t[4] || F(geq_p)(o)) F(sub_p)(o);
    F(copy)(r, o);
}
static inline void F(sqr)(u64 r[4], const u64 a[4]) { F(mul)(r, a, a); }

/* Montgomery form <-> canonical integer (little-endian limbs). */
static inline void F(to_canon)(u64 r[4], const u64 a[4]) {
    static const u64 one[4] = {1, 0, 0, 0};
    F(mul)(r, a, one);
}
static inline void F(from_canon)(u64 r[4], const u64 a[4]) { F(mul)(r, a, F(R2)); }
static inline void F(from_u64)(u64 r[4], u64 v) {
    u64 t[4] = {v, 0, 0, 0};
    F(from_canon)(r, t);
}

/* a^e for a canonical 256-bit exponent e (square-and-multiply, MSB first). */
static void F(pow)(u64 r[4], const u64 a[4], const u64 e[4]) {
    u64 acc[4], base[4];
    F(one)(acc);
    F(copy)(base, a);
    for (int i = 255; i >= 0; i--) {
        F(sqr)(acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) F(mul)(acc, acc, base);
    }
    F(copy)(r, acc);
}
/* Inverse by Fermat (a^(p-2)); returns 0 for a == 0 (callers check). */
static void F(inv)(u64 r[4], const u64 a[4]) {
    u64 e[4] = {F(P)[0] - 2, F(P)[1], F(P)[2], F(P)[3]}; /* p ends in ...0001, no borrow */
    F(pow)(r, a, e);
}

/* Reduce a 256-bit little-endian integer mod p and return it in Montgomery form.
 * This is `from_le_bytes_mod_order` on 32 bytes (group.rs:60, main.rs:28): the result is just
 * the integer mod p; p > 2^254 so at most 3 subtractions are needed. */
static void F(from_le_bytes_mod_order)(u64 r[4], const unsigned char b[32]) {
    u64 t[4];
    for (int i = 0; i < 4; i++) {
        u64 v = 0;
        for (int j = 7; j >= 0; j--) v = (v << 8) | b[8 * i + j];
        t[i] = v;
    }
    while (F(geq_p)(t)) F(sub_p)(t);
    F(from_canon)(r, t);
}
static void F(to_le_bytes)(unsigned char b[32], const u64 a[4]) {
    u64 t[4];
    F(to_canon)(t, a);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * i + j] = (unsigned char)(t[i] >> (8 * j));
}

static void F(init)(void) {
    /* -p^-1 mod 2^64 by Newton iteration */
    u64 p0 = F(P)[0], x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
    F(INV) = (u64)0 - x;
    /* R = 2^256 mod p, R2 = 2^512 mod p by repeated modular doubling of 1 */
    u64 t[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; i++) {
        /* t < p < 2^255 so 2t fits in 256 bits */
        u64 c = 0;
        for (int k = 0; k < 4; k++) {
            u64 n = (t[k] << 1) | c;
            c = t[k] >> 63;
            t[k] = n;
        }
        if (F(geq_p)(t)) F(sub_p)(t);
        if (i == 255) F(copy)(F(R), t);
    }
    F(copy)(F(R2), t);
}

#undef F
#undef FP_CAT
#undef FP_CAT_
#undef FP_NAME
#undef FP_MOD0
#undef FP_MOD1
#undef FP_MOD2
#undef FP_MOD3
