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
static inline void F(add_c)(u64 r[4], const u64 a[4], const u64 b[4]) {
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
static inline void F(sub_c)(u64 r[4], const u64 a[4], const u64 b[4]) {
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
#if defined(__x86_64__) && !defined(ORC_NO_ASM) && FP_MOD2 == 0 && FP_MOD3 == 0x4000000000000000ULL
/* the same two operations as ADD/ADC and SUB/SBB chains with a conditional move / a masked add-back (definitions: above) */
static inline void F(add)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 s0, s1, s2, s3, d0, d1, d2, d3, l;
    __asm__("movq (%[ap]), %[s0]\n\t"  "movq 8(%[ap]), %[s1]\n\t"  "movq 16(%[ap]), %[s2]\n\t"  "movq 24(%[ap]), %[s3]\n\t"
            "addq (%[bp]), %[s0]\n\t"  "adcq 8(%[bp]), %[s1]\n\t"  "adcq 16(%[bp]), %[s2]\n\t"  "adcq 24(%[bp]), %[s3]\n\t"
            "movabsq $0x4000000000000000, %[l]\n\t"
            "movq %[s0], %[d0]\n\t"  "movq %[s1], %[d1]\n\t"  "movq %[s2], %[d2]\n\t"  "movq %[s3], %[d3]\n\t"
            "subq %[kp0], %[d0]\n\t"  "sbbq %[kp1], %[d1]\n\t"  "sbbq $0, %[d2]\n\t"  "sbbq %[l], %[d3]\n\t"
            "cmovcq %[s0], %[d0]\n\t"  "cmovcq %[s1], %[d1]\n\t"  "cmovcq %[s2], %[d2]\n\t"  "cmovcq %[s3], %[d3]\n\t"
            : [s0] "=&r"(s0), [s1] "=&r"(s1), [s2] "=&r"(s2), [s3] "=&r"(s3), [d0] "=&r"(d0), [d1] "=&r"(d1), [d2] "=&r"(d2),
              [d3] "=&r"(d3), [l] "=&r"(l)
            : [ap] "r"(a), [bp] "r"(b), "m"(*(const u64(*)[4])a), "m"(*(const u64(*)[4])b), [kp0] "m"(F(P)[0]), [kp1] "m"(F(P)[1])
            : "cc");
    r[0] = d0, r[1] = d1, r[2] = d2, r[3] = d3;
}
static inline void F(sub)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 d0, d1, d2, d3, m, q0, q1, q3;
    __asm__("movq (%[ap]), %[d0]\n\t"  "movq 8(%[ap]), %[d1]\n\t"  "movq 16(%[ap]), %[d2]\n\t"  "movq 24(%[ap]), %[d3]\n\t"
            "subq (%[bp]), %[d0]\n\t"  "sbbq 8(%[bp]), %[d1]\n\t"  "sbbq 16(%[bp]), %[d2]\n\t"  "sbbq 24(%[bp]), %[d3]\n\t"
            "sbbq %[m], %[m]\n\t"
            "movq %[kp0], %[q0]\n\t"  "movq %[kp1], %[q1]\n\t"  "movabsq $0x4000000000000000, %[q3]\n\t"
            "andq %[m], %[q0]\n\t"  "andq %[m], %[q1]\n\t"  "andq %[m], %[q3]\n\t"
            "addq %[q0], %[d0]\n\t"  "adcq %[q1], %[d1]\n\t"  "adcq $0, %[d2]\n\t"  "adcq %[q3], %[d3]\n\t"
            : [d0] "=&r"(d0), [d1] "=&r"(d1), [d2] "=&r"(d2), [d3] "=&r"(d3), [m] "=&r"(m), [q0] "=&r"(q0), [q1] "=&r"(q1),
              [q3] "=&r"(q3)
            : [ap] "r"(a), [bp] "r"(b), "m"(*(const u64(*)[4])a), "m"(*(const u64(*)[4])b), [kp0] "m"(F(P)[0]), [kp1] "m"(F(P)[1])
            : "cc");
    r[0] = d0, r[1] = d1, r[2] = d2, r[3] = d3;
}
#else
static inline void F(add)(u64 r[4], const u64 a[4], const u64 b[4]) { F(add_c)(r, a, b); }
static inline void F(sub)(u64 r[4], const u64 a[4], const u64 b[4]) { F(sub_c)(r, a, b); }
#endif
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
static inline void F(mul_c)(u64 r[4], const u64 a[4], const u64 b[4]) {
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
#if defined(__x86_64__) && defined(__BMI2__) && !defined(ORC_NO_ASM) && FP_MOD2 == 0 && FP_MOD3 == 0x4000000000000000ULL
/* The same CIOS recurrence with MULX and one ADC chain per row, so that the CPU arm of the benchmark runs at the speed of a
 * tuned field library (arkworks' Montgomery backend is of this kind) instead of the compiler's rendering of the loop above:
 * ~26 instead of ~46 ns per multiplication.  p = p0 + p1 2^64 + 2^254, so one reduction step adds m*(p0 + p1 2^64) (two
 * products) and m 2^254 (two shifts); a, b < p keeps the running value below 2p: five accumulators, the limb a reduction
 * step clears is the next row's top limb.  F(mul_c) above is the definition; tests/test_oracle_golden.py compares the two. */
#define ORC_ROW(bi, T0, T1, T2, T3, T4)                                                                              \
    "movq " bi "(%[bp]), %%rdx\n\t"                                                                                       \
    "mulx (%[ap]), %[l], %[h0]\n\t"  "addq %[l], %[" T0 "]\n\t"                                                        \
    "mulx 8(%[ap]), %[l], %[h1]\n\t"  "adcq %[l], %[" T1 "]\n\t"                                                        \
    "mulx 16(%[ap]), %[l], %[h2]\n\t"  "adcq %[l], %[" T2 "]\n\t"                                                        \
    "mulx 24(%[ap]), %[l], %[" T4 "]\n\t"  "adcq %[l], %[" T3 "]\n\t"  "adcq $0, %[" T4 "]\n\t"                          \
    "addq %[h0], %[" T1 "]\n\t"  "adcq %[h1], %[" T2 "]\n\t"  "adcq %[h2], %[" T3 "]\n\t"  "adcq $0, %[" T4 "]\n\t"
#define ORC_RED(T0, T1, T2, T3, T4)                                                                                  \
    "movq %[" T0 "], %%rdx\n\t"  "imulq %[kinv], %%rdx\n\t"                                                         \
    "mulx %[kp0], %[h2], %[h0]\n\t"  "mulx %[kp1], %[l], %[h1]\n\t"  "addq %[l], %[h0]\n\t"  "adcq $0, %[h1]\n\t"     \
    "movq %%rdx, %[l]\n\t"  "shlq $62, %[l]\n\t"  "shrq $2, %%rdx\n\t"                                               \
    "addq %[h2], %[" T0 "]\n\t"  "adcq %[h0], %[" T1 "]\n\t"  "adcq %[h1], %[" T2 "]\n\t"  "adcq %[l], %[" T3 "]\n\t"  \
    "adcq %%rdx, %[" T4 "]\n\t"
static inline void F(mul)(u64 r[4], const u64 a[4], const u64 b[4]) {
    u64 t0, t1, t2, t3, t4, l, h0, h1, h2;
    __asm__("xorl %k[t0], %k[t0]\n\t"  "xorl %k[t1], %k[t1]\n\t"  "xorl %k[t2], %k[t2]\n\t"  "xorl %k[t3], %k[t3]\n\t"
            ORC_ROW("0", "t0", "t1", "t2", "t3", "t4") ORC_RED("t0", "t1", "t2", "t3", "t4")
            ORC_ROW("8", "t1", "t2", "t3", "t4", "t0") ORC_RED("t1", "t2", "t3", "t4", "t0")
            ORC_ROW("16", "t2", "t3", "t4", "t0", "t1") ORC_RED("t2", "t3", "t4", "t0", "t1")
            ORC_ROW("24", "t3", "t4", "t0", "t1", "t2") ORC_RED("t3", "t4", "t0", "t1", "t2")
            /* (t4, t0, t1, t2) < 2p: subtract p unless that borrows */
            "movq %[t4], %[h0]\n\t"  "movq %[t0], %[h1]\n\t"  "movq %[t1], %[h2]\n\t"  "movq %[t2], %[l]\n\t"
            "movabsq $0x4000000000000000, %[t3]\n\t"
            "subq %[kp0], %[h0]\n\t"  "sbbq %[kp1], %[h1]\n\t"  "sbbq $0, %[h2]\n\t"  "sbbq %[t3], %[l]\n\t"
            "cmovcq %[t4], %[h0]\n\t"  "cmovcq %[t0], %[h1]\n\t"  "cmovcq %[t1], %[h2]\n\t"  "cmovcq %[t2], %[l]\n\t"
            : [t0] "=&r"(t0), [t1] "=&r"(t1), [t2] "=&r"(t2), [t3] "=&r"(t3), [t4] "=&r"(t4), [l] "=&r"(l), [h0] "=&r"(h0),
              [h1] "=&r"(h1), [h2] "=&r"(h2)
            : [ap] "r"(a), [bp] "r"(b), "m"(*(const u64(*)[4])a), "m"(*(const u64(*)[4])b), [kp0] "m"(F(P)[0]), [kp1] "m"(F(P)[1]),
              [kinv] "m"(F(INV))
            : "rdx", "cc");
    r[0] = h0, r[1] = h1, r[2] = h2, r[3] = l;
}
#undef ORC_ROW
#undef ORC_RED
#else
static inline void F(mul)(u64 r[4], const u64 a[4], const u64 b[4]) { F(mul_c)(r, a, b); }
#endif
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
