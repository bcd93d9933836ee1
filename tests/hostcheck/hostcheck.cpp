// tests/hostcheck/hostcheck.cpp -- compiles the product's __host__ __device__ field / curve headers as
// plain C++ (g++) so their logic can be checked against the oracle on a machine without a GPU.
// Test-only; never shipped, never linked into libhalo_b200.so.
#include <cstring>
#include "../../halo-accumulation_b200/csrc/ec.cuh"
#include "../../halo-accumulation_b200/csrc/glv.cuh"
using namespace halo;
// count of operand pairs on which the implementations of the host field core disagree: fp_mul / fp_add / fp_sub as
// dispatched (x86-64: MULX / ADC chains in inline assembly), the 64-bit C versions, the portable 32-bit-limb multiplication;
// also with the result aliasing either operand and with a == b
template <class P>
static uint64_t impl_cross(const uint32_t* a, const uint32_t* b, uint64_t n) {
    uint64_t bad = 0;
#if !defined(HALO_FP_FORCE_PORTABLE)
    for (uint64_t i = 0; i < n; i++) {
        typedef fp_t<P> F;
        F x, y, m0, m1, m2, s0, s1, d0, d1, t;
        memcpy(&x, a + 8 * i, 32); memcpy(&y, b + 8 * i, 32);
        fp_mul(m0, x, y); fp_mul_host64<P>(m1.v, x.v, y.v); fp_mul_portable(m2, x, y);
        fp_add(s0, x, y); host64::add<P>(s1.v, x.v, y.v);
        fp_sub(d0, x, y); host64::sub<P>(d1.v, x.v, y.v);
        bad += !fp_eq(m0, m1) || !fp_eq(m0, m2) || !fp_eq(s0, s1) || !fp_eq(d0, d1);
        t = x; fp_mul(t, t, y); bad += !fp_eq(t, m0);
        t = y; fp_mul(t, x, t); bad += !fp_eq(t, m0);
        t = x; fp_add(t, t, y); bad += !fp_eq(t, s0);
        t = y; fp_sub(t, x, t); bad += !fp_eq(t, d0);
        F q0, q1; fp_mul(q0, x, x); fp_mul_portable(q1, x, x); t = x; fp_sqr(t, t); bad += !fp_eq(q0, q1) || !fp_eq(t, q0);
    }
#else
    (void)a; (void)b; (void)n;
#endif
    return bad;
}
extern "C" {
void hc_fp_mul(int which, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    if (which) { fr_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); fp_mul(z, x, y); memcpy(r, &z, 32); }
    else { fq_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); fp_mul(z, x, y); memcpy(r, &z, 32); }
}
void hc_fp_addsub(int which, int sub, const uint32_t* a, const uint32_t* b, uint32_t* r) {
    if (which) { fr_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); if (sub) fp_sub(z, x, y); else fp_add(z, x, y); memcpy(r, &z, 32); }
    else { fq_t x, y, z; memcpy(&x, a, 32); memcpy(&y, b, 32); if (sub) fp_sub(z, x, y); else fp_add(z, x, y); memcpy(r, &z, 32); }
}
void hc_fp_inv(int which, const uint32_t* a, uint32_t* r) {
    if (which) { fr_t x, z; memcpy(&x, a, 32); fp_inv(z, x); memcpy(r, &z, 32); }
    else { fq_t x, z; memcpy(&x, a, 32); fp_inv(z, x); memcpy(r, &z, 32); }
}
// count of inputs (n Montgomery residues) on which the division-step inversion (fp_inv) and the Fermat power (fp_inv_fermat)
// disagree, or a * fp_inv(a) != 1
uint64_t hc_fp_inv_cross(int which, const uint32_t* a, uint64_t n) {
    uint64_t bad = 0;
    for (uint64_t i = 0; i < n; i++) {
        if (which) { fr_t x, y, z, o, c; memcpy(&x, a + 8 * i, 32); fp_inv(y, x); fp_inv_fermat(z, x); fp_one(o); fp_mul(c, x, y);
                     bad += !fp_eq(y, z) || (!fp_is_zero(x) && !fp_eq(c, o)) || (fp_is_zero(x) && !fp_is_zero(y)); }
        else { fq_t x, y, z, o, c; memcpy(&x, a + 8 * i, 32); fp_inv(y, x); fp_inv_fermat(z, x); fp_one(o); fp_mul(c, x, y);
               bad += !fp_eq(y, z) || (!fp_is_zero(x) && !fp_eq(c, o)) || (fp_is_zero(x) && !fp_is_zero(y)); }
    }
    return bad;
}
uint64_t hc_fp_impl_cross(int which, const uint32_t* a, const uint32_t* b, uint64_t n) {
    return which ? impl_cross<ScalarParams>(a, b, n) : impl_cross<BaseParams>(a, b, n);
}
int hc_fp_uses_x64_asm(void) {
#if defined(HALO_FP_X64_ASM)
    return host64::g_has_bmi2 ? 2 : 1;
#else
    return 0;
#endif
}
void hc_fp_canon(int which, int to, const uint32_t* a, uint32_t* r) {
    if (which) { fr_t x; if (to) { memcpy(&x, a, 32); fp_to_canon(r, x); } else { fp_from_canon(x, a); memcpy(r, &x, 32); } }
    else { fq_t x; if (to) { memcpy(&x, a, 32); fp_to_canon(r, x); } else { fp_from_canon(x, a); memcpy(r, &x, 32); } }
}
// sum_i (neg_i ? -P_i : P_i) with madd, returned as Jacobian
void hc_madd_chain(const uint32_t* aff, const uint8_t* neg, uint64_t n, uint32_t* out_jac) {
    xyzz_t acc; xyzz_set_inf(acc);
    for (uint64_t i = 0; i < n; i++) { affine_t p; memcpy(&p, aff + 16 * i, 64); xyzz_madd(acc, p, neg && neg[i]); }
    jac_t j; xyzz_to_jac(j, acc); memcpy(out_jac, &j, 96);
}
// sum of Jacobian inputs via xyzz_add (+ optional doublings of the total)
void hc_add_chain(const uint32_t* jac, uint64_t n, int dbls, uint32_t* out_jac) {
    xyzz_t acc; xyzz_set_inf(acc);
    for (uint64_t i = 0; i < n; i++) { jac_t p; memcpy(&p, jac + 24 * i, 96); xyzz_t q; jac_to_xyzz(q, p); xyzz_add(acc, q); }
    for (int i = 0; i < dbls; i++) xyzz_dbl(acc, acc);
    jac_t j; xyzz_to_jac(j, acc); memcpy(out_jac, &j, 96);
}
void hc_to_affine(const uint32_t* jac, uint32_t* aff) {
    jac_t p; memcpy(&p, jac, 96); xyzz_t q; jac_to_xyzz(q, p); affine_t a; xyzz_to_affine(a, q); memcpy(aff, &a, 64);
}
// `Projective * Fr` of the host layer: GLV split + joint sparse form (csrc/glv.cuh); k is a Montgomery scalar
void hc_mul_glv(const uint32_t* jac, const uint32_t* k_mont, uint32_t* out_jac) {
    jac_t p; memcpy(&p, jac, 96); xyzz_t q, r; jac_to_xyzz(q, p); fr_t k; memcpy(&k, k_mont, 32); xyzz_mul_glv(r, q, k);
    jac_t j; xyzz_to_jac(j, r); memcpy(out_jac, &j, 96);
}
void hc_mul(const uint32_t* jac, const uint32_t* k_canon, uint32_t* out_jac) {
    jac_t p; memcpy(&p, jac, 96); xyzz_t q, r; jac_to_xyzz(q, p); xyzz_mul_canon(r, q, k_canon);
    jac_t j; xyzz_to_jac(j, r); memcpy(out_jac, &j, 96);
}
}
