// fp.cuh -- 255-bit Montgomery field core for the two Pasta moduli (K1 in SURVEY.md section 2b).
//
// Replaces, for the hot path, arkworks' `Fp<MontBackend<_, 4>>` arithmetic beneath every call site
// listed in SURVEY.md section 8a (group.rs:13-37, pcdl.rs:195-227, pcdl.rs:56-77).  Elements use the
// reference's in-memory layout bit for bit: 256-bit little-endian Montgomery residues with
// R = 2^256 (consts.rs:4-21, main.rs:47-53), here viewed as 8 x u32 limbs so every product is a
// 32-bit IMAD / IMAD.WIDE on the sm_100a integer pipe.
//
// Both moduli have the shape  p = 2^254 + t  with  t < 2^126  and  p = 1 (mod 2^32):
//   limbs(p) = [1, P1, P2, P3, 0, 0, 0, 0x40000000]   and   -p^-1 mod 2^32 = 0xffffffff,
// so one Montgomery reduction step is  m = -T[i]  (no multiply), three real 32x32 products
// (m*P1, m*P2, m*P3) and a shift (m * 2^30).  See DESIGN.md section "K1".
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define HALO_HD __host__ __device__ __forceinline__
#include "fp_mul_asm.cuh"
#else
#define HALO_HD inline
#endif

namespace halo {

struct FqParams {  // Pallas base field (point coordinates)
    static constexpr uint32_t P1 = 0x992d30edu, P2 = 0x094cf91bu, P3 = 0x224698fcu;
    static constexpr uint64_t INV64 = 0x992d30ecffffffffull;  // -p^-1 mod 2^64 (host path)
    HALO_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xfffffffdu, 0x34786d38u, 0xe41914adu, 0x992c350bu,
                                   0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
        return v[i];
    }
    HALO_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0x0000000fu, 0x8c78ecb3u, 0x8b0de0e7u, 0xd7d30dbdu,
                                   0xc3c95d18u, 0x7797a99bu, 0x7b9cb714u, 0x096d41afu};
        return v[i];
    }
};
struct FrParams {  // Pallas scalar field (= Vesta base field)
    static constexpr uint32_t P1 = 0x8c46eb21u, P2 = 0x0994a8ddu, P3 = 0x224698fcu;
    static constexpr uint64_t INV64 = 0x8c46eb20ffffffffull;
    HALO_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t v[8] = {0xfffffffdu, 0x5b2b3e9cu, 0xe3420567u, 0x992c350bu,
                                   0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
        return v[i];
    }
    HALO_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t v[8] = {0x0000000fu, 0xfc9678ffu, 0x891a16e3u, 0x67bb433du,
                                   0x04ccf590u, 0x7fae2310u, 0x7ccfdaa9u, 0x096d41afu};
        return v[i];
    }
};

template <class P>
HALO_HD constexpr uint32_t fp_mod(int i) {
    return i == 0 ? 1u : i == 1 ? P::P1 : i == 2 ? P::P2 : i == 3 ? P::P3 : i == 7 ? 0x40000000u : 0u;
}

template <class P>
struct alignas(16) fp_t {
    uint32_t v[8];
};
// The curve of this build.  Pallas (default): coordinates in Fq, scalars in Fr.  -DHALO_CURVE_VESTA swaps the roles:
// Vesta is y^2 = x^3 + 5 over Pallas' scalar field with Pallas' base field as its scalar field (the Pasta cycle), so
// every kernel, the host layer and the C ABI are the same sources with `fq_t` = coordinate field, `fr_t` = scalar field.
#if defined(HALO_CURVE_VESTA)
using BaseParams = FrParams;
using ScalarParams = FqParams;
#define HALO_CURVE_NAME "vesta"
#else
using BaseParams = FqParams;
using ScalarParams = FrParams;
#define HALO_CURVE_NAME "pallas"
#endif
using fq_t = fp_t<BaseParams>;    // coordinate field of the curve
using fr_t = fp_t<ScalarParams>;  // scalar field of the curve

template <class P>
HALO_HD void fp_zero(fp_t<P>& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
template <class P>
HALO_HD void fp_one(fp_t<P>& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = P::one(i);
}
template <class P>
HALO_HD bool fp_is_zero(const fp_t<P>& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
template <class P>
HALO_HD bool fp_eq(const fp_t<P>& a, const fp_t<P>& b) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}

// ---- raw 256-bit helpers ---------------------------------------------------------------------
// r = a + b, returns carry out
HALO_HD uint32_t add8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c;
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    return (uint32_t)c;
#endif
}
// r = a - b, returns borrow out (1 if a < b)
HALO_HD uint32_t sub8(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
#if defined(__CUDA_ARCH__)
    uint32_t c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    return c & 1u;  // subc.u32 0,0 with borrow gives 0xffffffff
#else
    uint64_t br = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a[i] - b[i] - br;
        r[i] = (uint32_t)d;
        br = (d >> 32) & 1;
    }
    return (uint32_t)br;
#endif
}

template <class P>
HALO_HD void fp_mod_limbs(uint32_t m[8]) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = fp_mod<P>(i);
}

// if (x >= p) x -= p      (x < 2p)
template <class P>
HALO_HD void fp_reduce_once(uint32_t x[8]) {
    uint32_t m[8], t[8];
    fp_mod_limbs<P>(m);
    uint32_t borrow = sub8(t, x, m);
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = borrow ? x[i] : t[i];
}

#if !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
// Host path (little-endian x86-64 / aarch64): the same bytes seen as 4 x u64.  The host glue of the hot path is a few
// hundred point operations per call (transcript points, C', the Horner finish of every variable-base MSM: ~255
// doublings); with 32-bit-limb loops a scalar multiplication on the host cost 0.3 ms.
namespace host64 {
typedef unsigned __int128 u128;
template <class P>
struct Mod {
    static constexpr uint64_t p0 = 1ull | ((uint64_t)P::P1 << 32), p1 = (uint64_t)P::P2 | ((uint64_t)P::P3 << 32), p2 = 0,
                              p3 = 0x4000000000000000ull;
};
inline void load(uint64_t o[4], const uint32_t v[8]) { __builtin_memcpy(o, v, 32); }
inline void store(uint32_t v[8], const uint64_t o[4]) { __builtin_memcpy(v, o, 32); }
inline uint64_t add4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    u128 c = (u128)a[0] + b[0];
    r[0] = (uint64_t)c;
    c = (c >> 64) + a[1] + b[1];
    r[1] = (uint64_t)c;
    c = (c >> 64) + a[2] + b[2];
    r[2] = (uint64_t)c;
    c = (c >> 64) + a[3] + b[3];
    r[3] = (uint64_t)c;
    return (uint64_t)(c >> 64);
}
inline uint64_t sub4(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {  // returns the borrow (0 / 1)
    u128 d = (u128)a[0] - b[0];
    r[0] = (uint64_t)d;
    d = (u128)a[1] - b[1] - (uint64_t)((d >> 64) & 1);
    r[1] = (uint64_t)d;
    d = (u128)a[2] - b[2] - (uint64_t)((d >> 64) & 1);
    r[2] = (uint64_t)d;
    d = (u128)a[3] - b[3] - (uint64_t)((d >> 64) & 1);
    r[3] = (uint64_t)d;
    return (uint64_t)((d >> 64) & 1);
}
template <class P>
inline void add(uint32_t r[8], const uint32_t a32[8], const uint32_t b32[8]) {
    uint64_t a[4], b[4], s[4], d[4];
    const uint64_t p[4] = {Mod<P>::p0, Mod<P>::p1, Mod<P>::p2, Mod<P>::p3};
    load(a, a32);
    load(b, b32);
    add4(s, a, b);  // a, b < p < 2^255: no carry out
    const uint64_t borrow = sub4(d, s, p);
    for (int i = 0; i < 4; i++) s[i] = borrow ? s[i] : d[i];
    store(r, s);
}
template <class P>
inline void sub(uint32_t r[8], const uint32_t a32[8], const uint32_t b32[8]) {
    uint64_t a[4], b[4], d[4], t[4];
    const uint64_t p[4] = {Mod<P>::p0, Mod<P>::p1, Mod<P>::p2, Mod<P>::p3};
    load(a, a32);
    load(b, b32);
    const uint64_t borrow = sub4(d, a, b);
    add4(t, d, p);
    for (int i = 0; i < 4; i++) d[i] = borrow ? t[i] : d[i];
    store(r, d);
}
#if defined(__x86_64__) && defined(__GNUC__) && !defined(HALO_FP_NO_X64_ASM)
// x86-64: the same two operations as straight ADD/ADC, SUB/SBB chains with a conditional move (the compiler turns the
// 128-bit carry idiom above into ~40 instructions each; a point doubling on the host is 9 multiplications and 12 of these).
typedef uint64_t __attribute__((may_alias, aligned(4))) u64m;
template <class P>
inline void add_x64(uint32_t r32[8], const uint32_t a32[8], const uint32_t b32[8]) {
    static const uint64_t K[2] = {Mod<P>::p0, Mod<P>::p1};
    const u64m* a = reinterpret_cast<const u64m*>(a32);
    const u64m* b = reinterpret_cast<const u64m*>(b32);
    uint64_t s0, s1, s2, s3, d0, d1, d2, d3, l;
    __asm__("movq (%[ap]), %[s0]\n\t"
            "movq 8(%[ap]), %[s1]\n\t"
            "movq 16(%[ap]), %[s2]\n\t"
            "movq 24(%[ap]), %[s3]\n\t"
            "addq (%[bp]), %[s0]\n\t"
            "adcq 8(%[bp]), %[s1]\n\t"
            "adcq 16(%[bp]), %[s2]\n\t"
            "adcq 24(%[bp]), %[s3]\n\t"  // a, b < p < 2^255: no carry out
            "movabsq $0x4000000000000000, %[l]\n\t"
            "movq %[s0], %[d0]\n\t"
            "movq %[s1], %[d1]\n\t"
            "movq %[s2], %[d2]\n\t"
            "movq %[s3], %[d3]\n\t"
            "subq %[kp0], %[d0]\n\t"
            "sbbq %[kp1], %[d1]\n\t"
            "sbbq $0, %[d2]\n\t"
            "sbbq %[l], %[d3]\n\t"
            "cmovcq %[s0], %[d0]\n\t"
            "cmovcq %[s1], %[d1]\n\t"
            "cmovcq %[s2], %[d2]\n\t"
            "cmovcq %[s3], %[d3]\n\t"
            : [s0] "=&r"(s0), [s1] "=&r"(s1), [s2] "=&r"(s2), [s3] "=&r"(s3), [d0] "=&r"(d0), [d1] "=&r"(d1), [d2] "=&r"(d2),
              [d3] "=&r"(d3), [l] "=&r"(l)
            : [ap] "r"(a), [bp] "r"(b), "m"(*reinterpret_cast<const u64m(*)[4]>(a)), "m"(*reinterpret_cast<const u64m(*)[4]>(b)),
              [kp0] "m"(K[0]), [kp1] "m"(K[1])
            : "cc");
    u64m* r = reinterpret_cast<u64m*>(r32);
    r[0] = d0, r[1] = d1, r[2] = d2, r[3] = d3;
}
template <class P>
inline void sub_x64(uint32_t r32[8], const uint32_t a32[8], const uint32_t b32[8]) {
    static const uint64_t K[2] = {Mod<P>::p0, Mod<P>::p1};
    const u64m* a = reinterpret_cast<const u64m*>(a32);
    const u64m* b = reinterpret_cast<const u64m*>(b32);
    uint64_t d0, d1, d2, d3, m, q0, q1, q3;
    __asm__("movq (%[ap]), %[d0]\n\t"
            "movq 8(%[ap]), %[d1]\n\t"
            "movq 16(%[ap]), %[d2]\n\t"
            "movq 24(%[ap]), %[d3]\n\t"
            "subq (%[bp]), %[d0]\n\t"
            "sbbq 8(%[bp]), %[d1]\n\t"
            "sbbq 16(%[bp]), %[d2]\n\t"
            "sbbq 24(%[bp]), %[d3]\n\t"
            "sbbq %[m], %[m]\n\t"  // all ones if a < b: add p back
            "movq %[kp0], %[q0]\n\t"
            "movq %[kp1], %[q1]\n\t"
            "movabsq $0x4000000000000000, %[q3]\n\t"
            "andq %[m], %[q0]\n\t"
            "andq %[m], %[q1]\n\t"
            "andq %[m], %[q3]\n\t"
            "addq %[q0], %[d0]\n\t"
            "adcq %[q1], %[d1]\n\t"
            "adcq $0, %[d2]\n\t"
            "adcq %[q3], %[d3]\n\t"
            : [d0] "=&r"(d0), [d1] "=&r"(d1), [d2] "=&r"(d2), [d3] "=&r"(d3), [m] "=&r"(m), [q0] "=&r"(q0), [q1] "=&r"(q1),
              [q3] "=&r"(q3)
            : [ap] "r"(a), [bp] "r"(b), "m"(*reinterpret_cast<const u64m(*)[4]>(a)), "m"(*reinterpret_cast<const u64m(*)[4]>(b)),
              [kp0] "m"(K[0]), [kp1] "m"(K[1])
            : "cc");
    u64m* r = reinterpret_cast<u64m*>(r32);
    r[0] = d0, r[1] = d1, r[2] = d2, r[3] = d3;
}
#define HALO_FP_X64_ADDSUB 1
#endif
}  // namespace host64
#endif

template <class P>
HALO_HD void fp_add(fp_t<P>& r, const fp_t<P>& a, const fp_t<P>& b) {
#if !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE) && defined(HALO_FP_X64_ADDSUB)
    host64::add_x64<P>(r.v, a.v, b.v);
    return;
#elif !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
    host64::add<P>(r.v, a.v, b.v);
    return;
#endif
    uint32_t s[8];
    add8(s, a.v, b.v);  // a, b < p < 2^255: no carry out
    fp_reduce_once<P>(s);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = s[i];
}
template <class P>
HALO_HD void fp_sub(fp_t<P>& r, const fp_t<P>& a, const fp_t<P>& b) {
#if !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE) && defined(HALO_FP_X64_ADDSUB)
    host64::sub_x64<P>(r.v, a.v, b.v);
    return;
#elif !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
    host64::sub<P>(r.v, a.v, b.v);
    return;
#endif
    uint32_t d[8], m[8], t[8];
    uint32_t borrow = sub8(d, a.v, b.v);
    fp_mod_limbs<P>(m);
    add8(t, d, m);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = borrow ? t[i] : d[i];
}
template <class P>
HALO_HD void fp_dbl(fp_t<P>& r, const fp_t<P>& a) {
    fp_add(r, a, a);
}
template <class P>
HALO_HD void fp_neg(fp_t<P>& r, const fp_t<P>& a) {
    uint32_t m[8], t[8];
    fp_mod_limbs<P>(m);
    sub8(t, m, a.v);
    bool z = fp_is_zero(a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = z ? 0u : t[i];
}
// r = neg ? -a : a   (branch-free select used by the signed-digit bucket accumulation)
template <class P>
HALO_HD void fp_cneg(fp_t<P>& r, const fp_t<P>& a, bool neg) {
    fp_t<P> n;
    fp_neg(n, a);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = neg ? n.v[i] : a.v[i];
}

// ---- Montgomery multiplication -----------------------------------------------------------------
// Montgomery reduction of a 16-limb value T < p * 2^256: returns T * 2^-256 mod p in r (fully reduced).
// Per step i: m = -T[i] (because -p^-1 = -1 mod 2^32); T += m * p * 2^(32 i).  With
// p = 1 + 2^32 * (P1 + 2^32 P2 + 2^64 P3) + 2^254 the limb T[i] cancels to zero with carry
// (T[i] != 0), the three products land on limbs i+1..i+4 and m * 2^30 on limbs i+7, i+8.
template <class P>
HALO_HD void fp_mont_reduce(uint32_t r[8], uint32_t T[16]) {
    uint32_t top = 0;  // carry out of limb 15 (T + sum m_i p 2^(32i) < 2p * 2^256 fits 16 limbs + 1 bit)
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = 0u - T[i];
        uint64_t c = (T[i] != 0) ? 1u : 0u;
        c += (uint64_t)m * P::P1 + T[i + 1];
        T[i + 1] = (uint32_t)c;
        c >>= 32;
        c += (uint64_t)m * P::P2 + T[i + 2];
        T[i + 2] = (uint32_t)c;
        c >>= 32;
        c += (uint64_t)m * P::P3 + T[i + 3];
        T[i + 3] = (uint32_t)c;
        c >>= 32;
#pragma unroll
        for (int j = 4; j < 7; j++) {
            c += T[i + j];
            T[i + j] = (uint32_t)c;
            c >>= 32;
        }
        c += ((uint64_t)m << 30) + T[i + 7];
        T[i + 7] = (uint32_t)c;
        c >>= 32;
#pragma unroll
        for (int j = i + 8; j < 16; j++) {
            c += T[j];
            T[j] = (uint32_t)c;
            c >>= 32;
        }
        top += (uint32_t)c;
    }
    // result = T[8..16) (+ top * 2^256) < 2p
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = T[8 + i];
    if (top) {
        uint32_t m[8], t[8];
        fp_mod_limbs<P>(m);
        sub8(t, x, m);
#pragma unroll
        for (int i = 0; i < 8; i++) x[i] = t[i];
    } else {
        fp_reduce_once<P>(x);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = x[i];
}

// 8x8 schoolbook product into 16 limbs
HALO_HD void mul8x8(uint32_t T[16], const uint32_t a[8], const uint32_t b[8]) {
#pragma unroll
    for (int i = 0; i < 16; i++) T[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a[j] * b[i] + T[i + j];
            T[i + j] = (uint32_t)c;
            c >>= 32;
        }
        T[i + 8] = (uint32_t)c;
    }
}

#if !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
// Host path (transcript glue, the O(255)-doubling Horner finish of an MSM): 4 x u64 CIOS on the same bytes, unrolled, with the
// shape of the modulus folded in (limb 2 is zero, limb 3 is 2^62: that product is a shift).
template <class P>
inline void fp_mul_host64(uint32_t r32[8], const uint32_t a32[8], const uint32_t b32[8]) {
    using host64::u128;
    typedef host64::Mod<P> M;
    uint64_t a[4], b[4];
    host64::load(a, a32);
    host64::load(b, b32);
    uint64_t t0 = 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0;
#define HALO_H64_ROUND(bi)                                  \
    {                                                       \
        u128 c = (u128)a[0] * (bi) + t0;                    \
        t0 = (uint64_t)c;                                   \
        c = (c >> 64) + (u128)a[1] * (bi) + t1;             \
        t1 = (uint64_t)c;                                   \
        c = (c >> 64) + (u128)a[2] * (bi) + t2;             \
        t2 = (uint64_t)c;                                   \
        c = (c >> 64) + (u128)a[3] * (bi) + t3;             \
        t3 = (uint64_t)c;                                   \
        c = (c >> 64) + t4;                                 \
        t4 = (uint64_t)c;                                   \
        const uint64_t t5 = (uint64_t)(c >> 64);            \
        const uint64_t m = t0 * P::INV64;                   \
        c = (u128)m * M::p0 + t0;                           \
        c = (c >> 64) + (u128)m * M::p1 + t1;               \
        t0 = (uint64_t)c;                                   \
        c = (c >> 64) + t2;                                 \
        t1 = (uint64_t)c;                                   \
        c = (c >> 64) + ((u128)m << 62) + t3;               \
        t2 = (uint64_t)c;                                   \
        c = (c >> 64) + t4;                                 \
        t3 = (uint64_t)c;                                   \
        t4 = t5 + (uint64_t)(c >> 64);                      \
    }
    HALO_H64_ROUND(b[0])
    HALO_H64_ROUND(b[1])
    HALO_H64_ROUND(b[2])
    HALO_H64_ROUND(b[3])
#undef HALO_H64_ROUND
    const uint64_t t[4] = {t0, t1, t2, t3}, p[4] = {M::p0, M::p1, M::p2, M::p3};
    uint64_t d[4];
    const uint64_t borrow = host64::sub4(d, t, p);
    const bool ge = t4 != 0 || borrow == 0;
    for (int i = 0; i < 4; i++) d[i] = ge ? d[i] : t[i];
    host64::store(r32, d);
}
#endif

#if !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE) && defined(__x86_64__) && defined(__GNUC__) && !defined(HALO_FP_NO_X64_ASM)
#define HALO_FP_X64_ASM 1
// x86-64 with BMI2 (checked once at run time; every server CPU since 2013): the same CIOS rounds written with MULX, which leaves
// the flags alone, so each round is one row of four products folded by a single ADC chain and one reduction step
// (m * p = m * (p0 + p1 2^64) + m 2^254: two products and two shifts).  a, b < p < 2^255 keeps the running value below 2p, so
// five accumulators suffice and the limb that becomes zero in a reduction step is the next round's top limb (the register
// names rotate, nothing moves).  ~150 instructions against ~390 from the compiler for fp_mul_host64: the Horner finish of a
// variable-base MSM and the verifier's scalar multiplications are chains of these.
namespace host64 {
typedef uint64_t __attribute__((may_alias, aligned(4))) u64a;
inline bool cpu_has_bmi2() {
    unsigned a = 0, b = 0, c = 0, d = 0;
    __asm__ volatile("cpuid" : "=a"(a), "=b"(b), "=c"(c), "=d"(d) : "a"(0), "c"(0));
    if (a < 7) return false;
    __asm__ volatile("cpuid" : "=a"(a), "=b"(b), "=c"(c), "=d"(d) : "a"(7), "c"(0));
    return (b >> 8) & 1;
}
inline const bool g_has_bmi2 = cpu_has_bmi2();  // (read as false before its initialiser has run: the portable path is taken)
}  // namespace host64
#define HALO_X64_ROW(bi, T0, T1, T2, T3, T4) /* T4 is zero on entry: the high half of the last product lands in it directly */ \
    "movq " bi "(%[bp]), %%rdx\n\t"               \
    "mulx (%[ap]), %[l], %[h0]\n\t"             \
    "addq %[l], %[" T0 "]\n\t"                \
    "mulx 8(%[ap]), %[l], %[h1]\n\t"             \
    "adcq %[l], %[" T1 "]\n\t"                \
    "mulx 16(%[ap]), %[l], %[h2]\n\t"             \
    "adcq %[l], %[" T2 "]\n\t"                \
    "mulx 24(%[ap]), %[l], %[" T4 "]\n\t"         \
    "adcq %[l], %[" T3 "]\n\t"                \
    "adcq $0, %[" T4 "]\n\t"                  \
    "addq %[h0], %[" T1 "]\n\t"               \
    "adcq %[h1], %[" T2 "]\n\t"               \
    "adcq %[h2], %[" T3 "]\n\t"               \
    "adcq $0, %[" T4 "]\n\t"
#define HALO_X64_RED(T0, T1, T2, T3, T4) /* leaves T0 = 0: the next round's top limb */ \
    "movq %[" T0 "], %%rdx\n\t"           \
    "imulq %[kinv], %%rdx\n\t"            \
    "mulx %[kp0], %[h2], %[h0]\n\t"       \
    "mulx %[kp1], %[l], %[h1]\n\t"        \
    "addq %[l], %[h0]\n\t"                \
    "adcq $0, %[h1]\n\t"                  \
    "movq %%rdx, %[l]\n\t"                \
    "shlq $62, %[l]\n\t"                  \
    "shrq $2, %%rdx\n\t"                  \
    "addq %[h2], %[" T0 "]\n\t"           \
    "adcq %[h0], %[" T1 "]\n\t"           \
    "adcq %[h1], %[" T2 "]\n\t"           \
    "adcq %[l], %[" T3 "]\n\t"            \
    "adcq %%rdx, %[" T4 "]\n\t"
template <class P>
inline void fp_mul_x64(uint32_t r32[8], const uint32_t a32[8], const uint32_t b32[8]) {
    typedef host64::Mod<P> M;
    static_assert(M::p2 == 0 && M::p3 == 0x4000000000000000ull, "modulus shape: 2^254 + (126 bits)");
    static const uint64_t K[3] = {M::p0, M::p1, P::INV64};
    const host64::u64a* a = reinterpret_cast<const host64::u64a*>(a32);
    const host64::u64a* b = reinterpret_cast<const host64::u64a*>(b32);
    uint64_t t0, t1, t2, t3, t4, l, h0, h1, h2;  // nine registers + rdx, so the statement fits wherever it is inlined
    __asm__("xorl %k[t0], %k[t0]\n\t"
            "xorl %k[t1], %k[t1]\n\t"
            "xorl %k[t2], %k[t2]\n\t"
            "xorl %k[t3], %k[t3]\n\t"  //
            HALO_X64_ROW("0", "t0", "t1", "t2", "t3", "t4") HALO_X64_RED("t0", "t1", "t2", "t3", "t4")  //
            HALO_X64_ROW("8", "t1", "t2", "t3", "t4", "t0") HALO_X64_RED("t1", "t2", "t3", "t4", "t0")  //
            HALO_X64_ROW("16", "t2", "t3", "t4", "t0", "t1") HALO_X64_RED("t2", "t3", "t4", "t0", "t1")  //
            HALO_X64_ROW("24", "t3", "t4", "t0", "t1", "t2") HALO_X64_RED("t3", "t4", "t0", "t1", "t2")
            // the value is (t4, t0, t1, t2) < 2p: subtract p once if that does not borrow
            "movq %[t4], %[h0]\n\t"
            "movq %[t0], %[h1]\n\t"
            "movq %[t1], %[h2]\n\t"
            "movq %[t2], %[l]\n\t"
            "movabsq $0x4000000000000000, %[t3]\n\t"
            "subq %[kp0], %[h0]\n\t"
            "sbbq %[kp1], %[h1]\n\t"
            "sbbq $0, %[h2]\n\t"
            "sbbq %[t3], %[l]\n\t"
            "cmovcq %[t4], %[h0]\n\t"
            "cmovcq %[t0], %[h1]\n\t"
            "cmovcq %[t1], %[h2]\n\t"
            "cmovcq %[t2], %[l]\n\t"
            : [t0] "=&r"(t0), [t1] "=&r"(t1), [t2] "=&r"(t2), [t3] "=&r"(t3), [t4] "=&r"(t4), [l] "=&r"(l), [h0] "=&r"(h0),
              [h1] "=&r"(h1), [h2] "=&r"(h2)
            : [ap] "r"(a), [bp] "r"(b), "m"(*reinterpret_cast<const host64::u64a(*)[4]>(a)),
              "m"(*reinterpret_cast<const host64::u64a(*)[4]>(b)), [kp0] "m"(K[0]), [kp1] "m"(K[1]), [kinv] "m"(K[2])
            : "rdx", "cc");
    host64::u64a* r = reinterpret_cast<host64::u64a*>(r32);
    r[0] = h0, r[1] = h1, r[2] = h2, r[3] = l;
}
#undef HALO_X64_ROW
#undef HALO_X64_RED
#endif

// Portable 32-bit-limb version (reference semantics of the device algorithm; also what the host check compiles).
template <class P>
HALO_HD void fp_mul_portable(fp_t<P>& r, const fp_t<P>& a, const fp_t<P>& b) {
    uint32_t T[16];
    mul8x8(T, a.v, b.v);
    fp_mont_reduce<P>(r.v, T);
}

#ifndef HALO_FP_MUL_VARIANT
#define HALO_FP_MUL_VARIANT 0  // 0 and 2 measure ~1 % faster than 1 across the MSM and IPA kernels (3, 4 slower)
#endif

#if defined(__CUDACC__)
// Out-of-line copy of the multiplication (arguments and result by value: they travel in registers, no local memory).
// Kernels whose body would otherwise inline 10-20 copies of the 210-instruction multiplication overflow the
// instruction cache (ncu: sm__icc_request_hit_rate 84 %, "no instruction" the third largest stall of k_accumulate).
template <class P>
__device__ __noinline__ fp_t<P> fp_mul_call(fp_t<P> a, fp_t<P> b) {
    fp_t<P> r;
    fp_mul_asm<P, HALO_FP_MUL_VARIANT>(r.v, a.v, b.v);
    return r;
}
#endif

template <class P>
HALO_HD void fp_mul(fp_t<P>& r, const fp_t<P>& a, const fp_t<P>& b) {
#if defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE) && defined(HALO_FP_MUL_CALL)
    r = fp_mul_call<P>(a, b);
#elif defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
    uint32_t o[8];
    fp_mul_asm<P, HALO_FP_MUL_VARIANT>(o, a.v, b.v);  // generated straight-line PTX, csrc/fp_mul_asm.cuh
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
#elif !defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
#if defined(HALO_FP_X64_ASM)
    if (host64::g_has_bmi2) {
        fp_mul_x64<P>(r.v, a.v, b.v);
        return;
    }
#endif
    fp_mul_host64<P>(r.v, a.v, b.v);
#else
    fp_mul_portable(r, a, b);
#endif
}
template <class P>
HALO_HD void fp_sqr(fp_t<P>& r, const fp_t<P>& a) {
#if defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE) && defined(HALO_FP_MUL_CALL)
    r = fp_mul_call<P>(a, a);
#elif defined(__CUDA_ARCH__) && !defined(HALO_FP_FORCE_PORTABLE)
    uint32_t o[8];
    fp_sqr_asm<P, HALO_FP_MUL_VARIANT>(o, a.v);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = o[i];
#else
    fp_mul(r, a, a);
#endif
}

// Montgomery form <-> canonical integer
template <class P>
HALO_HD void fp_to_canon(uint32_t out[8], const fp_t<P>& a) {
    uint32_t T[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        T[i] = a.v[i];
        T[8 + i] = 0;
    }
    fp_mont_reduce<P>(out, T);
}
template <class P>
HALO_HD void fp_from_canon(fp_t<P>& r, const uint32_t in[8]) {
    fp_t<P> t, r2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        t.v[i] = in[i];
        r2.v[i] = P::r2(i);
    }
    fp_mul(r, t, r2);
}
template <class P>
HALO_HD void fp_from_u32(fp_t<P>& r, uint32_t x) {
    uint32_t t[8] = {x, 0, 0, 0, 0, 0, 0, 0};
    fp_from_canon(r, t);
}

// a^(p-2) (Fermat): 254 squarings + ~65 multiplications, one dependent chain.  Kept as the independent cross-check of
// fp_inv below (tests) -- a lone warp needs 0.105 ms for it (profiles/r02_reduce_launches.csv, k_inv).
template <class P>
HALO_HD void fp_inv_fermat(fp_t<P>& r, const fp_t<P>& a) {
    // exponent p - 2: limbs(p) with limb0 = 1 - 2 -> borrow: p - 2 = [0xffffffff, P1-1, P2, P3, 0,0,0,0x40000000]
    uint32_t e[8] = {0xffffffffu, P::P1 - 1u, P::P2, P::P3, 0u, 0u, 0u, 0x40000000u};
    fp_t<P> acc;
    fp_one(acc);
    for (int i = 254; i >= 0; i--) {
        fp_sqr(acc, acc);
        if ((e[i >> 5] >> (i & 31)) & 1u) fp_mul(acc, acc, a);
    }
    r = acc;
}

// ---- inversion by division steps (Bernstein-Yang "safegcd", the 30-bit-limb formulation used for 256-bit moduli) --------
// Every serial point normalisation, the top of every batched inversion (k_inv: once per pair-tree pass, twice per generator
// fold) and the transcript's challenge inversions are ONE dependent chain; as a Fermat power that chain is ~78 000
// instructions.  The division-step algorithm replaces it by 20 rounds of { 30 branch-free steps on single 32-bit words that
// build a 2 x 2 transition matrix; apply the matrix to (f, g) and, modulo p, to (d, e) -- nine signed 30-bit limbs each }:
// ~15 000 instructions, no data-dependent branch (all lanes of a warp stay converged), 600 >= 590 steps suffice for any
// 256-bit input.  Integers only: the input is the Montgomery residue aR, the integer inverse (aR)^-1 is turned into the
// Montgomery residue a^-1 R by one multiplication with R^3.  inv(0) = 0, like the Fermat power.
namespace dsinv {
constexpr int32_t M30 = 0x3fffffff;
struct s30 {
    int32_t v[9];
};
template <class P>
HALO_HD constexpr int32_t mod30(int i) {  // limb i of p in 30-bit limbs (p = 2^254 + t, t < 2^126: limbs 5..7 are zero)
    // bits [30 i, 30 i + 30) of p, from the 32-bit limbs fp_mod<P>(.)
    const int lo = 30 * i, w = lo >> 5, sh = lo & 31;
    uint64_t x = (uint64_t)fp_mod<P>(w) >> sh;
    if (sh > 2 && w + 1 < 8) x |= (uint64_t)fp_mod<P>(w + 1) << (32 - sh);
    return (int32_t)(x & (uint64_t)M30);
}
HALO_HD void to30(s30& r, const uint32_t a[8]) {
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int lo = 30 * i, w = lo >> 5, sh = lo & 31;
        uint64_t x = (uint64_t)a[w] >> sh;
        if (sh > 2 && w + 1 < 8) x |= (uint64_t)a[w + 1] << (32 - sh);
        r.v[i] = (int32_t)(x & (uint64_t)M30);
    }
}
HALO_HD void from30(uint32_t r[8], const s30& a) {  // a normalised: limbs in [0, 2^30), value < 2^256
#pragma unroll
    for (int w = 0; w < 8; w++) {
        const int lo = 32 * w, i = lo / 30, sh = lo % 30;  // bit lo of the value is bit sh of limb i
        uint64_t x = (uint64_t)(uint32_t)a.v[i] >> sh;
        x |= (uint64_t)(uint32_t)a.v[i + 1] << (30 - sh);
        if (i + 2 < 9 && 60 - sh < 32) x |= (uint64_t)(uint32_t)a.v[i + 2] << (60 - sh);
        r[w] = (uint32_t)x;
    }
}
struct trans {
    int32_t u, v, q, r;
};
// 30 division steps on the low words; returns the new zeta and the transition matrix scaled by 2^30
HALO_HD int32_t divsteps30(int32_t zeta, uint32_t f0, uint32_t g0, trans& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        uint32_t m1 = (uint32_t)(zeta >> 31);       // all ones iff zeta < 0
        const uint32_t m2 = 0u - (g & 1u);          // all ones iff g odd
        const uint32_t x = (f ^ m1) - m1, y = (u ^ m1) - m1, z = (v ^ m1) - m1;  // (+-f, +-u, +-v)
        g += x & m2;
        q += y & m2;
        r += z & m2;
        m1 &= m2;                                    // swap iff zeta < 0 and g odd
        zeta = (int32_t)(((uint32_t)zeta ^ m1) - 1u);
        f += g & m1;
        u += q & m1;
        v += r & m1;
        g >>= 1;
        u <<= 1;
        v <<= 1;
    }
    t.u = (int32_t)u, t.v = (int32_t)v, t.q = (int32_t)q, t.r = (int32_t)r;
    return zeta;
}
// (d, e) <- t (d, e) / 2^30 (mod p), entries stay in (-2p, p)
template <class P>
HALO_HD void update_de(s30& d, s30& e, const trans& t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = u * d.v[0] + v * e.v[0], ce = q * d.v[0] + r * e.v[0];
    // p = 1 (mod 2^30), so p^-1 mod 2^30 = 1: choose md, me so that the low 30 bits of t (d, e) + p (md, me) vanish
    md -= (int32_t)(((uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
    me -= (int32_t)(((uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
    cd += (int64_t)mod30<P>(0) * md;
    ce += (int64_t)mod30<P>(0) * me;
    cd >>= 30;
    ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += u * d.v[i] + v * e.v[i];
        ce += q * d.v[i] + r * e.v[i];
        if (mod30<P>(i) != 0) {
            cd += (int64_t)mod30<P>(i) * md;
            ce += (int64_t)mod30<P>(i) * me;
        }
        d.v[i - 1] = (int32_t)cd & M30;
        cd >>= 30;
        e.v[i - 1] = (int32_t)ce & M30;
        ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}
// (f, g) <- t (f, g) / 2^30 (exact)
HALO_HD void update_fg(s30& f, s30& g, const trans& t) {
    const int64_t u = t.u, v = t.v, q = t.q, r = t.r;
    int64_t cf = u * f.v[0] + v * g.v[0], cg = q * f.v[0] + r * g.v[0];
    cf >>= 30;
    cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cf += u * f.v[i] + v * g.v[i];
        cg += q * f.v[i] + r * g.v[i];
        f.v[i - 1] = (int32_t)cf & M30;
        cf >>= 30;
        g.v[i - 1] = (int32_t)cg & M30;
        cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}
// r in (-2p, p) -> [0, p), negated first when sign < 0
template <class P>
HALO_HD void normalize(s30& r, int32_t sign) {
    int32_t c;
    const int32_t add1 = r.v[8] >> 31, neg = sign >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        r.v[i] += mod30<P>(i) & add1;
        r.v[i] = (r.v[i] ^ neg) - neg;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c = r.v[i] >> 30;
        r.v[i] &= M30;
        r.v[i + 1] += c;
    }
    const int32_t add2 = r.v[8] >> 31;
#pragma unroll
    for (int i = 0; i < 9; i++) r.v[i] += mod30<P>(i) & add2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c = r.v[i] >> 30;
        r.v[i] &= M30;
        r.v[i + 1] += c;
    }
}
}  // namespace dsinv

template <class P>
HALO_HD void fp_inv(fp_t<P>& r, const fp_t<P>& a) {
    using namespace dsinv;
    s30 d, e, f, g;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        d.v[i] = 0;
        e.v[i] = i == 0 ? 1 : 0;
        f.v[i] = mod30<P>(i);
    }
    to30(g, a.v);
    int32_t zeta = -1;
#pragma unroll 1
    for (int it = 0; it < 20; it++) {
        trans t;
        zeta = divsteps30(zeta, (uint32_t)f.v[0], (uint32_t)g.v[0], t);
        update_de<P>(d, e, t);
        update_fg(f, g, t);
    }
    // g = 0 and f = +-gcd = +-1 (or +-p when a = 0, where d = 0): the integer inverse is sign(f) * d
    normalize<P>(d, f.v[8]);
    fp_t<P> x, r3, r2;
    from30(x.v, d);
#pragma unroll
    for (int i = 0; i < 8; i++) r2.v[i] = P::r2(i);
    fp_mul(r3, r2, r2);  // R^2 * R^2 / R = R^3
    fp_mul(r, x, r3);    // (aR)^-1 * R^3 / R = a^-1 R
}

}  // namespace halo
