// host/group.hpp -- host mirror of code/src/group.rs: type aliases, the dot products (dispatched to the
// device through the C ABI), and the Fiat-Shamir macros rho_0! / rho_1! (host side, group.rs:41-89).
//
// Host arithmetic here is O(lg n) glue (challenges, inversions, a handful of point operations); the
// O(n) work is behind halo_b200.h.  Field / curve code is shared with the device (csrc/fp.cuh, ec.cuh,
// sha3.cuh compile as plain C++).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/halo_b200.h"
#include "../csrc/ec.cuh"
#include "../csrc/glv.cuh"
#include "../csrc/sha3.cuh"

namespace halo {

// group.rs:7-10
using PallasScalar = fr_t;
struct PallasPoint {  // Projective; kept as XYZZ on the host, Jacobian on the wire
    xyzz_t p;
};
using PallasPoly = std::vector<PallasScalar>;  // DensePolynomial coefficients, low degree first
struct PolyView {  // borrowed coefficients (no copy of 32 n bytes at the C boundary)
    const PallasScalar* data_;
    size_t size_;
    PolyView(const PallasScalar* d, size_t n) : data_(d), size_(n) {}
    PolyView(const PallasPoly& p) : data_(p.data()), size_(p.size()) {}
    const PallasScalar* data() const { return data_; }
    size_t size() const { return size_; }
    const PallasScalar& operator[](size_t i) const { return data_[i]; }
};

struct HaloFailure : std::runtime_error {  // carries a HALO_E* / HALO_REJECT_* code; the analogue of panic / Err
    int code;
    HaloFailure(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void ensure(bool ok, int code, const char* msg) {
    if (!ok) throw HaloFailure(code, msg);
}
inline void check_rc(halo_ctx* ctx, int rc) {
    if (rc != HALO_OK) throw HaloFailure(rc, halo_last_error(ctx));
}

// ---- scalars -----------------------------------------------------------------------------------
inline PallasScalar scalar_zero() { PallasScalar s; fp_zero(s); return s; }
inline PallasScalar scalar_one() { PallasScalar s; fp_one(s); return s; }
inline PallasScalar scalar_load(const uint64_t* p) { PallasScalar s; std::memcpy(&s, p, 32); return s; }
inline void scalar_store(uint64_t* p, const PallasScalar& s) { std::memcpy(p, &s, 32); }
inline PallasScalar operator*(const PallasScalar& a, const PallasScalar& b) { PallasScalar r; fp_mul(r, a, b); return r; }
inline PallasScalar operator+(const PallasScalar& a, const PallasScalar& b) { PallasScalar r; fp_add(r, a, b); return r; }
inline PallasScalar operator-(const PallasScalar& a, const PallasScalar& b) { PallasScalar r; fp_sub(r, a, b); return r; }
inline PallasScalar operator-(const PallasScalar& a) { PallasScalar r; fp_neg(r, a); return r; }
inline bool operator==(const PallasScalar& a, const PallasScalar& b) { return fp_eq(a, b); }
inline PallasScalar scalar_inverse(const PallasScalar& a) { PallasScalar r; fp_inv(r, a); return r; }

// ---- points --------------------------------------------------------------------------------------
inline PallasPoint point_load(const uint64_t* jac12) {
    jac_t j;
    std::memcpy(&j, jac12, 96);
    PallasPoint r;
    jac_to_xyzz(r.p, j);
    return r;
}
inline void point_store(uint64_t* jac12, const PallasPoint& a) {
    jac_t j;
    xyzz_to_jac(j, a.p);
    std::memcpy(jac12, &j, 96);
}
// ---- validation at the C boundary ------------------------------------------------------------------
// The reference's types cannot hold a non-canonical residue or an off-curve point (arkworks validates when a value is
// deserialised, and every in-memory value is built by field / group operations).  Raw limbs arriving through the C ABI
// carry no such guarantee, so the verifier-side entry points check them once: limbs < modulus, and for a Jacobian point
// z == 0 (infinity) or Y^2 = X^3 + 5 Z^6.  Failure is HALO_EINVAL (malformed input), never a verifier decision.
template <class P>
inline bool fp_is_canonical(const fp_t<P>& a) {
    uint32_t m[8], t[8];
    fp_mod_limbs<P>(m);
    return sub8(t, a.v, m) != 0;  // borrow <=> a < modulus
}
inline bool scalar_valid(const uint64_t* p) { return fp_is_canonical(scalar_load(p)); }
inline bool point_valid(const uint64_t* jac12) {
    jac_t j;
    std::memcpy(&j, jac12, 96);
    if (!fp_is_canonical(j.x) || !fp_is_canonical(j.y) || !fp_is_canonical(j.z)) return false;
    if (fp_is_zero(j.z)) return true;  // infinity (arkworks: any (x, y, 0))
    fq_t lhs, rhs, z2, z6, five;
    fp_sqr(lhs, j.y);
    fp_sqr(rhs, j.x);
    fp_mul(rhs, rhs, j.x);
    fp_sqr(z2, j.z);
    fp_sqr(z6, z2);
    fp_mul(z6, z6, z2);
    fp_from_u32(five, 5);
    fp_mul(z6, z6, five);
    fp_add(rhs, rhs, z6);
    return fp_eq(lhs, rhs);
}
inline PallasScalar scalar_load_checked(const uint64_t* p, const char* what) {
    ensure(p != nullptr && scalar_valid(p), HALO_EINVAL, what);
    return scalar_load(p);
}
inline PallasPoint point_load_checked(const uint64_t* jac12, const char* what) {
    ensure(jac12 != nullptr && point_valid(jac12), HALO_EINVAL, what);
    return point_load(jac12);
}

inline PallasPoint point_zero() { PallasPoint r; xyzz_set_inf(r.p); return r; }
inline PallasPoint operator+(const PallasPoint& a, const PallasPoint& b) { PallasPoint r = a; xyzz_add(r.p, b.p); return r; }
inline PallasPoint operator-(const PallasPoint& a) { PallasPoint r = a; if (!xyzz_is_inf(r.p)) xyzz_neg(r.p); return r; }
inline PallasPoint operator-(const PallasPoint& a, const PallasPoint& b) { return a + (-b); }
// `Projective * Fr`
inline PallasPoint operator*(const PallasPoint& a, const PallasScalar& k) {
    PallasPoint r;
    xyzz_mul_glv(r.p, a.p, k);  // GLV + joint sparse form (csrc/glv.cuh): half the doublings and additions of double-and-add
    return r;
}
// Projective equality (representation independent)
inline bool operator==(const PallasPoint& a, const PallasPoint& b) {
    bool ia = xyzz_is_inf(a.p), ib = xyzz_is_inf(b.p);
    if (ia || ib) return ia && ib;
    fq_t l, r;
    fp_mul(l, a.p.x, b.p.zz);
    fp_mul(r, b.p.x, a.p.zz);
    if (!fp_eq(l, r)) return false;
    fp_mul(l, a.p.y, b.p.zzz);
    fp_mul(r, b.p.y, a.p.zzz);
    return fp_eq(l, r);
}

// ark-serialize `serialize_compressed` of a short-Weierstrass point for Pallas (ark-ec 0.5): normalise,
// then Fp::serialize_with_flags(x, SWFlags) -> ceil((255 + 2) / 8) = 33 bytes: 32 bytes little-endian canonical x
// and a final byte holding only the flags: 0x80 if y > -y (as canonical integers), 0x40 for infinity (x = 0).
// [arkworks is not in the reference tree; restated from the published crate, see DESIGN.md "parity unpinned".]
inline void serialize_compressed(const PallasPoint& a, uint8_t out[33]) {
    std::memset(out, 0, 33);
    affine_t aff;
    xyzz_to_affine(aff, a.p);
    if (xyzz_is_inf(a.p)) {
        out[32] = 0x40;
        return;
    }
    uint32_t x[8], y[8], ny[8];
    fq_t negy;
    fp_to_canon(x, aff.x);
    fp_to_canon(y, aff.y);
    fp_neg(negy, aff.y);
    fp_to_canon(ny, negy);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(x[i] >> (8 * b));
    bool gt = false;
    for (int i = 7; i >= 0; i--) {
        if (y[i] != ny[i]) {
            gt = y[i] > ny[i];
            break;
        }
    }
    if (gt) out[32] = 0x80;
}

// Montgomery's trick: all inverses for one field inversion (zero entries stay zero).
template <class P>
inline void batch_inverse(std::vector<fp_t<P>>& v) {
    std::vector<fp_t<P>> prefix(v.size());
    fp_t<P> acc;
    fp_one(acc);
    for (size_t i = 0; i < v.size(); i++) {
        prefix[i] = acc;
        if (!fp_is_zero(v[i])) fp_mul(acc, acc, v[i]);
    }
    fp_t<P> inv;
    fp_inv(inv, acc);
    for (size_t i = v.size(); i-- > 0;) {
        if (fp_is_zero(v[i])) continue;
        fp_t<P> t;
        fp_mul(t, inv, prefix[i]);
        fp_mul(inv, inv, v[i]);
        v[i] = t;
    }
}
// Normalise many points with one inversion; infinity -> (0, 0).
inline std::vector<affine_t> batch_to_affine(const std::vector<PallasPoint>& pts) {
    std::vector<fq_t> d(pts.size());
    for (size_t i = 0; i < pts.size(); i++) fp_mul(d[i], pts[i].p.zz, pts[i].p.zzz);  // zero for infinity
    batch_inverse(d);
    std::vector<affine_t> out(pts.size());
    for (size_t i = 0; i < pts.size(); i++) {
        if (xyzz_is_inf(pts[i].p)) {
            affine_set_inf(out[i]);
            continue;
        }
        fq_t t;
        fp_mul(t, d[i], pts[i].p.zzz);  // 1 / zz
        fp_mul(out[i].x, pts[i].p.x, t);
        fp_mul(t, d[i], pts[i].p.zz);   // 1 / zzz
        fp_mul(out[i].y, pts[i].p.y, t);
    }
    return out;
}
// compressed form of an already normalised point (same bytes as serialize_compressed)
inline void serialize_compressed_affine(const affine_t& aff, uint8_t out[33]) {
    std::memset(out, 0, 33);
    if (affine_is_inf(aff)) {
        out[32] = 0x40;
        return;
    }
    uint32_t x[8], y[8], ny[8];
    fq_t negy;
    fp_to_canon(x, aff.x);
    fp_to_canon(y, aff.y);
    fp_neg(negy, aff.y);
    fp_to_canon(ny, negy);
    for (int i = 0; i < 8; i++)
        for (int b = 0; b < 4; b++) out[4 * i + b] = (uint8_t)(x[i] >> (8 * b));
    bool gt = false;
    for (int i = 7; i >= 0; i--) {
        if (y[i] != ny[i]) {
            gt = y[i] > ny[i];
            break;
        }
    }
    if (gt) out[32] = 0x80;
}

// ---- Fiat-Shamir transcript: group.rs:41-64 (rho_0!, tag 0) and :66-89 (rho_1!, tag 1) -----------------
class Transcript {
    std::vector<uint8_t> data_;

public:
    Transcript& point(const PallasPoint& p) {
        uint8_t b[33];
        serialize_compressed(p, b);
        data_.insert(data_.end(), b, b + 33);
        return *this;
    }
    Transcript& point_affine(const affine_t& a) {  // a point normalised earlier (batch_to_affine)
        uint8_t b[33];
        serialize_compressed_affine(a, b);
        data_.insert(data_.end(), b, b + 33);
        return *this;
    }
    Transcript& scalar(const PallasScalar& s) {  // Fr::serialize_compressed: 32 bytes little-endian canonical
        uint32_t c[8];
        fp_to_canon(c, s);
        for (int i = 0; i < 8; i++)
            for (int b = 0; b < 4; b++) data_.push_back((uint8_t)(c[i] >> (8 * b)));
        return *this;
    }
    Transcript& u64(uint64_t v) {  // Vec length prefix / usize
        for (int b = 0; b < 8; b++) data_.push_back((uint8_t)(v >> (8 * b)));
        return *this;
    }
    Transcript& u8(uint8_t v) {  // Option discriminant
        data_.push_back(v);
        return *this;
    }
    // SHA3-256(data || u32_le(tag)) -> from_le_bytes_mod_order (group.rs:52-60)
    PallasScalar finish(uint32_t tag) {
        for (int b = 0; b < 4; b++) data_.push_back((uint8_t)(tag >> (8 * b)));
        uint64_t dg[4];
        sha3_256(data_.data(), data_.size(), dg);
        uint32_t s[8], m[8], t[8];
        for (int j = 0; j < 4; j++) {
            s[2 * j] = (uint32_t)dg[j];
            s[2 * j + 1] = (uint32_t)(dg[j] >> 32);
        }
        fp_mod_limbs<ScalarParams>(m);
        for (int it = 0; it < 3; it++) {  // r > 2^254: at most three subtractions
            uint32_t borrow = sub8(t, s, m);
            for (int j = 0; j < 8; j++) s[j] = borrow ? s[j] : t[j];
        }
        PallasScalar r;
        fp_from_canon(r, s);
        return r;
    }
};

// ---- group.rs functions ----------------------------------------------------------------------------------
// group.rs:13-15
inline PallasScalar scalar_dot(halo_ctx* ctx, const PallasScalar* xs, const PallasScalar* ys, uint64_t n) {
    uint64_t out[4];
    check_rc(ctx, halo_scalar_dot(ctx, reinterpret_cast<const uint64_t*>(xs), reinterpret_cast<const uint64_t*>(ys), n, out));
    return scalar_load(out);
}
// group.rs:18-21 (Jacobian bases on the wire)
inline PallasPoint point_dot(halo_ctx* ctx, const PallasScalar* xs, const std::vector<PallasPoint>& Gs, uint64_t n) {
    if (n <= 4) {  // a handful of terms (acc.rs:178 has m + 1 <= 3): host double-and-add, no device round trip
        PallasPoint acc = point_zero();
        for (uint64_t i = 0; i < n; i++) acc = acc + Gs[i] * xs[i];
        return acc;
    }
    std::vector<uint64_t> jac(12 * n);
    for (uint64_t i = 0; i < n; i++) point_store(&jac[12 * i], Gs[i]);
    uint64_t out[12];
    check_rc(ctx, halo_msm_jac(ctx, jac.data(), reinterpret_cast<const uint64_t*>(xs), n, out));
    return point_load(out);
}
// group.rs:29-37
inline std::vector<PallasScalar> construct_powers(const PallasScalar& z, uint64_t n) {
    std::vector<PallasScalar> zs(n);
    PallasScalar cur = scalar_one();
    for (uint64_t i = 0; i < n; i++) {
        zs[i] = cur;
        cur = cur * z;
    }
    return zs;
}

// consts.rs S, H of the context
inline void params_SH(halo_ctx* ctx, PallasPoint& S, PallasPoint& H) {
    uint64_t s[12], h[12];
    check_rc(ctx, halo_get_SH(ctx, s, h));
    S = point_load(s);
    H = point_load(h);
}

}  // namespace halo
