// host/acc.cpp -- see acc.hpp.  Control flow follows code/src/acc.rs step by step.
#include "acc.hpp"

namespace halo {
namespace acc {

static uint32_t ilog2(uint64_t n) {
    uint32_t l = 0;
    while (((uint64_t)1 << (l + 1)) <= n) l++;
    return l;
}

PallasPoly AccumulatedHPolys::get_poly(halo_ctx* ctx, uint32_t lg_n) const {
    size_t n = (size_t)1 << lg_n;
    PallasPoly out(n);
    std::vector<PallasScalar> xis;
    xis.reserve(hs.size() * (lg_n + 1));
    for (const auto& h : hs) {
        ensure(h.xis.size() == lg_n + 1, HALO_EINVAL, "HPoly size mismatch");
        xis.insert(xis.end(), h.xis.begin(), h.xis.end());
    }
    size_t n_h0 = have_h0 ? (h_0.size() < n ? h_0.size() : n) : 0;
    check_rc(ctx, halo_h_lincomb(ctx, reinterpret_cast<const uint64_t*>(h_0.data()), n_h0,
                                 reinterpret_cast<const uint64_t*>(alphas.data()),
                                 reinterpret_cast<const uint64_t*>(xis.data()), hs.size(), lg_n,
                                 reinterpret_cast<uint64_t*>(out.data())));
    return out;
}

uint64_t AccumulatedHPolys::get_poly_resident(halo_ctx* ctx, uint32_t lg_n) const {
    size_t n = (size_t)1 << lg_n;
    std::vector<PallasScalar> xis;
    xis.reserve(hs.size() * (lg_n + 1));
    for (const auto& h : hs) {
        ensure(h.xis.size() == lg_n + 1, HALO_EINVAL, "HPoly size mismatch");
        xis.insert(xis.end(), h.xis.begin(), h.xis.end());
    }
    size_t n_h0 = have_h0 ? (h_0.size() < n ? h_0.size() : n) : 0;
    uint64_t deg = 0;
    check_rc(ctx, halo_h_lincomb_resident(ctx, reinterpret_cast<const uint64_t*>(h_0.data()), n_h0,
                                          reinterpret_cast<const uint64_t*>(alphas.data()),
                                          reinterpret_cast<const uint64_t*>(xis.data()), hs.size(), lg_n, &deg));
    return deg;
}

PallasScalar AccumulatedHPolys::eval(const PallasScalar& z) const {
    PallasScalar v = scalar_zero();
    if (have_h0) {  // h_0.evaluate(z), Horner
        PallasScalar acc = scalar_zero();
        for (size_t i = h_0.size(); i-- > 0;) acc = acc * z + h_0[i];
        v = v + acc;
    }
    for (size_t i = 0; i < hs.size(); i++) v = v + hs[i].eval(z) * alphas[i + 1];
    return v;
}

// Field order of the derive: h_0: Option<PallasPoly>, hs: Vec<HPoly>, alpha: Option<PallasScalar>, alphas: Vec<PallasScalar>.
// Option -> 1 byte then the value; Vec -> u64 LE length then elements; DensePolynomial -> its coeffs Vec (trailing
// zeros trimmed by from_coefficients_vec); HPoly -> its xis Vec.  [ark-serialize 0.5, restated; parity unpinned]
void AccumulatedHPolys::serialize(Transcript& t) const {
    t.u8(have_h0 ? 1 : 0);
    if (have_h0) {
        size_t len = h_0.size();
        while (len > 0 && fp_is_zero(h_0[len - 1])) len--;
        t.u64(len);
        for (size_t i = 0; i < len; i++) t.scalar(h_0[i]);
    }
    t.u64(hs.size());
    for (const auto& h : hs) {
        t.u64(h.xis.size());
        for (const auto& x : h.xis) t.scalar(x);
    }
    t.u8(have_alpha ? 1 : 0);
    if (have_alpha) t.scalar(alpha);
    t.u64(alphas.size());
    for (const auto& a : alphas) t.scalar(a);
}

struct CommonOut {
    PallasPoint C_bar;
    uint64_t d;
    PallasScalar z;
    AccumulatedHPolys hs;
};

// acc.rs:135-188
static CommonOut common_subroutine(halo_ctx* ctx, uint64_t d, const std::vector<Instance>& qs, const AccumulatorHiding& pi_V) {
    size_t m = qs.size();
    AccumulatedHPolys hs(m);
    std::vector<PallasPoint> Us;
    Us.reserve(m + 1);
    // (2) parse pi_V (:146-149)
    hs.h_0 = pi_V.h;
    hs.have_h0 = true;
    Us.push_back(pi_V.U);
    // (3) and (4) need m + 1 small MSMs, none of which depends on another's result: one device round trip for all of
    // them (pcdl::succinct_check_many), the `ensure!`s raised afterwards in the reference's order.  Input that the
    // reference would panic on goes through the one-by-one path so that the first failure is the reference's.
    bool batched = false;
    if (m >= 1 && pi_V.h.size() <= 4096) {
        try {
            std::vector<pcdl::Query> queries;
            for (const auto& q : qs) queries.push_back(pcdl::Query{&q.C, q.d, &q.z, &q.v, &q.pi});
            PolyView h0(pi_V.h);
            pcdl::SuccinctMany sm = pcdl::succinct_check_many(ctx, queries, &h0, d);
            batched = true;
            // (3) U_0 == PCDL.Commit(h_0, d, None) (:152-155)
            ensure(pi_V.U == sm.commitment, HALO_REJECT_U0, "U_0 != PCDL.Commit_rho0(ck^(1)_PC, h_0; w = bot)");
            // (4) (:158-170)
            for (size_t i = 0; i < m; i++) {
                ensure(sm.accept[i], HALO_REJECT_SUCCINCT, "C_(log_n) != CM.Commit_Sigma(c || v')");  // pcdl.rs:307-310
                hs.hs.push_back(sm.hu[i].first);
                Us.push_back(sm.hu[i].second);
                ensure(qs[i].d == d, HALO_REJECT_D, "d_i != d");  // :169
            }
        } catch (const HaloFailure& e) {
            if (batched || e.code <= HALO_REJECT_SUCCINCT) throw;  // a decision, already in the reference's order
            hs.hs.clear();
            Us.resize(1);
        }
    }
    if (!batched) {
        // (3) U_0 == PCDL.Commit(h_0, d, None) (:152-155)
        ensure(pi_V.U == pcdl::commit(ctx, pi_V.h, d, nullptr), HALO_REJECT_U0, "U_0 != PCDL.Commit_rho0(ck^(1)_PC, h_0; w = bot)");
        // (4) (:158-170)
        for (const auto& q : qs) {
            auto hu = pcdl::succinct_check(ctx, q.C, q.d, q.z, q.v, q.pi);
            hs.hs.push_back(hu.first);
            Us.push_back(hu.second);
            ensure(q.d == d, HALO_REJECT_D, "d_i != d");  // :169
        }
    }
    // (6) alpha = rho_1(hs) (:173)
    Transcript t;
    hs.serialize(t);
    hs.set_alpha(t.finish(1));
    // (8) C = sum alpha^i U_i (:178)
    PallasPoint C = point_dot(ctx, hs.alphas.data(), Us, Us.size() < hs.alphas.size() ? Us.size() : hs.alphas.size());
    // (9) z = rho_1(C, alpha) (:181)
    PallasScalar z = Transcript().point(C).scalar(hs.alpha).finish(1);
    // (10) C_bar = C + w S (:184)
    PallasPoint S, H;
    params_SH(ctx, S, H);
    PallasPoint C_bar = C + S * pi_V.w;
    return CommonOut{C_bar, d, z, std::move(hs)};
}

Accumulator prover(halo_ctx* ctx, uint64_t d, const std::vector<Instance>& qs, const PallasPoly& h_0, const PallasScalar& w,
                   PolyView q, const PallasScalar& w_bar) {
    ensure(h_0.size() == 2, HALO_EINVAL, "h_0 must be PallasPoly::rand(1): two coefficients");  // :192
    // U_0 = PCDL.Commit(h_0, d, None) (:195)
    PallasPoint U_0 = pcdl::commit(ctx, h_0, d, nullptr);
    AccumulatorHiding pi_V{h_0, U_0, w};  // :198-199
    CommonOut c = common_subroutine(ctx, d, qs, pi_V);  // :202
    PallasScalar v = c.hs.eval(c.z);                    // :205
    // pi = PCDL.Open(h(X), C_bar, d, z; w) (:209)
    // h.get_poly() (:85-94) is expanded on the device and opened from there
    uint64_t deg = c.hs.get_poly_resident(ctx, ilog2(d + 1));
    pcdl::EvalProof pi = pcdl::open_resident(ctx, deg, c.C_bar, d, c.z, &w, &q, &w_bar);
    return Accumulator{c.C_bar, c.d, c.z, v, pi, pi_V};  // :212-219
}

void verifier(halo_ctx* ctx, uint64_t D, const std::vector<Instance>& qs, const Accumulator& acc) {
    CommonOut c = common_subroutine(ctx, D, qs, acc.pi_V);  // :234
    ensure(c.C_bar == acc.C_bar, HALO_REJECT_CBAR, "C_bar' != C_bar");  // :237
    ensure(c.z == acc.z, HALO_REJECT_Z, "z' = z");                       // :238
    ensure(c.d == acc.d, HALO_REJECT_D, "d' = d");                       // :239
    ensure(c.hs.eval(acc.z) == acc.v, HALO_REJECT_V, "h(z) = v");        // :240
}

void decider(halo_ctx* ctx, const Accumulator& acc) {
    pcdl::check(ctx, acc.C_bar, acc.d, acc.z, acc.v, acc.pi);  // :254
}

Instance instance_from_c(const halo_instance& q) {
    return Instance{point_load(q.C), q.d, scalar_load(q.z), scalar_load(q.v), pcdl::proof_from_c(q.pi)};
}
Accumulator accumulator_from_c(const halo_accumulator& a) {
    AccumulatorHiding pv{{scalar_load(a.h0[0]), scalar_load(a.h0[1])}, point_load(a.U0), scalar_load(a.w)};
    return Accumulator{point_load(a.C_bar), a.d, scalar_load(a.z), scalar_load(a.v), pcdl::proof_from_c(a.pi), pv};
}
void accumulator_to_c(const Accumulator& a, halo_accumulator& out) {
    std::memset(&out, 0, sizeof out);
    point_store(out.C_bar, a.C_bar);
    out.d = a.d;
    scalar_store(out.z, a.z);
    scalar_store(out.v, a.v);
    pcdl::proof_to_c(a.pi, out.pi);
    for (size_t i = 0; i < 2 && i < a.pi_V.h.size(); i++) scalar_store(out.h0[i], a.pi_V.h[i]);
    point_store(out.U0, a.pi_V.U);
    scalar_store(out.w, a.pi_V.w);
}

}  // namespace acc
}  // namespace halo
