// host/pcdl.cpp -- see pcdl.hpp.  Control flow follows code/src/pcdl.rs step by step; every O(n) step is a
// call into libhalo_b200.so.
#include "pcdl.hpp"

#include "pedersen.hpp"

namespace halo {
namespace pcdl {

static bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }
static uint32_t ilog2(uint64_t n) {
    uint32_t l = 0;
    while (((uint64_t)1 << (l + 1)) <= n) l++;
    return l;
}
static uint64_t degree(PolyView p) {  // DensePolynomial::degree (trailing zeros trimmed)
    uint64_t n = p.size();
    while (n > 0 && fp_is_zero(p[n - 1])) n--;
    return n ? n - 1 : 0;
}

PallasPoly HPoly::get_poly(halo_ctx* ctx) const {
    uint32_t lg_n = (uint32_t)xis.size() - 1;
    PallasPoly out((size_t)1 << lg_n);
    check_rc(ctx, halo_h_expand(ctx, reinterpret_cast<const uint64_t*>(xis.data()), lg_n, reinterpret_cast<uint64_t*>(out.data())));
    return out;
}

PallasScalar HPoly::eval(const PallasScalar& z) const {
    size_t lg_n = xis.size() - 1;
    PallasScalar one = scalar_one();
    PallasScalar v = one + xis[lg_n] * z;
    PallasScalar z_i = z;
    for (size_t i = 1; i < lg_n; i++) {
        z_i = z_i * z_i;
        v = v * (one + xis[lg_n - i] * z_i);
    }
    return v;
}

PallasPoint commit(halo_ctx* ctx, PolyView p, uint64_t d, const PallasScalar* w) {
    uint64_t n = d + 1;
    ensure(is_pow2(n), HALO_EINVAL, "d+1 is not a power of 2");  // pcdl.rs:102
    ensure(degree(p) <= d, HALO_EINVAL, "p.degree() > d");       // pcdl.rs:103
    ensure(n <= halo_num_generators(ctx), HALO_EINVAL, "d > D");  // pcdl.rs:104
    uint64_t n_coeffs = p.size() < n ? p.size() : n;
    // coeffs.resize(n, ZERO) (pcdl.rs:106-107): the zero tail is implied, not uploaded
    return pedersen::commit(ctx, w, nullptr, n, p.data(), n_coeffs, true);
}

static EvalProof open_impl(halo_ctx* ctx, const PolyView* p, uint64_t deg, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                           const PallasScalar* w, const PolyView* q, const PallasScalar* w_bar);

EvalProof open(halo_ctx* ctx, PolyView p, const PallasPoint& C, uint64_t d, const PallasScalar& z, const PallasScalar* w,
               const PolyView* q, const PallasScalar* w_bar) {
    return open_impl(ctx, &p, degree(p), C, d, z, w, q, w_bar);
}
EvalProof open_resident(halo_ctx* ctx, uint64_t deg, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                        const PallasScalar* w, const PolyView* q, const PallasScalar* w_bar) {
    return open_impl(ctx, nullptr, deg, C, d, z, w, q, w_bar);
}

static EvalProof open_impl(halo_ctx* ctx, const PolyView* p, uint64_t deg, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                           const PallasScalar* w, const PolyView* q, const PallasScalar* w_bar) {
    uint64_t n = d + 1;
    ensure(is_pow2(n), HALO_EINVAL, "d+1 is not a power of 2");  // pcdl.rs:130
    ensure(deg <= d, HALO_EINVAL, "p.degree() > d");             // pcdl.rs:131
    ensure(n <= halo_num_generators(ctx), HALO_EINVAL, "d > D");  // pcdl.rs:132
    uint32_t lg_n = ilog2(n);
    PallasPoint S, H;
    params_SH(ctx, S, H);

    // device state: c (zero padded), G, z-powers; v = p(z) (pcdl.rs:135, :183-186)
    halo_ipa* st = nullptr;
    uint64_t vbuf[4];
    if (p) {
        uint64_t n_coeffs = p->size() < n ? p->size() : n;
        check_rc(ctx, halo_ipa_begin(ctx, reinterpret_cast<const uint64_t*>(p->data()), n_coeffs, n,
                                     reinterpret_cast<const uint64_t*>(&z), &st, vbuf));
    } else {
        check_rc(ctx, halo_ipa_begin_resident(ctx, n, reinterpret_cast<const uint64_t*>(&z), &st, vbuf));
    }
    struct Guard {
        halo_ipa* s;
        ~Guard() { halo_ipa_destroy(s); }
    } guard{st};
    PallasScalar v = scalar_load(vbuf);

    EvalProof pi;
    PallasPoint C_prime = C;
    if (w) {  // pcdl.rs:137-164
        ensure(q && w_bar, HALO_EINVAL, "hiding open needs the random polynomial q and w_bar");
        ensure(deg >= 1 && q->size() == deg, HALO_ELEN, "q must have deg(p) coefficients (PallasPoly::rand(p.degree() - 1))");
        // p_bar = q (X - z); C_bar = commit(p_bar, d, w_bar)  (:140-149)
        uint64_t cb[12];
        check_rc(ctx, halo_ipa_blind_commit(st, reinterpret_cast<const uint64_t*>(q->data()), q->size(), cb));
        PallasPoint C_bar = S * (*w_bar) + point_load(cb);
        // alpha = rho_0(C, z, v, C_bar)  (:153)
        PallasScalar a = Transcript().point(C).scalar(z).scalar(v).point(C_bar).finish(0);
        // p' = p + alpha p_bar  (:156)
        check_rc(ctx, halo_ipa_blind_apply(st, reinterpret_cast<const uint64_t*>(&a)));
        // w' = w_bar alpha + w ; C' = C + alpha C_bar - w' S  (:159-162)
        PallasScalar w_prime = (*w_bar) * a + (*w);
        C_prime = C + C_bar * a - S * w_prime;
        pi.hiding = true;
        pi.C_bar = C_bar;
        pi.w_prime = w_prime;
    }

    // xi_0 = rho_0(C', z, v); H' = xi_0 H  (:180-181)
    PallasScalar xi_i = Transcript().point(C_prime).scalar(z).scalar(v).finish(0);
    PallasPoint H_prime = H * xi_i;
    uint64_t hp[12];
    point_store(hp, H_prime);
    check_rc(ctx, halo_ipa_set_hprime(st, hp));

    pi.Ls.reserve(lg_n);
    pi.Rs.reserve(lg_n);
    for (uint32_t round = 0; round < lg_n; round++) {  // :195-227
        uint64_t Lb[12], Rb[12];
        check_rc(ctx, halo_ipa_round_lr(st, Lb, Rb));  // :199-209
        PallasPoint L = point_load(Lb), R = point_load(Rb);
        pi.Ls.push_back(L);
        pi.Rs.push_back(R);
        PallasScalar xi_next = Transcript().scalar(xi_i).point(L).point(R).finish(0);  // :212
        ensure(!fp_is_zero(xi_next), HALO_EINVAL, "challenge is zero (inverse().unwrap())");  // :213
        PallasScalar xi_next_inv = scalar_inverse(xi_next);
        xi_i = xi_next;
        check_rc(ctx, halo_ipa_round_fold(st, reinterpret_cast<const uint64_t*>(&xi_next),
                                          reinterpret_cast<const uint64_t*>(&xi_next_inv)));  // :216-224
    }
    uint64_t Ub[12], cb[4];
    check_rc(ctx, halo_ipa_finish(st, Ub, cb));  // :230-231
    pi.U = point_load(Ub);
    pi.c = scalar_load(cb);
    return pi;
}

// Everything of succinct_check (pcdl.rs:252-314) up to its group equation, which is left as one small MSM:
//   C_lg == c U + v' H'   <=>   C' + sum_k scalars[k] * aff[k] == 0
struct SuccinctPrep {
    HPoly h;
    PallasPoint U, C_prime;
    std::vector<affine_t> aff;
    std::vector<uint8_t> inf;
    std::vector<PallasScalar> scalars;
};
static SuccinctPrep succinct_prepare(halo_ctx* ctx, const PallasPoint& C, uint64_t d, const PallasScalar& z, const PallasScalar& v,
                                     const EvalProof& pi) {
    uint64_t n = d + 1;
    ensure(is_pow2(n), HALO_EINVAL, "d+1 is not a power of 2!");           // :261
    ensure(n <= halo_num_generators(ctx), HALO_EINVAL, "d was larger than D!");  // :262
    uint32_t lg_n = ilog2(n);
    ensure(pi.Ls.size() == lg_n && pi.Rs.size() == lg_n, HALO_EINVAL, "proof has the wrong number of rounds");
    PallasPoint S, H;
    params_SH(ctx, S, H);

    // C' = C + alpha C_bar - w' S  (:272-279)
    PallasPoint C_prime = C;
    if (pi.hiding) {
        PallasScalar a = Transcript().point(C).scalar(z).scalar(v).point(pi.C_bar).finish(0);
        C_prime = C + pi.C_bar * a - S * pi.w_prime;
    }
    // xi_0, challenges (:282-293); the group arithmetic of :285-298 and :307-310 is one small MSM:
    //   C_lg == c U + v' H'   <=>   C' + (xi_0 v - xi_0 v') H + sum_i (xi_{i+1}^-1 L_i + xi_{i+1} R_i) - c U == 0
    // all points that enter the transcript or the MSM are normalised with ONE field inversion (the reference normalises
    // each point separately inside serialize_compressed; same bytes)
    std::vector<PallasPoint> pts;
    pts.reserve(2 * (size_t)lg_n + 2);
    for (uint32_t i = 0; i < lg_n; i++) {
        pts.push_back(pi.Ls[i]);
        pts.push_back(pi.Rs[i]);
    }
    pts.push_back(H);
    pts.push_back(pi.U);
    std::vector<affine_t> aff = batch_to_affine(pts);
    std::vector<uint8_t> inf(aff.size());
    for (size_t i = 0; i < aff.size(); i++) inf[i] = affine_is_inf(aff[i]) ? 1 : 0;
    std::vector<PallasScalar> xis;
    xis.reserve(lg_n + 1);
    xis.push_back(Transcript().point(C_prime).scalar(z).scalar(v).finish(0));
    for (uint32_t i = 0; i < lg_n; i++) {
        PallasScalar xi_next = Transcript().scalar(xis[i]).point_affine(aff[2 * i]).point_affine(aff[2 * i + 1]).finish(0);  // :293
        ensure(!fp_is_zero(xi_next), HALO_EINVAL, "challenge is zero (inverse().unwrap())");                                 // :297
        xis.push_back(xi_next);
    }
    std::vector<PallasScalar> inv(xis.begin() + 1, xis.end());
    batch_inverse(inv);  // xi_{i+1}^-1, one inversion for all rounds
    std::vector<PallasScalar> scalars;
    scalars.reserve(2 * (size_t)lg_n + 2);
    for (uint32_t i = 0; i < lg_n; i++) {
        scalars.push_back(inv[i]);      // xi_{i+1}^-1 L_i
        scalars.push_back(xis[i + 1]);  // xi_{i+1}   R_i
    }
    HPoly h(xis);                                   // :301
    PallasScalar v_prime = pi.c * h.eval(z);        // :304
    scalars.push_back(xis[0] * v - xis[0] * v_prime);  // v H' - v' H' with H' = xi_0 H  (:285, :288, :308)
    scalars.push_back(-pi.c);                          // - c U
    return SuccinctPrep{std::move(h), pi.U, C_prime, std::move(aff), std::move(inf), std::move(scalars)};
}

std::pair<HPoly, PallasPoint> succinct_check(halo_ctx* ctx, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                                             const PallasScalar& v, const EvalProof& pi) {
    SuccinctPrep sp = succinct_prepare(ctx, C, d, z, v, pi);
    uint64_t out[12];
    check_rc(ctx, halo_msm(ctx, reinterpret_cast<const uint64_t*>(sp.aff.data()), sp.inf.data(),
                           reinterpret_cast<const uint64_t*>(sp.scalars.data()), sp.aff.size(), out));
    PallasPoint lhs = sp.C_prime + point_load(out);
    ensure(xyzz_is_inf(lhs.p), HALO_REJECT_SUCCINCT, "C_(log_n) != CM.Commit_Sigma(c || v')");  // :307-310
    return {sp.h, sp.U};                                                                        // :313
}

SuccinctMany succinct_check_many(halo_ctx* ctx, const std::vector<Query>& qs, const PolyView* commit_p, uint64_t commit_d) {
    std::vector<SuccinctPrep> preps;
    preps.reserve(qs.size());
    for (const Query& q : qs) preps.push_back(succinct_prepare(ctx, *q.C, q.d, *q.z, *q.v, *q.pi));
    std::vector<halo_msm_desc> descs;
    for (const SuccinctPrep& sp : preps)
        descs.push_back(halo_msm_desc{reinterpret_cast<const uint64_t*>(sp.aff.data()), sp.inf.data(),
                                      reinterpret_cast<const uint64_t*>(sp.scalars.data()), sp.aff.size(), 0});
    if (commit_p) {  // pcdl::commit without hiding: <coeffs, GS[0..)> over the resident generators
        uint64_t n = commit_d + 1;
        ensure(is_pow2(n), HALO_EINVAL, "d+1 is not a power of 2");          // pcdl.rs:102
        ensure(degree(*commit_p) <= commit_d, HALO_EINVAL, "p.degree() > d");  // pcdl.rs:103
        ensure(n <= halo_num_generators(ctx), HALO_EINVAL, "d > D");          // pcdl.rs:104
        uint64_t n_coeffs = commit_p->size() < n ? commit_p->size() : n;
        ensure(n_coeffs <= 4096, HALO_EINVAL, "succinct_check_many: the commitment riding along must be short");
        descs.push_back(halo_msm_desc{nullptr, nullptr, reinterpret_cast<const uint64_t*>(commit_p->data()), n_coeffs, 0});
    }
    std::vector<uint64_t> out(12 * descs.size() + 12);
    check_rc(ctx, halo_msm_multi(ctx, descs.data(), (uint32_t)descs.size(), out.data()));
    SuccinctMany r;
    for (size_t i = 0; i < preps.size(); i++) {
        PallasPoint lhs = preps[i].C_prime + point_load(&out[12 * i]);
        r.accept.push_back(xyzz_is_inf(lhs.p));  // :307-310
        r.hu.emplace_back(preps[i].h, preps[i].U);
    }
    if (commit_p) r.commitment = point_load(&out[12 * preps.size()]);
    return r;
}

void check(halo_ctx* ctx, const PallasPoint& C, uint64_t d, const PallasScalar& z, const PallasScalar& v,
           const EvalProof& pi) {
    // succinct_check (:332) and comm = pedersen::commit(None, GS[0..d+1], h.get_poly().coeffs) (:338).  The challenges
    // (hence h) need only the transcript, so the small MSM of the succinct check and the expansion + MSM <G, h> run
    // concurrently on the device; the two `ensure!`s are evaluated in the reference's order.
    SuccinctPrep sp = succinct_prepare(ctx, C, d, z, v, pi);
    uint64_t out_h[12], out_s[12];
    check_rc(ctx, halo_h_msm_with(ctx, reinterpret_cast<const uint64_t*>(sp.h.xis.data()), (uint32_t)sp.h.xis.size() - 1,
                                  reinterpret_cast<const uint64_t*>(sp.aff.data()), sp.inf.data(),
                                  reinterpret_cast<const uint64_t*>(sp.scalars.data()), sp.aff.size(), out_h, out_s));
    PallasPoint lhs = sp.C_prime + point_load(out_s);
    ensure(xyzz_is_inf(lhs.p), HALO_REJECT_SUCCINCT, "C_(log_n) != CM.Commit_Sigma(c || v')");  // :307-310
    ensure(sp.U == point_load(out_h), HALO_REJECT_U, "U != CM.Commit(ck, h_vec)");                // :339
}

EvalProof proof_from_c(const halo_eval_proof& p) {
    EvalProof r;
    ensure(p.lg_n <= HALO_MAX_LG, HALO_EINVAL, "lg_n too large");
    for (uint32_t i = 0; i < p.lg_n; i++) {
        r.Ls.push_back(point_load(p.Ls[i]));
        r.Rs.push_back(point_load(p.Rs[i]));
    }
    r.U = point_load(p.U);
    r.c = scalar_load(p.c);
    r.hiding = p.hiding != 0;
    if (r.hiding) {
        r.C_bar = point_load(p.C_bar);
        r.w_prime = scalar_load(p.w_prime);
    }
    return r;
}
void proof_to_c(const EvalProof& p, halo_eval_proof& out) {
    std::memset(&out, 0, sizeof out);
    out.lg_n = (uint32_t)p.Ls.size();
    out.hiding = p.hiding ? 1 : 0;
    for (size_t i = 0; i < p.Ls.size(); i++) {
        point_store(out.Ls[i], p.Ls[i]);
        point_store(out.Rs[i], p.Rs[i]);
    }
    point_store(out.U, p.U);
    scalar_store(out.c, p.c);
    if (p.hiding) {
        point_store(out.C_bar, p.C_bar);
        scalar_store(out.w_prime, p.w_prime);
    }
}

}  // namespace pcdl
}  // namespace halo
