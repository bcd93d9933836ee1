// host/capi_host.cpp -- C view (include/halo_pcdl.h) of the C++ host layer, for ctypes / C callers.
#include "../../include/halo_pcdl.h"

#include "acc.hpp"
#include "pcdl.hpp"
#include "pedersen.hpp"

using namespace halo;

namespace {
thread_local std::string g_host_error;
template <class F>
int guarded(F&& f) {
    try {
        f();
        return HALO_OK;
    } catch (const HaloFailure& e) {
        g_host_error = e.what();
        return e.code;
    } catch (const std::bad_alloc&) {
        g_host_error = "out of host memory";
        return HALO_ENOMEM;
    } catch (const std::exception& e) {
        g_host_error = e.what();
        return HALO_EINVAL;
    } catch (...) {  // nothing unwinds across the C ABI
        g_host_error = "unexpected internal exception";
        return HALO_EINVAL;
    }
}
// Verifier-side inputs are raw limbs from the caller: canonical residues and points on the curve, or HALO_EINVAL
// (see host/group.hpp "validation at the C boundary").
void validate(const halo_eval_proof& pi) {
    ensure(pi.lg_n <= HALO_MAX_LG, HALO_EINVAL, "proof: lg_n out of range");
    for (uint32_t i = 0; i < pi.lg_n; i++)
        ensure(point_valid(pi.Ls[i]) && point_valid(pi.Rs[i]), HALO_EINVAL, "proof: L / R is not a canonical point on the curve");
    ensure(point_valid(pi.U), HALO_EINVAL, "proof: U is not a canonical point on the curve");
    ensure(scalar_valid(pi.c), HALO_EINVAL, "proof: c is not a canonical scalar");
    if (pi.hiding) {
        ensure(point_valid(pi.C_bar), HALO_EINVAL, "proof: C_bar is not a canonical point on the curve");
        ensure(scalar_valid(pi.w_prime), HALO_EINVAL, "proof: w' is not a canonical scalar");
    }
}
void validate(const halo_instance& q) {
    ensure(point_valid(q.C), HALO_EINVAL, "instance: C is not a canonical point on the curve");
    ensure(scalar_valid(q.z) && scalar_valid(q.v), HALO_EINVAL, "instance: z / v is not a canonical scalar");
    validate(q.pi);
}
void validate(const halo_accumulator& a) {
    ensure(point_valid(a.C_bar) && point_valid(a.U0), HALO_EINVAL, "accumulator: C_bar / U_0 is not a canonical point on the curve");
    ensure(scalar_valid(a.z) && scalar_valid(a.v) && scalar_valid(a.w) && scalar_valid(a.h0[0]) && scalar_valid(a.h0[1]), HALO_EINVAL,
           "accumulator: z / v / w / h_0 is not a canonical scalar");
    validate(a.pi);
}
PolyView poly_from(const uint64_t* c, uint64_t n) { return PolyView(reinterpret_cast<const PallasScalar*>(c), n); }
}  // namespace

extern "C" {

const char* halo_host_last_error(void) { return g_host_error.c_str(); }

int halo_pedersen_commit(halo_ctx* ctx, const uint64_t* w, const uint64_t* gs_affine, uint64_t n_gs, const uint64_t* ms,
                         uint64_t n_ms, uint64_t out_jac[12]) {
    return guarded([&] {
        PallasScalar ws;
        if (w) ws = scalar_load(w);
        PallasPoint r = pedersen::commit(ctx, w ? &ws : nullptr, gs_affine, n_gs, reinterpret_cast<const PallasScalar*>(ms), n_ms);
        point_store(out_jac, r);
    });
}

int halo_pcdl_commit(halo_ctx* ctx, const uint64_t* coeffs, uint64_t n_coeffs, uint64_t d, const uint64_t* w, uint64_t out_jac[12]) {
    return guarded([&] {
        PallasScalar ws;
        if (w) ws = scalar_load(w);
        point_store(out_jac, pcdl::commit(ctx, poly_from(coeffs, n_coeffs), d, w ? &ws : nullptr));
    });
}

int halo_pcdl_open(halo_ctx* ctx, const uint64_t* coeffs, uint64_t n_coeffs, const uint64_t C_jac[12], uint64_t d,
                   const uint64_t z[4], const uint64_t* w, const uint64_t* q, uint64_t n_q, const uint64_t* w_bar,
                   halo_eval_proof* pi) {
    return guarded([&] {
        PallasScalar ws, wbs;
        PolyView qp(nullptr, 0);
        if (w) {
            ensure(q && w_bar, HALO_EINVAL, "hiding open needs q and w_bar");
            ws = scalar_load(w);
            wbs = scalar_load(w_bar);
            qp = poly_from(q, n_q);
        }
        pcdl::EvalProof p = pcdl::open(ctx, poly_from(coeffs, n_coeffs), point_load(C_jac), d, scalar_load(z), w ? &ws : nullptr,
                                       w ? &qp : nullptr, w ? &wbs : nullptr);
        pcdl::proof_to_c(p, *pi);
    });
}

int halo_pcdl_succinct_check(halo_ctx* ctx, const uint64_t C_jac[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                             const halo_eval_proof* pi, uint64_t* xis_out, uint64_t U_out[12]) {
    return guarded([&] {
        ensure(pi != nullptr, HALO_EINVAL, "null proof");
        validate(*pi);
        auto hu = pcdl::succinct_check(ctx, point_load_checked(C_jac, "C is not a canonical point on the curve"), d,
                                       scalar_load_checked(z, "z is not a canonical scalar"),
                                       scalar_load_checked(v, "v is not a canonical scalar"), pcdl::proof_from_c(*pi));
        if (xis_out) std::memcpy(xis_out, hu.first.xis.data(), hu.first.xis.size() * 32);
        if (U_out) point_store(U_out, hu.second);
    });
}

int halo_pcdl_check(halo_ctx* ctx, const uint64_t C_jac[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                    const halo_eval_proof* pi) {
    return guarded([&] {
        ensure(pi != nullptr, HALO_EINVAL, "null proof");
        validate(*pi);
        pcdl::check(ctx, point_load_checked(C_jac, "C is not a canonical point on the curve"), d,
                    scalar_load_checked(z, "z is not a canonical scalar"), scalar_load_checked(v, "v is not a canonical scalar"),
                    pcdl::proof_from_c(*pi));
    });
}

int halo_h_eval(const uint64_t* xis, uint32_t lg_n, const uint64_t z[4], uint64_t out[4]) {
    return guarded([&] {
        std::vector<PallasScalar> x(lg_n + 1);
        std::memcpy(x.data(), xis, (lg_n + 1) * 32);
        scalar_store(out, pcdl::HPoly(x).eval(scalar_load(z)));
    });
}

int halo_acc_prover(halo_ctx* ctx, uint64_t d, const halo_instance* qs, uint64_t m, const uint64_t h0[2][4], const uint64_t w[4],
                    const uint64_t* q, uint64_t n_q, const uint64_t w_bar[4], halo_accumulator* acc) {
    return guarded([&] {
        ensure(qs != nullptr || m == 0, HALO_EINVAL, "null instance list");
        std::vector<acc::Instance> v;
        for (uint64_t i = 0; i < m; i++) {
            validate(qs[i]);
            v.push_back(acc::instance_from_c(qs[i]));
        }
        acc::Accumulator a = acc::prover(ctx, d, v, PallasPoly{scalar_load(h0[0]), scalar_load(h0[1])}, scalar_load(w),
                                         poly_from(q, n_q), scalar_load(w_bar));
        acc::accumulator_to_c(a, *acc);
    });
}

int halo_acc_verifier(halo_ctx* ctx, uint64_t d, const halo_instance* qs, uint64_t m, const halo_accumulator* acc) {
    return guarded([&] {
        ensure(acc != nullptr && (qs != nullptr || m == 0), HALO_EINVAL, "null accumulator / instance list");
        validate(*acc);
        std::vector<acc::Instance> v;
        for (uint64_t i = 0; i < m; i++) {
            validate(qs[i]);
            v.push_back(acc::instance_from_c(qs[i]));
        }
        acc::verifier(ctx, d, v, acc::accumulator_from_c(*acc));
    });
}

int halo_acc_decider(halo_ctx* ctx, const halo_accumulator* acc) {
    return guarded([&] {
        ensure(acc != nullptr, HALO_EINVAL, "null accumulator");
        validate(*acc);
        acc::decider(ctx, acc::accumulator_from_c(*acc));
    });
}

void halo_acc_to_instance(const halo_accumulator* acc, halo_instance* q) {
    std::memcpy(q->C, acc->C_bar, 96);
    q->d = acc->d;
    std::memcpy(q->z, acc->z, 32);
    std::memcpy(q->v, acc->v, 32);
    q->pi = acc->pi;
}

void halo_point_serialize_compressed(const uint64_t p_jac[12], uint8_t out[33]) { serialize_compressed(point_load(p_jac), out); }

}  // extern "C"
