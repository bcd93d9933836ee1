// host/pcdl.hpp -- host mirror of code/src/pcdl.rs (Bulletproofs-style polynomial commitment, DL based).
#pragma once
#include "../../include/halo_pcdl.h"
#include "group.hpp"

namespace halo {
namespace pcdl {

// pcdl.rs:22-30
struct EvalProof {
    std::vector<PallasPoint> Ls, Rs;
    PallasPoint U;
    PallasScalar c;
    bool hiding = false;  // C_bar / w_prime are Some(..)
    PallasPoint C_bar;
    PallasScalar w_prime;
};

// pcdl.rs:44-92
struct HPoly {
    std::vector<PallasScalar> xis;
    explicit HPoly(std::vector<PallasScalar> x) : xis(std::move(x)) {}
    // pcdl.rs:56-77: coefficient vector (n = 2^lg n elements), expanded on the device
    PallasPoly get_poly(halo_ctx* ctx) const;
    // pcdl.rs:79-91
    PallasScalar eval(const PallasScalar& z) const;
};

// pcdl.rs:99-110
PallasPoint commit(halo_ctx* ctx, PolyView p, uint64_t d, const PallasScalar* w);
// pcdl.rs:120-242; rng draws made explicit: q (deg p coefficients), w_bar
EvalProof open(halo_ctx* ctx, PolyView p, const PallasPoint& C, uint64_t d, const PallasScalar& z, const PallasScalar* w,
               const PolyView* q, const PallasScalar* w_bar);
// Same for a polynomial of degree `deg` that already sits on the device (halo_h_lincomb_resident): acc.rs:209.
EvalProof open_resident(halo_ctx* ctx, uint64_t deg, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                        const PallasScalar* w, const PolyView* q, const PallasScalar* w_bar);
// pcdl.rs:252-314; throws HaloFailure(HALO_REJECT_SUCCINCT) on reject
std::pair<HPoly, PallasPoint> succinct_check(halo_ctx* ctx, const PallasPoint& C, uint64_t d, const PallasScalar& z,
                                             const PallasScalar& v, const EvalProof& pi);
// the succinct checks of several instances (acc.rs:158-170) and one unhidden commitment (acc.rs:153) with a
// single device round trip (halo_msm_multi).  Nothing is decided here: `accept[i]` is instance i's group equation
// (pcdl.rs:307-310) and the caller raises the reference's `ensure!`s in the reference's order.  Malformed input throws
// exactly as succinct_check / commit would, but possibly out of order -- callers fall back to the one-by-one path then.
struct Query {
    const PallasPoint* C;
    uint64_t d;
    const PallasScalar* z;
    const PallasScalar* v;
    const EvalProof* pi;
};
struct SuccinctMany {
    std::vector<std::pair<HPoly, PallasPoint>> hu;  // what succinct_check returns per instance
    std::vector<bool> accept;
    PallasPoint commitment;                          // commit(commit_p, commit_d, None) when commit_p != nullptr
};
SuccinctMany succinct_check_many(halo_ctx* ctx, const std::vector<Query>& qs, const PolyView* commit_p, uint64_t commit_d);
// pcdl.rs:323-342
void check(halo_ctx* ctx, const PallasPoint& C, uint64_t d, const PallasScalar& z, const PallasScalar& v,
           const EvalProof& pi);

// wire conversion
EvalProof proof_from_c(const halo_eval_proof& p);
void proof_to_c(const EvalProof& p, halo_eval_proof& out);

}  // namespace pcdl
}  // namespace halo
