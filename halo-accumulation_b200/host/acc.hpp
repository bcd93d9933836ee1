// host/acc.hpp -- host mirror of code/src/acc.rs (ASDL accumulation scheme).
#pragma once
#include "pcdl.hpp"

namespace halo {
namespace acc {

// acc.rs:21-28
struct Instance {
    PallasPoint C;
    uint64_t d;
    PallasScalar z, v;
    pcdl::EvalProof pi;
};
// acc.rs:54-59
struct AccumulatorHiding {
    PallasPoly h;  // degree-1 polynomial h_0
    PallasPoint U;
    PallasScalar w;
};
// acc.rs:43-51
struct Accumulator {
    PallasPoint C_bar;
    uint64_t d;
    PallasScalar z, v;
    pcdl::EvalProof pi;
    AccumulatorHiding pi_V;
};
// impl From<Accumulator> for Instance (acc.rs:121-131)
inline Instance to_instance(const Accumulator& a) { return Instance{a.C_bar, a.d, a.z, a.v, a.pi}; }

// acc.rs:61-107
struct AccumulatedHPolys {
    bool have_h0 = false;
    PallasPoly h_0;
    std::vector<pcdl::HPoly> hs;
    bool have_alpha = false;
    PallasScalar alpha;
    std::vector<PallasScalar> alphas;
    size_t alphas_capacity = 0;
    explicit AccumulatedHPolys(size_t capacity) : alphas_capacity(capacity + 1) { hs.reserve(capacity); }  // :69-76
    void set_alpha(const PallasScalar& a) {                                                                   // :79-82
        alphas = construct_powers(a, alphas_capacity);
        alpha = a;
        have_alpha = true;
    }
    PallasPoly get_poly(halo_ctx* ctx, uint32_t lg_n) const;  // :85-94, on the device
    uint64_t get_poly_resident(halo_ctx* ctx, uint32_t lg_n) const;  // same, left on the device; returns the degree
    PallasScalar eval(const PallasScalar& z) const;           // :97-106
    void serialize(Transcript& t) const;                      // #[derive(CanonicalSerialize)] (:61)
};

// acc.rs:190-220; rng draws explicit in the reference's order: h_0 (2 coefficients, :192), w (:198), open's q, w_bar
Accumulator prover(halo_ctx* ctx, uint64_t d, const std::vector<Instance>& qs, const PallasPoly& h_0, const PallasScalar& w,
                   PolyView q, const PallasScalar& w_bar);
// acc.rs:223-243
void verifier(halo_ctx* ctx, uint64_t D, const std::vector<Instance>& qs, const Accumulator& acc);
// acc.rs:245-255
void decider(halo_ctx* ctx, const Accumulator& acc);

Instance instance_from_c(const halo_instance& q);
Accumulator accumulator_from_c(const halo_accumulator& a);
void accumulator_to_c(const Accumulator& a, halo_accumulator& out);

}  // namespace acc
}  // namespace halo
