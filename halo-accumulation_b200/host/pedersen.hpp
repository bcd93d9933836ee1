// host/pedersen.hpp -- host mirror of code/src/pedersen.rs.
#pragma once
#include "group.hpp"

namespace halo {
namespace pedersen {

// pedersen.rs:6-20  commit(w, Gs, ms) = <ms, Gs> (+ w * S).
// Gs == nullptr selects the context's resident generators GS[0..n_gs) (what every reference caller passes:
// pcdl.rs:109, :338); otherwise caller-supplied affine points.  Trailing zero scalars contribute nothing, so a
// caller may pass n_ms < n_gs meaning "zero-padded to n_gs" (pcdl.rs:106-107) without uploading the zeros.
inline PallasPoint commit(halo_ctx* ctx, const PallasScalar* w, const uint64_t* Gs_affine, uint64_t n_gs,
                          const PallasScalar* ms, uint64_t n_ms, bool zero_padded = false) {
    if (!zero_padded) ensure(n_gs == n_ms, HALO_ELEN, "Length did not match for pedersen commitment");  // pedersen.rs:7-12
    uint64_t out[12];
    if (Gs_affine)
        check_rc(ctx, halo_msm(ctx, Gs_affine, nullptr, reinterpret_cast<const uint64_t*>(ms), n_ms, out));
    else
        check_rc(ctx, halo_msm_gens(ctx, reinterpret_cast<const uint64_t*>(ms), 0, n_ms, out));
    PallasPoint acc = point_load(out);
    if (w) {
        PallasPoint S, H;
        params_SH(ctx, S, H);
        acc = S * (*w) + acc;  // pedersen.rs:15-16
    }
    return acc;
}

}  // namespace pedersen
}  // namespace halo
