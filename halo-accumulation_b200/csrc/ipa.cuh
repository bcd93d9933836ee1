// ipa.cuh -- device state of one PCDL opening (pcdl.rs:183-231).
#pragma once
#include <vector>

#include "common.cuh"

struct halo_ipa {
    halo_ctx* ctx = nullptr;
    uint64_t n = 0;    // d + 1
    uint64_t cur = 0;  // current vector length (n, n/2, ..., 1)
    uint32_t lg_n = 0;
    uint32_t round = 0;
    bool have_hprime = false;
    bool lr_done = false;
    halo::fr_t z;
    // device buffers live in the context (ctx->ipa_*): G = affine working copy of GS[0..n) (pcdl.rs:185), cs / zs =
    // coefficient and z-power vectors (pcdl.rs:183-186), pbar = hiding polynomial (pcdl.rs:140-142),
    // tail = [affine H'] then [fr dot_l, fr dot_r] and dot partials
    bool have_pbar = false;
    bool frozen = false;  // generator vector frozen at M0 elements (ctx->ipa_frozen holds s, tL, tR)
    uint32_t M0 = 0;
    // deferred head: the first `defer` rounds leave the generators untouched (L / R over GS itself with per-index
    // coefficients, fixed-base tables when present); k_fold_multi then materialises G^(defer) in one joint pass
    // Later stages repeat the trick on the materialised vector (ctx->ipa_G, variable base): while the vector is longer than
    // the frozen-tail length, the next `stage_D` rounds run as per-index-coefficient MSMs over it and ONE k_fold_multi takes
    // it down 2^stage_D-fold (2^20: rounds 3-6 over G^(3), then G^(7) of 8192 elements, which the frozen tail keeps).
    int defer = 0;
    int stage_D = 0;             // rounds of the current deferred stage
    uint32_t stage_first = 0;    // its first round
    bool stage_over_gens = false;  // head stage: the vector is still GS[0..n) (FIXED-base tables usable)
    bool deferred = false, fixed_ok = false;
    halo::affine_t hprime;  // affine H' on the host (the FIXED-base L / R add dot * H' there)
    std::vector<halo::fr_t> defer_xis;
};

namespace halo {
// k_fold_multi of the second translation unit (ipa_fold.cu: field multiplication out of line), for folds with many outputs.
void launch_fold_multi_call(cudaStream_t st, unsigned grid, const affine_t* G0, uint64_t n, int D, const fq_t* bx, const affine_t* diff,
                            const uint8_t* ops, int n_ops, xyzz_t* sums, fq_t* den);
}  // namespace halo
