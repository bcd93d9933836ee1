// msm.cuh -- internal interface of the MSM engine (msm.cu).
#pragma once
#include "common.cuh"

namespace halo {
// One MSM problem, all pointers device resident: sum_{i<n} scalars[i] * bases[i] plus an optional short "tail"
// of extra (base, scalar) pairs held elsewhere in memory (used for the H' term of the IPA rounds, pcdl.rs:204,208).
struct MsmInput {
    const affine_t* bases = nullptr;
    const fr_t* scalars = nullptr;
    uint32_t n = 0;
    const affine_t* tail_bases = nullptr;
    const fr_t* tail_scalars = nullptr;
    uint32_t n_tail = 0;
};
// Window sums S_w (XYZZ, device).
void msm_window_sums(halo_ctx* ctx, const MsmInput& in, const MsmPlan& plan, xyzz_t* d_wsums_out);
// Up to 4 MSMs enqueued back to back with a single synchronisation; results on the host.
void msm_batch(halo_ctx* ctx, const MsmInput* ins, int count, xyzz_t* outs);
void msm_finish_host(const xyzz_t* wsums, const MsmPlan& plan, xyzz_t& out);
// Full MSM with device-resident inputs; synchronises the context stream and returns the point on the host.
void msm_device(halo_ctx* ctx, const affine_t* d_bases, const fr_t* d_scalars, uint64_t n, xyzz_t& out);
}  // namespace halo
