// msm.cuh -- internal interface of the MSM engine (msm.cu).
#pragma once
#include "common.cuh"

namespace halo {
// Window sums S_w (XYZZ, device) for sum_i scalars[i] * bases[i]; all buffers device resident.
void msm_window_sums(halo_ctx* ctx, const affine_t* d_bases, const fr_t* d_scalars, uint32_t n, const MsmPlan& plan,
                     xyzz_t* d_wsums_out);
void msm_finish_host(const xyzz_t* wsums, const MsmPlan& plan, xyzz_t& out);
// Full MSM with device-resident inputs; synchronises the context stream and returns the point on the host.
void msm_device(halo_ctx* ctx, const affine_t* d_bases, const fr_t* d_scalars, uint64_t n, xyzz_t& out);
}  // namespace halo
