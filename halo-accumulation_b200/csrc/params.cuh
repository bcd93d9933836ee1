// params.cuh -- internal interface of params.cu (K6 generator derivation).
#pragma once
#include "common.cuh"
namespace halo {
void params_ensure_table(halo_ctx* ctx);
// d_out[i] = P_{start+i} (Montgomery affine), i < count; asynchronous on ctx->stream
void params_derive_points(halo_ctx* ctx, uint64_t start, uint64_t count, affine_t* d_out);
// number of records among d_pts[0..n) that are not canonical affine points on the curve (synchronises ctx->stream)
uint64_t params_count_off_curve(halo_ctx* ctx, const affine_t* d_pts, uint64_t n);
}  // namespace halo
