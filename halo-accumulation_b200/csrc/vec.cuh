// vec.cuh -- internal interface of vec.cu (Fr vector kernels and the h(X) expansion).
#pragma once
#include "common.cuh"
namespace halo {
constexpr uint32_t VEC_DOT_MAX_BLOCKS = 1184;  // 8 CTAs per SM x 148 SMs
// d_out[j] = z^j, j < n
void vec_powers(halo_ctx* ctx, const fr_t& z, uint64_t n, fr_t* d_out);
// *d_out = sum a[i] b[i]; d_partials: scratch of VEC_DOT_MAX_BLOCKS elements; everything device resident
void vec_dot(halo_ctx* ctx, const fr_t* d_a, const fr_t* d_b, uint64_t n, fr_t* d_partials, fr_t* d_out);
// c[j] += xi_inv c[j+m]; z[j] += xi z[j+m], j < m
void vec_fold_scalars(halo_ctx* ctx, fr_t* d_c, fr_t* d_z, uint64_t m, const fr_t& xi, const fr_t& xi_inv);
// d_out[j] (+)= scale * prod_{b: bit b of j} xis[lg_n - b], j < 2^lg_n   (xis on the host)
void vec_h_expand(halo_ctx* ctx, const fr_t* xis, int lg_n, const fr_t& scale, bool accumulate, fr_t* d_out);
void vec_pbar(halo_ctx* ctx, const fr_t* d_q, uint64_t n_q, const fr_t& z, uint64_t n, fr_t* d_out);
void vec_axpy(halo_ctx* ctx, fr_t* d_y, const fr_t* d_x, const fr_t& alpha, uint64_t n);
}  // namespace halo
