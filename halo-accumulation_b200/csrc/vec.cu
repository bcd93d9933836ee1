// vec.cu -- Fr vector kernels: dot products, powers, folds, and the h(X) expansion (K5).
//
// Replaces the serial scalar loops of the reference:
//   group.rs:13-15  scalar_dot            -> k_dot_partial + k_dot_final
//   group.rs:29-37  construct_powers      -> k_powers
//   pcdl.rs:221-223 c / z folds           -> k_fold_scalars   (HBM bound: 2 x (64 B read + 32 B write) per j)
//   pcdl.rs:56-77   HPoly::get_poly       -> k_h_expand       (closed form pinned by pcdl.rs:496-508:
//                                             coeff[j] = prod_{b: bit b of j set} xi_{lg n - b})
//   acc.rs:85-94    AccumulatedHPolys::get_poly -> k_h_expand with scale = alpha^{i+1}, accumulating
//   pcdl.rs:140-142,156  p_bar = q (X - z), p' = p + alpha p_bar -> k_pbar, k_axpy
#include "common.cuh"
#include "vec.cuh"

namespace halo {

constexpr int VEC_THREADS = 256;

// ---- powers ------------------------------------------------------------------------------------
struct PowTable {
    fr_t p[40];  // p[i] = z^(2^i)
};
constexpr int POW_CHUNK = 32;

__global__ void __launch_bounds__(VEC_THREADS) k_powers(PowTable tab, uint64_t n, fr_t* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t j0 = t * POW_CHUNK;
    if (j0 >= n) return;
    fr_t cur;
    fp_one(cur);
#pragma unroll 1
    for (int i = 0; i < 40; i++)
        if ((j0 >> i) & 1) fp_mul(cur, cur, tab.p[i]);
    fr_t z = tab.p[0];
#pragma unroll 1
    for (int k = 0; k < POW_CHUNK && j0 + k < n; k++) {
        out[j0 + k] = cur;
        fp_mul(cur, cur, z);
    }
}

void vec_powers(halo_ctx* ctx, const fr_t& z, uint64_t n, fr_t* d_out) {
    if (n == 0) return;
    PowTable tab;
    tab.p[0] = z;
    for (int i = 1; i < 40; i++) fp_sqr(tab.p[i], tab.p[i - 1]);
    uint64_t threads = (n + POW_CHUNK - 1) / POW_CHUNK;
    k_powers<<<(unsigned)((threads + VEC_THREADS - 1) / VEC_THREADS), VEC_THREADS, 0, ctx->stream>>>(tab, n, d_out);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

// ---- dot product ---------------------------------------------------------------------------------
__device__ __forceinline__ void block_sum_fr(fr_t& v, fr_t* sm) {
    const int j = threadIdx.x;
    sm[j] = v;
    __syncthreads();
    for (int stride = VEC_THREADS / 2; stride >= 1; stride >>= 1) {
        if (j < stride) {
            fp_add(v, v, sm[j + stride]);
            sm[j] = v;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(VEC_THREADS) k_dot_partial(const fr_t* __restrict__ a, const fr_t* __restrict__ b, uint64_t n,
                                                             fr_t* __restrict__ partials) {
    __shared__ fr_t sm[VEC_THREADS];
    fr_t acc, t;
    fp_zero(acc);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        fp_mul(t, a[i], b[i]);
        fp_add(acc, acc, t);
    }
    block_sum_fr(acc, sm);
    if (threadIdx.x == 0) partials[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(VEC_THREADS) k_dot_final(const fr_t* __restrict__ partials, uint32_t count, fr_t* __restrict__ out) {
    __shared__ fr_t sm[VEC_THREADS];
    fr_t acc;
    fp_zero(acc);
    for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) fp_add(acc, acc, partials[i]);
    block_sum_fr(acc, sm);
    if (threadIdx.x == 0) *out = acc;
}

void vec_dot(halo_ctx* ctx, const fr_t* d_a, const fr_t* d_b, uint64_t n, fr_t* d_partials, fr_t* d_out) {
    uint32_t blocks = (uint32_t)((n + VEC_THREADS - 1) / VEC_THREADS);
    if (blocks > VEC_DOT_MAX_BLOCKS) blocks = VEC_DOT_MAX_BLOCKS;
    if (blocks == 0) blocks = 1;
    k_dot_partial<<<blocks, VEC_THREADS, 0, ctx->stream>>>(d_a, d_b, n, d_partials);
    k_dot_final<<<1, VEC_THREADS, 0, ctx->stream>>>(d_partials, blocks, d_out);
    ctx->kernel_launches += 2;
    HALO_CUDA(cudaGetLastError());
}

// ---- scalar folds (pcdl.rs:221-223) ------------------------------------------------------------------
__global__ void __launch_bounds__(VEC_THREADS) k_fold_scalars(fr_t* __restrict__ c, fr_t* __restrict__ z, uint64_t m, fr_t xi,
                                                              fr_t xi_inv) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    fr_t lo, hi, t;
    lo = c[j];
    hi = c[j + m];
    fp_mul(t, hi, xi_inv);
    fp_add(lo, lo, t);
    c[j] = lo;
    lo = z[j];
    hi = z[j + m];
    fp_mul(t, hi, xi);
    fp_add(lo, lo, t);
    z[j] = lo;
}
void vec_fold_scalars(halo_ctx* ctx, fr_t* d_c, fr_t* d_z, uint64_t m, const fr_t& xi, const fr_t& xi_inv) {
    if (m == 0) return;
    k_fold_scalars<<<(unsigned)((m + VEC_THREADS - 1) / VEC_THREADS), VEC_THREADS, 0, ctx->stream>>>(d_c, d_z, m, xi, xi_inv);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

// ---- h(X) expansion (K5) -------------------------------------------------------------------------------
struct XiTable {
    fr_t x[32];  // x[b] = xi_{lg_n - b}: the factor contributed by bit b of the coefficient index
};

// Each thread produces 8 consecutive coefficients: one product for the high index bits, then the 8 subset
// products of the three low-bit factors (7 multiplications).
__global__ void __launch_bounds__(VEC_THREADS) k_h_expand(XiTable tab, int lg_n, uint64_t n, fr_t scale, int accumulate,
                                                          fr_t* __restrict__ out) {
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t j0 = t * 8;
    if (j0 >= n) return;
    fr_t v[8];
    v[0] = scale;
#pragma unroll 1
    for (int b = 3; b < lg_n; b++)
        if ((j0 >> b) & 1) fp_mul(v[0], v[0], tab.x[b]);
    fp_mul(v[1], v[0], tab.x[0]);
    fp_mul(v[2], v[0], tab.x[1]);
    fp_mul(v[3], v[2], tab.x[0]);
    fp_mul(v[4], v[0], tab.x[2]);
    fp_mul(v[5], v[4], tab.x[0]);
    fp_mul(v[6], v[4], tab.x[1]);
    fp_mul(v[7], v[6], tab.x[0]);
#pragma unroll
    for (int s = 0; s < 8; s++) {
        if (j0 + s < n) {
            if (accumulate) {
                fr_t o = out[j0 + s];
                fp_add(o, o, v[s]);
                out[j0 + s] = o;
            } else {
                out[j0 + s] = v[s];
            }
        }
    }
}

void vec_h_expand(halo_ctx* ctx, const fr_t* xis /*host, lg_n + 1*/, int lg_n, const fr_t& scale, bool accumulate,
                  fr_t* d_out) {
    XiTable tab;
    for (int b = 0; b < 32; b++) {
        if (b < lg_n)
            tab.x[b] = xis[lg_n - b];
        else
            fp_one(tab.x[b]);
    }
    uint64_t n = (uint64_t)1 << lg_n;
    uint64_t threads = (n + 7) / 8;
    k_h_expand<<<(unsigned)((threads + VEC_THREADS - 1) / VEC_THREADS), VEC_THREADS, 0, ctx->stream>>>(tab, lg_n, n, scale,
                                                                                                     accumulate ? 1 : 0, d_out);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

// ---- hiding polynomial helpers ---------------------------------------------------------------------------
// p_bar = q * (X - z), zero-padded to n: p_bar[i] = q[i-1] - z q[i]   (pcdl.rs:140-142)
__global__ void __launch_bounds__(VEC_THREADS) k_pbar(const fr_t* __restrict__ q, uint64_t n_q, fr_t z, uint64_t n,
                                                      fr_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t acc, t;
    fp_zero(acc);
    if (i >= 1 && i - 1 < n_q) acc = q[i - 1];
    if (i < n_q) {
        fp_mul(t, z, q[i]);
        fp_sub(acc, acc, t);
    }
    out[i] = acc;
}
void vec_pbar(halo_ctx* ctx, const fr_t* d_q, uint64_t n_q, const fr_t& z, uint64_t n, fr_t* d_out) {
    k_pbar<<<(unsigned)((n + VEC_THREADS - 1) / VEC_THREADS), VEC_THREADS, 0, ctx->stream>>>(d_q, n_q, z, n, d_out);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

// y += alpha * x   (pcdl.rs:156)
__global__ void __launch_bounds__(VEC_THREADS) k_axpy(fr_t* __restrict__ y, const fr_t* __restrict__ x, fr_t alpha, uint64_t n) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t t, o = y[i];
    fp_mul(t, x[i], alpha);
    fp_add(o, o, t);
    y[i] = o;
}
void vec_axpy(halo_ctx* ctx, fr_t* d_y, const fr_t* d_x, const fr_t& alpha, uint64_t n) {
    if (n == 0) return;
    k_axpy<<<(unsigned)((n + VEC_THREADS - 1) / VEC_THREADS), VEC_THREADS, 0, ctx->stream>>>(d_y, d_x, alpha, n);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

}  // namespace halo
