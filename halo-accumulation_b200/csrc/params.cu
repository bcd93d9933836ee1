// params.cu -- K6: public-parameter derivation on device.
//
// Restates main.rs:18-45: P_k = [SHA3-256(genesis || k as u64 LE) mod r] * (-1, 2); S = P_0, H = P_1,
// GS[i] = P_{i+2}.  The reference does this offline and pastes 16 384 points into consts.rs
// (limitation at report/report.md:2081-2086); here a fixed-base table (32 windows x 255 multiples of the
// generator) turns each derivation into <= 32 mixed adds + one inversion, so n = 2^24 takes a fraction
// of a second and every GPU can derive its own slice.
#include "common.cuh"
#include "params.cuh"
#include "sha3.cuh"

namespace halo {

__constant__ uint8_t c_genesis[60];
static const char GENESIS[] = "To understand recursion, one must first understand recursion";
static_assert(sizeof(GENESIS) - 1 == 60, "genesis string length");

constexpr int FT_WINDOWS = 32;  // 8-bit windows over a 256-bit scalar
constexpr int FT_ENTRIES = 255;

// table[w][d-1] = d * 2^(8w) * (-1, 2), affine
__global__ void __launch_bounds__(64) k_fixed_table(affine_t* __restrict__ table) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= FT_WINDOWS * FT_ENTRIES) return;
    int w = idx / FT_ENTRIES, d = idx % FT_ENTRIES + 1;
    uint32_t k[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    k[w >> 2] = (uint32_t)d << (8 * (w & 3));
    xyzz_t g, r;
    fq_t one;
    fp_one(one);
    fp_neg(g.x, one);       // x = -1
    fp_dbl(g.y, one);       // y = 2
    fp_one(g.zz);
    fp_one(g.zzz);
    xyzz_mul_canon(r, g, k);
    affine_t a;
    xyzz_to_affine(a, r);
    table[idx] = a;
}

__device__ __noinline__ void xyzz_madd_nl(xyzz_t& acc, const affine_t& q) { xyzz_madd(acc, q, false); }

__global__ void __launch_bounds__(128) k_derive_points(const affine_t* __restrict__ table, uint64_t start, uint64_t count,
                                                       affine_t* __restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint64_t kidx = start + i;
    uint8_t msg[68];
#pragma unroll
    for (int j = 0; j < 60; j++) msg[j] = c_genesis[j];
#pragma unroll
    for (int j = 0; j < 8; j++) msg[60 + j] = (uint8_t)(kidx >> (8 * j));  // usize::to_le_bytes (main.rs:24)
    uint64_t dg[4];
    sha3_256(msg, 68, dg);
    // from_le_bytes_mod_order (main.rs:28): the 256-bit little-endian digest reduced mod r (r > 2^254: <= 3 subtractions)
    uint32_t s[8], m[8], t[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        s[2 * j] = (uint32_t)dg[j];
        s[2 * j + 1] = (uint32_t)(dg[j] >> 32);
    }
    fp_mod_limbs<ScalarParams>(m);
    for (int it = 0; it < 3; it++) {
        uint32_t borrow = sub8(t, s, m);
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = borrow ? s[j] : t[j];
    }
    xyzz_t acc;
    xyzz_set_inf(acc);
    for (int w = 0; w < FT_WINDOWS; w++) {
        uint32_t d = (s[w >> 2] >> (8 * (w & 3))) & 0xffu;
        if (d) {
            affine_t p = table[w * FT_ENTRIES + (int)d - 1];
            xyzz_madd_nl(acc, p);
        }
    }
    affine_t a;
    xyzz_to_affine(a, acc);
    out[i] = a;
}

// Generator store (capi.cu: halo_load_generators_file): a record is accepted only if it is a canonical Montgomery residue
// pair on y^2 = x^3 + 5 (the point at infinity is not a generator).  bad[0] counts the records that are not.
__global__ void __launch_bounds__(256) k_count_off_curve(const affine_t* __restrict__ pts, uint64_t n, unsigned long long* bad) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p = pts[i];
    uint32_t m[8], t[8];
    fp_mod_limbs<BaseParams>(m);
    bool ok = sub8(t, p.x.v, m) != 0 && sub8(t, p.y.v, m) != 0;  // both coordinates < modulus
    fq_t lhs, rhs, five;
    fp_sqr(lhs, p.y);
    fp_sqr(rhs, p.x);
    fp_mul(rhs, rhs, p.x);
    fp_from_u32(five, 5);
    fp_add(rhs, rhs, five);
    ok = ok && fp_eq(lhs, rhs);
    if (!ok) atomicAdd(bad, 1ull);
}

uint64_t params_count_off_curve(halo_ctx* ctx, const affine_t* d_pts, uint64_t n) {
    if (n == 0) return 0;
    ctx->stage_misc.reserve(256);
    auto* d_bad = ctx->stage_misc.as<unsigned long long>();
    HALO_CUDA(cudaMemsetAsync(d_bad, 0, 8, ctx->stream));
    k_count_off_curve<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_pts, n, d_bad);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
    unsigned long long bad = 0;
    HALO_CUDA(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    return bad;
}

void params_ensure_table(halo_ctx* ctx) {
    if (ctx->fixed_table.p) return;
    HALO_CUDA(cudaMemcpyToSymbolAsync(c_genesis, GENESIS, 60, 0, cudaMemcpyHostToDevice, ctx->stream));
    ctx->fixed_table.reserve((size_t)FT_WINDOWS * FT_ENTRIES * sizeof(affine_t));
    int total = FT_WINDOWS * FT_ENTRIES;
    k_fixed_table<<<(total + 63) / 64, 64, 0, ctx->stream>>>(ctx->fixed_table.as<affine_t>());
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

void params_derive_points(halo_ctx* ctx, uint64_t start, uint64_t count, affine_t* d_out) {
    if (count == 0) return;
    params_ensure_table(ctx);
    uint64_t grid = (count + 127) / 128;
    k_derive_points<<<(unsigned)grid, 128, 0, ctx->stream>>>(ctx->fixed_table.as<affine_t>(), start, count, d_out);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
}

}  // namespace halo
