// ipa.cu -- K3/K4/K7: the IPA rounds of PCDL.open on the device.
//
// Replaces the loop body pcdl.rs:195-227.  Per round, with m = cur / 2:
//   K3  dot_l = <c_hi, z_lo>, dot_r = <c_lo, z_hi>                       (group.rs:13-15, pcdl.rs:203,207)
//       L = <c_hi, G_lo> + dot_l H',  R = <c_lo, G_hi> + dot_r H'        (pcdl.rs:204,208) -- two Pippenger MSMs
//       whose H' term rides along as a one-element tail, enqueued back to back with one synchronisation
//   K4  G_j <- G_j + xi G_{j+m}   (shared-scalar multiply-add, signed NAF digits of xi broadcast to all
//       threads -> divergence-free), c_j <- c_j + xi^-1 c_{j+m}, z_j <- z_j + xi z_{j+m}   (pcdl.rs:216-224)
//   K7  the folded generators are normalised back to affine inside the fold kernel so every later MSM keeps the
//       8M+2S mixed addition (the reference instead pays one inversion per element per MSM, group.rs:19).
// Folds and MSM rounds stay on one GPU by design (sequential rounds with a hash between them, pcdl.rs:212).
#include "ipa.cuh"

#include <cstring>

#include "../../include/halo_b200.h"
#include "msm.cuh"
#include "vec.cuh"

using namespace halo;

namespace halo {

struct NafDigits {
    int8_t d[257];  // signed NAF digits of the canonical challenge, LSB first
    int16_t top;    // index of the most significant non-zero digit (-1 if xi == 0)
};

// K4 point part: G[j] <- affine(G[j] + xi * G[j + m]).  One thread per output element; the digit loop is
// uniform across the grid (same xi for every element of a round).
__global__ void __launch_bounds__(128) k_fold_points(affine_t* __restrict__ G, uint64_t m, NafDigits naf) {
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    affine_t hi = G[j + m];
    xyzz_t acc;
    xyzz_set_inf(acc);
#pragma unroll 1
    for (int i = naf.top; i >= 0; i--) {
        xyzz_dbl(acc, acc);
        int d = naf.d[i];
        if (d != 0) xyzz_madd(acc, hi, d < 0);
    }
    affine_t lo = G[j];
    xyzz_madd(acc, lo, false);
    affine_t out;
    xyzz_to_affine(out, acc);
    G[j] = out;
}

static void make_naf(const fr_t& xi, NafDigits& naf) {
    uint32_t k[9];
    fp_to_canon(k, xi);
    k[8] = 0;
    naf.top = -1;
    for (int i = 0; i < 257; i++) {
        int d = 0;
        if (k[0] & 1u) {
            d = 2 - (int)(k[0] & 3u);  // +1 or -1 so that (k - d) is divisible by 4
            if (d < 0) {
                // k += 1
                for (int l = 0; l < 9; l++)
                    if (++k[l] != 0) break;
            } else {
                k[0] -= 1;  // odd, no borrow
            }
            naf.top = (int16_t)i;
        }
        naf.d[i] = (int8_t)d;
        for (int l = 0; l < 8; l++) k[l] = (k[l] >> 1) | (k[l + 1] << 31);
        k[8] >>= 1;
    }
}

}  // namespace halo

#define IPA_TRY(st)  \
    halo_ctx* ctx = (st)->ctx; \
    try {             \
        HALO_CUDA(cudaSetDevice(ctx->device));
#define IPA_CATCH                                           \
    }                                                       \
    catch (const halo::CudaError& e) {                      \
        ctx->last_error = std::string("CUDA error: ") + cudaGetErrorString(e.err) + " (" + e.what + ")"; \
        return HALO_ECUDA;                                  \
    }                                                       \
    return HALO_OK;

static inline bool is_pow2(uint64_t n) { return n && !(n & (n - 1)); }

extern "C" {

int halo_ipa_begin(halo_ctx* ctx, const uint64_t* coeffs, uint64_t n_coeffs, uint64_t n, const uint64_t z[4], halo_ipa** out,
                   uint64_t v_out[4]) {
    if (!ctx || !out || !z || (!coeffs && n_coeffs)) return HALO_EINVAL;
    *out = nullptr;
    if (!is_pow2(n) || n > ctx->n_gens || n_coeffs > n) {  // pcdl.rs:128-132
        ctx->last_error = "halo_ipa_begin: n must be a power of two <= resident generators and >= n_coeffs";
        return HALO_EINVAL;
    }
    halo_ipa* st = new halo_ipa();
    st->ctx = ctx;
    try {
        HALO_CUDA(cudaSetDevice(ctx->device));
        st->n = st->cur = n;
        while (((uint64_t)1 << st->lg_n) < n) st->lg_n++;
        memcpy(&st->z, z, 32);
        st->G.reserve(n * sizeof(affine_t));
        st->cs.reserve(n * sizeof(fr_t));
        st->zs.reserve(n * sizeof(fr_t));
        st->tail.reserve(sizeof(affine_t) + (2 + VEC_DOT_MAX_BLOCKS) * sizeof(fr_t));
        cudaStream_t s = ctx->stream;
        HALO_CUDA(cudaMemcpyAsync(st->G.p, ctx->gens.p, n * sizeof(affine_t), cudaMemcpyDeviceToDevice, s));  // pcdl.rs:185
        HALO_CUDA(cudaMemsetAsync(st->cs.p, 0, n * sizeof(fr_t), s));                                         // pcdl.rs:183-184
        if (n_coeffs) HALO_CUDA(cudaMemcpyAsync(st->cs.p, coeffs, n_coeffs * sizeof(fr_t), cudaMemcpyHostToDevice, s));
        vec_powers(ctx, st->z, n, st->zs.as<fr_t>());  // pcdl.rs:186
        if (v_out) {                                   // v = p(z) = <c, z-powers>  (pcdl.rs:135)
            fr_t* scal = reinterpret_cast<fr_t*>(reinterpret_cast<char*>(st->tail.p) + sizeof(affine_t));
            vec_dot(ctx, st->cs.as<fr_t>(), st->zs.as<fr_t>(), n, scal + 2, scal);
            HALO_CUDA(cudaMemcpyAsync(v_out, scal, 32, cudaMemcpyDeviceToHost, s));
        }
        HALO_CUDA(cudaStreamSynchronize(s));
    } catch (const halo::CudaError& e) {
        ctx->last_error = std::string("CUDA error: ") + cudaGetErrorString(e.err) + " (" + e.what + ")";
        delete st;
        return HALO_ECUDA;
    }
    *out = st;
    return HALO_OK;
}

void halo_ipa_destroy(halo_ipa* st) {
    if (!st) return;
    cudaSetDevice(st->ctx->device);
    cudaStreamSynchronize(st->ctx->stream);
    st->G.release();
    st->cs.release();
    st->zs.release();
    st->pbar.release();
    st->tail.release();
    delete st;
}

int halo_ipa_blind_commit(halo_ipa* st, const uint64_t* q, uint64_t n_q, uint64_t out_jac[12]) {
    if (!st || !q || !out_jac) return HALO_EINVAL;
    if (n_q == 0 || n_q >= st->n || st->round != 0) {
        st->ctx->last_error = "halo_ipa_blind_commit: need 0 < n_q < n before the first round";
        return HALO_EINVAL;
    }
    IPA_TRY(st)
    st->pbar.reserve(st->n * sizeof(fr_t));
    ctx->stage_scalars.reserve(n_q * sizeof(fr_t));
    HALO_CUDA(cudaMemcpyAsync(ctx->stage_scalars.p, q, n_q * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    vec_pbar(ctx, ctx->stage_scalars.as<fr_t>(), n_q, st->z, st->n, st->pbar.as<fr_t>());
    xyzz_t r;
    msm_gens_device(ctx, st->pbar.as<fr_t>(), 0, st->n, r);  // commit(p_bar) without the w_bar S term
    jac_t j;
    xyzz_to_jac(j, r);
    memcpy(out_jac, &j, 96);
    IPA_CATCH
}

int halo_ipa_blind_apply(halo_ipa* st, const uint64_t alpha[4]) {
    if (!st || !alpha || !st->pbar.p || st->round != 0) return HALO_EINVAL;
    IPA_TRY(st)
    fr_t a;
    memcpy(&a, alpha, 32);
    vec_axpy(ctx, st->cs.as<fr_t>(), st->pbar.as<fr_t>(), a, st->n);  // p' = p + alpha p_bar  (pcdl.rs:156)
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    st->pbar.release();
    IPA_CATCH
}

int halo_ipa_set_hprime(halo_ipa* st, const uint64_t Hprime_jac[12]) {
    if (!st || !Hprime_jac) return HALO_EINVAL;
    IPA_TRY(st)
    jac_t j;
    memcpy(&j, Hprime_jac, 96);
    xyzz_t x;
    jac_to_xyzz(x, j);
    affine_t a;
    xyzz_to_affine(a, x);
    HALO_CUDA(cudaMemcpyAsync(st->tail.p, &a, sizeof a, cudaMemcpyHostToDevice, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    st->have_hprime = true;
    IPA_CATCH
}

int halo_ipa_round_lr(halo_ipa* st, uint64_t L_jac[12], uint64_t R_jac[12]) {
    if (!st || !L_jac || !R_jac) return HALO_EINVAL;
    if (!st->have_hprime || st->cur < 2 || st->lr_done) {
        st->ctx->last_error = "halo_ipa_round_lr: call order (set_hprime, then lr/fold alternating, lg n rounds)";
        return HALO_ESTATE;
    }
    IPA_TRY(st)
    const uint64_t m = st->cur / 2;
    affine_t* G = st->G.as<affine_t>();
    fr_t* c = st->cs.as<fr_t>();
    fr_t* z = st->zs.as<fr_t>();
    affine_t* Hp = st->tail.as<affine_t>();
    fr_t* scal = reinterpret_cast<fr_t*>(reinterpret_cast<char*>(st->tail.p) + sizeof(affine_t));
    vec_dot(ctx, c + m, z, m, scal + 2, scal);          // dot_l = <c_r, z_l>
    vec_dot(ctx, c, z + m, m, scal + 2, scal + 1);      // dot_r = <c_l, z_r>
    MsmInput in[2];
    in[0].bases = G;       // L = <c_r, g_l> + dot_l H'
    in[0].scalars = c + m;
    in[1].bases = G + m;   // R = <c_l, g_r> + dot_r H'
    in[1].scalars = c;
    for (int k = 0; k < 2; k++) {
        in[k].n = (uint32_t)m;
        in[k].tail_bases = Hp;
        in[k].tail_scalars = scal + k;
        in[k].n_tail = 1;
    }
    xyzz_t out[2];
    msm_batch(ctx, in, 2, out);
    jac_t j;
    xyzz_to_jac(j, out[0]);
    memcpy(L_jac, &j, 96);
    xyzz_to_jac(j, out[1]);
    memcpy(R_jac, &j, 96);
    st->lr_done = true;
    IPA_CATCH
}

int halo_ipa_round_fold(halo_ipa* st, const uint64_t xi[4], const uint64_t xi_inv[4]) {
    if (!st || !xi || !xi_inv) return HALO_EINVAL;
    if (!st->lr_done) {
        st->ctx->last_error = "halo_ipa_round_fold: round_lr must precede fold";
        return HALO_ESTATE;
    }
    IPA_TRY(st)
    const uint64_t m = st->cur / 2;
    fr_t x, xinv;
    memcpy(&x, xi, 32);
    memcpy(&xinv, xi_inv, 32);
    NafDigits naf;
    make_naf(x, naf);
    k_fold_points<<<(unsigned)((m + 127) / 128), 128, 0, ctx->stream>>>(st->G.as<affine_t>(), m, naf);
    ctx->kernel_launches++;
    HALO_CUDA(cudaGetLastError());
    vec_fold_scalars(ctx, st->cs.as<fr_t>(), st->zs.as<fr_t>(), m, x, xinv);
    st->cur = m;
    st->round++;
    st->lr_done = false;
    IPA_CATCH
}

int halo_ipa_finish(halo_ipa* st, uint64_t U_jac[12], uint64_t c_out[4]) {
    if (!st || !U_jac || !c_out) return HALO_EINVAL;
    if (st->cur != 1) {
        st->ctx->last_error = "halo_ipa_finish: rounds remaining";
        return HALO_ESTATE;
    }
    IPA_TRY(st)
    affine_t u;
    HALO_CUDA(cudaMemcpyAsync(&u, st->G.p, sizeof u, cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaMemcpyAsync(c_out, st->cs.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    HALO_CUDA(cudaStreamSynchronize(ctx->stream));
    xyzz_t x;
    xyzz_from_affine(x, u);
    jac_t j;
    xyzz_to_jac(j, x);
    memcpy(U_jac, &j, 96);  // U = G_(lg n)[0]  (pcdl.rs:230)
    IPA_CATCH
}

}  // extern "C"
