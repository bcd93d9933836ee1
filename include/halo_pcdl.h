/*
 * halo_pcdl.h -- C view of the host layer above the CUDA ABI: the reference's public API for the hot
 * path, same names, argument meaning and error behaviour (code/src/pedersen.rs, pcdl.rs, acc.rs), with the
 * data-parallel work dispatched to libhalo_b200.so (include/halo_b200.h).
 *
 * The reference is Rust; no Rust toolchain exists in this image, so the host side is C++
 * (halo-accumulation_b200/host/ *.hpp, namespaces halo::pedersen / halo::pcdl / halo::acc) and this header
 * exposes it to C / ctypes for the parity tests and the benchmark.  A Rust build would keep its own
 * pcdl.rs / acc.rs and bind halo_b200.h directly.
 *
 * Randomness: the reference draws from `rng` inside open (pcdl.rs:141,146) and prover (acc.rs:192,198).
 * Here the caller passes those draws explicitly, in the reference's draw order, so that results are
 * reproducible and comparable bit for bit.
 *
 * Return value: HALO_OK (accept / success), a HALO_E* error (halo_b200.h) -- the analogue of the
 * reference's assert!/panic -- or a HALO_REJECT_* code -- the analogue of its `ensure!` Err.
 */
#ifndef HALO_PCDL_H
#define HALO_PCDL_H
#include "halo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define HALO_MAX_LG 32

#define HALO_REJECT_SUCCINCT (-10) /* "C_(log_n) != CM.Commit_Sigma(c || v')"         pcdl.rs:307-310 */
#define HALO_REJECT_U (-11)        /* "U != CM.Commit(ck, h_vec)"                      pcdl.rs:339    */
#define HALO_REJECT_U0 (-12)       /* "U_0 != PCDL.Commit(ck, h_0; w = bot)"           acc.rs:152-155 */
#define HALO_REJECT_D (-13)        /* "d_i != d" / "d' = d"                            acc.rs:169,239 */
#define HALO_REJECT_CBAR (-14)     /* "C_bar' != C_bar"                                acc.rs:237     */
#define HALO_REJECT_Z (-15)        /* "z' = z"                                         acc.rs:238     */
#define HALO_REJECT_V (-17)        /* "h(z) = v"                                       acc.rs:240     */

/* pcdl.rs:22-30 EvalProof */
typedef struct {
    uint32_t lg_n;
    uint32_t hiding; /* C_bar / w_prime are Some(..) */
    uint64_t Ls[HALO_MAX_LG][12];
    uint64_t Rs[HALO_MAX_LG][12];
    uint64_t U[12];
    uint64_t c[4];
    uint64_t C_bar[12];
    uint64_t w_prime[4];
} halo_eval_proof;

/* acc.rs:21-28 Instance */
typedef struct {
    uint64_t C[12];
    uint64_t d;
    uint64_t z[4];
    uint64_t v[4];
    halo_eval_proof pi;
} halo_instance;

/* acc.rs:43-59 Accumulator with pi_V = AccumulatorHiding { h (degree 1), U, w } */
typedef struct {
    uint64_t C_bar[12];
    uint64_t d;
    uint64_t z[4];
    uint64_t v[4];
    halo_eval_proof pi;
    uint64_t h0[2][4];
    uint64_t U0[12];
    uint64_t w[4];
} halo_accumulator;

/* pedersen.rs:6-20  commit(w, Gs, ms); gs_affine == NULL means GS[0..n_gs) of the context. */
int halo_pedersen_commit(halo_ctx *ctx, const uint64_t *w /*nullable*/, const uint64_t *gs_affine, uint64_t n_gs,
                         const uint64_t *ms, uint64_t n_ms, uint64_t out_jac[12]);
/* pcdl.rs:99-110 */
int halo_pcdl_commit(halo_ctx *ctx, const uint64_t *coeffs, uint64_t n_coeffs, uint64_t d, const uint64_t *w /*nullable*/,
                     uint64_t out_jac[12]);
/* pcdl.rs:120-242; hiding iff w != NULL, then q (n_q = deg p coefficients) and w_bar are the rng draws. */
int halo_pcdl_open(halo_ctx *ctx, const uint64_t *coeffs, uint64_t n_coeffs, const uint64_t C_jac[12], uint64_t d,
                   const uint64_t z[4], const uint64_t *w /*nullable*/, const uint64_t *q, uint64_t n_q,
                   const uint64_t *w_bar, halo_eval_proof *pi);
/* pcdl.rs:252-314; on accept writes the HPoly challenges xis[lg_n+1][4] and U. */
int halo_pcdl_succinct_check(halo_ctx *ctx, const uint64_t C_jac[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                             const halo_eval_proof *pi, uint64_t *xis_out, uint64_t U_out[12]);
/* pcdl.rs:323-342 */
int halo_pcdl_check(halo_ctx *ctx, const uint64_t C_jac[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                    const halo_eval_proof *pi);
/* HPoly::eval pcdl.rs:79-91 */
int halo_h_eval(const uint64_t *xis, uint32_t lg_n, const uint64_t z[4], uint64_t out[4]);

/* acc.rs:190-220; draws in order: h0 (2 coefficients), w, then open's q and w_bar. */
int halo_acc_prover(halo_ctx *ctx, uint64_t d, const halo_instance *qs, uint64_t m, const uint64_t h0[2][4],
                    const uint64_t w[4], const uint64_t *q, uint64_t n_q, const uint64_t w_bar[4], halo_accumulator *acc);
/* acc.rs:223-243 */
int halo_acc_verifier(halo_ctx *ctx, uint64_t d, const halo_instance *qs, uint64_t m, const halo_accumulator *acc);
/* acc.rs:245-255 */
int halo_acc_decider(halo_ctx *ctx, const halo_accumulator *acc);
/* impl From<Accumulator> for Instance, acc.rs:121-131 */
void halo_acc_to_instance(const halo_accumulator *acc, halo_instance *q);

/* Fiat-Shamir helpers exposed for the parity tests (group.rs:41-89). */
void halo_point_serialize_compressed(const uint64_t p_jac[12], uint8_t out[33]);

#ifdef __cplusplus
}
#endif
#endif /* HALO_PCDL_H */
