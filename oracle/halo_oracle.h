/*
 * oracle/halo_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the halo-accumulation hot path (reference: rasmus-kirk/halo-accumulation,
 * code/src/{group,pedersen,pcdl,acc,main}.rs) plus the arkworks 0.5.0 behaviour underneath it
 * (ark-ff / ark-ec / ark-pallas / ark-serialize, pinned in code/Cargo.lock:51-167; sources are
 * NOT in the reference tree, so their published algorithms are restated here).
 *
 * PARITY PINNING: anchored to the reference's only golden data, the 16 386 points of
 * code/src/consts.rs (S, H, GS[0..16384]); see tests/test_oracle_golden.py.  Commitment / proof /
 * challenge VALUES are "parity unpinned": the reference stores no expected bytes for them and
 * cannot be executed here (no Rust toolchain).  See DESIGN.md section "Oracle".
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libhalo_b200.so) never links or calls it.
 *
 * Data layout everywhere: field elements are 4 x u64 little-endian limbs in Montgomery form
 * (R = 2^256), exactly arkworks' in-memory `Fp.0.0`.  Affine point = x[4] | y[4] (+ separate
 * infinity byte).  Jacobian point = x[4] | y[4] | z[4], infinity <=> z == 0.
 */
#ifndef HALO_ORACLE_H
#define HALO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LG 32

/* pcdl.rs:22-30 EvalProof */
typedef struct {
    uint32_t lg_n;
    uint32_t hiding;               /* C_bar / w_prime are Some(..) */
    uint64_t Ls[ORC_MAX_LG][12];   /* Jacobian */
    uint64_t Rs[ORC_MAX_LG][12];
    uint64_t U[12];
    uint64_t c[4];
    uint64_t C_bar[12];
    uint64_t w_prime[4];
} orc_eval_proof;

/* acc.rs:21-28 Instance */
typedef struct {
    uint64_t C[12];
    uint64_t d;
    uint64_t z[4];
    uint64_t v[4];
    orc_eval_proof pi;
} orc_instance;

/* acc.rs:43-59 Accumulator + AccumulatorHiding (pi_V = (h_0, U_0, w), h_0 = rand degree-1 poly) */
typedef struct {
    uint64_t C_bar[12];
    uint64_t d;
    uint64_t z[4];
    uint64_t v[4];
    orc_eval_proof pi;
    uint64_t h0[2][4];
    uint64_t U0[12];
    uint64_t w[4];
} orc_accumulator;

/* error codes shared by the verifier-style functions (0 = accept) */
enum {
    ORC_OK = 0,
    ORC_EINVAL = -1,        /* d+1 not a power of two / d too large (pcdl.rs:102-104,261-262) */
    ORC_ELEN = -2,          /* length mismatch (pedersen.rs:7-12) */
    ORC_REJECT_SUCCINCT = -10,  /* pcdl.rs:307-310 */
    ORC_REJECT_U = -11,         /* pcdl.rs:339 */
    ORC_REJECT_U0 = -12,        /* acc.rs:152-155 */
    ORC_REJECT_D = -13,         /* acc.rs:169 */
    ORC_REJECT_CBAR = -14,      /* acc.rs:237 */
    ORC_REJECT_Z = -15,         /* acc.rs:238 */
    ORC_REJECT_V = -17          /* acc.rs:240 */
};

void orc_init(void);
int orc_num_threads(void);

/* ---- hashing ---- */
void orc_sha3_256(const uint8_t *msg, uint64_t len, uint8_t out[32]);

/* ---- field arithmetic (which: 0 = Fq base field, 1 = Fr scalar field) ---- */
void orc_fp_mul(int which, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
uint64_t orc_fp_mul_cross(int which, const uint64_t *a, const uint64_t *b, uint64_t n); /* asm vs C definition, count of mismatches */
void orc_fp_add(int which, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_fp_sub(int which, const uint64_t a[4], const uint64_t b[4], uint64_t r[4]);
void orc_fp_inv(int which, const uint64_t a[4], uint64_t r[4]);
void orc_fp_to_canon(int which, const uint64_t a[4], uint64_t r[4]);
void orc_fp_from_canon(int which, const uint64_t a[4], uint64_t r[4]);
void orc_fp_from_le_bytes_mod_order(int which, const uint8_t b[32], uint64_t r[4]);

/* ---- curve ---- */
void orc_pt_add(const uint64_t a[12], const uint64_t b[12], uint64_t r[12]);
void orc_pt_add_affine(const uint64_t a[12], const uint64_t b_aff[8], int b_inf, uint64_t r[12]);
void orc_pt_double(const uint64_t a[12], uint64_t r[12]);
void orc_pt_mul(const uint64_t p[12], const uint64_t k[4], uint64_t r[12]);
int orc_pt_eq(const uint64_t a[12], const uint64_t b[12]);
/* returns infinity flag */
int orc_pt_to_affine(const uint64_t p[12], uint64_t aff[8]);
void orc_pt_from_affine(const uint64_t aff[8], int inf, uint64_t p[12]);
int orc_pt_on_curve_affine(const uint64_t aff[8]);
/* arkworks serialize_compressed of a point: 33 bytes (see .c) */
void orc_pt_serialize_compressed(const uint64_t p[12], uint8_t out[33]);

/* ---- public parameters: main.rs:18-45 ---- */
/* P_k = [SHA3-256(genesis || k as u64 LE) mod r] * (-1, 2), k in [start, start+count) -> affine */
void orc_derive_points(uint64_t start, uint64_t count, uint64_t *out_affine /*[count][8]*/);
/* same points through a fixed-base table of (-1, 2) (setup of the benchmark's CPU arm; checked against the line above) */
void orc_derive_points_fast(uint64_t start, uint64_t count, uint64_t *out_affine /*[count][8]*/);
/* sum_i scalars[i] * G_{first+i} for the DERIVED generators via their known discrete logs:
 * (sum_i scalars[i] * s_{first+i+2}) * (-1, 2).  Property check, cost O(n) hashes. */
void orc_msm_derived_by_dlog(uint64_t first, const uint64_t *scalars, uint64_t n, int threads, uint64_t out[12]);
/* S = P_0, H = P_1 (Jacobian with z = 1), GS[i] = P_{i+2}  */
void orc_set_params(const uint64_t S[12], const uint64_t H[12], const uint64_t *gs_affine, uint64_t n);
void orc_derive_params(uint64_t n); /* derive and install S, H, GS[0..n) */
const uint64_t *orc_params_gs(void);
void orc_params_SH(uint64_t S[12], uint64_t H[12]);
uint64_t orc_params_n(void);

/* ---- group.rs ---- */
/* group.rs:24-26 point_dot_affine -> VariableBaseMSM::msm_unchecked, arkworks-shaped Pippenger.
 * threads <= 1: serial (the reference's configuration); > 1: windows spread over OpenMP threads. */
void orc_msm_affine(const uint64_t *bases_affine, const uint8_t *inf /*nullable*/, const uint64_t *scalars,
                    uint64_t n, int threads, uint64_t out[12]);
/* naive sum of double-and-add products, independent cross-check of the Pippenger */
void orc_msm_naive(const uint64_t *bases_affine, const uint8_t *inf, const uint64_t *scalars, uint64_t n,
                   uint64_t out[12]);
/* group.rs:18-21 point_dot: per-element into_affine then MSM */
void orc_point_dot(const uint64_t *scalars, const uint64_t *points_jac, uint64_t n, int threads, uint64_t out[12]);
/* group.rs:13-15 */
void orc_scalar_dot(const uint64_t *xs, const uint64_t *ys, uint64_t n, uint64_t out[4]);
/* group.rs:29-37 */
void orc_construct_powers(const uint64_t z[4], uint64_t n, uint64_t *out);

/* ---- pedersen.rs:6-20 ---- */
int orc_pedersen_commit(const uint64_t *w /*nullable*/, const uint64_t *gs_affine, uint64_t n_gs,
                        const uint64_t *ms, uint64_t n_ms, int threads, uint64_t out[12]);

/* ---- pcdl.rs ---- */
/* HPoly::get_poly pcdl.rs:56-77: coefficient vector of prod_{i<lg n}(1 + xi_{lg n - i} X^{2^i}) */
void orc_h_get_poly(const uint64_t *xis /*[lg_n+1][4]*/, uint32_t lg_n, uint64_t *out /*[n][4]*/);
/* HPoly::eval pcdl.rs:79-91 */
void orc_h_eval(const uint64_t *xis, uint32_t lg_n, const uint64_t z[4], uint64_t out[4]);
/* pcdl.rs:99-110 */
int orc_pcdl_commit(const uint64_t *coeffs, uint64_t n_coeffs, uint64_t d, const uint64_t *w /*nullable*/,
                    int threads, uint64_t out[12]);
/* pcdl.rs:120-242.  Randomness is explicit (the reference draws it from `rng` at :141 and :146):
 * q = the deg(p)-1 polynomial (n_q = deg(p) coefficients), w_bar; both ignored when w == NULL. */
int orc_pcdl_open(const uint64_t *p_coeffs, uint64_t n_coeffs, const uint64_t C[12], uint64_t d,
                  const uint64_t z[4], const uint64_t *w /*nullable*/, const uint64_t *q_coeffs, uint64_t n_q,
                  const uint64_t *w_bar, int threads, orc_eval_proof *pi);
/* pcdl.rs:252-314: on accept writes xis[lg_n+1][4] (the HPoly) and U */
int orc_pcdl_succinct_check(const uint64_t C[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                            const orc_eval_proof *pi, uint64_t *xis_out, uint64_t U_out[12]);
/* pcdl.rs:323-342 */
int orc_pcdl_check(const uint64_t C[12], uint64_t d, const uint64_t z[4], const uint64_t v[4],
                   const orc_eval_proof *pi, int threads);

/* ---- acc.rs ---- */
/* acc.rs:190-220.  Randomness explicit, in the reference's draw order: h0 (2 coeffs, :192), w (:198),
 * then open's q (n_q coeffs) and w_bar. */
int orc_acc_prover(uint64_t d, const orc_instance *qs, uint64_t m, const uint64_t h0[2][4], const uint64_t w[4],
                   const uint64_t *q_coeffs, uint64_t n_q, const uint64_t w_bar[4], int threads,
                   orc_accumulator *acc);
/* acc.rs:223-243 */
int orc_acc_verifier(uint64_t d, const orc_instance *qs, uint64_t m, const orc_accumulator *acc, int threads);
/* acc.rs:245-255 */
int orc_acc_decider(const orc_accumulator *acc, int threads);
/* From<Accumulator> for Instance, acc.rs:121-131 */
void orc_acc_to_instance(const orc_accumulator *acc, orc_instance *q);

#ifdef __cplusplus
}
#endif
#endif
