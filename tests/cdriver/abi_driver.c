/*
 * tests/cdriver/abi_driver.c -- TEST INFRASTRUCTURE: a plain C11 consumer of the drop-in boundary.
 *
 * The reference crate would bind include/halo_b200.h from Rust; no Rust toolchain exists in this image,
 * so this is the closest compiled stand-in: a C program that includes the headers AS C, links libhalo_b200.so and
 * libhalo_host.so the way a build.rs would (-lhalo_b200), and drives the path of BASELINE config 1 through the ABI:
 *   generators -> MSM (halo_msm_gens, halo_msm, halo_msm_multi) -> pcdl commit / hiding open / check ->
 *   acc prover / verifier / decider,
 * every result compared with the CPU oracle (oracle/liboracle.so, linked here as the checker only).
 *
 * Exit status: 0 all equal; 1 a mismatch (printed); 2 no usable CUDA device (the library has no CPU fallback, and this
 * program must say so instead of computing anything: tests/test_cdriver.py asserts exactly that on a machine without GPU).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/halo_b200.h"
#include "../../include/halo_pcdl.h"
#include "../../oracle/halo_oracle.h"

#define CHECK(cond, what)                                                  \
    do {                                                                   \
        if (!(cond)) {                                                     \
            fprintf(stderr, "MISMATCH: %s (line %d)\n", what, __LINE__);   \
            return 1;                                                      \
        }                                                                  \
    } while (0)
#define OK(ctx, call)                                                                              \
    do {                                                                                           \
        int rc_ = (call);                                                                          \
        if (rc_ != 0) {                                                                            \
            fprintf(stderr, "%s failed: %d (%s)\n", #call, rc_, halo_last_error(ctx));             \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

/* seeded scalars: x(tag, i) = from_le_bytes_mod_order(SHA3-256(tag || i)), the input rule of tests/golden/make_kat.py */
static void scalars(const char *tag, uint64_t n, uint64_t *out /*[n][4] Montgomery*/) {
    uint8_t msg[64], dg[32];
    size_t tl = strlen(tag);
    memcpy(msg, tag, tl);
    for (uint64_t i = 0; i < n; i++) {
        for (int b = 0; b < 8; b++) msg[tl + b] = (uint8_t)(i >> (8 * b));
        orc_sha3_256(msg, tl + 8, dg);
        orc_fp_from_le_bytes_mod_order(1, dg, out + 4 * i);
    }
}

int main(void) {
    enum { LG = 10, N = 1 << LG, D = N - 1, LEN = N - 100 };
    halo_ctx *ctx = NULL;
    int rc = halo_ctx_create(0, N, &ctx);
    if (rc != HALO_OK || !ctx) {
        fprintf(stderr, "halo_ctx_create: %d -- no usable CUDA device; libhalo_b200 (%s) has no CPU fallback\n", rc, halo_curve_name());
        return 2;
    }
    orc_init();
    OK(ctx, halo_derive_generators(ctx, N));
    static uint64_t gs[N][8], S[12], H[12];
    OK(ctx, halo_get_generators(ctx, 0, N, &gs[0][0]));
    OK(ctx, halo_get_SH(ctx, S, H));
    orc_derive_params(N); /* the oracle derives its own parameters: main.rs:18-45 */
    CHECK(memcmp(gs, orc_params_gs(), sizeof gs) == 0, "generators (bytes)");
    uint64_t So[12], Ho[12];
    orc_params_SH(So, Ho);
    CHECK(orc_pt_eq(S, So) && orc_pt_eq(H, Ho), "S / H");

    /* group.rs:24-26 */
    static uint64_t p[N][4], q[N][4];
    uint64_t got[12], exp[12];
    scalars("cdriver/p", LEN, &p[0][0]);
    OK(ctx, halo_msm_gens(ctx, &p[0][0], 0, LEN, got));
    orc_msm_affine(&gs[0][0], NULL, &p[0][0], LEN, 1, exp);
    CHECK(orc_pt_eq(got, exp), "halo_msm_gens");
    OK(ctx, halo_msm(ctx, &gs[17][0], NULL, &p[0][0], 333, got));
    orc_msm_affine(&gs[17][0], NULL, &p[0][0], 333, 1, exp);
    CHECK(orc_pt_eq(got, exp), "halo_msm");
    halo_msm_desc descs[3] = {{&gs[5][0], NULL, &p[0][0], 40, 0}, {NULL, NULL, &p[7][0], 2, 3}, {&gs[0][0], NULL, &p[1][0], 64, 0}};
    uint64_t multi[3][12];
    OK(ctx, halo_msm_multi(ctx, descs, 3, &multi[0][0]));
    orc_msm_affine(&gs[5][0], NULL, &p[0][0], 40, 1, exp);
    CHECK(orc_pt_eq(multi[0], exp), "halo_msm_multi[0]");
    orc_msm_affine(&gs[3][0], NULL, &p[7][0], 2, 1, exp);
    CHECK(orc_pt_eq(multi[1], exp), "halo_msm_multi[1] (resident generators)");
    orc_msm_affine(&gs[0][0], NULL, &p[1][0], 64, 1, exp);
    CHECK(orc_pt_eq(multi[2], exp), "halo_msm_multi[2]");

    /* pcdl.rs: commit, hiding open, check */
    uint64_t w[4], z[4], wbar[4], v[4], C[12], Co[12];
    scalars("cdriver/w", 1, w);
    scalars("cdriver/z", 1, z);
    scalars("cdriver/wbar", 1, wbar);
    scalars("cdriver/q", LEN - 1, &q[0][0]);
    OK(ctx, halo_pcdl_commit(ctx, &p[0][0], LEN, D, w, C));
    CHECK(orc_pcdl_commit(&p[0][0], LEN, D, w, 1, Co) == 0 && orc_pt_eq(C, Co), "pcdl::commit");
    static uint64_t zs[N][4];
    orc_construct_powers(z, LEN, &zs[0][0]);
    orc_scalar_dot(&p[0][0], &zs[0][0], LEN, v);
    uint64_t v_gpu[4];
    OK(ctx, halo_scalar_dot(ctx, &p[0][0], &zs[0][0], LEN, v_gpu));
    CHECK(memcmp(v, v_gpu, 32) == 0, "scalar_dot");
    static halo_eval_proof pi;
    static orc_eval_proof pio;
    OK(ctx, halo_pcdl_open(ctx, &p[0][0], LEN, C, D, z, w, &q[0][0], LEN - 1, wbar, &pi));
    CHECK(orc_pcdl_open(&p[0][0], LEN, C, D, z, w, &q[0][0], LEN - 1, wbar, 1, &pio) == 0, "oracle open");
    CHECK(pi.lg_n == LG && pio.lg_n == LG && pi.hiding == 1, "proof shape");
    for (int i = 0; i < LG; i++) CHECK(orc_pt_eq(pi.Ls[i], pio.Ls[i]) && orc_pt_eq(pi.Rs[i], pio.Rs[i]), "L_i / R_i");
    CHECK(orc_pt_eq(pi.U, pio.U) && memcmp(pi.c, pio.c, 32) == 0, "U / c");
    CHECK(orc_pt_eq(pi.C_bar, pio.C_bar) && memcmp(pi.w_prime, pio.w_prime, 32) == 0, "C_bar / w'");
    OK(ctx, halo_pcdl_check(ctx, C, D, z, v, &pi));
    CHECK(sizeof(halo_eval_proof) == sizeof(orc_eval_proof), "struct layout");
    CHECK(orc_pcdl_check(C, D, z, v, (const orc_eval_proof *)&pi, 1) == 0, "oracle accepts the GPU proof");
    halo_eval_proof bad = pi;
    bad.c[0] ^= 1;
    CHECK(halo_pcdl_check(ctx, C, D, z, v, &bad) == HALO_REJECT_SUCCINCT, "a flipped proof is rejected");

    /* acc.rs: one accumulation step and the decider */
    static halo_instance inst;
    static halo_accumulator acc;
    memcpy(inst.C, C, 96);
    inst.d = D;
    memcpy(inst.z, z, 32);
    memcpy(inst.v, v, 32);
    inst.pi = pi;
    uint64_t h0[2][4], wa[4], wb2[4];
    static uint64_t qa[N][4];
    scalars("cdriver/h0", 2, &h0[0][0]);
    scalars("cdriver/wa", 1, wa);
    scalars("cdriver/wb2", 1, wb2);
    scalars("cdriver/qa", N - 1, &qa[0][0]);
    OK(ctx, halo_acc_prover(ctx, D, &inst, 1, h0, wa, &qa[0][0], N - 1, wb2, &acc));
    OK(ctx, halo_acc_verifier(ctx, D, &inst, 1, &acc));
    OK(ctx, halo_acc_decider(ctx, &acc));
    CHECK(sizeof(halo_accumulator) == sizeof(orc_accumulator) && sizeof(halo_instance) == sizeof(orc_instance), "struct layout");
    CHECK(orc_acc_verifier(D, (const orc_instance *)&inst, 1, (const orc_accumulator *)&acc, 1) == 0, "oracle verifier accepts");
    CHECK(orc_acc_decider((const orc_accumulator *)&acc, 1) == 0, "oracle decider accepts");
    static orc_accumulator acco;
    CHECK(orc_acc_prover(D, (const orc_instance *)&inst, 1, h0, wa, &qa[0][0], N - 1, wb2, 1, &acco) == 0, "oracle prover");
    CHECK(orc_pt_eq(acc.C_bar, acco.C_bar) && memcmp(acc.z, acco.z, 32) == 0 && memcmp(acc.v, acco.v, 32) == 0, "accumulator");
    CHECK(orc_pt_eq(acc.pi.U, acco.pi.U) && memcmp(acc.pi.c, acco.pi.c, 32) == 0, "accumulator proof");

    printf("abi_driver ok (%s): %llu kernel launches\n", halo_curve_name(), (unsigned long long)halo_kernel_launches(ctx));
    halo_ctx_destroy(ctx);
    return 0;
}
