/*
 * tests/cdriver/mgpu_driver.c -- TEST INFRASTRUCTURE: a plain C11 consumer of the multi-GPU part of the drop-in boundary.
 *
 * What a Rust caller of group.rs:24-26 (`point_dot_affine` over GS[0..n), pedersen.rs:14) binds when it has more than one
 * GPU: halo_mgpu_create(devices, g, n_total, window) once, then halo_mgpu_msm_gens(scalars, n) per commitment.  The shard
 * by point slice, the per-device copies, the ncclAllGather and the ordered sum all happen inside libhalo_b200.so.
 * Every result is compared with the CPU oracle (linked here as the checker only), both its arkworks-shaped Pippenger over
 * the whole point set and the discrete-log property of the derived generators.
 *
 * usage: mgpu_driver.bin <g>     exit status 0 all equal; 1 mismatch / error; 2 fewer than g usable devices
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/halo_b200.h"
#include "../../oracle/halo_oracle.h"

void orc_derive_points_fast(uint64_t start, uint64_t count, uint64_t *out_affine);

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

int main(int argc, char **argv) {
    int g = argc > 1 ? atoi(argv[1]) : 2;
    if (g < 1 || g > 8) return 1;
    int devs[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    const uint64_t N = (1u << 18) + 77; /* uneven slices */
    orc_init();
    uint64_t *gs = malloc(N * 64), *sc = malloc(N * 32);
    if (!gs || !sc) return 1;
    orc_derive_points_fast(2, N, gs);
    scalars("halo-b200-mgpu/s", N, sc);
    for (int window = -1; window <= 0; window++) { /* -1: variable base; 0: FIXED-base tables, automatic window */
        halo_mgpu *m = NULL;
        int rc = halo_mgpu_create(devs, g, N, window, &m);
        if (rc != HALO_OK || !m) {
            fprintf(stderr, "halo_mgpu_create(g = %d): %d -- needs %d usable CUDA devices and libnccl.so.2; no CPU fallback\n", g, rc, g);
            return 2;
        }
        if (halo_mgpu_size(m) != g) return 1;
        const uint64_t sizes[4] = {N, N - 12345, 5000, 0};
        for (int k = 0; k < 4; k++) {
            uint64_t n = sizes[k], got[12], exp[12], dl[12];
            rc = halo_mgpu_msm_gens(m, sc, n, got);
            if (rc != HALO_OK) {
                fprintf(stderr, "halo_mgpu_msm_gens(n = %llu): %d (%s)\n", (unsigned long long)n, rc, halo_mgpu_last_error(m));
                return 1;
            }
            orc_msm_affine(gs, NULL, sc, n, 8, exp);
            orc_msm_derived_by_dlog(0, sc, n, 8, dl);
            if (!orc_pt_eq(got, exp) || !orc_pt_eq(got, dl) || !halo_points_equal(got, exp)) {
                fprintf(stderr, "MISMATCH: window %d, n = %llu\n", window, (unsigned long long)n);
                return 1;
            }
        }
        halo_mgpu_destroy(m);
    }
    free(gs);
    free(sc);
    printf("mgpu_driver ok: %d GPUs, NCCL %d\n", g, halo_nccl_version());
    return 0;
}
