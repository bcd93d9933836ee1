// sha3.cuh -- SHA3-256 (FIPS 202) for short messages, host and device.
// Replaces the `sha3` crate's Sha3_256 at main.rs:22-25 (generator derivation, on device in K6) and
// group.rs:52-55 / :77-80 (Fiat-Shamir transcript, on the host).
#pragma once
#include <stdint.h>
#include "fp.cuh"

namespace halo {

HALO_HD uint64_t rotl64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }

HALO_HD void keccak_f1600(uint64_t st[25]) {
    const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    for (int round = 0; round < 24; round++) {
        uint64_t C[5], D[5], B[25];
#pragma unroll
        for (int x = 0; x < 5; x++) C[x] = st[x] ^ st[x + 5] ^ st[x + 10] ^ st[x + 15] ^ st[x + 20];
#pragma unroll
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rotl64(C[(x + 1) % 5], 1);
#pragma unroll
        for (int i = 0; i < 25; i++) st[i] ^= D[i % 5];
#pragma unroll
        for (int x = 0; x < 5; x++)
#pragma unroll
            for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rotl64(st[x + 5 * y], ROT[x + 5 * y]);
#pragma unroll
        for (int y = 0; y < 5; y++)
#pragma unroll
            for (int x = 0; x < 5; x++) st[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        st[0] ^= RC[round];
    }
}

// One-shot SHA3-256 of an arbitrary-length message; digest as 4 little-endian u64 words.
HALO_HD void sha3_256(const uint8_t* msg, uint64_t len, uint64_t digest[4]) {
    const int RATE = 136;
    uint64_t st[25];
    for (int i = 0; i < 25; i++) st[i] = 0;
    while (len >= (uint64_t)RATE) {
        for (int i = 0; i < RATE / 8; i++) {
            uint64_t v = 0;
            for (int j = 7; j >= 0; j--) v = (v << 8) | msg[8 * i + j];
            st[i] ^= v;
        }
        keccak_f1600(st);
        msg += RATE;
        len -= RATE;
    }
    for (uint64_t i = 0; i < len; i++) st[i >> 3] ^= (uint64_t)msg[i] << (8 * (i & 7));
    st[len >> 3] ^= (uint64_t)0x06 << (8 * (len & 7));
    st[(RATE - 1) >> 3] ^= (uint64_t)0x80 << (8 * ((RATE - 1) & 7));
    keccak_f1600(st);
    for (int i = 0; i < 4; i++) digest[i] = st[i];
}

}  // namespace halo
