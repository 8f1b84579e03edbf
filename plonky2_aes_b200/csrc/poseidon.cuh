// Poseidon-Goldilocks width-12 permutation (x^7, 4+22+4 rounds) for host and device.
// Replaces plonky2/src/hash/{poseidon.rs,poseidon_goldilocks.rs,hashing.rs} of the pinned
// dependency (/root/reference/Cargo.toml:12; entered from e.g.
// /root/reference/aes-gcm/src/circuit_gcm.rs:781 `data.prove(pw)` and directly from
// /root/reference/feistel/src/lib.rs:106 `hash_n_to_hash_no_pad`).
//
// One permutation per thread.  The MDS layer exploits the small circulant coefficients (<= 41):
// state words are split in 32-bit halves, each half is accumulated with IMAD.WIDE into a 64-bit
// sum (12 * 41 * 2^32 < 2^42, no overflow), and the two sums are recombined with one fold.
#pragma once
#include "gl64.cuh"

#if defined(__CUDACC__)
__constant__ gl_t POSEIDON_RC_DEV[360] = {
#include "poseidon_rc.inc"
};
#endif
static const gl_t POSEIDON_RC_HOST[360] = {
#include "poseidon_rc.inc"
};

#if defined(__CUDA_ARCH__)
#define POSEIDON_RC POSEIDON_RC_DEV
#else
#define POSEIDON_RC POSEIDON_RC_HOST
#endif

GL_HD gl_t poseidon_sbox(gl_t x) {
    gl_t x2 = gl_mul_lazy(x, x);
    gl_t x4 = gl_mul_lazy(x2, x2);
    gl_t x3 = gl_mul_lazy(x, x2);
    return gl_mul_lazy(x3, x4);
}

// s: any u64 residues in, lazy residues out
GL_HD void poseidon_mds(gl_t s[12]) {
    const uint32_t C[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    uint32_t lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) { lo[i] = (uint32_t)s[i]; hi[i] = (uint32_t)(s[i] >> 32); }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        uint64_t al = 0, ah = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (uint64_t)lo[(i + r) % 12] * C[i];
            ah += (uint64_t)hi[(i + r) % 12] * C[i];
        }
        if (r == 0) { al += (uint64_t)lo[0] * 8u; ah += (uint64_t)hi[0] * 8u; }
        // value = al + ah * 2^32  (ah < 2^42)
        uint64_t low = al + (ah << 32);
        uint64_t carry = low < al ? 1 : 0;
        uint64_t top = (ah >> 32) + carry;           // < 2^11, weight 2^64 = EPS
        uint64_t t = top * GL_EPS;                   // < 2^43
        uint64_t res = low + t;
        if (res < t) res += GL_EPS;
        s[r] = res;
    }
}

GL_HD void poseidon_permute(gl_t s[12]) {
    int r = 0;
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_lazy(s[i], POSEIDON_RC[12 * r + i]));
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int k = 0; k < 22; k++, r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add_lazy(s[i], POSEIDON_RC[12 * r + i]);
        s[0] = poseidon_sbox(s[0]);
        poseidon_mds(s);
    }
#pragma unroll 1
    for (int k = 0; k < 4; k++, r++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(gl_add_lazy(s[i], POSEIDON_RC[12 * r + i]));
        poseidon_mds(s);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_canon(s[i]);
}

// two_to_one(l, r): Poseidon([l, r, 0, 0, 0, 0])[0..4]
GL_HD void poseidon_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]) {
    gl_t s[12];
#pragma unroll
    for (int i = 0; i < 4; i++) { s[i] = l[i]; s[4 + i] = r[i]; s[8 + i] = 0; }
    poseidon_permute(s);
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = s[i];
}
