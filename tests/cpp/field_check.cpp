// Host-side check of csrc/gl64.cuh (the same header the device code compiles): gl_mul_pow2<K> against
// gl_mul by 2^K for every K, the power-of-two roots of unity the last NTT pass relies on, and the
// add / sub / canon identities.  Built and run by tests/test_field_host.py.
#include "../../plonky2_aes_b200/csrc/gl64.cuh"
#include <stdio.h>
#include <stdlib.h>

static gl_t rnd() { return ((gl_t)rand() << 42) ^ ((gl_t)rand() << 21) ^ (gl_t)rand(); }
template <int K> static int chk() {
    const gl_t w = gl_pow(2, K);
    const gl_t xs[] = {0, 1, GL_P - 1, GL_P - 2, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFF00000000ULL, 0x8000000000000000ULL};
    int bad = 0;
    for (int i = 0; i < 20000; i++) {
        gl_t x = gl_canon(i < 8 ? xs[i] : rnd());
        if (x >= GL_P) x -= GL_P;
        if (gl_mul_pow2<K>(x) != gl_mul(x, w)) bad++;
    }
    if (bad) printf("K=%d bad=%d\n", K, bad);
    return bad;
}
template <int K> struct All { static int run() { return chk<K>() + All<K - 1>::run(); } };
template <> struct All<0> { static int run() { return 0; } };

int main() {
    int bad = All<95>::run();
    // plonky2's roots of unity of order 16 and 64 are 2^156 and 2^39 (2 has order 192)
    if (gl_root_of_unity(4) != gl_pow(2, 156)) { printf("w16 mismatch\n"); bad++; }
    if (gl_root_of_unity(6) != gl_pow(2, 39)) { printf("w64 mismatch\n"); bad++; }
    if (gl_pow(2, 192) != 1 || gl_pow(2, 96) != GL_P - 1) { printf("order of 2 mismatch\n"); bad++; }
    for (int i = 0; i < 100000; i++) {
        gl_t a = gl_canon(rnd()), b = gl_canon(rnd());
        if (a >= GL_P) a -= GL_P;
        if (b >= GL_P) b -= GL_P;
        if (gl_sub(gl_add(a, b), b) != a) bad++;
        if (gl_add(gl_sub(a, b), b) != a) bad++;
        if (gl_mul(a, gl_inv(a ? a : 1)) != (a ? 1 : 0) && a) bad++;
    }
    printf("bad %d\n", bad);
    return bad != 0;
}
