// Cost of the building blocks of the Poseidon permutation on sm_100a, in cycles per warp and SM sub-partition:
// S-box, 128-bit product, accumulator read-out, I2F conversion, one full round, one partial-round pair, and the
// whole permutation.  256 threads x 3 blocks per SM like the Merkle kernels.  Build with -DP2G_DIAG_NO_MDS /
// -DP2G_DIAG_NO_SBOX for the integer-only / FP64-only halves.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../plonky2_aes_b200/csrc pos_parts.cu -o pos_parts
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "poseidon.cuh"
#ifndef TAG
#define TAG "base"
#endif
#define ITERS 512
#ifndef TPB
#define TPB 256
#endif
#ifndef MINB
#define MINB 3
#endif

template <int OP>
__global__ void __launch_bounds__(TPB, MINB) k(gl_t* out, uint32_t iters) {
#if defined(__CUDA_ARCH__)
    gl_t s[12];
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = (gl_t)g * 0x9E3779B97F4A7C15ull + i * 0x1234567ull;
    if (OP == 0) {              // 12 independent S-boxes per iteration
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = poseidon_sbox(s[i]);
        }
    } else if (OP == 1) {       // 12 multiplies (product + fold)
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = pmul_v1(s[i], s[i]);
        }
    } else if (OP == 2) {       // 12 products without the fold (words xor-ed)
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) {
                uint32_t l0, l1, h0, h1; pmul128(s[i], s[i] ^ it, l0, l1, h0, h1);
                s[i] = gl_pack(l0 ^ h0, l1 ^ h1);
            }
        }
    } else if (OP == 3) {       // 12 x (read-out + 2 I2F + 2 DADD)
        double dl[12], dh[12];
#pragma unroll
        for (int i = 0; i < 12; i++) { dl[i] = 4503599627370496.0 + (double)(uint32_t)s[i]; dh[i] = 4503599627370496.0 + (double)(uint32_t)(s[i] >> 32); }
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 12; i++) {
                const gl_t v = pos_readout(dl[i], dh[i]);
                dl[i] = __dadd_rn((double)(uint32_t)v, 4503599627370496.0 + 12345.0);
                dh[i] = __dadd_rn((double)(uint32_t)(v >> 32), 4503599627370496.0 + 777.0);
            }
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = __double_as_longlong(dl[i]) ^ __double_as_longlong(dh[i]);
    } else if (OP == 4) {       // 24 x (I2F + LOP3)
        uint32_t w[24];
#pragma unroll
        for (int i = 0; i < 12; i++) gl_unpack(s[i], w[2 * i], w[2 * i + 1]);
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 24; i++) {
                const double d = (double)w[i];
                w[i] = (uint32_t)__double2loint(d) ^ (uint32_t)__double2hiint(d);
            }
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pack(w[2 * i], w[2 * i + 1]);
    } else if (OP == 5) {       // 24 x LOP3 (to subtract from OP 4)
        uint32_t w[24];
#pragma unroll
        for (int i = 0; i < 12; i++) gl_unpack(s[i], w[2 * i], w[2 * i + 1]);
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 24; i++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[i]) : "r"(w[(i + 1) % 24]), "r"(it));
        }
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_pack(w[2 * i], w[2 * i + 1]);
    } else if (OP == 6) {       // full rounds only
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) poseidon_round<true>(s, 1 + (it & 3));
    } else if (OP == 7) {       // partial pairs only
        const int zero = pos_lane_zero();
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) poseidon_partial_pair(s, it % 11, zero);
    } else if (OP == 8) {       // whole permutation
#pragma unroll 1
        for (uint32_t it = 0; it < iters; it++) poseidon_permute_lazy(s);
    }
    gl_t acc = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) acc ^= s[i];
    out[g] = acc;
#endif
}
template <int OP> void run(const char* name, double units, gl_t* d, int sms, uint32_t iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * MINB * 4;
    k<OP><<<blocks, TPB>>>(d, iters);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<OP><<<blocks, TPB>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double warp_units = (double)blocks * (TPB / 32) * iters * units;
    const double cyc = best * 1e-3 * 1.965e9 * sms * 4 / warp_units;
    printf("{\"tag\": \"%s\", \"op\": \"%s\", \"ms\": %.3f, \"cycles_per_unit_per_smsp\": %.1f}\n", TAG, name, best, cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    gl_t* d; cudaMalloc(&d, (size_t)sms * MINB * 4 * TPB * 8);
    run<0>("sbox", 12, d, sms, ITERS); run<1>("square+fold", 12, d, sms, ITERS * 4); run<2>("product only", 12, d, sms, ITERS * 4);
    run<3>("readout+2 i2f+2 dadd", 12, d, sms, ITERS * 4); run<4>("i2f+lop3", 24, d, sms, ITERS * 4); run<5>("lop3", 24, d, sms, ITERS * 4);
    run<6>("full round", 1, d, sms, ITERS); run<7>("partial pair", 1, d, sms, ITERS); run<8>("permutation", 1, d, sms, 64);
    // digest of the permutation run: equal for every arithmetic variant
    const size_t n = (size_t)sms * MINB * 4 * TPB; gl_t* h = (gl_t*)malloc(n * 8);
    cudaMemcpy(h, d, n * 8, cudaMemcpyDeviceToHost);
    gl_t dig = 0; for (size_t i = 0; i < n; i++) dig = dig * 0x100000001B3ull ^ h[i];
    printf("{\"tag\": \"%s\", \"perm_digest\": \"%016llx\"}\n", TAG, (unsigned long long)dig);
    return cudaDeviceSynchronize() != cudaSuccess;
}
