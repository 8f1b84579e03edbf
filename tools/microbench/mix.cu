// instruction-mix microbenchmark: how IMAD.WIDE.U32 co-issues with IADD3 / carry chains on sm_100a
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define ITERS 2048
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[8], b[8], c[8], d[8]; uint64_t w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x; c[i] = a[i] ^ 0x55; d[i] = b[i] + 9; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i]));                       // IMAD.WIDE RZ acc, independent adds below
            if (OP == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(seed));
            if (OP == 1) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i]));
                           asm volatile("add.u32 %0, %0, %2; add.u32 %1, %1, %2; " : "+r"(c[i]), "+r"(d[i]) : "r"(seed)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed)); } // 1:3
            if (OP == 2) asm volatile("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %2; addc.u32 %3, %3, 0;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(c[i]));   // carry chain of 3
            if (OP == 3) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i]));
                           asm volatile("add.cc.u32 %0, %0, %2; addc.cc.u32 %1, %1, %2; addc.u32 %3, %3, 0;" : "+r"(c[i]), "+r"(d[i]) : "r"(seed), "r"(a[i])); }   // 1 wide + 3-chain
            if (OP == 4) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, %2; }" : "+r"(a[i]) : "r"(b[i]), "r"(seed));          // isetp + predicated add
            if (OP == 5) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(c[i]) : "r"(seed)); } // wide + lop3
            if (OP == 6) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed)); asm volatile("add.u32 %0, %0, %2; add.u32 %1, %1, %2;" : "+r"(c[i]), "+r"(d[i]) : "r"(seed)); } // imad.lo + 2 iadd
            if (OP == 7) { asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[(i+1)&7]) : "r"(c[i]), "r"(d[i])); } // 2 independent wides
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP> void run(const char* name, int sass, uint32_t* d, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d, 12345);
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double groups = (double)blocks * 8 * ITERS * 8;
    double cyc = ms * 1e-3 * 1.965e9 * sms * 4 / groups;
    printf("{\"mix\": \"%s\", \"sass_per_group\": %d, \"cycles_per_group_per_smsp\": %.2f}\n", name, sass, cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
    run<0>("mul.wide + iadd3 (independent regs)", 2, d, sms); run<1>("mul.wide + 3 iadd3", 4, d, sms);
    run<2>("add.cc/addc.cc/addc chain", 3, d, sms); run<3>("mul.wide + 3-carry-chain", 4, d, sms);
    run<4>("isetp + @p iadd", 2, d, sms); run<5>("mul.wide + lop3", 2, d, sms); run<6>("imad.lo + 2 iadd3", 3, d, sms);
    run<7>("2 mul.wide", 2, d, sms);
    return cudaDeviceSynchronize() != cudaSuccess;
}
