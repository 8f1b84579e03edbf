// Issue-rate microbenchmark by operand form on sm_100a: does the number of register operands an instruction reads
// (register-file bandwidth) limit the issue rate, and how do carry chains issue?  Each case is an unrolled stream of
// independent instructions over 16 register sets; 8 warps per SM sub-partition.  cycles = per instruction per SMSP.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a rf.cu -o rf
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define ITERS 4096
#define NR 16
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, uint32_t iters) {
    uint32_t a[NR], b[NR], c[NR], d[NR]; double x[NR], y[NR], z[NR]; uint64_t w[NR];
#pragma unroll
    for (int i = 0; i < NR; i++) {
        a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x + threadIdx.x * 13; c[i] = a[i] ^ 0x55; d[i] = b[i] + 9; w[i] = ((uint64_t)d[i] << 32) | c[i];
        x[i] = (double)a[i]; y[i] = (double)b[i] * 1e-9; z[i] = (double)c[i];
    }
#pragma unroll 1
    for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NR; i++) {
            const int j = (i + 1) % NR, l = (i + 5) % NR;
#define LO(v) ((uint32_t)(v))
#define HI(v) ((uint32_t)((v) >> 32))
#define DF "fma.rn.f64 %0, %1, 0d3FF0000010000000, %0;"
            if (OP == 0) asm volatile("add.u32 %0, %0, 12345;" : "+r"(a[i]));                                            // 1 register operand
            if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[j]));                                   // 2
            if (OP == 2) a[i] = a[i] + a[j] + a[l];                                                                      // one IADD3 with 3 register operands
            if (OP == 3) asm volatile("lop3.b32 %0, %0, 0x5555, 0x3333, 0x96;" : "+r"(a[i]));                            // LOP3, 1 register
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[j]), "r"(a[l]));             // LOP3, 3 registers
            if (OP == 5) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l])));           // IMAD.WIDE, addend RZ
            if (OP == 6) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l])));       // 64-bit addend
            if (OP == 7) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(x[i]) : "d"(x[j]), "d"(y[l]));                 // DFMA, 3 register pairs
            if (OP == 8) asm volatile(DF : "+d"(x[i]) : "d"(x[j]));                                                       // DFMA, 2 pairs + constant
            if (OP == 9) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[j]), "r"(b[l]));        // 64-bit add, 2 instructions
            if (OP == 10) asm volatile("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, 0;" : "+r"(a[i]), "+r"(b[i]), "+r"(c[i]) : "r"(a[j]), "r"(b[l]));  // 3-long chain
            if (OP == 11) asm volatile("sub.cc.u32 %0, %0, %2; subc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[j]), "r"(b[l]));
            if (OP == 12) { w[i] = w[i] + w[j]; }                                                                        // compiler's 64-bit add
            if (OP == 13) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(x[i]) : "r"(a[i])); a[i] = (uint32_t)__double2hiint(x[i]); }   // I2F.F64.U32
            if (OP == 14) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(x[i]) : "d"(x[j]));                               // DADD
            if (OP == 15) { asm volatile(DF : "+d"(x[i]) : "d"(x[j])); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(a[j])); }   // DFMA + IADD3
            if (OP == 16) { asm volatile(DF : "+d"(x[i]) : "d"(x[j])); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[j]), "r"(b[l])); }   // DFMA + 64-bit add
            if (OP == 17) { asm volatile(DF : "+d"(x[i]) : "d"(x[j])); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l]))); }   // DFMA + IMAD.WIDE
            if (OP == 18) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l]))); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[j]), "r"(b[l])); }   // IMAD.WIDE + 64-bit add
            if (OP == 19) { asm volatile(DF : "+d"(x[i]) : "d"(x[j])); asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l]))); asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[j]), "r"(b[l])); }   // all three
            if (OP == 20) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l]))); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[j]), "r"(a[l])); }   // IMAD.WIDE + LOP3
            if (OP == 21) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(HI(w[l]))); a[i] = a[i] + a[j] + a[l]; }   // IMAD.WIDE + IADD3
            if (OP == 22) { asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(b[j]), "r"(b[l])); a[i] = a[i] + a[j] + a[l]; }   // LOP3 + IADD3
            if (OP == 24) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));               // R x R + R64, invariant multiplicands
            if (OP == 25) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(seed + i));           // R x UR + R64
            if (OP == 26) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(LO(w[j])), "r"(b[i]));           // varying multiplicand
            if (OP == 27) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                 // IMAD R,R,R
            if (OP == 28) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile(DF : "+d"(x[i]) : "d"(x[j])); }   // + DFMA
            if (OP == 29) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); c[i] = c[i] + c[j] + d[l]; }   // + IADD3 (3 operands)
            if (OP == 30) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(c[j])); }   // + IADD3 (2 operands)
            if (OP == 31) w[i] = (uint64_t)LO(w[j]) * HI(w[l]) + w[i];                                                    // compiler's fused multiply-add
            if (OP == 32) w[i] = (uint64_t)LO(w[j]) * HI(w[l]);                                                           // compiler's wide multiply
            if (OP == 33) { w[i] = (uint64_t)LO(w[j]) * HI(w[l]) + w[i]; asm volatile(DF : "+d"(x[i]) : "d"(x[j])); }
            if (OP == 34) { w[i] = (uint64_t)LO(w[j]) * HI(w[l]) + w[i]; a[i] = a[i] + a[j] + b[l]; }
            if (OP == 35) { w[i] = (uint64_t)LO(w[j]) * HI(w[l]) + w[i]; a[i] = a[i] + a[j]; }
            if (OP == 36) { w[i] = (uint64_t)LO(w[j]) * HI(w[l]) + w[i]; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[j]), "r"(a[l])); }
            if (OP == 23) { asm volatile(DF : "+d"(x[i]) : "d"(x[j])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(a[j]), "r"(a[l])); }   // DFMA + LOP3
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NR; i++) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(x[i]) ^ (uint32_t)__double2hiint(x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP> void run(const char* name, int sass, uint32_t* d, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 4;                      // 4 blocks of 256 threads per SM = 8 warps per SMSP
    k<OP><<<blocks, 256>>>(d, 12345, 64);
    cudaEventRecord(e0); k<OP><<<blocks, 256>>>(d, 12345, ITERS); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double groups = (double)blocks * 8 * ITERS * NR;
    const double cyc = ms * 1e-3 * 1.965e9 * sms * 4 / groups;
    printf("{\"case\": \"%s\", \"sass_per_group\": %d, \"cycles_per_group\": %.2f, \"cycles_per_instruction\": %.2f}\n", name, sass, cyc, cyc / sass);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); const int sms = p.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 4 * 256 * 4);
    run<0>("iadd3 r,imm", 1, d, sms); run<1>("iadd3 r,r", 1, d, sms); run<2>("iadd3 r,r,r", 1, d, sms);
    run<3>("lop3 r,imm,imm", 1, d, sms); run<4>("lop3 r,r,r", 1, d, sms);
    run<5>("imad.wide r,r,RZ", 1, d, sms); run<6>("imad.wide r,r,r64", 1, d, sms);
    run<7>("dfma r,r,r", 1, d, sms); run<8>("dfma r,const,r", 1, d, sms);
    run<9>("add.cc+addc", 2, d, sms); run<10>("add.cc+addc.cc+addc", 3, d, sms); run<11>("sub.cc+subc", 2, d, sms);
    run<12>("64-bit add (compiler)", 2, d, sms); run<13>("i2f.f64.u32", 1, d, sms); run<14>("dadd", 1, d, sms);
    run<15>("dfma + iadd3", 2, d, sms); run<16>("dfma + add64", 3, d, sms); run<17>("dfma + imad.wide", 2, d, sms);
    run<18>("imad.wide + add64", 3, d, sms); run<19>("dfma + imad.wide + add64", 4, d, sms);
    run<20>("imad.wide + lop3", 2, d, sms); run<21>("imad.wide + iadd3", 2, d, sms); run<22>("lop3 + iadd3", 2, d, sms); run<23>("dfma + lop3", 2, d, sms);
    run<31>("imad.wide fused R,R,R64", 1, d, sms); run<32>("imad.wide R,R,RZ (C)", 1, d, sms); run<33>("imad.wide fused + dfma", 2, d, sms);
    run<34>("imad.wide fused + iadd3(3)", 2, d, sms); run<35>("imad.wide fused + iadd3(2)", 2, d, sms); run<36>("imad.wide fused + lop3", 2, d, sms);
    run<24>("imad.wide RxR+R64 (invariant a,b)", 1, d, sms); run<25>("imad.wide RxUR+R64", 1, d, sms); run<26>("imad.wide RxR+R64 (varying a)", 1, d, sms);
    run<27>("imad R,R,R", 1, d, sms); run<28>("imad.wide + dfma", 2, d, sms); run<29>("imad.wide + iadd3(3)", 2, d, sms); run<30>("imad.wide + iadd3(2)", 2, d, sms);
    return cudaDeviceSynchronize() != cudaSuccess;
}
