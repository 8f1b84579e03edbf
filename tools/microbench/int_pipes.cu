// Integer-pipe throughput microbenchmark for B200 (sm_100a): the roofline denominators for
// the Goldilocks kernels (Poseidon, NTT butterflies, quotient evaluation).
// Each kernel runs ILP independent dependency chains per thread; reports warp-instructions
// per cycle per SM sub-partition (SMSP) = total_warp_instr / (elapsed_cycles * 4 * SMs).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[ILP], b[ILP]; uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x; w[i] = ((uint64_t)a[i] << 32) | b[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));                 // IMAD
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));              // IMAD.WIDE.U32
            if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));                 // IMAD.HI
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                                   // IADD3
            if (OP == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed)); // IADD3 + IADD3.X
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));             // LOP3
            if (OP == 6) asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(a[i]) : "r"(b[i]));                         // SHF
            if (OP == 7) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a[i]) : "r"(b[i]), "r"(seed)); // ISETP+SEL
            if (OP == 8) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % ILP] | 1));                // 64-bit mul.lo
            if (OP == 9) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(0xFFFFFFFF00000001ULL));               // 64-bit mul.hi
            if (OP == 10) { double d = __longlong_as_double(w[i]); asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(1.0000001)); w[i] = __double_as_longlong(d); } // DFMA
            if (OP == 11) asm volatile("mad.wide.u32 %0, %1, %2, %0; add.u32 %1, %1, %3;" : "+l"(w[i]), "+r"(a[i]) : "r"(b[i]), "r"(seed)); // IMAD.WIDE + IADD3 mix (dual pipe)
            if (OP == 12) asm volatile("mad.lo.u32 %0, %0, %2, %3; add.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed)); // IMAD + IADD3 mix
            if (OP == 13) asm volatile("add.cc.u64 %0, %0, %1; " : "+l"(w[i]) : "l"((uint64_t)seed));                    // 64-bit add (2 SASS)
            if (OP == 15) { double d = __longlong_as_double(w[(i + 4) % ILP]); asm volatile("mad.wide.u32 %0, %2, %3, %0; fma.rn.f64 %1, %1, %4, %4;" : "+l"(w[i]), "+d"(d) : "r"(a[i]), "r"(b[i]), "d"(1.0000001)); w[(i + 4) % ILP] = __double_as_longlong(d); }
            if (OP == 16) { double d = __longlong_as_double(w[i]); asm volatile("fma.rn.f64 %0, %0, %2, %2; add.u32 %1, %1, %3;" : "+d"(d), "+r"(a[i]) : "d"(1.0000001), "r"(seed)); w[i] = __double_as_longlong(d); }
            if (OP == 17) { double d = __longlong_as_double(w[(i + 4) % ILP]); asm volatile("mad.wide.u32 %0, %2, %3, %0; fma.rn.f64 %1, %1, %4, %4; add.u32 %2, %2, %5; add.u32 %3, %3, %5;" : "+l"(w[i]), "+d"(d), "+r"(a[i]), "+r"(b[i]) : "d"(1.0000001), "r"(seed)); w[(i + 4) % ILP] = __double_as_longlong(d); }
            if (OP == 18) asm volatile("mad.wide.u32 %0, %1, 41, %0;" : "+l"(w[i]) : "r"(a[i]));
            if (OP == 14) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed + 1)); // mad.cc chain
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, int sass_per_op, uint32_t* d_out, int sms, double clock_ghz_hint) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 8;
    k<OP><<<blocks, 256>>>(d_out, 12345);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d_out, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_ops = (double)blocks * 8 * ITERS * ILP;     // 8 warps per block
    double ops_per_s = warp_ops / (ms * 1e-3);
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"warp_ops_per_s\": %.4e, \"warp_ops_per_ns_per_smsp\": %.4f, \"sass_per_op\": %d}\n",
           name, ms, ops_per_s, ops_per_s / 1e9 / (sms * 4), sass_per_op);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
    run<0>("imad.lo", 1, d, sms, 0); run<1>("imad.wide.u32", 1, d, sms, 0); run<2>("imad.hi", 1, d, sms, 0);
    run<3>("iadd3", 1, d, sms, 0); run<4>("iadd3+iadd3.x", 2, d, sms, 0); run<5>("lop3", 1, d, sms, 0);
    run<6>("shf", 1, d, sms, 0); run<7>("isetp+sel", 2, d, sms, 0); run<8>("mul.lo.u64", 0, d, sms, 0);
    run<9>("mul.hi.u64", 0, d, sms, 0); run<10>("dfma", 1, d, sms, 0); run<11>("imad.wide+iadd3", 2, d, sms, 0);
    run<12>("imad.lo+iadd3", 2, d, sms, 0); run<13>("add.u64", 2, d, sms, 0); run<14>("mad.lo.cc+madc.hi", 2, d, sms, 0);
    run<15>("imad.wide+dfma", 2, d, sms, 0); run<16>("dfma+iadd3", 2, d, sms, 0); run<17>("imad.wide+dfma+2iadd3", 4, d, sms, 0); run<18>("imad.wide.imm", 1, d, sms, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
