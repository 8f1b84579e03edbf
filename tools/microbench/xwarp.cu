// Cross-warp pipe sharing on sm_100a: warps of one SM sub-partition run DIFFERENT pure instruction
// streams (role = warp index mod 2).  If two instruction classes use separate pipes the mixed run
// takes max(tA, tB); if they share one it takes tA + tB.
// Classes: 0 = DFMA, 1 = IMAD.WIDE.U32, 2 = carry chain (IADD3 / IADD3.X), 3 = LOP3, 4 = plain IADD3,
// 5 = IMAD (32-bit low), 6 = IMAD.HI, 7 = ISETP + SEL
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#define ITERS 4096
template <int OP> __device__ __forceinline__ void body(uint32_t (&a)[8], uint32_t (&b)[8], uint64_t (&w)[8], double (&f)[8], uint32_t seed) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(1.0000001), "d"(0.5));
        if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));
        if (OP == 2) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %2;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed));
        if (OP == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
        if (OP == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed));
        if (OP == 5) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
        if (OP == 6) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
        if (OP == 7) asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %2, p; }" : "+r"(a[i]) : "r"(b[i]), "r"(seed));
    }
}
template <int OPA, int OPB>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[8], b[8]; uint64_t w[8]; double f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i + blockIdx.x; w[i] = a[i]; f[i] = 1.0 + i; }
    const bool roleA = ((threadIdx.x >> 5) & 1) == 0;
    if (roleA) { for (int it = 0; it < ITERS; it++) body<OPA>(a, b, w, f, seed); }
    else       { for (int it = 0; it < ITERS; it++) body<OPB>(a, b, w, f, seed); }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) acc ^= a[i] ^ b[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)__double2loint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OPA, int OPB> void run(const char* name, uint32_t* d, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int blocks = sms * 4;               // 4 x 8 warps = 8 warps per sub-partition, 4 of each role
    k<OPA, OPB><<<blocks, 256>>>(d, 12345);
    cudaEventRecord(e0); k<OPA, OPB><<<blocks, 256>>>(d, 12345); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    // per sub-partition: 4 warps of each role, each issuing ITERS*8 ops (carry chain = 2 SASS per op)
    double cyc_per_smsp = ms * 1e-3 * 1.965e9;
    printf("{\"roles\": \"%s\", \"ms\": %.3f, \"cycles_per_op_pair\": %.2f}\n", name, ms, cyc_per_smsp / (4.0 * ITERS * 8));
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); int sms = p.multiProcessorCount;
    uint32_t* d; cudaMalloc(&d, (size_t)sms * 4 * 256 * 4);
    run<0, 0>("dfma | dfma", d, sms); run<1, 1>("imad.wide | imad.wide", d, sms); run<2, 2>("carry2 | carry2", d, sms);
    run<3, 3>("lop3 | lop3", d, sms); run<4, 4>("iadd3 | iadd3", d, sms);
    run<0, 1>("dfma | imad.wide", d, sms); run<0, 2>("dfma | carry2", d, sms); run<0, 3>("dfma | lop3", d, sms);
    run<1, 2>("imad.wide | carry2", d, sms); run<1, 3>("imad.wide | lop3", d, sms); run<1, 4>("imad.wide | iadd3", d, sms);
    run<2, 3>("carry2 | lop3", d, sms); run<0, 4>("dfma | iadd3", d, sms);
    run<5, 5>("imad.lo | imad.lo", d, sms); run<6, 6>("imad.hi | imad.hi", d, sms); run<7, 7>("isetp+sel | isetp+sel", d, sms);
    run<5, 3>("imad.lo | lop3", d, sms); run<5, 2>("imad.lo | carry2", d, sms); run<6, 3>("imad.hi | lop3", d, sms);
    run<5, 1>("imad.lo | imad.wide", d, sms); run<5, 0>("imad.lo | dfma", d, sms); run<6, 0>("imad.hi | dfma", d, sms);
    return cudaDeviceSynchronize() != cudaSuccess;
}
