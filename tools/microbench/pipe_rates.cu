// Micro-benchmark: per-SM issue rate of the instructions the int8 GEMM epilogue is made of (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int OP> __global__ void k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8]; uint64_t b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 8 + i; b[i] = ((uint64_t)(0x3f800000u + i) << 32) | (0x3f800000u + threadIdx.x); }
    const uint64_t w = 0x3f8000013f800001ull, c = 0x3400000034000000ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a[i]));                       // I2FP
            if (OP == 1) asm volatile("add.s32 %0, %0, 0x4B400000;" : "+r"(a[i]));                  // IADD / VIADD
            if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(0x3f800001u), "r"(0x34000000u));   // FFMA
            if (OP == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b[i]) : "l"(w), "l"(c));  // FFMA2
            if (OP == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(w));             // FMUL2
            if (OP == 5) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(b[i]) : "l"(c));             // FADD2
            if (OP == 6) asm volatile("mad.lo.s32 %0, %0, 3, 0x4B400000;" : "+r"(a[i]));            // IMAD
            if (OP == 7) { asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(a[i])); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b[i]) : "l"(w), "l"(c)); }  // mix
            if (OP == 8) { asm volatile("add.s32 %0, %0, 0x4B400000;" : "+r"(a[i])); asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(b[i]) : "l"(w), "l"(c)); }
            if (OP == 9) asm volatile("lop3.b32 %0, %0, 0x4B400000, 0x7, 0x96;" : "+r"(a[i]));      // LOP3
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r ^= a[i] ^ (uint32_t)b[i] ^ (uint32_t)(b[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int OP> void run(const char* name, int ops_per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out; cudaMalloc(&out, sms * 1024 * 4);
    const int iters = 20000;
    k<OP><<<sms, 1024>>>(out, 100, 1); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<OP><<<sms, 1024>>>(out, iters, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr_per_sm = (double)iters * 8 * ops_per_iter * 32;     // 32 warps per SM
    double cycles = ms * 1e-3 * clk * 1e3;                               // at the nominal max clock
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)  = %6.1f lanes/clk/SM\n", name, ms, warp_instr_per_sm / cycles, clk / 1000,
           32 * warp_instr_per_sm / cycles);
    cudaFree(out);
}
int main() {
    run<0>("I2FP (cvt.rn.f32.s32)", 1); run<1>("IADD imm (add.s32)", 1); run<9>("LOP3", 1); run<6>("IMAD", 1); run<2>("FFMA", 1);
    run<3>("FFMA2 (fma.rn.f32x2)", 1); run<4>("FMUL2", 1); run<5>("FADD2", 1); run<7>("I2FP + FFMA2 interleaved", 2); run<8>("IADD + FFMA2 interleaved", 2);
    return 0;
}
