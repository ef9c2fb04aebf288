// Micro-benchmarks behind DESIGN.md's cost model: does an FP64 (or packed f32x2) warp-instruction leave the SMSP's
// issue port free for other instructions while it occupies its pipe?   nvcc -arch=sm_100a -O3 issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
template <int MODE> __global__ void __launch_bounds__(256) k(float* sink, unsigned long long iters) {
    double d[CHAINS], e[CHAINS], g[CHAINS]; float f[CHAINS]; unsigned u[CHAINS]; unsigned long long p[CHAINS];
    for (int i = 0; i < CHAINS; i++) { e[i] = 0.999 + 1e-6 * (threadIdx.x + i); g[i] = 1e-7 * (threadIdx.x + 3 * i); d[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; p[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 0.5f); }
    double xd2 = 0.5; const double xd = 0.999999, yd = 1e-6; const float xf = 0.999999f, yf = 1e-6f;
    const unsigned long long xp = ((unsigned long long)__float_as_uint(xf) << 32) | __float_as_uint(xf), yp = ((unsigned long long)__float_as_uint(yf) << 32) | __float_as_uint(yf);
    for (unsigned long long it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (MODE == 0 || MODE == 1) d[i] = fma(d[i], xd, yd);                                   // DFMA
                if (MODE == 1 || MODE == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(r)); // ALU
                if (MODE == 3 || MODE == 4) f[i] = fmaf(f[i], xf, yf);                                  // FFMA
                if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(r));
                if (MODE == 5 || MODE == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xp), "l"(yp));  // FFMA2
                if (MODE == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(u[(i + 1) % CHAINS]), "r"(r));
                if (MODE == 12) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(g[i]));      // DFMA, 3 distinct register pairs
                if (MODE == 13) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(yd));        // DFMA, 2 distinct + constant
                if (MODE == 14) asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(d[i]) : "d"(e[i]));                  // DFMA, a*a+c
                if (MODE == 15) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(e[i]));                     // DMUL 2 distinct
                if (MODE == 7) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(xd));                      // DMUL
                if (MODE == 8) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(yd));                      // DADD
                if (MODE == 9) asm volatile("{.reg .pred p; setp.lt.f64 p, %0, %1; selp.b32 %2, %2, %3, p;}" : "+d"(d[i]), "+d"(xd2), "+r"(u[i]) : "r"(r));  // DSETP + SEL
                if (MODE == 10) { d[i] = fma(d[i], xd, yd); asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.f32 %2, %2, %3, p;}" : "+r"(u[i]) : "r"(r), "f"(f[i]), "f"(xf)); }  // DFMA + ISETP + FSEL
                if (MODE == 11) { d[i] = fma(d[i], xd, yd); d[(i + 1) % CHAINS] = fma(d[(i + 1) % CHAINS], xd, yd); f[i] = fmaf(f[i], xf, yf); }   // 2 DFMA + 1 FFMA
            }
        }
    }
    double s = xd2; for (int i = 0; i < CHAINS; i++) s += e[i] + g[i] + d[i] + f[i] + u[i] + (double)p[i];
    if (s == -1.2345) sink[0] = (float)s;
}
template <int MODE> void run(const char* name, float* sink, double ops_per_inner) {
    const unsigned long long iters = 20000; const int blocks = 148 * 8, threads = 256;
    k<MODE><<<blocks, threads>>>(sink, 200); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(sink, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double inner = (double)blocks * threads * iters * 8 * CHAINS;   // thread-level inner iterations
    double warp_inner_per_smsp_cycle = inner / 32.0 / (148 * 4) / (ms * 1e-3 * 1.965e9);
    printf("%-28s %8.3f ms   %7.2f G inner/s   cycles per warp-inner per SMSP = %.3f\n", name, ms, inner / ms * 1e-6, 1.0 / warp_inner_per_smsp_cycle);
}
int main() {
    float* sink; cudaMalloc(&sink, 256);
    run<0>("DFMA", sink, 1); run<2>("LOP3", sink, 1); run<1>("DFMA + LOP3", sink, 2);
    run<12>("DFMA 3 distinct regs", sink, 1); run<13>("DFMA 2 regs + const", sink, 1); run<14>("DFMA a*a+c (2 regs)", sink, 1); run<15>("DMUL 2 distinct regs", sink, 1);
    run<7>("DMUL", sink, 1); run<8>("DADD", sink, 1); run<9>("DSETP + SEL", sink, 2); run<10>("DFMA + ISETP + FSEL", sink, 3); run<11>("2 DFMA + FFMA", sink, 3);
    run<3>("FFMA", sink, 1); run<4>("FFMA + LOP3", sink, 2); run<5>("FFMA2 (f32x2)", sink, 1); run<6>("FFMA2 + LOP3", sink, 2);
    return 0;
}
