// Micro-benchmarks behind DESIGN.md's cost model: does an FP64 (or packed f32x2) warp-instruction leave the SMSP's
// issue port free for other instructions while it occupies its pipe?   nvcc -arch=sm_100a -O3 issue_model.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
template <int MODE> __global__ void __launch_bounds__(256) k(float* sink, unsigned long long iters) {
    double d[CHAINS]; float f[CHAINS]; unsigned u[CHAINS]; unsigned long long p[CHAINS];
    for (int i = 0; i < CHAINS; i++) { d[i] = threadIdx.x * 1e-3 + i; f[i] = threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; p[i] = ((unsigned long long)__float_as_uint(f[i]) << 32) | __float_as_uint(f[i] + 0.5f); }
    const double xd = 0.999999, yd = 1e-6; const float xf = 0.999999f, yf = 1e-6f;
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
            }
        }
    }
    double s = 0; for (int i = 0; i < CHAINS; i++) s += d[i] + f[i] + u[i] + (double)p[i];
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
    run<3>("FFMA", sink, 1); run<4>("FFMA + LOP3", sink, 2); run<5>("FFMA2 (f32x2)", sink, 1); run<6>("FFMA2 + LOP3", sink, 2);
    return 0;
}
