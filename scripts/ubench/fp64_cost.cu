// fp64_cost.cu — what do the building blocks of the f64 march cost on B200, in SMSP cycles per warp, as a function of
// resident warps per scheduler? Blocks: trig_pair (16 DFMA + 3 DMUL + 1 DADD), rhs_ks_u, the whole implicit-midpoint
// step, and synthetic chains that isolate DFMA forms (register / uniform / immediate operands).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../blackhole-simulation_b200/csrc fp64_cost.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "gvt_device.cuh"
using namespace gvt;
struct P { TrigTable trig; double M, a, k0, k1, k2, k3; };

template <int MODE> __global__ void k(const __grid_constant__ P p, float* sink, int iters) {
    HoleRay<double> c; c.trig = &p.trig; c.set_hole(p.M, p.a); c.set_ray(-1.0, 2.0 + 1e-3 * threadIdx.x);
    Ray<double> y; y.t = 0; y.ph = 0; y.r = 20.0 + 1e-3 * threadIdx.x; y.th = 1.2 + 1e-4 * threadIdx.x; y.pr = 0.5; y.pth = 0.3;
    double acc = 0.0, z = 0.3 + 1e-4 * threadIdx.x, w = 1.1, v = 0.7 + 1e-5 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { double a, sc; trig_pair(p.trig, y.th, a, sc); y.th = fma(a, 1e-9, y.th); acc += sc; }
        if (MODE == 1) { DerivU<double> d = rhs_ks_u<double, false, false>(c, y.r, z, w, y.pr, y.pth); y.r = fma(d.dr, 1e-9, y.r); y.pr = fma(d.dpr, 1e-9, y.pr); y.pth = fma(d.dpth, 1e-9, y.pth); acc += d.isig; }
        if (MODE == 2) { step_symplectic<double, 1, false>(c, y, 1e-3); }
        if (MODE == 3) {  // 16 dependent DFMAs: p = p*z + K (uniform-register coefficients), two chains
#pragma unroll
            for (int i = 0; i < 8; i++) { z = fma(z, v, p.k0); w = fma(w, v, p.k1); }
        }
        if (MODE == 4) {  // 16 DFMAs with three distinct register operands, two chains
#pragma unroll
            for (int i = 0; i < 8; i++) { z = fma(z, v, w); w = fma(w, v, z); }
        }
        if (MODE == 5) {  // 16 DFMAs with immediates
#pragma unroll
            for (int i = 0; i < 8; i++) { z = fma(z, 0.999999, 1e-7); w = fma(w, 0.999998, 2e-7); }
        }
        if (MODE == 6) {  // 16 independent-ish DMULs
#pragma unroll
            for (int i = 0; i < 8; i++) { z = z * v; w = w * v; }
        }
        if (MODE == 7) {  // 4 chains of DFMA with uniform coefficient
#pragma unroll
            for (int i = 0; i < 4; i++) { z = fma(z, v, p.k0); w = fma(w, v, p.k1); acc = fma(acc, v, p.k2); y.r = fma(y.r, v, p.k3); }
        }
    }
    const double s = acc + y.r + y.th + y.pr + y.pth + z + w;
    if (s == -1.2345) sink[0] = (float)s;
}
template <int MODE> void run(const char* name, const P& p, float* sink, double fp64_ops, int wps) {
    const int iters = 4000, threads = 128 * wps, blocks = 148;       // one CTA per SM, wps warps per scheduler
    k<MODE><<<blocks, threads>>>(p, sink, 50); cudaDeviceSynchronize();
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a); k<MODE><<<blocks, threads>>>(p, sink, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double cyc = ms * 1e-3 * 1.965e9;                           // SM cycles (boost clock)
    const double per_warp_iter = cyc / iters / wps;                  // SMSP cycles per warp-iteration
    printf("%-34s warps/SMSP %d: %8.1f cycles per warp-iteration  (%.2f per FP64 op, latency per iteration %.0f)\n", name, wps, per_warp_iter,
           per_warp_iter / fp64_ops, cyc / iters);
}
int main() {
    float* sink; cudaMalloc(&sink, 256);
    P p; const TrigTable tt = GVT_TRIG_TABLE_INIT; p.trig = tt; p.M = 1.0; p.a = 0.999; p.k0 = 1e-7; p.k1 = 2e-7; p.k2 = 3e-7; p.k3 = 4e-7;
    for (int wps : {1, 2, 4, 6, 8}) {
        run<0>("trig_pair (20 ops)", p, sink, 20 + 2, wps);
        run<1>("rhs_ks_u (28 ops)", p, sink, 28 + 4, wps);
        run<2>("step_symplectic (~160 ops)", p, sink, 160, wps);
        run<3>("16 DFMA reg,reg,uniform 2 chains", p, sink, 16, wps);
        run<7>("16 DFMA reg,reg,uniform 4 chains", p, sink, 16, wps);
        run<4>("16 DFMA 3 regs 2 chains", p, sink, 16, wps);
        run<5>("16 DFMA imm 2 chains", p, sink, 16, wps);
        run<6>("16 DMUL 2 chains", p, sink, 16, wps);
    }
    return 0;
}
