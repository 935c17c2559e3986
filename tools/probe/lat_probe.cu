// Dependent-chain latency probe (dev tool): cycles per op for the instructions on the lattice chain.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <int OP> __global__ void k(double* out, long long* cyc, double seed, float fseed) {
    double a = seed + threadIdx.x * 1e-9, b = 1.0000001;
    float f = fseed + threadIdx.x * 1e-6f, g = 1.0001f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) a = a + b;                                   // DADD
        if (OP == 1) a = fma(a, b, b);                            // DFMA
        if (OP == 2) { f = (float)a; a = (double)f + 0.0; }        // F2F.F32.F64 + F2F.F64.F32 (+DADD folded?)
        if (OP == 3) f = f + g;                                   // FADD
        if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
        if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
        if (OP == 6) f = __shfl_up_sync(0xffffffffu, f, 1);
        if (OP == 7) a = __shfl_up_sync(0xffffffffu, a, 1);
        if (OP == 8) { asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(a)); asm volatile("cvt.f64.f32 %0, %1;" : "=d"(a) : "f"(f)); }  // both conversions
        if (OP == 12) { float h = f + g; float e = h - f; float lo = g - e; f = h + lo; }  // fast two-sum chain (3 dependent FADD)
        if (OP == 9) a = fmax(a, b) + 1e-9;                       // DSETP/select + DADD
        if (OP == 10) f = fmaf(f, g, g);
        if (OP == 11) { asm volatile("bar.sync 0;"); f += 1.f; }
    }
    long long t1 = clock64();
    out[threadIdx.x] = a + f;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8192); cudaMallocManaged(&cyc, 8);
    const char* names[] = {"DADD", "DFMA", "F2F.f32<-f64 + F2F.f64<-f32 + DADD", "FADD", "MUFU.EX2", "MUFU.LG2", "SHFL f32", "SHFL f64 (2x32)", "F2F both ways", "fmax(double)+DADD", "FFMA", "BAR.SYNC(320 thr)+FADD", "two-sum (4 FADD)"};
#define RUN(OP, TH) k<OP><<<1, TH>>>(out, cyc, 1.5, 0.5f); cudaDeviceSynchronize(); k<OP><<<1, TH>>>(out, cyc, 1.5, 0.5f); cudaDeviceSynchronize(); printf("%-40s %6.1f cycles/iter (%d threads)\n", names[OP], (double)cyc[0] / N, TH);
    RUN(0, 32) RUN(1, 32) RUN(2, 32) RUN(3, 32) RUN(4, 32) RUN(5, 32) RUN(6, 32) RUN(7, 32) RUN(8, 32) RUN(9, 32) RUN(10, 32) RUN(11, 320) RUN(11, 64)
    RUN(0, 320) RUN(8, 320) RUN(12, 32)
    return 0;
}
