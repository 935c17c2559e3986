// umma_probe.cu — checks warp-transducer_b200/csrc/rnnt_umma.cuh (tcgen05 tf32x3 GEMM) against a double
// precision CPU product for the operand-layout combinations the additive-joint kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/umma_probe tools/probe/umma_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../warp-transducer_b200/csrc/rnnt_umma.cuh"

using namespace b200rnnt::umma;

template <bool A_MN, bool B_MN, int NT, int KS, int AMODE = (A_MN ? 0 : 1), int BMODE = (B_MN ? 0 : 1)>
double run_case(const char* name, int batch, int M, int Nn, int K, int slices) {
    // logical A[M][K], B[Nn][K]; stored K-major ([mn][k]) or MN-major ([k][mn])
    std::vector<float> hA((size_t)batch * M * K), hB((size_t)batch * Nn * K);
    srand(7);
    for (auto& x : hA) x = (float)rand() / RAND_MAX;
    for (auto& x : hB) x = (float)rand() / RAND_MAX;
    float *dA, *dB, *dO;
    cudaMalloc(&dA, hA.size() * 4);
    cudaMalloc(&dB, hB.size() * 4);
    const size_t no = (size_t)batch * slices * M * Nn;
    cudaMalloc(&dO, no * 4);
    cudaMemset(dO, 0xff, no * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    Operand A{dA, (long long)M * K, A_MN ? 1 : K, A_MN ? M : 1, M};
    Operand B{dB, (long long)Nn * K, B_MN ? 1 : K, B_MN ? Nn : 1, Nn};
    auto kern = gemm_kernel<AMODE, BMODE, NT, KS>;
    const size_t smem = gemm_smem_bytes<NT, KS>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(slices * ((Nn + NT - 1) / NT), (M + 127) / 128, batch);
    kern<<<grid, kThreads, smem>>>(A, B, K, slices, Epilogue{nullptr, 0, 0, 0, dO, (long long)M * Nn, (long long)slices * M * Nn, Nn, 1});
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%-28s: CUDA error %s\n", name, cudaGetErrorString(e));
        exit(1);
    }
    std::vector<float> hO(no);
    cudaMemcpy(hO.data(), dO, no * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int b = 0; b < batch; ++b)
        for (int m = 0; m < M; ++m)
            for (int n = 0; n < Nn; ++n) {
                double ref = 0, got = 0;
                for (int k = 0; k < K; ++k) {
                    const double a = A_MN ? hA[((size_t)b * K + k) * M + m] : hA[((size_t)b * M + m) * K + k];
                    const double bb = B_MN ? hB[((size_t)b * K + k) * Nn + n] : hB[((size_t)b * Nn + n) * K + k];
                    ref += a * bb;
                }
                for (int s = 0; s < slices; ++s) got += hO[(((size_t)b * slices + s) * M + m) * Nn + n];
                const double err = fabs(got - ref) / fmax(fabs(ref), 1e-30);
                if (!(err <= worst)) worst = std::isnan(err) ? 1e30 : err;
            }
    // the tensor core accumulates in fp32 rounding toward zero: with all-positive operands the error grows
    // linearly with the number of MMAs into one accumulator (3 per 8 k values), 2^-24 each
    const double mmas = 3.0 * ((K + slices - 1) / slices + 7) / 8;
    const double tol = 1e-5 + 1.5 * mmas * 5.96e-8;
    printf("%-28s: batch %d M %d N %d K %d slices %d smem %zu B -> max rel err %.3e (tol %.1e, %.0f MMAs/accumulator) %s\n", name, batch,
           M, Nn, K, slices, smem, worst, tol, mmas, worst < tol ? "OK" : "MISMATCH");
    cudaFree(dA), cudaFree(dB), cudaFree(dO);
    return worst;
}

int main() {
    run_case<false, false, 32, 24>("S  (A k-contig scalar, B k-contig scalar)", 2, 150, 21, 1000, 3);
    run_case<true, false, 160, 24>("dF (A mn-contig, B k-contig)", 2, 200, 150, 21, 1);
    run_case<true, true, 32, 24>("dG (A mn-contig, B mn-contig)", 2, 200, 21, 150, 1);
    run_case<false, false, 32, 32, 2, 2>("S  full size, float4 fetch", 4, 150, 21, 5000, 8);
    run_case<false, false, 64, 32, 2, 2>("S  N=64, float4 fetch", 2, 70, 66, 1000, 3);
    run_case<true, true, 64, 24>("dG N=64", 2, 200, 66, 70, 1);
    run_case<true, false, 64, 24>("dF N=64 tiled", 2, 200, 150, 21, 1);
    run_case<false, false, 32, 32, 2, 2>("S  one stage", 2, 150, 21, 24, 1);
    run_case<false, false, 32, 32, 2, 2>("S  two stages", 2, 150, 21, 60, 1);
    run_case<true, false, 192, 24>("dF N=192", 2, 200, 150, 21, 1);
    run_case<true, false, 256, 24>("dF N=256 tiled", 1, 200, 300, 40, 1);
    run_case<true, false, 64, 24, 3, 1>("dF float4 transposing fetch", 2, 200, 150, 21, 1);
    run_case<true, true, 32, 24, 3, 0>("dG float4 transposing fetch", 2, 200, 21, 150, 1);
    run_case<true, false, 192, 24, 3, 1>("dF N=192 float4 transposing", 2, 328, 150, 21, 1);
    return 0;
}
