// umma_rate_probe.cu — how long does one tcgen05.mma kind::tf32 (M = 128, K = 8) take as a function of N,
// back to back from shared memory operands in the K-major no-swizzle layout of rnnt_umma.cuh?
// One CTA per SM issues R MMAs into the same accumulator, commits, waits; clock64 around it.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/umma_rate_probe tools/probe/umma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../warp-transducer_b200/csrc/rnnt_umma.cuh"

using namespace b200rnnt::umma;

template <int N>
__global__ void __launch_bounds__(128) rate_kernel(int reps, long long* cycles, int ctas_sharing) {
    constexpr int KS = 32;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long done;
    __shared__ uint32_t slot;
    constexpr uint32_t A_BYTES = TileGeom::bytes(128, KS), B_BYTES = TileGeom::bytes(N, KS);
    for (uint32_t i = threadIdx.x; i < (A_BYTES + B_BYTES) / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1.0f;
    constexpr uint32_t COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    if (threadIdx.x < 32) tmem_alloc(s32(&slot), COLS);
    if (threadIdx.x == 0) bar_init(s32(&done), 1);
    fence_smem_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = slot;
    constexpr uint32_t idesc = instr_desc_tf32(128, N, false, false);
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint32_t a = s32(smem), b = s32(smem) + A_BYTES;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int j = 0; j < KS / 8; ++j) {
                const uint64_t ad = smem_desc(a + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                const uint64_t bd = smem_desc(b + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                mma_tf32(tm, ad, bd, idesc, 1);
            }
        }
        mma_commit(s32(&done));
        bar_wait(s32(&done), 0);
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, COLS);
}

template <int N> void run(int ctas_per_sm) {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int reps = 2000;
    const size_t smem = TileGeom::bytes(128, 32) + TileGeom::bytes(N, 32);
    cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long* d;
    const int grid = sms * ctas_per_sm;
    cudaMalloc(&d, grid * sizeof(long long));
    rate_kernel<N><<<grid, 128, smem>>>(reps, d, ctas_per_sm);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("N=%d: %s\n", N, cudaGetErrorString(e));
        exit(1);
    }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (long long c : h) mean += (double)c;
    mean /= grid;
    const double per = mean / (reps * 4.0);
    printf("M=128 N=%3d K=8 tf32, %d CTA/SM: %.1f cycles per MMA per CTA  (%.0f flop/clk/SM aggregate)\n", N, ctas_per_sm, per,
           128.0 * N * 8 * 2 / per * ctas_per_sm);
    cudaFree(d);
}

int main() {
    run<32>(1);
    run<64>(1);
    run<128>(1);
    run<256>(1);
    run<32>(2);
    run<64>(2);
    run<32>(4);
    return 0;
}
