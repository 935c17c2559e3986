// umma_layout_probe.cu — discovers where tcgen05.mma (kind::tf32, no swizzle) reads element (mn,k) of an
// MN-major operand.  The operand image in shared memory holds its own WORD INDEX at every word; the other
// operand is a one-hot selector, so D reveals the address map.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe/umma_layout_probe tools/probe/umma_layout_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../warp-transducer_b200/csrc/rnnt_umma.cuh"
using namespace b200rnnt::umma;

// one MMA (K = 8): A image / B image are raw byte images copied to smem; descriptors built from params
__global__ void __launch_bounds__(128)
one_mma(const float* a_img, int a_words, const float* b_img, int b_words, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo,
        uint32_t b_sbo, int a_mn, int b_mn, int N, float* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* sa = reinterpret_cast<float*>(smem);
    float* sb = reinterpret_cast<float*>(smem + 32768);
    __shared__ __align__(8) unsigned long long bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < a_words; i += 128) sa[i] = a_img[i];
    for (int i = threadIdx.x; i < b_words; i += 128) sb[i] = b_img[i];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) tmem_alloc(s32(&slot), 32);
    if (threadIdx.x == 0) bar_init(s32(&bar), 1);
    fence_smem_async();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t td = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = instr_desc_tf32(128, N, a_mn != 0, b_mn != 0);
        mma_tf32(td, smem_desc(s32(sa), a_lbo, a_sbo), smem_desc(s32(sb), b_lbo, b_sbo), idesc, 0);
        mma_commit(s32(&bar));
    }
    bar_wait(s32(&bar), 0);
    fence_after();
    float v[16];
    tmem_ld16(td + ((uint32_t)(warp * 32) << 16), v);
    for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = v[i];
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(td, 32);
}

static void run(const char* title, const std::vector<float>& a, const std::vector<float>& b, uint32_t a_lbo, uint32_t a_sbo,
                uint32_t b_lbo, uint32_t b_sbo, int a_mn, int b_mn, bool show_rows) {
    float *da, *db, *dout;
    cudaMalloc(&da, a.size() * 4), cudaMalloc(&db, b.size() * 4), cudaMalloc(&dout, 128 * 16 * 4);
    cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(one_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    one_mma<<<1, 128, 65536>>>(da, (int)a.size(), db, (int)b.size(), a_lbo, a_sbo, b_lbo, b_sbo, a_mn, b_mn, 16, dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("== %s  (A: lbo %u sbo %u %s | B: lbo %u sbo %u %s) -> %s\n", title, a_lbo, a_sbo, a_mn ? "MN" : "K", b_lbo, b_sbo,
           b_mn ? "MN" : "K", cudaGetErrorString(e));
    if (e != cudaSuccess) exit(1);
    std::vector<float> o(128 * 16);
    cudaMemcpy(o.data(), dout, o.size() * 4, cudaMemcpyDeviceToHost);
    if (show_rows) {
        const int rows[] = {0, 1, 2, 3, 4, 5, 7, 8, 9, 12, 16, 31, 32, 33, 64, 127};
        for (int m : rows) {
            printf("   m=%3d :", m);
            for (int n = 0; n < 8; ++n) printf(" %6.0f", o[m * 16 + n]);
            printf("\n");
        }
    } else {
        for (int m = 0; m < 8; ++m) {
            printf("   k=%d :", m);
            for (int n = 0; n < 16; ++n) printf(" %6.0f", o[m * 16 + n]);
            printf("\n");
        }
    }
    cudaFree(da), cudaFree(db), cudaFree(dout);
}

int main() {
    // known-good K-major no-swizzle one-hot: element (mn,k) at (mn%8)*16 + (mn/8)*SBO + (k/4)*LBO + (k%4)*4, LBO 128, SBO 256
    auto onehot_k = [](int MN) {
        std::vector<float> img(2048, 0.0f);
        for (int mn = 0; mn < MN; ++mn)
            for (int k = 0; k < 8; ++k)
                img[((mn % 8) * 16 + (mn / 8) * 256 + (k / 4) * 128 + (k % 4) * 4) / 4] = (mn == k) ? 1.0f : 0.0f;
        return img;
    };
    std::vector<float> idx(2048);
    for (int i = 0; i < 2048; ++i) idx[i] = (float)i;   // exact in tf32 (11-bit significand)

    // sanity: A K-major index image, B one-hot -> D[m][k] = word index of A(m,k)
    run("A K-major discovery", idx, onehot_k(16), 128, 256, 128, 256, 0, 0, true);
    // A MN-major: which field is the stride between 4-element MN chunks?
    run("A MN-major, lbo=4096 sbo=128", idx, onehot_k(16), 4096, 128, 128, 256, 1, 0, true);
    run("A MN-major, lbo=128 sbo=4096", idx, onehot_k(16), 128, 4096, 128, 256, 1, 0, true);
    run("A MN-major, lbo=256 sbo=128", idx, onehot_k(16), 256, 128, 128, 256, 1, 0, true);
    run("A MN-major, lbo=128 sbo=256", idx, onehot_k(16), 128, 256, 128, 256, 1, 0, true);
    run("A MN-major, lbo=144 sbo=4608", idx, onehot_k(16), 144, 4608, 128, 256, 1, 0, true);
    run("A MN-major, lbo=4608 sbo=144", idx, onehot_k(16), 4608, 144, 128, 256, 1, 0, true);
    // B MN-major discovery: A one-hot K-major (rows m<8 select k=m) -> D[m][n] = word index of B(n, k=m)
    run("B MN-major, lbo=4096 sbo=128", onehot_k(128), idx, 128, 256, 4096, 128, 0, 1, false);
    run("B MN-major, lbo=128 sbo=4096", onehot_k(128), idx, 128, 256, 128, 4096, 0, 1, false);
    run("B MN-major, lbo=256 sbo=128", onehot_k(128), idx, 128, 256, 256, 128, 0, 1, false);
    run("B MN-major, lbo=128 sbo=256", onehot_k(128), idx, 128, 256, 128, 256, 0, 1, false);
    return 0;
}
