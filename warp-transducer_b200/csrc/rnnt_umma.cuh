// rnnt_umma.cuh — batched fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05.mma
// kind::tf32, accumulator in tensor memory) for the three contractions of the additive-joint RNN-T
// loss (rnnt_joint.cuh):
//
//   S [t,u]  = sum_v Ef[t,v] Eg[u,v]          M = t tile, N = u,      K = v (split over CTAs)
//   P [v,t]  = sum_u Eg[u,v] Wm[t,u]          M = v tile, N = t,      K = u      dF[t,v] = Ef[t,v] P[v,t]
//   Q [v,u]  = sum_t Ef[t,v] Wm[t,u]          M = v tile, N = u,      K = t      dG[u,v] = Eg[u,v] Q[v,u]
//
// In P and Q the VOCABULARY index is the M dimension, i.e. the tensor-memory lane: the epilogue thread of
// lane v walks the accumulator's columns and its global accesses are coalesced across the warp.
//
// fp32 accuracy on tf32 tensor cores: every operand element x is split by the producer threads into
// hi = tf32(x) (cvt.rna) and lo = x - hi (exact in fp32; the tensor core keeps its top 11 bits), and
// each k-step issues three MMAs into the same accumulator: hi*hi + hi*lo + lo*hi.  The dropped lo*lo
// term is 2^-22 relative; the sums feed a logarithm (S) and gradients checked to 1e-4 (P, Q).
//
// Operands are staged global -> registers (split) -> shared memory in the UMMA canonical NO-SWIZZLE
// ("interleave") layouts, in units of 16-byte chunks (cute/atom/mma_traits_sm100.hpp:167-203):
//   K-major  : ((8,n),2):((1,SBO),LBO)          element (mn,k) at (mn%8)*16 + (mn/8)*SBO + (k/4)*LBO + (k%4)*4
//   MN-major : ((1,n),(8,k)):((X,SBO),(1,LBO))  element (mn,k) at (mn%4)*4 + (mn/4)*SBO + (k%8)*16 + (k/8)*LBO
// so whichever index is contiguous in global memory stays contiguous in shared memory and no transpose
// is ever needed.  The strides are padded (144 B instead of 128 B) so the scalar stores of a warp hit
// 32 different banks.
//
// One CTA = 128 threads = 4 warps = the 4 lane quarters of the 128-lane accumulator.  All threads
// produce a k-stage, thread 0 issues its MMAs and commits them to an mbarrier, everybody waits for the
// commit before the stage buffers are overwritten; the epilogue reads the accumulator with tcgen05.ld.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200rnnt {
namespace umma {

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {   // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_smem_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar), "r"(parity)
                     : "memory");
}
// 16 consecutive fp32 columns of this thread's accumulator lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors (cute/arch/mma_sm100_desc.hpp) --------------------------------------------------
// shared-memory matrix descriptor, no swizzle: start[0,14) LBO[16,30) SBO[32,46) (all >> 4), version 1 at [46,48)
__host__ __device__ constexpr uint64_t smem_desc(uint32_t start_bytes, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((start_bytes >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}
// instruction descriptor for kind::tf32: D fp32 [4,6)=1, A tf32 [7,10)=2, B tf32 [10,13)=2,
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kPad = 144;   // padded 16-byte-chunk group stride: 8 chunks (128 B) + one chunk of padding

// Geometry of one operand tile [MN x KS] in shared memory.
template <bool MN_MAJOR> struct TileGeom;
template <> struct TileGeom<false> {   // K-major
    // chunk (mn/8, k/4) of 8 rows x 16 B; k-chunks kPad apart, 8-row groups (KS/4)*kPad apart
    static __host__ __device__ constexpr uint32_t lbo(int /*MN*/, int /*KS*/) { return kPad; }
    static __host__ __device__ constexpr uint32_t sbo(int /*MN*/, int KS) { return (uint32_t)(KS / 4) * kPad; }
    static __host__ __device__ constexpr uint32_t bytes(int MN, int KS) { return (uint32_t)(MN / 8) * sbo(MN, KS); }
    static __device__ __forceinline__ uint32_t off(int mn, int k, int MN, int KS) {
        return (uint32_t)(mn & 7) * 16 + (uint32_t)(mn >> 3) * sbo(MN, KS) + (uint32_t)(k >> 2) * kPad + (uint32_t)(k & 3) * 4;
    }
    // descriptor start offset of k-step j (8 values of k = two 16-byte chunks)
    static __host__ __device__ constexpr uint32_t kstep(int j, int /*MN*/, int /*KS*/) { return (uint32_t)(2 * j) * kPad; }
};
template <> struct TileGeom<true> {    // MN-major
    // chunk (mn/4, k/8) of 8 k-rows x 16 B; mn-chunks kPad apart, k-groups (MN/4)*kPad apart
    static __host__ __device__ constexpr uint32_t sbo(int /*MN*/, int /*KS*/) { return kPad; }
    static __host__ __device__ constexpr uint32_t lbo(int MN, int /*KS*/) { return (uint32_t)(MN / 4) * kPad; }
    static __host__ __device__ constexpr uint32_t bytes(int MN, int KS) { return (uint32_t)(KS / 8) * lbo(MN, KS); }
    static __device__ __forceinline__ uint32_t off(int mn, int k, int MN, int KS) {
        return (uint32_t)(mn & 3) * 4 + (uint32_t)(mn >> 2) * kPad + (uint32_t)(k & 7) * 16 + (uint32_t)(k >> 3) * lbo(MN, KS);
    }
    static __host__ __device__ constexpr uint32_t kstep(int j, int MN, int KS) { return (uint32_t)j * lbo(MN, KS); }
};

// element (mn, k) of batch b lives at p + b*batch + mn*s_mn + k*s_k; rows mn >= mn_valid and columns k >= k_valid read as 0
struct Operand {
    const float* p;
    long long batch;
    int s_mn, s_k;
    int mn_valid;
};

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

// Fill one stage of an operand: tile rows [mn0, mn0+MN), k in [k0, k0+KS), hi and lo copies.
// The thread index runs along the index that is contiguous in GLOBAL memory, so loads coalesce.
template <bool MN_MAJOR, int MN, int KS>
__device__ __forceinline__ void fill_stage(unsigned char* hi, unsigned char* lo, const Operand& op, const float* base,
                                           int mn0, int k0, int k_end) {
    using G = TileGeom<MN_MAJOR>;
    for (int i = threadIdx.x; i < MN * KS; i += blockDim.x) {
        int mn, k;
        if (MN_MAJOR) { mn = i % MN; k = i / MN; } else { k = i % KS; mn = i / KS; }
        float x = 0.0f;
        if (mn0 + mn < op.mn_valid && k0 + k < k_end)
            x = __ldg(base + (long long)(mn0 + mn) * op.s_mn + (long long)(k0 + k) * op.s_k);
        float h, l;
        split_tf32(x, h, l);
        const uint32_t o = G::off(mn, k, MN, KS);
        *reinterpret_cast<float*>(hi + o) = h;
        *reinterpret_cast<float*>(lo + o) = l;
    }
}

// D[128 x N] = sum_{k in [kbeg,kend)} A[m0+m][k] * B[n][k]   for batch blockIdx.z, M tile blockIdx.y,
// k slice blockIdx.x.  Epi(b, slice, m (global row), n, value) is called for every element with m < A.mn_valid
// and n < B.mn_valid.
template <bool A_MN, bool B_MN, int N, int KS, typename Epi>
__global__ void __launch_bounds__(128)
gemm_kernel(const Operand A, const Operand B, int K, int slices, const Epi epi, const int variant = 0) {
    static_assert(N % 16 == 0 && N >= 16 && N <= 256 && KS % 8 == 0, "UMMA shape");
    using GA = TileGeom<A_MN>;
    using GB = TileGeom<B_MN>;
    constexpr uint32_t A_BYTES = GA::bytes(128, KS), B_BYTES = GB::bytes(N, KS);
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + A_BYTES;
    unsigned char* b_hi = a_lo + A_BYTES;
    unsigned char* b_lo = b_hi + B_BYTES;
    __shared__ __align__(8) unsigned long long mma_done;
    __shared__ uint32_t tmem_base_slot;

    const int b = blockIdx.z, m0 = blockIdx.y * 128, slice = blockIdx.x;
    const int kper = ((K + slices - 1) / slices + KS - 1) / KS * KS;
    const int kbeg = slice * kper, kend = min(K, kbeg + kper);
    const int warp = threadIdx.x >> 5;

    if (warp == 0) tmem_alloc(s32(&tmem_base_slot), TMEM_COLS);
    if (threadIdx.x == 0) bar_init(s32(&mma_done), 1);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    const uint32_t idesc = instr_desc_tf32(128, N, A_MN, B_MN);
    const float* abase = A.p + (long long)b * A.batch;
    const float* bbase = B.p + (long long)b * B.batch;

    uint32_t parity = 0, accum = 0;
    for (int k0 = kbeg; k0 < kend; k0 += KS) {
        fill_stage<A_MN, 128, KS>(a_hi, a_lo, A, abase, m0, k0, kend);
        fill_stage<B_MN, N, KS>(b_hi, b_lo, B, bbase, 0, k0, kend);
        fence_smem_async();      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
#pragma unroll
            for (int j = 0; j < KS / 8; ++j) {
                // (variant bits: probe-only switches that swap the two stride fields of a descriptor)
                const uint32_t al_ = (variant & 1) ? GA::sbo(128, KS) : GA::lbo(128, KS), as_ = (variant & 1) ? GA::lbo(128, KS) : GA::sbo(128, KS);
                const uint32_t bl_ = (variant & 2) ? GB::sbo(N, KS) : GB::lbo(N, KS), bs_ = (variant & 2) ? GB::lbo(N, KS) : GB::sbo(N, KS);
                const uint64_t ah = smem_desc(s32(a_hi) + GA::kstep(j, 128, KS), al_, as_);
                const uint64_t al = smem_desc(s32(a_lo) + GA::kstep(j, 128, KS), al_, as_);
                const uint64_t bh = smem_desc(s32(b_hi) + GB::kstep(j, N, KS), bl_, bs_);
                const uint64_t bl = smem_desc(s32(b_lo) + GB::kstep(j, N, KS), bl_, bs_);
                mma_tf32(tmem_d, ah, bh, idesc, accum);
                mma_tf32(tmem_d, ah, bl, idesc, 1);
                mma_tf32(tmem_d, al, bh, idesc, 1);
                accum = 1;
            }
            mma_commit(s32(&mma_done));
        }
        bar_wait(s32(&mma_done), parity);   // the stage buffers are free again (and, last time, D is complete)
        parity ^= 1;
    }
    fence_after();

    // epilogue: warp w owns accumulator lanes [32w, 32w+32); lane = row m of the tile
    const int m = m0 + warp * 32 + (threadIdx.x & 31);
#pragma unroll 1
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        if (kbeg < kend || true) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (m < A.mn_valid && c0 + i < B.mn_valid) epi(b, slice, m, c0 + i, kbeg < kend ? v[i] : 0.0f);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

template <bool A_MN, bool B_MN, int N, int KS>
constexpr size_t gemm_smem_bytes() {
    return 2 * (size_t)TileGeom<A_MN>::bytes(128, KS) + 2 * (size_t)TileGeom<B_MN>::bytes(N, KS);
}

}  // namespace umma
}  // namespace b200rnnt
