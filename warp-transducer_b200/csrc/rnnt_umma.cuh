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
// hi = tf32(x) (truncated) and lo = x - hi (exact in fp32; the tensor core keeps its top 11 bits), and
// each k-step issues three MMAs into the same accumulator: hi*hi + hi*lo + lo*hi.  The dropped lo*lo
// term is 2^-20 relative; the sums feed a logarithm (S) and gradients checked to 1e-4 (P, Q).
//
// Operands are staged global -> registers (split) -> shared memory in the UMMA canonical K-major
// NO-SWIZZLE ("interleave") layout, in units of 16-byte chunks (cute/atom/mma_traits_sm100.hpp:190-203):
//   ((8,n),2):((1,SBO),LBO)     element (mn,k) at (mn%8)*16 + (mn/8)*SBO + (k/4)*LBO + (k%4)*4
// (verified element by element on a B200 with tools/probe/umma_layout_probe.cu).  The chunk stride is
// padded (144 B instead of 128 B) so that the scalar stores of a warp hit 32 different banks.
//
// One CTA = 256 threads = 8 warps, two per lane quarter of the 128-lane accumulator.  All threads
// produce the k-stages (one shared-memory stage buffer, the next stage's loads in flight in registers),
// thread 0 issues the MMAs of a stage and commits them to the mbarrier; the epilogue reads the
// accumulator with tcgen05.ld.  gemm_kernel is the generic form (S, and dF / dG beyond 32 label positions);
// grad_fused_kernel computes P and Q in one pass over Ef (see there).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

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
// Spin on test_wait: a tensor-core commit arrives within a few hundred cycles.  (try_wait with a suspend-time
// hint compiles to SYNCS.TRYWAIT + NANOSLEEP of the hint: with 1000 ns every wait for a ~0.3 us product was
// rounded up to whole microseconds - measured: it WAS the run time of the contraction kernels.)
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
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
// The same descriptor as two 32-bit words: only the start-address field (low word, bits [0,14)) changes
// between the k-steps of a tile, so a kernel keeps one low word per tile and adds a constant per step.
__host__ __device__ constexpr uint32_t smem_desc_lo(uint32_t start_bytes, uint32_t lbo_bytes) {
    return ((start_bytes >> 4) & 0x3fff) | (((lbo_bytes >> 4) & 0x3fff) << 16);
}
__host__ __device__ constexpr uint32_t smem_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fff) | (1u << 14); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo, uint32_t hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
    return d;
}
// instruction descriptor for kind::tf32: D fp32 [4,6)=1, A tf32 [7,10)=2, B tf32 [10,13)=2,
// a_major bit 15, b_major bit 16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t instr_desc_tf32(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

constexpr int kPad = 144;   // padded stride between the 16-byte k-chunks of a row group: 8 chunks (128 B) + 16 B

// Geometry of one operand tile [MN x KS] in shared memory: K-major, no swizzle.  Chunk (mn/8, k/4) holds
// 8 rows x 16 B; the k-chunks of a row group are kPad apart (LBO), row groups (KS/4)*kPad apart (SBO).
// (MN-major tf32 operands were tried: tcgen05.mma kind::tf32 returns zeros for them in the no-swizzle
//  layouts on sm_100a - tools/probe/umma_layout_probe.cu - so operands that are MN-contiguous in global
//  memory are transposed on their way into shared memory instead.)
struct TileGeom {
    static __host__ __device__ constexpr uint32_t lbo() { return kPad; }
    // row-group stride: the k-chunks of a group plus a pad that makes it 16 (mod 128) bytes, so that 16-byte
    // stores of 8 lanes to rows 4 apart in consecutive row groups (StageT4) fall on 8 different bank groups
    static __host__ __device__ constexpr uint32_t sbo(int KS) {
        return (uint32_t)(KS / 4) * kPad + (16u + 128u - ((uint32_t)(KS / 4) * kPad) % 128u) % 128u;
    }
    static __host__ __device__ constexpr uint32_t bytes(int MN, int KS) { return (uint32_t)(MN / 8) * sbo(KS); }
    static __device__ __forceinline__ uint32_t off(int mn, int k, int KS) {
        return (uint32_t)(mn & 7) * 16 + (uint32_t)(mn >> 3) * sbo(KS) + (uint32_t)(k >> 2) * kPad + (uint32_t)(k & 3) * 4;
    }
    // descriptor start offset of k-step j (8 values of k = two 16-byte chunks)
    static __host__ __device__ constexpr uint32_t kstep(int j) { return (uint32_t)(2 * j) * kPad; }
};

// element (mn, k) of batch b lives at p + b*batch + mn*s_mn + k*s_k (one of the two strides is 1);
// rows mn >= mn_valid and columns k >= the slice end read as 0
struct Operand {
    const float* p;
    long long batch;
    int s_mn, s_k;
    int mn_valid;
};

// hi = x truncated to tf32 (one LOP3; cvt.rna.tf32 is a four-instruction sequence on sm_100 and the staging
// loops are issue-bound), lo = x - hi exactly; the tensor core then drops at most 2^-20 |x| from lo.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
    lo = x - hi;
}

// One operand's share of a k-stage, staged through registers.  init() fixes, once per CTA, this
// thread's first element (global pointer, shared-memory offset) - every further element of the thread is a
// COMPILE-TIME step away in shared memory and a constant stride away in global memory, so the stage loop
// carries no index arithmetic.  load() issues all of the thread's global loads for the stage at k0 (they
// stay in flight while the previous stage is multiplied), store() splits into hi / lo and writes the two
// shared-memory copies.
//  MODE 2 (k-contiguous, 16-byte aligned rows): float4 loads, one 16-byte chunk per store (KS in {8,16,32,64}).
//  MODE 1 (k-contiguous, unaligned rows, KS <= 32): scalar; lane = k, warp = row.
//  MODE 0 (mn-contiguous): scalar; a warp covers 8 consecutive mn x 4 consecutive k - four full 32-byte
//         sectors per load instruction, 32 different banks per store (the transposition happens here).
//  MODE 3 (mn-contiguous, 16-byte aligned rows): float4 loads of 4 consecutive mn at one k; a warp covers
//         16 mn x 8 k (eight 64-byte row pieces per load instruction); the four values go to four
//         consecutive 16-byte chunks of the K-major layout (immediate offsets), bank-conflict free at KS=24.
// Rows beyond mn_valid are neither loaded nor stored: a row of D depends on its own row of A only (and a
// column on its own row of B), and the epilogue never reads those rows/columns - so M tiles with few real
// rows (T=150: the second tile has 22) cost what their real rows cost.  The k padding IS zero-filled.
constexpr int kThreads = 256;   // 8 warps: two per accumulator lane quarter (they split the columns in the epilogue)
template <int MODE, int MN, int KS> struct StageRegs {
    static constexpr int CH = KS / 4;                 // 16-byte chunks per row
    static constexpr int RP = kThreads / (CH > 0 ? CH : 1);   // MODE 2: rows per pass
    static constexpr int MB = MN / 8;                 // MODE 0: 8-row blocks per k-group
    static constexpr int Q = MB >= 8 ? MB / 8 : 1;    // MODE 0, MB >= 8: passes over mn per k-group
    static constexpr int KB_STEP = MB >= 8 ? 1 : 8 / MB;   // MODE 0: k-groups advanced per pass (MB < 8: several at once)
    static constexpr int MT = MN / 16 > 0 ? MN / 16 : 1;   // MODE 3: 16-row tiles along mn
    static constexpr int KT = KS / 8;                     // MODE 3: 8-wide tiles along k
    static constexpr int PER = MODE == 2   ? (MN + RP - 1) / RP
                               : MODE == 1 ? MN / 8
                               : MODE == 3 ? (MT * KT + 7) / 8
                               : MB >= 8   ? Q * CH
                                           : (CH + KB_STEP - 1) / KB_STEP;
    static_assert(MODE != 3 || (MN % 16 == 0 && (MT >= 8 ? MT % 8 == 0 : 8 % MT == 0)), "MODE 3 needs MN in {16..128} or a multiple of 128");
    static_assert(MODE != 2 || (kThreads % CH == 0 && RP % 8 == 0), "MODE 2 needs KS in {8,16,32,64}");
    static_assert(MODE != 1 || KS <= 32, "MODE 1 needs KS <= 32");
    static_assert(MODE != 0 || (MB >= 8 ? MB % 8 == 0 : 8 % MB == 0), "MODE 0 needs MN in {8,16,32,64} or a multiple of 64");

    float4 v4[(MODE == 2 || MODE == 3) ? PER : 1];
    float v1[(MODE == 2 || MODE == 3) ? 1 : PER];
    const float* g0;      // this thread's element 0 at k0 = 0
    int g_mn, g_k;        // global strides (elements) along mn / k; offsets inside one batch item fit 32 bits
    uint32_t o0;          // shared-memory offset of element 0
    int mn_first, k_first, mn_lim;

    __device__ __forceinline__ void init(const Operand& op, const float* base, int mn0) {
        const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
        g_mn = op.s_mn, g_k = op.s_k;
        if (MODE == 2) {
            mn_first = tid / CH;
            k_first = 4 * (tid % CH);
        } else if (MODE == 1) {
            mn_first = w;
            k_first = lane;
        } else if (MODE == 3) {
            // warp tile = 16 mn x 8 k; tile index of pass j: w + 8j -> (mt, kt) = (ti % MT, ti / MT)
            mn_first = (MT >= 8 ? w : w % MT) * 16 + (tid & 3) * 4;
            k_first = (MT >= 8 ? 0 : w / MT) * 8 + ((tid >> 2) & 7);
        } else {
            const int mm = tid & 7, kk = (tid >> 3) & 3;
            mn_first = (MB >= 8 ? w : w % MB) * 8 + mm;
            k_first = (MB >= 8 ? 0 : w / MB) * 4 + kk;
        }
        o0 = TileGeom::off(mn_first, k_first, KS);
        g0 = base + ((mn0 + mn_first) * g_mn + k_first * g_k);
        mn_lim = op.mn_valid - mn0 - mn_first;   // element j is a real row iff its mn step < mn_lim
    }
    // (mn step, k step) of element j relative to element 0 - compile-time
    static __device__ __forceinline__ constexpr int dmn(int j) {
        return MODE == 2 ? j * RP : MODE == 1 ? 8 * j : MODE == 3 ? (MT >= 8 ? 128 * (j % (MT / 8)) : 0)
                                                      : MB >= 8   ? 64 * (j % Q)
                                                                  : 0;
    }
    static __device__ __forceinline__ constexpr int dk(int j) {
        return MODE == 2 ? 0 : MODE == 1 ? 0 : MODE == 3 ? (MT >= 8 ? 8 * (j / (MT / 8)) : 8 * (8 / MT) * j)
                                                      : MB >= 8   ? 4 * (j / Q)
                                                                  : 4 * KB_STEP * j;
    }
    __device__ __forceinline__ void load(int k0, int k_end) {
        const float* g = g0 + k0 * g_k;
        if (MODE == 2 && k0 + KS <= k_end) {
            // every stage but the last of a k range: no k raggedness - one predicated 16-byte load per row
            // (CTA-uniform branch; the general form below costs ~3x the instructions in an issue-bound loop)
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const bool in = dmn(j) < mn_lim && mn_first + dmn(j) < MN;
                v4[j] = in ? __ldg(reinterpret_cast<const float4*>(g + dmn(j) * g_mn)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int kj = k0 + k_first + dk(j);
            const bool in = dmn(j) < mn_lim && kj < k_end && (MODE != 1 || k_first < KS) &&
                            (MODE == 2 ? mn_first + dmn(j) < MN : k_first + dk(j) < KS);
            const float* p = g + (dmn(j) * g_mn + dk(j) * g_k);
            if (MODE == 3) {
                v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in) {
                    if (dmn(j) + 3 < mn_lim) {
                        v4[j] = __ldg(reinterpret_cast<const float4*>(p));
                    } else {   // ragged end of the mn range
                        v4[j].x = __ldg(p);
                        if (dmn(j) + 1 < mn_lim) v4[j].y = __ldg(p + 1);
                        if (dmn(j) + 2 < mn_lim) v4[j].z = __ldg(p + 2);
                    }
                }
            } else if (MODE == 2) {
                v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (in) {
                    if (kj + 3 < k_end) {
                        v4[j] = __ldg(reinterpret_cast<const float4*>(p));
                    } else {   // ragged end of the k range
                        v4[j].x = __ldg(p);
                        if (kj + 1 < k_end) v4[j].y = __ldg(p + 1);
                        if (kj + 2 < k_end) v4[j].z = __ldg(p + 2);
                    }
                }
            } else {
                v1[j] = in ? __ldg(p) : 0.0f;
            }
        }
    }
    __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo) const {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            // offset step of element j: row groups are SBO apart, k-chunks kPad apart (compile-time)
            const uint32_t o = o0 + (uint32_t)(dmn(j) / 8) * TileGeom::sbo(KS) + (uint32_t)(dk(j) / 4) * kPad;
            // rows past mn_valid are skipped (see above); the k padding of real rows is written as zeros
            const bool slot = (MODE == 2 ? mn_first + dmn(j) < MN : k_first + dk(j) < KS) && (MODE != 1 || k_first < KS) &&
                              dmn(j) < mn_lim;
            if (slot) {
                if (MODE == 3) {
                    const float xs[4] = {v4[j].x, v4[j].y, v4[j].z, v4[j].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {   // mn + c sits one 16-byte chunk further (same 8-row group)
                        float h, l;
                        split_tf32(xs[c], h, l);
                        *reinterpret_cast<float*>(hi + o + 16 * c) = h;
                        *reinterpret_cast<float*>(lo + o + 16 * c) = l;
                    }
                } else if (MODE == 2) {
                    float4 h, l;
                    split_tf32(v4[j].x, h.x, l.x);
                    split_tf32(v4[j].y, h.y, l.y);
                    split_tf32(v4[j].z, h.z, l.z);
                    split_tf32(v4[j].w, h.w, l.w);
                    *reinterpret_cast<float4*>(hi + o) = h;
                    *reinterpret_cast<float4*>(lo + o) = l;
                } else {
                    float h, l;
                    split_tf32(v1[j], h, l);
                    *reinterpret_cast<float*>(hi + o) = h;
                    *reinterpret_cast<float*>(lo + o) = l;
                }
            }
        }
    }
};

// MN-contiguous operand with 16-byte aligned rows, transposed IN REGISTERS: a thread fetches a 4 (k) x 4 (mn)
// block as four float4 loads (a warp: 32 adjacent blocks of one k group = four fully coalesced 512-byte row
// pieces), and writes it as four 16-byte chunks per copy (one per mn: the four k values of a K-major chunk) -
// 8 STS.128 instead of the 32 scalar stores of StageRegs MODE 3.  The hi / lo split here truncates
// (hi = x & ~0x1fff, lo = x - hi, exact): one LOP3 instead of the four-instruction cvt.rna sequence; the
// tensor core then drops at most 2^-20 of x from lo, instead of 2^-21 with rounding.
// Same interface as StageRegs.  (ncu round 2: the fused gradient kernel executed 1080 warp instructions per
// warp and 32-frame chunk, most of them in MODE 3 staging.)
template <int MN, int KS> struct StageT4 {
    static constexpr int MG = MN / 4, KG = KS / 4, UNITS = MG * KG;
    static constexpr int PER = (UNITS + kThreads - 1) / kThreads;
    static_assert(MN % 128 == 0 || MN == 32 || MN == 64, "a warp covers 32 adjacent blocks of one k group");
    float4 v[PER][4];
    const float* g0;
    int g_k;
    uint32_t o0;
    int mn_lim, k_first;
    bool active;
    __device__ __forceinline__ void init(const Operand& op, const float* base, int mn0) {
        const int tid = threadIdx.x;
        const int mg = tid % MG, kg = tid / MG;
        g_k = op.s_k;
        k_first = 4 * kg;
        active = tid < UNITS || PER > 1;
        o0 = TileGeom::off(4 * mg, 4 * kg, KS);
        g0 = base + (mn0 + 4 * mg + k_first * g_k);
        mn_lim = op.mn_valid - mn0 - 4 * mg;   // > 0: the block's four rows exist (mn_valid is a multiple of 4)
    }
    __device__ __forceinline__ void load(int k0, int k_end) {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int kk = k_first + p * 4 * (kThreads / MG);
            const float* gp = g0 + (k0 + p * 4 * (kThreads / MG)) * g_k;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool in = active && mn_lim > 0 && kk < KS && k0 + kk + j < k_end;
                v[p][j] = in ? __ldg(reinterpret_cast<const float4*>(gp + j * g_k)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    static __device__ __forceinline__ void split(float x, float& hi, float& lo) {
        hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);
        lo = x - hi;
    }
    __device__ __forceinline__ void store(unsigned char* hi, unsigned char* lo) const {
#pragma unroll
        for (int p = 0; p < PER; ++p) {
            const int kk = k_first + p * 4 * (kThreads / MG);
            if (!(active && mn_lim > 0 && kk < KS)) continue;
            const uint32_t o = o0 + (uint32_t)(p * (kThreads / MG)) * kPad;
            const float xs[4][4] = {{v[p][0].x, v[p][1].x, v[p][2].x, v[p][3].x},
                                    {v[p][0].y, v[p][1].y, v[p][2].y, v[p][3].y},
                                    {v[p][0].z, v[p][1].z, v[p][2].z, v[p][3].z},
                                    {v[p][0].w, v[p][1].w, v[p][2].w, v[p][3].w}};
#pragma unroll
            for (int c = 0; c < 4; ++c) {   // row mn + c: one 16-byte chunk further inside the 8-row group
                float4 h, l;
                split(xs[c][0], h.x, l.x);
                split(xs[c][1], h.y, l.y);
                split(xs[c][2], h.z, l.z);
                split(xs[c][3], h.w, l.w);
                *reinterpret_cast<float4*>(hi + o + 16 * c) = h;
                *reinterpret_cast<float4*>(lo + o + 16 * c) = l;
            }
        }
    }
};

// What the epilogue does with accumulator element (m, n) of batch b, k slice `slice`:
//   out[slice*out_s + b*out_b + m*out_m + n*out_n] = acc * (in ? in[b*in_b + m*in_m + n*in_n] : 1)
// The `in` values of a group of columns are fetched before the accumulator is waited for.
struct Epilogue {
    const float* in;
    long long in_b;
    int in_m, in_n;      // strides inside one batch item: 32-bit
    float* out;
    long long out_s, out_b;
    int out_m, out_n;
};

// D[128 x N] = sum_{k in slice} A[m0+m][k] * B[n0+n][k]   for batch blockIdx.z, M tile blockIdx.y, and
// blockIdx.x = k slice + slices * N tile.  A_MODE / B_MODE: how the operand is fetched (StageRegs).
// One shared-memory stage buffer per CTA (kBufs): the footprint stays small, so 4-6 CTAs are resident per SM
// and cover each other's load -> split -> multiply chains; within a CTA the global loads of stage s+1 are in
// flight in registers while stage s is multiplied.  Elements with m >= A.mn_valid or n >= B.mn_valid are skipped.
constexpr int kBufs = 1;
template <int A_MODE, int B_MODE, int N, int KS>
__global__ void __launch_bounds__(kThreads, 3)
gemm_kernel(const Operand A, const Operand B, int K, int slices, const Epilogue epi) {
    static_assert(N % 32 == 0 && N >= 32 && N <= 256 && KS % 8 == 0, "UMMA shape (two epilogue warps per lane quarter)");
    constexpr uint32_t A_BYTES = TileGeom::bytes(128, KS), B_BYTES = TileGeom::bytes(N, KS);
    constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mma_done[2];
    __shared__ uint32_t tmem_base_slot;

    const int b = blockIdx.z, m0 = blockIdx.y * 128, slice = blockIdx.x % slices, n0 = (blockIdx.x / slices) * N;
    const int kper = ((K + slices - 1) / slices + KS - 1) / KS * KS;
    const int kbeg = slice * kper, kend = min(K, kbeg + kper);
    const int nstages = kbeg < kend ? (kend - kbeg + KS - 1) / KS : 0;
    const int warp = threadIdx.x >> 5;

    StageRegs<A_MODE, 128, KS> ra;
    StageRegs<B_MODE, N, KS> rb;
    ra.init(A, A.p + (long long)b * A.batch, m0);
    rb.init(B, B.p + (long long)b * B.batch, n0);
    if (nstages > 0) {   // first stage's loads go out before anything else
        ra.load(kbeg, kend);
        rb.load(kbeg, kend);
    }
    if (warp == 0) tmem_alloc(s32(&tmem_base_slot), TMEM_COLS);
    if (threadIdx.x == 0) {
        bar_init(s32(&mma_done[0]), 1);
        bar_init(s32(&mma_done[1]), 1);
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_d = tmem_base_slot;
    constexpr uint32_t idesc = instr_desc_tf32(128, N, false, false);

    for (int s = 0; s < nstages; ++s) {
        const int buf = s % kBufs;
        unsigned char* a_hi = smem + buf * STAGE_BYTES;
        unsigned char* a_lo = a_hi + A_BYTES;
        unsigned char* b_hi = a_lo + A_BYTES;
        unsigned char* b_lo = b_hi + B_BYTES;
        // the MMAs of the stage that last used this buffer have finished reading it
        if (s >= kBufs) bar_wait(s32(&mma_done[buf]), (uint32_t)(((s - kBufs) / kBufs) & 1));
        ra.store(a_hi, a_lo);
        rb.store(b_hi, b_lo);
        if (s + 1 < nstages) {   // next stage's loads fly during the barrier and the MMAs
            ra.load(kbeg + (s + 1) * KS, kend);
            rb.load(kbeg + (s + 1) * KS, kend);
        }
        fence_smem_async();      // generic-proxy stores -> visible to the tensor core's async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
#pragma unroll
            for (int j = 0; j < KS / 8; ++j) {
                const uint64_t ah = smem_desc(s32(a_hi) + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                const uint64_t al = smem_desc(s32(a_lo) + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                const uint64_t bh = smem_desc(s32(b_hi) + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                const uint64_t bl = smem_desc(s32(b_lo) + TileGeom::kstep(j), TileGeom::lbo(), TileGeom::sbo(KS));
                mma_tf32(tmem_d, ah, bh, idesc, (s | j) != 0);
                mma_tf32(tmem_d, ah, bl, idesc, 1);
                mma_tf32(tmem_d, al, bh, idesc, 1);
            }
            mma_commit(s32(&mma_done[buf]));
        }
    }

    // epilogue: warps w and w+4 own accumulator lanes [32(w%4), +32) and one half of the columns each;
    // lane = row m of the tile.  The `in` values of a group of 16 columns are fetched before the accumulator
    // is waited for / read; all addresses are a base plus a running stride.
    const int quarter = warp & 3, half = warp >> 2;
    const int m = m0 + quarter * 32 + (threadIdx.x & 31);
    const bool mrow = m < A.mn_valid;
    constexpr int HALF = N / 2;
    const float* ip = epi.in ? epi.in + b * epi.in_b + (m * epi.in_m + (n0 + half * HALF) * epi.in_n) : nullptr;
    float* op = epi.out + slice * epi.out_s + b * epi.out_b + (m * epi.out_m + (n0 + half * HALF) * epi.out_n);
    const int nlim = B.mn_valid - n0 - half * HALF;   // columns of this half that exist
    bool waited = false;
#pragma unroll 1
    for (int c = 0; c < HALF; c += 16) {
        float pre[16], v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pre[i] = (ip && mrow && c + i < nlim) ? __ldg(ip + (c + i) * epi.in_n) : 1.0f;
        if (!waited) {
            // commits complete in issue order: the last stage's barrier covers all of them
            if (nstages > 0) bar_wait(s32(&mma_done[(nstages - 1) % kBufs]), (uint32_t)(((nstages - 1) / kBufs) & 1));
            fence_after();
            waited = true;
        }
        tmem_ld16(tmem_d + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * HALF + c), v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (mrow && c + i < nlim) op[(c + i) * epi.out_n] = nstages > 0 ? v[i] * pre[i] : 0.0f;
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, TMEM_COLS);
}

// =================================================================================================
// Both gradient contractions of the additive joint in ONE pass over Ef (the [N,T,V] factor that dominates
// the traffic).  CTA = (utterance b, 128 vocabulary entries v0..v0+127 = the accumulator lanes).  Time is
// walked in chunks of TC frames; per chunk
//
//   Q[v,u] += sum_{t in chunk} Ef[t,v] Wm[t,u]        accumulates over all chunks     (K = t)
//   P[v,t]  = sum_u            Eg[u,v] Wm[t,u]        complete after this chunk       (K = u, N = t in chunk)
//   dF[t,v] = Ef[t,v] * P[v,t]                         written per chunk
//
// and after the last chunk  dG[u,v] = Eg[u,v] * Q[v,u].  Ef is read from DRAM once (the chunk's epilogue
// re-reads the tile it has just staged - an L1/L2 hit) instead of once per contraction, and the P / Q products
// of a chunk are issued back to back by one thread.  Tensor memory: Q in columns [0,NU), P (two buffers) behind it.
// Shared memory: EgT (staged once), EfT / WmTU / WmUT of the current chunk, each as hi and lo tf32 copies.
// =================================================================================================
struct GradFused {
    const float *ef, *eg, *wm;   // [N,T,V], [N,U,V], [N,T,kWmPad]: Wm rows zero-padded to kWmPad label positions
    float *dF, *dG;              // [N,T,V], [N,U,V]
    int T, U, V;
};
constexpr int kWmPad = 32;       // the weights' row pitch (16-byte aligned rows: both Wm operands are fetched as float4)
template <int NU, int TC> struct GradFusedGeom {
    static constexpr int KU = NU;   // k extent of the P product (label positions, padded)
    static constexpr uint32_t A1 = TileGeom::bytes(128, KU), A2 = TileGeom::bytes(128, TC);
    static constexpr uint32_t B1 = TileGeom::bytes(TC, KU), B2 = TileGeom::bytes(NU, TC);
    static constexpr uint32_t total = 2 * (A1 + A2 + B1 + B2);
};
template <int NU, int TC, int A_MODE>   // A_MODE 3: rows of Ef / Eg 16-byte aligned (V % 4 == 0), else 0
__global__ void __launch_bounds__(kThreads, 2)
grad_fused_kernel(const GradFused g) {
    static_assert(NU == 32 && TC == 32, "instantiated shape: up to 32 label positions, 32 frames per chunk");
    using G = GradFusedGeom<NU, TC>;
    constexpr int KU = G::KU;
    constexpr uint32_t TMEM_COLS = 128;   // Q [0,NU), P double-buffered [NU, NU + 2 TC)
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long mma_done;
    __shared__ uint32_t tmem_base_slot;
    unsigned char* a1_hi = smem;
    unsigned char* a1_lo = a1_hi + G::A1;
    unsigned char* a2_hi = a1_lo + G::A1;
    unsigned char* a2_lo = a2_hi + G::A2;
    unsigned char* b1_hi = a2_lo + G::A2;
    unsigned char* b1_lo = b1_hi + G::B1;
    unsigned char* b2_hi = b1_lo + G::B1;
    unsigned char* b2_lo = b2_hi + G::B2;

    const int b = blockIdx.z, v0 = blockIdx.x * 128;
    const int T = g.T, U = g.U, V = g.V;
    const int warp = threadIdx.x >> 5;
    const float* ef_b = g.ef + (long long)b * T * V;
    const float* eg_b = g.eg + (long long)b * U * V;
    const float* wm_b = g.wm + (long long)b * T * kWmPad;
    const Operand EgT{g.eg, 0, 1, V, V}, EfT{g.ef, 0, 1, V, V};
    const Operand WmTU{g.wm, 0, kWmPad, 1, T};        // (n = t, k = u): k-contiguous, aligned rows
    const Operand WmUT{g.wm, 0, 1, kWmPad, kWmPad};   // (n = u, k = t): n-contiguous, transposed in registers
    const int nchunks = (T + TC - 1) / TC;

    typename std::conditional<A_MODE == 3, StageT4<128, KU>, StageRegs<0, 128, KU>>::type r_eg;
    typename std::conditional<A_MODE == 3, StageT4<128, TC>, StageRegs<0, 128, TC>>::type r_ef;
    StageRegs<2, TC, KU> r_w1;   // one float4 per thread
    StageT4<NU, TC> r_w2;        // 64 threads, a 4x4 block each
    r_eg.init(EgT, eg_b, v0);
    r_ef.init(EfT, ef_b, v0);
    r_w2.init(WmUT, wm_b, 0);
    r_eg.load(0, U);
    r_ef.load(0, T);
    r_w1.init(WmTU, wm_b, 0);
    r_w1.load(0, kWmPad);
    r_w2.load(0, T);
    if (warp == 0) tmem_alloc(s32(&tmem_base_slot), TMEM_COLS);
    if (threadIdx.x == 0) bar_init(s32(&mma_done), 1);
    r_eg.store(a1_hi, a1_lo);
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_q = tmem_base_slot, tmem_p = tmem_base_slot + NU;
    constexpr uint32_t idesc_q = instr_desc_tf32(128, NU, false, false);
    constexpr uint32_t idesc_p = instr_desc_tf32(128, TC, false, false);

    // low descriptor words of the eight operand tiles (the buffers never move)
    const uint32_t d_a1h = smem_desc_lo(s32(a1_hi), kPad), d_a1l = smem_desc_lo(s32(a1_lo), kPad);
    const uint32_t d_a2h = smem_desc_lo(s32(a2_hi), kPad), d_a2l = smem_desc_lo(s32(a2_lo), kPad);
    const uint32_t d_b1h = smem_desc_lo(s32(b1_hi), kPad), d_b1l = smem_desc_lo(s32(b1_lo), kPad);
    const uint32_t d_b2h = smem_desc_lo(s32(b2_hi), kPad), d_b2l = smem_desc_lo(s32(b2_lo), kPad);
    // epilogue geometry: warps w and w+4 own accumulator lanes [32(w%4), +32) and 16 columns each
    const int quarter = warp & 3, half = warp >> 2;
    const int m = v0 + quarter * 32 + (threadIdx.x & 31);
    const bool mrow = m < V;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;

    // dF of chunk c: Ef values first (L1/L2 hits), then the accumulator (P buffer c & 1)
    auto epilogue_p = [&](int c) {
        const int tb = c * TC + half * 16;
        const float* ip = ef_b + (tb * V + m);
        float* op = g.dF + (long long)b * T * V + (tb * V + m);
        float pre[16], v[16];
        const uint32_t taddr = tmem_p + (uint32_t)((c & 1) * TC) + lane_addr + (uint32_t)(half * 16);
        // (warp-uniform branch: tcgen05.ld is a warp-collective instruction)
        if (__all_sync(0xffffffffu, mrow) && tb + 16 <= T) {   // the common case: every row and all 16 frames exist
#pragma unroll
            for (int i = 0; i < 16; ++i) pre[i] = __ldg(ip + i * V);
            tmem_ld16(taddr, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) op[i * V] = v[i] * pre[i];
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pre[i] = (mrow && tb + i < T) ? __ldg(ip + i * V) : 0.0f;
            tmem_ld16(taddr, v);   // (warp-collective: executed by every lane, rows that do not exist included)
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (mrow && tb + i < T) op[i * V] = v[i] * pre[i];
        }
    };
    // Software pipeline: the products of chunk c run on the tensor core while the threads write chunk c-1's
    // dF (P is double-buffered in tensor memory) and the loads of chunk c+1 are in flight.
    for (int c = 0; c < nchunks; ++c) {
        const int t0 = c * TC;
        if (c > 0) {   // chunk c-1's products have read the stage buffers and written P[(c-1)&1]
            bar_wait(s32(&mma_done), (uint32_t)((c - 1) & 1));
            fence_after();
        }
        r_ef.store(a2_hi, a2_lo);
        r_w1.store(b1_hi, b1_lo);
        r_w2.store(b2_hi, b2_lo);
        if (c + 1 < nchunks) {
            r_ef.load(t0 + TC, T);
            r_w1.init(WmTU, wm_b, t0 + TC);
            r_w1.load(0, kWmPad);
            r_w2.load(t0 + TC, T);
        }
        fence_smem_async();
        fence_before();          // epilogue c-2's tcgen05.ld of P[c&1] precede the barrier
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after();
            const uint32_t tp = tmem_p + (uint32_t)((c & 1) * TC);
#pragma unroll
            for (int j = 0; j < TC / 8; ++j) {
                constexpr uint32_t hi = smem_desc_hi(TileGeom::sbo(TC));
                const uint32_t st = TileGeom::kstep(j) >> 4;
                const uint64_t ah = desc64(d_a2h + st, hi), al = desc64(d_a2l + st, hi);
                const uint64_t bh = desc64(d_b2h + st, hi), bl = desc64(d_b2l + st, hi);
                mma_tf32(tmem_q, ah, bh, idesc_q, (c | j) != 0);
                mma_tf32(tmem_q, ah, bl, idesc_q, 1);
                mma_tf32(tmem_q, al, bh, idesc_q, 1);
            }
#pragma unroll
            for (int j = 0; j < KU / 8; ++j) {
                constexpr uint32_t hi = smem_desc_hi(TileGeom::sbo(KU));
                const uint32_t st = TileGeom::kstep(j) >> 4;
                const uint64_t ah = desc64(d_a1h + st, hi), al = desc64(d_a1l + st, hi);
                const uint64_t bh = desc64(d_b1h + st, hi), bl = desc64(d_b1l + st, hi);
                mma_tf32(tp, ah, bh, idesc_p, j != 0);
                mma_tf32(tp, ah, bl, idesc_p, 1);
                mma_tf32(tp, al, bh, idesc_p, 1);
            }
            mma_commit(s32(&mma_done));
        }
        if (c > 0) epilogue_p(c - 1);
    }
    bar_wait(s32(&mma_done), (uint32_t)((nchunks - 1) & 1));
    fence_after();
    epilogue_p(nchunks - 1);
    // dG: Q is complete (the last commit covered every product)
    {
        const int ub = half * 16;
        const float* ip = eg_b + (ub * V + m);
        float* op = g.dG + (long long)b * U * V + (ub * V + m);
        float pre[16], v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pre[i] = (mrow && ub + i < U) ? __ldg(ip + i * V) : 0.0f;
        tmem_ld16(tmem_q + lane_addr + (uint32_t)ub, v);
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (mrow && ub + i < U) op[i * V] = v[i] * pre[i];
    }
    fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_slot, TMEM_COLS);
}

template <int N, int KS>
constexpr size_t gemm_smem_bytes() {   // kBufs stage buffers, each with hi and lo copies of both operand tiles
    return kBufs * (2 * (size_t)TileGeom::bytes(128, KS) + 2 * (size_t)TileGeom::bytes(N, KS));
}

}  // namespace umma
}  // namespace b200rnnt
