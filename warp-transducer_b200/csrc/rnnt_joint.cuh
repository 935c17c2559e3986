// rnnt_joint.cuh — RNN-T loss for the ADDITIVE joint network, logits never materialised
// (the reference's own timing script builds its logits this way,
// pytorch_binding/test/test_time.py:73: acts = trans.unsqueeze(2) + pred.unsqueeze(1); the
// gradients w.r.t. the two factors are docs/rnnt_notes.tex:147-153).
//
//   h[b,t,u,k] = f[b,t,k] + g[b,u,k]
//
// Because exp(f+g) = exp(f)·exp(g), every V-length reduction of the standard path factorises:
//   S(t,u)   = sum_k e^{f_tk - mf_t} e^{g_uk - mg_u}            = (Ef · Eg^T)[t,u]
//   lse(t,u) = mf_t + mg_u + log S(t,u)
//   dL/df_tk = Ef_tk · sum_u Wm(t,u) Eg_uk  - (blank / label terms),   Wm = e^{alpha+beta-ll} / S
//   dL/dg_uk = Eg_uk · sum_t Wm(t,u) Ef_tk  - (blank / label terms)
// i.e. three small batched fp32 GEMMs around the SAME lattice kernel, and HBM traffic of
// O(N (T+U) V) instead of O(N T U V): 0.44 GB instead of 24 GB on the README large-vocabulary shape.
// fp32 FMA GEMMs, not tensor cores: the sums feed a logarithm and need ~1e-6 relative accuracy,
// and the whole contraction is only 3 x 2 GFMA.
// Limitation (documented in DESIGN.md): the factor-wise maxima bound the terms by 1 but not from
// below; if max_k(f+g) is more than ~85 nats under mf+mg the fp32 sum underflows.
#pragma once
#include "rnnt_kernels.cuh"
#include "rnnt_lattice.cuh"
#include "rnnt_umma.cuh"

namespace b200rnnt {

struct JointDims {
    int N, T, U, V, blank;
};

// ---- J1: per row of a factor: max and exp(x - max) ------------------------------------------------
// one warp per row of V elements (rows = N*T for f, N*U for g)
__global__ void __launch_bounds__(256)
joint_prep_kernel(const float* __restrict__ x, float* __restrict__ e, float* __restrict__ mx, int rows, int V) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* p = x + (size_t)row * V;
    float m = -INFINITY;
    for (int k = lane; k < V; k += 32) m = fmaxf(m, __ldg(p + k));
    m = group_max<32>(m);
    const float mz = (m == -INFINITY) ? 0.0f : m;
    float* q = e + (size_t)row * V;
    for (int k = lane; k < V; k += 32) q[k] = Real<float>::exp(__ldg(p + k) - mz);
    if (lane == 0) mx[row] = m;
}

// Same, one CTA per row with the whole row in registers (V % 4 == 0, V <= 256*4*NV): the factor is read
// ONCE (16-byte loads, all in flight before first use) and its exponentials written once with 16-byte
// stores - the streaming shape of rowstats_row_kernel.
template <int NV>
__global__ void __launch_bounds__(256)
joint_prep_row_kernel(const float* __restrict__ x, float* __restrict__ e, float* __restrict__ mx, int V) {
    __shared__ float sh[8];
    const int row = blockIdx.x, nv = V >> 2;
    const float4* p = reinterpret_cast<const float4*>(x + (size_t)row * V);
    float4 v[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = threadIdx.x + j * 256;
        v[j] = i < nv ? __ldg(p + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < NV; ++j) m = fmaxf(m, fmaxf(fmaxf(v[j].x, v[j].y), fmaxf(v[j].z, v[j].w)));
    m = group_max<32>(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    m = sh[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) m = fmaxf(m, sh[w]);
    const float mz = (m == -INFINITY) ? 0.0f : m;
    float4* q = reinterpret_cast<float4*>(e + (size_t)row * V);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        const int i = threadIdx.x + j * 256;
        if (i < nv) {
            float4 o;
            o.x = Real<float>::exp(v[j].x - mz);
            o.y = Real<float>::exp(v[j].y - mz);
            o.z = Real<float>::exp(v[j].z - mz);
            o.w = Real<float>::exp(v[j].w - mz);
            q[i] = o;
        }
    }
    if (threadIdx.x == 0) mx[row] = m;
}

// ---- generic batched fp32 GEMM  C[b](m,n) = sum_k A[b](m,k) * B[b](k,n), strided operands ----------
// 64x64 tile, 16-deep k-chunks through shared memory, 256 threads x (4x4) outputs.  Epilogue
// functor Epi(b, m, n, acc) writes the result.
struct Operand {
    const float* p;
    size_t batch;  // elements between batches
    int s_outer;   // stride of the m (A) / n (B) index
    int s_k;       // stride of the k index
};

template <typename Epi, int TN = 64, int KC = 16>
__global__ void __launch_bounds__(256)
joint_gemm_kernel(Operand A, Operand B, int M, int Nn, int Kfull, int slices, Epi epi) {
    // 64 x TN output tile (TN = 32 for skinny right-hand sides), KC-deep k-chunks
    constexpr int PN = TN / 16;  // output columns per thread
    __shared__ float sa[KC][64 + 1], sb[KC][TN + 1];
    // blockIdx.z = batch * slices + k-slice; each slice covers a KC-aligned range of K
    const int b = blockIdx.z / slices, ks = blockIdx.z - b * slices;
    const int kper = ((Kfull + slices - 1) / slices + KC - 1) / KC * KC;
    const int kbeg = ks * kper;
    const int K = min(Kfull, kbeg + kper);
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * TN;
    const float* a = A.p + (size_t)b * A.batch;
    const float* bb = B.p + (size_t)b * B.batch;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads, 4 x PN outputs each
    float acc[4][PN] = {};
    for (int k0 = kbeg; k0 < K; k0 += KC) {
        // cooperative loads; the index order follows whichever stride is 1 so reads coalesce
        for (int i = threadIdx.x; i < 64 * KC; i += 256) {
            int r, kk;
            if (A.s_k == 1) { kk = i % KC; r = i / KC; } else { r = i & 63; kk = i >> 6; }
            const int m = m0 + r, k = k0 + kk;
            sa[kk][r] = (m < M && k < K) ? __ldg(a + (size_t)m * A.s_outer + (size_t)k * A.s_k) : 0.0f;
        }
        for (int i = threadIdx.x; i < TN * KC; i += 256) {
            int r, kk;
            if (B.s_k == 1) { kk = i % KC; r = i / KC; } else { r = i % TN; kk = i / TN; }
            const int n = n0 + r, k = k0 + kk;
            sb[kk][r] = (n < Nn && k < K) ? __ldg(bb + (size_t)n * B.s_outer + (size_t)k * B.s_k) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; ++kk) {
            float av[4], bv[PN];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = sa[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < PN; ++j) bv[j] = sb[kk][tx * PN + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < PN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < PN; ++j) {
            const int m = m0 + ty * 4 + i, n = n0 + tx * PN + j;
            if (m < M && n < Nn) epi(blockIdx.z, m, n, acc[i][j]);
        }
}

// ---- J2: split-K partial sums (deterministic: one slab per K-slice, summed in fixed order) ---------
struct EpiPartial {
    float* part;  // [slices][N,T,U]
    size_t cells; // N*T*U
    int T, U, slices;
    __device__ void operator()(int bz, int t, int u, float acc) const {
        const int b = bz / slices, ks = bz - b * slices;
        part[(size_t)ks * cells + ((size_t)b * T + t) * U + u] = acc;
    }
};

// ---- J2 epilogue: S -> lse, lattice log-prob pair (diagonal-major), keep 1/S -----------------------
struct EpiStats {
    const float *f, *g, *mf, *mg;
    const int *labels, *xlen, *ylen;
    float* inv_s;  // [N,T,U]
    float4* lp2;   // diagonal-major lattice factors (Lat<float>::fac)
    JointDims jd;
    Dims d;        // lattice geometry (maxT = T, maxU = U)
    __device__ void operator()(int b, int t, int u, float S) const {
        int Tb, Ub;
        utt_extent(d, xlen, ylen, b, Tb, Ub);
        const size_t cell = ((size_t)b * jd.T + t) * jd.U + u;
        if (t >= Tb || u >= Ub) {
            inv_s[cell] = 0.0f;
            return;
        }
        const float mft = mf[(size_t)b * jd.T + t], mgu = mg[(size_t)b * jd.U + u];
        const float lse = mft + mgu + logf(S);
        inv_s[cell] = 1.0f / S;
        const float* fr = f + ((size_t)b * jd.T + t) * jd.V;
        const float* gr = g + ((size_t)b * jd.U + u) * jd.V;
        const float lpb = (__ldg(fr + jd.blank) + __ldg(gr + jd.blank)) - lse;
        float lpl = 0.0f;
        const bool has_label = u < Ub - 1;
        if (has_label) {
            const int y = __ldg(labels + (size_t)b * (jd.U > 1 ? jd.U - 1 : 0) + u);
            lpl = (__ldg(fr + y) + __ldg(gr + y)) - lse;
        }
        lp2[skew(d, b, t, u)] = make_fac(lpb, lpl, has_label);
    }
};

__global__ void __launch_bounds__(256)
joint_stats_kernel(const float* __restrict__ part, int slices, const EpiStats epi) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= epi.d.rows) return;
    uint32_t bt, u, b, t;
    epi.d.divU.divmod(r, bt, u);
    epi.d.divT.divmod(bt, b, t);
    float S = 0.0f;
    for (int ks = 0; ks < slices; ++ks) S += part[(size_t)ks * epi.d.rows + r];
    epi((int)b, (int)t, (int)u, S);
}

// ---- J3: per cell weights from the lattices ---------------------------------------------------------
//   Wm = e^{alpha+beta-ll} / S ;  Bk = blank-transition occupancy ;  Lb = label-transition occupancy
__global__ void __launch_bounds__(256)
joint_weights_kernel(const float4* __restrict__ lp2, const LogVal* __restrict__ alphas,
                     const LogVal* __restrict__ betas, const LogVal* __restrict__ llf,
                     const float* __restrict__ inv_s, const int* __restrict__ xlen,
                     const int* __restrict__ ylen, float* __restrict__ Wm, float* __restrict__ Bk,
                     float* __restrict__ Lb, const float scale_in, const float* __restrict__ scale_vec,
                     const Dims d, const int wm_pitch) {
    // Wm rows have `wm_pitch` >= maxU entries (zero beyond maxU: the tensor-core kernel fetches them as aligned
    // float4 rows); Bk / Lb / inv_s are [N,T,maxU].  One thread per Wm entry.
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (uint32_t)d.N * d.maxT * wm_pitch) return;
    const uint32_t bt = q / (uint32_t)wm_pitch, u = q % (uint32_t)wm_pitch;
    if (u >= (uint32_t)d.maxU) {
        Wm[q] = 0.0f;
        return;
    }
    const uint32_t r = bt * d.maxU + u;
    uint32_t b, t;
    d.divT.divmod(bt, b, t);
    int Tb, Ub;
    utt_extent(d, xlen, ylen, b, Tb, Ub);
    float w = 0.0f, bk = 0.0f, lb = 0.0f;
    const float scale = scale_vec ? scale_in * __ldg(scale_vec + b) : scale_in;
    if ((int)t < Tb && (int)u < Ub) {
        // everything in the exp2 domain: log2 occupancy = exact integer part + small float part
        const float4 fc = lp2[skew(d, b, t, u)];
        const size_t q = cell(d, b, t, u);
        const LogVal a = alphas[q], ll = llf[b], bq = betas[q];
        const int oe = a.e - ll.e;
        const float ol = a.l - ll.l;
        const float lpb2 = (float)__float_as_int(fc.y) + log2f(fc.x);   // log2 p_blank
        w = scale * exp2f((float)(oe + bq.e) + (ol + bq.l)) * inv_s[r];
        if ((int)t < Tb - 1) {
            const LogVal bn = betas[q + d.maxU];
            bk = scale * exp2f((float)(oe + bn.e) + (ol + bn.l) + lpb2);
        } else if ((int)u == Ub - 1) {
            bk = scale * exp2f((float)oe + ol + lpb2);
        }
        if ((int)u < Ub - 1) {
            const LogVal bn = betas[q + 1];
            const float lpl2 = (float)__float_as_int(fc.w) + log2f(fc.z);
            lb = scale * exp2f((float)(oe + bn.e) + (ol + bn.l) + lpl2);
        }
    }
    Wm[q] = w;
    Bk[r] = bk;
    Lb[r] = lb;
}

// ---- J4/J5: out[b,r,v] = Eout[b,r,v] * sum_s W(r,s) * Ein[b,s,v]  (thin contraction over s) ----------
//   dF: r = t, s = u, W(r,s) = Wm[b,t,u], Ein = Eg, Eout = Ef
//   dG: r = u, s = t, W(r,s) = Wm[b,t,u] (transposed access), Ein = Ef, Eout = Eg
// One thread per vocabulary column (coalesced over v), RT output rows per block held in registers,
// the W tile broadcast from shared memory: RT FMAs per 4-byte load of Ein.
constexpr int kJointRT = 16, kJointSC = 64;
__global__ void __launch_bounds__(256)
joint_thin_kernel(const float* __restrict__ Wm, int w_stride_r, int w_stride_s, size_t w_batch,
                  const float* __restrict__ Ein, const float* __restrict__ Eout, float* __restrict__ out,
                  int R, int S, int V) {
    __shared__ float sw[kJointSC][kJointRT];
    const int b = blockIdx.z, r0 = blockIdx.y * kJointRT;
    const int v = blockIdx.x * 256 + threadIdx.x;
    const float* w = Wm + (size_t)b * w_batch;
    const float* ein = Ein + (size_t)b * S * V;
    float acc[kJointRT];
#pragma unroll
    for (int i = 0; i < kJointRT; ++i) acc[i] = 0.0f;
    for (int s0 = 0; s0 < S; s0 += kJointSC) {
        for (int i = threadIdx.x; i < kJointSC * kJointRT; i += 256) {
            const int ss = i / kJointRT, rr = i % kJointRT;
            const int s = s0 + ss, r = r0 + rr;
            sw[ss][rr] = (s < S && r < R) ? __ldg(w + (size_t)r * w_stride_r + (size_t)s * w_stride_s) : 0.0f;
        }
        __syncthreads();
        if (v < V) {
            const int smax = min(kJointSC, S - s0);
            for (int ss = 0; ss < smax; ++ss) {
                const float e = __ldg(ein + (size_t)(s0 + ss) * V + v);
#pragma unroll
                for (int i = 0; i < kJointRT; ++i) acc[i] = fmaf(sw[ss][i], e, acc[i]);
            }
        }
        __syncthreads();
    }
    if (v < V) {
#pragma unroll
        for (int i = 0; i < kJointRT; ++i) {
            const int r = r0 + i;
            if (r < R) {
                const size_t o = ((size_t)b * R + r) * V + v;
                out[o] = __ldg(Eout + o) * acc[i];
            }
        }
    }
}

// generic-GEMM form of the same products, used when V is too short to give every thread a column
struct EpiGrad {
    const float* e;  // Ef [N,T,V] (or Eg [N,U,V])
    float* out;      // dF (or dG), same shape
    int rows, V;     // rows per batch (T or U)
    __device__ void operator()(int b, int m, int n, float acc) const {
        const size_t i = ((size_t)b * rows + m) * V + n;
        out[i] = e[i] * acc;
    }
};

// ---- J6: the blank / label terms (sparse in k) -------------------------------------------------------
// One WARP per (b,t) row of dF and per (b,u) row of dG.  The blank term is a warp sum; the label terms are
// one read-modify-write per distinct label: lanes holding the same label are found with a warp match, the
// lowest of them adds the group's values in lane order (deterministic, no atomics) and updates the row.
// Chunks of 32 labels are processed in order with a warp barrier between them.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// row[key] -= sum of val over the lanes with this key (lanes with key < 0 hold nothing); whole warp calls
__device__ __forceinline__ void warp_scatter_sub(float* row, int key, float val) {
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(peers) - 1;
    float acc = 0.0f;
    // every lane walks its own peer set in lane order; shuffles are executed by the whole warp
    for (int src = 0; src < 32; ++src) {
        const float v = __shfl_sync(0xffffffffu, val, src);
        if ((peers >> src) & 1u) acc += v;
    }
    if (key >= 0 && lane == leader) row[key] -= acc;
}

__global__ void __launch_bounds__(128)
joint_sparse_f_kernel(float* __restrict__ dF, const float* __restrict__ Bk, const float* __restrict__ Lb,
                      const int* __restrict__ labels, const int* __restrict__ ylen, const JointDims jd) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (b,t)
    const int lane = threadIdx.x & 31;
    if (i >= jd.N * jd.T) return;
    const int b = i / jd.T;
    const int Ub = min(max(__ldg(ylen + b) + 1, 1), jd.U);
    float* row = dF + (size_t)i * jd.V;
    const float* bk = Bk + (size_t)i * jd.U;
    const float* lb = Lb + (size_t)i * jd.U;
    float sb = 0.0f;
    for (int u = lane; u < Ub; u += 32) sb += bk[u];
    sb = warp_sum(sb);
    if (lane == 0) row[jd.blank] -= sb;
    for (int u0 = 0; u0 < Ub - 1; u0 += 32) {
        __syncwarp();
        const int u = u0 + lane;
        const bool has = u < Ub - 1;
        warp_scatter_sub(row, has ? __ldg(labels + (size_t)b * (jd.U - 1) + u) : -1, has ? lb[u] : 0.0f);
    }
}

__global__ void __launch_bounds__(128)
joint_sparse_g_kernel(float* __restrict__ dG, const float* __restrict__ Bk, const float* __restrict__ Lb,
                      const int* __restrict__ labels, const int* __restrict__ xlen,
                      const int* __restrict__ ylen, const JointDims jd) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // (b,u)
    const int lane = threadIdx.x & 31;
    if (i >= jd.N * jd.U) return;
    const int b = i / jd.U, u = i % jd.U;
    const int Tb = min(max(__ldg(xlen + b), 1), jd.T);
    const int Ub = min(max(__ldg(ylen + b) + 1, 1), jd.U);
    if (u >= Ub) return;
    float* row = dG + (size_t)i * jd.V;
    float sb = 0.0f, sl = 0.0f;
    for (int t = lane; t < Tb; t += 32) {
        const size_t c = ((size_t)b * jd.T + t) * jd.U + u;
        sb += Bk[c];
        sl += Lb[c];
    }
    sb = warp_sum(sb);
    sl = warp_sum(sl);
    if (lane == 0) {
        row[jd.blank] -= sb;
        if (u < Ub - 1) row[__ldg(labels + (size_t)b * (jd.U - 1) + u)] -= sl;
    }
}

}  // namespace b200rnnt
