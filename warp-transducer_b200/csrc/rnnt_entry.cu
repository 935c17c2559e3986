// rnnt_entry.cu — C-ABI of libwarprnnt.so (declared in include/rnnt.h) and the host-side
// orchestration of the three kernels.  Replaces reference src/rnnt_entrypoint.cpp and
// include/detail/gpu_rnnt.h (GpuRNNT<T>::compute_cost_and_score) for loc == RNNT_GPU.
//
// Per call, on options.stream:  [stage host labels/lengths if needed] -> rowstats -> lattice
// (alpha || beta) -> grad -> costs to host + ONE stream synchronise (sync API) / nothing (async API).
// No memset pass, no intermediate host synchronisation.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rnnt.h"
#include "rnnt_chunk.cuh"
#include "rnnt_joint.cuh"
#include "rnnt_kernels.cuh"
#include "rnnt_lattice.cuh"

using namespace b200rnnt;

namespace {

thread_local int g_last_launches = 0;
thread_local bool g_pdl = false;   // launch the dependent kernels of the current call with programmatic stream serialization
thread_local bool g_layout_tunv = false;   // set only inside rnnt_b200_loss_async_layout*

// Optional per-kernel timing (bench.py's roofline leg): when enabled, events are recorded on the
// call's own stream around each of the three kernels, one event set per call (pooled), so a timed
// loop needs no host synchronisation; rnnt_b200_profile_collect() averages them afterwards.
struct EventSet {
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    bool valid[4] = {false, false, false, false};
};
thread_local bool g_profile = false;
thread_local std::vector<EventSet> g_sets;
thread_local size_t g_used = 0;   // sets recorded since the last collect
inline void profile_begin_call() {
    if (!g_profile) return;
    if (g_used == g_sets.size()) g_sets.emplace_back();
    EventSet& es = g_sets[g_used++];
    for (int i = 0; i < 4; ++i) es.valid[i] = false;
}
inline void mark(int i, cudaStream_t s) {
    if (!g_profile || g_used == 0) return;
    EventSet& es = g_sets[g_used - 1];
    if (!es.e[i]) cudaEventCreate(&es.e[i]);
    es.valid[i] = cudaEventRecord(es.e[i], s) == cudaSuccess;
}
// ms of {rowstats, lattice, grad} for one recorded call; false where not measured
inline int read_set(EventSet& es, float* ms3) {
    int n = 0;
    for (int i = 0; i < 3; ++i) {
        ms3[i] = -1.0f;
        if (es.valid[i] && es.valid[i + 1] && cudaEventSynchronize(es.e[i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms3[i], es.e[i], es.e[i + 1]) == cudaSuccess)
            ++n;
    }
    return n;
}

// Dev aid, RNNT_B200_TIMELINE=1: timed events around every launch of the grouped (overlapped) schedule,
// printed to stderr relative to the call's start.  Synchronises the call; never on in production.
struct Timeline {
    std::vector<std::pair<std::string, cudaEvent_t>> ev;
    bool on = false;
    void tick(const char* what, int k, cudaStream_t s) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, s);
        ev.emplace_back(std::string(what) + std::to_string(k), e);
    }
    void dump() {
        if (!on || ev.empty()) return;
        cudaDeviceSynchronize();
        for (auto& p : ev) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[0].second, p.second);
            fprintf(stderr, "[timeline] %-14s %8.3f ms\n", p.first.c_str(), ms);
        }
        for (auto& p : ev) cudaEventDestroy(p.second);
        ev.clear();
    }
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Launch `kernel`; with pdl it may start while the previous kernel of the stream is still running (its
// CTAs block in griddepcontrol.wait until that kernel has completed and its writes are visible).
template <typename... KArgs, typename... Args>
void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}

// cudaFuncSetAttribute once per (kernel, attribute, device) and host thread (again only for a larger value).  Keyed by the kernel's
// address: template instantiations with identical signatures share a function-pointer TYPE, so a static
// flag inside a generic lambda would be shared between them.
void func_attr_once(const void* kernel, cudaFuncAttribute attr, int value) {
    struct Key { const void* k; int attr, dev, value; };
    thread_local std::vector<Key> done;
    int dev = 0;
    cudaGetDevice(&dev);
    for (Key& e : done)
        if (e.k == kernel && e.attr == (int)attr && e.dev == dev) {
            if (value > e.value) {   // e.g. a larger dynamic shared-memory request than any before
                cudaFuncSetAttribute(kernel, attr, value);
                e.value = value;
            }
            return;
        }
    cudaFuncSetAttribute(kernel, attr, value);
    done.push_back(Key{kernel, (int)attr, dev, value});
}

// Side streams for overlapping the (latency-bound, few-SM) lattice kernel of one group of
// utterances with the (bandwidth-bound) streaming passes of the others.  High priority so the
// lattice CTAs are placed as soon as short-lived streaming CTAs retire.  Fork/join with events on
// the caller's stream only, so the call stays capturable and stream-ordered for the caller.
constexpr int kMaxGroups = 8;    // upper bound (RNNT_B200_GROUPS)
constexpr int kAutoGroups = 4;   // what the overlap heuristic picks
struct SidePool {
    cudaStream_t stream[kMaxGroups] = {};
    cudaEvent_t forked[kMaxGroups] = {};
    cudaEvent_t joined[kMaxGroups] = {};
    int device = -1;
    bool ok = false;
};
SidePool& side_pool() {
    // one pool per (host thread, device): a thread that alternates devices finds its pools again instead
    // of re-creating (and leaking) streams and events on every switch
    constexpr int kMaxDevices = 64;
    thread_local SidePool pools[kMaxDevices];
    int dev = 0;
    cudaGetDevice(&dev);
    SidePool& pool = pools[dev >= 0 && dev < kMaxDevices ? dev : 0];
    if (pool.device != dev) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        pool.ok = true;
        for (int g = 0; g < kMaxGroups; ++g) {
            pool.ok &= cudaStreamCreateWithPriority(&pool.stream[g], cudaStreamNonBlocking, hi) == cudaSuccess;
            pool.ok &= cudaEventCreateWithFlags(&pool.forked[g], cudaEventDisableTiming) == cudaSuccess;
            pool.ok &= cudaEventCreateWithFlags(&pool.joined[g], cudaEventDisableTiming) == cudaSuccess;
        }
        pool.device = dev;
    }
    return pool;
}

// ---- workspace carve-up (all sections 256-B aligned) -------------------------------------------
struct Workspace {
    void* stat;     // pair<T>  [rows]   (row max, log sum exp)
    void* lp2;      // Lat<T>::fac [lat]  per-cell transition factors (16 B), diagonal-major
    void* alphas;   // Lat<T>::val [lat]  lat = N*(maxT+maxU-1)*maxU   (8 B: LogVal for fp32 - cell-major, using N*maxT*maxU
                    //                    of the entries -, double for fp64 - diagonal-major)
    void* betas;    // Lat<T>::val [lat]
    void* llf;      // Lat<T>::val [N]
    void* llb;      // Lat<T>::val [N]
    void* costs;    // T [N]
    int* labels;    // staging for host-side integer inputs
    int* ylen;
    int* xlen;
    size_t bytes;
};

Workspace carve(void* base, size_t rows, size_t lat, int N, int maxU, size_t dtype) {
    Workspace w;
    size_t off = align_up(reinterpret_cast<uintptr_t>(base), 256) - reinterpret_cast<uintptr_t>(base);
    char* p = static_cast<char*>(base);
    auto take = [&](size_t n) {
        void* q = p ? p + off : nullptr;
        off = align_up(off + n, 256);
        return q;
    };
    w.stat = take(rows * 2 * dtype);
    w.lp2 = take(lat * 16);
    w.alphas = take(lat * 8);
    w.betas = take(lat * 8);
    w.llf = take((size_t)N * 8);
    w.llb = take((size_t)N * 8);
    w.costs = take(N * dtype);
    w.labels = static_cast<int*>(take((size_t)N * (maxU > 1 ? maxU - 1 : 1) * sizeof(int)));
    w.ylen = static_cast<int*>(take(N * sizeof(int)));
    w.xlen = static_cast<int*>(take(N * sizeof(int)));
    // slack: a base pointer that is not 256-aligned, plus the speculative lattice reads one diagonal past
    // the last utterance (row_grad_setup_spec: at most (maxU + 1) * 8 bytes beyond `betas`)
    w.bytes = off + 256 + 16 * 1024;
    return w;
}

struct DeviceInfo {
    int sms = 0;
};
const DeviceInfo& device_info() {
    thread_local int cached_dev = -1;
    thread_local DeviceInfo info;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&info.sms, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return info;
}

bool is_device_pointer(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}


// ---- streaming-kernel dispatch on (vector width, row length) -------------------------------------
// Long rows: one CTA per row (grid = rows).  Short rows: register tiles, 32/LPR rows per warp.
// Both grids are non-persistent on purpose (see rnnt_kernels.cuh).
template <typename T, int VEC, int NV, typename IO>
void launch_rowstats_row(const IO* acts, const int* labels, const int* xlen, const int* ylen,
                         const Workspace& w, const Dims& d, cudaStream_t s) {
    rowstats_row_kernel<T, VEC, NV, IO><<<d.rows, RowThreads<IO>::value, 0, s>>>(
        acts, labels, xlen, ylen, static_cast<typename Real<T>::pair*>(w.stat),
        static_cast<typename Lat<T>::fac*>(w.lp2), d);
    ++g_last_launches;
}

template <typename T, int VEC, int NV, typename IO>
void launch_grad_row(const IO* acts, IO* grads, const int* labels, const int* xlen, const int* ylen,
                     const Workspace& w, T scale, const T* scale_vec, const Dims& d, cudaStream_t s) {
    auto k = (scale != T(1) || scale_vec) ? grad_row_kernel<T, VEC, NV, true, IO>
                                          : grad_row_kernel<T, VEC, NV, false, IO>;
    launch_k(k, dim3(d.rows), dim3(RowThreads<IO>::value), 0, s, g_pdl, acts, grads, labels, xlen, ylen,
             static_cast<const typename Real<T>::pair*>(w.stat), static_cast<const typename Lat<T>::val*>(w.alphas),
             static_cast<const typename Lat<T>::val*>(w.betas), static_cast<const typename Lat<T>::val*>(w.llf), scale,
             scale_vec, d);
    ++g_last_launches;
}

template <typename T, int VEC, int LPR, typename IO>
void launch_rowstats_tile(const IO* acts, const int* labels, const int* xlen, const int* ylen,
                          const Workspace& w, const Dims& d, cudaStream_t s) {
    const uint64_t warps = ((uint64_t)d.rows * LPR + 31) / 32;
    rowstats_tile_kernel<T, VEC, LPR, IO><<<(unsigned)((warps + 7) / 8), 256, 0, s>>>(
        acts, labels, xlen, ylen, static_cast<typename Real<T>::pair*>(w.stat),
        static_cast<typename Lat<T>::fac*>(w.lp2), d);
    ++g_last_launches;
}

template <typename T, int VEC, int LPR, typename IO>
void launch_grad_tile(const IO* acts, IO* grads, const int* labels, const int* xlen, const int* ylen,
                      const Workspace& w, T scale, const T* scale_vec, const Dims& d, cudaStream_t s) {
    auto k = (scale != T(1) || scale_vec) ? grad_tile_kernel<T, VEC, LPR, true, IO>
                                          : grad_tile_kernel<T, VEC, LPR, false, IO>;
    const uint64_t warps = ((uint64_t)d.rows * LPR + 31) / 32;
    launch_k(k, dim3((unsigned)((warps + 7) / 8)), dim3(256), 0, s, g_pdl, acts, grads, labels, xlen, ylen,
             static_cast<const typename Real<T>::pair*>(w.stat), static_cast<const typename Lat<T>::val*>(w.alphas),
             static_cast<const typename Lat<T>::val*>(w.betas), static_cast<const typename Lat<T>::val*>(w.llf), scale,
             scale_vec, d);
    ++g_last_launches;
}

// lanes per row for the register-tile kernels: the fewest lanes (>= 2) whose kVPL = 8 vector
// registers hold the row - measured on B200: V=28 -> 2 lanes, V=50 (float2) -> 4 lanes per row
inline int pick_lpr(int nv) {
    static const int forced = [] {
        const char* e = getenv("RNNT_B200_LPR");  // tuning hook
        return e ? atoi(e) : 0;
    }();
    int lpr = 2;
    while (lpr < 32 && lpr * kVPL < nv) lpr *= 2;
    if (forced >= 1 && forced <= 32 && (forced & (forced - 1)) == 0 && forced * kVPL >= nv) lpr = forced;
    return lpr;
}

template <typename T, int VEC, typename IO>
void stream_passes(const IO* acts, IO* grads, const int* labels, const int* xlen, const int* ylen,
                   const Workspace& w, T scale, const T* scale_vec, const Dims& d, cudaStream_t s, int pass) {
    const int nv = d.V / VEC;
    if (nv > 32 * kVPL) {  // long rows: CTA per row; NV = vectors per thread per trip
        const int per_thread = (nv + RowThreads<IO>::value - 1) / RowThreads<IO>::value;
#define B200_ROW(NVV)                                                                             \
    do {                                                                                          \
        if (pass == 1) launch_rowstats_row<T, VEC, NVV, IO>(acts, labels, xlen, ylen, w, d, s);       \
        else launch_grad_row<T, VEC, NVV, IO>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s);       \
    } while (0)
        if (sizeof(IO) <= 4 && VEC == 16 / (int)sizeof(IO)) {  // 16-B fast paths get an exact register count
            switch (per_thread) {
                case 1: B200_ROW(1); break;
                case 2: B200_ROW(2); break;
                case 3: B200_ROW(3); break;
                case 4: B200_ROW(4); break;
                case 5: B200_ROW(5); break;
                case 6: B200_ROW(6); break;
                default: B200_ROW(8); break;
            }
        } else {
            if (per_thread <= 2) B200_ROW(2);
            else if (per_thread <= 4) B200_ROW(4);
            else B200_ROW(8);
        }
#undef B200_ROW
        return;
    }
#define B200_TILE(L)                                                                              \
    case L:                                                                                       \
        if (pass == 1) launch_rowstats_tile<T, VEC, L, IO>(acts, labels, xlen, ylen, w, d, s);        \
        else launch_grad_tile<T, VEC, L, IO>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s);        \
        break;
    switch (pick_lpr(nv)) {
        B200_TILE(2)
        B200_TILE(4)
        B200_TILE(8)
        B200_TILE(16)
        B200_TILE(32)
    }
#undef B200_TILE
}

// ---- short rows (<= 512 B): chunk kernels (rnnt_chunk.cuh), TMA bulk staging of R consecutive rows ----
// Threads per row: as FEW as keep a thread's share at <= 32 elements (two for even V, so that the walk can use
// 8-byte pairs) - every lane of a row repeats the row's fixed work (mapping, shuffles, reductions), and at
// V = 28 that fixed work dominated: 4 lanes/row 0.099 ms, 2 lanes/row 0.079 ms for the whole C2 call.  Bank
// conflicts are dealt with by the lane mapping (chunk_walk_cost), not by the lane count.
// RNNT_B200_CHUNK=0 routes short rows to the register-tile kernels.
inline int pick_tpr(int V) {
    int tpr = 1;
    while (tpr < 32 && V > 32 * tpr) tpr *= 2;
    if (tpr == 1 && V % 2 == 0 && V > 16) tpr = 2;
    return tpr;
}
// Shared-memory wavefronts of one warp-wide access of the chunk kernels' element walk under either
// lane -> (row, slice) mapping (ChunkMap in rnnt_chunk.cuh).  A warp's access is served in phases of
// 128 bytes' worth of lanes; within a phase lanes that fall on the same bank group serialise.
inline int chunk_walk_cost(int V, int tpr, int elt, bool hmajor) {
    const int unit = V % 2 == 0 ? 2 * elt : elt;   // bytes per lane access (pairs for even V)
    const int per_phase = 128 / unit;
    const int rpw = 32 / tpr;
    int cost = 0;
    for (int p0 = 0; p0 < 32; p0 += per_phase) {
        int cnt[32] = {0}, worst = 0;
        for (int l = p0; l < p0 + per_phase; ++l) {
            const int il = hmajor ? l % rpw : l / tpr, h = hmajor ? l / rpw : l % tpr;
            const long addr = (long)il * V * elt / unit + h;
            worst = std::max(worst, ++cnt[addr % per_phase]);
        }
        cost += worst;
    }
    return cost;
}
inline bool chunk_enabled() {
    static const bool on = [] { const char* e = getenv("RNNT_B200_CHUNK"); return !(e && atoi(e) == 0); }();
    return on;
}
template <typename T>
bool chunk_pass(const T* acts, T* grads, const int* labels, const int* xlen, const int* ylen,
                const Workspace& w, T scale, const T* scale_vec, const Dims& d, cudaStream_t s, int pass) {
    using Pair = typename Real<T>::pair;
    using Fac = typename Lat<T>::fac;
    using Val = typename Lat<T>::val;
    if (!chunk_enabled() || (size_t)d.V * sizeof(T) > 512) return false;
    if (reinterpret_cast<uintptr_t>(acts) % 16 || (pass == 2 && reinterpret_cast<uintptr_t>(grads) % 16)) return false;
    // threads per chunk CTA (tuning hook RNNT_B200_CHUNK_NT): fp64 always 128
    static const int forced_nt = [] { const char* e = getenv("RNNT_B200_CHUNK_NT"); return e ? atoi(e) : 0; }();
    int nt = sizeof(T) >= 8 ? 128 : 256;
    if (sizeof(T) == 4 && (forced_nt == 64 || forced_nt == 128 || forced_nt == 256)) nt = forced_nt;
    static const int forced_tpr = [] { const char* e = getenv("RNNT_B200_CHUNK_TPR"); return e ? atoi(e) : 0; }();
    int tpr = pick_tpr(d.V);
    if (forced_tpr == 1 || forced_tpr == 2 || forced_tpr == 4 || forced_tpr == 8) tpr = forced_tpr;   // tuning hook
    if (nt / tpr < 4) nt = 4 * tpr;   // a chunk is at least 4 rows (16-byte aligned chunk starts)
    const int rows_per = nt / tpr;
    const unsigned grid = (unsigned)(((uint64_t)d.rows + rows_per - 1) / rows_per);
    const size_t smem = (size_t)rows_per * d.V * sizeof(T);
    const bool scaled = scale != T(1) || scale_vec;
    static const uint32_t wait_ns = [] { const char* e = getenv("RNNT_B200_CHUNK_WAIT_NS"); return e ? (uint32_t)atoi(e) : 2000u; }();
    static const int forced_map = [] { const char* e = getenv("RNNT_B200_CHUNK_MAP"); return e ? atoi(e) : -1; }();
    int hmajor = chunk_walk_cost(d.V, tpr, (int)sizeof(T), true) < chunk_walk_cost(d.V, tpr, (int)sizeof(T), false);
    if (forced_map == 0 || forced_map == 1) hmajor = forced_map;   // tuning hook
    // 8+ chunk CTAs per SM need most of the shared memory: ask for the largest carve-out once per kernel
    auto prefer_smem = [](auto kernel) {
        func_attr_once(reinterpret_cast<const void*>(kernel), cudaFuncAttributePreferredSharedMemoryCarveout,
                       cudaSharedmemCarveoutMaxShared);
        return kernel;
    };
    auto go = [&](auto tpr_c, auto nt_c) {
        constexpr int TPR = decltype(tpr_c)::value, NT = decltype(nt_c)::value;
        if constexpr (NT / TPR >= 4) {
            if (pass == 1)
                prefer_smem(rowstats_chunk_kernel<T, TPR, NT>)<<<grid, NT, smem, s>>>(
                    acts, labels, xlen, ylen, static_cast<Pair*>(w.stat), static_cast<Fac*>(w.lp2), d, hmajor, wait_ns);
            else if (scaled)
                launch_k(prefer_smem(grad_chunk_kernel<T, TPR, NT, true>), dim3(grid), dim3(NT), smem, s, g_pdl, acts, grads,
                         labels, xlen, ylen, static_cast<const Pair*>(w.stat), static_cast<const Val*>(w.alphas),
                         static_cast<const Val*>(w.betas), static_cast<const Val*>(w.llf), scale, scale_vec, d, hmajor, wait_ns);
            else
                launch_k(prefer_smem(grad_chunk_kernel<T, TPR, NT, false>), dim3(grid), dim3(NT), smem, s, g_pdl, acts, grads,
                         labels, xlen, ylen, static_cast<const Pair*>(w.stat), static_cast<const Val*>(w.alphas),
                         static_cast<const Val*>(w.betas), static_cast<const Val*>(w.llf), scale, scale_vec, d, hmajor, wait_ns);
        }
    };
    auto with_rpt = [&](auto tpr_c) {
        if constexpr (sizeof(T) >= 8) {
            go(tpr_c, std::integral_constant<int, 128>{});
        } else {
            if (nt == 64) go(tpr_c, std::integral_constant<int, 64>{});
            else if (nt == 128) go(tpr_c, std::integral_constant<int, 128>{});
            else go(tpr_c, std::integral_constant<int, 256>{});
        }
    };
    switch (tpr) {
        case 1: with_rpt(std::integral_constant<int, 1>{}); break;
        case 2: with_rpt(std::integral_constant<int, 2>{}); break;
        case 4: with_rpt(std::integral_constant<int, 4>{}); break;
        case 8: with_rpt(std::integral_constant<int, 8>{}); break;
        case 16: with_rpt(std::integral_constant<int, 16>{}); break;
        default: with_rpt(std::integral_constant<int, 32>{}); break;
    }
    ++g_last_launches;
    return true;
}

template <typename T, typename IO>
void stream_pass(const IO* acts, IO* grads, const int* labels, const int* xlen, const int* ylen,
                 const Workspace& w, T scale, const T* scale_vec, const Dims& d, cudaStream_t s, int pass) {
    if constexpr (std::is_same<T, IO>::value) {
        if (chunk_pass<T>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s, pass)) return;
    }
    // widest vector the row pitch and the base pointers allow (16-B vectors on the fast path)
    const uintptr_t mis = reinterpret_cast<uintptr_t>(acts) | reinterpret_cast<uintptr_t>(grads) |
                          ((uintptr_t)d.V * sizeof(IO));
    constexpr int kMaxVec = 16 / sizeof(IO);
    if (mis % 16 == 0)
        stream_passes<T, kMaxVec, IO>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s, pass);
    else if (sizeof(IO) == 4 && mis % 8 == 0)
        stream_passes<T, (sizeof(IO) == 4 ? 2 : 1), IO>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s, pass);
    else
        stream_passes<T, 1, IO>(acts, grads, labels, xlen, ylen, w, scale, scale_vec, d, s, pass);
}

// What a call does.  The reference API is FULL (stats -> lattice -> grad in one call); the
// operator splits a training step into FORWARD (stats + both lattices, costs out) and BACKWARD
// (gradient pass only, reading the lattices the forward left in the workspace).
enum Phase { kFull = 0, kForward = 1, kBackward = 2 };

template <typename IO>
rnntStatus_t run(const IO* acts, IO* grads, const int* labels, const int* ylen, const int* xlen,
                 int V, int N, typename ComputeOf<IO>::type* costs, bool async,
                 typename ComputeOf<IO>::type scale, const typename ComputeOf<IO>::type* scale_vec,
                 Phase phase, bool want_beta, void* workspace, rnntOptions opt) {
    using T = typename ComputeOf<IO>::type;  // arithmetic type (float for the 16-bit storage types)
    if (acts == nullptr || labels == nullptr || ylen == nullptr || xlen == nullptr ||
        (costs == nullptr && phase != kBackward) || workspace == nullptr || V <= 0 || N <= 0 ||
        opt.maxT <= 0 || opt.maxU <= 0 || (phase == kBackward && grads == nullptr))
        return RNNT_STATUS_INVALID_VALUE;  // reference src/rnnt_entrypoint.cpp:49-59
    if (opt.loc == RNNT_CPU) {
        fprintf(stderr, "b200-rnnt: CPU execution requested, but this library is the CUDA path only\n");
        return RNNT_STATUS_EXECUTION_FAILED;
    }
    if (opt.loc != RNNT_GPU) return RNNT_STATUS_INVALID_VALUE;  // :90-92
    const uint64_t rows64 = (uint64_t)N * opt.maxT * opt.maxU;
    if (rows64 >= (1ull << 31) || opt.blank_label < 0 || opt.blank_label >= V)
        return RNNT_STATUS_INVALID_VALUE;
    if (opt.maxU > 1024) {
        fprintf(stderr, "b200-rnnt: maxU > 1024 is not supported (the reference launches maxU threads per block and has the same limit)\n");
        return RNNT_STATUS_INVALID_VALUE;
    }

    g_last_launches = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(opt.stream);
    const size_t lat = (size_t)N * (opt.maxT + opt.maxU - 1) * opt.maxU;
    Workspace w = carve(workspace, rows64, lat, N, opt.maxU, sizeof(T));

    Dims d;
    d.N = N;
    d.maxT = opt.maxT;
    d.maxU = opt.maxU;
    d.V = V;
    d.blank = opt.blank_label;
    d.rows = (uint32_t)rows64;
    d.divU = FastDiv(opt.maxU);
    d.divT = FastDiv(opt.maxT);
    d.divN = FastDiv(N);
    // options.batch_first is IGNORED on purpose, exactly as the reference's GPU path ignores it
    // (include/detail/gpu_rnnt_kernel.h:7): the reference's own tests/test_gpu.cu:42-50 and
    // tests/test_time.cu pass it zero-initialised (false) with [N,T,U,V] data, so honouring or
    // rejecting it here would break every caller written against the reference.  The [T,U,N,V]
    // layout the reference's CPU path indexes (include/detail/cpu_rnnt.h:139-144) is available through
    // the explicit extension entry rnnt_b200_loss_async_layout (layout = RNNT_B200_LAYOUT_TUNV).
    d.tmajor = g_layout_tunv ? 1 : 0;

    // Integer inputs: device pointers are used in place; host pointers (the header's literal
    // contract, reference include/rnnt.h:84-89) are staged through the workspace.
    if (!async) {
        const size_t nl = (size_t)N * (opt.maxU > 1 ? opt.maxU - 1 : 0);
        if (!is_device_pointer(labels)) {
            if (nl && cudaMemcpyAsync(w.labels, labels, nl * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            labels = w.labels;
        }
        if (!is_device_pointer(ylen)) {
            if (cudaMemcpyAsync(w.ylen, ylen, N * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            ylen = w.ylen;
        }
        if (!is_device_pointer(xlen)) {
            if (cudaMemcpyAsync(w.xlen, xlen, N * sizeof(int), cudaMemcpyHostToDevice, s) != cudaSuccess)
                return RNNT_STATUS_MEMOPS_FAILED;
            xlen = w.xlen;
        }
    }

    // ---- batch groups: utterances are independent, every array is batch-major, so a group is the
    // same problem on offset pointers.  One group = the plain in-order pipeline.
    const size_t cell_stride = (size_t)opt.maxT * opt.maxU;            // rows per utterance
    const size_t lat_stride = (size_t)(opt.maxT + opt.maxU - 1) * opt.maxU;
    const size_t lab_stride = opt.maxU > 1 ? opt.maxU - 1 : 0;
    using Pair = typename Real<T>::pair;
    T* cdev = async ? costs : static_cast<T*>(w.costs);
    struct Group {
        Workspace w;
        Dims d;
        const IO* acts;
        IO* grads;
        const int *labels, *xlen, *ylen;
        T* costs;
        const T* scale_vec;
    };
    auto make_group = [&](int b0, int nb) {
        Group g;
        g.w = w;
        g.w.stat = static_cast<Pair*>(w.stat) + (size_t)b0 * cell_stride;
        g.w.lp2 = static_cast<char*>(w.lp2) + (size_t)b0 * lat_stride * 16;
        // fp32 lattices are cell-major [b][t][u], fp64 ones diagonal-major (rnnt_kernels.cuh: cell / skew)
        const size_t val_stride = sizeof(T) == 4 ? cell_stride : lat_stride;
        g.w.alphas = static_cast<char*>(w.alphas) + (size_t)b0 * val_stride * 8;
        g.w.betas = static_cast<char*>(w.betas) + (size_t)b0 * val_stride * 8;
        g.w.llf = static_cast<char*>(w.llf) + (size_t)b0 * 8;
        g.w.llb = static_cast<char*>(w.llb) + (size_t)b0 * 8;
        g.d = d;
        g.d.N = nb;
        g.d.rows = (uint32_t)((size_t)nb * cell_stride);
        g.acts = acts + (size_t)b0 * cell_stride * V;
        g.grads = grads ? grads + (size_t)b0 * cell_stride * V : nullptr;
        g.labels = labels + (size_t)b0 * lab_stride;
        g.xlen = xlen + b0;
        g.ylen = ylen + b0;
        g.costs = cdev ? cdev + b0 : nullptr;
        g.scale_vec = scale_vec ? scale_vec + b0 : nullptr;
        return g;
    };
    const bool with_beta = grads || want_beta;
    bool co_running = false;   // set when the lattice of one group shares the GPU with the streaming passes of others
    auto launch_lattice = [&](const Group& g, cudaStream_t st) {
        const int threads = (opt.maxU + 31) / 32 * 32;
        dim3 grid(g.d.N, with_beta ? 2 : 1);
        if constexpr (sizeof(T) == 4) {
            // fp32: linear-domain wavefront with explicit exponents, COLS columns per lane (rnnt_lattice.cuh)
            const int lthreads = lattice_threads(opt.maxU);
            const int depth = lattice_ring_depth(opt.maxU, co_running);
            const size_t ring = lattice_ring_bytes(opt.maxU, depth);
            auto launch = [&](auto kernel, int static_smem) {
                if (ring + static_smem > 48 * 1024)
                    func_attr_once(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
                launch_k(kernel, grid, dim3(lthreads), ring, st, g_pdl, static_cast<const float4*>(g.w.lp2), g.xlen, g.ylen,
                         static_cast<LogVal*>(g.w.alphas), static_cast<LogVal*>(g.w.betas), static_cast<LogVal*>(g.w.llf),
                         static_cast<LogVal*>(g.w.llb), g.costs, g.d);
            };
            if (opt.maxU <= 32) launch(lattice_lin_kernel<1, false, 8>, 64);
            else if (opt.maxU <= 64) launch(lattice_lin_kernel<2, false, 8>, 64);
            else if (depth == 32) launch(lattice_lin_kernel<1, true, 32>, kLinStaticSmem);
            else if (depth == 16) launch(lattice_lin_kernel<1, true, 16>, kLinStaticSmem);
            else launch(lattice_lin_kernel<1, true, 8>, kLinStaticSmem);
        } else {
            // fp64: log-domain wavefront (rnnt_kernels.cuh)
            const size_t ring = (size_t)kRing * threads * sizeof(double2);
            auto launch = [&](auto kernel) {
                // opt-in when static + dynamic shared memory exceed the 48 KB default (maxU >= 353);
                // the attribute is per device and the call is rare and cheap next to such a wavefront
                if (ring + kLatticeStaticSmem > 48 * 1024)
                    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
                launch_k(kernel, grid, dim3(threads), ring, st, g_pdl, static_cast<const double2*>(g.w.lp2), g.xlen, g.ylen,
                         static_cast<double*>(g.w.alphas), static_cast<double*>(g.w.betas), static_cast<double*>(g.w.llf),
                         static_cast<double*>(g.w.llb), g.costs, g.d);
            };
            if (threads > 32) launch(lattice_kernel<double, true>);
            else launch(lattice_kernel<double, false>);
        }
        ++g_last_launches;
    };

    // Overlap decision: the wavefront costs ~0.25 us per anti-diagonal whatever the batch; the
    // streaming passes cost 12 B/elt at ~6.9 TB/s.  Worth splitting only when the lattice is a
    // visible share of a call that is long enough to amortise the extra launches.
    // Measured on B200 (N=64,T=1500,U=301,V=50): 3.88 -> 3.77 ms with 4 groups; the co-running
    // lattice CTAs slow the streaming kernels a little, so the gain is modest, and there is none
    // when no pass 2 follows in the same call (loss-only / operator forward) - those stay in order.
    int groups = 1;
    if (phase == kFull && grads && N >= 2 * kAutoGroups && !d.tmajor) {
        static const int forced = [] { const char* e = getenv("RNNT_B200_GROUPS"); return e ? atoi(e) : 0; }();
        const double lattice_us = 0.25 * (opt.maxT + opt.maxU) + 20.0;
        const double stream_us = (double)rows64 * V * sizeof(IO) * (grads ? 3.0 : 1.0) / 6.9e6;
        if (stream_us > 400.0 && lattice_us > 0.08 * stream_us) groups = kAutoGroups;
        if (forced >= 1 && forced <= kMaxGroups && forced <= N) groups = forced;
        if (groups > 1 && !side_pool().ok) groups = 1;
    }

    profile_begin_call();
    mark(0, s);
    // EXPERIMENTAL, off unless RNNT_B200_PDL=1: launch the lattice and gradient kernels with programmatic
    // stream serialization, so their prologues (and the gradient kernel's first wave of logit loads) overlap
    // the tail of the kernel before them; each waits (griddepcontrol.wait) before it touches that kernel's
    // output.  Measured on B200: C2 0.128 vs 0.130 ms, C3/C4 unchanged - the three kernels are each bound by
    // their own latency chains, not by the launch gaps - and one 16-bit parity case differed, so it stays
    // opt-in.  (The event markers of the profiling mode would serialise the kernels anyway.)
    static const bool pdl_env = [] { const char* e = getenv("RNNT_B200_PDL"); return e && atoi(e) != 0; }();
    g_pdl = pdl_env && groups == 1 && !g_profile;
    if (groups == 1) {
        const Group g = make_group(0, N);
        if (phase != kBackward) {
            // pass 1: log-softmax statistics + (blank, label) log-prob gather
            stream_pass<T, IO>(g.acts, nullptr, g.labels, g.xlen, g.ylen, g.w, scale, g.scale_vec, g.d, s, 1);
            mark(1, s);
            // lattice: alpha (and beta when gradients are or will be wanted)
            launch_lattice(g, s);
        } else {
            mark(1, s);
        }
        mark(2, s);
        // pass 2: dense gradient (+ zeros on padding)
        if (grads && phase != kForward) {
            stream_pass<T, IO>(g.acts, g.grads, g.labels, g.xlen, g.ylen, g.w, scale, g.scale_vec, g.d, s, 2);
            mark(3, s);
        }
    } else {
        SidePool& pool = side_pool();
        co_running = true;
        bool fork_ok = true;   // a failed fork/join would leave the gradient pass unordered against a lattice
        static const bool tl_env = [] { const char* e = getenv("RNNT_B200_TIMELINE"); return e && atoi(e) != 0; }();
        Timeline tl;
        tl.on = tl_env;
        tl.tick("start", 0, s);
        Group gs[kMaxGroups];
        for (int k = 0; k < groups; ++k) {
            const int b0 = (int)((int64_t)N * k / groups), b1 = (int)((int64_t)N * (k + 1) / groups);
            gs[k] = make_group(b0, b1 - b0);
        }
        // main stream: pass 1 of every group back to back; each group's lattice forks off behind it
        for (int k = 0; k < groups; ++k) {
            stream_pass<T, IO>(gs[k].acts, nullptr, gs[k].labels, gs[k].xlen, gs[k].ylen, gs[k].w, scale,
                               gs[k].scale_vec, gs[k].d, s, 1);
            tl.tick("rowstats_end", k, s);
            fork_ok &= cudaEventRecord(pool.forked[k], s) == cudaSuccess;
            fork_ok &= cudaStreamWaitEvent(pool.stream[k], pool.forked[k], 0) == cudaSuccess;
            tl.tick("lattice_beg", k, pool.stream[k]);
            launch_lattice(gs[k], pool.stream[k]);
            tl.tick("lattice_end", k, pool.stream[k]);
            fork_ok &= cudaEventRecord(pool.joined[k], pool.stream[k]) == cudaSuccess;
        }
        mark(1, s);
        // main stream: join each lattice, then that group's pass 2 (or just join, forward-only)
        for (int k = 0; k < groups; ++k) {
            fork_ok &= cudaStreamWaitEvent(s, pool.joined[k], 0) == cudaSuccess;
            if (k == 0) mark(2, s);
            tl.tick("grad_beg", k, s);
            if (grads && phase != kForward)
                stream_pass<T, IO>(gs[k].acts, gs[k].grads, gs[k].labels, gs[k].xlen, gs[k].ylen, gs[k].w,
                                   scale, gs[k].scale_vec, gs[k].d, s, 2);
            tl.tick("grad_end", k, s);
        }
        tl.dump();
        if (grads && phase != kForward) mark(3, s);
        if (!fork_ok) {
            cudaStreamSynchronize(s);
            return RNNT_STATUS_EXECUTION_FAILED;
        }
    }

    g_pdl = false;
    if (cudaGetLastError() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    if (async) return RNNT_STATUS_SUCCESS;

    // costs to the caller (host memory in the reference contract) and the call's single sync
    const cudaMemcpyKind kind = is_device_pointer(costs) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    if (cudaMemcpyAsync(costs, w.costs, N * sizeof(T), kind, s) != cudaSuccess)
        return RNNT_STATUS_MEMOPS_FAILED;
    if (cudaStreamSynchronize(s) != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    return RNNT_STATUS_SUCCESS;
}

// ---- additive-joint variant (rnnt_joint.cuh) ----------------------------------------------------
constexpr int kJointSlices = 16;  // max split-K slabs of the S = Ef.Eg^T contraction (K = V is the long axis)
inline int joint_slices(int V) {
    static const int forced = [] { const char* e = getenv("RNNT_B200_JOINT_SLICES"); return e ? atoi(e) : 0; }();
    if (forced >= 1 && forced <= kJointSlices) return forced;   // tuning hook
    return std::max(1, std::min(kJointSlices, V / 320));
}
struct JointWorkspace {
    float *ef, *eg, *mf, *mg, *inv_s, *wm, *bk, *lb, *part;
    float4* lp2;
    LogVal *alphas, *betas, *llf, *llb;
    size_t bytes;
};
JointWorkspace carve_joint(void* base, int N, int T, int U, int V) {
    JointWorkspace w;
    size_t off = align_up(reinterpret_cast<uintptr_t>(base), 256) - reinterpret_cast<uintptr_t>(base);
    char* p = static_cast<char*>(base);
    auto take = [&](size_t n) {
        void* q = p ? p + off : nullptr;
        off = align_up(off + n, 256);
        return q;
    };
    const size_t C = (size_t)N * T * U, D = (size_t)N * (T + U - 1) * U;
    w.ef = static_cast<float*>(take((size_t)N * T * V * 4));
    w.eg = static_cast<float*>(take((size_t)N * U * V * 4));
    w.mf = static_cast<float*>(take((size_t)N * T * 4));
    w.mg = static_cast<float*>(take((size_t)N * U * 4));
    w.inv_s = static_cast<float*>(take(C * 4));
    w.wm = static_cast<float*>(take(std::max(C, (size_t)N * T * umma::kWmPad) * 4));   // rows padded for the fused gradient kernel
    w.bk = static_cast<float*>(take(C * 4));
    w.lb = static_cast<float*>(take(C * 4));
    w.part = static_cast<float*>(take(C * 4 * kJointSlices));
    w.lp2 = static_cast<float4*>(take(D * 16));
    w.alphas = static_cast<LogVal*>(take(D * 8));
    w.betas = static_cast<LogVal*>(take(D * 8));
    w.llf = static_cast<LogVal*>(take((size_t)N * 8));
    w.llb = static_cast<LogVal*>(take((size_t)N * 8));
    w.bytes = off + 256;
    return w;
}

// ---- tensor-core contractions of the additive joint (rnnt_umma.cuh) ---------------------------------
// N (accumulator columns per CTA) is the smallest instantiated width that holds `n`, tiled beyond 256.
inline bool joint_umma_enabled() {
    static const bool on = [] { const char* e = getenv("RNNT_B200_JOINT_SIMT"); return !(e && atoi(e) != 0); }();
    return on;
}
template <int A_MODE, int B_MODE, int KS>
void launch_umma(const umma::Operand& A, const umma::Operand& B, int m, int n, int K, int slices, int batch,
                 const umma::Epilogue& epi, cudaStream_t s, int max_tile = 256) {
    auto go = [&](auto kernel, int NT, size_t smem) {
        func_attr_once(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid((unsigned)(slices * ((n + NT - 1) / NT)), (unsigned)((m + 127) / 128), (unsigned)batch);
        kernel<<<grid, umma::kThreads, smem, s>>>(A, B, K, slices, epi);
    };
    if (n <= 32) go(umma::gemm_kernel<A_MODE, B_MODE, 32, KS>, 32, umma::gemm_smem_bytes<32, KS>());
    else if (n <= 64 || max_tile <= 64) go(umma::gemm_kernel<A_MODE, B_MODE, 64, KS>, 64, umma::gemm_smem_bytes<64, KS>());
    else if (n <= 128) go(umma::gemm_kernel<A_MODE, B_MODE, 128, KS>, 128, umma::gemm_smem_bytes<128, KS>());
    else if (n <= 192) go(umma::gemm_kernel<A_MODE, B_MODE, 192, KS>, 192, umma::gemm_smem_bytes<192, KS>());
    else go(umma::gemm_kernel<A_MODE, B_MODE, 256, KS>, 256, umma::gemm_smem_bytes<256, KS>());
}

rnntStatus_t run_add_joint(const float* f, const float* g, float* dF, float* dG, const int* labels,
                           const int* ylen, const int* xlen, int V, int N, float* costs, float scale,
                           const float* scale_vec, Phase phase, bool want_beta, void* workspace,
                           rnntOptions opt) {
    if (!f || !g || !labels || !ylen || !xlen || (!costs && phase != kBackward) || !workspace || V <= 0 ||
        N <= 0 || opt.maxT <= 0 || opt.maxU <= 0 || (dF == nullptr) != (dG == nullptr) ||
        (phase == kBackward && !dF))
        return RNNT_STATUS_INVALID_VALUE;
    if (opt.loc != RNNT_GPU) return opt.loc == RNNT_CPU ? RNNT_STATUS_EXECUTION_FAILED : RNNT_STATUS_INVALID_VALUE;
    const int T = opt.maxT, U = opt.maxU;
    const uint64_t rows64 = (uint64_t)N * T * U;
    if (rows64 >= (1ull << 31) || U > 1024 || opt.blank_label < 0 || opt.blank_label >= V)
        return RNNT_STATUS_INVALID_VALUE;
    // the contraction kernels index inside one utterance's factor with 32-bit offsets
    if ((uint64_t)std::max(T, U) * (uint64_t)V >= (1ull << 31)) return RNNT_STATUS_INVALID_VALUE;
    g_last_launches = 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(opt.stream);
    JointWorkspace w = carve_joint(workspace, N, T, U, V);
    JointDims jd{N, T, U, V, opt.blank_label};
    Dims d;
    d.N = N;
    d.maxT = T;
    d.maxU = U;
    d.V = V;
    d.blank = opt.blank_label;
    d.rows = (uint32_t)rows64;
    d.divU = FastDiv(U);
    d.divT = FastDiv(T);
    d.divN = FastDiv(N);
    d.tmajor = 0;
    const bool want_grad = dF != nullptr && phase != kForward;
    const bool with_beta = dF != nullptr || want_beta;

    if (phase != kBackward) {
    // J1: factor-wise max and exponentials
    auto prep = [&](const float* x, float* e, float* mx, int rows) {
        const int per = (V / 4 + 255) / 256;   // float4 per thread when one CTA owns a row
        const bool vec = V % 4 == 0 && per <= 8 && V >= 1024 && reinterpret_cast<uintptr_t>(x) % 16 == 0;
        if (!vec) joint_prep_kernel<<<(rows + 7) / 8, 256, 0, s>>>(x, e, mx, rows, V);
        else if (per <= 1) joint_prep_row_kernel<1><<<rows, 256, 0, s>>>(x, e, mx, V);
        else if (per <= 2) joint_prep_row_kernel<2><<<rows, 256, 0, s>>>(x, e, mx, V);
        else if (per <= 4) joint_prep_row_kernel<4><<<rows, 256, 0, s>>>(x, e, mx, V);
        else if (per <= 5) joint_prep_row_kernel<5><<<rows, 256, 0, s>>>(x, e, mx, V);
        else joint_prep_row_kernel<8><<<rows, 256, 0, s>>>(x, e, mx, V);
    };
    prep(f, w.ef, w.mf, N * T);
    prep(g, w.eg, w.mg, N * U);
    // J2: S = Ef . Eg^T in kJointSlices deterministic K-slabs, then lse + lattice log-prob pairs
    {
        const int slices = joint_slices(V);
        if (joint_umma_enabled()) {
            // tcgen05: M = t (tiles of 128), N = u, K = v split into `slices` slabs
            umma::Operand A{w.ef, (long long)T * V, V, 1, T}, B{w.eg, (long long)U * V, V, 1, U};
            // float4 operand fetches when every row of Ef / Eg starts on a 16-byte boundary
            // partial sums of slice ks land in slab ks: part[ks][b][t][u]
            const umma::Epilogue epi{nullptr, 0, 0, 0, w.part, (long long)rows64, (long long)T * U, U, 1};
            if (V % 4 == 0)
                launch_umma<2, 2, 32>(A, B, T, U, V, slices, N, epi, s);
            else
                launch_umma<1, 1, 32>(A, B, T, U, V, slices, N, epi, s);
        } else {
            Operand A{w.ef, (size_t)T * V, V, 1}, B{w.eg, (size_t)U * V, V, 1};
            dim3 grid((U + 63) / 64, (T + 63) / 64, N * slices);
            if (U <= 32) {
                grid.x = (U + 31) / 32;
                joint_gemm_kernel<EpiPartial, 32, 32><<<grid, 256, 0, s>>>(
                    A, B, T, U, V, slices, EpiPartial{w.part, (size_t)rows64, T, U, slices});
            } else {
                joint_gemm_kernel<EpiPartial, 64, 32><<<grid, 256, 0, s>>>(
                    A, B, T, U, V, slices, EpiPartial{w.part, (size_t)rows64, T, U, slices});
            }
        }
        EpiStats epi{f, g, w.mf, w.mg, labels, xlen, ylen, w.inv_s, w.lp2, jd, d};
        joint_stats_kernel<<<(d.rows + 255) / 256, 256, 0, s>>>(w.part, slices, epi);
    }
    // lattice (same kernel as the dense path)
    {
        dim3 grid(N, with_beta ? 2 : 1);
        const int lthreads = lattice_threads(U);
        const size_t ring = lattice_ring_bytes(U, 8);
        auto launch = [&](auto kernel, int static_smem) {
            if (ring + static_smem > 48 * 1024)
                func_attr_once(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring);
            kernel<<<grid, lthreads, ring, s>>>(w.lp2, xlen, ylen, w.alphas, w.betas, w.llf, w.llb, costs, d);
        };
        if (U <= 32) launch(lattice_lin_kernel<1, false, 8>, 64);
        else if (U <= 64) launch(lattice_lin_kernel<2, false, 8>, 64);
        else launch(lattice_lin_kernel<1, true, 8>, kLinStaticSmem);
    }
    g_last_launches += 5;
    }  // phase != kBackward
    if (want_grad) {
        static const bool fused = [] { const char* e = getenv("RNNT_B200_JOINT_FUSED"); return !(e && atoi(e) == 0); }();
        const bool use_fused = fused && joint_umma_enabled() && U <= umma::kWmPad &&
                               (uint64_t)N * T * umma::kWmPad < (1ull << 31);   // padded weights are indexed with 32 bits
        const int wm_pitch = use_fused ? umma::kWmPad : U;
        const unsigned wm_entries = (unsigned)N * T * wm_pitch;
        joint_weights_kernel<<<(wm_entries + 255) / 256, 256, 0, s>>>(w.lp2, w.alphas, w.betas, w.llf, w.inv_s, xlen, ylen,
                                                                     w.wm, w.bk, w.lb, scale, scale_vec, d, wm_pitch);
        if (joint_umma_enabled()) {
            // tcgen05, vocabulary index on the accumulator lanes (coalesced epilogue):
            //   dF[t,v] = Ef[t,v] * sum_u Eg[u,v] Wm[t,u]      M = v, N = t, K = u
            //   dG[u,v] = Eg[u,v] * sum_t Ef[t,v] Wm[t,u]      M = v, N = u, K = t
            umma::Operand EgT{w.eg, (long long)U * V, 1, V, V}, EfT{w.ef, (long long)T * V, 1, V, V};
            umma::Operand WmTU{w.wm, (long long)T * U, U, 1, T};   // (n = t, k = u): k-contiguous
            umma::Operand WmUT{w.wm, (long long)T * U, 1, U, U};   // (n = u, k = t): n-contiguous
            // 64-column accumulator tiles: small tensor-memory / shared-memory footprint -> several CTAs per SM,
            // and the whole tile's Ef fetches are in flight before the accumulator is read
            static const int df_tile = [] { const char* e = getenv("RNNT_B200_DF_TILE"); return e ? atoi(e) : 64; }();
            if (use_fused) {
                // both contractions in one pass over Ef (rnnt_umma.cuh: grad_fused_kernel)
                const umma::GradFused gf{w.ef, w.eg, w.wm, dF, dG, T, U, V};
                static const size_t pad = [] { const char* e = getenv("RNNT_B200_FUSED_PAD_SMEM"); return e ? (size_t)atoi(e) : (size_t)0; }();
                const size_t smem = umma::GradFusedGeom<32, 32>::total + pad;   // pad: residency experiment hook
                auto go = [&](auto kernel) {
                    func_attr_once(reinterpret_cast<const void*>(kernel), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kernel<<<dim3((unsigned)((V + 127) / 128), 1, (unsigned)N), umma::kThreads, smem, s>>>(gf);
                };
                if (V % 4 == 0) go(umma::grad_fused_kernel<32, 32, 3>);
                else go(umma::grad_fused_kernel<32, 32, 0>);
            } else {
            const umma::Epilogue epf{w.ef, (long long)T * V, 1, V, dF, 0, (long long)T * V, 1, V};
            if (V % 4 == 0) launch_umma<3, 1, 24>(EgT, WmTU, V, T, U, 1, N, epf, s, df_tile);   // 16-byte aligned rows
            else launch_umma<0, 1, 24>(EgT, WmTU, V, T, U, 1, N, epf, s, df_tile);
            const umma::Epilogue epg{w.eg, (long long)U * V, 1, V, dG, 0, (long long)U * V, 1, V};
            if (V % 4 == 0) launch_umma<3, 0, 24>(EfT, WmUT, V, U, T, 1, N, epg, s);
            else launch_umma<0, 0, 24>(EfT, WmUT, V, U, T, 1, N, epg, s);
            }
        } else if (V >= 512) {  // long vocabulary: one thread per column, thin contraction
            {   // dF[t,v] = Ef[t,v] * sum_u Wm[t,u] Eg[u,v]
                dim3 grid((V + 255) / 256, (T + kJointRT - 1) / kJointRT, N);
                joint_thin_kernel<<<grid, 256, 0, s>>>(w.wm, U, 1, (size_t)T * U, w.eg, w.ef, dF, T, U, V);
            }
            {   // dG[u,v] = Eg[u,v] * sum_t Wm[t,u] Ef[t,v]
                dim3 grid((V + 255) / 256, (U + kJointRT - 1) / kJointRT, N);
                joint_thin_kernel<<<grid, 256, 0, s>>>(w.wm, 1, U, (size_t)T * U, w.ef, w.eg, dG, U, T, V);
            }
        } else {  // short vocabulary: tiled GEMM
            {
                Operand A{w.wm, (size_t)T * U, U, 1}, B{w.eg, (size_t)U * V, 1, V};
                dim3 grid((V + 63) / 64, (T + 63) / 64, N);
                joint_gemm_kernel<EpiGrad><<<grid, 256, 0, s>>>(A, B, T, V, U, 1, EpiGrad{w.ef, dF, T, V});
            }
            {
                Operand A{w.wm, (size_t)T * U, 1, U}, B{w.ef, (size_t)T * V, 1, V};
                dim3 grid((V + 63) / 64, (U + 63) / 64, N);
                joint_gemm_kernel<EpiGrad><<<grid, 256, 0, s>>>(A, B, U, V, T, 1, EpiGrad{w.eg, dG, U, V});
            }
        }
        joint_sparse_f_kernel<<<(N * T + 3) / 4, 128, 0, s>>>(dF, w.bk, w.lb, labels, ylen, jd);
        joint_sparse_g_kernel<<<(N * U + 3) / 4, 128, 0, s>>>(dG, w.bk, w.lb, labels, xlen, ylen, jd);
        g_last_launches += 5;
    }
    return cudaGetLastError() == cudaSuccess ? RNNT_STATUS_SUCCESS : RNNT_STATUS_EXECUTION_FAILED;
}

}  // namespace

extern "C" {

int get_warprnnt_version(void) { return 1; }

const char* rnntGetStatusString(rnntStatus_t status) {
    switch (status) {
        case RNNT_STATUS_SUCCESS: return "no error";
        case RNNT_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
        case RNNT_STATUS_INVALID_VALUE: return "invalid value";
        case RNNT_STATUS_EXECUTION_FAILED: return "execution failed";
        case RNNT_STATUS_UNKNOWN_ERROR:
        default: return "unknown error";
    }
}

rnntStatus_t compute_rnnt_loss(const float* const activations, float* gradients,
                               const int* const flat_labels, const int* const label_lengths,
                               const int* const input_lengths, int alphabet_size, int minibatch,
                               float* costs, void* workspace, rnntOptions options) {
    return run<float>(activations, gradients, flat_labels, label_lengths, input_lengths,
                      alphabet_size, minibatch, costs, false, 1.0f, nullptr, kFull, false, workspace, options);
}

rnntStatus_t compute_rnnt_loss_fp64(const double* const activations, double* gradients,
                                    const int* const flat_labels,
                                    const int* const label_lengths,
                                    const int* const input_lengths, int alphabet_size,
                                    int minibatch, double* costs, void* workspace,
                                    rnntOptions options) {
    return run<double>(activations, gradients, flat_labels, label_lengths, input_lengths,
                       alphabet_size, minibatch, costs, false, 1.0, nullptr, kFull, false, workspace, options);
}

rnntStatus_t compute_rnnt_loss_async(const float* const activations, float* gradients,
                                     const int* const flat_labels,
                                     const int* const label_lengths,
                                     const int* const input_lengths, int alphabet_size,
                                     int minibatch, float* costs_device, float grad_scale,
                                     void* workspace, rnntOptions options) {
    return run<float>(activations, gradients, flat_labels, label_lengths, input_lengths,
                      alphabet_size, minibatch, costs_device, true, grad_scale, nullptr, kFull, false, workspace, options);
}

rnntStatus_t compute_rnnt_loss_async_fp64(const double* const activations, double* gradients,
                                          const int* const flat_labels,
                                          const int* const label_lengths,
                                          const int* const input_lengths, int alphabet_size,
                                          int minibatch, double* costs_device, double grad_scale,
                                          void* workspace, rnntOptions options) {
    return run<double>(activations, gradients, flat_labels, label_lengths, input_lengths,
                       alphabet_size, minibatch, costs_device, true, grad_scale, nullptr, kFull, false, workspace, options);
}

rnntStatus_t rnnt_b200_forward(const float* const activations, const int* const flat_labels,
                               const int* const label_lengths, const int* const input_lengths,
                               int alphabet_size, int minibatch, float* costs_device,
                               int prepare_backward, void* workspace, rnntOptions options) {
    return run<float>(activations, nullptr, flat_labels, label_lengths, input_lengths, alphabet_size,
                      minibatch, costs_device, true, 1.0f, nullptr, kForward, prepare_backward != 0,
                      workspace, options);
}

rnntStatus_t rnnt_b200_forward_fp64(const double* const activations, const int* const flat_labels,
                                    const int* const label_lengths, const int* const input_lengths,
                                    int alphabet_size, int minibatch, double* costs_device,
                                    int prepare_backward, void* workspace, rnntOptions options) {
    return run<double>(activations, nullptr, flat_labels, label_lengths, input_lengths, alphabet_size,
                       minibatch, costs_device, true, 1.0, nullptr, kForward, prepare_backward != 0,
                       workspace, options);
}

rnntStatus_t rnnt_b200_backward(const float* const activations, float* gradients,
                                const int* const flat_labels, const int* const label_lengths,
                                const int* const input_lengths, int alphabet_size, int minibatch,
                                const float* grad_costs_device, float grad_scale, void* workspace,
                                rnntOptions options) {
    return run<float>(activations, gradients, flat_labels, label_lengths, input_lengths, alphabet_size,
                      minibatch, nullptr, true, grad_scale, grad_costs_device, kBackward, false,
                      workspace, options);
}

rnntStatus_t rnnt_b200_backward_fp64(const double* const activations, double* gradients,
                                     const int* const flat_labels, const int* const label_lengths,
                                     const int* const input_lengths, int alphabet_size,
                                     int minibatch, const double* grad_costs_device,
                                     double grad_scale, void* workspace, rnntOptions options) {
    return run<double>(activations, gradients, flat_labels, label_lengths, input_lengths,
                       alphabet_size, minibatch, nullptr, true, grad_scale, grad_costs_device,
                       kBackward, false, workspace, options);
}

// ---- explicit activation layout ([N,T,U,V] or [T,U,N,V]) ------------------------------------------
rnntStatus_t rnnt_b200_loss_async_layout(int layout, const float* activations, float* gradients,
                                         const int* flat_labels, const int* label_lengths,
                                         const int* input_lengths, int alphabet_size, int minibatch,
                                         float* costs_device, float grad_scale, void* workspace,
                                         rnntOptions options) {
    if (layout != RNNT_B200_LAYOUT_NTUV && layout != RNNT_B200_LAYOUT_TUNV) return RNNT_STATUS_INVALID_VALUE;
    g_layout_tunv = layout == RNNT_B200_LAYOUT_TUNV;
    const rnntStatus_t st = run<float>(activations, gradients, flat_labels, label_lengths, input_lengths,
                                       alphabet_size, minibatch, costs_device, true, grad_scale, nullptr, kFull,
                                       false, workspace, options);
    g_layout_tunv = false;
    return st;
}

rnntStatus_t rnnt_b200_loss_async_layout_fp64(int layout, const double* activations, double* gradients,
                                              const int* flat_labels, const int* label_lengths,
                                              const int* input_lengths, int alphabet_size, int minibatch,
                                              double* costs_device, double grad_scale, void* workspace,
                                              rnntOptions options) {
    if (layout != RNNT_B200_LAYOUT_NTUV && layout != RNNT_B200_LAYOUT_TUNV) return RNNT_STATUS_INVALID_VALUE;
    g_layout_tunv = layout == RNNT_B200_LAYOUT_TUNV;
    const rnntStatus_t st = run<double>(activations, gradients, flat_labels, label_lengths, input_lengths,
                                        alphabet_size, minibatch, costs_device, true, grad_scale, nullptr, kFull,
                                        false, workspace, options);
    g_layout_tunv = false;
    return st;
}

// ---- 16-bit storage (bf16 / fp16 logits and gradients, fp32 arithmetic and costs) ---------------
rnntStatus_t rnnt_b200_loss_async_16(int dtype, const void* activations, void* gradients,
                                     const int* flat_labels, const int* label_lengths,
                                     const int* input_lengths, int alphabet_size, int minibatch,
                                     float* costs_device, float grad_scale, void* workspace,
                                     rnntOptions options) {
    if (dtype == RNNT_B200_BF16)
        return run<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(activations),
                                  static_cast<__nv_bfloat16*>(gradients), flat_labels, label_lengths,
                                  input_lengths, alphabet_size, minibatch, costs_device, true, grad_scale,
                                  nullptr, kFull, false, workspace, options);
    if (dtype == RNNT_B200_FP16)
        return run<__half>(static_cast<const __half*>(activations), static_cast<__half*>(gradients),
                           flat_labels, label_lengths, input_lengths, alphabet_size, minibatch,
                           costs_device, true, grad_scale, nullptr, kFull, false, workspace, options);
    return RNNT_STATUS_INVALID_VALUE;
}

rnntStatus_t rnnt_b200_forward_16(int dtype, const void* activations, const int* flat_labels,
                                  const int* label_lengths, const int* input_lengths,
                                  int alphabet_size, int minibatch, float* costs_device,
                                  int prepare_backward, void* workspace, rnntOptions options) {
    if (dtype == RNNT_B200_BF16)
        return run<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(activations), nullptr, flat_labels,
                                  label_lengths, input_lengths, alphabet_size, minibatch, costs_device,
                                  true, 1.0f, nullptr, kForward, prepare_backward != 0, workspace, options);
    if (dtype == RNNT_B200_FP16)
        return run<__half>(static_cast<const __half*>(activations), nullptr, flat_labels, label_lengths,
                           input_lengths, alphabet_size, minibatch, costs_device, true, 1.0f, nullptr,
                           kForward, prepare_backward != 0, workspace, options);
    return RNNT_STATUS_INVALID_VALUE;
}

rnntStatus_t rnnt_b200_backward_16(int dtype, const void* activations, void* gradients,
                                   const int* flat_labels, const int* label_lengths,
                                   const int* input_lengths, int alphabet_size, int minibatch,
                                   const float* grad_costs_device, float grad_scale, void* workspace,
                                   rnntOptions options) {
    if (dtype == RNNT_B200_BF16)
        return run<__nv_bfloat16>(static_cast<const __nv_bfloat16*>(activations),
                                  static_cast<__nv_bfloat16*>(gradients), flat_labels, label_lengths,
                                  input_lengths, alphabet_size, minibatch, nullptr, true, grad_scale,
                                  grad_costs_device, kBackward, false, workspace, options);
    if (dtype == RNNT_B200_FP16)
        return run<__half>(static_cast<const __half*>(activations), static_cast<__half*>(gradients),
                           flat_labels, label_lengths, input_lengths, alphabet_size, minibatch, nullptr,
                           true, grad_scale, grad_costs_device, kBackward, false, workspace, options);
    return RNNT_STATUS_INVALID_VALUE;
}

// ---- additive joint network, logits never materialised -------------------------------------------
rnntStatus_t rnnt_b200_add_joint_loss(const float* trans, const float* pred, float* grad_trans,
                                      float* grad_pred, const int* flat_labels,
                                      const int* label_lengths, const int* input_lengths,
                                      int alphabet_size, int minibatch, float* costs_device,
                                      float grad_scale, void* workspace, rnntOptions options) {
    return run_add_joint(trans, pred, grad_trans, grad_pred, flat_labels, label_lengths, input_lengths,
                         alphabet_size, minibatch, costs_device, grad_scale, nullptr, kFull, false, workspace,
                         options);
}

rnntStatus_t rnnt_b200_add_joint_forward(const float* trans, const float* pred, const int* flat_labels,
                                         const int* label_lengths, const int* input_lengths,
                                         int alphabet_size, int minibatch, float* costs_device,
                                         int prepare_backward, void* workspace, rnntOptions options) {
    return run_add_joint(trans, pred, nullptr, nullptr, flat_labels, label_lengths, input_lengths,
                         alphabet_size, minibatch, costs_device, 1.0f, nullptr, kForward,
                         prepare_backward != 0, workspace, options);
}

rnntStatus_t rnnt_b200_add_joint_backward(const float* trans, const float* pred, float* grad_trans,
                                          float* grad_pred, const int* flat_labels,
                                          const int* label_lengths, const int* input_lengths,
                                          int alphabet_size, int minibatch, const float* grad_costs_device,
                                          float grad_scale, void* workspace, rnntOptions options) {
    return run_add_joint(trans, pred, grad_trans, grad_pred, flat_labels, label_lengths, input_lengths,
                         alphabet_size, minibatch, nullptr, grad_scale, grad_costs_device, kBackward, false,
                         workspace, options);
}

rnntStatus_t rnnt_b200_add_joint_workspace_size(int maxT, int maxU, int minibatch, int alphabet_size,
                                                size_t* size_bytes) {
    if (minibatch <= 0 || maxT <= 0 || maxU <= 0 || alphabet_size <= 0 || size_bytes == nullptr)
        return RNNT_STATUS_INVALID_VALUE;
    *size_bytes = carve_joint(nullptr, minibatch, maxT, maxU, alphabet_size).bytes;
    return RNNT_STATUS_SUCCESS;
}

rnntStatus_t get_workspace_size(int maxT, int maxU, int minibatch, bool gpu, size_t* size_bytes,
                                size_t dtype_size) {
    if (minibatch <= 0 || maxT <= 0 || maxU <= 0 || size_bytes == nullptr)
        return RNNT_STATUS_INVALID_VALUE;  // reference src/rnnt_entrypoint.cpp:102-105
    if (dtype_size != sizeof(double)) dtype_size = sizeof(float);
    const size_t rows = (size_t)minibatch * maxT * maxU;
    if (!gpu) {
        // No CPU path here; the reference's figure (alphas, betas, 2-wide log-prob cache) is
        // returned so host-side sizing code keeps working.  src/rnnt_entrypoint.cpp:113-118
        *size_bytes = dtype_size * rows * 4;
        return RNNT_STATUS_SUCCESS;
    }
    const size_t lat = (size_t)minibatch * (maxT + maxU - 1) * maxU;
    *size_bytes = carve(nullptr, rows, lat, minibatch, maxU, dtype_size).bytes;
    return RNNT_STATUS_SUCCESS;
}

rnntStatus_t get_rnnt_workspace_size(int maxT, int maxU, int minibatch, bool gpu,
                                     size_t* size_bytes, size_t dtype_size) {
    return get_workspace_size(maxT, maxU, minibatch, gpu, size_bytes, dtype_size);
}

int rnnt_b200_last_launch_count(void) { return g_last_launches; }

// Debug / test hook: forward and backward log-likelihoods (natural log) left in a workspace by the last
// call that produced both lattices.  The reference asserts their agreement in debug builds
// (include/detail/cpu_rnnt.h:167-170).  Synchronises the device.
rnntStatus_t rnnt_b200_debug_log_likelihoods(const void* workspace, int maxT, int maxU, int minibatch,
                                             size_t dtype_size, double* llf_host, double* llb_host) {
    if (!workspace || !llf_host || !llb_host || maxT <= 0 || maxU <= 0 || minibatch <= 0)
        return RNNT_STATUS_INVALID_VALUE;
    if (dtype_size != sizeof(double)) dtype_size = sizeof(float);
    const size_t rows = (size_t)minibatch * maxT * maxU;
    const size_t lat = (size_t)minibatch * (maxT + maxU - 1) * maxU;
    const Workspace w = carve(const_cast<void*>(workspace), rows, lat, minibatch, maxU, dtype_size);
    std::vector<unsigned char> f((size_t)minibatch * 8), b((size_t)minibatch * 8);
    if (cudaDeviceSynchronize() != cudaSuccess) return RNNT_STATUS_EXECUTION_FAILED;
    if (cudaMemcpy(f.data(), w.llf, f.size(), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(b.data(), w.llb, b.size(), cudaMemcpyDeviceToHost) != cudaSuccess)
        return RNNT_STATUS_MEMOPS_FAILED;
    for (int i = 0; i < minibatch; ++i) {
        if (dtype_size == sizeof(double)) {
            llf_host[i] = reinterpret_cast<const double*>(f.data())[i];
            llb_host[i] = reinterpret_cast<const double*>(b.data())[i];
        } else {
            const LogVal lf = reinterpret_cast<const LogVal*>(f.data())[i];
            const LogVal lb = reinterpret_cast<const LogVal*>(b.data())[i];
            llf_host[i] = ((double)lf.e + (double)lf.l) * 0.6931471805599453;
            llb_host[i] = ((double)lb.e + (double)lb.l) * 0.6931471805599453;
        }
    }
    return RNNT_STATUS_SUCCESS;
}

void rnnt_b200_set_profiling(int enabled) { g_profile = enabled != 0; }

int rnnt_b200_last_kernel_ms(float* ms3) {
    if (!ms3 || g_used == 0) return 0;
    return read_set(g_sets[g_used - 1], ms3);
}

int rnnt_b200_profile_collect(float* ms3_mean) {
    if (!ms3_mean) return 0;
    double acc[3] = {0, 0, 0};
    int cnt[3] = {0, 0, 0};
    for (size_t k = 0; k < g_used; ++k) {
        float ms[3];
        read_set(g_sets[k], ms);
        for (int i = 0; i < 3; ++i)
            if (ms[i] >= 0) {
                acc[i] += ms[i];
                ++cnt[i];
            }
    }
    for (int i = 0; i < 3; ++i) ms3_mean[i] = cnt[i] ? (float)(acc[i] / cnt[i]) : -1.0f;
    const int calls = (int)g_used;
    g_used = 0;
    return calls;
}

int rnnt_b200_debug_policy(int what, int a, int b) {
    switch (what) {
        case 0: return (size_t)a * (size_t)b > 512 || a <= 0 ? 0 : pick_tpr(a);
        case 1: {
            if ((size_t)a * (size_t)b > 512 || a <= 0) return 0;
            const int tpr = pick_tpr(a);
            return chunk_walk_cost(a, tpr, b, true) < chunk_walk_cost(a, tpr, b, false) ? 1 : 0;
        }
        case 2: return lattice_cols(a);
        case 3: return lattice_threads(a);
        case 4: return lattice_ring_depth(a, b != 0);
        case 5: return joint_slices(a);
        default: return -1;
    }
}

const char* rnnt_b200_build_info(void) { return "b200-rnnt sm_100a built " __DATE__ " " __TIME__; }

}  // extern "C"
