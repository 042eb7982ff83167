// gfs_p2p.cu — K5b: replica reconcile as ONE kernel over NVLink peer memory.
//
// The exchange step of a replicated multi-GPU run.  Every rank maps every other rank's replica (CUDA IPC between
// one-process-per-GPU ranks, plain peer access inside one process) and one kernel per rank does the whole exchange:
//
//   start barrier   block b of rank r signals block b of every peer: "my SGD slice is complete" (it is: the kernel
//                   is stream-ordered after the SGD kernel) and waits for theirs;
//   reduce+scatter  rank r owns the 16-byte vectors [nvec r/G, nvec (r+1)/G): for each it loads the G replicas' values
//                   (G-1 of them over NVLink: one 16-byte load per peer, all G issued before the first use), forms
//                   x_sync + sum of displacements / #replicas that moved the element (the "moved-replica mean" of
//                   DESIGN.md §6, in f64) and stores the result into all G replicas (16-byte stores) and into its OWN
//                   x_sync — the base of a slice is only ever read by the slice's owner, so x_sync is kept per slice
//                   by its owner and there is no refresh pass over the whole replica (until round 2's last change
//                   there was one: 160 MB of local traffic per reconcile);
//   end barrier     all of this rank's peer stores are visible (fence.sys + release) before any peer goes on.
//
// Per rank and reconcile: n(G-1)/G elements read and n(G-1)/G written over NVLink — the volume of a ring all-reduce —
// in one launch, no staging buffer, local traffic 3n/G reads + 2n/G writes.
//
// Barrier flags live in the region itself, one u32 per (phase, block, source rank), written by the source with
// st.release.sys and polled by the owner with ld.acquire.sys; tags increase by one per reconcile, so no reset is
// needed.  Every spin is bounded.  A barrier that does not complete within `spin_cap` polls FAILS THE RUN: the block
// raises the error word of every rank's region (its own and, over NVLink, its peers'), every later reconcile kernel
// returns at once when it sees the word set, and gfs_p2p_region_check — called by the replica runner after every
// schedule it enqueues and before any result is handed out — turns it into GFS_ERR_CUDA.  After a timeout the
// replicas are in an undefined state (other blocks may already have scattered); the contract is "an error, never a
// hang and never a silently wrong result", not "untouched replicas".
//
// OVERLAPPED form (rc_p2p_async; gfs_p2p_reconcile_async): the same rule without stopping the SGD.  After an SGD slice the
// rank copies its replica into a snapshot x_snap (local, 30 us) and starts the next slice at once; on a second stream a
// small kernel (128 threads per block, <= 32 registers: one warp of 1024 registers per SM sub-partition, which is exactly
// what the three resident SGD blocks leave free there — 64 threads x 64 registers is the same total and does NOT fit) meets
// its peers at the same barriers, forms the new common base B' = x_sync + moved-replica mean of (x_snap_g - x_sync) for
// its slice, and for EVERY rank g — its own included — adds (B' - x_snap_g) to the live replica with red.add (over NVLink
// for the peers) and stores B' into its own x_sync.  The live replica is then B' + whatever the rank has done since
// its snapshot: peers' contributions arrive a fraction of an epoch late instead of stopping everybody for the exchange.
// "Moved" is |x_snap - x_sync| above 2.5 ulps, because x + (B' - x) is B' only up to one rounding.  The next snapshot waits
// (stream event) until this kernel — whose end barrier says every peer's corrections have landed — has finished.
//
// One device, several replicas (tests): kernels that wait on one another must not be separate launches on one GPU
// (nothing guarantees they run at the same time), so gfs_p2p_reconcile_local runs all ranks as ONE cooperative
// launch in which block group g plays rank g.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gfasort_cuda.h"

namespace gfs { void set_error(const std::string& s); }

namespace {

constexpr uint32_t P2P_MAX_RANKS = GFS_P2P_MAX_RANKS;
constexpr uint32_t P2P_MAX_BLOCKS = 256;
constexpr uint32_t P2P_THREADS = 512;
constexpr uint64_t P2P_ALIGN = 256;
constexpr uint32_t P2P_EMU_MAX = 8;         // ranks one emulated launch can hold (kernel parameter space: 8 x 360 B)

struct P2pArgs {
    void* x[P2P_MAX_RANKS];                  // replicas, by rank (own entry = local pointer)
    uint32_t* flags[P2P_MAX_RANKS];          // [phase 0/1][block][source rank]
    unsigned long long* err[P2P_MAX_RANKS];  // every rank's error word
    void* snap[P2P_MAX_RANKS];               // overlapped form: every rank's snapshot x_snap
    void* xs;                                // local x_sync
    double inv_div[P2P_MAX_RANKS + 1];       // div_pow != 1 (GFASORT_RC_POW): 1 / m^div_pow for m replicas that moved an element
    uint32_t use_inv_div;                    // 0: divide by the number of replicas that moved it (the moved-replica mean)
    uint64_t nvec;                           // 16-byte vectors per replica (arrays are padded to 256 B)
    uint64_t spin_cap;
    uint32_t rank, world, tag, blocks;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// one 16-byte vector of T
template <typename T> struct V16;
template <> struct V16<double> {
    static constexpr int N = 2;
    double v[2];
    static __device__ __forceinline__ V16 ld_sys(const void* p) {       // peer (or local) replica: bypass L1
        V16 r;
        asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.v[0]), "=d"(r.v[1]) : "l"(p) : "memory");
        return r;
    }
    static __device__ __forceinline__ void st(void* p, const V16& r) {
        asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]) : "memory");
    }
};
template <> struct V16<float> {
    static constexpr int N = 4;
    float v[4];
    static __device__ __forceinline__ V16 ld_sys(const void* p) {
        V16 r;
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p) : "memory");
        return r;
    }
    static __device__ __forceinline__ void st(void* p, const V16& r) {
        asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]) : "memory");
    }
};

// vectors [slice_begin(r), slice_begin(r + 1)) belong to rank r: n/G each, the first n%G ranks one more
__device__ __host__ __forceinline__ uint64_t slice_begin(uint64_t n, uint32_t world, uint32_t r) {
    const uint64_t q = n / world, rem = n % world;
    return q * r + (r < rem ? r : rem);
}

// Block b of this rank meets block b of every peer.  PHASE 1 (end) publishes this block's peer stores first.
template <int PHASE>
__device__ bool p2p_barrier(const P2pArgs& a, uint32_t b) {
    if (PHASE == 1) __threadfence_system();
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < a.world) {
        const uint32_t peer = threadIdx.x;
        const size_t slot = ((size_t)PHASE * P2P_MAX_BLOCKS + b) * P2P_MAX_RANKS;
        st_release_sys(a.flags[peer] + slot + a.rank, a.tag);
        const uint32_t* mine = a.flags[a.rank] + slot + peer;
        uint64_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - a.tag) < 0) {
            if (++spins > a.spin_cap) { ok = false; break; }
            __nanosleep(64);
        }
    }
    return __syncthreads_and(ok) != 0;
}

// a barrier timed out: fail the run on every rank (see the header)
__device__ void p2p_raise(const P2pArgs& a) {
    if (threadIdx.x < a.world) atomicAdd_system(a.err[threadIdx.x], 1ull);
}

// the moved-replica mean of one 16-byte vector; W > 0: world known at compile time (all loads issued up front)
template <typename T, int W>
__device__ __forceinline__ void reduce_scatter_vec(const P2pArgs& a, uint64_t i) {
    using V = V16<T>;
    const V s = *reinterpret_cast<const V*>(static_cast<const char*>(a.xs) + i * 16);   // the base, kept by the slice's owner
    V nv;
    if constexpr (W > 0) {
        V v[W];
#pragma unroll
        for (int g = 0; g < W; ++g) v[g] = V::ld_sys(static_cast<const char*>(a.x[g]) + i * 16);
#pragma unroll
        for (int k = 0; k < V::N; ++k) {
            double sum = 0.0; uint32_t moved = 0; T only = s.v[k];
#pragma unroll
            for (int g = 0; g < W; ++g)
                if (v[g].v[k] != s.v[k]) { sum += (double)v[g].v[k] - (double)s.v[k]; ++moved; only = v[g].v[k]; }
            nv.v[k] = moved <= 1 ? only                                                 // exact when 0 or 1 replicas moved it
                      : (T)((double)s.v[k] + (a.use_inv_div ? sum * a.inv_div[moved] : sum / (double)moved));
        }
#pragma unroll
        for (int g = 0; g < W; ++g) V::st(static_cast<char*>(a.x[g]) + i * 16, nv);
    } else {
        double sum[V::N]; uint32_t moved[V::N]; T only[V::N];
#pragma unroll
        for (int k = 0; k < V::N; ++k) { sum[k] = 0.0; moved[k] = 0; only[k] = s.v[k]; }
        for (uint32_t g = 0; g < a.world; ++g) {
            const V v = V::ld_sys(static_cast<const char*>(a.x[g]) + i * 16);
#pragma unroll
            for (int k = 0; k < V::N; ++k)
                if (v.v[k] != s.v[k]) { sum[k] += (double)v.v[k] - (double)s.v[k]; ++moved[k]; only[k] = v.v[k]; }
        }
#pragma unroll
        for (int k = 0; k < V::N; ++k)
            nv.v[k] = moved[k] <= 1 ? only[k]
                      : (T)((double)s.v[k] + (a.use_inv_div ? sum[k] * a.inv_div[moved[k]] : sum[k] / (double)moved[k]));
        for (uint32_t g = 0; g < a.world; ++g) V::st(static_cast<char*>(a.x[g]) + i * 16, nv);
    }
    *reinterpret_cast<V*>(static_cast<char*>(a.xs) + i * 16) = nv;       // the new base of this slice (only its owner ever reads it)
}

template <typename T>
__device__ void rc_p2p_body(const P2pArgs& a, uint32_t b) {
    using V = V16<T>;
    __shared__ int s_dead;
    if (threadIdx.x == 0) s_dead = ld_relaxed_sys_u64(a.err[a.rank]) != 0;     // an earlier reconcile failed: the run is over
    __syncthreads();
    if (s_dead) return;
    if (!p2p_barrier<0>(a, b)) { p2p_raise(a); return; }
    const uint64_t lo = slice_begin(a.nvec, a.world, a.rank), hi = slice_begin(a.nvec, a.world, a.rank + 1);
    const uint64_t stride = (uint64_t)a.blocks * blockDim.x;
    const uint64_t first = (uint64_t)b * blockDim.x + threadIdx.x;
    switch (a.world) {
        case 2: for (uint64_t i = lo + first; i < hi; i += stride) reduce_scatter_vec<T, 2>(a, i); break;
        case 4: for (uint64_t i = lo + first; i < hi; i += stride) reduce_scatter_vec<T, 4>(a, i); break;
        case 8: for (uint64_t i = lo + first; i < hi; i += stride) reduce_scatter_vec<T, 8>(a, i); break;
        default: for (uint64_t i = lo + first; i < hi; i += stride) reduce_scatter_vec<T, 0>(a, i); break;
    }
    // every peer's stores into MY replica have landed when this barrier opens.  There is no refresh pass: x_sync is only
    // ever read by the owner of a slice, for that slice, and the owner stored the new base itself (reduce_scatter_vec)
    if (!p2p_barrier<1>(a, b)) { p2p_raise(a); return; }
}

template <typename T>
__global__ void __launch_bounds__(P2P_THREADS) rc_p2p(const P2pArgs a) { rc_p2p_body<T>(a, blockIdx.x); }

// all ranks in one cooperative launch (replicas on one device): block group g plays rank g
struct P2pEmuArgs { P2pArgs r[P2P_EMU_MAX]; };
template <typename T>
__global__ void __launch_bounds__(P2P_THREADS) rc_p2p_emulated(const P2pEmuArgs e) {
    const uint32_t blocks = e.r[0].blocks;
    rc_p2p_body<T>(e.r[blockIdx.x / blocks], blockIdx.x % blocks);
}

// ---- overlapped form -------------------------------------------------------------------------------------------
constexpr uint32_t P2P_ASYNC_THREADS = 128;
template <typename T> __device__ __forceinline__ bool moved_beyond_rounding(T v, T s);
template <> __device__ __forceinline__ bool moved_beyond_rounding<double>(double v, double s) {
    return fabs(v - s) > fabs(s) * 5.6e-16 + 1e-300;             // 2.5 ulps: x + (B' - x) returns to B' only up to one rounding
}
template <> __device__ __forceinline__ bool moved_beyond_rounding<float>(float v, float s) {
    return fabsf(v - s) > fabsf(s) * 3e-7f + 1e-37f;
}
template <typename T> __device__ __forceinline__ void red_add(T* p, T v);
template <> __device__ __forceinline__ void red_add<double>(double* p, double v) {
    asm volatile("red.relaxed.sys.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
template <> __device__ __forceinline__ void red_add<float>(float* p, float v) {
    asm volatile("red.relaxed.sys.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// G > 0: the world size is a compile-time constant — all G snapshots of a vector are loaded at once (G loads in flight
// per thread over NVLink) and stay in registers for the second half, so every snapshot crosses the link ONCE.
// G == 0: any world size, two sequential passes over the snapshots.
template <typename T, int G>
__device__ void rc_p2p_async_body(const P2pArgs& a, uint32_t b) {
    using V = V16<T>;
    __shared__ int s_dead;
    if (threadIdx.x == 0) s_dead = ld_relaxed_sys_u64(a.err[a.rank]) != 0;
    __syncthreads();
    if (s_dead) return;
    if (!p2p_barrier<0>(a, b)) { p2p_raise(a); return; }          // every rank's snapshot of this round is complete
    const uint64_t lo = slice_begin(a.nvec, a.world, a.rank), hi = slice_begin(a.nvec, a.world, a.rank + 1);
    const uint64_t stride = (uint64_t)a.blocks * blockDim.x;
    for (uint64_t i = lo + (uint64_t)b * blockDim.x + threadIdx.x; i < hi; i += stride) {
        const V s = *reinterpret_cast<const V*>(static_cast<const char*>(a.xs) + i * 16);
        double sum[V::N]; uint32_t moved[V::N]; T only[V::N];
#pragma unroll
        for (int k = 0; k < V::N; ++k) { sum[k] = 0.0; moved[k] = 0; only[k] = s.v[k]; }
        if constexpr (G > 0) {
            V v[G];
#pragma unroll
            for (int g = 0; g < G; ++g) v[g] = V::ld_sys(static_cast<const char*>(a.snap[g]) + i * 16);
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int k = 0; k < V::N; ++k)
                    if (moved_beyond_rounding<T>(v[g].v[k], s.v[k])) { sum[k] += (double)v[g].v[k] - (double)s.v[k]; ++moved[k]; only[k] = v[g].v[k]; }
            V nv;
#pragma unroll
            for (int k = 0; k < V::N; ++k) nv.v[k] = moved[k] <= 1 ? only[k] : (T)((double)s.v[k] + sum[k] / (double)moved[k]);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                T* xg = reinterpret_cast<T*>(static_cast<char*>(a.x[g]) + i * 16);
#pragma unroll
                for (int k = 0; k < V::N; ++k) {
                    const T d = (T)((double)nv.v[k] - (double)v[g].v[k]);
                    if (d != T(0)) red_add<T>(xg + k, d);         // the rank keeps what it did since its snapshot
                }
            }
            *reinterpret_cast<V*>(static_cast<char*>(a.xs) + i * 16) = nv;
        } else {
            for (uint32_t g = 0; g < a.world; ++g) {
                const V v = V::ld_sys(static_cast<const char*>(a.snap[g]) + i * 16);
#pragma unroll
                for (int k = 0; k < V::N; ++k)
                    if (moved_beyond_rounding<T>(v.v[k], s.v[k])) { sum[k] += (double)v.v[k] - (double)s.v[k]; ++moved[k]; only[k] = v.v[k]; }
            }
            V nv;
#pragma unroll
            for (int k = 0; k < V::N; ++k) nv.v[k] = moved[k] <= 1 ? only[k] : (T)((double)s.v[k] + sum[k] / (double)moved[k]);
            for (uint32_t g = 0; g < a.world; ++g) {
                const V v = V::ld_sys(static_cast<const char*>(a.snap[g]) + i * 16);
                T* xg = reinterpret_cast<T*>(static_cast<char*>(a.x[g]) + i * 16);
#pragma unroll
                for (int k = 0; k < V::N; ++k) {
                    const T d = (T)((double)nv.v[k] - (double)v.v[k]);
                    if (d != T(0)) red_add<T>(xg + k, d);
                }
            }
            *reinterpret_cast<V*>(static_cast<char*>(a.xs) + i * 16) = nv;
        }
    }
    if (!p2p_barrier<1>(a, b)) { p2p_raise(a); return; }          // every peer's corrections to MY replica have landed
}
// 4 warps x 32 x 32 registers: each SM sub-partition has 1024 registers left beside three resident SGD blocks (24 warps x
// 80 registers = 15360 of its 16384).  Measured: a 64-thread x 64-register block never became resident next to the SGD.
template <typename T, int G>
__global__ void __launch_bounds__(P2P_ASYNC_THREADS, 16) rc_p2p_async(const P2pArgs a) { rc_p2p_async_body<T, G>(a, blockIdx.x); }
template <typename T, int G>
__global__ void __launch_bounds__(P2P_ASYNC_THREADS, 16) rc_p2p_async_emulated(const P2pEmuArgs e) {
    const uint32_t blocks = e.r[0].blocks;
    rc_p2p_async_body<T, G>(e.r[blockIdx.x / blocks], blockIdx.x % blocks);
}
template <typename T> const void* async_kernel(uint32_t world, bool emulated) {
    switch (world) {
        case 2: return emulated ? (const void*)rc_p2p_async_emulated<T, 2> : (const void*)rc_p2p_async<T, 2>;
        case 4: return emulated ? (const void*)rc_p2p_async_emulated<T, 4> : (const void*)rc_p2p_async<T, 4>;
        case 8: return emulated ? (const void*)rc_p2p_async_emulated<T, 8> : (const void*)rc_p2p_async<T, 8>;
        default: return emulated ? (const void*)rc_p2p_async_emulated<T, 0> : (const void*)rc_p2p_async<T, 0>;
    }
}

uint64_t align_up(uint64_t v) { return (v + P2P_ALIGN - 1) / P2P_ALIGN * P2P_ALIGN; }

}  // namespace

struct gfs_p2p_region {
    int device = 0;
    uint64_t n = 0;
    uint32_t elem_bytes = 8;
    uint32_t rank = 0, world = 1, blocks = 1, tag = 0;
    size_t off_xs = 0, off_snap = 0, off_flags = 0, off_err = 0, bytes = 0;
    char* base = nullptr;                               // this rank's allocation
    char* peer_base[P2P_MAX_RANKS] = {};                // every rank's allocation as mapped here
    bool ipc_opened[P2P_MAX_RANKS] = {};
    bool connected = false;
    uint64_t spin_cap = 0;
    double div_pow = 1.0;                               // GFASORT_RC_POW: divisor = (#replicas that moved the element)^div_pow
};

#define P2P_CUDA(call)                                                                                       \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            gfs::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__));                      \
            return GFS_ERR_CUDA;                                                                             \
        }                                                                                                    \
    } while (0)

extern "C" int gfs_p2p_region_create(int32_t device, uint64_t n, uint32_t elem_bytes, uint32_t max_blocks, gfs_p2p_region** out) {
    if (!out) { gfs::set_error("gfs_p2p_region_create: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    if (elem_bytes != 4 && elem_bytes != 8) { gfs::set_error("gfs_p2p_region_create: elem_bytes must be 4 or 8"); return GFS_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        gfs::set_error("no CUDA device available — libgfasort_cuda has no CPU fallback");
        return GFS_ERR_NO_DEVICE;
    }
    if (device < 0) P2P_CUDA(cudaGetDevice(&device));
    if (device >= count) { gfs::set_error("gfs_p2p_region_create: device ordinal out of range"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(device));
    gfs_p2p_region* r = new gfs_p2p_region();
    r->device = device; r->n = n; r->elem_bytes = elem_bytes;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    uint32_t blocks = (uint32_t)(sms > 0 ? sms : 1);                       // one resident block per SM: the barriers need co-residency
    if (max_blocks) blocks = std::min(blocks, max_blocks);
    r->blocks = std::max(1u, std::min(blocks, P2P_MAX_BLOCKS));
    const char* cap = std::getenv("GFASORT_P2P_SPIN_CAP");
    r->spin_cap = cap && *cap ? std::strtoull(cap, nullptr, 10) : (1ull << 24);   // x >= 64 ns: seconds, not forever
    // experiment (profiles/r2_experiments.md §6): 1 = the moved-replica mean (default), 0 = the sum of the displacements;
    // the stop-the-world kernel only
    const char* dp = std::getenv("GFASORT_RC_POW");
    if (dp && *dp) r->div_pow = std::min(1.0, std::max(0.0, std::strtod(dp, nullptr)));
    const uint64_t arr = align_up(std::max<uint64_t>(n, 1) * elem_bytes);  // padded: the kernel works on whole 16-byte vectors
    r->off_xs = arr;
    r->off_snap = 2 * arr;
    r->off_flags = 3 * arr;
    r->off_err = r->off_flags + align_up((uint64_t)2 * P2P_MAX_BLOCKS * P2P_MAX_RANKS * 4);
    r->bytes = r->off_err + P2P_ALIGN;
    cudaError_t e = cudaMalloc(&r->base, r->bytes);
    if (e != cudaSuccess) { gfs::set_error(std::string("gfs_p2p_region_create: cudaMalloc failed: ") + cudaGetErrorString(e)); delete r; return GFS_ERR_CUDA; }
    e = cudaMemset(r->base, 0, r->bytes);
    if (e != cudaSuccess) { gfs::set_error(std::string("gfs_p2p_region_create: cudaMemset failed: ") + cudaGetErrorString(e)); cudaFree(r->base); delete r; return GFS_ERR_CUDA; }
    r->peer_base[0] = r->base;
    *out = r;
    return GFS_OK;
}

extern "C" int gfs_p2p_region_snap_ptr(gfs_p2p_region* r, void** x_snap) {
    if (!r || !x_snap) { gfs::set_error("gfs_p2p_region_snap_ptr: null argument"); return GFS_ERR_INVALID; }
    *x_snap = r->base + r->off_snap;
    return GFS_OK;
}
extern "C" int gfs_p2p_region_ptrs(gfs_p2p_region* r, void** x, void** x_sync, uint64_t* region_bytes) {
    if (!r) { gfs::set_error("gfs_p2p_region_ptrs: null region"); return GFS_ERR_INVALID; }
    if (x) *x = r->base;
    if (x_sync) *x_sync = r->base + r->off_xs;
    if (region_bytes) *region_bytes = r->bytes;
    return GFS_OK;
}

// blob = cudaIpcMemHandle_t (64 bytes) + n (u64) + elem_bytes (u32) + blocks (u32): what a peer must agree on
extern "C" int gfs_p2p_region_ipc_handle(gfs_p2p_region* r, uint8_t* handle /*GFS_P2P_HANDLE_BYTES*/) {
    if (!r || !handle) { gfs::set_error("gfs_p2p_region_ipc_handle: null argument"); return GFS_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) + 16 == GFS_P2P_HANDLE_BYTES, "blob = 64-byte IPC handle + 16 bytes of shape");
    P2P_CUDA(cudaSetDevice(r->device));
    cudaIpcMemHandle_t h;
    P2P_CUDA(cudaIpcGetMemHandle(&h, r->base));
    std::memcpy(handle, &h, sizeof h);
    std::memcpy(handle + 64, &r->n, 8);
    std::memcpy(handle + 72, &r->elem_bytes, 4);
    std::memcpy(handle + 76, &r->blocks, 4);
    return GFS_OK;
}

extern "C" int gfs_p2p_region_connect_ipc(gfs_p2p_region* r, const uint8_t* handles /*world x GFS_P2P_HANDLE_BYTES, rank order*/,
                                          uint32_t world, uint32_t rank) {
    if (!r || !handles) { gfs::set_error("gfs_p2p_region_connect_ipc: null argument"); return GFS_ERR_INVALID; }
    if (world == 0 || world > P2P_MAX_RANKS || rank >= world) { gfs::set_error("gfs_p2p_region_connect_ipc: bad world / rank"); return GFS_ERR_INVALID; }
    if (r->connected) { gfs::set_error("gfs_p2p_region_connect_ipc: region already connected"); return GFS_ERR_INVALID; }
    // every peer must have the same element count, element type and grid: a mismatch would leave blocks without
    // a partner at the barriers (a guaranteed timeout) or reduce different elements
    for (uint32_t g = 0; g < world; ++g) {
        const uint8_t* b = handles + (size_t)g * GFS_P2P_HANDLE_BYTES;
        uint64_t n; uint32_t eb, bl;
        std::memcpy(&n, b + 64, 8); std::memcpy(&eb, b + 72, 4); std::memcpy(&bl, b + 76, 4);
        if (n != r->n || eb != r->elem_bytes || bl != r->blocks) {
            gfs::set_error("gfs_p2p_region_connect_ipc: rank " + std::to_string(g) + " differs in size, element type or grid (n " +
                           std::to_string(n) + " vs " + std::to_string(r->n) + ", elem_bytes " + std::to_string(eb) + " vs " +
                           std::to_string(r->elem_bytes) + ", blocks " + std::to_string(bl) + " vs " + std::to_string(r->blocks) + ")");
            return GFS_ERR_INVALID;
        }
    }
    P2P_CUDA(cudaSetDevice(r->device));
    r->rank = rank; r->world = world;
    for (uint32_t g = 0; g < world; ++g) {
        if (g == rank) { r->peer_base[g] = r->base; continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)g * GFS_P2P_HANDLE_BYTES, sizeof h);
        void* p = nullptr;
        P2P_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        r->peer_base[g] = static_cast<char*>(p);
        r->ipc_opened[g] = true;
    }
    r->connected = true;
    return GFS_OK;
}

extern "C" int gfs_p2p_region_connect_local(gfs_p2p_region* const* regions, uint32_t world) {
    if (!regions || world == 0 || world > P2P_MAX_RANKS) { gfs::set_error("gfs_p2p_region_connect_local: bad argument"); return GFS_ERR_INVALID; }
    for (uint32_t g = 0; g < world; ++g) {
        if (!regions[g] || regions[g]->connected) { gfs::set_error("gfs_p2p_region_connect_local: null or already connected region"); return GFS_ERR_INVALID; }
        if (regions[g]->n != regions[0]->n || regions[g]->elem_bytes != regions[0]->elem_bytes || regions[g]->blocks != regions[0]->blocks) {
            gfs::set_error("gfs_p2p_region_connect_local: regions differ in size, element type or grid");
            return GFS_ERR_INVALID;
        }
    }
    // enabling peer access costs milliseconds per ordered device pair (G (G-1) of them): one host thread per device
    std::vector<std::string> errs(world);
    std::vector<int> rcs(world, GFS_OK);
    {
        std::vector<std::thread> pool;
        for (uint32_t a = 0; a < world; ++a)
            pool.emplace_back([&, a] {
                auto fail = [&](const std::string& m, int code) { errs[a] = m; rcs[a] = code; };
                if (cudaSetDevice(regions[a]->device) != cudaSuccess) return fail("cudaSetDevice failed", GFS_ERR_CUDA);
                for (uint32_t b = 0; b < world; ++b) {
                    if (regions[b]->device == regions[a]->device) continue;
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, regions[a]->device, regions[b]->device) != cudaSuccess || !can)
                        return fail("gfs_p2p_region_connect_local: devices cannot access each other's memory", GFS_ERR_INVALID);
                    cudaError_t e = cudaDeviceEnablePeerAccess(regions[b]->device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                        return fail(std::string("cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(e), GFS_ERR_CUDA);
                    (void)cudaGetLastError();
                }
            });
        for (auto& th : pool) th.join();
    }
    for (uint32_t a = 0; a < world; ++a)
        if (rcs[a]) { gfs::set_error(errs[a]); return rcs[a]; }
    for (uint32_t a = 0; a < world; ++a) {
        for (uint32_t b = 0; b < world; ++b) regions[a]->peer_base[b] = regions[b]->base;
        regions[a]->rank = a; regions[a]->world = world; regions[a]->connected = true;
    }
    return GFS_OK;
}

// Asynchronous on `stream`: x_sync <- x.  Called once after the initial positions have been uploaded into x
// (every replica starts from the same positions, so the snapshots start equal — the invariant the kernel relies on).
extern "C" int gfs_p2p_region_snapshot(gfs_p2p_region* r, void* stream) {
    if (!r) { gfs::set_error("gfs_p2p_region_snapshot: null region"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    P2P_CUDA(cudaMemcpyAsync(r->base + r->off_xs, r->base, r->n * r->elem_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GFS_OK;
}

static P2pArgs make_args(gfs_p2p_region* r, uint32_t blocks) {
    P2pArgs a{};
    for (uint32_t g = 0; g < r->world; ++g) {
        a.x[g] = r->peer_base[g];
        a.snap[g] = r->peer_base[g] + r->off_snap;
        a.flags[g] = reinterpret_cast<uint32_t*>(r->peer_base[g] + r->off_flags);
        a.err[g] = reinterpret_cast<unsigned long long*>(r->peer_base[g] + r->off_err);
    }
    a.xs = r->base + r->off_xs;
    a.nvec = (r->n * r->elem_bytes + 15) / 16;
    a.use_inv_div = r->div_pow != 1.0;
    for (uint32_t m = 0; m <= P2P_MAX_RANKS; ++m) a.inv_div[m] = m ? 1.0 / std::pow((double)m, r->div_pow) : 1.0;
    a.spin_cap = r->spin_cap;
    a.rank = r->rank; a.world = r->world;
    a.tag = ++r->tag;
    a.blocks = blocks;
    return a;
}

// Asynchronous on `stream`.  Every rank calls it once per reconcile, in the same order.  Ranks must be on
// DIFFERENT devices (kernels of one device that wait on one another may never run together): replicas that
// share a device go through gfs_p2p_reconcile_local.
extern "C" int gfs_p2p_reconcile(gfs_p2p_region* r, void* stream) {
    if (!r) { gfs::set_error("gfs_p2p_reconcile: null region"); return GFS_ERR_INVALID; }
    if (!r->connected) { gfs::set_error("gfs_p2p_reconcile: region is not connected to its peers"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    const P2pArgs a = make_args(r, r->blocks);
    if (r->elem_bytes == 8) rc_p2p<double><<<r->blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(a);
    else rc_p2p<float><<<r->blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(a);
    P2P_CUDA(cudaGetLastError());
    return GFS_OK;
}

// Overlapped form, step 1 (asynchronous on `stream`, the stream the SGD runs on): x_snap <- x.
extern "C" int gfs_p2p_region_snapshot_x(gfs_p2p_region* r, void* stream) {
    if (!r) { gfs::set_error("gfs_p2p_region_snapshot_x: null region"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    P2P_CUDA(cudaMemcpyAsync(r->base + r->off_snap, r->base, r->n * r->elem_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return GFS_OK;
}
// Overlapped form, step 2 (asynchronous on `stream`, a SECOND stream that waits for the snapshot): the exchange over the
// snapshots; corrections are added to the live replicas while their owners go on sampling.
extern "C" int gfs_p2p_reconcile_async(gfs_p2p_region* r, void* stream) {
    if (!r) { gfs::set_error("gfs_p2p_reconcile_async: null region"); return GFS_ERR_INVALID; }
    if (!r->connected) { gfs::set_error("gfs_p2p_reconcile_async: region is not connected to its peers"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    P2pArgs a = make_args(r, r->blocks);
    const void* fn = r->elem_bytes == 8 ? async_kernel<double>(r->world, false) : async_kernel<float>(r->world, false);
    void* kargs[] = {(void*)&a};
    P2P_CUDA(cudaLaunchKernel(fn, dim3(r->blocks), dim3(P2P_ASYNC_THREADS), kargs, 0, (cudaStream_t)stream));
    return GFS_OK;
}

// All `world` replicas live on ONE device and were connected with gfs_p2p_region_connect_local: one cooperative
// launch in which block group g plays rank g (same barriers, same partition, same arithmetic as G ranks on G GPUs).
static int reconcile_local(gfs_p2p_region* const* regions, uint32_t world, void* stream, bool overlapped);
extern "C" int gfs_p2p_reconcile_local(gfs_p2p_region* const* regions, uint32_t world, void* stream) {
    return reconcile_local(regions, world, stream, false);
}
// the overlapped form's kernel for replicas that share one device (tests of its arithmetic): x_snap must have been taken
extern "C" int gfs_p2p_reconcile_async_local(gfs_p2p_region* const* regions, uint32_t world, void* stream) {
    return reconcile_local(regions, world, stream, true);
}
static int reconcile_local(gfs_p2p_region* const* regions, uint32_t world, void* stream, bool overlapped) {
    if (!regions || world == 0 || world > P2P_EMU_MAX) { gfs::set_error("gfs_p2p_reconcile_local: 1..8 regions"); return GFS_ERR_INVALID; }
    for (uint32_t g = 0; g < world; ++g) {
        if (!regions[g] || !regions[g]->connected || regions[g]->world != world || regions[g]->rank != g ||
            regions[g]->device != regions[0]->device) {
            gfs::set_error("gfs_p2p_reconcile_local: regions must be connected to each other, in rank order, on one device");
            return GFS_ERR_INVALID;
        }
    }
    P2P_CUDA(cudaSetDevice(regions[0]->device));
    const bool f64 = regions[0]->elem_bytes == 8;
    const void* fn = overlapped ? (f64 ? async_kernel<double>(world, true) : async_kernel<float>(world, true))
                                : (f64 ? (const void*)rc_p2p_emulated<double> : (const void*)rc_p2p_emulated<float>);
    const uint32_t threads = overlapped ? P2P_ASYNC_THREADS : P2P_THREADS;
    int per_sm = 0, sms = 0;
    P2P_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, 0));
    P2P_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, regions[0]->device));
    const uint32_t blocks = std::min<uint32_t>(regions[0]->blocks, (uint32_t)(per_sm * sms) / world);
    if (blocks == 0) { gfs::set_error("gfs_p2p_reconcile_local: the device cannot hold all ranks at once"); return GFS_ERR_INVALID; }
    P2pEmuArgs e{};
    for (uint32_t g = 0; g < world; ++g) e.r[g] = make_args(regions[g], blocks);
    void* kargs[] = {(void*)&e};
    P2P_CUDA(cudaLaunchCooperativeKernel(fn, dim3(blocks * world), dim3(threads), kargs, 0, (cudaStream_t)stream));
    return GFS_OK;
}

// Blocking: non-zero (GFS_ERR_CUDA) when a barrier of any earlier reconcile — on any rank — timed out.
extern "C" int gfs_p2p_region_check(gfs_p2p_region* r) {
    if (!r) { gfs::set_error("gfs_p2p_region_check: null region"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    unsigned long long err = 0;
    P2P_CUDA(cudaMemcpy(&err, r->base + r->off_err, 8, cudaMemcpyDeviceToHost));
    if (err) {
        gfs::set_error("peer-memory reconcile: a barrier timed out (a rank did not reach the reconcile, or its kernel was not "
                       "resident); the replicas are in an undefined state and the run has failed");
        return GFS_ERR_CUDA;
    }
    return GFS_OK;
}

extern "C" void gfs_p2p_region_free(gfs_p2p_region* r) {
    if (!r) return;
    cudaSetDevice(r->device);
    for (uint32_t g = 0; g < P2P_MAX_RANKS; ++g)
        if (r->ipc_opened[g]) cudaIpcCloseMemHandle(r->peer_base[g]);
    cudaFree(r->base);
    delete r;
}
