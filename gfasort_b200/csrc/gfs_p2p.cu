// gfs_p2p.cu — K5b: replica reconcile as ONE kernel over NVLink peer memory.
//
// The default reconcile of a multi-GPU run is rc_pack -> NCCL all-reduce -> rc_apply: three passes over the
// replica, a 2n-float staging buffer and two kernel boundaries around a library collective.  Here every rank
// maps every other rank's replica (CUDA IPC between the one-process-per-GPU ranks, or plain peer access inside
// one process) and one kernel per rank does the whole exchange:
//
//   start barrier   block b of rank r signals block b of every peer: "my SGD slice is complete" (it is: the kernel
//                   is stream-ordered after the SGD kernel) and waits for theirs;
//   reduce+scatter  rank r owns elements [n r/G, n (r+1)/G): for each it loads the G replicas' values (G-1 of them
//                   over NVLink, coalesced), forms x_sync + sum of displacements / #replicas that moved the element
//                   (the "moved-replica mean" of DESIGN.md §6, in f64 instead of a f32 staging buffer), and stores
//                   the result into all G replicas;
//   end barrier     all of this rank's peer stores are visible (fence.sys + release) before any peer goes on;
//   refresh         x_sync <- x over the whole local replica (local HBM traffic only).
//
// Per rank and reconcile: n(G-1)/G elements read and written over NVLink — the volume of a ring all-reduce — in one
// launch, with no staging buffer and 3 local passes (read x_sync slice, read x, write x_sync) instead of 7.
//
// Barrier flags live in the region itself, one u32 per (phase, block, source rank), written by the source with
// st.release.sys and polled by the owner with ld.acquire.sys; tags increase by one per reconcile, so no reset is
// needed.  Every spin is bounded: a barrier that does not complete within `spin_cap` polls raises the region's
// error flag and the kernel returns without touching any replica (gfs_p2p_region_check turns it into an error).
//
// STATUS: opt-in (`--reconcile p2p` / ReplicaRun(mode="p2p")); the NCCL path stays the default until this one
// has been measured on 2 and 8 GPUs.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gfasort_cuda.h"

namespace gfs { void set_error(const std::string& s); }

namespace {

constexpr uint32_t P2P_MAX_RANKS = GFS_P2P_MAX_RANKS;
constexpr uint32_t P2P_MAX_BLOCKS = 256;
constexpr uint32_t P2P_THREADS = 512;
constexpr uint64_t P2P_ALIGN = 256;

struct P2pArgs {
    void* x[P2P_MAX_RANKS];            // replicas, by rank (own entry = local pointer)
    uint32_t* flags[P2P_MAX_RANKS];    // [phase 0/1][block][source rank]
    void* xs;                          // local x_sync
    unsigned long long* err;           // local watchdog counter
    uint64_t n;
    uint64_t spin_cap;
    uint32_t rank, world, tag;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
template <typename T> __device__ __forceinline__ T ld_peer(const T* p);
template <> __device__ __forceinline__ double ld_peer<double>(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
template <> __device__ __forceinline__ float ld_peer<float>(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// elements [slice_begin(r), slice_begin(r + 1)) belong to rank r: n/G each, the first n%G ranks one more
__device__ __forceinline__ uint64_t slice_begin(uint64_t n, uint32_t world, uint32_t r) {
    const uint64_t q = n / world, rem = n % world;
    return q * r + (r < rem ? r : rem);
}

// Block b of this rank meets block b of every peer.  PHASE 1 (end) publishes this block's peer stores first.
template <int PHASE>
__device__ bool p2p_barrier(const P2pArgs& a) {
    if (PHASE == 1) __threadfence_system();
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < a.world) {
        const uint32_t peer = threadIdx.x;
        const size_t slot = ((size_t)PHASE * P2P_MAX_BLOCKS + blockIdx.x) * P2P_MAX_RANKS;
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

template <typename T>
__global__ void __launch_bounds__(P2P_THREADS) rc_p2p(const P2pArgs a) {
    if (!p2p_barrier<0>(a)) {
        if (threadIdx.x == 0) atomicAdd(a.err, 1ull);
        return;
    }
    T* const xs = static_cast<T*>(a.xs);
    const uint64_t lo = slice_begin(a.n, a.world, a.rank), hi = slice_begin(a.n, a.world, a.rank + 1);
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (uint64_t i = lo + first; i < hi; i += stride) {
        const T s = xs[i];                                    // identical on every replica (invariant of the reconcile)
        double sum = 0.0;
        uint32_t moved = 0;
        T only = s;                                           // the value of the one replica that moved it, if just one did
        for (uint32_t g = 0; g < a.world; ++g) {
            const T v = ld_peer<T>(static_cast<const T*>(a.x[g]) + i);
            if (v != s) { sum += (double)v - (double)s; ++moved; only = v; }
        }
        const T nv = moved <= 1 ? only : (T)((double)s + sum / (double)moved);      // exact when 0 or 1 replicas moved it
        for (uint32_t g = 0; g < a.world; ++g) static_cast<T*>(a.x[g])[i] = nv;
    }
    if (!p2p_barrier<1>(a)) {
        if (threadIdx.x == 0) atomicAdd(a.err, 1ull);
        return;
    }
    // Refresh the local snapshot (local traffic only).  The barriers pair block b with block b of every peer, so
    // this thread may only read what the SAME (block, thread) of rank g wrote: walk every rank's slice with the
    // partition the data phase used.
    const T* const x = static_cast<const T*>(a.x[a.rank]);
    for (uint32_t g = 0; g < a.world; ++g) {
        const uint64_t glo = slice_begin(a.n, a.world, g), ghi = slice_begin(a.n, a.world, g + 1);
        for (uint64_t i = glo + first; i < ghi; i += stride) xs[i] = ld_peer<T>(x + i);
    }
}

uint64_t align_up(uint64_t v) { return (v + P2P_ALIGN - 1) / P2P_ALIGN * P2P_ALIGN; }

}  // namespace

struct gfs_p2p_region {
    int device = 0;
    uint64_t n = 0;
    uint32_t elem_bytes = 8;
    uint32_t rank = 0, world = 1, blocks = 1, tag = 0;
    size_t off_xs = 0, off_flags = 0, off_err = 0, bytes = 0;
    char* base = nullptr;                               // this rank's allocation
    char* peer_base[P2P_MAX_RANKS] = {};                // every rank's allocation as mapped here
    bool ipc_opened[P2P_MAX_RANKS] = {};
    bool connected = false;
    uint64_t spin_cap = 0;
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
    const uint64_t arr = align_up(std::max<uint64_t>(n, 1) * elem_bytes);
    r->off_xs = arr;
    r->off_flags = 2 * arr;
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

extern "C" int gfs_p2p_region_ptrs(gfs_p2p_region* r, void** x, void** x_sync, uint64_t* region_bytes) {
    if (!r) { gfs::set_error("gfs_p2p_region_ptrs: null region"); return GFS_ERR_INVALID; }
    if (x) *x = r->base;
    if (x_sync) *x_sync = r->base + r->off_xs;
    if (region_bytes) *region_bytes = r->bytes;
    return GFS_OK;
}

extern "C" int gfs_p2p_region_ipc_handle(gfs_p2p_region* r, uint8_t* handle /*GFS_P2P_HANDLE_BYTES*/) {
    if (!r || !handle) { gfs::set_error("gfs_p2p_region_ipc_handle: null argument"); return GFS_ERR_INVALID; }
    static_assert(sizeof(cudaIpcMemHandle_t) == GFS_P2P_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
    P2P_CUDA(cudaSetDevice(r->device));
    cudaIpcMemHandle_t h;
    P2P_CUDA(cudaIpcGetMemHandle(&h, r->base));
    std::memcpy(handle, &h, sizeof h);
    return GFS_OK;
}

extern "C" int gfs_p2p_region_connect_ipc(gfs_p2p_region* r, const uint8_t* handles /*world x 64, rank order*/, uint32_t world,
                                          uint32_t rank) {
    if (!r || !handles) { gfs::set_error("gfs_p2p_region_connect_ipc: null argument"); return GFS_ERR_INVALID; }
    if (world == 0 || world > P2P_MAX_RANKS || rank >= world) { gfs::set_error("gfs_p2p_region_connect_ipc: bad world / rank"); return GFS_ERR_INVALID; }
    if (r->connected) { gfs::set_error("gfs_p2p_region_connect_ipc: region already connected"); return GFS_ERR_INVALID; }
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
    for (uint32_t a = 0; a < world; ++a) {
        P2P_CUDA(cudaSetDevice(regions[a]->device));
        for (uint32_t b = 0; b < world; ++b) {
            if (regions[b]->device != regions[a]->device) {
                int can = 0;
                P2P_CUDA(cudaDeviceCanAccessPeer(&can, regions[a]->device, regions[b]->device));
                if (!can) { gfs::set_error("gfs_p2p_region_connect_local: devices cannot access each other's memory"); return GFS_ERR_INVALID; }
                cudaError_t e = cudaDeviceEnablePeerAccess(regions[b]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { gfs::set_error(std::string("cudaDeviceEnablePeerAccess failed: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
                (void)cudaGetLastError();
            }
            regions[a]->peer_base[b] = regions[b]->base;
        }
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

// Asynchronous on `stream`.  Every rank calls it once per reconcile, in the same order.
extern "C" int gfs_p2p_reconcile(gfs_p2p_region* r, void* stream) {
    if (!r) { gfs::set_error("gfs_p2p_reconcile: null region"); return GFS_ERR_INVALID; }
    if (!r->connected) { gfs::set_error("gfs_p2p_reconcile: region is not connected to its peers"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    P2pArgs a{};
    for (uint32_t g = 0; g < r->world; ++g) {
        a.x[g] = r->peer_base[g];
        a.flags[g] = reinterpret_cast<uint32_t*>(r->peer_base[g] + r->off_flags);
    }
    a.xs = r->base + r->off_xs;
    a.err = reinterpret_cast<unsigned long long*>(r->base + r->off_err);
    a.n = r->n; a.spin_cap = r->spin_cap;
    a.rank = r->rank; a.world = r->world;
    a.tag = ++r->tag;
    if (r->elem_bytes == 8) rc_p2p<double><<<r->blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(a);
    else rc_p2p<float><<<r->blocks, P2P_THREADS, 0, (cudaStream_t)stream>>>(a);
    P2P_CUDA(cudaGetLastError());
    return GFS_OK;
}

// Blocking: non-zero (GFS_ERR_CUDA) when a barrier of any earlier reconcile timed out.
extern "C" int gfs_p2p_region_check(gfs_p2p_region* r) {
    if (!r) { gfs::set_error("gfs_p2p_region_check: null region"); return GFS_ERR_INVALID; }
    P2P_CUDA(cudaSetDevice(r->device));
    unsigned long long err = 0;
    P2P_CUDA(cudaMemcpy(&err, r->base + r->off_err, 8, cudaMemcpyDeviceToHost));
    if (err) {
        gfs::set_error("peer-memory reconcile: a barrier timed out (a rank did not reach the reconcile, or its kernel was not resident)");
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
