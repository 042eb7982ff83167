// gfs_internal.h — host-side objects shared by the translation units of libgfasort_cuda.so
// (gfs_lib.cu: index + sessions + C ABI; gfs_multi.cu: replicated multi-GPU runs).  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gfasort_cuda.h"
#include "gfs_device.cuh"

namespace gfs {

void set_error(const std::string& s);

#define GFS_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            gfs::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                           std::to_string(__LINE__) + ")");                                         \
            return GFS_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

// Scoped device buffer / stream: transient allocations are released on every return path.
template <typename T> struct DevBuf {
    T* p = nullptr;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { cudaFree(p); }
    void release() { cudaFree(p); p = nullptr; }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    cudaError_t up(const T* h, size_t n) { return cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T* h, size_t n) { return cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};
struct ScopedStream {
    cudaStream_t st = nullptr;
    ScopedStream() = default;
    ScopedStream(const ScopedStream&) = delete;
    ScopedStream& operator=(const ScopedStream&) = delete;
    ~ScopedStream() { if (st) cudaStreamDestroy(st); }
    cudaError_t create() { return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking); }
};

inline double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
inline long env_long(const char* name, long dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    return std::strtol(v, nullptr, 10);
}

// coordinate stride per node end for `dims` layout dimensions (0 = the 1D sort)
inline uint32_t coord_stride(uint32_t dims) { return dims <= 1 ? 1u : dims == 2 ? 2u : dims <= 4 ? 4u : 8u; }

}  // namespace gfs

struct gfs_index {
    int device = 0;
    uint64_t S = 0, P = 0, N = 0;
    uint64_t max_path_steps = 0;
    bool any_multi_step = false;
    gfs::StepRec* d_recs = nullptr;
    uint64_t* d_first_step = nullptr;   // P+1
    uint64_t* d_path_len = nullptr;     // P
    uint32_t* d_new_of_old = nullptr;   // N, null when not relabelled
    uint32_t* d_old_of_new = nullptr;   // N
    void* d_build_arena = nullptr;      // the build's transient device memory, released with the index
    // the zeta table of the last (theta, space, space_max, q) asked for: every session of a run needs the same one
    mutable std::mutex zeta_mu;
    mutable std::vector<double> zeta_cache;
    mutable double zeta_key[4] = {0, 0, 0, 0};
    std::vector<uint64_t> h_first_step;
    double build_seconds = 0, h2d_seconds = 0, kernel_seconds = 0, alloc_seconds = 0, relabel_seconds = 0;
    uint64_t launches = 0;
    // A multi-GPU index (gfs_index_build under GFASORT_GPUS > 1) is a directory of per-device shards:
    // shard g holds the records of the paths that step slice g overlaps.  The fields above
    // then describe the WHOLE graph (S, P, N, h_first_step, max_path_steps) and the device pointers are null.
    std::vector<gfs_index*> shards;
    std::vector<gfs_shard_plan> plans;
};

struct gfs_sgd_session {
    const gfs_index* ix = nullptr;
    gfs_sgd_params params{};
    uint32_t dims = 0;          // 0 = 1D
    uint32_t DS = 1;            // coordinate stride per node end
    bool f64 = true;
    bool aggregate = true;
    int device = 0;
    uint32_t grid = 0, block = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    void* d_pos = nullptr;
    bool own_pos = false;
    uint64_t n_elems = 0;       // elements in d_pos
    double2* d_zetas = nullptr; uint32_t zlen = 0;      // 2 x zlen {zeta, den} pairs (gfs_lib.cu h_zeta_tables)
    gfs::EpochDesc* d_epochs = nullptr; uint32_t n_epochs = 0;
    uint64_t* d_attempts = nullptr;
    unsigned long long* d_counters = nullptr;
    double* d_stage = nullptr;  // f64 staging for nD conversions
    size_t smem_bytes = 0;
    uint32_t zeta_smem = 0;     // experiment: zeta entries per theta staged in shared memory (GFASORT_ZETA_SMEM)
    uint64_t rng_thread_base = 0;
    // kernel timing: a ring of event pairs, so that enqueueing a launch never waits for the previous one (a host that
    // blocks per launch leaves the GPU idle between an epoch's reconcile and the next epoch's kernel)
    static constexpr int EV_RING = 16;
    cudaEvent_t ev0[EV_RING] = {}, ev1[EV_RING] = {};
    uint32_t ev_head = 0, ev_tail = 0;      // pairs [ev_tail, ev_head) are recorded and not yet read
    double kernel_ms = 0.0;
    uint64_t launches = 0;
    double h2d_s = 0, d2h_s = 0;
    int inflight = 2;           // terms in flight per thread (kernel template parameter K)
    uint32_t coherent = 32;     // sweep schedule: groups of this many lanes sample consecutive steps (0 = off)
    uint64_t window_steps = 0;  // 0 = static schedule
    uint32_t chunk_updates = 128;
    unsigned long long* d_work = nullptr;
    void* d_saved = nullptr;    // gfs_sgd_session_save snapshot of the positions
    uint64_t samp_base = 0, samp_len = 0;
};

// gfs_multi.cu: gfs_sgd_1d / gfs_sgd_nd on a multi-GPU index (one replica per shard, peer-memory reconcile)
int gfs_multi_run_whole(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg, uint32_t dims,
                        double* pos_inout, gfs_stats* stats);
