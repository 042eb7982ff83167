// gfs_lib.cu — libgfasort_cuda.so: kernels + C ABI (include/gfasort_cuda.h) for the path-guided SGD
// hot path of pangenome/gfasort on B200 (sm_100a).
//
// This file: host-side objects (index, session), the schedule / zeta-table twins of the reference's
// scalar helpers, and every `extern "C"` entry point.  The kernels live in the headers it includes:
//   gfs_device.cuh          Philox4x32-10, fast_precise_pow, DirtyZipfian, record loads, warp group sums
//   gfs_kernels_index.cuh   K1  path index (k1_*) and node relabelling (rl_*).   PathIndex::from_graph, sgd.rs:34-71
//   gfs_kernels_sgd.cuh     K2/K3 sgd_kernel<CT, D, DS, AGG, K>: the persistent, software-pipelined term kernel
//                           (1D `Y`, f64; nD `L`, f32/f64).   Worker loops sgd.rs:442-584, 988-1156; checkers :366-407, :925-955
//   gfs_kernels_aux.cuh     K4 stress_kernel (calculate_layout_stress, :1196-1283), layout conversions,
//                           K5 rc_pack / rc_apply (multi-GPU reconcile), K6 rs_* (order by position), debug kernels
// Host code in other files: gfs_synth.cpp (synthetic graphs), gfs_host_graph.cpp (linear-time `g` / `s`),
// gfs_io.cpp (flat GFA ingest, buffered writers).
//
// There is no CPU implementation of any kernel in this library: if CUDA is unavailable every compute
// entry point returns an error.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "gfs_internal.h"

namespace gfs {

static thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }

// =============================================================================================
// Host twins of the reference's scalar helpers (schedule / zeta table / per-epoch constants).
// Compiled with -ffp-contract=off; x86-64 baseline has no FMA, so nothing is fused.
// =============================================================================================
static inline int32_t h_f64_as_i32(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}
static double h_fast_precise_pow(double a, double b) {      // src/sgd.rs:155-182
    int32_t e = h_f64_as_i32(b);
    uint64_t bits; std::memcpy(&bits, &a, 8);
    int32_t diff = (int32_t)((uint32_t)(bits >> 32) - 1072632447u);
    int32_t nh = h_f64_as_i32((b - (double)e) * (double)diff + 1072632447.0);
    uint64_t fb = ((uint64_t)(int64_t)nh) << 32;
    double frac; std::memcpy(&frac, &fb, 8);
    double base = a, r = 1.0;
    for (int32_t ex = e; ex != 0; ex >>= 1) { if (ex & 1) r *= base; base *= base; }
    return r * frac;
}
static void h_schedule(const gfs_sgd_params& p, std::vector<double>& etas) {   // src/sgd.rs:617-638
    const double w_min = 1.0 / p.eta_max, w_max = 1.0;
    const double eta_max = 1.0 / w_min, eta_min = p.eps / w_max;
    const double lambda = std::log(eta_max / eta_min) / ((double)p.iter_max - 1.0);
    etas.resize(p.iter_max + 1);
    for (uint64_t t = 0; t <= p.iter_max; ++t) {
        int64_t dt = (int64_t)t - (int64_t)p.iter_with_max_learning_rate;
        if (dt < 0) dt = -dt;
        etas[t] = eta_max * std::exp(-lambda * (double)dt);
    }
}
// src/sgd.rs:311-331.  The serial summation order is kept (it defines the values); the loop stops
// at the largest jump_space any path can produce (max_path_steps), since entries beyond it are
// unreachable (jump_space = min(space, rank) <= max_path_steps - 1, src/sgd.rs:462,477).
// The reference's loop is one dependent chain of `space` iterations (3e8 at config 3; 1e7 even after the cap).
// Only the ADDITIONS depend on one another: the terms fpp(1/i, theta) are computed block by block on several host
// threads, then added up in the reference's order by one — same operations, same order, same bits, a few
// milliseconds instead of 0.1 s (config 3) / 1 s (config 5) per session.
static void h_zetas(const gfs_sgd_params& p, uint64_t max_path_steps, std::vector<double>& z) {
    const uint64_t sm = p.space_max, q = p.space_quantization_step ? p.space_quantization_step : 1;
    const uint64_t full = ((p.space <= sm) ? p.space : sm + (p.space - sm) / q + 1) + 1;
    const uint64_t reach = std::min(p.space, max_path_steps);
    const uint64_t reach_idx = reach > sm ? sm + (reach - sm) / q + 1 : reach;
    const uint64_t n = std::min(full, reach_idx + 2);
    z.assign(n, 0.0);
    constexpr uint64_t BLK = 1u << 16;
    const int T = (int)std::max<long>(1, std::min<long>(8, (long)std::thread::hardware_concurrency()));
    const double theta = p.theta;
    std::vector<double> terms((size_t)std::min<uint64_t>(reach, BLK * T));
    double acc = 0.0;
    for (uint64_t i0 = 1; i0 <= reach; i0 += BLK * T) {
        const uint64_t cnt = std::min<uint64_t>(BLK * T, reach - i0 + 1);
        auto fill = [&](uint64_t lo, uint64_t hi) { for (uint64_t k = lo; k < hi; ++k) terms[k] = h_fast_precise_pow(1.0 / (double)(i0 + k), theta); };
        if (cnt < BLK || T == 1) fill(0, cnt);
        else {
            std::vector<std::thread> pool;
            const uint64_t per = (cnt + T - 1) / T;
            for (int t = 0; t < T; ++t) { const uint64_t lo = std::min(cnt, per * t), hi = std::min(cnt, per * (t + 1)); if (hi > lo) pool.emplace_back(fill, lo, hi); }
            for (auto& th : pool) th.join();
        }
        for (uint64_t k = 0; k < cnt; ++k) {
            const uint64_t i = i0 + k;
            acc += terms[k];
            if (i <= sm) { if (i < n) z[i] = acc; }
            if (i >= sm && (i - sm) % q == 0) {
                const uint64_t idx = sm + 1 + (i - sm) / q;
                if (idx < n) z[idx] = acc;
            }
        }
    }
}

// What the kernels read: for each of the run's two thetas (params.theta while warm, 0.001 while cooling, sgd.rs:394)
// a table of {zetas[k], 1 - zeta2theta/zetas[k]} — the second member is the denominator of DirtyZipfian's eta
// (sgd.rs:133-134), computed here with the same two IEEE operations the reference performs per sample.
static void h_zeta_tables(const gfs_sgd_params& p, uint64_t max_path_steps, std::vector<double2>& out, uint32_t& zlen,
                          const gfs_index* cache_in = nullptr) {
    std::vector<double> z;
    if (cache_in) {
        std::lock_guard<std::mutex> lk(cache_in->zeta_mu);
        const double key[4] = {p.theta, (double)p.space, (double)p.space_max, (double)p.space_quantization_step};
        if (cache_in->zeta_cache.empty() || std::memcmp(key, cache_in->zeta_key, sizeof key) != 0) {
            h_zetas(p, max_path_steps, cache_in->zeta_cache);
            std::memcpy(cache_in->zeta_key, key, sizeof key);
        }
        z = cache_in->zeta_cache;
    } else {
        h_zetas(p, max_path_steps, z);
    }
    zlen = (uint32_t)z.size();
    out.resize(2 * z.size());
    const double thetas[2] = {p.theta, 0.001};
    for (int t = 0; t < 2; ++t) {
        const double z2 = 1.0 + h_fast_precise_pow(0.5, thetas[t]);
        for (size_t k = 0; k < z.size(); ++k) out[t * z.size() + k] = make_double2(z[k], 1.0 - z2 / z[k]);
    }
}

static void h_epochs(const gfs_sgd_params& p, std::vector<EpochDesc>& out) {
    std::vector<double> etas;
    h_schedule(p, etas);
    const uint64_t first_cooling = (uint64_t)std::floor(p.cooling_start * (double)p.iter_max);   // sgd.rs:297
    out.resize(p.iter_max + 1);
    for (uint64_t e = 0; e <= p.iter_max; ++e) {
        EpochDesc d{};
        d.eta = etas[e];
        d.cooling = e > first_cooling ? 1u : 0u;                    // strict (sgd.rs:393)
        d.ztab = d.cooling;
        const double theta = d.cooling ? 0.001 : p.theta;           // sgd.rs:394
        d.zc.theta = theta;
        d.zc.one_minus_theta = 1.0 - theta;
        d.zc.alpha = 1.0 / (1.0 - theta);
        d.zc.z2 = 1.0 + h_fast_precise_pow(0.5, theta);
        d.zc.alpha_e = h_f64_as_i32(d.zc.alpha);
        d.zc.alpha_frac = d.zc.alpha - (double)d.zc.alpha_e;
        d.updates = p.min_term_updates;
        out[e] = d;
    }
}

}  // namespace gfs

#include "gfs_kernels_index.cuh"
#include "gfs_kernels_sgd.cuh"
#include "gfs_kernels_aux.cuh"

// =============================================================================================
// Host-side objects
// =============================================================================================
using namespace gfs;

static int select_device(int dev) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                  " — libgfasort_cuda has no CPU fallback");
        return GFS_ERR_NO_DEVICE;
    }
    if (dev < 0) {
        long envd = env_long("GFASORT_DEVICE", -1);
        if (envd >= 0) dev = (int)envd;
        else { GFS_CUDA(cudaGetDevice(&dev)); }
    }
    if (dev >= n) { set_error("device ordinal out of range"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(dev));
    return GFS_OK;
}

extern "C" const char* gfs_last_error(void) { return g_last_error.c_str(); }

extern "C" const char* gfs_device_info(void) {
    static thread_local std::string info;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { info = "{\"devices\": 0}"; return info.c_str(); }
    int dev = 0; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    char buf[512];
    std::snprintf(buf, sizeof buf,
                  "{\"devices\": %d, \"device\": %d, \"name\": \"%s\", \"cc\": \"%d.%d\", \"sms\": %d, \"l2_bytes\": %d, "
                  "\"global_mem\": %zu, \"persisting_l2_max\": %d, \"access_policy_max_window\": %d}",
                  n, dev, p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize, (size_t)p.totalGlobalMem,
                  p.persistingL2CacheMaxSize, p.accessPolicyMaxWindowSize);
    info = buf;
    return info.c_str();
}

// ---------------------------------------------------------------------------------------------
// index build
// ---------------------------------------------------------------------------------------------
static int radix_sort_pairs(uint64_t*& k0, uint64_t*& k1, uint32_t*& v0, uint32_t*& v1, uint32_t* hist, uint64_t n, int passes,
                            cudaStream_t st, uint64_t* launches);

// Host threads that copy a pageable source into a pinned bounce buffer, slice by slice: one thread tops out well
// below what a PCIe 5 x16 link drains.  The workers live for one index build and spin (yielding) between copies,
// which come a few hundred microseconds apart.
struct CopyPool {
    std::vector<std::thread> workers;
    std::atomic<uint64_t> seq{0};
    std::atomic<uint32_t> done{0};
    std::atomic<bool> stop{false};
    char* dst = nullptr; const char* src = nullptr; size_t len = 0;
    int T = 1;
    void slice(int t) const {
        const size_t per = ((len + T - 1) / T + 4095) & ~(size_t)4095;
        const size_t lo = std::min(len, per * t), hi = std::min(len, per * (t + 1));
        if (hi > lo) std::memcpy(dst + lo, src + lo, hi - lo);
    }
    explicit CopyPool(int threads) : T(std::max(threads, 1)) {
        for (int t = 1; t < T; ++t)
            workers.emplace_back([this, t] {
                uint64_t seen = 0;
                for (;;) {
                    uint64_t cur;
                    while ((cur = seq.load(std::memory_order_acquire)) == seen) {
                        if (stop.load(std::memory_order_relaxed)) return;
                        std::this_thread::yield();
                    }
                    seen = cur;
                    slice(t);
                    done.fetch_add(1, std::memory_order_release);
                }
            });
    }
    void copy(void* d, const void* s, size_t n) {
        dst = static_cast<char*>(d); src = static_cast<const char*>(s); len = n;
        done.store(0, std::memory_order_relaxed);
        seq.fetch_add(1, std::memory_order_release);
        slice(0);
        while (done.load(std::memory_order_acquire) != (uint32_t)(T - 1)) std::this_thread::yield();
    }
    ~CopyPool() {
        stop.store(true);
        for (auto& w : workers) w.join();
    }
};

// GFASORT_BUILD_TRACE=1: phase times of the index build on stderr
static thread_local double g_trace_last = 0.0;
static void trace_mark(const char* what) {
    if (g_trace_last == 0.0) return;
    const double t = now_s();
    std::fprintf(stderr, "[gfs_index_build]   %-26s +%.4f s\n", what, t - g_trace_last);
    g_trace_last = t;
}

static int relabel_rewrite(gfs_index* ix, cudaStream_t st) {
    const uint32_t N = (uint32_t)ix->N;
    if (ix->S) rl_rewrite<<<(unsigned)((ix->S + 255) / 256), 256, 0, st>>>(ix->d_recs, ix->S, N, ix->d_new_of_old);
    ix->launches += ix->S ? 1 : 0;
    GFS_CUDA(cudaStreamSynchronize(st));
    GFS_CUDA(cudaGetLastError());
    trace_mark("rl_rewrite + sync");
    return GFS_OK;
}

// Relabel a built index with the caller's permutation new_of_old[N] (host pointer).
static int index_relabel_given(gfs_index* ix, const uint32_t* given, cudaStream_t st) {
    if (ix->N == 0) return GFS_OK;
    const uint32_t N = (uint32_t)ix->N;
    if (!ix->d_new_of_old) GFS_CUDA(cudaMalloc(&ix->d_new_of_old, (size_t)N * 4));
    if (!ix->d_old_of_new) GFS_CUDA(cudaMalloc(&ix->d_old_of_new, (size_t)N * 4));
    GFS_CUDA(cudaMemcpyAsync(ix->d_new_of_old, given, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    rl_invert<<<(N + 255) / 256, 256, 0, st>>>(ix->d_new_of_old, N, ix->d_old_of_new);
    ix->launches += 1;
    return relabel_rewrite(ix, st);
}

// Relabel by order of first appearance: a stable radix sort of the N first-occurrence keys K1 left in the
// node table (32-bit keys: 4 passes), never-visited nodes last.
// sort scratch of the relabelling for N nodes: k0[N] u64 | k1[N] u64 | v1[N] u32 | hist[256 * n_blocks] u32
static uint64_t relabel_scratch_bytes(uint64_t N) {
    const uint64_t n_blocks = (N + RS_TILE - 1) / RS_TILE;
    auto up256 = [](uint64_t v) { return (v + 255) / 256 * 256; };
    return up256(N * 8) * 2 + up256(N * 4) + up256(256 * n_blocks * 4);
}
// ix->d_new_of_old / d_old_of_new are allocated by the caller; `scratch` holds relabel_scratch_bytes(N) bytes.
static int index_relabel_first_occ(gfs_index* ix, const uint32_t* d_first_key, char* scratch, cudaStream_t st) {
    if (ix->N == 0) return GFS_OK;
    const uint32_t N = (uint32_t)ix->N;
    auto up256 = [](uint64_t v) { return (v + 255) / 256 * 256; };
    uint64_t* k0 = reinterpret_cast<uint64_t*>(scratch);
    uint64_t* k1 = reinterpret_cast<uint64_t*>(scratch + up256((uint64_t)N * 8));
    uint32_t* v1 = reinterpret_cast<uint32_t*>(scratch + 2 * up256((uint64_t)N * 8));
    uint32_t* hist = reinterpret_cast<uint32_t*>(scratch + 2 * up256((uint64_t)N * 8) + up256((uint64_t)N * 4));
    uint32_t* v0 = ix->d_old_of_new;
    rl_keys<<<(N + 255) / 256, 256, 0, st>>>(d_first_key, N, k0, v0);
    ix->launches += 1;
    int rc = radix_sort_pairs(k0, k1, v0, v1, hist, N, 4, st, &ix->launches);     // even number of passes: result in v0
    if (rc) return rc;
    rl_invert<<<(N + 255) / 256, 256, 0, st>>>(ix->d_old_of_new, N, ix->d_new_of_old);
    ix->launches += 1;
    trace_mark("relabel: sort enqueued");
    return relabel_rewrite(ix, st);
}

// PathIndex::from_graph (src/sgd.rs:34-71) for paths [path_begin, path_end) on `device`.
// The step array is streamed through two device staging buffers in chunks: the copy of chunk c+1 (stream
// s_copy) runs under the K1 kernel of chunk c (stream s_k).  A page-locked source is read by the copy engine
// directly; a pageable one (a Rust Vec) goes through two pinned bounce buffers filled by several host threads,
// so that the host-side staging of chunk c+1 also runs under the DMA of chunk c.
// relabel_mode: 0 none, 1 first-appearance order, 2 the given permutation.
template <typename HT>
static int build_shard_impl(const HT* step_handles, const uint64_t* path_first_step, const uint32_t* node_len, uint64_t S,
                            uint64_t P, uint64_t N, uint64_t path_begin, uint64_t path_end, int32_t device,
                            int32_t relabel_mode, const uint32_t* new_of_old, gfs_index** out) {
    if (!out) { set_error("gfs_index_build: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    if (!path_first_step || (S && !step_handles) || (N && !node_len)) { set_error("gfs_index_build: null input array"); return GFS_ERR_INVALID; }
    if (path_begin > path_end || path_end > P) { set_error("gfs_index_build: bad path range"); return GFS_ERR_INVALID; }
    if (N >= (1ull << 31)) { set_error("gfs_index_build: N must be < 2^31"); return GFS_ERR_INVALID; }
    if (path_first_step[0] != 0 || path_first_step[P] != S) { set_error("gfs_index_build: path_first_step must start at 0 and end at S"); return GFS_ERR_INVALID; }
    for (uint64_t p = 0; p < P; ++p) {
        if (path_first_step[p + 1] < path_first_step[p]) { set_error("gfs_index_build: path_first_step not monotone"); return GFS_ERR_INVALID; }
        if (path_first_step[p + 1] - path_first_step[p] >= (1ull << 32)) { set_error("gfs_index_build: a path has >= 2^32 steps"); return GFS_ERR_INVALID; }
    }
    if (path_end - path_begin >= (1ull << 31)) { set_error("gfs_index_build: too many paths"); return GFS_ERR_INVALID; }
    if (relabel_mode == 2 && !new_of_old) { set_error("gfs_index_build: relabel_mode 2 needs a permutation"); return GFS_ERR_INVALID; }
    int rc = select_device(device);
    if (rc) return rc;

    const double t_begin = now_s();
    const bool trace = env_long("GFASORT_BUILD_TRACE", 0) != 0;
    double t_last = t_begin;
    auto mark = [&](const char* what) {
        if (!trace) return;
        const double t = now_s();
        std::fprintf(stderr, "[gfs_index_build] %-28s +%.4f s (at %.4f)\n", what, t - t_last, t - t_begin);
        t_last = t;
    };
    gfs_index* ix = new gfs_index();
    cudaGetDevice(&ix->device);
    const uint64_t s_begin = path_first_step[path_begin], s_end = path_first_step[path_end];
    ix->S = s_end - s_begin; ix->P = path_end - path_begin; ix->N = N;
    ix->h_first_step.resize(ix->P + 1);
    for (uint64_t p = 0; p <= ix->P; ++p) ix->h_first_step[p] = path_first_step[path_begin + p] - s_begin;
    for (uint64_t p = 0; p < ix->P; ++p) {
        const uint64_t c = ix->h_first_step[p + 1] - ix->h_first_step[p];
        ix->max_path_steps = std::max(ix->max_path_steps, c);
        if (c > 1) ix->any_multi_step = true;
    }
    auto fail = [&](int code) { gfs_index_free(ix); return code; };
#define IX_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__)); return fail(GFS_ERR_CUDA); } } while (0)

    ScopedStream sk, sc;
    IX_CUDA(sk.create());
    IX_CUDA(sc.create());
    cudaStream_t s_k = sk.st, s_copy = sc.st;
    const bool first_occ = relabel_mode == 1;
    uint32_t key_shift = 0;
    while ((ix->S >> key_shift) >= 0xffffffffull) ++key_shift;

    // chunks of at most CH steps (a multiple of the tile, so no tile straddles two chunks)
    uint64_t CH = (uint64_t)env_long("GFASORT_INDEX_CHUNK", 1l << 24);
    CH = std::max<uint64_t>(K1_TILE, (CH / K1_TILE) * K1_TILE);
    const uint64_t chunk_cap = std::min<uint64_t>(CH, std::max<uint64_t>(ix->S, 1));
    const uint64_t n_chunks = (ix->S + CH - 1) / CH;
    const uint64_t tiles_total = (ix->S + K1_TILE - 1) / K1_TILE;

    IX_CUDA(cudaMalloc(&ix->d_first_step, (ix->P + 1) * 8));
    IX_CUDA(cudaMalloc(&ix->d_path_len, std::max<uint64_t>(ix->P, 1) * 8));
    // records and staged handles are padded to whole tiles: K1 works without bounds checks (padding handles are all-ones,
    // i.e. "missing node"; padding records are written and never read)
    IX_CUDA(cudaMalloc(&ix->d_recs, std::max<uint64_t>(tiles_total * K1_TILE, 1) * sizeof(StepRec)));
    // everything transient in ONE allocation: [descriptors | flags | node table (length + visited bit) | node_len staging | first keys | 2 handle buffers]
    const uint64_t chunk_alloc = (chunk_cap + K1_TILE - 1) / K1_TILE * K1_TILE;
    auto up256 = [](uint64_t v) { return (v + 255) / 256 * 256; };
    const uint64_t o_desc = 0, o_flag = o_desc + up256((tiles_total + 1) * 8), o_len = o_flag + 256, o_vis = o_len + up256(N * 4 + 4),
                   o_key = o_vis + up256(N * 4 + 4), o_h0 = o_key + up256(N * 4 + 4), o_h1 = o_h0 + up256(chunk_alloc * sizeof(HT)),
                   k1_bytes = o_h1 + (n_chunks > 1 ? up256(chunk_alloc * sizeof(HT)) : 0),
                   // the relabelling's sort scratch reuses the handle buffers (idle once K1 has drained)
                   arena_bytes = std::max(k1_bytes, relabel_mode == 1 ? o_h0 + relabel_scratch_bytes(N) : 0);
    // cudaMalloc / cudaFree calls are kept to a handful per build, all before the first copy: on this pool single
    // calls sporadically stall for 0.5-1.7 s when issued between kernels (profiles/r2_index_build.md); the arena is
    // handed to the index and freed with it
    struct Arena { char* p = nullptr; } arena;
    IX_CUDA(cudaMalloc(&ix->d_build_arena, std::max<uint64_t>(arena_bytes, 256)));
    arena.p = static_cast<char*>(ix->d_build_arena);
    if (relabel_mode != 0 && N) {
        IX_CUDA(cudaMalloc(&ix->d_new_of_old, (size_t)N * 4));
        IX_CUDA(cudaMalloc(&ix->d_old_of_new, (size_t)N * 4));
    }
    uint64_t* d_desc = reinterpret_cast<uint64_t*>(arena.p + o_desc);
    unsigned int* d_flag = reinterpret_cast<unsigned int*>(arena.p + o_flag);
    uint32_t* d_node_tbl = reinterpret_cast<uint32_t*>(arena.p + o_len);
    uint32_t* d_node_len = reinterpret_cast<uint32_t*>(arena.p + o_vis);       // staging of the caller's node_len
    uint32_t* d_first_key = reinterpret_cast<uint32_t*>(arena.p + o_key);
    HT* d_h[2] = {reinterpret_cast<HT*>(arena.p + o_h0), reinterpret_cast<HT*>(arena.p + o_h1)};
    IX_CUDA(cudaMemsetAsync(arena.p, 0, o_len, s_k));                                   // descriptors, flags
    if (env_long("GFASORT_K1_NO_LOOKBACK", 0)) {                                        // timing experiment: offsets are WRONG
        const unsigned int one = 1;
        IX_CUDA(cudaMemcpyAsync(d_flag + 1, &one, sizeof one, cudaMemcpyHostToDevice, s_k));
        IX_CUDA(cudaStreamSynchronize(s_k));
    }
    IX_CUDA(cudaMemsetAsync(d_first_key, 0xff, o_h0 - o_key, s_k));                     // 0xffffffff = never visited
    IX_CUDA(cudaMemsetAsync(ix->d_path_len, 0, std::max<uint64_t>(ix->P, 1) * 8, s_k));
    mark("device allocations");
    const double t_h2d0 = now_s();
    ix->alloc_seconds = t_h2d0 - t_begin;
    IX_CUDA(cudaMemcpyAsync(ix->d_first_step, ix->h_first_step.data(), (ix->P + 1) * 8, cudaMemcpyHostToDevice, s_k));
    if (N) {
        IX_CUDA(cudaMemcpyAsync(d_node_len, node_len, N * 4, cudaMemcpyHostToDevice, s_k));
        k1_init_table<<<(unsigned)((N + 255) / 256), 256, 0, s_k>>>(d_node_len, (uint32_t)N, d_node_tbl, d_flag);
        ix->launches += 1;
    }

    // is the caller's step array page-locked?  (cudaHostAlloc / cudaHostRegister memory: the copy engine reads it directly)
    bool src_pinned = false;
    if (ix->S) {
        cudaPointerAttributes attr{};
        if (cudaPointerGetAttributes(&attr, step_handles + s_begin) == cudaSuccess) src_pinned = attr.type == cudaMemoryTypeHost;
        else (void)cudaGetLastError();
    }
    // pageable source: NB bounce buffers of BB bytes (page-locking memory costs ~1 ms per MB, so they are small and
    // decoupled from the device chunk: a chunk arrives as several sub-copies)
    constexpr int NB = 4;
    const size_t BB = (size_t)std::max<long>(1l << 20, env_long("GFASORT_BOUNCE_BYTES", 8l << 20));
    struct Pinned { void* p = nullptr; ~Pinned() { if (p) cudaFreeHost(p); } } pin;
    const int copy_threads = (int)std::max<long>(1, env_long("GFASORT_COPY_THREADS", std::min<long>(8, std::max<long>(1, (long)std::thread::hardware_concurrency() / 2))));
    if (ix->S && !src_pinned) IX_CUDA(cudaHostAlloc(&pin.p, NB * BB, cudaHostAllocDefault));
    std::unique_ptr<CopyPool> pool;
    if (pin.p) pool.reset(new CopyPool(copy_threads));
    mark("pinned bounce buffers");
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr}, ev_b[NB] = {};
    std::vector<cudaEvent_t> ev_t;      // kernel timing: one pair per chunk
    auto drop_events = [&]() {
        for (int b = 0; b < 2; ++b) { if (ev_h2d[b]) cudaEventDestroy(ev_h2d[b]); if (ev_k[b]) cudaEventDestroy(ev_k[b]); }
        for (int b = 0; b < NB; ++b) if (ev_b[b]) cudaEventDestroy(ev_b[b]);
        for (cudaEvent_t e : ev_t) cudaEventDestroy(e);
    };
#define IXE_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__)); drop_events(); return fail(GFS_ERR_CUDA); } } while (0)
    for (int b = 0; b < 2; ++b) {
        IXE_CUDA(cudaEventCreateWithFlags(&ev_h2d[b], cudaEventDisableTiming));
        IXE_CUDA(cudaEventCreateWithFlags(&ev_k[b], cudaEventDisableTiming));
    }
    if (pin.p) for (int b = 0; b < NB; ++b) IXE_CUDA(cudaEventCreateWithFlags(&ev_b[b], cudaEventDisableTiming));
    uint64_t n_sub = 0;                 // bounce sub-copies issued so far (buffer n_sub % NB is next)
    for (uint64_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c & 1);
        const uint64_t c0 = c * CH, clen = std::min(CH, ix->S - c0);
        const size_t bytes = clen * sizeof(HT);
        const HT* src = step_handles + s_begin + c0;
        if (c >= 2) IXE_CUDA(cudaStreamWaitEvent(s_copy, ev_k[b], 0));          // the kernel of chunk c-2 is done with this buffer
        if (src_pinned) {
            IXE_CUDA(cudaMemcpyAsync(d_h[b], src, bytes, cudaMemcpyHostToDevice, s_copy));
        } else {
            for (size_t off = 0; off < bytes; off += BB, ++n_sub) {
                const int k = (int)(n_sub % NB);
                const size_t len = std::min(BB, bytes - off);
                char* stage = static_cast<char*>(pin.p) + (size_t)k * BB;
                if (n_sub >= NB) IXE_CUDA(cudaEventSynchronize(ev_b[k]));       // the DMA that last read this bounce buffer has finished
                pool->copy(stage, reinterpret_cast<const char*>(src) + off, len);
                IXE_CUDA(cudaMemcpyAsync(reinterpret_cast<char*>(d_h[b]) + off, stage, len, cudaMemcpyHostToDevice, s_copy));
                IXE_CUDA(cudaEventRecord(ev_b[k], s_copy));
            }
        }
        const uint64_t cpad = (clen + K1_TILE - 1) / K1_TILE * K1_TILE;
        if (cpad > clen) IXE_CUDA(cudaMemsetAsync(d_h[b] + clen, 0xff, (cpad - clen) * sizeof(HT), s_copy));   // the last tile's padding
        IXE_CUDA(cudaEventRecord(ev_h2d[b], s_copy));
        IXE_CUDA(cudaStreamWaitEvent(s_k, ev_h2d[b], 0));
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        IXE_CUDA(cudaEventCreate(&t0)); ev_t.push_back(t0);
        IXE_CUDA(cudaEventCreate(&t1)); ev_t.push_back(t1);
        IXE_CUDA(cudaEventRecord(t0, s_k));
        const unsigned n_tiles = (unsigned)((clen + K1_TILE - 1) / K1_TILE);
        if (first_occ)
            k1_scan_write<HT, true><<<n_tiles, K1_BLOCK, 0, s_k>>>(d_h[b], d_node_tbl, d_first_key, (uint32_t)N, ix->d_first_step,
                                                                    (uint32_t)ix->P, c0, ix->S, d_desc, d_flag, key_shift, ix->d_recs, ix->d_path_len);
        else
            k1_scan_write<HT, false><<<n_tiles, K1_BLOCK, 0, s_k>>>(d_h[b], d_node_tbl, d_first_key, (uint32_t)N, ix->d_first_step,
                                                                     (uint32_t)ix->P, c0, ix->S, d_desc, d_flag, key_shift, ix->d_recs, ix->d_path_len);
        if (ix->P > 1) {
            k1_fix_path_starts<<<(unsigned)(ix->P - 1), K1_THREADS, 0, s_k>>>(ix->d_first_step, (uint32_t)ix->P, c0, c0 + clen, ix->d_recs, ix->d_path_len);
            ix->launches += 1;
        }
        IXE_CUDA(cudaGetLastError());
        IXE_CUDA(cudaEventRecord(t1, s_k));
        IXE_CUDA(cudaEventRecord(ev_k[b], s_k));
        ix->launches += 1;
    }
    mark("chunks enqueued");
    IXE_CUDA(cudaStreamSynchronize(s_k));
    IXE_CUDA(cudaGetLastError());
    mark("copy + K1 drained");
    ix->h2d_seconds = now_s() - t_h2d0;          // wall time of the streamed copy + K1 (they overlap)
    for (size_t k = 0; k + 1 < ev_t.size(); k += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ev_t[k], ev_t[k + 1]) == cudaSuccess) ix->kernel_seconds += ms * 1e-3;
    }
    drop_events();
    mark("events read and destroyed");
#undef IXE_CUDA
    // relabelling reads the watchdog flag with its own results (one blocking copy at the end instead of one here)
    const double t_rl0 = now_s();
    g_trace_last = trace ? t_rl0 : 0.0;
    if (relabel_mode == 1) rc = index_relabel_first_occ(ix, d_first_key, arena.p + o_h0, s_k);
    else if (relabel_mode == 2) rc = index_relabel_given(ix, new_of_old, s_k);
    if (rc) return fail(rc);
    ix->relabel_seconds = now_s() - t_rl0;
    g_trace_last = 0.0;
    mark("relabel");
    {
        unsigned int tk[3] = {0, 0, 0};
        IX_CUDA(cudaMemcpy(tk, d_flag, sizeof tk, cudaMemcpyDeviceToHost));
        if (tk[0]) { set_error("gfs_index_build: the scan's look-back watchdog tripped (a tile never published its prefix)"); return fail(GFS_ERR_CUDA); }
        if (tk[2]) { set_error("gfs_index_build: a node is 2^31 bp or longer"); return fail(GFS_ERR_INVALID); }
    }
    mark("watchdog flag read");
    ix->build_seconds = now_s() - t_begin;
    *out = ix;
    return GFS_OK;
#undef IX_CUDA
}

// ---- multi-GPU index (GFASORT_GPUS > 1): one shard per device, built concurrently, one shared relabelling ----------
template <typename HT>
static int build_multi_impl(const HT* step_handles, const uint64_t* path_first_step, const uint32_t* node_len, uint64_t S,
                            uint64_t P, uint64_t N, uint32_t G, int32_t relabel, gfs_index** out) {
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device available — libgfasort_cuda has no CPU fallback"); return GFS_ERR_NO_DEVICE; }
    if ((int)G > ndev) { set_error("GFASORT_GPUS=" + std::to_string(G) + " but only " + std::to_string(ndev) + " CUDA devices are visible"); return GFS_ERR_INVALID; }
    if (G > GFS_P2P_MAX_RANKS) { set_error("GFASORT_GPUS too large"); return GFS_ERR_INVALID; }
    if (!path_first_step || path_first_step[0] != 0 || path_first_step[P] != S) { set_error("gfs_index_build: path_first_step must start at 0 and end at S"); return GFS_ERR_INVALID; }
    const double t0 = now_s();
    gfs_index* mix = new gfs_index();
    mix->device = 0; mix->S = S; mix->P = P; mix->N = N;
    mix->h_first_step.assign(path_first_step, path_first_step + P + 1);
    for (uint64_t p = 0; p < P; ++p) {
        const uint64_t c = path_first_step[p + 1] - path_first_step[p];
        mix->max_path_steps = std::max(mix->max_path_steps, c);
        if (c > 1) mix->any_multi_step = true;
    }
    mix->shards.assign(G, nullptr);
    mix->plans.resize(G);
    for (uint32_t g = 0; g < G; ++g) gfs_shard_plan_make(path_first_step, P, g, G, &mix->plans[g]);
    std::vector<int> rcs(G, GFS_OK);
    std::vector<std::string> errs(G);
    {   // every device pulls its own shard over its own PCIe link; shard 0 also derives the permutation
        std::vector<std::thread> pool;
        for (uint32_t g = 0; g < G; ++g)
            pool.emplace_back([&, g] {
                const gfs_shard_plan& pl = mix->plans[g];
                rcs[g] = build_shard_impl<HT>(step_handles, path_first_step, node_len, S, P, N, pl.path_begin, pl.path_end, (int32_t)g,
                                              (relabel && g == 0) ? 1 : 0, nullptr, &mix->shards[g]);
                if (rcs[g]) errs[g] = g_last_error;
            });
        for (auto& th : pool) th.join();
    }
    for (uint32_t g = 0; g < G; ++g)
        if (rcs[g]) { set_error("shard " + std::to_string(g) + ": " + errs[g]); gfs_index_free(mix); return rcs[g]; }
    if (relabel && N) {     // replicas are reconciled element-wise: all shards use shard 0's node order
        std::vector<uint32_t> perm(N);
        int rc = gfs_index_export_relabel(mix->shards[0], perm.data());
        if (rc) { gfs_index_free(mix); return rc; }
        std::vector<std::thread> pool;
        for (uint32_t g = 1; g < G; ++g)
            pool.emplace_back([&, g] {
                rcs[g] = gfs_index_apply_relabel(mix->shards[g], perm.data());
                if (rcs[g]) errs[g] = g_last_error;
            });
        for (auto& th : pool) th.join();
        for (uint32_t g = 1; g < G; ++g)
            if (rcs[g]) { set_error("shard " + std::to_string(g) + ": " + errs[g]); gfs_index_free(mix); return rcs[g]; }
    }
    for (uint32_t g = 0; g < G; ++g) {
        mix->launches += mix->shards[g]->launches;
        mix->kernel_seconds = std::max(mix->kernel_seconds, mix->shards[g]->kernel_seconds);
        mix->h2d_seconds = std::max(mix->h2d_seconds, mix->shards[g]->h2d_seconds);
        mix->alloc_seconds = std::max(mix->alloc_seconds, mix->shards[g]->alloc_seconds);
        mix->relabel_seconds = std::max(mix->relabel_seconds, mix->shards[g]->relabel_seconds);
    }
    mix->build_seconds = now_s() - t0;
    *out = mix;
    return GFS_OK;
}

template <typename HT>
static int index_build_env(const HT* step_handles, const uint64_t* path_first_step, const uint32_t* node_len, uint64_t S, uint64_t P,
                           uint64_t N, gfs_index** out) {
    if (!out) { set_error("gfs_index_build: out is null"); return GFS_ERR_INVALID; }
    const long gpus = env_long("GFASORT_GPUS", 1);
    const int relabel = env_long("GFASORT_RELABEL", 1) ? 1 : 0;
    if (gpus > 1) return build_multi_impl<HT>(step_handles, path_first_step, node_len, S, P, N, (uint32_t)gpus, relabel, out);
    return build_shard_impl<HT>(step_handles, path_first_step, node_len, S, P, N, 0, P, -1, relabel, nullptr, out);
}

extern "C" int gfs_index_build_shard(const uint64_t* step_handles, const uint64_t* path_first_step,
                                     const uint32_t* node_len, uint64_t S, uint64_t P, uint64_t N,
                                     uint64_t path_begin, uint64_t path_end, int32_t device, int32_t relabel_mode,
                                     const uint32_t* new_of_old, gfs_index** out) {
    return build_shard_impl<uint64_t>(step_handles, path_first_step, node_len, S, P, N, path_begin, path_end, device, relabel_mode, new_of_old, out);
}
extern "C" int gfs_index_build_shard32(const uint32_t* step_handles, const uint64_t* path_first_step,
                                       const uint32_t* node_len, uint64_t S, uint64_t P, uint64_t N,
                                       uint64_t path_begin, uint64_t path_end, int32_t device, int32_t relabel_mode,
                                       const uint32_t* new_of_old, gfs_index** out) {
    return build_shard_impl<uint32_t>(step_handles, path_first_step, node_len, S, P, N, path_begin, path_end, device, relabel_mode, new_of_old, out);
}
extern "C" int gfs_index_build(const uint64_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                               uint64_t S, uint64_t P, uint64_t N, gfs_index** out) {
    return index_build_env<uint64_t>(step_handles, path_first_step, node_len, S, P, N, out);
}
extern "C" int gfs_index_build32(const uint32_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                                 uint64_t S, uint64_t P, uint64_t N, gfs_index** out) {
    return index_build_env<uint32_t>(step_handles, path_first_step, node_len, S, P, N, out);
}

// Page-locked host memory for the flattened step array: the copy engine reads it directly, so a host that flattens
// its paths straight into such a buffer skips the bounce-buffer staging a pageable array needs.
extern "C" int gfs_host_alloc(uint64_t bytes, void** out) {
    if (!out) { set_error("gfs_host_alloc: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    int rc = select_device(-1);
    if (rc) return rc;
    GFS_CUDA(cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable));
    return GFS_OK;
}
extern "C" void gfs_host_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" int gfs_index_apply_relabel(gfs_index* ix, const uint32_t* new_of_old) {
    if (!ix || !new_of_old) { set_error("gfs_index_apply_relabel: null argument"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) { set_error("gfs_index_apply_relabel: not for a multi-GPU index (its shards already share one order)"); return GFS_ERR_INVALID; }
    if (ix->d_new_of_old) { set_error("gfs_index_apply_relabel: the index is already relabelled"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    ScopedStream ss;
    GFS_CUDA(ss.create());
    return index_relabel_given(ix, new_of_old, ss.st);
}

extern "C" int gfs_index_export_relabel(const gfs_index* ix, uint32_t* new_of_old) {
    if (!ix || !new_of_old) { set_error("gfs_index_export_relabel: null argument"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) return gfs_index_export_relabel(ix->shards[0], new_of_old);
    GFS_CUDA(cudaSetDevice(ix->device));
    if (ix->d_new_of_old) GFS_CUDA(cudaMemcpy(new_of_old, ix->d_new_of_old, ix->N * 4, cudaMemcpyDeviceToHost));
    else for (uint64_t i = 0; i < ix->N; ++i) new_of_old[i] = (uint32_t)i;
    return GFS_OK;
}

extern "C" void gfs_index_free(gfs_index* ix) {
    if (!ix) return;
    for (gfs_index* sh : ix->shards) gfs_index_free(sh);
    if (ix->shards.empty()) {
        cudaSetDevice(ix->device);
        cudaFree(ix->d_recs); cudaFree(ix->d_first_step); cudaFree(ix->d_path_len);
        cudaFree(ix->d_new_of_old); cudaFree(ix->d_old_of_new); cudaFree(ix->d_build_arena);
    }
    delete ix;
}

extern "C" int gfs_index_dims(const gfs_index* ix, uint64_t* S, uint64_t* P, uint64_t* N, uint64_t* max_path_steps) {
    if (!ix) { set_error("gfs_index_dims: null index"); return GFS_ERR_INVALID; }
    if (S) *S = ix->S;
    if (P) *P = ix->P;
    if (N) *N = ix->N;
    if (max_path_steps) *max_path_steps = ix->max_path_steps;
    return GFS_OK;
}

extern "C" int gfs_index_build_info(const gfs_index* ix, double* build_seconds, double* copy_seconds, double* kernel_seconds,
                                    double* alloc_seconds, double* relabel_seconds, uint64_t* launches, uint32_t* n_devices) {
    if (!ix) { set_error("gfs_index_build_info: null index"); return GFS_ERR_INVALID; }
    if (build_seconds) *build_seconds = ix->build_seconds;
    if (copy_seconds) *copy_seconds = ix->h2d_seconds;
    if (kernel_seconds) *kernel_seconds = ix->kernel_seconds;
    if (alloc_seconds) *alloc_seconds = ix->alloc_seconds;
    if (relabel_seconds) *relabel_seconds = ix->relabel_seconds;
    if (launches) *launches = ix->launches;
    if (n_devices) *n_devices = ix->shards.empty() ? 1u : (uint32_t)ix->shards.size();
    return GFS_OK;
}

extern "C" int gfs_index_export(const gfs_index* ix, uint64_t* step_pos, uint64_t* path_len) {
    if (!ix) { set_error("gfs_index_export: null index"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) {      // shard g holds paths [path_begin, path_end); a path two shards share has equal values in both
        for (size_t g = 0; g < ix->shards.size(); ++g) {
            const gfs_shard_plan& pl = ix->plans[g];
            int rc = gfs_index_export(ix->shards[g], step_pos ? step_pos + pl.first_step : nullptr, path_len ? path_len + pl.path_begin : nullptr);
            if (rc) return rc;
        }
        return GFS_OK;
    }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (step_pos && ix->S) {
        DevBuf<uint64_t> d;
        const uint64_t CH = 1ull << 26;
        GFS_CUDA(d.alloc(std::min(CH, ix->S)));
        for (uint64_t c0 = 0; c0 < ix->S; c0 += CH) {
            const uint64_t n = std::min(CH, ix->S - c0);
            k1_export_pos<<<(unsigned)((n + 255) / 256), 256>>>(ix->d_recs + c0, n, d.p);
            GFS_CUDA(cudaGetLastError());
            GFS_CUDA(cudaMemcpy(step_pos + c0, d.p, n * 8, cudaMemcpyDeviceToHost));
        }
    }
    if (path_len && ix->P) GFS_CUDA(cudaMemcpy(path_len, ix->d_path_len, ix->P * 8, cudaMemcpyDeviceToHost));
    return GFS_OK;
}

extern "C" int gfs_index_export_records(const gfs_index* ix, uint64_t* step_handle, uint32_t* step_node_len) {
    if (!ix || !step_handle || !step_node_len) { set_error("gfs_index_export_records: null argument"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) {
        for (size_t g = 0; g < ix->shards.size(); ++g) {
            const gfs_shard_plan& pl = ix->plans[g];
            int rc = gfs_index_export_records(ix->shards[g], step_handle + pl.first_step, step_node_len + pl.first_step);
            if (rc) return rc;
        }
        return GFS_OK;
    }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (!ix->S) return GFS_OK;
    DevBuf<uint64_t> dh; DevBuf<uint32_t> dl;
    const uint64_t CH = 1ull << 26;
    const uint64_t cap = std::min(CH, ix->S);
    GFS_CUDA(dh.alloc(cap));
    GFS_CUDA(dl.alloc(cap));
    for (uint64_t c0 = 0; c0 < ix->S; c0 += CH) {
        const uint64_t n = std::min(CH, ix->S - c0);
        k1_export_hl<<<(unsigned)((n + 255) / 256), 256>>>(ix->d_recs + c0, n, (uint32_t)ix->N, ix->d_old_of_new, dh.p, dl.p);
        GFS_CUDA(cudaGetLastError());
        GFS_CUDA(cudaMemcpy(step_handle + c0, dh.p, n * 8, cudaMemcpyDeviceToHost));
        GFS_CUDA(cudaMemcpy(step_node_len + c0, dl.p, n * 4, cudaMemcpyDeviceToHost));
    }
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// SGD sessions
// ---------------------------------------------------------------------------------------------
typedef void (*sgd_kernel_fn)(const SgdArgs);

template <typename CT, bool AGG, int K>
static sgd_kernel_fn pick_nd(uint32_t dims, uint32_t& DS) {
    switch (dims) {
        case 1: DS = 1; return sgd_kernel<CT, 1, 1, AGG, K>;
        case 2: DS = 2; return sgd_kernel<CT, 2, 2, AGG, K>;
        case 3: DS = 4; return sgd_kernel<CT, 3, 4, AGG, K>;
        case 4: DS = 4; return sgd_kernel<CT, 4, 4, AGG, K>;
        case 5: DS = 8; return sgd_kernel<CT, 5, 8, AGG, 1>;      // wide coordinates: one term in flight
        case 6: DS = 8; return sgd_kernel<CT, 6, 8, AGG, 1>;
        case 7: DS = 8; return sgd_kernel<CT, 7, 8, AGG, 1>;
        case 8: DS = 8; return sgd_kernel<CT, 8, 8, AGG, 1>;
        default: return nullptr;
    }
}
template <int K>
static sgd_kernel_fn pick_kernel_k(uint32_t dims, bool f64, bool agg, uint32_t& DS) {
    if (dims == 0) { DS = 1; return agg ? sgd_kernel<double, 0, 1, true, K> : sgd_kernel<double, 0, 1, false, K>; }
    if (f64) return agg ? pick_nd<double, true, K>(dims, DS) : pick_nd<double, false, K>(dims, DS);
    return agg ? pick_nd<float, true, K>(dims, DS) : pick_nd<float, false, K>(dims, DS);
}
static sgd_kernel_fn pick_kernel(uint32_t dims, bool f64, bool agg, int inflight, uint32_t& DS) {
    return inflight >= 2 ? pick_kernel_k<2>(dims, f64, agg, DS) : pick_kernel_k<1>(dims, f64, agg, DS);
}

static KernelGraph make_kgraph(const gfs_index* ix, const gfs_sgd_params& p, const double2* d_zetas, uint32_t zlen) {
    KernelGraph g{};
    g.recs = ix->d_recs; g.first_step = ix->d_first_step; g.zetas = d_zetas;
    g.S = ix->S; g.P = (uint32_t)ix->P; g.N = (uint32_t)ix->N; g.zlen = zlen;
    g.space = (uint32_t)std::min<uint64_t>(p.space, 0xffffffffull);
    g.space_max = (uint32_t)std::min<uint64_t>(p.space_max, 0xffffffffull);
    g.q = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(p.space_quantization_step, 1), 0xffffffffull);
    g.q_is_100 = g.q == 100 ? 1u : 0u;
    g.blk_shift = 0;
    while (((ix->S ? ix->S - 1 : 0) >> g.blk_shift) >= BLK_TABLE) ++g.blk_shift;
    g.samp_base = 0; g.samp_len = ix->S;
    return g;
}

static int validate_params(const gfs_sgd_params* p) {
    if (!p) { set_error("params is null"); return GFS_ERR_INVALID; }
    if (p->iter_max >= (1ull << 31)) { set_error("iter_max too large"); return GFS_ERR_INVALID; }
    if (!(p->theta < 1.0)) { set_error("theta must be < 1"); return GFS_ERR_INVALID; }
    return GFS_OK;
}

extern "C" void gfs_sgd_session_destroy(gfs_sgd_session* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->own_pos) cudaFree(s->d_pos);
    cudaFree(s->d_zetas); cudaFree(s->d_epochs); cudaFree(s->d_attempts); cudaFree(s->d_counters); cudaFree(s->d_stage);
    cudaFree(s->d_work); cudaFree(s->d_saved);
    for (int k = 0; k < gfs_sgd_session::EV_RING; ++k) {
        if (s->ev0[k]) cudaEventDestroy(s->ev0[k]);
        if (s->ev1[k]) cudaEventDestroy(s->ev1[k]);
    }
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int gfs_sgd_session_create(const gfs_index* ix, const gfs_sgd_params* params, uint32_t dims,
                                      const gfs_launch_cfg* cfg, gfs_sgd_session** out) {
    if (!out) { set_error("gfs_sgd_session_create: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    if (!ix) { set_error("gfs_sgd_session_create: null index"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) {
        set_error("gfs_sgd_session_create: this index spans several GPUs (GFASORT_GPUS); use gfs_sgd_1d / gfs_sgd_nd, or one "
                  "gfs_replica per shard");
        return GFS_ERR_INVALID;
    }
    int rc = validate_params(params);
    if (rc) return rc;
    if (dims > 8) { set_error("gfs_sgd_session_create: dims must be <= 8"); return GFS_ERR_INVALID; }
    if (!ix->any_multi_step) { set_error("no paths with multiple steps found"); return GFS_ERR_NO_VALID_PATH; }
    if (ix->N == 0) { set_error("graph has no nodes"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));

    gfs_sgd_session* s = new gfs_sgd_session();
    s->ix = ix; s->params = *params; s->dims = dims; s->device = ix->device;
    int agg = cfg && cfg->aggregate >= 0 ? cfg->aggregate : (int)env_long("GFASORT_AGGREGATE", 1);
    int f64 = cfg && cfg->layout_f64 >= 0 ? cfg->layout_f64 : (int)env_long("GFASORT_LAYOUT_F64", 0);
    s->aggregate = agg != 0;
    s->f64 = dims == 0 ? true : (f64 != 0);
    s->rng_thread_base = cfg ? cfg->rng_thread_base : 0;
    s->samp_base = 0; s->samp_len = ix->S;
    if (cfg && cfg->sample_end > cfg->sample_begin) {
        if (cfg->sample_end > ix->S) { set_error("gfs_sgd_session_create: sample range outside the index"); delete s; return GFS_ERR_INVALID; }
        s->samp_base = cfg->sample_begin; s->samp_len = cfg->sample_end - cfg->sample_begin;
    }
    auto fail = [&](int code) { gfs_sgd_session_destroy(s); return code; };
#define SS_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__)); return fail(GFS_ERR_CUDA); } } while (0)

    if (cfg && cfg->stream) { s->stream = (cudaStream_t)cfg->stream; s->own_stream = false; }
    else { SS_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking)); s->own_stream = true; }
    for (int k = 0; k < gfs_sgd_session::EV_RING; ++k) {
        SS_CUDA(cudaEventCreate(&s->ev0[k]));
        SS_CUDA(cudaEventCreate(&s->ev1[k]));
    }

    // terms in flight per thread: 2 for real runs; 1 (strictly sequential per thread, like one reference
    // worker) when the caller asks for a handful of threads, which is what the bit-exact tests do
    {
        const long want_threads = cfg && cfg->total_threads ? (long)cfg->total_threads : env_long("GFASORT_THREADS", 0);
        s->inflight = (int)env_long("GFASORT_INFLIGHT", (want_threads > 0 && want_threads <= 32) ? 1 : 2);
    }
    sgd_kernel_fn fn = pick_kernel(dims, s->f64, s->aggregate, s->inflight, s->DS);
    if (!fn) return fail(GFS_ERR_INVALID);

    // schedule, zeta table
    std::vector<EpochDesc> epochs; h_epochs(*params, epochs);
    std::vector<double2> zetas; h_zeta_tables(*params, ix->max_path_steps, zetas, s->zlen, ix);
    s->n_epochs = (uint32_t)epochs.size();
    SS_CUDA(cudaMalloc(&s->d_epochs, epochs.size() * sizeof(EpochDesc)));
    SS_CUDA(cudaMalloc(&s->d_zetas, std::max<size_t>(zetas.size(), 1) * sizeof(double2)));
    SS_CUDA(cudaMemcpyAsync(s->d_epochs, epochs.data(), epochs.size() * sizeof(EpochDesc), cudaMemcpyHostToDevice, s->stream));
    SS_CUDA(cudaMemcpyAsync(s->d_zetas, zetas.data(), zetas.size() * sizeof(double2), cudaMemcpyHostToDevice, s->stream));
    SS_CUDA(cudaStreamSynchronize(s->stream));   // host vectors go out of scope

    // launch shape: persistent, every block co-resident
    const uint32_t n_fs = (ix->P + 1 <= SMEM_FS_MAX) ? (uint32_t)ix->P + 1 : 0;
    s->smem_bytes = n_fs ? (size_t)n_fs * 8 + (size_t)BLK_TABLE * 2 : 0;
    s->zeta_smem = (uint32_t)std::min<long>(std::max<long>(env_long("GFASORT_ZETA_SMEM", 0), 0), 6000);      // experiment; default off
    if (s->zeta_smem) s->smem_bytes = ((s->smem_bytes + 15) & ~(size_t)15) + (size_t)2 * std::min<uint32_t>(s->zeta_smem, s->zlen) * sizeof(double2);
    SS_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem_bytes));
    int per_sm = 0, sms = 0;
    SS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, SGD_BLOCK, s->smem_bytes));
    SS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
    if (per_sm < 1) { set_error("SGD kernel does not fit on an SM"); return fail(GFS_ERR_CUDA); }
    const uint64_t max_threads = (uint64_t)per_sm * sms * SGD_BLOCK;
    uint64_t want = cfg && cfg->total_threads ? cfg->total_threads : (uint64_t)env_long("GFASORT_THREADS", 0);
    if (want == 0) {
        // auto: full occupancy, but never more threads than there is work for (>= 8 updates per thread
        // per epoch) and never more terms in flight (3 pipeline stages x K per thread) than half the nodes:
        // beyond that, early epochs (mu = 1) work from positions that are too stale and converge slower
        const uint64_t by_work = std::max<uint64_t>(params->min_term_updates / 8, 32);
        const uint64_t by_nodes = std::max<uint64_t>(ix->N / (2ull * 3ull * (uint64_t)std::max(s->inflight, 1)), 32);
        want = std::min(max_threads, std::min(by_work, by_nodes));
    }
    want = std::min(want, max_threads);
    if (want >= SGD_BLOCK) { s->block = SGD_BLOCK; s->grid = (uint32_t)(want / SGD_BLOCK); }
    else { s->block = (uint32_t)std::max<uint64_t>(want, 1); s->grid = 1; }
    const uint64_t T = (uint64_t)s->grid * s->block;

    s->n_elems = dims == 0 ? ix->N : ix->N * 2 * s->DS;
    const size_t esz = s->f64 ? 8 : 4;
    if (cfg && cfg->device_positions) { s->d_pos = cfg->device_positions; s->own_pos = false; }
    else { SS_CUDA(cudaMalloc(&s->d_pos, s->n_elems * esz)); s->own_pos = true; }
    SS_CUDA(cudaMalloc(&s->d_attempts, T * 8));
    SS_CUDA(cudaMemsetAsync(s->d_attempts, 0, T * 8, s->stream));
    SS_CUDA(cudaMalloc(&s->d_counters, 24));
    SS_CUDA(cudaMemsetAsync(s->d_counters, 0, 24, s->stream));

    // schedule: a sliding sampling window when the step table is much larger than L2
    // (GFASORT_WINDOW: window length in steps, 0 = static schedule with steps ~ U[0,S), -1 = auto)
    {
        long w = cfg && cfg->total_threads == 1 ? 0 : env_long("GFASORT_WINDOW", -1);
        // auto: graphs whose records fit in half of L2 need no window; otherwise 2^20 steps (16 MB of
        // records, several times the number of terms in flight) but never more than 1/8 of the range
        if (w < 0) w = (ix->S * sizeof(StepRec) > (64ull << 20)) ? (long)std::min<uint64_t>(1ull << 20, s->samp_len / 8) : 0;
        // a window is at least 1024 steps; window + one warp of consecutive steps must fit in the sampled range
        // (sample_s1 wraps a step past the end of the range exactly once), else the static schedule is used
        uint64_t ws = w > 0 ? std::max<uint64_t>((uint64_t)w, 1024) : 0;
        if (ws + 32 > s->samp_len) ws = 0;
        s->window_steps = ws;
        s->chunk_updates = (uint32_t)std::max<long>(1, env_long("GFASORT_CHUNK", 256));
        {   // GFASORT_COHERENT: 0 = off, 1 = default group (32 lanes), else the group size rounded down to a power of two <= 32
            long c = env_long("GFASORT_COHERENT", 1);
            uint32_t g = c <= 0 ? 0u : (c == 1 ? 32u : (uint32_t)std::min<long>(c, 32));
            while (g & (g - 1)) g &= g - 1;
            s->coherent = g < 2 ? 0u : g;
        }
        SS_CUDA(cudaMalloc(&s->d_work, 8));
    }

    *out = s;
    return GFS_OK;
#undef SS_CUDA
}

extern "C" int gfs_sgd_session_upload(gfs_sgd_session* s, const double* positions) {
    if (!s || !positions) { set_error("gfs_sgd_session_upload: null argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    const double t0 = now_s();
    const uint64_t N = s->ix->N;
    const uint32_t ends = s->dims == 0 ? 1 : 2, D = s->dims == 0 ? 1 : s->dims;
    const uint32_t* perm = s->ix->d_new_of_old;
    if (!perm && s->f64 && s->DS == D) {
        GFS_CUDA(cudaMemcpyAsync(s->d_pos, positions, N * ends * D * 8, cudaMemcpyHostToDevice, s->stream));
    } else {
        if (!s->d_stage) GFS_CUDA(cudaMalloc(&s->d_stage, N * ends * D * 8));
        GFS_CUDA(cudaMemcpyAsync(s->d_stage, positions, N * ends * D * 8, cudaMemcpyHostToDevice, s->stream));
        const uint64_t n = N * ends * s->DS;
        const unsigned grid = (unsigned)((n + 255) / 256);
        if (s->f64) pos_to_device<double><<<grid, 256, 0, s->stream>>>(s->d_stage, (double*)s->d_pos, N, ends, D, s->DS, perm);
        else pos_to_device<float><<<grid, 256, 0, s->stream>>>(s->d_stage, (float*)s->d_pos, N, ends, D, s->DS, perm);
        GFS_CUDA(cudaGetLastError());
    }
    GFS_CUDA(cudaStreamSynchronize(s->stream));
    s->h2d_s += now_s() - t0;
    return GFS_OK;
}

// reads (blocking) the oldest `n` recorded event pairs into kernel_ms
static int session_read_events(gfs_sgd_session* s, uint32_t n) {
    for (; n > 0 && s->ev_tail != s->ev_head; --n, ++s->ev_tail) {
        const uint32_t k = s->ev_tail % gfs_sgd_session::EV_RING;
        GFS_CUDA(cudaEventSynchronize(s->ev1[k]));
        float ms = 0;
        GFS_CUDA(cudaEventElapsedTime(&ms, s->ev0[k], s->ev1[k]));
        s->kernel_ms += ms;
    }
    return GFS_OK;
}
static int session_flush_events(gfs_sgd_session* s) { return session_read_events(s, s->ev_head - s->ev_tail); }

extern "C" int gfs_sgd_session_download(gfs_sgd_session* s, double* positions) {
    if (!s || !positions) { set_error("gfs_sgd_session_download: null argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    int rc = session_flush_events(s);
    if (rc) return rc;
    const double t0 = now_s();
    const uint64_t N = s->ix->N;
    const uint32_t ends = s->dims == 0 ? 1 : 2, D = s->dims == 0 ? 1 : s->dims;
    const uint32_t* perm = s->ix->d_new_of_old;
    if (!perm && s->f64 && s->DS == D) {
        GFS_CUDA(cudaMemcpyAsync(positions, s->d_pos, N * ends * D * 8, cudaMemcpyDeviceToHost, s->stream));
    } else {
        if (!s->d_stage) GFS_CUDA(cudaMalloc(&s->d_stage, N * ends * D * 8));
        const uint64_t n = N * ends * D;
        const unsigned grid = (unsigned)((n + 255) / 256);
        if (s->f64) pos_from_device<double><<<grid, 256, 0, s->stream>>>((const double*)s->d_pos, s->d_stage, N, ends, D, s->DS, perm);
        else pos_from_device<float><<<grid, 256, 0, s->stream>>>((const float*)s->d_pos, s->d_stage, N, ends, D, s->DS, perm);
        GFS_CUDA(cudaGetLastError());
        GFS_CUDA(cudaMemcpyAsync(positions, s->d_stage, n * 8, cudaMemcpyDeviceToHost, s->stream));
    }
    GFS_CUDA(cudaStreamSynchronize(s->stream));
    s->d2h_s += now_s() - t0;
    return GFS_OK;
}

extern "C" int gfs_sgd_session_run(gfs_sgd_session* s, uint64_t epoch_begin, uint64_t epoch_end, uint32_t slice,
                                   uint32_t n_slices) {
    if (!s) { set_error("gfs_sgd_session_run: null session"); return GFS_ERR_INVALID; }
    if (epoch_begin > epoch_end || epoch_end > s->n_epochs || n_slices == 0 || slice >= n_slices) {
        set_error("gfs_sgd_session_run: bad epoch range or slice"); return GFS_ERR_INVALID;
    }
    if (epoch_begin == epoch_end) return GFS_OK;
    GFS_CUDA(cudaSetDevice(s->device));
    if (s->ev_head - s->ev_tail == gfs_sgd_session::EV_RING) {       // ring full: read the oldest pair (long finished, normally)
        int rc = session_read_events(s, 1);
        if (rc) return rc;
    }
    uint32_t DS;
    sgd_kernel_fn fn = pick_kernel(s->dims, s->f64, s->aggregate, s->inflight, DS);
    SgdArgs a{};
    a.g = make_kgraph(s->ix, s->params, s->d_zetas, s->zlen);
    a.g.samp_base = s->samp_base; a.g.samp_len = s->samp_len;
    a.g.coherent = s->window_steps > 0 ? s->coherent : 0u;
    a.epochs = s->d_epochs;
    a.epoch_begin = (uint32_t)epoch_begin; a.epoch_end = (uint32_t)epoch_end;
    a.slice = slice; a.n_slices = n_slices;
    a.attempt_ctr = s->d_attempts; a.counters = s->d_counters;
    a.seed_lo = (uint32_t)s->params.seed; a.seed_hi = (uint32_t)(s->params.seed >> 32);
    a.tid_base = (uint32_t)s->rng_thread_base;
    a.positions = s->d_pos;
    a.window_steps = s->window_steps; a.chunk_updates = s->chunk_updates;
    a.work_ctr = s->d_work;
    a.zeta_smem = s->zeta_smem;
    {
        // generous bound on loop iterations per warp: 64x its fair share of the launch's attempts + slack
        const uint64_t m = s->params.min_term_updates / n_slices + 1;
        const uint64_t T = (uint64_t)s->grid * s->block;
        a.iter_cap = ((m / T + 1) * (epoch_end - epoch_begin)) * 64 + (1ull << 16);
    }
    GFS_CUDA(cudaMemsetAsync(s->d_work, 0, 8, s->stream));
    void* kargs[] = {(void*)&a};
    const uint32_t ek = s->ev_head % gfs_sgd_session::EV_RING;
    GFS_CUDA(cudaEventRecord(s->ev0[ek], s->stream));
    GFS_CUDA(cudaLaunchCooperativeKernel((const void*)fn, dim3(s->grid), dim3(s->block), kargs, s->smem_bytes, s->stream));
    GFS_CUDA(cudaEventRecord(s->ev1[ek], s->stream));
    s->ev_head += 1;
    s->launches += 1;
    return GFS_OK;
}

// Device-side snapshot / restore of the positions (asynchronous, on the session's stream): lets a
// caller rerun the schedule from the same start without another host->device copy.
extern "C" int gfs_sgd_session_save(gfs_sgd_session* s) {
    if (!s) { set_error("gfs_sgd_session_save: null session"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    const size_t bytes = s->n_elems * (s->f64 ? 8 : 4);
    if (!s->d_saved) GFS_CUDA(cudaMalloc(&s->d_saved, bytes));
    GFS_CUDA(cudaMemcpyAsync(s->d_saved, s->d_pos, bytes, cudaMemcpyDeviceToDevice, s->stream));
    return GFS_OK;
}
extern "C" int gfs_sgd_session_restore(gfs_sgd_session* s) {
    if (!s || !s->d_saved) { set_error("gfs_sgd_session_restore: nothing saved"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    const size_t bytes = s->n_elems * (s->f64 ? 8 : 4);
    GFS_CUDA(cudaMemcpyAsync(s->d_pos, s->d_saved, bytes, cudaMemcpyDeviceToDevice, s->stream));
    return GFS_OK;
}

extern "C" int gfs_sgd_session_sync(gfs_sgd_session* s) {
    if (!s) { set_error("gfs_sgd_session_sync: null session"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    GFS_CUDA(cudaStreamSynchronize(s->stream));
    return session_flush_events(s);
}

extern "C" int gfs_sgd_session_positions(gfs_sgd_session* s, void** dev_ptr, uint64_t* n_elems, uint32_t* elem_bytes) {
    if (!s) { set_error("gfs_sgd_session_positions: null session"); return GFS_ERR_INVALID; }
    if (dev_ptr) *dev_ptr = s->d_pos;
    if (n_elems) *n_elems = s->n_elems;
    if (elem_bytes) *elem_bytes = s->f64 ? 8 : 4;
    return GFS_OK;
}

extern "C" int gfs_sgd_session_stats(gfs_sgd_session* s, gfs_stats* st) {
    if (!s || !st) { set_error("gfs_sgd_session_stats: null argument"); return GFS_ERR_INVALID; }
    int rc = gfs_sgd_session_sync(s);
    if (rc) return rc;
    unsigned long long c[3] = {0, 0, 0};
    GFS_CUDA(cudaMemcpy(c, s->d_counters, 24, cudaMemcpyDeviceToHost));
    if (c[2] != 0) {
        set_error("SGD kernel watchdog tripped: a warp exceeded its iteration bound (sampling cannot find valid terms?)");
        return GFS_ERR_CUDA;
    }
    std::memset(st, 0, sizeof *st);
    st->applied_updates = c[0]; st->attempts = c[1];
    st->epochs = s->n_epochs; st->launches = s->launches;
    st->kernel_seconds = s->kernel_ms * 1e-3;
    st->h2d_seconds = s->h2d_s; st->d2h_seconds = s->d2h_s;
    st->grid = s->grid; st->block = s->block; st->coord_bytes = s->f64 ? 8 : 4;
    st->n_devices = 1; st->window_steps = s->window_steps; st->coherent = s->window_steps > 0 ? s->coherent : 0u;
    st->syncs_per_epoch = 0;
    return GFS_OK;
}

static int run_whole(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg, uint32_t dims,
                     double* pos_inout, gfs_stats* stats) {
    if (!pos_inout) { set_error("positions buffer is null"); return GFS_ERR_INVALID; }
    if (ix && !ix->shards.empty()) return gfs_multi_run_whole(ix, params, cfg, dims, pos_inout, stats);    // GFASORT_GPUS > 1
    const double t0 = now_s();
    gfs_sgd_session* s = nullptr;
    int rc = gfs_sgd_session_create(ix, params, dims, cfg, &s);
    if (rc) return rc;   // GFS_ERR_NO_VALID_PATH: positions untouched (sgd.rs:258-261)
    rc = gfs_sgd_session_upload(s, pos_inout);
    if (!rc) rc = gfs_sgd_session_run(s, 0, s->n_epochs, 0, 1);
    if (!rc) rc = gfs_sgd_session_sync(s);
    if (!rc) rc = gfs_sgd_session_download(s, pos_inout);
    gfs_stats st{};
    if (!rc) rc = gfs_sgd_session_stats(s, &st);
    st.total_seconds = now_s() - t0;
    if (stats && !rc) *stats = st;
    gfs_sgd_session_destroy(s);
    return rc;
}

extern "C" int gfs_sgd_1d_cfg(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg,
                              double* x_inout, gfs_stats* stats) {
    return run_whole(ix, params, cfg, 0, x_inout, stats);
}
extern "C" int gfs_sgd_1d(const gfs_index* ix, const gfs_sgd_params* params, double* x_inout, gfs_stats* stats) {
    return run_whole(ix, params, nullptr, 0, x_inout, stats);
}
extern "C" int gfs_sgd_nd_cfg(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg,
                              uint32_t dims, double* coords_inout, gfs_stats* stats) {
    if (dims < 1 || dims > 8) { set_error("gfs_sgd_nd: dims must be in 1..8"); return GFS_ERR_INVALID; }
    return run_whole(ix, params, cfg, dims, coords_inout, stats);
}
extern "C" int gfs_sgd_nd(const gfs_index* ix, const gfs_sgd_params* params, uint32_t dims, double* coords_inout,
                          gfs_stats* stats) {
    return gfs_sgd_nd_cfg(ix, params, nullptr, dims, coords_inout, stats);
}

// ---------------------------------------------------------------------------------------------
// stress
// ---------------------------------------------------------------------------------------------
// One index's share of the sampled stress: sample k draws a step s ~ U[0, total_steps) of the WHOLE graph; this index
// (whose local step 0 is global step `step_offset`) evaluates the samples with s in [step_begin, step_end) and adds
// (sum of squared relative errors, sum of absolute relative errors, count) to sums[3].  A path's steps must all lie
// in this index, which holds for a shard and the step slice it samples.
static int stress_partial(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords, uint64_t samples,
                          uint64_t seed, uint64_t total_steps, uint64_t step_offset, uint64_t step_begin, uint64_t step_end,
                          double* sums) {
    GFS_CUDA(cudaSetDevice(ix->device));
    if (total_steps < 2 || samples == 0 || step_end <= step_begin) return GFS_OK;            // sgd.rs:1220-1222
    const uint32_t stride = layout_order ? 2 * dims : dims;
    DevBuf<double> d_coords, d_partial;
    const size_t n_coords = (size_t)ix->N * stride;
    GFS_CUDA(d_coords.alloc(n_coords));
    GFS_CUDA(d_coords.up(coords, n_coords));
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ix->device);
    const unsigned grid = (unsigned)std::min<uint64_t>((samples + STRESS_BLOCK - 1) / STRESS_BLOCK, (uint64_t)sms * 8);
    GFS_CUDA(d_partial.alloc((size_t)grid * 3));
    gfs_sgd_params dummy{};
    KernelGraph g = make_kgraph(ix, dummy, nullptr, 0);
    stress_kernel<<<grid, STRESS_BLOCK>>>(g, ix->d_old_of_new, d_coords.p, dims, stride, samples, (uint32_t)seed, (uint32_t)(seed >> 32),
                                          total_steps, step_offset, step_begin, step_end, d_partial.p);
    GFS_CUDA(cudaGetLastError());
    std::vector<double> part((size_t)grid * 3);
    GFS_CUDA(d_partial.down(part.data(), part.size()));
    for (unsigned b = 0; b < grid; ++b) { sums[0] += part[b * 3]; sums[1] += part[b * 3 + 1]; sums[2] += part[b * 3 + 2]; }
    return GFS_OK;
}

extern "C" int gfs_stress_partial(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords, uint64_t samples,
                                  uint64_t seed, uint64_t total_steps, uint64_t step_offset, uint64_t step_begin, uint64_t step_end,
                                  double* sums3) {
    if (!ix || !coords || dims < 1 || !sums3) { set_error("gfs_stress_partial: bad argument"); return GFS_ERR_INVALID; }
    if (!ix->shards.empty()) { set_error("gfs_stress_partial: call it on one shard (gfs_stress handles a multi-GPU index)"); return GFS_ERR_INVALID; }
    if (step_begin < step_offset || step_end > step_offset + ix->S) { set_error("gfs_stress_partial: step range outside the index"); return GFS_ERR_INVALID; }
    return stress_partial(ix, dims, layout_order, coords, samples, seed, total_steps, step_offset, step_begin, step_end, sums3);
}

extern "C" int gfs_stress(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords,
                          uint64_t samples, uint64_t seed, double* rms_rel, double* mean_abs_rel, uint64_t* counted) {
    if (!ix || !coords || dims < 1) { set_error("gfs_stress: bad argument"); return GFS_ERR_INVALID; }
    if (rms_rel) *rms_rel = 0;
    if (mean_abs_rel) *mean_abs_rel = 0;
    if (counted) *counted = 0;
    double sums[3] = {0, 0, 0};
    if (!ix->shards.empty()) {      // every shard evaluates the samples that fall into its step slice: all paths, same sample as one GPU
        for (size_t g = 0; g < ix->shards.size(); ++g) {
            const gfs_shard_plan& pl = ix->plans[g];
            int rc = stress_partial(ix->shards[g], dims, layout_order, coords, samples, seed, ix->S, pl.first_step, pl.sample_begin,
                                    pl.sample_end, sums);
            if (rc) return rc;
        }
    } else {
        int rc = stress_partial(ix, dims, layout_order, coords, samples, seed, ix->S, 0, 0, ix->S, sums);
        if (rc) return rc;
    }
    if (sums[2] > 0) {
        if (rms_rel) *rms_rel = std::sqrt(sums[0] / sums[2]);
        if (mean_abs_rel) *mean_abs_rel = sums[1] / sums[2];
    }
    if (counted) *counted = (uint64_t)sums[2];
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// debug / parity hooks
// ---------------------------------------------------------------------------------------------
extern "C" int gfs_debug_fast_precise_pow(const double* a, const double* b, double* out, uint64_t n) {
    int rc = select_device(-1); if (rc) return rc;
    DevBuf<double> da, db, dout;
    GFS_CUDA(da.alloc(n)); GFS_CUDA(db.alloc(n)); GFS_CUDA(dout.alloc(n));
    GFS_CUDA(da.up(a, n)); GFS_CUDA(db.up(b, n));
    dbg_fpp<<<(unsigned)((n + 255) / 256), 256>>>(da.p, db.p, dout.p, n);
    GFS_CUDA(cudaGetLastError());
    GFS_CUDA(dout.down(out, n));
    return GFS_OK;
}
extern "C" int gfs_debug_dirty_zipf(const uint64_t* zmax, const double* theta, const double* zeta, const double* u,
                                    uint64_t* out, uint64_t n) {
    int rc = select_device(-1); if (rc) return rc;
    DevBuf<uint64_t> dz, dout; DevBuf<double> dt, dze, du;
    GFS_CUDA(dz.alloc(n)); GFS_CUDA(dout.alloc(n)); GFS_CUDA(dt.alloc(n)); GFS_CUDA(dze.alloc(n)); GFS_CUDA(du.alloc(n));
    GFS_CUDA(dz.up(zmax, n)); GFS_CUDA(dt.up(theta, n)); GFS_CUDA(dze.up(zeta, n)); GFS_CUDA(du.up(u, n));
    dbg_zipf<<<(unsigned)((n + 255) / 256), 256>>>(dz.p, dt.p, dze.p, du.p, dout.p, n);
    GFS_CUDA(cudaGetLastError());
    GFS_CUDA(dout.down(out, n));
    return GFS_OK;
}
extern "C" int gfs_debug_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4, uint64_t n) {
    int rc = select_device(-1); if (rc) return rc;
    DevBuf<uint32_t> dc, dk, dout;
    GFS_CUDA(dc.alloc(4 * n)); GFS_CUDA(dk.alloc(2 * n)); GFS_CUDA(dout.alloc(4 * n));
    GFS_CUDA(dc.up(ctr4, 4 * n)); GFS_CUDA(dk.up(key2, 2 * n));
    dbg_philox<<<(unsigned)((n + 255) / 256), 256>>>(dc.p, dk.p, dout.p, n);
    GFS_CUDA(cudaGetLastError());
    GFS_CUDA(dout.down(out4, 4 * n));
    return GFS_OK;
}
extern "C" int gfs_debug_schedule(const gfs_sgd_params* params, double* etas) {
    int rc = validate_params(params); if (rc) return rc;
    std::vector<double> e; h_schedule(*params, e);
    std::memcpy(etas, e.data(), e.size() * 8);
    return GFS_OK;
}
// Host-only twin of gfs_debug_zetas (no index, no device): the table for a graph whose longest path has max_path_steps steps.
extern "C" int gfs_debug_zetas_host(const gfs_sgd_params* params, uint64_t max_path_steps, double* zetas, uint64_t cap, uint64_t* n) {
    int rc = validate_params(params); if (rc) return rc;
    std::vector<double> z; h_zetas(*params, max_path_steps, z);
    if (n) *n = z.size();
    if (zetas) std::memcpy(zetas, z.data(), std::min<uint64_t>(cap, z.size()) * 8);
    return GFS_OK;
}
extern "C" int gfs_debug_zetas(const gfs_index* ix, const gfs_sgd_params* params, double* zetas, uint64_t cap, uint64_t* n) {
    int rc = validate_params(params); if (rc) return rc;
    if (!ix) { set_error("gfs_debug_zetas: null index"); return GFS_ERR_INVALID; }
    std::vector<double> z; h_zetas(*params, ix->max_path_steps, z);
    if (n) *n = z.size();
    if (zetas) std::memcpy(zetas, z.data(), std::min<uint64_t>(cap, z.size()) * 8);
    return GFS_OK;
}
extern "C" int gfs_debug_trace_terms(const gfs_index* ix, const gfs_sgd_params* params, int32_t nd, uint64_t epoch,
                                     uint32_t tid, uint64_t attempt0, uint64_t count, uint8_t* valid, uint64_t* step_a,
                                     uint64_t* step_b, uint8_t* flags, double* dist) {
    int rc = validate_params(params); if (rc) return rc;
    if (!ix || epoch > params->iter_max) { set_error("gfs_debug_trace_terms: bad argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    std::vector<EpochDesc> epochs; h_epochs(*params, epochs);
    std::vector<double2> zetas; uint32_t zlen = 0; h_zeta_tables(*params, ix->max_path_steps, zetas, zlen);
    DevBuf<EpochDesc> de; DevBuf<double2> dz; DevBuf<double> dd; DevBuf<uint8_t> dv, df; DevBuf<uint64_t> da, db;
    GFS_CUDA(de.alloc(epochs.size())); GFS_CUDA(dz.alloc(zetas.size()));
    GFS_CUDA(de.up(epochs.data(), epochs.size())); GFS_CUDA(dz.up(zetas.data(), zetas.size()));
    GFS_CUDA(dv.alloc(count)); GFS_CUDA(df.alloc(count)); GFS_CUDA(da.alloc(count)); GFS_CUDA(db.alloc(count)); GFS_CUDA(dd.alloc(count));
    KernelGraph g = make_kgraph(ix, *params, dz.p, zlen);
    const unsigned grid = (unsigned)((count + 255) / 256);
    if (nd) dbg_trace<true><<<grid, 256>>>(g, de.p, (uint32_t)epoch, (uint32_t)params->seed, (uint32_t)(params->seed >> 32), tid, attempt0, count, dv.p, da.p, db.p, df.p, dd.p);
    else dbg_trace<false><<<grid, 256>>>(g, de.p, (uint32_t)epoch, (uint32_t)params->seed, (uint32_t)(params->seed >> 32), tid, attempt0, count, dv.p, da.p, db.p, df.p, dd.p);
    GFS_CUDA(cudaGetLastError());
    GFS_CUDA(dv.down(valid, count)); GFS_CUDA(da.down(step_a, count)); GFS_CUDA(db.down(step_b, count));
    GFS_CUDA(df.down(flags, count)); GFS_CUDA(dd.down(dist, count));
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// replica reconcile helpers
// ---------------------------------------------------------------------------------------------
// the device that owns `ptr` becomes current; returns its SM count (the helpers take raw device pointers and a
// caller stream, so nothing else says where to launch)
static int enter_device_of(const void* ptr, int* sms) {
    cudaPointerAttributes attr{};
    GFS_CUDA(cudaPointerGetAttributes(&attr, ptr));
    if (attr.type != cudaMemoryTypeDevice) { set_error("reconcile: not a device pointer"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(attr.device));
    GFS_CUDA(cudaDeviceGetAttribute(sms, cudaDevAttrMultiProcessorCount, attr.device));
    return GFS_OK;
}
extern "C" int gfs_reconcile_pack(const void* x, const void* x_sync, uint64_t n, uint32_t elem_bytes, float* buf, void* stream) {
    if (!x || !x_sync || !buf || (elem_bytes != 4 && elem_bytes != 8)) { set_error("gfs_reconcile_pack: bad argument"); return GFS_ERR_INVALID; }
    if (n == 0) return GFS_OK;
    int sms = 0, rc = enter_device_of(x, &sms);
    if (rc) return rc;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 16);
    if (elem_bytes == 8) rc_pack<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, (const double*)x_sync, n, buf);
    else rc_pack<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)x_sync, n, buf);
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}
extern "C" int gfs_reconcile_apply(void* x, void* x_sync, uint64_t n, uint32_t elem_bytes, const float* buf, void* stream) {
    if (!x || !x_sync || !buf || (elem_bytes != 4 && elem_bytes != 8)) { set_error("gfs_reconcile_apply: bad argument"); return GFS_ERR_INVALID; }
    if (n == 0) return GFS_OK;
    int sms = 0, rc = enter_device_of(x, &sms);
    if (rc) return rc;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)sms * 16);
    if (elem_bytes == 8) rc_apply<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)x, (double*)x_sync, n, buf);
    else rc_apply<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, (float*)x_sync, n, buf);
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// order by position
// ---------------------------------------------------------------------------------------------
// Stable LSD radix sort of n (key, value) pairs over the low 8*passes key bits (rs_* kernels).  The buffers are
// ping-ponged: after the call (k0, v0) name the sorted arrays (with an even number of passes, the ones passed in).
static int radix_sort_pairs(uint64_t*& k0, uint64_t*& k1, uint32_t*& v0, uint32_t*& v1, uint32_t* hist, uint64_t n, int passes,
                            cudaStream_t st, uint64_t* launches) {
    const uint32_t n_blocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    for (int pass = 0; pass < passes; ++pass) {
        rs_hist<<<n_blocks, RS_THREADS, 0, st>>>(k0, n, pass * 8, n_blocks, hist);
        rs_scan<<<1, 1024, 0, st>>>(hist, (uint64_t)256 * n_blocks);
        rs_scatter<<<n_blocks, RS_THREADS, 0, st>>>(k0, v0, n, pass * 8, n_blocks, hist, k1, v1);
        std::swap(k0, k1); std::swap(v0, v1);
        if (launches) *launches += 3;
    }
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}

// d_x: n positions on the device, in the caller's node order.  d_order: n dense indices.  Asynchronous on st.
static int sort_positions_device(const double* d_x, uint64_t n, uint32_t* d_order, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return GFS_OK;
    if (n >= (1ull << 32)) { set_error("sort: n must be < 2^32"); return GFS_ERR_INVALID; }
    const uint32_t n_blocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    DevBuf<uint64_t> b_k0, b_k1; DevBuf<uint32_t> b_v1, b_hist;
    GFS_CUDA(b_k0.alloc(n)); GFS_CUDA(b_k1.alloc(n));
    GFS_CUDA(b_v1.alloc(n)); GFS_CUDA(b_hist.alloc((size_t)256 * n_blocks));
    uint64_t *k0 = b_k0.p, *k1 = b_k1.p; uint32_t *v1 = b_v1.p;
    uint32_t* v0 = d_order;
    rs_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_x, n, k0, v0);
    if (launches) *launches += 1;
    int rc = radix_sort_pairs(k0, k1, v0, v1, b_hist.p, n, 8, st, launches);    // 8 passes: the result is back in d_order
    if (rc) return rc;
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("sort failed: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
    return GFS_OK;
}

extern "C" int gfs_sort_positions(const double* x, uint64_t n, uint32_t* order) {
    if ((n && !x) || (n && !order)) { set_error("gfs_sort_positions: null argument"); return GFS_ERR_INVALID; }
    int rc = select_device(-1); if (rc) return rc;
    if (n == 0) return GFS_OK;
    DevBuf<double> dx; DevBuf<uint32_t> dord;
    GFS_CUDA(dx.alloc(n)); GFS_CUDA(dord.alloc(n));
    GFS_CUDA(dx.up(x, n));
    rc = sort_positions_device(dx.p, n, dord.p, nullptr, nullptr);
    if (rc) return rc;
    GFS_CUDA(dord.down(order, n));
    return GFS_OK;
}

extern "C" int gfs_sgd_session_sort(gfs_sgd_session* s, uint32_t* order) {
    if (!s || !order) { set_error("gfs_sgd_session_sort: null argument"); return GFS_ERR_INVALID; }
    if (s->dims != 0) { set_error("gfs_sgd_session_sort: 1D sessions only"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(s->device));
    int rc = session_flush_events(s);
    if (rc) return rc;
    const uint64_t N = s->ix->N;
    // positions in the caller's node order (undo the internal relabelling), then sort
    if (!s->d_stage) GFS_CUDA(cudaMalloc(&s->d_stage, N * 8));
    pos_from_device<double><<<(unsigned)((N + 255) / 256), 256, 0, s->stream>>>((const double*)s->d_pos, s->d_stage, N, 1, 1, 1, s->ix->d_new_of_old);
    GFS_CUDA(cudaGetLastError());
    DevBuf<uint32_t> dord;
    GFS_CUDA(dord.alloc(N));
    rc = sort_positions_device(s->d_stage, N, dord.p, s->stream, &s->launches);
    if (rc) return rc;
    const double t0 = now_s();
    GFS_CUDA(dord.down(order, N));
    s->d2h_s += now_s() - t0;
    return GFS_OK;
}

extern "C" int gfs_sgd_sort_1d(const gfs_index* ix, const gfs_sgd_params* params, double* x_inout, uint32_t* order_out,
                               gfs_stats* stats) {
    if (!x_inout || !order_out) { set_error("gfs_sgd_sort_1d: null buffer"); return GFS_ERR_INVALID; }
    if (ix && !ix->shards.empty()) {                 // multi-GPU index: the replicated run, then the sort on one device
        int rc = gfs_sgd_1d(ix, params, x_inout, stats);
        if (!rc) rc = gfs_sort_positions(x_inout, ix->N, order_out);
        return rc;
    }
    const double t0 = now_s();
    gfs_sgd_session* s = nullptr;
    int rc = gfs_sgd_session_create(ix, params, 0, nullptr, &s);
    if (rc) return rc;
    rc = gfs_sgd_session_upload(s, x_inout);
    if (!rc) rc = gfs_sgd_session_run(s, 0, s->n_epochs, 0, 1);
    if (!rc) rc = gfs_sgd_session_sync(s);
    if (!rc) rc = gfs_sgd_session_sort(s, order_out);
    if (!rc) rc = gfs_sgd_session_download(s, x_inout);
    gfs_stats st{};
    if (!rc) rc = gfs_sgd_session_stats(s, &st);
    st.total_seconds = now_s() - t0;
    if (stats && !rc) *stats = st;
    gfs_sgd_session_destroy(s);
    return rc;
}
