// gfs_lib.cu — libgfasort_cuda.so: kernels + C ABI (include/gfasort_cuda.h) for the path-guided SGD
// hot path of pangenome/gfasort on B200 (sm_100a).
//
// Kernels
//   K1  k1_tile_sums / k1_scan_tiles / k1_path_base / k1_write_recs
//         path index: gather node lengths, scan them, emit one 16-byte StepRec per step.
//         Replaces PathIndex::from_graph (reference src/sgd.rs:34-71).
//   K2  sgd_kernel<ONE_D>   persistent 1D `Y` term loop, f64 positions.  Replaces the worker loop
//         src/sgd.rs:442-584 and the checker thread :366-407.
//   K3  sgd_kernel<ND>      persistent nD `L` term loop on [node][end][dim] coordinates (float or
//         double).  Replaces src/sgd.rs:988-1156 and :925-955.
//   K4  stress_kernel       sampled path stress.  Replaces calculate_layout_stress (:1196-1283).
//
// There is no CPU implementation of any of these in this library: if CUDA is unavailable every
// entry point returns an error.
#include <cuda_runtime.h>
#include <cooperative_groups.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gfasort_cuda.h"
#include "gfs_device.cuh"

namespace gfs {

static thread_local std::string g_last_error;
void set_error(const std::string& s) { g_last_error = s; }

#define GFS_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess) {                                                                   \
            set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" __FILE__ ":" + \
                      std::to_string(__LINE__) + ")");                                              \
            return GFS_ERR_CUDA;                                                                    \
        }                                                                                           \
    } while (0)

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
static long env_long(const char* name, long dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    return std::strtol(v, nullptr, 10);
}

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
static void h_zetas(const gfs_sgd_params& p, uint64_t max_path_steps, std::vector<double>& z) {
    const uint64_t sm = p.space_max, q = p.space_quantization_step ? p.space_quantization_step : 1;
    const uint64_t full = ((p.space <= sm) ? p.space : sm + (p.space - sm) / q + 1) + 1;
    const uint64_t reach = std::min(p.space, max_path_steps);
    const uint64_t reach_idx = reach > sm ? sm + (reach - sm) / q + 1 : reach;
    const uint64_t n = std::min(full, reach_idx + 2);
    z.assign(n, 0.0);
    double acc = 0.0;
    for (uint64_t i = 1; i <= reach; ++i) {
        acc += h_fast_precise_pow(1.0 / (double)i, p.theta);
        if (i <= sm) { if (i < n) z[i] = acc; }
        if (i >= sm && (i - sm) % q == 0) {
            uint64_t idx = sm + 1 + (i - sm) / q;
            if (idx < n) z[idx] = acc;
        }
    }
}

// One epoch of the schedule as the kernels see it.
struct EpochDesc {
    double eta;
    ZipfConsts zc;
    uint64_t updates;     // min_term_updates
    uint32_t cooling;
    uint32_t pad;
};

static void h_epochs(const gfs_sgd_params& p, std::vector<EpochDesc>& out) {
    std::vector<double> etas;
    h_schedule(p, etas);
    const uint64_t first_cooling = (uint64_t)std::floor(p.cooling_start * (double)p.iter_max);   // sgd.rs:297
    out.resize(p.iter_max + 1);
    for (uint64_t e = 0; e <= p.iter_max; ++e) {
        EpochDesc d{};
        d.eta = etas[e];
        d.cooling = e > first_cooling ? 1u : 0u;                    // strict (sgd.rs:393)
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

// =============================================================================================
// K1 — path index
// =============================================================================================
constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 8;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;

__device__ __forceinline__ uint32_t gathered_len(uint64_t h, const uint32_t* __restrict__ node_len, uint64_t N) {
    const uint64_t node = h >> 1;
    return node < N ? __ldg(node_len + node) : 0u;    // missing node => +0 (src/sgd.rs:52-54)
}

__device__ __forceinline__ uint64_t block_sum_u64(uint64_t v, uint64_t* warp_buf) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) warp_buf[w] = v;
    __syncthreads();
    uint64_t t = 0;
    if (threadIdx.x < 32) {
        t = threadIdx.x < (blockDim.x >> 5) ? warp_buf[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) warp_buf[0] = t;
    }
    __syncthreads();
    t = warp_buf[0];
    __syncthreads();
    return t;
}

// tile_sum[t] = sum of node lengths of the steps of tile t
__global__ void __launch_bounds__(K1_THREADS)
k1_tile_sums(const uint64_t* __restrict__ handles, const uint32_t* __restrict__ node_len, uint64_t S, uint64_t N,
             uint64_t* __restrict__ tile_sum) {
    __shared__ uint64_t wb[32];
    const uint64_t base = (uint64_t)blockIdx.x * K1_TILE;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
        if (i < S) s += gathered_len(handles[i], node_len, N);
    }
    s = block_sum_u64(s, wb);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = s;
}

// exclusive scan of tile sums, seeded with *carry; leaves the running total in *carry. One block.
__global__ void __launch_bounds__(1024)
k1_scan_tiles(uint64_t* __restrict__ tile_sum, uint64_t n_tiles, uint64_t* __restrict__ carry) {
    __shared__ uint64_t wsum[32];
    __shared__ uint64_t running;
    if (threadIdx.x == 0) running = *carry;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t base = 0; base < n_tiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sum[i] : 0;
        uint64_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            uint64_t ws = wsum[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint64_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            wsum[lane] = wi - ws;   // exclusive warp offsets
        }
        __syncthreads();
        const uint64_t excl = running + wsum[w] + (inc - v);
        if (i < n_tiles) tile_sum[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) running = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *carry = running;
}

// path_base[p] = global exclusive prefix at the first step of path p, for the paths that start
// inside [chunk_begin, chunk_end).  One block per path of the chunk's path range.
__global__ void __launch_bounds__(K1_THREADS)
k1_path_base(const uint64_t* __restrict__ handles /*chunk-local*/, const uint32_t* __restrict__ node_len,
             uint64_t N, const uint64_t* __restrict__ first_step, uint32_t p_begin, uint32_t p_end,
             uint64_t chunk_begin, uint64_t chunk_end, const uint64_t* __restrict__ tile_prefix /*chunk-local*/,
             uint64_t* __restrict__ path_base) {
    __shared__ uint64_t wb[32];
    const uint32_t p = p_begin + blockIdx.x;
    if (p >= p_end) return;
    const uint64_t s0 = first_step[p];
    if (s0 < chunk_begin || s0 >= chunk_end) return;   // block-uniform
    const uint64_t local = s0 - chunk_begin;
    const uint64_t tile = local / K1_TILE;
    const uint64_t tbase = tile * K1_TILE;
    uint64_t s = 0;
    for (uint64_t i = tbase + threadIdx.x; i < local; i += K1_THREADS) s += gathered_len(handles[i], node_len, N);
    s = block_sum_u64(s, wb);
    if (threadIdx.x == 0) path_base[p] = tile_prefix[tile] + s;
}

// Emit the records of one tile: pos = global prefix - path_base[path(step)].
__global__ void __launch_bounds__(K1_THREADS)
k1_write_recs(const uint64_t* __restrict__ handles /*chunk-local*/, const uint32_t* __restrict__ node_len,
              uint64_t N, const uint64_t* __restrict__ first_step, uint32_t P, uint64_t chunk_begin,
              uint64_t chunk_len, const uint64_t* __restrict__ tile_prefix, const uint64_t* __restrict__ path_base,
              StepRec* __restrict__ recs /*global index*/) {
    __shared__ uint32_t s_len[K1_TILE];
    __shared__ uint32_t s_nr[K1_TILE];
    __shared__ uint64_t s_pos[K1_TILE];
    __shared__ uint64_t wsum[K1_THREADS / 32];
    const uint64_t tbase = (uint64_t)blockIdx.x * K1_TILE;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const int j = k * K1_THREADS + threadIdx.x;
        const uint64_t i = tbase + j;
        uint32_t len = 0, nr = 0;
        if (i < chunk_len) {
            const uint64_t h = handles[i];
            len = gathered_len(h, node_len, N);
            const uint64_t node = h >> 1;
            nr = (uint32_t)(((node < N ? node : N) << 1) | (h & 1));
        }
        s_len[j] = len; s_nr[j] = nr;
    }
    __syncthreads();
    // thread t owns items [t*8, t*8+8)
    uint64_t loc[K1_ITEMS];
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) { loc[k] = tsum; tsum += s_len[threadIdx.x * K1_ITEMS + k]; }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint64_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint64_t woff = 0;
#pragma unroll
    for (int k = 0; k < K1_THREADS / 32; ++k) woff += (k < w) ? wsum[k] : 0;
    const uint64_t texcl = tile_prefix[blockIdx.x] + woff + (inc - tsum);
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) s_pos[threadIdx.x * K1_ITEMS + k] = texcl + loc[k];
    __syncthreads();
    // path of the tile's first and last step; most tiles lie inside one path
    const uint64_t g_first = chunk_begin + tbase;
    const uint64_t last_local = (tbase + K1_TILE <= chunk_len ? tbase + K1_TILE : chunk_len) - 1;
    const uint32_t p_first = find_path(first_step, P, g_first);
    const uint32_t p_last = find_path(first_step, P, chunk_begin + last_local);
    const uint64_t base_first = path_base[p_first];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const int j = k * K1_THREADS + threadIdx.x;
        const uint64_t i = tbase + j;
        if (i < chunk_len) {
            const uint64_t gi = chunk_begin + i;
            uint64_t pb = base_first;
            if (p_first != p_last) pb = path_base[find_path(first_step, P, gi)];
            StepRec r;
            r.node_rev = s_nr[j]; r.node_len = s_len[j]; r.pos = s_pos[j] - pb;
            recs[gi] = r;
        }
    }
}

__global__ void k1_path_len(const uint64_t* __restrict__ path_base, uint32_t P, uint64_t* __restrict__ path_len) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) path_len[p] = path_base[p + 1] - path_base[p];
}
__global__ void k1_set_u64(uint64_t* p, uint64_t idx, const uint64_t* src) { p[idx] = *src; }

__global__ void k1_export_pos(const StepRec* __restrict__ recs, uint64_t S, uint64_t* __restrict__ pos) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < S) pos[i] = recs[i].pos;
}
__global__ void k1_export_hl(const StepRec* __restrict__ recs, uint64_t S, uint32_t N, const uint32_t* __restrict__ old_of_new,
                             uint64_t* __restrict__ h, uint32_t* __restrict__ l) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const uint32_t nr = recs[i].node_rev;
    uint32_t node = nr >> 1;
    if (old_of_new && node < N) node = old_of_new[node];
    h[i] = ((uint64_t)node << 1) | (nr & 1u);
    l[i] = recs[i].node_len;
}

// ---------------------------------------------------------------------------------------------
// node relabelling: internal node index = order of first appearance along the paths, so that the
// positions of path-adjacent nodes share cache lines (the host's dense idx order is the GFA file
// order, which says nothing about adjacency).  Purely a storage permutation: uploads scatter through
// new_of_old, downloads gather back; no arithmetic changes.
// ---------------------------------------------------------------------------------------------
__global__ void rl_first_occ(const StepRec* __restrict__ recs, uint64_t S, uint32_t N, unsigned long long* __restrict__ first_occ) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const uint32_t node = recs[i].node_rev >> 1;
    if (node < N && first_occ[node] > i) atomicMin(first_occ + node, (unsigned long long)i);
}
// exclusive scan of one flag per thread-item across the block; returns the block total in *total
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t* wsum, uint32_t* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t off = 0, tot = 0;
    for (int k = 0; k < nw; ++k) { const uint32_t x = wsum[k]; if (k < w) off += x; tot += x; }
    __syncthreads();
    *total = tot;
    return off + inc - v;
}
// mode 0: item i is a step, flag = "first occurrence of its node"; mode 1: item i is a node, flag = "never visited"
template <int MODE>
__device__ __forceinline__ bool rl_flag(const StepRec* recs, const unsigned long long* first_occ, uint32_t N, uint64_t i) {
    if (MODE == 0) { const uint32_t node = recs[i].node_rev >> 1; return node < N && first_occ[node] == i; }
    return first_occ[i] == ~0ull;
}
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS)
rl_tile_count(const StepRec* __restrict__ recs, const unsigned long long* __restrict__ first_occ, uint32_t N, uint64_t n_items,
              uint64_t* __restrict__ tile_cnt) {
    __shared__ uint64_t wb[32];
    const uint64_t base = (uint64_t)blockIdx.x * K1_TILE;
    uint64_t c = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
        if (i < n_items) c += rl_flag<MODE>(recs, first_occ, N, i) ? 1 : 0;
    }
    c = block_sum_u64(c, wb);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = c;
}
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS)
rl_assign(const StepRec* __restrict__ recs, const unsigned long long* __restrict__ first_occ, uint32_t N, uint64_t n_items,
          const uint64_t* __restrict__ tile_prefix, uint32_t* __restrict__ new_of_old, uint32_t* __restrict__ old_of_new) {
    __shared__ uint32_t wsum[K1_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * K1_TILE + (uint64_t)threadIdx.x * K1_ITEMS;   // 8 consecutive items
    bool f[K1_ITEMS];
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) { f[k] = (base + k < n_items) && rl_flag<MODE>(recs, first_occ, N, base + k); cnt += f[k]; }
    uint32_t total;
    uint32_t off = block_excl_scan_u32(cnt, wsum, &total);
    uint64_t rank = tile_prefix[blockIdx.x] + off;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        if (f[k]) {
            const uint32_t old = MODE == 0 ? (recs[base + k].node_rev >> 1) : (uint32_t)(base + k);
            new_of_old[old] = (uint32_t)rank;
            old_of_new[rank] = old;
            ++rank;
        }
    }
}
__global__ void rl_rewrite(StepRec* __restrict__ recs, uint64_t S, uint32_t N, const uint32_t* __restrict__ new_of_old) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const uint32_t nr = recs[i].node_rev;
    const uint32_t node = nr >> 1;
    if (node < N) recs[i].node_rev = (new_of_old[node] << 1) | (nr & 1u);
}
__global__ void rl_invert(const uint32_t* __restrict__ new_of_old, uint32_t N, uint32_t* __restrict__ old_of_new) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) old_of_new[new_of_old[i]] = i;
}

// =============================================================================================
// term sampling (shared by K2, K3, trace) — SURVEY.md Appendix A steps 1-5'
// =============================================================================================
struct KernelGraph {
    const StepRec* recs;
    const uint64_t* first_step;   // P+1 (global memory copy)
    const double* zetas;          // zlen entries (global)
    uint64_t S;
    uint32_t P;
    uint32_t N;
    uint32_t zlen;
    uint32_t space;               // min(params.space, 2^32-1): compared with ranks < 2^32
    uint32_t space_max;
    uint32_t q;
    uint32_t q_is_100;            // 1: the quantisation step is the reference's 100 (constant division)
    uint32_t blk_shift;           // path-of-block table granularity: block = step >> blk_shift
    uint32_t coherent;            // 1: the lanes of a warp sample 32 consecutive steps (see sample_s1)
    uint64_t samp_base, samp_len; // sampled steps are drawn from [samp_base, samp_base + samp_len) (default 0, S)
};

constexpr uint32_t SMEM_FS_MAX = 2048;         // first_step entries staged per block (16 KB)
constexpr uint32_t BLK_TABLE = 4096;           // path-of-block entries staged per block (8 KB)

// step -> path.  With the tables in shared memory: one 16-bit lookup (path of the first step of the
// step's 2^shift-block) plus a short forward scan; otherwise a binary search over first_step.
struct PathLookup {
    const uint64_t* fs;           // P+1 entries, shared or global
    const uint16_t* blk;          // BLK_TABLE entries in shared memory, or nullptr
    uint32_t shift;
    uint32_t P;
    __device__ __forceinline__ uint32_t path_of(uint64_t s) const {
        if (blk) {
            uint32_t p = blk[(uint32_t)(s >> shift)];
            while (s >= fs[p + 1]) ++p;
            return p;
        }
        return find_path(fs, P, s);
    }
};

// One term being sampled.  The stages are straight-line (selects, predicated loads): S1 turns the
// draw into the sampled step and issues the zeta load, S2 turns it into the partner step.  The
// kernel then requests both records and applies the update two pipeline stages later.
struct Slot {
    StepRec a, b;
    uint64_t step_a, step_b;   // step indices
    uint64_t f;                // first step of the path
    uint64_t r23;              // second half of the Philox block
    double zeta;
    uint32_t n, ra, J;
    uint32_t coins;            // r.z
    bool zipf, back, live;     // live: a partner is drawn (n > 1 and the Zipf branch has room to move)
    bool other_a, other_b;
    bool valid;
};

// Draw slots of one Philox block r (see oracle/gfs_oracle.cpp PhiloxDraw):
//   step = mulhi64(r.y:r.x, S); u = ((r.w:r.z) >> 11) * 2^-53; uniform rank = mulhi64(r.w:r.z, n);
//   coins = bits 0..3 of r.z (zipf, back, end_a, end_b).
__device__ __forceinline__ void sample_s1(const KernelGraph& g, const PathLookup& pl, const EpochDesc& ep, uint4 r,
                                          uint64_t win_base, uint64_t win_len, bool active, unsigned warp_mask,
                                          int lane, Slot& t) {
    const uint64_t r01 = ((uint64_t)r.y << 32) | r.x;
    t.r23 = ((uint64_t)r.w << 32) | r.z;
    t.coins = r.z;
    // step ~ U[win_base, win_base + win_len) on the circular sampling range; the default window
    // (samp_base, samp_len) = (0, S) is the reference's U[0, S) (sgd.rs:444)
    uint64_t s = win_base + __umul64hi(r01, win_len);
    if (g.coherent) {
        // warp-coherent sampling (sweep schedule only): the warp's first lane draws the step, lane l takes
        // the l-th step after it.  Every step is still drawn with the same probability over a sweep, but
        // the 32 sampled records — and, with the node relabelling, most of their nodes' positions — are
        // adjacent in memory: one coalesced request instead of 32.  Partners stay independent per lane.
        s = __shfl_sync(warp_mask, s, __ffs(warp_mask) - 1) + (uint32_t)lane;
    }
    if (s >= g.samp_base + g.samp_len) s -= g.samp_len;
    t.step_a = s;
    const uint32_t p = pl.path_of(s);
    t.f = pl.fs[p];
    const uint32_t n = (uint32_t)(pl.fs[p + 1] - t.f);
    const uint32_t ra = (uint32_t)(s - t.f);
    t.n = n; t.ra = ra;
    t.zipf = ep.cooling || (t.coins & 1u);                                                 // sgd.rs:456
    t.back = ra > 0 && (((t.coins >> 1) & 1u) || ra == n - 1);                             // sgd.rs:460
    const bool fwd = !t.back && ra < n - 1;                                                // sgd.rs:475
    const bool moves = t.back || fwd;
    const uint32_t span = t.back ? ra : n - ra - 1;
    const uint32_t J = span < g.space ? span : g.space;
    t.J = J;
    uint32_t k = J;                                                                        // sgd.rs:463-467
    if (J > g.space_max) {
        const uint32_t over = J - g.space_max;
        k = g.space_max + (g.q_is_100 ? over / 100u : over / g.q) + 1;
    }
    k = k < g.zlen - 1 ? k : g.zlen - 1;                                                   // sgd.rs:469
    t.live = active && n > 1 && (!t.zipf || moves);        // n == 1 => continue (sgd.rs:448)
    t.zeta = 1.0;
    if (t.live && t.zipf) t.zeta = __ldg(g.zetas + k);
}

__device__ __forceinline__ void sample_s2(const KernelGraph& g, const EpochDesc& ep, Slot& t) {
    const uint32_t n = t.n, ra = t.ra;
    // u = (r23 >> 11) * 2^-53 (sgd.rs:136 through PhiloxDraw::unit)
    const double u = __dmul_rn((double)(t.r23 >> 11), 1.0 / 9007199254740992.0);
    const ZipfPre pre = dirty_zipf_pre(t.J, ep.zc);
    const uint32_t z = dirty_zipf_post(t.J, ep.zc, pre, t.zeta, u);
    const uint32_t room = n - 1 - ra;
    const uint32_t rb_back = ra >= z ? ra - z : 0u;                                        // saturating_sub
    const uint32_t rb_fwd = z < room ? ra + z : n - 1;                                     // min(ra + z, n - 1)
    const uint32_t rb_zipf = t.back ? rb_back : rb_fwd;
    const uint32_t rb_unif = (uint32_t)__umul64hi(t.r23, (uint64_t)n);                     // sgd.rs:493-494
    const uint32_t rb = t.zipf ? rb_zipf : rb_unif;
    t.valid = t.live && ra != rb;                                                          // sgd.rs:497
    t.step_b = t.valid ? t.f + rb : t.step_a;
    t.other_a = t.other_b = false;
}

// nD end choice (sgd.rs:1060-1077); needs both records.
__device__ __forceinline__ void sample_ends(Slot& t) {
    const bool rev_a = t.a.node_rev & 1u, rev_b = t.b.node_rev & 1u;
    bool ua = (t.coins >> 2) & 1u;
    if (ua) { t.a.pos += t.a.node_len; ua = !rev_a; } else { ua = rev_a; }
    bool ub = (t.coins >> 3) & 1u;
    if (ub) { t.b.pos += t.b.node_len; ub = !rev_b; } else { ub = rev_b; }
    t.other_a = ua; t.other_b = ub;
}

__device__ __forceinline__ double term_distance(const Slot& t) {
    return fabs(__dsub_rn(u52_to_f64(t.a.pos), u52_to_f64(t.b.pos)));                      // sgd.rs:509-513
}

// =============================================================================================
// coordinate access for K3
// =============================================================================================
template <typename CT> struct Arith;
template <> struct Arith<double> {
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    static __device__ __forceinline__ double sqrt(double a) { return __dsqrt_rn(a); }
};
template <> struct Arith<float> {
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};

// positions are written by atomics at L2 and read here: bypass the (incoherent) L1 with ld.cg
__device__ __forceinline__ double ld_pos(const double* p) {
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
template <typename CT, int DS> __device__ __forceinline__ void ld_coords(const CT* p, CT (&c)[DS]);
template <> __device__ __forceinline__ void ld_coords<float, 1>(const float* p, float (&c)[1]) {
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(c[0]) : "l"(p));
}
template <> __device__ __forceinline__ void ld_coords<float, 2>(const float* p, float (&c)[2]) {
    asm volatile("ld.global.cg.v2.f32 {%0,%1}, [%2];" : "=f"(c[0]), "=f"(c[1]) : "l"(p));
}
template <> __device__ __forceinline__ void ld_coords<float, 4>(const float* p, float (&c)[4]) {
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "l"(p));
}
template <> __device__ __forceinline__ void ld_coords<float, 8>(const float* p, float (&c)[8]) {
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3]) : "l"(p));
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c[4]), "=f"(c[5]), "=f"(c[6]), "=f"(c[7]) : "l"(p + 4));
}
template <> __device__ __forceinline__ void ld_coords<double, 1>(const double* p, double (&c)[1]) { c[0] = ld_pos(p); }
template <> __device__ __forceinline__ void ld_coords<double, 2>(const double* p, double (&c)[2]) {
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "l"(p));
}
template <> __device__ __forceinline__ void ld_coords<double, 4>(const double* p, double (&c)[4]) {
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(c[0]), "=d"(c[1]) : "l"(p));
    asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(c[2]), "=d"(c[3]) : "l"(p + 2));
}
template <> __device__ __forceinline__ void ld_coords<double, 8>(const double* p, double (&c)[8]) {
#pragma unroll
    for (int k = 0; k < 8; k += 2)
        asm volatile("ld.global.cg.v2.f64 {%0,%1}, [%2];" : "=d"(c[k]), "=d"(c[k + 1]) : "l"(p + k));
}

template <typename CT, int DS> __device__ __forceinline__ void red_coords(CT* p, const CT (&d)[DS]);
template <> __device__ __forceinline__ void red_coords<float, 1>(float* p, const float (&d)[1]) { atomicAdd(p, d[0]); }
template <> __device__ __forceinline__ void red_coords<float, 2>(float* p, const float (&d)[2]) {
    atomicAdd(reinterpret_cast<float2*>(p), make_float2(d[0], d[1]));          // red.global.add.v2.f32
}
template <> __device__ __forceinline__ void red_coords<float, 4>(float* p, const float (&d)[4]) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(d[0], d[1], d[2], d[3]));   // red.global.add.v4.f32
}
template <> __device__ __forceinline__ void red_coords<float, 8>(float* p, const float (&d)[8]) {
    atomicAdd(reinterpret_cast<float4*>(p), make_float4(d[0], d[1], d[2], d[3]));
    atomicAdd(reinterpret_cast<float4*>(p + 4), make_float4(d[4], d[5], d[6], d[7]));
}
template <> __device__ __forceinline__ void red_coords<double, 1>(double* p, const double (&d)[1]) { atomicAdd(p, d[0]); }
template <> __device__ __forceinline__ void red_coords<double, 2>(double* p, const double (&d)[2]) {
    atomicAdd(p, d[0]); atomicAdd(p + 1, d[1]);
}
template <> __device__ __forceinline__ void red_coords<double, 4>(double* p, const double (&d)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) atomicAdd(p + k, d[k]);
}
template <> __device__ __forceinline__ void red_coords<double, 8>(double* p, const double (&d)[8]) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(p + k, d[k]);
}

// =============================================================================================
// K2 / K3 — persistent SGD term kernel
// =============================================================================================
constexpr int SGD_BLOCK = 256;

struct SgdArgs {
    KernelGraph g;
    const EpochDesc* epochs;     // device array, iter_max+1 entries
    uint32_t epoch_begin, epoch_end;
    uint32_t slice, n_slices;    // run slice `slice` of n_slices equal parts of every epoch's updates
    uint64_t* attempt_ctr;       // per-thread Philox attempt counters (persist across launches)
    unsigned long long* counters;   // [0] applied, [1] attempts, [2] watchdog trips
    uint32_t seed_lo, seed_hi;
    uint32_t tid_base;
    void* positions;             // 1D: double[N]; nD: CT[N*2*DS]
    // sweep scheduling (window_steps > 0): warps claim chunks of `chunk_updates` updates from *work_ctr;
    // chunk c samples its steps from a window of `window_steps` steps that slides once over the step
    // array per epoch, so the records being sampled stay L2-resident.
    uint64_t window_steps;
    uint32_t chunk_updates;
    unsigned long long* work_ctr;
    uint64_t iter_cap;           // watchdog: a warp that loops more often than this sets counters[2] and stops
};

// 1D update of one warp's terms (sgd.rs:512-576), optionally merging lanes that hit the same node.
// r_x is the displacement computed from positions xi, xj that were loaded earlier (stage S3).
template <bool AGG>
__device__ __forceinline__ void apply_1d(double* X, unsigned warp_mask, int lane, bool valid, uint32_t i,
                                         uint32_t j, double d, double eta, double xi, double xj) {
    double r_x = 0.0;
    auto add = [&](double* p, double v) { atomicAdd(p, v); };        // red.global.add.f64 (result unused)
    if (valid) {
        const double mu = fmin(__dmul_rn(eta, __ddiv_rn(1.0, d)), 1.0);     // sgd.rs:518-520
        double dx = __dsub_rn(xi, xj);
        if (dx == 0.0) dx = 1e-9;                                            // sgd.rs:546-548
        const double mag = fabs(dx);
        const double delta = __dmul_rn(__dmul_rn(mu, __dsub_rn(mag, d)), 0.5);   // sgd.rs:552
        const double r = __ddiv_rn(delta, mag);
        r_x = __dmul_rn(r, dx);
    }
    if (AGG) {
        const unsigned vmask = __ballot_sync(warp_mask, valid);
        const unsigned mi = __match_any_sync(warp_mask, i) & vmask;
        const unsigned mj = __match_any_sync(warp_mask, j) & vmask;
        const bool dup = valid && (__popc(mi) > 1 || __popc(mj) > 1);
        if (!__any_sync(warp_mask, dup)) {               // common case: 64 distinct nodes in the warp
            if (valid) { add(X + i, -r_x); add(X + j, r_x); }
            return;
        }
        bool lead;
        const double si = group_sum(warp_mask, valid ? mi : 0u, -r_x, lane, lead);
        if (valid && lead) add(X + i, si);
        const double sj = group_sum(warp_mask, valid ? mj : 0u, r_x, lane, lead);
        if (valid && lead) add(X + j, sj);
    } else if (valid) {
        add(X + i, -r_x);                                                    // sgd.rs:575
        add(X + j, r_x);                                                     // sgd.rs:576
    }
}

// nD update (sgd.rs:1079-1149) on coordinates laid out [node][end][DS] (DS >= D, padded with zeros).
template <typename CT, int D, int DS, bool AGG>
__device__ __forceinline__ void apply_nd(CT* C, unsigned warp_mask, int lane, bool valid, uint32_t idx_i,
                                         uint32_t idx_j, double d, double eta, const CT (&ci)[DS], const CT (&cj)[DS]) {
    using A = Arith<CT>;
    CT di[DS], dj[DS];
#pragma unroll
    for (int k = 0; k < DS; ++k) { di[k] = CT(0); dj[k] = CT(0); }
    if (valid) {
        const CT mu = (CT)fmin(__dmul_rn(eta, __ddiv_rn(1.0, d)), 1.0);      // sgd.rs:1085-1086
        CT dl[DS];
        CT mag_sq = CT(0);
#pragma unroll
        for (int k = 0; k < D; ++k) { dl[k] = A::sub(ci[k], cj[k]); mag_sq = A::add(mag_sq, A::mul(dl[k], dl[k])); }
        if (mag_sq == CT(0)) { dl[0] = (CT)1e-9; mag_sq = (CT)1e-18; }       // sgd.rs:1116-1119
        const CT mag = A::sqrt(mag_sq);
        const CT delta = A::mul(A::mul(mu, A::sub(mag, (CT)d)), CT(0.5));    // sgd.rs:1125
        const CT r = A::div(delta, mag);
#pragma unroll
        for (int k = 0; k < D; ++k) { const CT rd = A::mul(r, dl[k]); di[k] = -rd; dj[k] = rd; }
    }
    if (AGG) {
        const unsigned vmask = __ballot_sync(warp_mask, valid);
        const unsigned mi = __match_any_sync(warp_mask, idx_i) & vmask;
        const unsigned mj = __match_any_sync(warp_mask, idx_j) & vmask;
        const bool dup = valid && (__popc(mi) > 1 || __popc(mj) > 1);
        if (!__any_sync(warp_mask, dup)) {
            if (valid) { red_coords<CT, DS>(C + (size_t)idx_i * DS, di); red_coords<CT, DS>(C + (size_t)idx_j * DS, dj); }
            return;
        }
        bool lead_i, lead_j;
#pragma unroll
        for (int k = 0; k < D; ++k) di[k] = group_sum(warp_mask, valid ? mi : 0u, di[k], lane, lead_i);
        if (valid && lead_i) red_coords<CT, DS>(C + (size_t)idx_i * DS, di);
#pragma unroll
        for (int k = 0; k < D; ++k) dj[k] = group_sum(warp_mask, valid ? mj : 0u, dj[k], lane, lead_j);
        if (valid && lead_j) red_coords<CT, DS>(C + (size_t)idx_j * DS, dj);
    } else if (valid) {
        red_coords<CT, DS>(C + (size_t)idx_i * DS, di);
        red_coords<CT, DS>(C + (size_t)idx_j * DS, dj);
    }
}

// D == 0: 1D `Y` (CT must be double).  D >= 1: nD `L`.  K: terms in flight per thread.
template <typename CT, int D, int DS, bool AGG, int K>
#ifndef GFS_K1_BLOCKS
#define GFS_K1_BLOCKS 4
#endif
__global__ void __launch_bounds__(SGD_BLOCK, (K > 1 ? 3 : GFS_K1_BLOCKS))
sgd_kernel(const SgdArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // shared: first_step (P+1 u64) + path-of-block table (BLK_TABLE u16) when the path table fits
    uint64_t* s_fs = reinterpret_cast<uint64_t*>(smem_raw);
    const bool tables = a.g.P + 1 <= SMEM_FS_MAX;
    const uint32_t n_fs = tables ? a.g.P + 1 : 0;
    uint16_t* s_blk = reinterpret_cast<uint16_t*>(smem_raw + (size_t)n_fs * 8);
    for (uint32_t k = threadIdx.x; k < n_fs; k += blockDim.x) s_fs[k] = a.g.first_step[k];
    __syncthreads();
    if (tables) {
        for (uint32_t k = threadIdx.x; k < BLK_TABLE; k += blockDim.x) {
            const uint64_t s0 = (uint64_t)k << a.g.blk_shift;
            s_blk[k] = (uint16_t)(s0 < a.g.S ? find_path(s_fs, a.g.P, s0) : a.g.P - 1);
        }
        __syncthreads();
    }
    PathLookup pl;
    pl.fs = tables ? s_fs : a.g.first_step;
    pl.blk = tables ? s_blk : nullptr;
    pl.shift = a.g.blk_shift;
    pl.P = a.g.P;

    const unsigned warp_mask = __activemask();
    const int lane = threadIdx.x & 31;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t T = gridDim.x * blockDim.x;
    uint64_t attempt = a.attempt_ctr[tid];
    uint64_t applied = 0, n_attempts = 0;
    const uint2 key = make_uint2(a.seed_lo, a.seed_hi);

    // ---- software pipeline -------------------------------------------------------------------------
    // Three terms per slot are in different stages at any time (K slots per thread):
    //   stage A   sample term i+3: Philox, path lookup, zeta load, Zipf arithmetic -> the two step indices
    //   stage L   request the two records of term i+2 (indices from the previous A)
    //   stage B1  term i+1: its records were requested one iteration ago and have had the whole of
    //             stage A (~350 instructions and an L2 round trip) to arrive from DRAM; term distance,
    //             validity, request the two positions
    //   stage B2  term i: its positions have had one iteration to arrive; compute the update, apply it (red)
    // Loop order is B2, B1, L, A, so that every register set is reloaded only after its consumer has run
    // and no in-flight value is ever moved.  A term's validity is final only in B1 (zero distance,
    // missing node), so a lane's quota is tracked optimistically: owed = target - done - in flight.
    struct InFlight {            // term whose records are being loaded (L -> B)
        StepRec a, b;
        double eta;
        uint32_t coins;
        bool valid;
    };
    struct Sampled {             // term whose steps are known (A -> L)
        uint64_t sa, sb;
        double eta;
        uint32_t coins;
        bool valid;
    };
    InFlight fl[K];
    Sampled sm[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        fl[k].valid = false; fl[k].coins = 0; fl[k].eta = 0.0;
        fl[k].a.node_rev = fl[k].a.node_len = 0; fl[k].a.pos = 0; fl[k].b = fl[k].a;
        sm[k].valid = false; sm[k].sa = sm[k].sb = 0; sm[k].coins = 0; sm[k].eta = 0.0;
    }
    uint64_t target = 0, done = 0;           // per lane: updates owed by the chunks claimed so far / applied

    // work claiming (warp-uniform): sets ep / win_base / win_len and raises `target`
    const bool sweep = a.window_steps != 0;
    const uint32_t n_lanes = __popc(warp_mask);
    const uint32_t lane_rank = __popc(warp_mask & ((1u << lane) - 1u));
    const int leader = __ffs(warp_mask) - 1;
    const uint32_t C = a.chunk_updates;
    EpochDesc ep = a.epochs[a.epoch_begin];
    uint64_t win_base = a.g.samp_base, win_len = a.g.samp_len;
    uint32_t e_cur = a.epoch_begin;
    // sweep schedule (see SgdArgs): chunks of C updates claimed from a global counter; chunk cc of an epoch
    // samples from the window starting at cc * samp_len / chunks_per_epoch.  Claiming in order keeps all
    // warps on neighbouring chunks (a static assignment lets them drift apart by SM speed), so the windows
    // in use at any moment cover about n_warps * C * samp_len / m + window_steps consecutive steps.
    const uint64_t m_sweep = ep.updates / a.n_slices + (a.slice < ep.updates % a.n_slices ? 1 : 0);
    const uint64_t cpe = sweep ? (m_sweep + C - 1) / C : 1;                           // chunks per epoch
    const uint64_t total_chunks = cpe * (uint64_t)(a.epoch_end - a.epoch_begin);
    const double steps_per_chunk = (double)a.g.samp_len / (double)cpe;
    uint64_t c_lo = 0;                        // first chunk of epoch e_cur
    bool first_claim = true;
    auto claim = [&]() -> bool {
        if (sweep) {
            unsigned long long c = 0;
            if (lane == leader) c = atomicAdd(a.work_ctr, 1ull);
            c = __shfl_sync(warp_mask, c, leader);
            if (c >= total_chunks) return false;
            while (c >= c_lo + cpe) { c_lo += cpe; ++e_cur; ep = a.epochs[e_cur]; }       // claims only move forward
            const uint64_t cc = c - c_lo;
            const uint64_t left = m_sweep - cc * C;
            const uint32_t n_upd = left < C ? (uint32_t)left : C;
            target += n_lanes == 32 ? (n_upd >> 5) + (lane_rank < (n_upd & 31u) ? 1u : 0u)
                                    : n_upd / n_lanes + (lane_rank < n_upd % n_lanes ? 1u : 0u);
            uint64_t off = (uint64_t)((double)cc * steps_per_chunk);
            if (off >= a.g.samp_len) off = a.g.samp_len - 1;
            win_base = a.g.samp_base + off;
            win_len = a.window_steps;
            return true;
        }
        // static schedule: one "chunk" per epoch; thread t applies floor(m/T) + (t < m%T) updates, steps ~ U[0,S)
        if (!first_claim) ++e_cur;
        first_claim = false;
        if (e_cur >= a.epoch_end) return false;
        ep = a.epochs[e_cur];
        const uint64_t m = ep.updates / a.n_slices + (a.slice < ep.updates % a.n_slices ? 1 : 0);
        target += m / T + (tid < m % T ? 1 : 0);
        return true;
    };
    bool more = true;

    struct Loaded {              // term whose positions are being loaded (B1 -> B2)
        CT ci[DS], cj[DS];
        double dist, eta;
        uint32_t idx_i, idx_j;
        bool ok;
    };
    Loaded xs[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        xs[k].ok = false; xs[k].dist = 0.0; xs[k].eta = 0.0; xs[k].idx_i = xs[k].idx_j = 0;
#pragma unroll
        for (int q = 0; q < DS; ++q) { xs[k].ci[q] = CT(0); xs[k].cj[q] = CT(0); }
    }

    uint64_t iters = 0;
    for (;;) {
        if (++iters > a.iter_cap) {          // never taken in a healthy run; turns a would-be hang into an error
            if (lane == leader) atomicAdd(a.counters + 2, 1ull);
            break;
        }
        // ---- B2: apply the terms whose positions were requested in the previous iteration
        bool any_ok = false;
#pragma unroll
        for (int k = 0; k < K; ++k) any_ok = any_ok || xs[k].ok;
#ifdef GFS_EXP_NOAPPLY
#pragma unroll
        for (int k = 0; k < K; ++k) { done += xs[k].ok ? 1u : 0u; if (xs[k].dist == 1.2345e-300) ((double*)a.positions)[0] = (double)xs[k].ci[0] + (double)xs[k].cj[0]; }
        any_ok = false;
#endif
        if (__any_sync(warp_mask, any_ok)) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if constexpr (D == 0) {
                    apply_1d<AGG>(reinterpret_cast<double*>(a.positions), warp_mask, lane, xs[k].ok, xs[k].idx_i, xs[k].idx_j,
                                  xs[k].dist, xs[k].eta, xs[k].ci[0], xs[k].cj[0]);
                } else {
                    apply_nd<CT, (D > 0 ? D : 1), DS, AGG>(reinterpret_cast<CT*>(a.positions), warp_mask, lane, xs[k].ok,
                                                           xs[k].idx_i, xs[k].idx_j, xs[k].dist, xs[k].eta, xs[k].ci, xs[k].cj);
                }
                done += xs[k].ok ? 1u : 0u;                                                // sgd.rs:579
            }
        }
        // ---- B1: the records requested in the previous iteration have arrived: term distance, validity,
        //          and the requests for the two positions
        uint32_t pending = 0;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            bool oa = false, ob = false;
            StepRec ra = fl[k].a, rb = fl[k].b;
            if (D > 0) {                                                                   // sgd.rs:1060-1077
                const bool rev_a = ra.node_rev & 1u, rev_b = rb.node_rev & 1u;
                oa = (fl[k].coins >> 2) & 1u;
                if (oa) { ra.pos += ra.node_len; oa = !rev_a; } else { oa = rev_a; }
                ob = (fl[k].coins >> 3) & 1u;
                if (ob) { rb.pos += rb.node_len; ob = !rev_b; } else { ob = rev_b; }
            }
            xs[k].dist = fabs(__dsub_rn(u52_to_f64(ra.pos), u52_to_f64(rb.pos)));          // sgd.rs:509-513
            xs[k].eta = fl[k].eta;
            const uint32_t na = ra.node_rev >> 1, nb = rb.node_rev >> 1;
            xs[k].ok = fl[k].valid && xs[k].dist != 0.0 && na < a.g.N && nb < a.g.N;       // sgd.rs:514, 525-538
            if constexpr (D == 0) {
                xs[k].idx_i = na; xs[k].idx_j = nb;
                const double* X = reinterpret_cast<const double*>(a.positions);
                xs[k].ci[0] = xs[k].cj[0] = 0.0;
                if (xs[k].ok) { xs[k].ci[0] = ld_pos(X + na); xs[k].cj[0] = ld_pos(X + nb); }
            } else {
                xs[k].idx_i = na * 2 + (oa ? 1u : 0u);                                     // sgd.rs:1099-1103
                xs[k].idx_j = nb * 2 + (ob ? 1u : 0u);
                const CT* Cc = reinterpret_cast<const CT*>(a.positions);
#pragma unroll
                for (int q = 0; q < DS; ++q) { xs[k].ci[q] = CT(0); xs[k].cj[q] = CT(0); }
                if (xs[k].ok) {
                    ld_coords<CT, DS>(Cc + (size_t)xs[k].idx_i * DS, xs[k].ci);
                    ld_coords<CT, DS>(Cc + (size_t)xs[k].idx_j * DS, xs[k].cj);
                }
            }
            pending += xs[k].ok ? 1u : 0u;
        }
        // ---- L: request the records of the terms sampled in the previous iteration
#pragma unroll
        for (int k = 0; k < K; ++k) {
            fl[k].valid = sm[k].valid; fl[k].coins = sm[k].coins; fl[k].eta = sm[k].eta;
            if (sm[k].valid) {
                fl[k].a = load_rec(a.g.recs + sm[k].sa);
                fl[k].b = load_rec(a.g.recs + sm[k].sb);
                ++pending;
            }
            sm[k].valid = false;
        }
        // ---- A: sample the next terms
        uint64_t owed = target - done - pending;
        // a new chunk is claimed when every lane has sampled its share; the static schedule (which the
        // bit-exact single-thread tests use) also waits until that share is confirmed applied, so that a
        // term never runs with the next epoch's eta
        const bool need_claim = sweep ? !__any_sync(warp_mask, owed != 0)
                                      : !__any_sync(warp_mask, owed != 0 || pending != 0);
        if (need_claim && more) {
            more = claim();
            owed = target - done - pending;
        }
        if (__any_sync(warp_mask, owed != 0)) {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const bool active = owed > (uint64_t)k;
                const uint4 r = philox4x32_10(make_uint4((uint32_t)attempt, (uint32_t)(attempt >> 32),
                                                         a.tid_base + tid, STREAM_SGD), key);
                // the Philox counter advances with every draw this lane makes — also on draws made only on
                // behalf of other lanes (warp-coherent steps, grouped partners): a lane that has finished
                // its share must not keep serving the same block to its neighbours
                attempt += (active || a.g.coherent) ? 1 : 0;
                n_attempts += active ? 1 : 0;
                Slot t;
                sample_s1(a.g, pl, ep, r, win_base, win_len, active, warp_mask, lane, t);
                sample_s2(a.g, ep, t);
                sm[k].sa = t.step_a; sm[k].sb = t.step_b; sm[k].valid = t.valid; sm[k].coins = t.coins; sm[k].eta = ep.eta;
            }
        } else if (!more && !__any_sync(warp_mask, pending != 0)) {
            break;
        }
    }
    applied = done;
    a.attempt_ctr[tid] = attempt;
    // counters: one atomic pair per full warp (partial warps: one pair per thread)
    uint64_t att = n_attempts;
    if (warp_mask == 0xffffffffu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            applied += __shfl_xor_sync(0xffffffffu, applied, o);
            att += __shfl_xor_sync(0xffffffffu, att, o);
        }
        if (lane != 0) return;
    }
    atomicAdd(a.counters + 0, (unsigned long long)applied);
    atomicAdd(a.counters + 1, (unsigned long long)att);
}

// =============================================================================================
// K4 — sampled stress (sgd.rs:1196-1283)
// =============================================================================================
constexpr int STRESS_BLOCK = 256;
// coords: stride_node doubles per node, the + end's `dims` coordinates first.
__global__ void __launch_bounds__(STRESS_BLOCK)
stress_kernel(KernelGraph g, const uint32_t* __restrict__ old_of_new, const double* __restrict__ coords, uint32_t dims,
              uint32_t stride_node, uint64_t samples, uint32_t seed_lo, uint32_t seed_hi, double* __restrict__ partial /*3 per block*/) {
    __shared__ double red[3][STRESS_BLOCK / 32];
    double sum = 0.0, sum_abs = 0.0, cnt = 0.0;
    const uint2 key = make_uint2(seed_lo, seed_hi);
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < samples; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)k, (uint32_t)(k >> 32), 0u, STREAM_STRESS), key);
        const uint64_t s = __umul64hi(((uint64_t)r.y << 32) | r.x, g.S);
        const uint32_t p = find_path(g.first_step, g.P, s);
        const uint64_t f = g.first_step[p];
        const uint32_t n = (uint32_t)(g.first_step[p + 1] - f);
        if (n < 2) continue;
        const uint32_t ra = (uint32_t)(s - f);
        const uint32_t rb = (uint32_t)__umul64hi(((uint64_t)r.w << 32) | r.z, (uint64_t)n);
        if (ra == rb) continue;
        const StepRec A = load_rec(g.recs + s), B = load_rec(g.recs + f + rb);
        const double dp = fabs(__dsub_rn((double)A.pos, (double)B.pos));
        if (dp == 0.0) continue;
        uint32_t ia = A.node_rev >> 1, ib = B.node_rev >> 1;
        if (ia >= g.N || ib >= g.N) continue;
        if (old_of_new) { ia = old_of_new[ia]; ib = old_of_new[ib]; }
        double sq = 0.0;
        for (uint32_t d = 0; d < dims; ++d) {
            const double dl = __dsub_rn(coords[(size_t)ia * stride_node + d], coords[(size_t)ib * stride_node + d]);
            sq = __dadd_rn(sq, __dmul_rn(dl, dl));
        }
        const double err = __dsub_rn(__dsqrt_rn(sq), dp);
        sum += __ddiv_rn(__dmul_rn(err, err), __dmul_rn(dp, dp));
        sum_abs += __ddiv_rn(fabs(err), dp);
        cnt += 1.0;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sum_abs += __shfl_xor_sync(0xffffffffu, sum_abs, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if (lane == 0) { red[0][w] = sum; red[1][w] = sum_abs; red[2][w] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0, a1 = 0, a2 = 0;
        for (int k = 0; k < STRESS_BLOCK / 32; ++k) { a0 += red[0][k]; a1 += red[1][k]; a2 += red[2][k]; }
        partial[blockIdx.x * 3 + 0] = a0; partial[blockIdx.x * 3 + 1] = a1; partial[blockIdx.x * 3 + 2] = a2;
    }
}

// =============================================================================================
// conversion kernels (host Layout order f64 <-> device [node][end][DS] CT)
// =============================================================================================
// src: host order, f64, `ends` node ends of D coordinates each per node (1D: ends = 1, D = 1).
// dst: device order (node relabelled through new_of_old when non-null), CT, stride DS per end.
template <typename CT>
__global__ void pos_to_device(const double* __restrict__ src, CT* __restrict__ dst, uint64_t N, uint32_t ends, uint32_t D,
                              uint32_t DS, const uint32_t* __restrict__ new_of_old) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_node = (uint64_t)ends * DS;
    if (i >= N * per_node) return;
    const uint64_t node = i / per_node; const uint32_t r = (uint32_t)(i % per_node);
    const uint32_t e = r / DS, k = r % DS;
    const uint64_t dn = new_of_old ? new_of_old[node] : node;
    dst[dn * per_node + r] = k < D ? (CT)src[(node * ends + e) * D + k] : CT(0);
}
template <typename CT>
__global__ void pos_from_device(const CT* __restrict__ src, double* __restrict__ dst, uint64_t N, uint32_t ends, uint32_t D,
                                uint32_t DS, const uint32_t* __restrict__ new_of_old) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t per_node = (uint64_t)ends * D;
    if (i >= N * per_node) return;
    const uint64_t node = i / per_node; const uint32_t r = (uint32_t)(i % per_node);
    const uint32_t e = r / D, k = r % D;
    const uint64_t dn = new_of_old ? new_of_old[node] : node;
    dst[i] = (double)src[(dn * ends + e) * DS + k];
}

// =============================================================================================
// K6 — order by position (the host side of path_sgd_sort, src/sgd.rs:659-671)
// =============================================================================================
// Stable LSD radix sort of (key = order-preserving image of the f64 position, value = dense idx), 8-bit
// digits, 8 passes.  Stability + values starting as 0..n-1 gives "ties by dense idx" (the reference sorts
// (idx, pos) pairs in HashMap iteration order with a stable sort, i.e. its tie order is unspecified).
constexpr int RS_THREADS = 256;
constexpr int RS_ITEMS = 16;                       // keys per thread, processed in index order
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;

// f64 -> u64 whose unsigned order is the numeric order; -0.0 == +0.0 (partial_cmp: Equal); NaN last.
__device__ __forceinline__ uint64_t f64_sort_key(double v) {
    if (v != v) return ~0ull;
    if (v == 0.0) v = 0.0;
    const uint64_t b = (uint64_t)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void rs_make_keys(const double* __restrict__ x, uint64_t n, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = f64_sort_key(x[i]); vals[i] = (uint32_t)i; }
}
// hist[d * n_blocks + b] = number of keys of tile b whose digit is d
__global__ void __launch_bounds__(RS_THREADS)
rs_hist(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint32_t n_blocks, uint32_t* __restrict__ hist) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}
// exclusive scan of hist in place (digit-major), one block
__global__ void __launch_bounds__(1024) rs_scan(uint32_t* __restrict__ hist, uint64_t m) {
    __shared__ uint32_t wsum[32];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint64_t b0 = 0; b0 < m; b0 += 1024) {
        const uint64_t i = b0 + threadIdx.x;
        const uint32_t v = i < m ? hist[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[w] = inc;
        __syncthreads();
        if (w == 0) {
            const uint32_t ws = wsum[lane];
            uint32_t wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += t; }
            wsum[lane] = wi - ws;
        }
        __syncthreads();
        const uint32_t excl = running + wsum[w] + inc - v;
        if (i < m) hist[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) running = excl + v;
        __syncthreads();
    }
}
// stable scatter: tile b writes its keys of digit d to offs[d * n_blocks + b] + (rank among the tile's digit-d keys)
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, int shift, uint32_t n_blocks,
           const uint32_t* __restrict__ offs, uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t bin_off[256];                       // next free slot of each digit for this tile
    __shared__ uint32_t warp_cnt[RS_THREADS / 32][256];
    bin_off[threadIdx.x] = offs[(size_t)threadIdx.x * n_blocks + blockIdx.x];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    for (int k = 0; k < RS_ITEMS; ++k) {                    // 256 consecutive keys at a time, in index order
#pragma unroll
        for (int q = 0; q < RS_THREADS / 32; ++q) warp_cnt[q][threadIdx.x] = 0;
        __syncthreads();
        const uint64_t i = base + (uint64_t)k * RS_THREADS + threadIdx.x;
        const bool ok = i < n;
        const uint64_t key = ok ? keys[i] : 0ull;
        const uint32_t d = ok ? ((uint32_t)(key >> shift) & 255u) : 256u;       // 256: matches only other padding lanes
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) warp_cnt[w][d] = __popc(peers);
        __syncthreads();
        {   // thread t owns digit t: turn per-warp counts into per-warp offsets, advance the tile's cursor
            uint32_t run = bin_off[threadIdx.x];
#pragma unroll
            for (int q = 0; q < RS_THREADS / 32; ++q) { const uint32_t c = warp_cnt[q][threadIdx.x]; warp_cnt[q][threadIdx.x] = run; run += c; }
            bin_off[threadIdx.x] = run;
        }
        __syncthreads();
        if (ok) {
            const uint32_t dst = warp_cnt[w][d] + rank;
            keys_out[dst] = key;
            vals_out[dst] = vals[i];
        }
        __syncthreads();
    }
}

// =============================================================================================
// K5 — replica reconcile helpers (multi-GPU; the all-reduce itself is NCCL, driven by the host)
// =============================================================================================
// pack:  buf[i] = float(x[i] - x_sync[i]),  buf[n + i] = (x[i] != x_sync[i])      (one f32 buffer, one all-reduce)
// apply: x[i] = x_sync[i] + buf[i] / max(buf[n + i], 1);  x_sync[i] = x[i]
// i.e. the mean of the displacements over the replicas that moved the element since the last sync.
template <typename CT>
__global__ void __launch_bounds__(256) rc_pack(const CT* __restrict__ x, const CT* __restrict__ xs, uint64_t n, float* __restrict__ buf) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const CT d = x[i] - xs[i];
        buf[i] = (float)d;
        buf[n + i] = d != CT(0) ? 1.0f : 0.0f;
    }
}
template <typename CT>
__global__ void __launch_bounds__(256) rc_apply(CT* __restrict__ x, CT* __restrict__ xs, uint64_t n, const float* __restrict__ buf) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float c = buf[n + i];
        const CT v = xs[i] + (CT)buf[i] / (CT)(c > 1.0f ? c : 1.0f);
        x[i] = v; xs[i] = v;
    }
}

// =============================================================================================
// debug kernels
// =============================================================================================
__global__ void dbg_fpp(const double* a, const double* b, double* out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fast_precise_pow(a[i], b[i]);
}
__global__ void dbg_zipf(const uint64_t* zmax, const double* theta, const double* zeta, const double* u, uint64_t* out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ZipfConsts zc;
    zc.theta = theta[i];
    zc.one_minus_theta = __dsub_rn(1.0, theta[i]);
    zc.alpha = __ddiv_rn(1.0, __dsub_rn(1.0, theta[i]));
    zc.z2 = __dadd_rn(1.0, fast_precise_pow(0.5, theta[i]));
    zc.alpha_e = __double2int_rz(zc.alpha);
    zc.alpha_frac = __dsub_rn(zc.alpha, (double)zc.alpha_e);
    out[i] = dirty_zipf((uint32_t)zmax[i], zc, zeta[i], u[i]);
}
__global__ void dbg_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out, uint64_t n) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 r = philox4x32_10(make_uint4(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3]),
                                  make_uint2(key[2 * i], key[2 * i + 1]));
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}
template <bool ND>
__global__ void dbg_trace(KernelGraph g, const EpochDesc* epochs, uint32_t epoch, uint32_t seed_lo, uint32_t seed_hi,
                          uint32_t tid, uint64_t attempt0, uint64_t count, uint8_t* valid, uint64_t* step_a,
                          uint64_t* step_b, uint8_t* flags, double* dist) {
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    const EpochDesc ep = epochs[epoch];
    const uint64_t attempt = attempt0 + k;
    const uint4 r = philox4x32_10(make_uint4((uint32_t)attempt, (uint32_t)(attempt >> 32), tid, STREAM_SGD),
                                  make_uint2(seed_lo, seed_hi));
    Slot t;
    PathLookup pl;
    pl.fs = g.first_step; pl.blk = nullptr; pl.shift = 0; pl.P = g.P;
    sample_s1(g, pl, ep, r, g.samp_base, g.samp_len, true, 0u, 0, t);      // g.coherent == 0 here
    sample_s2(g, ep, t);
    t.a = load_rec(g.recs + t.step_a);
    t.b = load_rec(g.recs + t.step_b);
    if (ND) sample_ends(t);
    const double d = term_distance(t);
    const bool ok = t.valid && d != 0.0;
    valid[k] = ok;
    step_a[k] = ok ? t.step_a : 0;
    step_b[k] = ok ? t.step_b : 0;
    flags[k] = ok ? (uint8_t)((t.other_a ? 1 : 0) | (t.other_b ? 2 : 0)) : 0;
    dist[k] = ok ? d : 0.0;
}

}  // namespace gfs

// =============================================================================================
// Host-side objects
// =============================================================================================
using namespace gfs;

struct gfs_index {
    int device = 0;
    uint64_t S = 0, P = 0, N = 0;
    uint64_t max_path_steps = 0;
    bool any_multi_step = false;
    StepRec* d_recs = nullptr;
    uint64_t* d_first_step = nullptr;   // P+1
    uint64_t* d_path_len = nullptr;     // P
    uint32_t* d_new_of_old = nullptr;   // N, null when not relabelled
    uint32_t* d_old_of_new = nullptr;   // N
    std::vector<uint64_t> h_first_step;
    double build_seconds = 0, h2d_seconds = 0;
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
    double* d_zetas = nullptr; uint32_t zlen = 0;
    EpochDesc* d_epochs = nullptr; uint32_t n_epochs = 0;
    uint64_t* d_attempts = nullptr;
    unsigned long long* d_counters = nullptr;
    double* d_stage = nullptr;  // f64 staging for nD conversions
    size_t smem_bytes = 0;
    uint64_t rng_thread_base = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double kernel_ms = 0.0;
    uint64_t launches = 0;
    double h2d_s = 0, d2h_s = 0;
    bool ev_pending = false;
    int inflight = 2;           // terms in flight per thread (kernel template parameter K)
    bool coherent = true;       // warp-coherent step sampling in the sweep schedule
    uint64_t window_steps = 0;  // 0 = static schedule
    uint32_t chunk_updates = 128;
    unsigned long long* d_work = nullptr;
    void* d_saved = nullptr;    // gfs_sgd_session_save snapshot of the positions
    uint64_t samp_base = 0, samp_len = 0;
};

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

// Relabel the nodes of a built index.  given == nullptr: order of first appearance in this index's
// steps (never-visited nodes last); else the caller's permutation (multi-GPU: one for all ranks).
static int index_relabel(gfs_index* ix, const uint32_t* given, cudaStream_t st) {
    if (ix->N == 0) return GFS_OK;
    const uint32_t N = (uint32_t)ix->N;
    GFS_CUDA(cudaMalloc(&ix->d_new_of_old, (size_t)N * 4));
    GFS_CUDA(cudaMalloc(&ix->d_old_of_new, (size_t)N * 4));
    if (given) {
        GFS_CUDA(cudaMemcpyAsync(ix->d_new_of_old, given, (size_t)N * 4, cudaMemcpyHostToDevice, st));
        rl_invert<<<(N + 255) / 256, 256, 0, st>>>(ix->d_new_of_old, N, ix->d_old_of_new);
    } else {
        unsigned long long* d_first = nullptr; uint64_t* d_tiles = nullptr; uint64_t* d_carry = nullptr;
        const uint64_t tiles_s = (ix->S + K1_TILE - 1) / K1_TILE, tiles_n = ((uint64_t)N + K1_TILE - 1) / K1_TILE;
        GFS_CUDA(cudaMalloc(&d_first, (size_t)N * 8));
        GFS_CUDA(cudaMalloc(&d_tiles, (std::max(tiles_s, tiles_n) + 1) * 8));
        GFS_CUDA(cudaMalloc(&d_carry, 8));
        GFS_CUDA(cudaMemsetAsync(d_first, 0xff, (size_t)N * 8, st));
        GFS_CUDA(cudaMemsetAsync(d_carry, 0, 8, st));
        if (ix->S) {
            rl_first_occ<<<(unsigned)((ix->S + 255) / 256), 256, 0, st>>>(ix->d_recs, ix->S, N, d_first);
            rl_tile_count<0><<<(unsigned)tiles_s, K1_THREADS, 0, st>>>(ix->d_recs, d_first, N, ix->S, d_tiles);
            k1_scan_tiles<<<1, 1024, 0, st>>>(d_tiles, tiles_s, d_carry);
            rl_assign<0><<<(unsigned)tiles_s, K1_THREADS, 0, st>>>(ix->d_recs, d_first, N, ix->S, d_tiles, ix->d_new_of_old, ix->d_old_of_new);
        }
        rl_tile_count<1><<<(unsigned)tiles_n, K1_THREADS, 0, st>>>(nullptr, d_first, N, N, d_tiles);
        k1_scan_tiles<<<1, 1024, 0, st>>>(d_tiles, tiles_n, d_carry);      // carry continues after the visited nodes
        rl_assign<1><<<(unsigned)tiles_n, K1_THREADS, 0, st>>>(nullptr, d_first, N, N, d_tiles, ix->d_new_of_old, ix->d_old_of_new);
        cudaError_t e = cudaStreamSynchronize(st);
        cudaFree(d_first); cudaFree(d_tiles); cudaFree(d_carry);
        if (e != cudaSuccess) { set_error(std::string("relabel failed: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
    }
    if (ix->S) rl_rewrite<<<(unsigned)((ix->S + 255) / 256), 256, 0, st>>>(ix->d_recs, ix->S, N, ix->d_new_of_old);
    GFS_CUDA(cudaStreamSynchronize(st));
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}

extern "C" int gfs_index_build_shard(const uint64_t* step_handles, const uint64_t* path_first_step,
                                     const uint32_t* node_len, uint64_t S, uint64_t P, uint64_t N,
                                     uint64_t path_begin, uint64_t path_end, int32_t device, int32_t relabel_mode,
                                     const uint32_t* new_of_old, gfs_index** out) {
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
    int rc = select_device(device);
    if (rc) return rc;

    const double t_begin = now_s();
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

    cudaStream_t st;
    IX_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint64_t* d_path_base = nullptr; uint64_t* d_carry = nullptr; uint32_t* d_node_len = nullptr;
    uint64_t* d_handles = nullptr; uint64_t* d_tiles = nullptr;
    IX_CUDA(cudaMalloc(&ix->d_first_step, (ix->P + 1) * 8));
    IX_CUDA(cudaMalloc(&ix->d_path_len, std::max<uint64_t>(ix->P, 1) * 8));
    IX_CUDA(cudaMalloc(&ix->d_recs, std::max<uint64_t>(ix->S, 1) * sizeof(StepRec)));
    IX_CUDA(cudaMalloc(&d_path_base, (ix->P + 1) * 8));
    IX_CUDA(cudaMalloc(&d_carry, 8));
    IX_CUDA(cudaMalloc(&d_node_len, std::max<uint64_t>(N, 1) * 4));
    IX_CUDA(cudaMemsetAsync(d_carry, 0, 8, st));
    IX_CUDA(cudaMemsetAsync(d_path_base, 0, (ix->P + 1) * 8, st));
    const double t_h2d0 = now_s();
    IX_CUDA(cudaMemcpyAsync(ix->d_first_step, ix->h_first_step.data(), (ix->P + 1) * 8, cudaMemcpyHostToDevice, st));
    if (N) IX_CUDA(cudaMemcpyAsync(d_node_len, node_len, N * 4, cudaMemcpyHostToDevice, st));
    double h2d = now_s() - t_h2d0;

    // chunks of at most CH steps (multiple of the tile) so that the transient handle buffer stays small
    uint64_t CH = (uint64_t)env_long("GFASORT_INDEX_CHUNK", 1l << 28);
    CH = std::max<uint64_t>(K1_TILE, (CH / K1_TILE) * K1_TILE);
    const uint64_t chunk_cap = std::min<uint64_t>(CH, std::max<uint64_t>(ix->S, 1));
    const uint64_t tiles_cap = (chunk_cap + K1_TILE - 1) / K1_TILE;
    IX_CUDA(cudaMalloc(&d_handles, chunk_cap * 8));
    IX_CUDA(cudaMalloc(&d_tiles, (tiles_cap + 1) * 8));
    uint32_t p_lo = 0;
    for (uint64_t c0 = 0; c0 < ix->S; c0 += CH) {
        const uint64_t clen = std::min(CH, ix->S - c0);
        const uint64_t n_tiles = (clen + K1_TILE - 1) / K1_TILE;
        const double t0 = now_s();
        IX_CUDA(cudaMemcpyAsync(d_handles, step_handles + s_begin + c0, clen * 8, cudaMemcpyHostToDevice, st));
        IX_CUDA(cudaStreamSynchronize(st));
        h2d += now_s() - t0;
        k1_tile_sums<<<(unsigned)n_tiles, K1_THREADS, 0, st>>>(d_handles, d_node_len, clen, N, d_tiles);
        k1_scan_tiles<<<1, 1024, 0, st>>>(d_tiles, n_tiles, d_carry);
        // paths whose first step lies in this chunk: [p_lo, p_hi)
        uint32_t p_hi = p_lo;
        while (p_hi < ix->P && ix->h_first_step[p_hi] < c0 + clen) ++p_hi;
        if (p_hi > p_lo)
            k1_path_base<<<p_hi - p_lo, K1_THREADS, 0, st>>>(d_handles, d_node_len, N, ix->d_first_step, p_lo, p_hi, c0,
                                                            c0 + clen, d_tiles, d_path_base);
        k1_write_recs<<<(unsigned)n_tiles, K1_THREADS, 0, st>>>(d_handles, d_node_len, N, ix->d_first_step, (uint32_t)ix->P,
                                                               c0, clen, d_tiles, d_path_base, ix->d_recs);
        IX_CUDA(cudaGetLastError());
        p_lo = p_hi;
    }
    // paths that start at S (empty tail paths) and the sentinel base[P] = total
    if (ix->P) {
        // every empty path at the very end has first_step == S: base = total
        for (uint32_t p = p_lo; p <= ix->P; ++p) k1_set_u64<<<1, 1, 0, st>>>(d_path_base, p, d_carry);
        k1_path_len<<<(unsigned)((ix->P + 255) / 256), 256, 0, st>>>(d_path_base, (uint32_t)ix->P, ix->d_path_len);
    }
    IX_CUDA(cudaStreamSynchronize(st));
    IX_CUDA(cudaGetLastError());
    cudaFree(d_handles); cudaFree(d_tiles); cudaFree(d_path_base); cudaFree(d_carry); cudaFree(d_node_len);
    if (relabel_mode == 2 && !new_of_old) { set_error("gfs_index_build: relabel_mode 2 needs a permutation"); return fail(GFS_ERR_INVALID); }
    if (relabel_mode == 1 || relabel_mode == 2) {
        rc = index_relabel(ix, relabel_mode == 2 ? new_of_old : nullptr, st);
        if (rc) return fail(rc);
    }
    cudaStreamDestroy(st);
    ix->h2d_seconds = h2d;
    ix->build_seconds = now_s() - t_begin;
    *out = ix;
    return GFS_OK;
#undef IX_CUDA
}

extern "C" int gfs_index_build(const uint64_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                               uint64_t S, uint64_t P, uint64_t N, gfs_index** out) {
    return gfs_index_build_shard(step_handles, path_first_step, node_len, S, P, N, 0, P, -1,
                                 env_long("GFASORT_RELABEL", 1) ? 1 : 0, nullptr, out);
}

extern "C" int gfs_index_export_relabel(const gfs_index* ix, uint32_t* new_of_old) {
    if (!ix || !new_of_old) { set_error("gfs_index_export_relabel: null argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (ix->d_new_of_old) GFS_CUDA(cudaMemcpy(new_of_old, ix->d_new_of_old, ix->N * 4, cudaMemcpyDeviceToHost));
    else for (uint64_t i = 0; i < ix->N; ++i) new_of_old[i] = (uint32_t)i;
    return GFS_OK;
}

extern "C" void gfs_index_free(gfs_index* ix) {
    if (!ix) return;
    cudaSetDevice(ix->device);
    cudaFree(ix->d_recs); cudaFree(ix->d_first_step); cudaFree(ix->d_path_len);
    cudaFree(ix->d_new_of_old); cudaFree(ix->d_old_of_new);
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

extern "C" int gfs_index_export(const gfs_index* ix, uint64_t* step_pos, uint64_t* path_len) {
    if (!ix) { set_error("gfs_index_export: null index"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (step_pos && ix->S) {
        uint64_t* d = nullptr;
        const uint64_t CH = 1ull << 26;
        GFS_CUDA(cudaMalloc(&d, std::min(CH, ix->S) * 8));
        for (uint64_t c0 = 0; c0 < ix->S; c0 += CH) {
            const uint64_t n = std::min(CH, ix->S - c0);
            k1_export_pos<<<(unsigned)((n + 255) / 256), 256>>>(ix->d_recs + c0, n, d);
            cudaError_t e = cudaMemcpy(step_pos + c0, d, n * 8, cudaMemcpyDeviceToHost);
            if (e != cudaSuccess) { cudaFree(d); set_error(std::string("export copy failed: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
        }
        cudaFree(d);
    }
    if (path_len && ix->P) GFS_CUDA(cudaMemcpy(path_len, ix->d_path_len, ix->P * 8, cudaMemcpyDeviceToHost));
    return GFS_OK;
}

extern "C" int gfs_index_export_records(const gfs_index* ix, uint64_t* step_handle, uint32_t* step_node_len) {
    if (!ix || !step_handle || !step_node_len) { set_error("gfs_index_export_records: null argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (!ix->S) return GFS_OK;
    uint64_t* dh = nullptr; uint32_t* dl = nullptr;
    const uint64_t CH = 1ull << 26;
    const uint64_t cap = std::min(CH, ix->S);
    GFS_CUDA(cudaMalloc(&dh, cap * 8));
    GFS_CUDA(cudaMalloc(&dl, cap * 4));
    for (uint64_t c0 = 0; c0 < ix->S; c0 += CH) {
        const uint64_t n = std::min(CH, ix->S - c0);
        k1_export_hl<<<(unsigned)((n + 255) / 256), 256>>>(ix->d_recs + c0, n, (uint32_t)ix->N, ix->d_old_of_new, dh, dl);
        cudaError_t e1 = cudaMemcpy(step_handle + c0, dh, n * 8, cudaMemcpyDeviceToHost);
        cudaError_t e2 = cudaMemcpy(step_node_len + c0, dl, n * 4, cudaMemcpyDeviceToHost);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { cudaFree(dh); cudaFree(dl); set_error("export copy failed"); return GFS_ERR_CUDA; }
    }
    cudaFree(dh); cudaFree(dl);
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

static KernelGraph make_kgraph(const gfs_index* ix, const gfs_sgd_params& p, const double* d_zetas, uint32_t zlen) {
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
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" int gfs_sgd_session_create(const gfs_index* ix, const gfs_sgd_params* params, uint32_t dims,
                                      const gfs_launch_cfg* cfg, gfs_sgd_session** out) {
    if (!out) { set_error("gfs_sgd_session_create: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    if (!ix) { set_error("gfs_sgd_session_create: null index"); return GFS_ERR_INVALID; }
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
    SS_CUDA(cudaEventCreate(&s->ev0));
    SS_CUDA(cudaEventCreate(&s->ev1));

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
    std::vector<double> zetas; h_zetas(*params, ix->max_path_steps, zetas);
    s->n_epochs = (uint32_t)epochs.size(); s->zlen = (uint32_t)zetas.size();
    SS_CUDA(cudaMalloc(&s->d_epochs, epochs.size() * sizeof(EpochDesc)));
    SS_CUDA(cudaMalloc(&s->d_zetas, std::max<size_t>(zetas.size(), 1) * 8));
    SS_CUDA(cudaMemcpyAsync(s->d_epochs, epochs.data(), epochs.size() * sizeof(EpochDesc), cudaMemcpyHostToDevice, s->stream));
    SS_CUDA(cudaMemcpyAsync(s->d_zetas, zetas.data(), zetas.size() * 8, cudaMemcpyHostToDevice, s->stream));
    SS_CUDA(cudaStreamSynchronize(s->stream));   // host vectors go out of scope

    // launch shape: persistent, every block co-resident
    const uint32_t n_fs = (ix->P + 1 <= SMEM_FS_MAX) ? (uint32_t)ix->P + 1 : 0;
    s->smem_bytes = n_fs ? (size_t)n_fs * 8 + (size_t)BLK_TABLE * 2 : 0;
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
        if ((uint64_t)w >= s->samp_len) w = 0;
        s->window_steps = w > 0 ? std::max<uint64_t>((uint64_t)w, 1024) : 0;
        s->chunk_updates = (uint32_t)std::max<long>(1, env_long("GFASORT_CHUNK", 256));
        s->coherent = env_long("GFASORT_COHERENT", 1) != 0;
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

static int session_flush_events(gfs_sgd_session* s) {
    if (s->ev_pending) {
        GFS_CUDA(cudaEventSynchronize(s->ev1));
        float ms = 0;
        GFS_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        s->kernel_ms += ms;
        s->ev_pending = false;
    }
    return GFS_OK;
}

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
    int rc = session_flush_events(s);
    if (rc) return rc;
    uint32_t DS;
    sgd_kernel_fn fn = pick_kernel(s->dims, s->f64, s->aggregate, s->inflight, DS);
    SgdArgs a{};
    a.g = make_kgraph(s->ix, s->params, s->d_zetas, s->zlen);
    a.g.samp_base = s->samp_base; a.g.samp_len = s->samp_len;
    a.g.coherent = (s->window_steps > 0 && s->coherent) ? 1u : 0u;
    a.epochs = s->d_epochs;
    a.epoch_begin = (uint32_t)epoch_begin; a.epoch_end = (uint32_t)epoch_end;
    a.slice = slice; a.n_slices = n_slices;
    a.attempt_ctr = s->d_attempts; a.counters = s->d_counters;
    a.seed_lo = (uint32_t)s->params.seed; a.seed_hi = (uint32_t)(s->params.seed >> 32);
    a.tid_base = (uint32_t)s->rng_thread_base;
    a.positions = s->d_pos;
    a.window_steps = s->window_steps; a.chunk_updates = s->chunk_updates;
    a.work_ctr = s->d_work;
    {
        // generous bound on loop iterations per warp: 64x its fair share of the launch's attempts + slack
        const uint64_t m = s->params.min_term_updates / n_slices + 1;
        const uint64_t T = (uint64_t)s->grid * s->block;
        a.iter_cap = ((m / T + 1) * (epoch_end - epoch_begin)) * 64 + (1ull << 16);
    }
    GFS_CUDA(cudaMemsetAsync(s->d_work, 0, 8, s->stream));
    void* kargs[] = {(void*)&a};
    GFS_CUDA(cudaEventRecord(s->ev0, s->stream));
    GFS_CUDA(cudaLaunchCooperativeKernel((const void*)fn, dim3(s->grid), dim3(s->block), kargs, s->smem_bytes, s->stream));
    GFS_CUDA(cudaEventRecord(s->ev1, s->stream));
    s->ev_pending = true;
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
    return GFS_OK;
}

static int run_whole(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg, uint32_t dims,
                     double* pos_inout, gfs_stats* stats) {
    if (!pos_inout) { set_error("positions buffer is null"); return GFS_ERR_INVALID; }
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
extern "C" int gfs_stress(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords,
                          uint64_t samples, uint64_t seed, double* rms_rel, double* mean_abs_rel, uint64_t* counted) {
    if (!ix || !coords || dims < 1) { set_error("gfs_stress: bad argument"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(ix->device));
    if (rms_rel) *rms_rel = 0;
    if (mean_abs_rel) *mean_abs_rel = 0;
    if (counted) *counted = 0;
    if (ix->S < 2 || samples == 0) return GFS_OK;            // sgd.rs:1220-1222
    const uint32_t stride = layout_order ? 2 * dims : dims;
    double* d_coords = nullptr; double* d_partial = nullptr;
    const size_t bytes = (size_t)ix->N * stride * 8;
    GFS_CUDA(cudaMalloc(&d_coords, bytes));
    cudaError_t e = cudaMemcpy(d_coords, coords, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(d_coords); set_error(std::string("gfs_stress copy: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
    const unsigned grid = (unsigned)std::min<uint64_t>((samples + STRESS_BLOCK - 1) / STRESS_BLOCK, 148 * 8);
    e = cudaMalloc(&d_partial, (size_t)grid * 3 * 8);
    if (e != cudaSuccess) { cudaFree(d_coords); set_error("gfs_stress alloc"); return GFS_ERR_CUDA; }
    gfs_sgd_params dummy{};
    KernelGraph g = make_kgraph(ix, dummy, nullptr, 0);
    stress_kernel<<<grid, STRESS_BLOCK>>>(g, ix->d_old_of_new, d_coords, dims, stride, samples, (uint32_t)seed, (uint32_t)(seed >> 32), d_partial);
    std::vector<double> part((size_t)grid * 3);
    e = cudaMemcpy(part.data(), d_partial, part.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_coords); cudaFree(d_partial);
    if (e != cudaSuccess) { set_error(std::string("gfs_stress kernel: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
    double s0 = 0, s1 = 0, c = 0;
    for (unsigned b = 0; b < grid; ++b) { s0 += part[b * 3]; s1 += part[b * 3 + 1]; c += part[b * 3 + 2]; }
    if (c > 0) {
        if (rms_rel) *rms_rel = std::sqrt(s0 / c);
        if (mean_abs_rel) *mean_abs_rel = s1 / c;
    }
    if (counted) *counted = (uint64_t)c;
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// debug / parity hooks
// ---------------------------------------------------------------------------------------------
template <typename T> struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
    cudaError_t up(const T* h, size_t n) { return cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T* h, size_t n) { return cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};

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
    std::vector<double> zetas; h_zetas(*params, ix->max_path_steps, zetas);
    DevBuf<EpochDesc> de; DevBuf<double> dz, dd; DevBuf<uint8_t> dv, df; DevBuf<uint64_t> da, db;
    GFS_CUDA(de.alloc(epochs.size())); GFS_CUDA(dz.alloc(zetas.size()));
    GFS_CUDA(de.up(epochs.data(), epochs.size())); GFS_CUDA(dz.up(zetas.data(), zetas.size()));
    GFS_CUDA(dv.alloc(count)); GFS_CUDA(df.alloc(count)); GFS_CUDA(da.alloc(count)); GFS_CUDA(db.alloc(count)); GFS_CUDA(dd.alloc(count));
    KernelGraph g = make_kgraph(ix, *params, dz.p, (uint32_t)zetas.size());
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
extern "C" int gfs_reconcile_pack(const void* x, const void* x_sync, uint64_t n, uint32_t elem_bytes, float* buf, void* stream) {
    if (!x || !x_sync || !buf || (elem_bytes != 4 && elem_bytes != 8)) { set_error("gfs_reconcile_pack: bad argument"); return GFS_ERR_INVALID; }
    if (n == 0) return GFS_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    if (elem_bytes == 8) rc_pack<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)x, (const double*)x_sync, n, buf);
    else rc_pack<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)x_sync, n, buf);
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}
extern "C" int gfs_reconcile_apply(void* x, void* x_sync, uint64_t n, uint32_t elem_bytes, const float* buf, void* stream) {
    if (!x || !x_sync || !buf || (elem_bytes != 4 && elem_bytes != 8)) { set_error("gfs_reconcile_apply: bad argument"); return GFS_ERR_INVALID; }
    if (n == 0) return GFS_OK;
    const unsigned grid = (unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 16);
    if (elem_bytes == 8) rc_apply<double><<<grid, 256, 0, (cudaStream_t)stream>>>((double*)x, (double*)x_sync, n, buf);
    else rc_apply<float><<<grid, 256, 0, (cudaStream_t)stream>>>((float*)x, (float*)x_sync, n, buf);
    GFS_CUDA(cudaGetLastError());
    return GFS_OK;
}

// ---------------------------------------------------------------------------------------------
// order by position
// ---------------------------------------------------------------------------------------------
// d_x: n positions on the device, in the caller's node order.  d_order: n dense indices.  Asynchronous on st.
static int sort_positions_device(const double* d_x, uint64_t n, uint32_t* d_order, cudaStream_t st, uint64_t* launches) {
    if (n == 0) return GFS_OK;
    if (n >= (1ull << 32)) { set_error("sort: n must be < 2^32"); return GFS_ERR_INVALID; }
    const uint32_t n_blocks = (uint32_t)((n + RS_TILE - 1) / RS_TILE);
    uint64_t *k0 = nullptr, *k1 = nullptr; uint32_t *v1 = nullptr, *hist = nullptr;
    GFS_CUDA(cudaMalloc(&k0, n * 8)); GFS_CUDA(cudaMalloc(&k1, n * 8));
    GFS_CUDA(cudaMalloc(&v1, n * 4)); GFS_CUDA(cudaMalloc(&hist, (size_t)256 * n_blocks * 4));
    uint32_t* v0 = d_order;
    rs_make_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_x, n, k0, v0);
    uint64_t nl = 1;
    for (int pass = 0; pass < 8; ++pass) {
        rs_hist<<<n_blocks, RS_THREADS, 0, st>>>(k0, n, pass * 8, n_blocks, hist);
        rs_scan<<<1, 1024, 0, st>>>(hist, (uint64_t)256 * n_blocks);
        rs_scatter<<<n_blocks, RS_THREADS, 0, st>>>(k0, v0, n, pass * 8, n_blocks, hist, k1, v1);
        std::swap(k0, k1); std::swap(v0, v1);
        nl += 3;
    }
    // 8 passes: the result is back in the buffers it started in (v0 == d_order)
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(k0); cudaFree(k1); cudaFree(v1 == d_order ? v0 : v1); cudaFree(hist);
    if (e != cudaSuccess) { set_error(std::string("sort failed: ") + cudaGetErrorString(e)); return GFS_ERR_CUDA; }
    if (launches) *launches += nl;
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
