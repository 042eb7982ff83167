// gfs_kernels_aux.cuh — K4 sampled stress, layout conversions, K6 radix sort, K5 reconcile, debug kernels.
// Part of libgfasort_cuda.so; included by gfs_lib.cu (one translation unit).  See DESIGN.md §4.
#pragma once
#include "gfs_device.cuh"
#include "gfs_kernels_sgd.cuh"

namespace gfs {

// =============================================================================================
// K4 — sampled stress (sgd.rs:1196-1283)
// =============================================================================================
constexpr int STRESS_BLOCK = 256;
// coords: stride_node doubles per node, the + end's `dims` coordinates first.  g is one index (the whole graph, or
// one shard whose local step 0 is global step `step_offset`).
__global__ void __launch_bounds__(STRESS_BLOCK)
stress_kernel(KernelGraph g, const uint32_t* __restrict__ old_of_new, const double* __restrict__ coords, uint32_t dims,
              uint32_t stride_node, uint64_t samples, uint32_t seed_lo, uint32_t seed_hi, uint64_t total_steps, uint64_t step_offset,
              uint64_t step_begin, uint64_t step_end, double* __restrict__ partial /*3 per block*/) {
    __shared__ double red[3][STRESS_BLOCK / 32];
    double sum = 0.0, sum_abs = 0.0, cnt = 0.0;
    const uint2 key = make_uint2(seed_lo, seed_hi);
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < samples; k += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4((uint32_t)k, (uint32_t)(k >> 32), 0u, STREAM_STRESS), key);
        // the step is drawn over the whole graph; this index evaluates the samples that land in [step_begin, step_end)
        const uint64_t sg = __umul64hi(((uint64_t)r.y << 32) | r.x, total_steps);
        if (sg < step_begin || sg >= step_end) continue;
        const uint64_t s = sg - step_offset;
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
    pl.fs = g.first_step; pl.blk = nullptr; pl.shift = 0; pl.P = g.P; pl.zs = nullptr; pl.zs_n = 0;
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
