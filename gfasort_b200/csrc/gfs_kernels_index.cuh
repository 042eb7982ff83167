// gfs_kernels_index.cuh — K1 path index kernels and the node relabelling kernels (reference src/sgd.rs:34-71).
// Part of libgfasort_cuda.so; included by gfs_lib.cu (one translation unit).  See DESIGN.md §4.
#pragma once
#include "gfs_device.cuh"

namespace gfs {

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
    uint64_t hh[K1_ITEMS];                  // all handle loads first, then all gathers: two round trips
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
        hh[k] = i < S ? __ldg(handles + i) : ~0ull;              // past the end: node >= N => length 0
    }
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) s += gathered_len(hh[k], node_len, N);
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
// Loads and stores are strided (thread t owns items k*256 + t: coalesced handle loads, one 16-byte
// store per record), the scan is blocked (thread t sums items [8t, 8t+8)); the two views meet in
// shared memory, padded by one word per 32 (lengths) / one entry per 16 (offsets) so that neither
// view has bank conflicts.  All eight handle loads are issued before the first gather, so a thread
// waits for two memory round trips, not sixteen.
__device__ __forceinline__ int k1_pad32(int j) { return j + (j >> 5); }
__device__ __forceinline__ int k1_pad16(int j) { return j + (j >> 4); }

__global__ void __launch_bounds__(K1_THREADS, 4)      // <= 64 registers: 4 blocks per SM
k1_write_recs(const uint64_t* __restrict__ handles /*chunk-local*/, const uint32_t* __restrict__ node_len,
              uint64_t N, const uint64_t* __restrict__ first_step, uint32_t P, uint64_t chunk_begin,
              uint64_t chunk_len, const uint64_t* __restrict__ tile_prefix, const uint64_t* __restrict__ path_base,
              StepRec* __restrict__ recs /*global index*/) {
    __shared__ uint32_t s_len[K1_TILE + K1_TILE / 32];
    __shared__ uint64_t s_pos[K1_TILE + K1_TILE / 16];
    __shared__ uint64_t wsum[K1_THREADS / 32];
    const uint64_t tbase = (uint64_t)blockIdx.x * K1_TILE;
    uint64_t hh[K1_ITEMS];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t i = tbase + (uint64_t)(k * K1_THREADS + threadIdx.x);
        hh[k] = i < chunk_len ? __ldg(handles + i) : ~0ull;      // past the end: node >= N => length 0
    }
    // path of the tile's first and last step (most tiles lie inside one path), found while the handle
    // loads are in flight: with few paths every thread tests one first_step entry and the block counts
    // the entries <= step (one load round trip); with many paths, a binary search.
    const uint64_t g_first = chunk_begin + tbase;
    const uint64_t g_last = chunk_begin + (tbase + K1_TILE <= chunk_len ? tbase + K1_TILE : chunk_len) - 1;
    uint32_t p_first, p_last;
    if (P <= 8 * K1_THREADS) {
        int c_first = 0, c_last = 0;
        for (uint32_t p0 = 0; p0 < P; p0 += K1_THREADS) {          // block-uniform trip count
            const uint32_t p = p0 + threadIdx.x;
            const uint64_t f = p < P ? __ldg(first_step + p) : ~0ull;
            c_first += __syncthreads_count(f <= g_first);
            c_last += __syncthreads_count(f <= g_last);
        }
        p_first = (uint32_t)c_first - 1;                            // first_step[0] = 0 <= step: count >= 1
        p_last = (uint32_t)c_last - 1;
    } else {
        p_first = find_path(first_step, P, g_first);
        p_last = find_path(first_step, P, g_last);
    }
    const uint64_t base_first = __ldg(path_base + p_first);
    const uint64_t tile_pre = __ldg(tile_prefix + blockIdx.x);
    uint32_t len[K1_ITEMS], nr[K1_ITEMS];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t node = hh[k] >> 1;
        len[k] = gathered_len(hh[k], node_len, N);
        nr[k] = (uint32_t)(((node < N ? node : N) << 1) | (hh[k] & 1));
        s_len[k1_pad32(k * K1_THREADS + threadIdx.x)] = len[k];
    }
    __syncthreads();
    // thread t owns items [t*8, t*8+8)
    uint64_t loc[K1_ITEMS];
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) { loc[k] = tsum; tsum += s_len[k1_pad32(threadIdx.x * K1_ITEMS + k)]; }
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
    const uint64_t texcl = tile_pre + woff + (inc - tsum);
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) s_pos[k1_pad16(threadIdx.x * K1_ITEMS + k)] = texcl + loc[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const int j = k * K1_THREADS + threadIdx.x;
        const uint64_t i = tbase + j;
        if (i < chunk_len) {
            const uint64_t gi = chunk_begin + i;
            uint64_t pb = base_first;
            if (p_first != p_last) pb = path_base[find_path(first_step, P, gi)];
            const uint64_t pos = s_pos[k1_pad16(j)] - pb;
            // StepRec {node_rev, node_len, pos} as one 16-byte store (see load_rec)
            *reinterpret_cast<uint4*>(recs + gi) = make_uint4(nr[k], len[k], (uint32_t)pos, (uint32_t)(pos >> 32));
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
    if (MODE == 0) {                        // all node loads first, then all first_occ gathers
        uint32_t node[K1_ITEMS];
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) {
            const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
            node[k] = i < n_items ? (recs[i].node_rev >> 1) : N;          // N: never a first occurrence
        }
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) {
            const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
            c += (node[k] < N && first_occ[node[k]] == i) ? 1 : 0;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) {
            const uint64_t i = base + (uint64_t)k * K1_THREADS + threadIdx.x;
            if (i < n_items) c += rl_flag<MODE>(recs, first_occ, N, i) ? 1 : 0;
        }
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
    uint32_t node[K1_ITEMS];
    uint32_t cnt = 0;
    if (MODE == 0) {                        // all node loads first, then all first_occ gathers
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) node[k] = base + k < n_items ? (recs[base + k].node_rev >> 1) : N;
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) { f[k] = node[k] < N && first_occ[node[k]] == base + k; cnt += f[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < K1_ITEMS; ++k) {
            node[k] = (uint32_t)(base + k);
            f[k] = (base + k < n_items) && rl_flag<MODE>(recs, first_occ, N, base + k);
            cnt += f[k];
        }
    }
    uint32_t total;
    uint32_t off = block_excl_scan_u32(cnt, wsum, &total);
    uint64_t rank = tile_prefix[blockIdx.x] + off;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        if (f[k]) {
            const uint32_t old = node[k];
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

}  // namespace gfs
