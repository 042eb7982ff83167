// gfs_kernels_index.cuh — K1 path index kernel and the node relabelling kernels (reference src/sgd.rs:34-71).
// Part of libgfasort_cuda.so; included by gfs_lib.cu (one translation unit).  See DESIGN.md §4.
#pragma once
#include "gfs_device.cuh"

namespace gfs {

// =============================================================================================
// K1 — path index: ONE pass over the steps (chained scan with decoupled look-back)
// =============================================================================================
// PathIndex::from_graph (src/sgd.rs:41-62) is, per path, an exclusive prefix sum of node lengths along
// the steps.  Over the concatenated step array that is a SEGMENTED exclusive scan with a segment start at
// every path's first step.  k1_scan_write does it in one pass: every tile of 2048 steps
//   loads its handles once (coalesced), gathers {node length, first-occurrence key} (one 8-byte entry per
//   node, L2-resident), scans the lengths in shared memory, publishes its aggregate in a per-tile
//   descriptor, looks back over its predecessors' descriptors for its exclusive prefix (Merrill & Garland's
//   decoupled look-back, with the segmented twist: a tile that contains a path start publishes its
//   INCLUSIVE value at once — the scan restarts inside it, so nothing before it matters to its successors),
//   and writes one 16-byte record per step {node<<1|rev, node_len, offset}.
// Fused into the same pass: the per-path lengths (PathInfo.length, sgd.rs:29) and, when the index is going
// to be relabelled, the first-occurrence key of every node (atomicMin, almost always skipped after the
// first path thanks to the key that came with the length gather).
// Algorithmic bytes per step: 8 (handle; 4 with 32-bit handles) + 4 (gathered length) + 16 (record).
constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 8;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;

// per-node table the kernel gathers from: node length and the node's first-occurrence key
// (step index >> key_shift of the first step that visits it; 0xffffffff = not visited yet)
struct __align__(8) NodeEnt { uint32_t len; uint32_t key; };

constexpr uint64_t K1_ST_AGG = 1ull << 62;      // descriptor holds the tile's sum; the tile has no path start
constexpr uint64_t K1_ST_INCL = 2ull << 62;     // descriptor holds the offset, in its path, of the step after the tile
constexpr uint64_t K1_VAL_MASK = (1ull << 62) - 1;
constexpr uint32_t K1_SPIN_CAP = 1u << 27;      // look-back polls before the kernel gives up (ticket[1] = 1: the build fails)

__device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ NodeEnt ld_node_ent(const NodeEnt* p) {
    NodeEnt e;
    asm volatile("ld.global.cg.v2.u32 {%0,%1}, [%2];" : "=r"(e.len), "=r"(e.key) : "l"(p));
    return e;
}

__global__ void k1_init_table(const uint32_t* __restrict__ node_len, uint32_t N, NodeEnt* __restrict__ tbl) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { NodeEnt e; e.len = node_len[i]; e.key = 0xffffffffu; tbl[i] = e; }
}

// shared-memory views: loads and stores are strided (thread t owns items k*256 + t: coalesced handle loads,
// one 16-byte store per record), the scan is blocked (thread t sums items [8t, 8t+8)); padded by one word per
// 32 (lengths) / one entry per 16 (offsets) so that neither view has bank conflicts.
__device__ __forceinline__ int k1_pad32(int j) { return j + (j >> 5); }
__device__ __forceinline__ int k1_pad16(int j) { return j + (j >> 4); }

// HT: handle type of the caller's step array (uint64_t = Handle as the reference stores it; uint32_t = the
// same value in 32 bits, dense idx < 2^31).  FIRST_OCC: maintain the first-occurrence keys.
// handles: this chunk's steps; chunk_begin (a multiple of K1_TILE): index-local step of handles[0].
// ticket[0]: tile counter of this launch (zeroed by the host before it); ticket[1]: look-back watchdog flag.
template <typename HT, bool FIRST_OCC>
__global__ void __launch_bounds__(K1_THREADS, 4)
k1_scan_write(const HT* __restrict__ handles, NodeEnt* __restrict__ tbl, uint32_t N, const uint64_t* __restrict__ first_step,
              uint32_t P, uint64_t chunk_begin, uint64_t chunk_len, uint64_t* __restrict__ desc, unsigned int* __restrict__ ticket,
              uint32_t key_shift, StepRec* __restrict__ recs, uint64_t* __restrict__ path_len) {
    __shared__ uint32_t s_len[K1_TILE + K1_TILE / 32];
    __shared__ uint64_t s_pos[K1_TILE + K1_TILE / 16];
    __shared__ uint64_t wsum[K1_THREADS / 32];
    __shared__ uint64_t s_prefix;
    __shared__ unsigned int s_tile;
    // tiles are handed out in order, so every predecessor of a running tile is running or done: the
    // look-back below can never wait for a block that has not been scheduled
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint64_t tbase = (uint64_t)s_tile * K1_TILE;              // chunk-local
    if (tbase >= chunk_len) return;
    const uint64_t gtile = (chunk_begin + tbase) / K1_TILE;
    uint64_t hh[K1_ITEMS];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t i = tbase + (uint64_t)(k * K1_THREADS + threadIdx.x);
        hh[k] = i < chunk_len ? (uint64_t)__ldg(handles + i) : ~0ull;      // past the end: node >= N => length 0
    }
    // path of the tile's first and last step (most tiles lie inside one path), found while the handle
    // loads are in flight: with few paths every thread tests one first_step entry and the block counts
    // the entries <= step (one load round trip); with many paths, a binary search.
    const uint64_t g_first = chunk_begin + tbase;
    const uint32_t n_here = (uint32_t)(tbase + K1_TILE <= chunk_len ? K1_TILE : chunk_len - tbase);
    const uint64_t g_last = g_first + n_here - 1;
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
    const uint64_t fs_first = __ldg(first_step + p_first);
    const bool start_first = fs_first == g_first;                   // the tile begins exactly at a path start
    const bool has_start = start_first || p_last != p_first;
    uint32_t len[K1_ITEMS], nr[K1_ITEMS];
    NodeEnt ent[K1_ITEMS];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {                            // all eight gathers in flight before the first use
        const uint64_t node = hh[k] >> 1;
        ent[k].len = 0; ent[k].key = 0;
        if (node < N) ent[k] = ld_node_ent(tbl + node);             // missing node => +0 (src/sgd.rs:52-54)
    }
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t node = hh[k] >> 1;
        len[k] = ent[k].len;
        nr[k] = (uint32_t)(((node < N ? node : N) << 1) | (hh[k] & 1));
        s_len[k1_pad32(k * K1_THREADS + threadIdx.x)] = ent[k].len;
        if (FIRST_OCC && node < N) {
            const uint64_t gi = g_first + (uint64_t)(k * K1_THREADS + threadIdx.x);
            const uint32_t key = (uint32_t)(gi >> key_shift);
            if (key < ent[k].key) atomicMin(&tbl[node].key, key);
        }
    }
    __syncthreads();
    // blocked view: thread t owns items [t*8, t*8+8)
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
    uint64_t woff = 0, tile_sum = 0;
#pragma unroll
    for (int k = 0; k < K1_THREADS / 32; ++k) { const uint64_t v = wsum[k]; woff += (k < w) ? v : 0; tile_sum += v; }
    const uint64_t texcl = woff + (inc - tsum);                     // tile-local exclusive prefix of item t*8
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) s_pos[k1_pad16(threadIdx.x * K1_ITEMS + k)] = texcl + loc[k];
    __syncthreads();
    // ---- publish, look back ---------------------------------------------------------------------
    if (w == 0) {
        if (has_start) {
            // the scan restarts at the tile's last path start: what follows the tile depends on nothing before it
            if (lane == 0) {
                const uint32_t j_ls = (uint32_t)(__ldg(first_step + p_last) - g_first);     // p_last's first step is in the tile
                st_desc(desc + gtile, K1_ST_INCL | (tile_sum - s_pos[k1_pad16((int)j_ls)]));
            }
        } else if (lane == 0) {
            st_desc(desc + gtile, K1_ST_AGG | tile_sum);
        }
        uint64_t T = 0;
        if (!start_first) {
            int64_t look = (int64_t)gtile - 1;
            for (;;) {
                const int64_t t = look - lane;
                uint64_t d = K1_ST_INCL;                            // before the first tile: offset 0
                if (t >= 0) {
                    uint32_t spins = 0;
                    do {
                        d = ld_desc(desc + t);
                        if (++spins > K1_SPIN_CAP) { atomicExch(ticket + 1, 1u); d = K1_ST_INCL; break; }   // watchdog: never a hang
                    } while ((d >> 62) == 0);
                }
                const unsigned incl = __ballot_sync(0xffffffffu, (d >> 62) == 2);
                const int stop = incl ? __ffs(incl) - 1 : 32;       // nearest predecessor with an inclusive value
                uint64_t v = lane <= stop ? (d & K1_VAL_MASK) : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                T += v;
                if (incl) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            if (!has_start) st_desc(desc + gtile, K1_ST_INCL | (T + tile_sum));
            s_prefix = T;
        }
    }
    __syncthreads();
    const uint64_t T = s_prefix;
    const uint64_t fs_next = __ldg(first_step + p_first + 1);       // single-path tiles: where the path ends
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const int j = k * K1_THREADS + threadIdx.x;
        if (j < (int)n_here) {
            const uint64_t gi = g_first + j;
            uint64_t pos, next_first;
            uint32_t q = p_first;
            if (p_first == p_last || gi < fs_next) {
                pos = T + s_pos[k1_pad16(j)];                        // T == 0 when the tile begins at a path start
                next_first = fs_next;
            } else {
                q = find_path(first_step, P, gi);
                const uint64_t fq = first_step[q];
                pos = s_pos[k1_pad16(j)] - s_pos[k1_pad16((int)(fq - g_first))];
                next_first = first_step[q + 1];
            }
            // StepRec {node_rev, node_len, pos} as one 16-byte store (see load_rec)
            *reinterpret_cast<uint4*>(recs + gi) = make_uint4(nr[k], len[k], (uint32_t)pos, (uint32_t)(pos >> 32));
            if (gi + 1 == next_first) path_len[q] = pos + len[k];   // the path's last step: PathInfo.length (sgd.rs:64-68)
        }
    }
}

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
// new_of_old, downloads gather back; no arithmetic changes.  The order comes from a stable radix sort
// (rs_* kernels, gfs_kernels_aux.cuh) of the N first-occurrence keys K1 left in the node table:
// never-visited nodes (key 0xffffffff) sort last, ties (key_shift > 0) by dense idx.
// ---------------------------------------------------------------------------------------------
__global__ void rl_keys(const NodeEnt* __restrict__ tbl, uint32_t N, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { keys[i] = tbl[i].key; vals[i] = i; }
}
__global__ void rl_rewrite(StepRec* __restrict__ recs, uint64_t S, uint32_t N, const uint32_t* __restrict__ new_of_old) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S) return;
    const uint32_t nr = recs[i].node_rev;
    const uint32_t node = nr >> 1;
    if (node < N) recs[i].node_rev = (__ldg(new_of_old + node) << 1) | (nr & 1u);
}
__global__ void rl_invert(const uint32_t* __restrict__ perm, uint32_t N, uint32_t* __restrict__ inverse) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) inverse[perm[i]] = i;
}

}  // namespace gfs
