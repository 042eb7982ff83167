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
//   loads its handles once (16-byte streaming loads), gathers the node lengths (4 bytes per step from an
//   L2-resident table), scans them in registers and with warp shuffles, publishes its aggregate in a per-tile
//   descriptor, looks back over its predecessors' descriptors for its exclusive prefix (Merrill & Garland's
//   decoupled look-back, with the segmented twist: a tile that contains a path start publishes its
//   INCLUSIVE value at once — the scan restarts inside it, so nothing before it matters to its successors),
//   and writes one 16-byte record per step {node<<1|rev, node_len, offset}.
// Fused into the same pass: the per-path lengths (PathInfo.length, sgd.rs:29) and, when the index is going
// to be relabelled, the first-occurrence key of every node (a visited bitmap gates an atomicMin that almost only
// the first path ever executes).
// Algorithmic bytes per step: 8 (handle; 4 with 32-bit handles) + 4 (gathered length) + 16 (record).
constexpr int K1_THREADS = 256;
constexpr int K1_ITEMS = 8;
constexpr int K1_TILE = K1_THREADS * K1_ITEMS;

// What the kernel gathers per step is ONE 4-byte word per node: bits 0..30 = node length, bit 31 = "visited" (the
// 40 MB table of config 3 stays L2-resident next to the streamed handles and records; the kernel is bound by L2 sector
// throughput — every random gather costs a whole 32-byte sector — so one gather per step, not two).  First
// occurrences: first_key[node] = (step index >> key_shift) of the first step that visits the node, 0xffffffff = never
// visited — an atomicMin issued only while the node's visited bit is still clear, i.e. almost only during the first path.
constexpr uint32_t K1_VISITED = 0x80000000u;
__global__ void k1_init_table(const uint32_t* __restrict__ node_len, uint32_t N, uint32_t* __restrict__ tbl, unsigned int* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint32_t l = node_len[i];
    if (l & K1_VISITED) atomicExch(flags + 2, 1u);           // a node of >= 2^31 bp: not representable next to the flag
    tbl[i] = l & ~K1_VISITED;
}

constexpr uint64_t K1_ST_AGG = 1ull << 62;      // descriptor holds the tile's sum; the tile has no path start
constexpr uint64_t K1_ST_INCL = 2ull << 62;     // descriptor holds the offset, in its path, of the step after the tile
constexpr uint64_t K1_VAL_MASK = (1ull << 62) - 1;
constexpr int K1_LB_WINDOWS = 4;                // look-back: 4 x 32 descriptors per round
constexpr uint32_t K1_SPIN_CAP = 1u << 27;      // look-back polls before the kernel gives up (flags[0] = 1: the build fails)

__device__ __forceinline__ uint64_t ld_desc(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_desc(uint64_t* p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// streamed once: handles in (evict-first), records out (evict-first) — keep L2 for the node-length table
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 v;
    asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_v4(void* p, uint4 v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// warp-level "largest p with first_step[p] <= g" for p in [0, P): every lane tests one entry per round and the warp
// counts them (first_step[0] = 0 <= g, so the count is >= 1).  Few paths: 1-3 rounds, no block barrier.
__device__ __forceinline__ uint32_t warp_find_path(const uint64_t* __restrict__ first_step, uint32_t P, uint64_t g, int lane) {
    if (P > 1024) return find_path(first_step, P, g);
    uint32_t c = 0;
    for (uint32_t p0 = 0; p0 < P; p0 += 32) {                       // warp-uniform trip count
        const uint32_t p = p0 + lane;
        const uint64_t f = p < P ? __ldg(first_step + p) : ~0ull;
        c += __popc(__ballot_sync(0xffffffffu, f <= g));
    }
    return c - 1;
}

template <typename HT> struct K1Load;
template <> struct K1Load<uint64_t> {       // 8 handles = 64 bytes = 4 x LDG.128
    static __device__ __forceinline__ void load(const uint64_t* p, uint64_t (&h)[K1_ITEMS]) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = ld_stream_v4(q + k);
            h[2 * k] = ((uint64_t)v.y << 32) | v.x;
            h[2 * k + 1] = ((uint64_t)v.w << 32) | v.z;
        }
    }
};
template <> struct K1Load<uint32_t> {       // 8 handles = 32 bytes = 2 x LDG.128
    static __device__ __forceinline__ void load(const uint32_t* p, uint64_t (&h)[K1_ITEMS]) {
        const uint4* q = reinterpret_cast<const uint4*>(p);
        const uint4 a = ld_stream_v4(q), b = ld_stream_v4(q + 1);
        h[0] = a.x; h[1] = a.y; h[2] = a.z; h[3] = a.w; h[4] = b.x; h[5] = b.y; h[6] = b.z; h[7] = b.w;
    }
};
// a padding handle (steps past the end of the index): 32-bit handles are widened, so test the narrow all-ones too
template <typename HT> __device__ __forceinline__ uint64_t k1_node_of(uint64_t h) {
    return sizeof(HT) == 4 && h == 0xffffffffull ? ~0ull : (h >> 1);
}

// HT: handle type of the caller's step array (uint64_t = Handle as the reference stores it; uint32_t = the
// same value in 32 bits, dense idx < 2^31).  FIRST_OCC: maintain the first-occurrence keys.
// Block = 8 worker warps + 1 look-back warp.  Worker thread t owns the 8 CONSECUTIVE steps [8t, 8t+8) of the tile:
// 16-byte handle loads, eight gathers in flight, a register scan, one warp-shuffle scan, warp totals through shared
// memory, one 16-byte store per record.  The ninth warp runs the decoupled look-back from the start, under the
// workers' loads and gathers, so that the tile's exclusive prefix is normally there when the workers need it.
// handles: this chunk's steps, PADDED to a whole number of tiles with all-ones handles (node >= N: length 0);
// recs is padded likewise (records past the last step are written and never read).
// chunk_begin (a multiple of K1_TILE): index-local step of handles[0];  S: steps in the index.
// flags[0]: look-back watchdog; flags[1]: skip the look-back (timing experiment); flags[2]: a node length >= 2^31.
constexpr int K1_BLOCK = K1_THREADS + 32;
template <typename HT, bool FIRST_OCC>
__global__ void __launch_bounds__(K1_BLOCK, 4)
k1_scan_write(const HT* __restrict__ handles, uint32_t* __restrict__ node_tbl,
              uint32_t* __restrict__ first_key, uint32_t N, const uint64_t* __restrict__ first_step,
              uint32_t P, uint64_t chunk_begin, uint64_t S, uint64_t* __restrict__ desc, unsigned int* __restrict__ flags,
              uint32_t key_shift, StepRec* __restrict__ recs, uint64_t* __restrict__ path_len) {
    __shared__ uint64_t s_wsum[K1_THREADS / 32];
    __shared__ uint64_t s_prefix;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // blocks are dispatched in index order, so every predecessor of a running tile is running or done: the
    // look-back can never wait for a block that has not been scheduled (and a watchdog bounds it anyway)
    const uint64_t g_first = chunk_begin + (uint64_t)blockIdx.x * K1_TILE;
    const uint64_t gtile = g_first / K1_TILE;
    const uint64_t g_last = (g_first + K1_TILE <= S ? g_first + K1_TILE : S) - 1;
    // path of the tile's first and last step (most tiles lie inside one path): every warp finds them for itself
    const uint32_t p_first = warp_find_path(first_step, P, g_first, lane);
    const uint32_t p_last = warp_find_path(first_step, P, g_last, lane);
    const uint64_t fs_first = __ldg(first_step + p_first);
    const bool start_first = fs_first == g_first;                   // the tile begins exactly at a path start
    const bool has_start = start_first || p_last != p_first;

    if (w == K1_THREADS / 32) {
        // ---- look-back warp ------------------------------------------------------------------------------
        // Every round reads K1_LB_WINDOWS x 32 predecessor descriptors at once (all loads in flight together) and
        // consumes them nearest window first, up to the nearest inclusive value.  The wider the round, the fewer L2
        // round trips separate a tile from the inclusive front — with hundreds of tiles in flight that chain, not
        // bandwidth, is what bounds a chained scan (32 per round capped this kernel at ~27 tiles/us).
        uint64_t T = 0;
        if (!start_first && !flags[1]) {                        // flags[1]: timing experiment only (GFASORT_K1_NO_LOOKBACK: wrong offsets)
            int64_t look = (int64_t)gtile - 1;
            bool found = false;
            while (!found) {
                uint64_t d[K1_LB_WINDOWS];
#pragma unroll
                for (int j = 0; j < K1_LB_WINDOWS; ++j) {
                    const int64_t t = look - 32 * j - lane;
                    d[j] = t >= 0 ? ld_desc(desc + t) : K1_ST_INCL;     // before the first tile: offset 0
                }
#pragma unroll
                for (int j = 0; j < K1_LB_WINDOWS; ++j) {
                    if (found) break;
                    const int64_t t = look - 32 * j - lane;
                    uint32_t spins = 0;
                    while ((d[j] >> 62) == 0) {                         // not published yet: poll
                        d[j] = ld_desc(desc + t);
                        if (++spins > K1_SPIN_CAP) { atomicExch(flags, 1u); d[j] = K1_ST_INCL; break; }   // watchdog: never a hang
                    }
                    const unsigned incl = __ballot_sync(0xffffffffu, (d[j] >> 62) == 2);
                    const int stop = incl ? __ffs(incl) - 1 : 32;       // nearest predecessor with an inclusive value
                    uint64_t v = lane <= stop ? (d[j] & K1_VAL_MASK) : 0;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    T += v;
                    found = incl != 0;
                }
                look -= 32 * K1_LB_WINDOWS;
            }
        }
        if (lane == 0) s_prefix = T;
        asm volatile("bar.sync 2, %0;" ::"n"(K1_BLOCK) : "memory");     // the workers' warp totals are in s_wsum (they passed barrier 1)
        if (lane == 0 && !has_start) {
            uint64_t tile_sum = 0;
#pragma unroll
            for (int k = 0; k < K1_THREADS / 32; ++k) tile_sum += s_wsum[k];
            st_desc(desc + gtile, K1_ST_INCL | (T + tile_sum));
        }
        return;
    }
    // ---- worker warps ------------------------------------------------------------------------------------
    const uint64_t my_first = g_first + (uint64_t)threadIdx.x * K1_ITEMS;      // this thread's first step
    uint64_t hh[K1_ITEMS];
    K1Load<HT>::load(handles + (my_first - chunk_begin), hh);
    uint32_t len[K1_ITEMS];
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {                            // all gathers in flight before the first use
        const uint64_t node = k1_node_of<HT>(hh[k]);
        len[k] = K1_VISITED;                                        // missing node => +0 (src/sgd.rs:52-54), nothing to mark
        // L2 load (the visited bits change during the launch; a stale clear bit only costs a redundant atomicMin)
        if (node < N) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(len[k]) : "l"(node_tbl + node));
    }
    uint32_t nr[K1_ITEMS];
    uint64_t loc[K1_ITEMS];
    uint64_t tsum = 0;
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t node = k1_node_of<HT>(hh[k]);
        nr[k] = (uint32_t)(((node < N ? node : N) << 1) | (hh[k] & 1));
        if (FIRST_OCC && !(len[k] & K1_VISITED)) {
            atomicOr(node_tbl + node, K1_VISITED);
            atomicMin(first_key + node, (uint32_t)((my_first + k) >> key_shift));
        }
        len[k] &= ~K1_VISITED;
        loc[k] = tsum;
        tsum += len[k];
    }
    uint64_t inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint64_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_wsum[w] = inc;
    asm volatile("bar.sync 1, %0;" ::"n"(K1_THREADS) : "memory");        // workers only
    uint64_t woff = 0, tile_sum = 0;
#pragma unroll
    for (int k = 0; k < K1_THREADS / 32; ++k) { const uint64_t v = s_wsum[k]; woff += (k < w) ? v : 0; tile_sum += v; }
    const uint64_t texcl = woff + (inc - tsum);                     // tile-local exclusive prefix of this thread's first step
    // publish at once what the successors can use: the tile's sum, or — when the scan restarts inside the tile — the
    // offset after its last path start (nothing before the tile matters to them then)
    uint32_t j_ls = 0;
    if (has_start) j_ls = (uint32_t)(__ldg(first_step + p_last) - g_first);      // p_last's first step is in the tile
    if (has_start) {
        // the thread that owns step j_ls knows its local prefix
        if (j_ls / K1_ITEMS == threadIdx.x) {
            uint64_t lp = 0;                                        // loc[j_ls % 8] without indexing a register array
#pragma unroll
            for (int k = 0; k < K1_ITEMS; ++k) lp += k < (int)(j_ls % K1_ITEMS) ? len[k] : 0u;
            st_desc(desc + gtile, K1_ST_INCL | (tile_sum - (texcl + lp)));
        }
    } else if (threadIdx.x == 0) {
        st_desc(desc + gtile, K1_ST_AGG | tile_sum);
    }
    asm volatile("bar.sync 2, %0;" ::"n"(K1_BLOCK) : "memory");          // the look-back warp has delivered s_prefix
    const uint64_t base = s_prefix + texcl;                          // s_prefix == 0 when the tile begins at a path start
    const uint64_t fs_next = __ldg(first_step + p_first + 1);       // single-path tiles: where the path ends
    StepRec* out = recs + my_first;
    // Offsets as if the whole tile continued p_first's path.  That is exact for every step of a single-path tile (all
    // but at most P tiles) and for p_first's steps in a tile where other paths start; the steps after such a start
    // are off by the tile-local prefix at that start, which k1_fix_path_starts subtracts afterwards.
#pragma unroll
    for (int k = 0; k < K1_ITEMS; ++k) {
        const uint64_t pos = base + loc[k];
        // StepRec {node_rev, node_len, pos} as one 16-byte store (see load_rec)
        st_stream_v4(out + k, make_uint4(nr[k], len[k], (uint32_t)pos, (uint32_t)(pos >> 32)));
        if (my_first + k + 1 == fs_next) path_len[p_first] = pos + len[k];         // the path's last step: PathInfo.length (sgd.rs:64-68)
    }
}

// The paths that start INSIDE a tile (not at its first step): k1_scan_write gave the steps from such a start to the
// end of the tile (or of the path) offsets that continue the previous path; subtract the offset it gave the start.
// One block per path; at most P - 1 blocks do anything, each touches < K1_TILE records.  Launched after the
// k1_scan_write of the chunk [c_begin, c_end) that holds the start.
__global__ void __launch_bounds__(K1_THREADS)
k1_fix_path_starts(const uint64_t* __restrict__ first_step, uint32_t P, uint64_t c_begin, uint64_t c_end, StepRec* __restrict__ recs,
                   uint64_t* __restrict__ path_len) {
    __shared__ uint64_t s0;
    const uint32_t p = blockIdx.x + 1;
    if (p >= P) return;
    const uint64_t fs = first_step[p], fe = first_step[p + 1];
    if (fs == fe || fs < c_begin || fs >= c_end || fs % K1_TILE == 0) return;       // empty / not this chunk / tile-aligned start (already exact)
    const uint64_t tile_end = (fs / K1_TILE + 1) * K1_TILE;
    const uint64_t end = fe < tile_end ? fe : tile_end;
    if (threadIdx.x == 0) s0 = recs[fs].pos;
    __syncthreads();
    const uint64_t off = s0;
    for (uint64_t i = fs + threadIdx.x; i < end; i += blockDim.x) {
        const uint64_t pos = recs[i].pos - off;
        recs[i].pos = pos;
        if (i + 1 == fe) path_len[p] = pos + recs[i].node_len;
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
__global__ void rl_keys(const uint32_t* __restrict__ first_key, uint32_t N, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) { keys[i] = first_key[i]; vals[i] = i; }
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
