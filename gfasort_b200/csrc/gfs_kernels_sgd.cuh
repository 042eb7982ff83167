// gfs_kernels_sgd.cuh — term sampling and the persistent, software-pipelined SGD term kernel K2/K3 (reference src/sgd.rs:442-584, 988-1156).
// Part of libgfasort_cuda.so; included by gfs_lib.cu (one translation unit).  See DESIGN.md §4.
#pragma once
#include "gfs_device.cuh"

namespace gfs {

// =============================================================================================
// term sampling (shared by K2, K3, trace) — SURVEY.md Appendix A steps 1-5'
// =============================================================================================
struct KernelGraph {
    const StepRec* recs;
    const uint64_t* first_step;   // P+1 (global memory copy)
    const double2* zetas;         // 2 tables (one per theta, EpochDesc::ztab) of zlen {zeta, 1 - zeta2theta/zeta} pairs (global)
    uint64_t S;
    uint32_t P;
    uint32_t N;
    uint32_t zlen;
    uint32_t space;               // min(params.space, 2^32-1): compared with ranks < 2^32
    uint32_t space_max;
    uint32_t q;
    uint32_t q_is_100;            // 1: the quantisation step is the reference's 100 (constant division)
    uint32_t blk_shift;           // path-of-block table granularity: block = step >> blk_shift
    uint32_t coherent;            // 0: every lane draws its own step; G = 2..32 (power of two): groups of G lanes
                                  // sample G consecutive steps (see sample_s1)
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
    // experiment (GFASORT_ZETA_SMEM, profiles/r2_experiments.md): the first zs_n entries of both zeta tables staged in
    // shared memory ([table][entry]); null = every zeta read goes to L2 through the read-only path (the default)
    const double2* zs;
    uint32_t zs_n;
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
    double zeta, den;          // table entry: zeta and 1 - zeta2theta/zeta
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
        // warp-coherent sampling (sweep schedule only): the first lane of every group of G lanes draws the step,
        // lane l of the group takes the l-th step after it.  Every step is still drawn with the same probability
        // over a sweep, but the G sampled records — and, with the node relabelling, most of their nodes'
        // positions — are adjacent in memory: one coalesced request instead of G.  Partners stay independent
        // per lane.  (A partial warp — only with a hand-picked thread count — falls back to one group led by
        // its first active lane.)
        const uint32_t G = g.coherent;
        const bool full = warp_mask == 0xffffffffu;
        const int src = full ? (lane & ~(int)(G - 1)) : __ffs(warp_mask) - 1;
        s = __shfl_sync(warp_mask, s, src) + (uint32_t)(full ? (lane & (int)(G - 1)) : lane);
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
    t.zeta = 1.0; t.den = 1.0;
    if (t.live && t.zipf) {
        const double2 zd = (pl.zs && k < pl.zs_n) ? pl.zs[ep.ztab * pl.zs_n + k] : __ldg(g.zetas + (size_t)ep.ztab * g.zlen + k);
        t.zeta = zd.x; t.den = zd.y;
    }
}

__device__ __forceinline__ void sample_s2(const KernelGraph& g, const EpochDesc& ep, Slot& t) {
    const uint32_t n = t.n, ra = t.ra;
    // u = (r23 >> 11) * 2^-53 (sgd.rs:136 through PhiloxDraw::unit)
    const double u = __dmul_rn((double)(t.r23 >> 11), 1.0 / 9007199254740992.0);
    const ZipfPre pre = dirty_zipf_pre(t.J, ep.zc);
    const uint32_t z = dirty_zipf_post(t.J, ep.zc, pre, t.zeta, t.den, u);
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
    uint32_t zeta_smem;          // experiment: zeta-table entries (per theta) staged in shared memory; 0 = none
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
constexpr int SGD_MIN_BLOCKS_K1 = 4, SGD_MIN_BLOCKS_K2 = 3;     // resident blocks per SM the register budget is set for
template <typename CT, int D, int DS, bool AGG, int K>
__global__ void __launch_bounds__(SGD_BLOCK, (K > 1 ? SGD_MIN_BLOCKS_K2 : SGD_MIN_BLOCKS_K1))
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
    pl.zs = nullptr; pl.zs_n = 0;
    if (a.zeta_smem) {
        // after the path tables, 16-byte aligned
        const size_t off = ((size_t)n_fs * 8 + (tables ? (size_t)BLK_TABLE * 2 : 0) + 15) & ~(size_t)15;
        double2* s_z = reinterpret_cast<double2*>(smem_raw + off);
        const uint32_t n = a.zeta_smem < a.g.zlen ? a.zeta_smem : a.g.zlen;
        for (uint32_t k = threadIdx.x; k < 2 * n; k += blockDim.x) s_z[k] = a.g.zetas[(size_t)(k / n) * a.g.zlen + (k % n)];
        __syncthreads();
        pl.zs = s_z; pl.zs_n = n;
    }

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
    // A sliced epoch (n_slices > 1: one launch per reconcile interval) is ONE sweep cut into n_slices consecutive pieces:
    // slice k runs the epoch's chunks [cpe_all k / n, cpe_all (k+1) / n), so the window positions of the slices follow one
    // another exactly as in a whole-epoch launch (restarting the sweep in every slice costs 6-7 % in L2 locality).
    const uint64_t m_sweep = ep.updates;                                             // the whole epoch's updates
    const uint64_t cpe_all = sweep ? (m_sweep + C - 1) / C : 1;                       // chunks per (whole) epoch
    const uint64_t c_first = cpe_all * a.slice / a.n_slices;                          // this slice's chunks of every epoch
    const uint64_t cpe = cpe_all * (a.slice + 1) / a.n_slices - c_first;
    const uint64_t total_chunks = cpe * (uint64_t)(a.epoch_end - a.epoch_begin);
    const double steps_per_chunk = (double)a.g.samp_len / (double)cpe_all;
    uint64_t c_lo = 0;                        // first chunk of epoch e_cur
    bool first_claim = true;
    auto claim = [&]() -> bool {
        if (sweep) {
            unsigned long long c = 0;
            if (lane == leader) c = atomicAdd(a.work_ctr, 1ull);
            c = __shfl_sync(warp_mask, c, leader);
            if (c >= total_chunks) return false;
            while (c >= c_lo + cpe) { c_lo += cpe; ++e_cur; ep = a.epochs[e_cur]; }       // claims only move forward
            const uint64_t cc = c_first + (c - c_lo);                                     // chunk index within the whole epoch
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

}  // namespace gfs
