// gfs_device.cuh — device-side building blocks shared by the SGD, stress and debug kernels.
//
// Everything numerical here is written with explicit round-to-nearest intrinsics (__dmul_rn,
// __dadd_rn, ...) so that nvcc cannot contract a*b+c into an FMA: the reference is Rust, which
// never fuses, and the parity tests compare these functions bit-for-bit with the CPU oracle.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gfs {

// One step of a path as the kernels read it: a single 16-byte, 16-byte-aligned record, so a random
// step costs exactly one 32-byte sector.  Replaces the reference's four parallel 8-byte arrays
// (src/sgd.rs:14-24); path id and rank are recovered from the P+1 first_step table instead.
struct __align__(16) StepRec {
    uint32_t node_rev;   // (dense node idx << 1) | is_reverse        (Handle, src/graph.rs:9-19)
    uint32_t node_len;   // sequence length of that node             (needed by the nD kernel, sgd.rs:1051-1058)
    uint64_t pos;        // nucleotide offset of the step in its path (step_to_position, sgd.rs:18)
};
static_assert(sizeof(StepRec) == 16, "StepRec must be one 16-byte record");

enum : uint32_t { STREAM_SGD = 1, STREAM_STRESS = 2 };

// ---- Philox4x32-10 (Salmon et al. SC'11) -------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// ---- fast_precise_pow (src/sgd.rs:155-182) -----------------------------------------------------
// a^b = a^e * "a^(b-e)", e = trunc(b): the integer part by square-and-multiply, the fractional part by
// the high-word bit trick.  __double2int_rz saturates and maps NaN to 0 exactly like Rust's `as i32`.
// frac_pow is the bit trick alone (it is the whole function when e == 0, since 1.0 * x == x).
__device__ __forceinline__ double frac_pow(double a, double bfrac) {
    const int diff = (int)((unsigned)__double2hiint(a) - 1072632447u);       // wrapping, as release Rust
    const int nh = __double2int_rz(__dadd_rn(__dmul_rn(bfrac, (double)diff), 1072632447.0));
    return __hiloint2double(nh, 0);
}
// a^e, e >= 0, with exactly the multiplication order of the reference's loop (src/sgd.rs:171-179):
// r = 1; for each bit of e from the LSB: if set r *= base; base *= base.
__device__ __forceinline__ double int_pow(double a, int e) {
    double base = a, r = 1.0;
    while (e != 0) {
        if (e & 1) r = __dmul_rn(r, base);
        base = __dmul_rn(base, base);
        e >>= 1;
    }
    return r;
}
// Same products in the same order for a compile-time exponent (the loop unrolls and the untaken
// multiplications disappear).  The sampler's alpha = 1/(1-theta) truncates to 99 for the
// reference's theta = 0.99 (1/(1-0.99) = 99.99999999999991) and to 1 for the cooling theta = 0.001.
template <int E>
__device__ __forceinline__ double int_pow_fixed(double a) {
    double base = a, r = 1.0;
    bool first = true;
#pragma unroll
    for (int e = E; e != 0; e >>= 1) {
        if (e & 1) { r = first ? base : __dmul_rn(r, base); first = false; }   // 1.0 * base == base exactly
        if (e >> 1) base = __dmul_rn(base, base);
    }
    return r;
}
__device__ __forceinline__ double fast_precise_pow(double a, double b) {
    const int e = __double2int_rz(b);
    return __dmul_rn(int_pow(a, e), frac_pow(a, __dsub_rn(b, (double)e)));
}

// Per-epoch constants of the Zipf sampler: pure functions of the current theta, computed once on
// the host with the same formulas (src/sgd.rs:132, 143, 471).
struct ZipfConsts {
    double theta;
    double one_minus_theta;   // 1.0 - theta            (exponent of fpp(2/n, .): integer part 0)
    double alpha;             // 1.0 / (1.0 - theta)
    double z2;                // 1.0 + fast_precise_pow(0.5, theta)
    double alpha_frac;        // alpha - trunc(alpha)
    int alpha_e;              // trunc(alpha) as the reference's `as i32`
    int pad;
};

// ---- DirtyZipfian::sample with min = 1, max = jump_space (src/sgd.rs:122-151) -------------------
// Written without branches (the two fast paths become selects) so that several terms can be in
// flight per thread, and split in two so that everything that does not need zeta — most of the
// work — can be scheduled while the zeta load is still in flight.
// __double2uint_rz saturates / maps negatives and NaN to 0 like Rust's `as u64` followed by
// .min(max) with max < 2^32.
struct ZipfPre { double n, num; };
__device__ __forceinline__ ZipfPre dirty_zipf_pre(uint32_t jump_space, const ZipfConsts& zc) {
    ZipfPre p;
    p.n = (double)jump_space;
    // 1 - theta has integer part 0 for every theta in (0, 1): fpp(2/n, 1-theta) is the bit trick alone
    const double t0 = __ddiv_rn(2.0, p.n);
    const double f1 = zc.one_minus_theta < 1.0 ? frac_pow(t0, zc.one_minus_theta) : fast_precise_pow(t0, zc.one_minus_theta);
    p.num = __dsub_rn(1.0, f1);
    return p;
}
// den = 1 - zeta2theta / zeta (sgd.rs:133-134): a function of the zeta table entry and the epoch's theta alone, so the
// kernels read it from the table ({zeta, den} pairs, one table per theta, filled on the host with the same two IEEE
// operations) instead of running a second dependent division behind the zeta load.
__device__ __forceinline__ uint32_t dirty_zipf_post(uint32_t jump_space, const ZipfConsts& zc, const ZipfPre& p,
                                                    double zeta, double den, double u) {
    const double uz = __dmul_rn(u, zeta);
    const double eta = __ddiv_rn(p.num, den);
    const double base = __dadd_rn(__dsub_rn(__dmul_rn(eta, u), eta), 1.0);
    double ip;
    if (zc.alpha_e == 99) ip = int_pow_fixed<99>(base);  // warp-uniform branches (per-epoch constant)
    else if (zc.alpha_e == 1) ip = base;                 // 1.0 * base
    else ip = int_pow(base, zc.alpha_e);
    const double pw = __dmul_rn(ip, frac_pow(base, zc.alpha_frac));
    const double result = __dadd_rn(1.0, __dmul_rn(p.n, pw));
    uint32_t z = __double2uint_rz(result);
    z = z > jump_space ? jump_space : z;
    z = uz < zc.z2 ? 2u : z;                             // sgd.rs:142-144 (min + 1, not clamped to max)
    z = uz < 1.0 ? 1u : z;                               // sgd.rs:139-141
    return z;
}
__device__ __forceinline__ uint32_t dirty_zipf(uint32_t jump_space, const ZipfConsts& zc, double zeta, double u) {
    return dirty_zipf_post(jump_space, zc, dirty_zipf_pre(jump_space, zc), zeta, __dsub_rn(1.0, __ddiv_rn(zc.z2, zeta)), u);
}

// exact u64 -> f64 for v < 2^52 (step offsets; gfs_index_build rejects longer paths): one DADD
// instead of a 64-bit I2F.
__device__ __forceinline__ double u52_to_f64(uint64_t v) {
    return __dsub_rn(__hiloint2double(0x43300000 | (int)(v >> 32), (int)(uint32_t)v), 4503599627370496.0);
}

// ---- memory access helpers ---------------------------------------------------------------------
// Step records are read-only and streamed once each: non-coherent path, no L1 allocation.
__device__ __forceinline__ StepRec load_rec(const StepRec* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    StepRec r;
    r.node_rev = v.x; r.node_len = v.y; r.pos = ((uint64_t)v.w << 32) | v.z;
    return r;
}
// largest p in [0, P) with first_step[p] <= s   (first_step has P+1 entries, first_step[P] = S > s)
__device__ __forceinline__ uint32_t find_path(const uint64_t* __restrict__ fs, uint32_t P, uint64_t s) {
    uint32_t lo = 0, hi = P;   // invariant: fs[lo] <= s < fs[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (fs[mid] <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// Sum over the lanes named in `peers` (the valid lanes sharing this lane's key); every lane of
// `active` calls, lanes outside every group pass a peers mask that does not contain themselves.
// Returns the group sum in the group's leader (lowest lane); other lanes get an undefined value.
template <typename T>
__device__ __forceinline__ T group_sum(unsigned active, unsigned peers, T v, int lane, bool& is_leader) {
    is_leader = (__ffs(peers) - 1) == lane;   // false for lanes not in their own mask
    // fast path: no lane of the warp shares a key
    const bool dup = __popc(peers) > 1;
    if (!__any_sync(active, dup)) return v;
    T acc = v;
    unsigned rest = peers & ~(1u << lane);
    // rounds: in round r each lane fetches the r-th other member of its group
    const int rounds = __reduce_max_sync(active, (unsigned)__popc(rest));
    for (int r = 0; r < rounds; ++r) {
        const int src = rest ? (__ffs(rest) - 1) : lane;
        const T o = __shfl_sync(active, v, src);
        if (rest) { acc += o; rest &= rest - 1; }
    }
    return acc;
}

// One epoch of the schedule as the kernels see it.
struct EpochDesc {
    double eta;
    ZipfConsts zc;
    uint64_t updates;     // min_term_updates
    uint32_t cooling;
    uint32_t ztab;        // which {zeta, den} table this epoch's theta uses (0 = params.theta, 1 = the cooling theta)
};

}  // namespace gfs
