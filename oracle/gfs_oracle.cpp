// gfs_oracle.cpp — CPU ORACLE for the path-guided SGD hot path of pangenome/gfasort.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (gfasort_b200/libgfasort_cuda.so) never links, imports or calls anything in oracle/.
//
// What it is: a C++ restatement of /root/reference/src/sgd.rs (gfasort v0.1.0), function by
// function, with the reference file:line each block follows.  The reference is safe Rust and
// there is no Rust toolchain in this image or on the GPU box, so the reference itself cannot
// be built (oracle/_ref does not exist); this restatement is both the parity checker and the
// "port" CPU baseline.
//
// PARITY STATUS: the reference's own tests hold NO golden vector, known-answer test or fixture
// value for this path (tests/integration_tests.rs only checks node/edge counts).  The
// arithmetic that lives in un-vendored crates (rand 0.9 `Uniform`, rand_xoshiro 0.7
// `Xoshiro256Plus::seed_from_u64`, rand_distr 0.5 `StandardNormal`; Cargo.lock is git-ignored)
// is restated here from the crates' published algorithms (xoshiro256+ / SplitMix64 seeding /
// widening-multiply rejection for bounded ints / 53-bit mantissa uniform).  The generator cores
// (SplitMix64, xoshiro256+) are pinned by their published known-answer vectors
// (tests/test_oracle.py::test_splitmix64_published_vectors, ::test_xoshiro256plus_published_vector);
// rand's `Uniform<usize>` rejection rule and rand_distr's ziggurat `StandardNormal` (replaced here by
// Marsaglia polar) are restated from the crates' documentation and stay unpinned.  => "parity
// unpinned" at the RNG boundary (whole-run parity is statistical); everything deterministic (path
// index, fast_precise_pow, DirtyZipfian given u, eta schedule, zeta table, X init, layout order) is
// pinned by the hand-derived known answers in tests/golden/known_answers.json.
//
// Two draw policies feed the SAME restated term loop:
//   * XoshiroDraw — the reference's RNG and lazy draw order (sgd.rs:431-494, 1062-1071)
//   * PhiloxDraw  — the GPU library's counter-based stream and fixed draw slots, so that the
//                   CUDA kernels' term sampling can be checked BIT-EXACTLY against this file
//                   (same (seed, tid, attempt) -> same (step_a, step_b, ends, d, mu)).
//
// Build: see oracle/Makefile (g++ -O3 -march=x86-64-v3 -ffp-contract=off -pthread).
// -ffp-contract=off matters: Rust never fuses a*b+c, so neither may this file.

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// Rust `as` cast semantics (saturating, NaN -> 0), used by fast_precise_pow / DirtyZipfian.
// ---------------------------------------------------------------------------------------------
inline int32_t f64_as_i32(double v) {
    if (std::isnan(v)) return 0;
    if (v >= 2147483647.0) return std::numeric_limits<int32_t>::max();
    if (v <= -2147483648.0) return std::numeric_limits<int32_t>::min();
    return (int32_t)v;
}
inline uint64_t f64_as_u64(double v) {
    if (std::isnan(v)) return 0;
    if (v <= 0.0) return 0;
    if (v >= 18446744073709551616.0) return std::numeric_limits<uint64_t>::max();
    return (uint64_t)v;
}
inline double bits_f64(uint64_t b) { double d; std::memcpy(&d, &b, 8); return d; }
inline uint64_t f64_bits(double d) { uint64_t b; std::memcpy(&b, &d, 8); return b; }

// sgd.rs:155-182  fast_precise_pow
inline double fast_precise_pow(double a, double b) {
    int32_t e = f64_as_i32(b);
    uint64_t bits = f64_bits(a);
    int32_t high = (int32_t)(uint32_t)(bits >> 32);
    // Rust release builds wrap on i32 overflow; do the subtraction in uint32.
    int32_t diff = (int32_t)((uint32_t)high - 1072632447u);
    int32_t new_high = f64_as_i32((b - (double)e) * (double)diff + 1072632447.0);
    uint64_t frac_bits = ((uint64_t)(int64_t)new_high) << 32;   // `new_high as u64` sign-extends
    double frac = bits_f64(frac_bits);
    double base = a;
    int32_t ex = e;
    double r = 1.0;
    while (ex != 0) {
        if (ex & 1) r *= base;
        base *= base;
        ex >>= 1;
    }
    return r * frac;
}

// sgd.rs:122-151  DirtyZipfian::sample, with the uniform draw `u` supplied by the caller.
inline uint64_t dirty_zipf(uint64_t zmin, uint64_t zmax, double theta, double zeta, double zeta2theta,
                           double u) {
    uint64_t n = zmax - zmin + 1;
    double alpha = 1.0 / (1.0 - theta);
    double eta = (1.0 - fast_precise_pow(2.0 / (double)n, 1.0 - theta)) / (1.0 - zeta2theta / zeta);
    double uz = u * zeta;
    if (uz < 1.0) return zmin;
    if (uz < 1.0 + fast_precise_pow(0.5, theta)) return zmin + 1;
    double result = (double)zmin + ((double)n * fast_precise_pow(eta * u - eta + 1.0, alpha));
    return std::min(f64_as_u64(result), zmax);
}

// sgd.rs:617-638  path_linear_sgd_schedule
void schedule(double w_min, double w_max, uint64_t iter_max, uint64_t iter_with_max_lr, double eps,
              double* etas) {
    double eta_max = 1.0 / w_min;
    double eta_min = eps / w_max;
    double lambda = std::log(eta_max / eta_min) / ((double)iter_max - 1.0);
    for (uint64_t t = 0; t <= iter_max; ++t) {
        int64_t dt = (int64_t)t - (int64_t)iter_with_max_lr;
        if (dt < 0) dt = -dt;
        etas[t] = eta_max * std::exp(-lambda * (double)dt);
    }
}

// sgd.rs:311-315  size of the zeta table
uint64_t zeta_size(uint64_t space, uint64_t space_max, uint64_t q) {
    uint64_t s = (space <= space_max) ? space : space_max + (space - space_max) / q + 1;
    return s + 1;
}
// sgd.rs:317-331  zeta table.  `iter_cap` (0 = none) stops the serial loop early; entries past
// the cap stay 0 exactly as unreached entries would if `space` were smaller.
void zetas_fill(uint64_t space, uint64_t space_max, uint64_t q, double theta, double* zetas,
                uint64_t size, uint64_t iter_cap) {
    for (uint64_t i = 0; i < size; ++i) zetas[i] = 0.0;
    double z = 0.0;
    uint64_t last = (iter_cap && iter_cap < space) ? iter_cap : space;
    for (uint64_t i = 1; i <= last; ++i) {
        z += fast_precise_pow(1.0 / (double)i, theta);
        if (i <= space_max) zetas[i] = z;
        if (i >= space_max && (i - space_max) % q == 0) {
            uint64_t idx = space_max + 1 + (i - space_max) / q;
            if (idx < size) zetas[idx] = z;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RNGs
// ---------------------------------------------------------------------------------------------
// rand_xoshiro 0.7: Xoshiro256Plus, seed_from_u64 = 4 outputs of SplitMix64(seed).
// SplitMix64 (Steele, Lea, Flood; Vigna's splitmix64.c): the seeding generator of rand_core's
// `seed_from_u64` for the xoshiro family.  Pinned by published vectors in tests/test_oracle.py.
static inline uint64_t splitmix64_next(uint64_t& x) {
    x += 0x9e3779b97f4a7c15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
struct Xoshiro256Plus {
    uint64_t s[4];
    static Xoshiro256Plus seed_from_u64(uint64_t seed) {
        Xoshiro256Plus r;
        uint64_t x = seed;
        for (int i = 0; i < 4; ++i) r.s[i] = splitmix64_next(x);
        return r;
    }
    static Xoshiro256Plus from_state(const uint64_t st[4]) {
        Xoshiro256Plus r;
        for (int i = 0; i < 4; ++i) r.s[i] = st[i];
        return r;
    }
    inline uint64_t next_u64() {
        uint64_t result = s[0] + s[3];
        uint64_t t = s[1] << 17;
        s[2] ^= s[0];
        s[3] ^= s[1];
        s[1] ^= s[2];
        s[0] ^= s[3];
        s[2] ^= t;
        s[3] = (s[3] << 45) | (s[3] >> 19);
        return result;
    }
    inline uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }   // upper bits, as the crate
    inline double next_f64() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    // rand 0.9 Uniform<usize>::new(0, n).sample: 32-bit lane when n-1 fits u32, else 64-bit;
    // widening multiply, reject when lo < (2^w - n) % n.
    inline uint64_t below(uint64_t n) {
        if (n - 1 <= 0xffffffffULL) {
            uint32_t range = (uint32_t)n;   // n == 2^32 -> range 0 -> any u32
            if (range == 0) return next_u32();
            uint32_t thresh = (uint32_t)(0u - range) % range;
            for (;;) {
                uint64_t m = (uint64_t)next_u32() * (uint64_t)range;
                if ((uint32_t)m >= thresh) return m >> 32;
            }
        } else {
            uint64_t thresh = (0ULL - n) % n;
            for (;;) {
                unsigned __int128 m = (unsigned __int128)next_u64() * (unsigned __int128)n;
                if ((uint64_t)m >= thresh) return (uint64_t)(m >> 64);
            }
        }
    }
    // N(0,1).  rand_distr's ziggurat tables are not restated; Marsaglia polar is used instead
    // (documented: the Gaussian init stream is statistical, not bit-level, parity).
    double spare = 0.0; bool has_spare = false;
    inline double normal() {
        if (has_spare) { has_spare = false; return spare; }
        double u, v, q;
        do {
            u = 2.0 * next_f64() - 1.0;
            v = 2.0 * next_f64() - 1.0;
            q = u * u + v * v;
        } while (q >= 1.0 || q == 0.0);
        double f = std::sqrt(-2.0 * std::log(q) / q);
        spare = v * f; has_spare = true;
        return u * f;
    }
};

// Philox4x32-10 (Salmon et al., SC'11; Random123 constants).
struct Philox {
    static inline void round(uint32_t c[4], uint32_t k0, uint32_t k1) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    }
    static inline void gen(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
        uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
        uint32_t k0 = key[0], k1 = key[1];
        for (int r = 0; r < 10; ++r) {
            round(c, k0, k1);
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
    }
};
inline uint64_t mulhi64(uint64_t a, uint64_t b) {
    return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
}

// ---------------------------------------------------------------------------------------------
// Graph view handed in from the harness (flat arrays; ids as in the reference's BidirectedGraph)
// ---------------------------------------------------------------------------------------------
struct GraphView {
    // nodes: Vec<Option<BiNode>> indexed by node id; present[id] != 0 <=> Some
    const uint8_t* present; const uint64_t* seq_len; uint64_t nodes_len;
    const uint64_t* node_order; uint64_t node_order_len;
    // paths: concatenated Handle(u64) = id<<1 | is_rev ; path p = [path_first[p], path_first[p+1])
    const uint64_t* steps; const uint64_t* path_first; uint64_t num_paths;
    inline bool has(uint64_t id) const { return id < nodes_len && present[id]; }
    inline uint64_t len(uint64_t id) const { return has(id) ? seq_len[id] : 0; }
    uint64_t node_count() const { uint64_t c = 0; for (uint64_t i = 0; i < nodes_len; ++i) c += present[i] != 0; return c; }
};

// sgd.rs:14-108  PathIndex (+ PathInfo): four parallel per-step arrays, faithful layout.
struct PathInfo { uint64_t step_count, length, first_step; };
struct PathIndex {
    std::vector<uint64_t> step_to_handle, step_to_position, step_to_path, step_to_rank;
    std::vector<PathInfo> paths;
    static PathIndex from_graph(const GraphView& g) {          // sgd.rs:34-71
        PathIndex ix;
        uint64_t S = g.path_first[g.num_paths];
        ix.step_to_handle.reserve(S); ix.step_to_position.reserve(S);
        ix.step_to_path.reserve(S); ix.step_to_rank.reserve(S);
        for (uint64_t p = 0; p < g.num_paths; ++p) {
            uint64_t first_step = ix.step_to_handle.size();
            uint64_t position = 0;
            uint64_t cnt = g.path_first[p + 1] - g.path_first[p];
            for (uint64_t r = 0; r < cnt; ++r) {
                uint64_t h = g.steps[g.path_first[p] + r];
                ix.step_to_handle.push_back(h);
                ix.step_to_position.push_back(position);
                ix.step_to_path.push_back(p);
                ix.step_to_rank.push_back(r);
                position += g.len(h >> 1);                     // missing node => +0 (:52-54)
            }
            ix.paths.push_back(PathInfo{cnt, position, first_step});
        }
        return ix;
    }
    uint64_t get_total_steps() const { return step_to_handle.size(); }
    uint64_t get_handle_of_step(uint64_t s) const { return step_to_handle[s]; }
    uint64_t get_position_of_step(uint64_t s) const { return step_to_position[s]; }
    uint64_t get_path_of_step(uint64_t s) const { return step_to_path[s]; }
    uint64_t get_rank_of_step(uint64_t s) const { return step_to_rank[s]; }
    uint64_t get_path_step_count(uint64_t p) const { return paths[p].step_count; }
    uint64_t get_step_at_path_position(uint64_t p, uint64_t r) const { return paths[p].first_step + r; }
    uint64_t num_paths() const { return paths.size(); }
    uint64_t get_path_length(uint64_t p) const { return paths[p].length; }
};

// handle_to_idx: HashMap<Handle, usize> (sgd.rs:272, 289).  Open addressing + a 64-bit mixer:
// one probe-sequence cache miss per lookup like hashbrown, cheaper than SipHash-1-3 — i.e. this
// baseline is, if anything, FASTER than the reference's map.
struct HandleMap {
    std::vector<uint64_t> keys; std::vector<uint64_t> vals; uint64_t mask = 0;
    static inline uint64_t mix(uint64_t x) {
        x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
        return x;
    }
    void init(uint64_t n) {
        uint64_t cap = 16; while (cap < n * 2) cap <<= 1;
        keys.assign(cap, ~0ULL); vals.assign(cap, 0); mask = cap - 1;
    }
    void insert(uint64_t k, uint64_t v) {
        uint64_t i = mix(k) & mask;
        while (keys[i] != ~0ULL && keys[i] != k) i = (i + 1) & mask;
        keys[i] = k; vals[i] = v;
    }
    inline bool get(uint64_t k, uint64_t& v) const {
        uint64_t i = mix(k) & mask;
        while (keys[i] != ~0ULL) { if (keys[i] == k) { v = vals[i]; return true; } i = (i + 1) & mask; }
        return false;
    }
};

// sgd.rs:276-284  node_ids = node_order, else sorted live ids
std::vector<uint64_t> node_ids_of(const GraphView& g) {
    std::vector<uint64_t> ids;
    if (g.node_order_len) { ids.assign(g.node_order, g.node_order + g.node_order_len); }
    else { for (uint64_t i = 0; i < g.nodes_len; ++i) if (g.present[i]) ids.push_back(i); }
    return ids;
}

struct Params {   // PathSGDParams / LayoutSGDParams (sgd.rs:196-212, 676-707), field for field
    uint64_t iter_max, iter_with_max_learning_rate, min_term_updates;
    double delta, eps, eta_max, theta;
    uint64_t space, space_max, space_quantization_step;
    double cooling_start;
    uint64_t nthreads; uint64_t progress; uint64_t seed;
};

struct Stats { uint64_t applied, attempts; double seconds; uint64_t epochs; };

// Shared control block (sgd.rs:340-346)
struct Control {
    std::atomic<uint64_t> term_updates{0}, iteration{0}, eta_bits{0}, theta_bits{0};
    std::atomic<bool> cooling{false}, work_todo{true};
    std::atomic<uint64_t> delta_max{0};
};

// ---------------------------------------------------------------------------------------------
// Draw policies
// ---------------------------------------------------------------------------------------------
struct XoshiroDraw {          // reference order, lazy draws
    Xoshiro256Plus rng;
    explicit XoshiroDraw(uint64_t seed) : rng(Xoshiro256Plus::seed_from_u64(seed)) {}
    inline void begin() {}
    inline uint64_t step(uint64_t S) { return rng.below(S); }          // sgd.rs:435,444
    inline bool coin_zipf() { return rng.below(2) == 1; }              // :456
    inline bool coin_back() { return rng.below(2) == 1; }              // :460
    inline double unit() { return rng.next_f64(); }                    // :136
    inline uint64_t rank(uint64_t n) { return rng.below(n); }          // :493-494
    inline bool coin_end_a() { return rng.below(2) == 1; }             // :1062
    inline bool coin_end_b() { return rng.below(2) == 1; }             // :1071
};
// GPU stream: one Philox4x32-10 block per attempt.
//   ctr = {attempt_lo, attempt_hi, tid, stream}, key = {seed_lo, seed_hi}
//   R01 = r1:r0 -> step = mulhi64(R01, S);  R23 = r3:r2 -> u = (R23>>11)*2^-53 | rank = mulhi64(R23, n)
//   coins = low bits of r2 (bit0 zipf, bit1 back, bit2 end_a, bit3 end_b); they sit below the 11
//   bits `u` discards.
struct PhiloxDraw {
    uint32_t key[2]; uint32_t tid; uint32_t stream; uint64_t attempt; uint32_t r[4];
    PhiloxDraw(uint64_t seed, uint32_t tid_, uint32_t stream_, uint64_t attempt0)
        : tid(tid_), stream(stream_), attempt(attempt0) { key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32); }
    inline void begin() {
        uint32_t ctr[4] = {(uint32_t)attempt, (uint32_t)(attempt >> 32), tid, stream};
        Philox::gen(ctr, key, r);
        ++attempt;
    }
    inline uint64_t r01() const { return ((uint64_t)r[1] << 32) | r[0]; }
    inline uint64_t r23() const { return ((uint64_t)r[3] << 32) | r[2]; }
    inline uint64_t step(uint64_t S) { return mulhi64(r01(), S); }
    inline bool coin_zipf() { return r[2] & 1u; }
    inline bool coin_back() { return (r[2] >> 1) & 1u; }
    inline double unit() { return (double)(r23() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t rank(uint64_t n) { return mulhi64(r23(), n); }
    inline bool coin_end_a() { return (r[2] >> 2) & 1u; }
    inline bool coin_end_b() { return (r[2] >> 3) & 1u; }
};

struct Term {   // everything one sampled term determines before positions are touched
    uint64_t step_a, step_b; uint64_t handle_a, handle_b;
    double pos_a, pos_b; bool other_a, other_b;
};

// sgd.rs:444-510 (1D) / :990-1077 (nD): sample one term.  Returns false for "continue".
template <class Draw>
inline bool sample_term(const PathIndex& ix, const GraphView* g_for_len, bool nd, bool cooling,
                        double current_theta, uint64_t space, uint64_t space_max, uint64_t q,
                        const std::vector<double>& zetas, Draw& d, Term& t) {
    d.begin();
    uint64_t step_idx = d.step(ix.get_total_steps());
    uint64_t path_idx = ix.get_path_of_step(step_idx);
    uint64_t path_step_count = ix.get_path_step_count(path_idx);
    if (path_step_count == 1) return false;
    uint64_t rank_a = ix.get_rank_of_step(step_idx);
    uint64_t rank_b = rank_a;
    if (cooling || d.coin_zipf()) {
        if (rank_a > 0 && (d.coin_back() || rank_a == path_step_count - 1)) {
            uint64_t jump_space = std::min(space, rank_a);
            uint64_t space_idx = jump_space > space_max ? space_max + (jump_space - space_max) / q + 1 : jump_space;
            space_idx = std::min<uint64_t>(space_idx, zetas.size() - 1);
            double z2 = 1.0 + fast_precise_pow(0.5, current_theta);
            uint64_t z_i = dirty_zipf(1, jump_space, current_theta, zetas[space_idx], z2, d.unit());
            rank_b = rank_a >= z_i ? rank_a - z_i : 0;                     // saturating_sub
        } else if (rank_a < path_step_count - 1) {
            uint64_t jump_space = std::min(space, path_step_count - rank_a - 1);
            uint64_t space_idx = jump_space > space_max ? space_max + (jump_space - space_max) / q + 1 : jump_space;
            space_idx = std::min<uint64_t>(space_idx, zetas.size() - 1);
            double z2 = 1.0 + fast_precise_pow(0.5, current_theta);
            uint64_t z_i = dirty_zipf(1, jump_space, current_theta, zetas[space_idx], z2, d.unit());
            rank_b = std::min(rank_a + z_i, path_step_count - 1);
        }
    } else {
        rank_b = d.rank(path_step_count);
    }
    if (rank_a == rank_b) return false;
    t.step_a = ix.get_step_at_path_position(path_idx, rank_a);
    t.step_b = ix.get_step_at_path_position(path_idx, rank_b);
    t.handle_a = ix.get_handle_of_step(t.step_a);
    t.handle_b = ix.get_handle_of_step(t.step_b);
    t.pos_a = (double)ix.get_position_of_step(t.step_a);
    t.pos_b = (double)ix.get_position_of_step(t.step_b);
    t.other_a = t.other_b = false;
    if (nd) {                                                             // sgd.rs:1051-1077
        double len_i = (double)g_for_len->len(t.handle_a >> 1);
        double len_j = (double)g_for_len->len(t.handle_b >> 1);
        bool rev_a = t.handle_a & 1, rev_b = t.handle_b & 1;
        bool ua = d.coin_end_a();
        if (ua) { t.pos_a += len_i; ua = !rev_a; } else { ua = rev_a; }
        bool ub = d.coin_end_b();
        if (ub) { t.pos_b += len_j; ub = !rev_b; } else { ub = rev_b; }
        t.other_a = ua; t.other_b = ub;
    }
    return true;
}

inline double load_f64(const std::atomic<uint64_t>& a) { return bits_f64(a.load(std::memory_order_relaxed)); }
inline void store_f64(std::atomic<uint64_t>& a, double v) { a.store(f64_bits(v), std::memory_order_relaxed); }

// sgd.rs:512-576: apply one 1D term.  Returns false for "continue" (term_dist == 0 / unknown handle).
inline bool apply_1d(const Term& t, double eta, const HandleMap& h2i, std::atomic<uint64_t>* X,
                     Control* ctl) {
    double term_dist = std::fabs(t.pos_a - t.pos_b);
    if (term_dist == 0.0) return false;
    double term_weight = 1.0 / term_dist;
    double mu = eta * term_weight;
    mu = std::fmin(mu, 1.0);                                              // Rust f64::min: NaN -> other operand
    uint64_t i, j;
    if (!h2i.get((t.handle_a >> 1) << 1, i)) return false;
    if (!h2i.get((t.handle_b >> 1) << 1, j)) return false;
    double x_i = load_f64(X[i]), x_j = load_f64(X[j]);
    double dx = x_i - x_j;
    if (dx == 0.0) dx = 1e-9;
    double mag = std::fabs(dx);
    double delta_update = mu * (mag - term_dist) / 2.0;
    if (ctl) {                                                            // :554-567 (dead, but costed)
        double delta_abs = std::fabs(delta_update);
        uint64_t cur = ctl->delta_max.load(std::memory_order_relaxed);
        while (delta_abs > bits_f64(cur)) {
            if (ctl->delta_max.compare_exchange_weak(cur, f64_bits(delta_abs), std::memory_order_relaxed)) break;
        }
    }
    double r = delta_update / mag;
    double r_x = r * dx;
    store_f64(X[i], load_f64(X[i]) - r_x);
    store_f64(X[j], load_f64(X[j]) + r_x);
    return true;
}

// sgd.rs:1079-1149: apply one nD term on coords[d][2*idx+end].
inline bool apply_nd(const Term& t, double eta, const HandleMap& h2i, uint64_t dims,
                     std::vector<std::vector<std::atomic<uint64_t>>>& C, Control* ctl, double* deltas) {
    double term_dist = std::fabs(t.pos_a - t.pos_b);
    if (term_dist == 0.0) return false;
    double term_weight = 1.0 / term_dist;
    double mu = std::fmin(eta * term_weight, 1.0);
    uint64_t i, j;
    if (!h2i.get((t.handle_a >> 1) << 1, i)) return false;
    if (!h2i.get((t.handle_b >> 1) << 1, j)) return false;
    uint64_t idx_i = i * 2 + (t.other_a ? 1 : 0), idx_j = j * 2 + (t.other_b ? 1 : 0);
    double mag_sq = 0.0;
    for (uint64_t d = 0; d < dims; ++d) {
        deltas[d] = load_f64(C[d][idx_i]) - load_f64(C[d][idx_j]);
        mag_sq += deltas[d] * deltas[d];
    }
    if (mag_sq == 0.0) { deltas[0] = 1e-9; mag_sq = 1e-18; }
    double mag = std::sqrt(mag_sq);
    double delta_update = mu * (mag - term_dist) / 2.0;
    if (ctl) {
        double delta_abs = std::fabs(delta_update);
        uint64_t cur = ctl->delta_max.load(std::memory_order_relaxed);
        while (delta_abs > bits_f64(cur)) {
            if (ctl->delta_max.compare_exchange_weak(cur, f64_bits(delta_abs), std::memory_order_relaxed)) break;
        }
    }
    double r = delta_update / mag;
    for (uint64_t d = 0; d < dims; ++d) {
        double r_d = r * deltas[d];
        double c_i = load_f64(C[d][idx_i]), c_j = load_f64(C[d][idx_j]);
        store_f64(C[d][idx_i], c_i - r_d);
        store_f64(C[d][idx_j], c_j + r_d);
    }
    return true;
}

struct Barrier {   // reusable barrier for the exact-count mode (C++17)
    std::mutex m; std::condition_variable cv; uint64_t n, count = 0, gen = 0;
    explicit Barrier(uint64_t n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        uint64_t g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

enum { MODE_REFERENCE = 0, MODE_EXACT = 1 };
enum { DRAW_XOSHIRO = 0, DRAW_PHILOX = 1 };

// bench.py's bounded CPU samples run single epochs of the reference's schedule (same eta, same cooling state as
// the epoch the GPU arm times): epochs [g_epoch_begin, min(g_epoch_end, iter_max + 1)) of the run are executed,
// the schedule itself (etas, first cooling epoch) is always the full one.  Default: the whole run, i.e. the reference.
static uint64_t g_epoch_begin = 0, g_epoch_end = ~0ull;

// One SGD run, 1D (dims == 0 -> X) or nD (dims >= 1 -> C).  mode REFERENCE = checker thread polling
// every 1 ms + free-running workers (sgd.rs:355-593 / 912-1164); mode EXACT = every epoch applies
// exactly min_term_updates updates (thread t does floor(M/T) + (t < M%T)), barrier between epochs.
template <class Draw>
void run_sgd(const GraphView& g, const Params& P, int mode, uint64_t dims, uint32_t philox_stream,
             const uint64_t* philox_tid_base, std::atomic<uint64_t>* X,
             std::vector<std::vector<std::atomic<uint64_t>>>* C, const PathIndex& ix,
             const HandleMap& h2i, Stats* st) {
    const bool nd = dims > 0;
    uint64_t first_cooling_iteration = (uint64_t)std::floor(P.cooling_start * (double)P.iter_max);  // :297
    std::vector<double> etas(P.iter_max + 1);
    schedule(1.0 / P.eta_max, 1.0, P.iter_max, P.iter_with_max_learning_rate, P.eps, etas.data());
    // zeta table; iterations capped at the largest reachable jump_space (SURVEY.md §7 "zeta table size")
    uint64_t max_steps = 0;
    for (auto& pi : ix.paths) max_steps = std::max(max_steps, pi.step_count);
    uint64_t zsz = zeta_size(P.space, P.space_max, P.space_quantization_step);
    uint64_t reach = std::min(P.space, max_steps);
    uint64_t reach_idx = reach > P.space_max ? P.space_max + (reach - P.space_max) / P.space_quantization_step + 1 : reach;
    uint64_t zalloc = std::min(zsz, reach_idx + 2);
    std::vector<double> zetas(zalloc);
    zetas_fill(P.space, P.space_max, P.space_quantization_step, P.theta, zetas.data(), zalloc, reach);
    // NOTE: the clamp `space_idx.min(zetas.len()-1)` (sgd.rs:469) can only bind when the table is
    // full-size; with the capped table every reachable index is < zalloc-1, so results are identical.

    const uint64_t e_begin = std::min<uint64_t>(g_epoch_begin, P.iter_max), e_end = std::min<uint64_t>(g_epoch_end, P.iter_max + 1);
    Control ctl;
    ctl.iteration.store(e_begin);
    ctl.eta_bits.store(f64_bits(etas[e_begin]));
    ctl.theta_bits.store(f64_bits(e_begin > first_cooling_iteration ? 0.001 : P.theta));
    ctl.cooling.store(e_begin > first_cooling_iteration);
    std::atomic<uint64_t> applied_total{0}, attempts_total{0};
    auto t0 = std::chrono::steady_clock::now();
    const uint64_t T = P.nthreads;

    auto make_draw = [&](uint64_t tid) -> Draw {
        if constexpr (std::is_same<Draw, XoshiroDraw>::value) return XoshiroDraw(P.seed + tid);   // :431-432
        else return PhiloxDraw(P.seed, (uint32_t)((philox_tid_base ? *philox_tid_base : 0) + tid), philox_stream, 0);
    };

    if (mode == MODE_REFERENCE) {
        std::thread checker([&] {                                           // :366-407
            while (ctl.work_todo.load(std::memory_order_relaxed)) {
                uint64_t cur = ctl.term_updates.load(std::memory_order_relaxed);
                if (cur >= P.min_term_updates) {
                    uint64_t new_iter = ctl.iteration.fetch_add(1, std::memory_order_relaxed) + 1;
                    if (new_iter > P.iter_max || new_iter >= e_end) {
                        ctl.work_todo.store(false, std::memory_order_relaxed);
                    } else {
                        if (new_iter < etas.size()) ctl.eta_bits.store(f64_bits(etas[new_iter]), std::memory_order_relaxed);
                        if (new_iter > first_cooling_iteration) {
                            ctl.theta_bits.store(f64_bits(0.001), std::memory_order_relaxed);
                            ctl.cooling.store(true, std::memory_order_relaxed);
                        }
                    }
                    ctl.term_updates.store(0, std::memory_order_relaxed);
                }
                std::this_thread::sleep_for(std::chrono::milliseconds(1));
            }
        });
        std::vector<std::thread> workers;
        for (uint64_t tid = 0; tid < T; ++tid) {
            workers.emplace_back([&, tid] {                                 // :429-590
                Draw d = make_draw(tid);
                std::vector<double> deltas(std::max<uint64_t>(dims, 1));
                uint64_t local = 0, applied = 0, attempts = 0;
                Term t;
                while (ctl.work_todo.load(std::memory_order_relaxed)) {
                    ++attempts;
                    bool cooling = ctl.cooling.load(std::memory_order_relaxed);
                    double theta = bits_f64(ctl.theta_bits.load(std::memory_order_relaxed));
                    if (!sample_term(ix, &g, nd, cooling, theta, P.space, P.space_max,
                                     P.space_quantization_step, zetas, d, t)) continue;
                    double eta = bits_f64(ctl.eta_bits.load(std::memory_order_relaxed));
                    bool ok = nd ? apply_nd(t, eta, h2i, dims, *C, &ctl, deltas.data())
                                 : apply_1d(t, eta, h2i, X, &ctl);
                    if (!ok) continue;
                    ++applied;
                    if (++local >= 1000) { ctl.term_updates.fetch_add(local, std::memory_order_relaxed); local = 0; }
                }
                if (local) ctl.term_updates.fetch_add(local, std::memory_order_relaxed);
                applied_total += applied; attempts_total += attempts;
            });
        }
        for (auto& w : workers) w.join();
        ctl.work_todo.store(false);
        checker.join();
    } else {
        Barrier bar(T ? T : 1);
        std::vector<std::thread> workers;
        for (uint64_t tid = 0; tid < T; ++tid) {
            workers.emplace_back([&, tid] {
                Draw d = make_draw(tid);
                std::vector<double> deltas(std::max<uint64_t>(dims, 1));
                uint64_t applied = 0, attempts = 0;
                Term t;
                for (uint64_t it = e_begin; it < e_end; ++it) {
                    double eta = etas[it];
                    bool cooling = it > first_cooling_iteration;
                    double theta = cooling ? 0.001 : P.theta;
                    uint64_t quota = P.min_term_updates / T + (tid < P.min_term_updates % T ? 1 : 0);
                    uint64_t done = 0;
                    while (done < quota) {
                        ++attempts;
                        if (!sample_term(ix, &g, nd, cooling, theta, P.space, P.space_max,
                                         P.space_quantization_step, zetas, d, t)) continue;
                        bool ok = nd ? apply_nd(t, eta, h2i, dims, *C, nullptr, deltas.data())
                                     : apply_1d(t, eta, h2i, X, nullptr);
                        if (ok) ++done;
                    }
                    applied += done;
                    if (T > 1) bar.wait();
                }
                applied_total += applied; attempts_total += attempts;
            });
        }
        for (auto& w : workers) w.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (st) {
        st->applied = applied_total; st->attempts = attempts_total;
        st->seconds = std::chrono::duration<double>(t1 - t0).count();
        st->epochs = e_end - e_begin;
    }
}

HandleMap build_h2i(const GraphView& g, const std::vector<uint64_t>& ids, bool live_only) {
    HandleMap m; m.init(ids.size() + 1);
    uint64_t idx = 0;
    for (uint64_t k = 0; k < ids.size(); ++k) {
        if (live_only) {                       // 1D: idx advances only for live nodes (sgd.rs:286-293)
            if (g.has(ids[k])) { m.insert(ids[k] << 1, idx); ++idx; }
        } else {                               // nD / stress: enumerate positions (sgd.rs:811-814)
            m.insert(ids[k] << 1, k);
        }
    }
    return m;
}

}  // namespace

// =============================================================================================
// C interface for the Python harness (ctypes).  All arrays caller-owned.
// =============================================================================================
extern "C" {

struct oracle_graph {
    const uint8_t* present; const uint64_t* seq_len; uint64_t nodes_len;
    const uint64_t* node_order; uint64_t node_order_len;
    const uint64_t* steps; const uint64_t* path_first; uint64_t num_paths;
};
struct oracle_params {
    uint64_t iter_max, iter_with_max_learning_rate, min_term_updates;
    double delta, eps, eta_max, theta;
    uint64_t space, space_max, space_quantization_step;
    double cooling_start;
    uint64_t nthreads, progress, seed;
};
struct oracle_stats { uint64_t applied, attempts; double seconds; uint64_t epochs; };

static GraphView view(const oracle_graph* g) {
    return GraphView{g->present, g->seq_len, g->nodes_len, g->node_order, g->node_order_len,
                     g->steps, g->path_first, g->num_paths};
}
static Params conv(const oracle_params* p) {
    return Params{p->iter_max, p->iter_with_max_learning_rate, p->min_term_updates, p->delta, p->eps,
                  p->eta_max, p->theta, p->space, p->space_max, p->space_quantization_step,
                  p->cooling_start, p->nthreads, p->progress, p->seed};
}

void oracle_set_epoch_window(uint64_t epoch_begin, uint64_t epoch_end) { g_epoch_begin = epoch_begin; g_epoch_end = epoch_end; }
double oracle_fast_precise_pow(double a, double b) { return fast_precise_pow(a, b); }
uint64_t oracle_dirty_zipf(uint64_t zmin, uint64_t zmax, double theta, double zeta, double zeta2theta, double u) {
    return dirty_zipf(zmin, zmax, theta, zeta, zeta2theta, u);
}
void oracle_schedule(double w_min, double w_max, uint64_t iter_max, uint64_t iter_with_max_lr, double eps, double* etas) {
    schedule(w_min, w_max, iter_max, iter_with_max_lr, eps, etas);
}
uint64_t oracle_zeta_size(uint64_t space, uint64_t space_max, uint64_t q) { return zeta_size(space, space_max, q); }
void oracle_zetas(uint64_t space, uint64_t space_max, uint64_t q, double theta, double* out, uint64_t size, uint64_t iter_cap) {
    zetas_fill(space, space_max, q, theta, out, size, iter_cap);
}
void oracle_philox4x32_10(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { Philox::gen(ctr, key, out); }
void oracle_xoshiro_u64(uint64_t seed, uint64_t n, uint64_t* out) {
    auto r = Xoshiro256Plus::seed_from_u64(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.next_u64();
}
void oracle_splitmix64(uint64_t seed, uint64_t n, uint64_t* out) {
    uint64_t x = seed;
    for (uint64_t i = 0; i < n; ++i) out[i] = splitmix64_next(x);
}
void oracle_xoshiro_from_state(const uint64_t* state4, uint64_t n, uint64_t* out) {
    auto r = Xoshiro256Plus::from_state(state4);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.next_u64();
}
void oracle_xoshiro_f64(uint64_t seed, uint64_t n, double* out) {
    auto r = Xoshiro256Plus::seed_from_u64(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.next_f64();
}
void oracle_xoshiro_below(uint64_t seed, uint64_t bound, uint64_t n, uint64_t* out) {
    auto r = Xoshiro256Plus::seed_from_u64(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = r.below(bound);
}

// PathIndex::from_graph (sgd.rs:34-71) -> the four per-step arrays + per-path info.
void oracle_path_index(const oracle_graph* og, uint64_t* step_to_handle, uint64_t* step_to_position,
                       uint64_t* step_to_path, uint64_t* step_to_rank, uint64_t* path_step_count,
                       uint64_t* path_length, uint64_t* path_first_step) {
    GraphView g = view(og);
    PathIndex ix = PathIndex::from_graph(g);
    uint64_t S = ix.get_total_steps();
    if (step_to_handle) std::memcpy(step_to_handle, ix.step_to_handle.data(), S * 8);
    if (step_to_position) std::memcpy(step_to_position, ix.step_to_position.data(), S * 8);
    if (step_to_path) std::memcpy(step_to_path, ix.step_to_path.data(), S * 8);
    if (step_to_rank) std::memcpy(step_to_rank, ix.step_to_rank.data(), S * 8);
    for (uint64_t p = 0; p < ix.num_paths(); ++p) {
        if (path_step_count) path_step_count[p] = ix.paths[p].step_count;
        if (path_length) path_length[p] = ix.paths[p].length;
        if (path_first_step) path_first_step[p] = ix.paths[p].first_step;
    }
}

// YgsParams::from_graph (ygs.rs:50-92) and LayoutSGDParams::from_graph (sgd.rs:733-763).
void oracle_params_from_graph(const oracle_graph* og, int layout, uint64_t nthreads, oracle_params* out) {
    GraphView g = view(og);
    PathIndex ix = PathIndex::from_graph(g);
    uint64_t sum = 0, max_steps = 0, max_len = 0;
    for (uint64_t p = 0; p < ix.num_paths(); ++p) {
        sum += ix.get_path_step_count(p);
        max_steps = std::max(max_steps, ix.get_path_step_count(p));
        max_len = std::max(max_len, ix.get_path_length(p));
    }
    out->iter_with_max_learning_rate = 0; out->delta = 0.0; out->eps = 0.01; out->theta = 0.99;
    out->space_quantization_step = 100; out->cooling_start = 0.5; out->nthreads = nthreads;
    out->progress = 0; out->seed = 9399220;
    out->eta_max = (double)(max_steps * max_steps);
    if (!layout) { out->iter_max = 100; out->min_term_updates = sum; out->space = max_len; out->space_max = 100; }
    else { out->iter_max = 30; out->min_term_updates = 10 * sum; out->space = max_steps; out->space_max = 1000; }
}

// X init (sgd.rs:264-294): prefix sum of node lengths in node_ids order, live nodes only.
// Returns the number of entries written (== live nodes in node_ids).
uint64_t oracle_init_x(const oracle_graph* og, double* x) {
    GraphView g = view(og);
    auto ids = node_ids_of(g);
    uint64_t len = 0, idx = 0;
    for (uint64_t id : ids) if (g.has(id)) { x[idx++] = (double)len; len += g.seq_len[id]; }
    return idx;
}

// Layout init (sgd.rs:816-854), written in Layout order coords[node*2*D + end*D + dim] (layout.rs:52-61).
void oracle_init_layout(const oracle_graph* og, uint64_t dims, uint64_t seed, double* coords) {
    GraphView g = view(og);
    auto ids = node_ids_of(g);
    uint64_t num_nodes = g.node_count();
    Xoshiro256Plus rng = Xoshiro256Plus::seed_from_u64(seed);
    uint64_t len = 0;
    double sqrt_n = std::sqrt((double)num_nodes * 2.0);
    for (uint64_t idx = 0; idx < ids.size(); ++idx) {
        if (!g.has(ids[idx])) continue;
        uint64_t nl = g.seq_len[ids[idx]];
        double* plus = coords + idx * 2 * dims; double* minus = plus + dims;
        plus[0] = (double)len;
        for (uint64_t d = 1; d < dims; ++d) plus[d] = rng.normal() * sqrt_n;
        minus[0] = (double)(len + nl);
        for (uint64_t d = 1; d < dims; ++d) minus[d] = rng.normal() * sqrt_n;
        len += nl;
    }
}

// A prebuilt PathIndex + handle map, so that bench.py's repeated baseline steps time the SGD loop
// and not the (serial) index construction.  The reference rebuilds the index on every call
// (sgd.rs:247); reusing it only makes this baseline faster than the reference.
struct oracle_index { PathIndex ix; HandleMap h2i_live; };
void* oracle_index_create(const oracle_graph* og) {
    GraphView g = view(og);
    oracle_index* oi = new oracle_index();
    oi->ix = PathIndex::from_graph(g);
    oi->h2i_live = build_h2i(g, node_ids_of(g), true);
    return oi;
}
void oracle_index_free(void* p) { delete (oracle_index*)p; }

static int sgd_1d_with_index(const GraphView& g, const Params& P, const PathIndex& ix, const HandleMap& h2i, int mode,
                             int draw, uint32_t philox_stream, uint64_t philox_tid_base, double* x_inout, uint64_t n_x,
                             oracle_stats* st) {
    bool valid = false;
    for (auto& pi : ix.paths) if (pi.step_count > 1) { valid = true; break; }   // :250-261
    if (!valid) return 1;
    std::vector<std::atomic<uint64_t>> X(n_x);
    for (uint64_t i = 0; i < n_x; ++i) X[i].store(f64_bits(x_inout[i]));
    oracle_stats dummy; Stats s{};
    if (draw == DRAW_XOSHIRO) run_sgd<XoshiroDraw>(g, P, mode, 0, 0, nullptr, X.data(), nullptr, ix, h2i, &s);
    else run_sgd<PhiloxDraw>(g, P, mode, 0, philox_stream, &philox_tid_base, X.data(), nullptr, ix, h2i, &s);
    for (uint64_t i = 0; i < n_x; ++i) x_inout[i] = bits_f64(X[i].load());
    if (!st) st = &dummy;
    st->applied = s.applied; st->attempts = s.attempts; st->seconds = s.seconds; st->epochs = s.epochs;
    return 0;
}

// path_linear_sgd (sgd.rs:237-614).  x_inout: N doubles (in: init, out: final), N = live nodes.
// mode 0 = reference (checker thread), 1 = exact-count epochs.  draw 0 = xoshiro, 1 = philox.
int oracle_path_linear_sgd(const oracle_graph* og, const oracle_params* op, int mode, int draw,
                           uint32_t philox_stream, uint64_t philox_tid_base, double* x_inout,
                           uint64_t n_x, oracle_stats* st) {
    GraphView g = view(og); Params P = conv(op);
    PathIndex ix = PathIndex::from_graph(g);                      // :247
    HandleMap h2i = build_h2i(g, node_ids_of(g), true);
    return sgd_1d_with_index(g, P, ix, h2i, mode, draw, philox_stream, philox_tid_base, x_inout, n_x, st);
}
int oracle_path_linear_sgd_ix(void* index, const oracle_graph* og, const oracle_params* op, int mode, int draw,
                              uint32_t philox_stream, uint64_t philox_tid_base, double* x_inout,
                              uint64_t n_x, oracle_stats* st) {
    oracle_index* oi = (oracle_index*)index;
    return sgd_1d_with_index(view(og), conv(op), oi->ix, oi->h2i_live, mode, draw, philox_stream, philox_tid_base,
                             x_inout, n_x, st);
}

// path_linear_sgd_layout (sgd.rs:773-1188).  coords_inout in Layout order (N*2*dims doubles).
int oracle_path_linear_sgd_layout(const oracle_graph* og, const oracle_params* op, uint64_t dims, int mode,
                                  int draw, uint32_t philox_stream, uint64_t philox_tid_base,
                                  double* coords_inout, uint64_t num_nodes, oracle_stats* st) {
    GraphView g = view(og); Params P = conv(op);
    PathIndex ix = PathIndex::from_graph(g);                      // :784
    bool valid = false;
    for (auto& pi : ix.paths) if (pi.step_count > 1) { valid = true; break; }
    if (!valid) return 1;
    auto ids = node_ids_of(g);
    HandleMap h2i = build_h2i(g, ids, false);
    std::vector<std::vector<std::atomic<uint64_t>>> C(dims);      // coords[dim][2*node+end] (:823-825)
    for (uint64_t d = 0; d < dims; ++d) {
        C[d] = std::vector<std::atomic<uint64_t>>(num_nodes * 2);
        for (uint64_t n = 0; n < num_nodes; ++n)
            for (uint64_t e = 0; e < 2; ++e)
                C[d][n * 2 + e].store(f64_bits(coords_inout[n * 2 * dims + e * dims + d]));
    }
    oracle_stats dummy; Stats s{};
    if (draw == DRAW_XOSHIRO) run_sgd<XoshiroDraw>(g, P, mode, dims, 0, nullptr, nullptr, &C, ix, h2i, &s);
    else run_sgd<PhiloxDraw>(g, P, mode, dims, philox_stream, &philox_tid_base, nullptr, &C, ix, h2i, &s);
    // Layout::from_vectors (layout.rs:39-69)
    for (uint64_t n = 0; n < num_nodes; ++n)
        for (uint64_t e = 0; e < 2; ++e)
            for (uint64_t d = 0; d < dims; ++d)
                coords_inout[n * 2 * dims + e * dims + d] = bits_f64(C[d][n * 2 + e].load());
    if (!st) st = &dummy;
    st->applied = s.applied; st->attempts = s.attempts; st->seconds = s.seconds; st->epochs = s.epochs;
    return 0;
}

// Trace of sampled terms under the Philox policy, for bit-exact comparison with the CUDA sampler.
// For k in [0,count): draw attempt `attempt0 + k` of thread `tid`; out_valid[k] = 0 if the reference
// loop would `continue` before touching positions, else 1 with step_a/step_b/flags/term_dist.
void oracle_trace_terms(const oracle_graph* og, const oracle_params* op, int nd, int cooling, double theta_cur,
                        uint32_t stream, uint32_t tid, uint64_t attempt0, uint64_t count,
                        uint8_t* out_valid, uint64_t* out_step_a, uint64_t* out_step_b, uint8_t* out_flags,
                        double* out_dist) {
    GraphView g = view(og); Params P = conv(op);
    PathIndex ix = PathIndex::from_graph(g);
    uint64_t max_steps = 0;
    for (auto& pi : ix.paths) max_steps = std::max(max_steps, pi.step_count);
    uint64_t zsz = zeta_size(P.space, P.space_max, P.space_quantization_step);
    std::vector<double> zetas(zsz);
    zetas_fill(P.space, P.space_max, P.space_quantization_step, P.theta, zetas.data(), zsz, std::min(P.space, max_steps));
    PhiloxDraw d(P.seed, tid, stream, attempt0);
    Term t;
    for (uint64_t k = 0; k < count; ++k) {
        bool ok = sample_term(ix, &g, nd != 0, cooling != 0, theta_cur, P.space, P.space_max,
                              P.space_quantization_step, zetas, d, t);
        double dist = ok ? std::fabs(t.pos_a - t.pos_b) : 0.0;
        if (ok && dist == 0.0) ok = false;
        out_valid[k] = ok;
        out_step_a[k] = ok ? t.step_a : 0; out_step_b[k] = ok ? t.step_b : 0;
        out_flags[k] = ok ? (uint8_t)((t.other_a ? 1 : 0) | (t.other_b ? 2 : 0)) : 0;
        out_dist[k] = dist;
    }
}

// calculate_layout_stress (sgd.rs:1196-1283).  coords in Layout order with `dims` dims; + ends.
// draw 0: xoshiro256+(12345) exactly as the reference.  draw 1: Philox(seed, tid = sample index,
// stream) — the GPU library's sample, for bit-level comparison.  Also returns mean |dl-dp|/dp
// (BASELINE.json's form) in *mean_abs_rel.
double oracle_layout_stress(const oracle_graph* og, uint64_t dims, const double* coords, uint64_t sample_count,
                            int draw, uint64_t seed, uint32_t stream, double* mean_abs_rel, uint64_t* counted) {
    GraphView g = view(og);
    PathIndex ix = PathIndex::from_graph(g);
    auto ids = node_ids_of(g);
    HandleMap h2i = build_h2i(g, ids, false);
    uint64_t total = ix.get_total_steps();
    if (total < 2) { if (mean_abs_rel) *mean_abs_rel = 0; if (counted) *counted = 0; return 0.0; }
    Xoshiro256Plus rng = Xoshiro256Plus::seed_from_u64(12345);
    double sum = 0.0, sum_abs = 0.0; uint64_t count = 0;
    for (uint64_t k = 0; k < sample_count; ++k) {
        uint64_t step_a, rank_b_draw_src = 0; uint32_t r[4] = {0, 0, 0, 0};
        if (draw == DRAW_XOSHIRO) step_a = rng.below(total);
        else {
            uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
            uint32_t ctr[4] = {(uint32_t)k, (uint32_t)(k >> 32), 0, stream};
            Philox::gen(ctr, key, r);
            step_a = mulhi64(((uint64_t)r[1] << 32) | r[0], total);
            rank_b_draw_src = ((uint64_t)r[3] << 32) | r[2];
        }
        uint64_t p = ix.get_path_of_step(step_a);
        uint64_t n = ix.get_path_step_count(p);
        if (n < 2) continue;
        uint64_t rank_a = ix.get_rank_of_step(step_a);
        uint64_t rank_b = draw == DRAW_XOSHIRO ? rng.below(n) : mulhi64(rank_b_draw_src, n);
        if (rank_a == rank_b) continue;
        uint64_t sa = ix.get_step_at_path_position(p, rank_a), sb = ix.get_step_at_path_position(p, rank_b);
        uint64_t ha = ix.get_handle_of_step(sa), hb = ix.get_handle_of_step(sb);
        double pa = (double)ix.get_position_of_step(sa), pb = (double)ix.get_position_of_step(sb);
        double path_dist = std::fabs(pa - pb);
        if (path_dist == 0.0) continue;
        uint64_t ia, ib;
        if (!h2i.get((ha >> 1) << 1, ia)) continue;
        if (!h2i.get((hb >> 1) << 1, ib)) continue;
        double sq = 0.0;                                          // Layout::distance (layout.rs:126-133)
        for (uint64_t d = 0; d < dims; ++d) {
            double dl = coords[ia * 2 * dims + d] - coords[ib * 2 * dims + d];
            sq += dl * dl;
        }
        double layout_dist = std::sqrt(sq);
        double err = layout_dist - path_dist;
        sum += (err * err) / (path_dist * path_dist);
        sum_abs += std::fabs(err) / path_dist;
        ++count;
    }
    if (mean_abs_rel) *mean_abs_rel = count ? sum_abs / (double)count : 0.0;
    if (counted) *counted = count;
    return count ? std::sqrt(sum / (double)count) : 0.0;
}

// path_sgd_sort's ordering step (sgd.rs:659-671): stable sort of idx by position, NaN compares
// Equal.  Ties: the reference's order is HashMap-iteration order (unspecified); idx order here.
void oracle_sort_by_position(const double* x, uint64_t n, uint64_t* order) {
    std::vector<uint64_t> idx(n);
    for (uint64_t i = 0; i < n; ++i) idx[i] = i;
    std::stable_sort(idx.begin(), idx.end(), [&](uint64_t a, uint64_t b) { return x[a] < x[b]; });
    std::memcpy(order, idx.data(), n * 8);
}

}  // extern "C"
