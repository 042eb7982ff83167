// gfs_synth.cpp — seeded synthetic pangenome graphs in the C ABI's own flat input format
// (bench / test inputs).  Host-only, multi-threaded.  Not part of the hot path.
//
// Model: a "bubble chain".  A backbone of shared nodes is interrupted, on average every 10
// backbone nodes, by a variant site: SNP bubble (two 1-bp allele nodes, 80 %), indel (one
// optional node of 1-50 bp, 15 %) or inversion (a run of 2-20 backbone nodes that carrier
// haplotypes traverse backwards with reversed orientation, 5 %).  Each site has an alt-allele
// frequency U(0.05, 0.5).  Each of P haplotype paths walks the whole chain and picks its allele
// per site independently, from a hash of (seed, path, site) — so a path's walk does not depend on
// how many threads generated it, or on which subset of paths a rank asked for.  Node ids are
// finally permuted (Fisher-Yates) so that the file order — the SGD's initial order — is scrambled.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#ifndef GFS_SYNTH_STANDALONE
#include <cuda_runtime.h>
#endif

#include "../../include/gfasort_cuda.h"

#ifdef GFS_SYNTH_STANDALONE
// libgfs_synth.so: the generator alone, host code only (no CUDA, nothing of the product library), so that the
// CPU reference arm of bench.py and the oracle tools can make the SAME graphs without mapping libgfasort_cuda.so.
namespace gfs {
static thread_local std::string g_synth_error;
void set_error(const std::string& s) { g_synth_error = s; }
}
extern "C" const char* gfs_synth_last_error(void) { return gfs::g_synth_error.c_str(); }
#else
namespace gfs { void set_error(const std::string& s); }
#endif

namespace {

struct SplitMix64 {
    uint64_t x;
    explicit SplitMix64(uint64_t s) : x(s) {}
    inline uint64_t next() {
        x += 0x9e3779b97f4a7c15ULL;
        uint64_t z = x;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    }
    inline double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint64_t below(uint64_t n) { return (uint64_t)(((unsigned __int128)next() * n) >> 64); }
};
inline uint64_t mix3(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t z = a * 0x9e3779b97f4a7c15ULL + b * 0xbf58476d1ce4e5b9ULL + c * 0x94d049bb133111ebULL + 0x2545F4914F6CDD1DULL;
    z = (z ^ (z >> 32)) * 0xd6e8feb86659fd93ULL;
    z = (z ^ (z >> 32)) * 0xd6e8feb86659fd93ULL;
    return z ^ (z >> 32);
}

enum : uint8_t { EL_BACKBONE = 0, EL_SNP = 1, EL_INDEL = 2, EL_INVERSION = 3 };
struct Element {
    uint32_t first_node;   // pre-permutation node number
    uint16_t count;        // nodes in the element (1, 2, 1, k)
    uint8_t kind;
    uint8_t pad;
    uint64_t alt_threshold;   // haplotype carries the alt allele iff hash < threshold
};

}  // namespace

namespace {
// chain structure + pre-permutation node lengths (serial, cheap)
void build_chain(uint64_t N, uint64_t seed_in, std::vector<Element>& chain, std::vector<uint32_t>& len) {
    struct { uint64_t seed; } spec_{seed_in};
    auto* spec = &spec_;
    // ---- chain structure (serial, cheap) ----
    SplitMix64 rng(spec->seed);
    chain.reserve((size_t)(N * 0.93) + 16);
    len.assign(N, 0);
    auto backbone_len = [&]() -> uint32_t {   // 1 + Geometric(mean 31), capped at 1024
        double u = rng.unit();
        double v = std::floor(std::log(1.0 - u) / std::log(1.0 - 1.0 / 32.0));
        uint64_t l = 1 + (uint64_t)std::max(0.0, v);
        return (uint32_t)std::min<uint64_t>(l, 1024);
    };
    uint64_t next = 0;   // next unused node number
    uint64_t until_site = 1 + rng.below(19);
    while (next < N) {
        const uint64_t left = N - next;
        if (until_site == 0 && left >= 3) {
            const double kind_draw = rng.unit();
            const double f = 0.05 + 0.45 * rng.unit();
            const uint64_t thr = (uint64_t)(f * 18446744073709551615.0);
            Element e{};
            e.first_node = (uint32_t)next; e.alt_threshold = thr;
            if (kind_draw < 0.80) {
                e.kind = EL_SNP; e.count = 2;
                len[next] = 1; len[next + 1] = 1;
            } else if (kind_draw < 0.95) {
                e.kind = EL_INDEL; e.count = 1;
                len[next] = 1 + (uint32_t)rng.below(50);
            } else {
                uint64_t k = 2 + rng.below(19);
                k = std::min(k, left);
                e.kind = EL_INVERSION; e.count = (uint16_t)k;
                for (uint64_t i = 0; i < k; ++i) len[next + i] = backbone_len();
            }
            next += e.count;
            chain.push_back(e);
            until_site = 1 + rng.below(19);   // mean 10 backbone nodes between sites
        } else {
            Element e{};
            e.first_node = (uint32_t)next; e.count = 1; e.kind = EL_BACKBONE;
            len[next] = backbone_len();
            ++next;
            chain.push_back(e);
            if (until_site) --until_site;
        }
    }
}

uint64_t walk_path(const std::vector<Element>& chain, const std::vector<uint32_t>& perm, uint64_t seed, uint64_t p, uint64_t* dst) {
    uint64_t n = 0;
    const size_t ne = chain.size();
    for (size_t ei = 0; ei < ne; ++ei) {
        const Element& e = chain[ei];
        if (e.kind == EL_BACKBONE) {
            if (dst) dst[n] = (uint64_t)perm[e.first_node] << 1;
            ++n;
            continue;
        }
        const bool alt = mix3(seed, p, ei) < e.alt_threshold;
        if (e.kind == EL_SNP) {
            if (dst) dst[n] = (uint64_t)perm[e.first_node + (alt ? 1 : 0)] << 1;
            ++n;
        } else if (e.kind == EL_INDEL) {
            if (alt) { if (dst) dst[n] = (uint64_t)perm[e.first_node] << 1; ++n; }
        } else {
            if (dst) {
                if (!alt) for (uint32_t i = 0; i < e.count; ++i) dst[n + i] = (uint64_t)perm[e.first_node + i] << 1;
                else for (uint32_t i = 0; i < e.count; ++i) dst[n + i] = ((uint64_t)perm[e.first_node + e.count - 1 - i] << 1) | 1;
            }
            n += e.count;
        }
    }
    return n;
}
template <class F> void parallel_for(uint64_t n, F fn) {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 4;
    const unsigned nt = (unsigned)std::min<uint64_t>(hw, std::max<uint64_t>(n, 1));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] { for (uint64_t k = t; k < n; k += nt) fn(k); });
    for (auto& x : th) x.join();
}
}  // namespace

struct gfs_synth_graph {
    uint64_t N = 0, P = 0, S = 0;
    uint64_t path_begin = 0, path_end = 0;
    std::vector<uint32_t> node_len;
    std::vector<uint64_t> path_first;   // path_end - path_begin + 1 entries, local to the generated range
    uint64_t* steps = nullptr;          // S handles (malloc'd: avoid value-initialising tens of GB)
    bool pinned = false;                // steps came from cudaHostAlloc
#ifndef GFS_SYNTH_STANDALONE
    ~gfs_synth_graph() { if (pinned) cudaFreeHost(steps); else std::free(steps); }
#else
    ~gfs_synth_graph() { std::free(steps); }
#endif
};

extern "C" int gfs_synth_create_range(const gfs_synth_spec* spec, uint64_t path_begin, uint64_t path_end,
                                      gfs_synth_graph** out);

extern "C" int gfs_synth_create(const gfs_synth_spec* spec, gfs_synth_graph** out) {
    if (!spec) { gfs::set_error("gfs_synth_create: null spec"); return GFS_ERR_INVALID; }
    return gfs_synth_create_range(spec, 0, spec->num_paths, out);
}

extern "C" int gfs_synth_create_range(const gfs_synth_spec* spec, uint64_t path_begin, uint64_t path_end,
                                      gfs_synth_graph** out) {
    if (!spec || !out) { gfs::set_error("gfs_synth_create: null argument"); return GFS_ERR_INVALID; }
    const uint64_t N = spec->num_nodes, P = spec->num_paths;
    if (N < 4 || N >= (1ull << 31) || P == 0 || path_begin > path_end || path_end > P) {
        gfs::set_error("gfs_synth_create: need 4 <= num_nodes < 2^31, num_paths >= 1, valid path range");
        return GFS_ERR_INVALID;
    }
    gfs_synth_graph* g = new (std::nothrow) gfs_synth_graph();
    if (!g) { gfs::set_error("gfs_synth_create: out of memory"); return GFS_ERR_INVALID; }
    g->N = N; g->P = P; g->path_begin = path_begin; g->path_end = path_end;

    std::vector<Element> chain;
    std::vector<uint32_t> len;
    build_chain(N, spec->seed, chain, len);
    // ---- id permutation ----
    std::vector<uint32_t> perm(N);
    for (uint64_t i = 0; i < N; ++i) perm[i] = (uint32_t)i;
    if (spec->permute_ids) {
        SplitMix64 prng(spec->seed ^ 0xA5A5A5A5DEADBEEFULL);
        for (uint64_t i = N - 1; i > 0; --i) {
            uint64_t j = prng.below(i + 1);
            std::swap(perm[i], perm[j]);
        }
    }
    g->node_len.resize(N);
    for (uint64_t i = 0; i < N; ++i) g->node_len[perm[i]] = len[i];

    // ---- paths: count, prefix, fill (threads over paths) ----
    const uint64_t np = path_end - path_begin;
    std::vector<uint64_t> count(np, 0);
    const uint64_t seed = spec->seed;
    auto walk = [&](uint64_t p, uint64_t* dst) -> uint64_t { return walk_path(chain, perm, seed, p, dst); };
    auto parallel_paths = [&](auto fn) { parallel_for(np, fn); };
    parallel_paths([&](uint64_t k) { count[k] = walk(path_begin + k, nullptr); });
    g->path_first.assign(np + 1, 0);
    for (uint64_t k = 0; k < np; ++k) g->path_first[k + 1] = g->path_first[k] + count[k];
    g->S = g->path_first[np];
    const size_t step_bytes = std::max<uint64_t>(g->S, 1) * sizeof(uint64_t);
#ifndef GFS_SYNTH_STANDALONE
    if (spec->pinned) {
        void* ptr = nullptr;
        if (cudaHostAlloc(&ptr, step_bytes, cudaHostAllocDefault) == cudaSuccess) { g->steps = (uint64_t*)ptr; g->pinned = true; }
        else cudaGetLastError();   // no device / no pinned memory: fall back to pageable
    }
#endif
    if (!g->steps) g->steps = (uint64_t*)std::malloc(step_bytes);
    if (!g->steps) { delete g; gfs::set_error("gfs_synth_create: out of memory for steps"); return GFS_ERR_INVALID; }
    parallel_paths([&](uint64_t k) { walk(path_begin + k, g->steps + g->path_first[k]); });
    *out = g;
    return GFS_OK;
}

extern "C" int gfs_synth_dims(const gfs_synth_graph* g, uint64_t* S, uint64_t* P, uint64_t* N) {
    if (!g) { gfs::set_error("gfs_synth_dims: null graph"); return GFS_ERR_INVALID; }
    if (S) *S = g->S;
    if (P) *P = g->path_end - g->path_begin;
    if (N) *N = g->N;
    return GFS_OK;
}

extern "C" int gfs_synth_arrays(const gfs_synth_graph* g, const uint64_t** step_handles,
                                const uint64_t** path_first_step, const uint32_t** node_len) {
    if (!g) { gfs::set_error("gfs_synth_arrays: null graph"); return GFS_ERR_INVALID; }
    if (step_handles) *step_handles = g->steps;
    if (path_first_step) *path_first_step = g->path_first.data();
    if (node_len) *node_len = g->node_len.data();
    return GFS_OK;
}

extern "C" void gfs_synth_free(gfs_synth_graph* g) { delete g; }

// Step count of every path without materialising any steps (ranks of a multi-GPU run need the
// global path_first_step to pick their slice).
extern "C" int gfs_synth_path_counts(const gfs_synth_spec* spec, uint64_t* counts) {
    if (!spec || !counts) { gfs::set_error("gfs_synth_path_counts: null argument"); return GFS_ERR_INVALID; }
    const uint64_t N = spec->num_nodes, P = spec->num_paths;
    if (N < 4 || N >= (1ull << 31) || P == 0) { gfs::set_error("gfs_synth_path_counts: bad spec"); return GFS_ERR_INVALID; }
    std::vector<Element> chain;
    std::vector<uint32_t> len;
    build_chain(N, spec->seed, chain, len);
    std::vector<uint32_t> perm;   // unused when dst == nullptr
    const uint64_t seed = spec->seed;
    parallel_for(P, [&](uint64_t p) { counts[p] = walk_path(chain, perm, seed, p, nullptr); });
    return GFS_OK;
}
