// gfs_host_graph.cpp — linear-time host versions of the two steps that consume `Y`'s order in the
// `Ygs` pipeline: grooming (`g`) and the heads-first topological sort (`s`).  SURVEY.md §8f-1.
//
// The reference's versions rescan the whole edge set per node (find_head_nodes,
// src/graph_ops.rs:1138-1183; groom_bfs_majority, src/groom.rs:202-275) or clone and sort it per
// processed handle (exact_odgi_topological_order, src/graph_ops.rs:1232-1485): O(N*E) and
// O(N*E log E), which cannot finish on the 1M-node graph of BASELINE.json's config 2.  These build
// the two adjacency lists once and emit EXACTLY the same order (tests/test_ygs_host.py compares them
// with a line-by-line restatement of the reference in oracle/ on the fixtures and on random
// bidirected graphs with cycles, inversions, self loops and missing nodes).
//
// Graph input, as flat arrays: present[nodes_len] (nodes: Vec<Option<BiNode>>), edges as
// (from, to) handle pairs (Handle = id << 1 | is_reverse, src/graph.rs:9-19; the reference keeps
// them in a HashSet, so they are unique), paths as one concatenated handle array + first-step table.
//
// Edge relations (the reference stores one of {edge, complement}; both are honoured everywhere):
//   e "goes to"   h  <=>  e.to == h   || e.from == flip(h)      (graph_ops.rs:1369-1374)
//   e "goes from" h  <=>  e.from == h || e.to == flip(h)        (graph_ops.rs:1377-1382)
//   next(e, h) = e.from == h ? e.to : flip(e.from)              (graph_ops.rs:1385-1392)
// Every per-handle scan of the reference walks the edges in sorted (from, to) order, so the
// adjacency lists are filled in that order.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <new>
#include <queue>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gfasort_cuda.h"

namespace gfs { void set_error(const std::string& s); }

// An exception (std::bad_alloc on a huge node id, ...) must not unwind through the C ABI.
static int gfs_host_exception(const char* where) noexcept {
    try { throw; }
    catch (const std::bad_alloc&) { try { gfs::set_error(std::string(where) + ": out of host memory"); } catch (...) {} }
    catch (const std::exception& e) { try { gfs::set_error(std::string(where) + ": " + e.what()); } catch (...) {} }
    catch (...) { try { gfs::set_error(std::string(where) + ": unknown exception"); } catch (...) {} }
    return GFS_ERR_INVALID;
}

namespace {

struct HostGraph {
    uint64_t nodes_len = 0;                 // node ids are < nodes_len
    const uint8_t* present = nullptr;
    const uint64_t* efrom = nullptr; const uint64_t* eto = nullptr; uint64_t E = 0;
    const uint64_t* steps = nullptr; const uint64_t* path_first = nullptr; uint64_t P = 0;
    uint64_t H = 0;                         // handle space: 2 * (largest node id anywhere + 1)
    std::vector<uint32_t> order;            // edge indices sorted by (from, to)
    std::vector<uint64_t> in_first, out_first;     // CSR offsets per handle (H + 1)
    std::vector<uint32_t> in_list, out_list;       // edge indices, in sorted-edge order

    bool has(uint64_t id) const { return id < nodes_len && present[id]; }

    int build() {
        if (E >= (1ull << 32)) { gfs::set_error("host graph: more than 2^32 edges"); return GFS_ERR_INVALID; }
        uint64_t max_id = nodes_len;
        for (uint64_t e = 0; e < E; ++e) max_id = std::max({max_id, (efrom[e] >> 1) + 1, (eto[e] >> 1) + 1});
        H = 2 * max_id;
        order.resize(E);
        for (uint64_t e = 0; e < E; ++e) order[e] = (uint32_t)e;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
            return efrom[a] != efrom[b] ? efrom[a] < efrom[b] : eto[a] < eto[b];
        });
        in_first.assign(H + 1, 0); out_first.assign(H + 1, 0);
        for (uint64_t e = 0; e < E; ++e) {
            const uint64_t t = eto[e], f = efrom[e];
            ++in_first[t + 1];  if ((f ^ 1) != t) ++in_first[(f ^ 1) + 1];
            ++out_first[f + 1]; if ((t ^ 1) != f) ++out_first[(t ^ 1) + 1];
        }
        for (uint64_t h = 0; h < H; ++h) { in_first[h + 1] += in_first[h]; out_first[h + 1] += out_first[h]; }
        in_list.resize(in_first[H]); out_list.resize(out_first[H]);
        std::vector<uint64_t> ip(in_first.begin(), in_first.end() - 1), op(out_first.begin(), out_first.end() - 1);
        for (uint32_t e : order) {
            const uint64_t t = eto[e], f = efrom[e];
            in_list[ip[t]++] = e;  if ((f ^ 1) != t) in_list[ip[f ^ 1]++] = e;
            out_list[op[f]++] = e; if ((t ^ 1) != f) out_list[op[t ^ 1]++] = e;
        }
        return GFS_OK;
    }

    // find_head_nodes (graph_ops.rs:1138-1183): forward handles with no edge going to them, sorted by
    // (earliest rank in any path, node id); nodes on no path sort last (usize::MAX).
    std::vector<uint64_t> heads() const {
        std::vector<uint64_t> hs;
        for (uint64_t id = 0; id < nodes_len; ++id)
            if (present[id] && in_first[(id << 1) + 1] == in_first[id << 1]) hs.push_back(id << 1);
        // earliest rank in any path (build_path_position_map :1111-1125) — needed for the heads only, and the heads
        // are few: one sequential pass over the steps against a byte map, instead of a random 8-byte read per step
        std::vector<uint8_t> is_head(nodes_len, 0);
        for (uint64_t h : hs) is_head[h >> 1] = 1;
        std::vector<uint64_t> min_pos(hs.empty() ? 0 : nodes_len, ~0ull);
        if (!hs.empty())
            for (uint64_t p = 0; p < P; ++p)
                for (uint64_t s = path_first[p]; s < path_first[p + 1]; ++s) {
                    const uint64_t id = steps[s] >> 1;
                    if (id < nodes_len && is_head[id]) {
                        const uint64_t r = s - path_first[p];
                        if (r < min_pos[id]) min_pos[id] = r;
                    }
                }
        std::stable_sort(hs.begin(), hs.end(), [&](uint64_t a, uint64_t b) {
            const uint64_t pa = min_pos[a >> 1], pb = min_pos[b >> 1];
            return pa != pb ? pa < pb : a < b;
        });
        return hs;
    }
};

int make_graph(HostGraph& g, const uint8_t* present, uint64_t nodes_len, const uint64_t* efrom, const uint64_t* eto,
               uint64_t E, const uint64_t* steps, const uint64_t* path_first, uint64_t P) {
    if ((nodes_len && !present) || (E && (!efrom || !eto)) || (P && !path_first)) {
        gfs::set_error("host graph: null input array"); return GFS_ERR_INVALID;
    }
    g.nodes_len = nodes_len; g.present = present; g.efrom = efrom; g.eto = eto; g.E = E;
    g.steps = steps; g.path_first = path_first; g.P = P;
    return g.build();
}

}  // namespace

extern "C" int gfs_find_head_nodes(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from,
                                   const uint64_t* edge_to, uint64_t E, const uint64_t* steps, const uint64_t* path_first,
                                   uint64_t P, uint64_t* heads_out, uint64_t* n_heads) try {
    HostGraph g;
    int rc = make_graph(g, present, nodes_len, edge_from, edge_to, E, steps, path_first, P);
    if (rc) return rc;
    const std::vector<uint64_t> hs = g.heads();
    if (n_heads) *n_heads = hs.size();
    if (heads_out) std::memcpy(heads_out, hs.data(), hs.size() * 8);
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_find_head_nodes"); }

// groom(use_bfs = true) (groom.rs:49-199 + groom_bfs_majority :202-275): BFS from the heads over both
// edge forms, neighbours in (node id, orientation) order; a node reached through its reverse handle is
// flipped.  Output: every present node in increasing id, as a reverse handle when flipped.
extern "C" int gfs_groom_order(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from,
                               const uint64_t* edge_to, uint64_t E, const uint64_t* steps, const uint64_t* path_first,
                               uint64_t P, uint64_t* order_out /* one per present node */, uint64_t* n_flipped) try {
    HostGraph g;
    int rc = make_graph(g, present, nodes_len, edge_from, edge_to, E, steps, path_first, P);
    if (rc) return rc;
    std::vector<uint8_t> visited(g.H / 2, 0), flipped(g.H / 2, 0);
    std::vector<uint64_t> seeds = g.heads();
    if (seeds.empty())                                        // no heads: the first node, forward (:117-128)
        for (uint64_t id = 0; id < nodes_len; ++id) if (present[id]) { seeds.push_back(id << 1); break; }
    std::vector<uint64_t> queue, next;
    uint64_t scan = 0;                                        // next candidate for "first unvisited node"
    for (;;) {
        if (seeds.empty()) {                                  // new component from the first unvisited node (:137-157)
            while (scan < nodes_len && (!present[scan] || visited[scan])) ++scan;
            if (scan >= nodes_len) break;
            seeds.push_back(scan << 1);
        }
        queue.clear();
        for (uint64_t s : seeds)
            if (!visited[s >> 1]) { queue.push_back(s); visited[s >> 1] = 1; if (s & 1) flipped[s >> 1] = 1; }
        for (size_t qi = 0; qi < queue.size(); ++qi) {
            const uint64_t cur = queue[qi];
            next.clear();
            if (cur < g.H)
                for (uint64_t k = g.out_first[cur]; k < g.out_first[cur + 1]; ++k) {
                    const uint32_t e = g.out_list[k];
                    next.push_back(edge_from[e] == cur ? edge_to[e] : (edge_from[e] ^ 1));   // direct form first (:227-236)
                }
            std::sort(next.begin(), next.end());              // by (node id, is_reverse) (:240)
            for (uint64_t nx : next)
                if (!visited[nx >> 1]) {
                    visited[nx >> 1] = 1;
                    if (nx & 1) flipped[nx >> 1] = 1;
                    queue.push_back(nx);
                }
        }
        seeds.clear();
    }
    uint64_t k = 0, nf = 0;
    for (uint64_t id = 0; id < nodes_len; ++id)
        if (present[id]) { order_out[k++] = (id << 1) | flipped[id]; nf += flipped[id]; }
    if (n_flipped) *n_flipped = nf;
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_groom_order"); }

// exact_odgi_topological_order(use_heads = true, use_tails = false) (graph_ops.rs:1232-1485): the modified
// Kahn's algorithm — heads first, the ready set processed smallest handle first, every node handled through
// its FORWARD handle, incoming edges masked only when their source has been placed, blocked successors kept
// as cycle-breaking seeds (smallest (node, orientation) first), then any unvisited handle in the same order.
extern "C" int gfs_topological_order(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from,
                                     const uint64_t* edge_to, uint64_t E, const uint64_t* steps, const uint64_t* path_first,
                                     uint64_t P, uint64_t* order_out /* one per present node */, uint64_t* n_out) try {
    HostGraph g;
    int rc = make_graph(g, present, nodes_len, edge_from, edge_to, E, steps, path_first, P);
    if (rc) return rc;
    std::vector<uint8_t> unv(g.H, 0), emitted(g.H / 2, 0), masked(E, 0);
    uint64_t n_unv = 0;
    for (uint64_t id = 0; id < nodes_len; ++id) if (present[id]) { unv[id << 1] = unv[(id << 1) | 1] = 1; n_unv += 2; }
    std::vector<uint32_t> cnt_in(g.H);                        // unmasked edges going to each handle
    for (uint64_t h = 0; h < g.H; ++h) cnt_in[h] = (uint32_t)(g.in_first[h + 1] - g.in_first[h]);
    auto mask = [&](uint32_t e) {
        masked[e] = 1;
        const uint64_t t = edge_to[e], f1 = edge_from[e] ^ 1;
        --cnt_in[t]; if (f1 != t) --cnt_in[f1];
    };
    auto take = [&](uint64_t h) { if (unv[h]) { unv[h] = 0; --n_unv; } if (unv[h ^ 1]) { unv[h ^ 1] = 0; --n_unv; } };
    std::priority_queue<uint64_t, std::vector<uint64_t>, std::greater<uint64_t>> S;     // BTreeSet<Handle>, min first
    std::set<uint64_t> seeds;                                  // Vec + contains() + sort + remove(0)
    uint64_t scan = 0;                                         // smallest possibly-unvisited handle
    uint64_t k = 0;
    for (uint64_t h : g.heads()) { S.push(h); take(h); }       // :1273-1283
    while (n_unv != 0 || !S.empty()) {
        if (S.empty()) {
            while (!seeds.empty() && S.empty()) {              // :1303-1317
                const uint64_t h = *seeds.begin();
                seeds.erase(seeds.begin());
                if (unv[h]) { S.push(h); take(h); }
            }
            if (S.empty() && n_unv != 0) {                     // :1322-1344
                while (!unv[scan]) ++scan;
                S.push(scan); take(scan);
            }
        }
        while (!S.empty()) {
            const uint64_t handle = S.top(); S.pop();
            const uint64_t fh = handle & ~1ull;                // always processed in forward orientation (:1356)
            if (!emitted[fh >> 1]) { emitted[fh >> 1] = 1; order_out[k++] = fh; }
            // incoming edges whose source node has been placed are consumed (:1396-1424)
            for (uint64_t q = g.in_first[fh]; q < g.in_first[fh + 1]; ++q) {
                const uint32_t e = g.in_list[q];
                if (masked[e]) continue;
                const uint64_t src = edge_to[e] == fh ? (edge_from[e] >> 1) : (edge_to[e] >> 1);
                if (!unv[src << 1] && !unv[(src << 1) | 1]) mask(e);
            }
            // outgoing edges (:1428-1475)
            for (uint64_t q = g.out_first[fh]; q < g.out_first[fh + 1]; ++q) {
                const uint32_t e = g.out_list[q];
                if (masked[e]) continue;
                mask(e);
                const uint64_t nx = edge_from[e] == fh ? edge_to[e] : (edge_from[e] ^ 1);
                if (nx < g.H && unv[nx]) {
                    if (cnt_in[nx] == 0) { S.push(nx); take(nx); }
                    else seeds.insert(nx);
                }
            }
        }
    }
    if (n_out) *n_out = k;
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_topological_order"); }

// Flat form of the handle rewrites that follow every pipeline step — apply_ordering (graph_ops.rs:1939-2025),
// apply_node_id_mapping (:36-84) and the orientation flips of apply_grooming_with_reorder (groom.rs:533-605) —
// over path steps or edge ends, in place: id -> new_id[id] where id < table_len and new_id[id] != UINT64_MAX
// (other handles keep their id), then orientation ^= flip[id] (flip indexed by the OLD id, may be NULL).
// The reference does this through HashMap<usize, usize> lookups per step (1e9 SipHash probes at config 3).
extern "C" int gfs_remap_handles(uint64_t* handles, uint64_t n, const uint64_t* new_id, uint64_t table_len,
                                 const uint8_t* flip, uint64_t flip_len) try {
    if (n && !handles) { gfs::set_error("gfs_remap_handles: null handles"); return GFS_ERR_INVALID; }
    if (table_len && !new_id) { gfs::set_error("gfs_remap_handles: null table"); return GFS_ERR_INVALID; }
    auto work = [&](uint64_t lo, uint64_t hi) {
        for (uint64_t i = lo; i < hi; ++i) {
            const uint64_t h = handles[i];
            uint64_t id = h >> 1, rev = h & 1;
            if (flip && id < flip_len) rev ^= (uint64_t)(flip[id] & 1);
            if (id < table_len && new_id[id] != ~0ull) id = new_id[id];
            handles[i] = (id << 1) | rev;
        }
    };
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned nt = (unsigned)std::min<uint64_t>(hw ? hw : 4, std::max<uint64_t>(n >> 20, 1));   // >= 1M handles per thread
    if (nt <= 1) { work(0, n); return GFS_OK; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
    for (auto& x : th) x.join();
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_remap_handles"); }

// The edge set a graph has when only its paths are known (synthetic inputs; what a GFA writer emitting L lines in
// path order would store): every pair of consecutive steps, one edge per {edge, complement} class as add_edge keeps
// them (graph_ops.rs:626-638), in order of first occurrence and in the form a path first walks it.
struct gfs_edge_list { std::vector<uint64_t> from, to; };

extern "C" int gfs_edges_from_paths(const uint64_t* steps, const uint64_t* path_first, uint64_t P, gfs_edge_list** out) try {
    if (!out || (P && !path_first)) { gfs::set_error("gfs_edges_from_paths: null argument"); return GFS_ERR_INVALID; }
    *out = nullptr;
    std::unique_ptr<gfs_edge_list> el(new gfs_edge_list());
    // open addressing over the canonical (smaller) form of each class; 16-byte keys, grows at 1/2 load
    struct Slot { uint64_t a, b; };
    const uint64_t EMPTY = ~0ull;                             // no handle pair is (EMPTY, EMPTY): ids are < 2^63
    std::vector<Slot> table((size_t)1 << 16, Slot{EMPTY, EMPTY});
    uint64_t used = 0;
    auto hash = [](uint64_t a, uint64_t b) {
        uint64_t x = a * 0x9e3779b97f4a7c15ULL ^ (b + 0x7f4a7c15ULL + (a << 6));
        x ^= x >> 31; x *= 0xbf58476d1ce4e5b9ULL; return x ^ (x >> 29);
    };
    auto insert = [&](std::vector<Slot>& t, uint64_t a, uint64_t b) -> bool {      // true when (a, b) was not there
        const uint64_t mask = t.size() - 1;
        for (uint64_t i = hash(a, b) & mask;; i = (i + 1) & mask) {
            if (t[i].a == a && t[i].b == b) return false;
            if (t[i].a == EMPTY && t[i].b == EMPTY) { t[i] = Slot{a, b}; return true; }
        }
    };
    for (uint64_t p = 0; p < P; ++p)
        for (uint64_t s = path_first[p]; s + 1 < path_first[p + 1]; ++s) {
            const uint64_t a = steps[s], b = steps[s + 1];
            const uint64_t ca = b ^ 1, cb = a ^ 1;            // the complement edge
            const bool swap = ca < a || (ca == a && cb < b);
            if (insert(table, swap ? ca : a, swap ? cb : b)) {
                el->from.push_back(a); el->to.push_back(b);
                if (++used * 2 > table.size()) {
                    std::vector<Slot> bigger(table.size() * 4, Slot{EMPTY, EMPTY});
                    for (const Slot& q : table) if (!(q.a == EMPTY && q.b == EMPTY)) insert(bigger, q.a, q.b);
                    table.swap(bigger);
                }
            }
        }
    *out = el.release();
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_edges_from_paths"); }

extern "C" int gfs_edge_list_get(const gfs_edge_list* el, const uint64_t** edge_from, const uint64_t** edge_to, uint64_t* n_edges) {
    if (!el) { gfs::set_error("gfs_edge_list_get: null list"); return GFS_ERR_INVALID; }
    if (edge_from) *edge_from = el->from.data();
    if (edge_to) *edge_to = el->to.data();
    if (n_edges) *n_edges = el->from.size();
    return GFS_OK;
}
extern "C" void gfs_edge_list_free(gfs_edge_list* el) { delete el; }
