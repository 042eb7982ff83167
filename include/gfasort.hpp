// gfasort.hpp — the host side of the drop-in boundary, in C++17 (header-only), above the C ABI of
// libgfasort_cuda.so (include/gfasort_cuda.h).
//
// The reference (pangenome/gfasort v0.1.0) is a Rust crate with no FFI; its maintainers would keep their
// Rust signatures and let the bodies call the C ABI (INTEGRATION.md shows that binding).  There is no Rust
// toolchain in this image, so this header restates the reference's public interface for the hot path in
// C++ — same names, same argument meaning, same error behaviour — and does exactly what the Rust side of
// the boundary would do: build the dense node numbering, flatten the paths, compute the
// initial positions, call the library, wrap the result.  All compute happens in the CUDA library; nothing
// here falls back to the CPU: where the Rust host would `panic!` on a non-zero return, this throws
// gfasort::Error.
//
//   reference item (file:line)                                  here
//   Handle, reverse_complement, BiNode, BiPath, BiEdge          graph.rs:9-200
//   BidirectedGraph {nodes, edges, paths, node_order}           graph_ops.rs:10-16, 503-510
//     add_node / add_edge / node_count / has_edge / build_path  graph_ops.rs:535, 613-638, 649, 684
//     apply_node_id_mapping / apply_ordering                    graph_ops.rs:36-84, 1939-2025
//     find_head_nodes / exact_odgi_topological_order            graph_ops.rs:1138-1183, 1232-1485  -> gfs_find_head_nodes / gfs_topological_order
//     count_edge_directions / write_gfa                         graph_ops.rs:1215-1227, 693-738
//     groom / apply_grooming_with_reorder                       groom.rs:49-275, 533-605           -> gfs_groom_order
//   gfa_parser::load_gfa / write_gfa                            gfa_parser.rs:9-170
//   PathIndex::from_graph + 9 accessors                         sgd.rs:14-107                      -> gfs_index_build / gfs_index_export
//   PathSGDParams, LayoutSGDParams(+from_graph)                 sgd.rs:196-234, 676-763
//   path_linear_sgd / path_sgd_sort                             sgd.rs:237-614, 641-672            -> gfs_sgd_1d / gfs_sgd_sort_1d
//   path_linear_sgd_layout / calculate_layout_stress            sgd.rs:773-1188, 1196-1283         -> gfs_sgd_nd / gfs_stress
//   Layout                                                      layout.rs:17-245                   (write_tsv -> gfs_layout_write_tsv)
//   YgsParams(+default, from_graph), ygs_sort, sgd_sort_only,
//   groom_only, topological_sort_only                           ygs.rs:16-206
//
// Differences from the reference that a caller can observe (all stated in DESIGN.md):
//   * the SGD draws its random numbers from Philox4x32-10 keyed by `seed`, not from xoshiro256+(seed + tid):
//     results are statistically, not bit-wise, those of the reference; `nthreads` is accepted and inert.
//   * an epoch applies exactly `min_term_updates` updates (the reference: at least, by a 1 ms polling race).
//   * path_sgd_sort breaks position ties by dense index (the reference: HashMap iteration order).
//   * the layout's Gaussian initial noise comes from xoshiro256+(seed) through a polar transform, not
//     through rand_distr's ziggurat (same distribution, different stream).
#ifndef GFASORT_HPP
#define GFASORT_HPP

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

#include "gfasort_cuda.h"

namespace gfasort {

/// Thrown where the Rust host would panic on a non-zero return of the C ABI.
class Error : public std::runtime_error {
public:
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc) {
    if (rc != GFS_OK) throw Error(rc, std::string("libgfasort_cuda: ") + gfs_last_error());
}

// ------------------------------------------------------------------------------------------------
// graph.rs
// ------------------------------------------------------------------------------------------------
/// graph.rs:9-63: node id << 1 | is_reverse.
struct Handle {
    uint64_t value = 0;
    static Handle make(size_t node_id, bool is_reverse) { return Handle{((uint64_t)node_id << 1) | (is_reverse ? 1u : 0u)}; }
    static Handle forward(size_t node_id) { return make(node_id, false); }
    static Handle reverse(size_t node_id) { return make(node_id, true); }
    static Handle from_u64(uint64_t v) { return Handle{v}; }
    size_t node_id() const { return (size_t)(value >> 1); }
    bool is_reverse() const { return (value & 1) == 1; }
    char orientation_char() const { return is_reverse() ? '-' : '+'; }
    Handle flip() const { return Handle{value ^ 1}; }
    uint64_t as_u64() const { return value; }
    std::string to_string() const { return std::to_string(node_id()) + orientation_char(); }
    bool operator==(const Handle& o) const { return value == o.value; }
    bool operator!=(const Handle& o) const { return value != o.value; }
    bool operator<(const Handle& o) const { return value < o.value; }
};

/// graph.rs:73-86.
inline std::vector<uint8_t> reverse_complement(const std::vector<uint8_t>& seq) {
    std::vector<uint8_t> out(seq.rbegin(), seq.rend());
    for (auto& b : out) {
        switch (b) {
            case 'A': case 'a': b = 'T'; break;
            case 'T': case 't': b = 'A'; break;
            case 'C': case 'c': b = 'G'; break;
            case 'G': case 'g': b = 'C'; break;
            case 'N': case 'n': b = 'N'; break;
            default: break;
        }
    }
    return out;
}

/// graph.rs:89-130.
struct BiNode {
    size_t id = 0;
    std::vector<uint8_t> sequence;
    std::optional<uint64_t> rank;
    BiNode() = default;
    BiNode(size_t i, std::vector<uint8_t> s) : id(i), sequence(std::move(s)) {}
    std::vector<uint8_t> get_sequence(bool is_reverse) const { return is_reverse ? reverse_complement(sequence) : sequence; }
};
/// graph.rs:132-175.
struct BiPath {
    std::string name;
    std::vector<Handle> steps;
    BiPath() = default;
    explicit BiPath(std::string n) : name(std::move(n)) {}
    void add_step(Handle h) { steps.push_back(h); }
};
/// graph.rs:177-200.
struct BiEdge {
    Handle from, to;
    BiEdge() = default;
    BiEdge(Handle f, Handle t) : from(f), to(t) {}
    bool operator==(const BiEdge& o) const { return from == o.from && to == o.to; }
};
struct BiEdgeHash {
    size_t operator()(const BiEdge& e) const {
        uint64_t x = e.from.value * 0x9e3779b97f4a7c15ull ^ (e.to.value + 0x7f4a7c15ull + (e.from.value << 6));
        x ^= x >> 31; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 29;
        return (size_t)x;
    }
};

// ------------------------------------------------------------------------------------------------
// graph_ops.rs — the container, and the operations the `Ygs` pipeline applies to it
// ------------------------------------------------------------------------------------------------
struct BidirectedGraph {
    std::vector<std::optional<BiNode>> nodes;                  // index = node id
    std::unordered_set<BiEdge, BiEdgeHash> edges;
    std::vector<BiPath> paths;
    std::vector<size_t> node_order;                            // add_node order (GFA file order)

    size_t node_count() const {                                // graph_ops.rs:535-537
        size_t n = 0;
        for (const auto& x : nodes) n += x.has_value();
        return n;
    }
    size_t total_sequence_length() const {                     // :527-532
        size_t n = 0;
        for (const auto& x : nodes) if (x) n += x->sequence.size();
        return n;
    }
    void add_node(size_t id, std::vector<uint8_t> sequence) {  // :613-623
        if (id >= nodes.size()) nodes.resize(id + 1);
        if (!nodes[id]) node_order.push_back(id);              // only a NEW node enters node_order
        nodes[id] = BiNode(id, std::move(sequence));
    }
    void add_node(size_t id, const std::string& sequence) { add_node(id, std::vector<uint8_t>(sequence.begin(), sequence.end())); }
    void add_edge(Handle from, Handle to) {                    // :626-638: one of {edge, complement} is stored
        BiEdge e(from, to), c(to.flip(), from.flip());
        if (!edges.count(e) && !edges.count(c)) edges.insert(e);
    }
    bool has_edge(Handle from, Handle to) const {              // :649-653
        return edges.count(BiEdge(from, to)) || edges.count(BiEdge(to.flip(), from.flip()));
    }
    std::optional<std::vector<uint8_t>> get_sequence(Handle h) const {    // :641-646
        if (h.node_id() >= nodes.size() || !nodes[h.node_id()]) return std::nullopt;
        return nodes[h.node_id()]->get_sequence(h.is_reverse());
    }
    void build_path(std::string name, const std::vector<std::pair<size_t, bool>>& steps) {   // :684-690
        BiPath p(std::move(name));
        for (auto& s : steps) p.add_step(Handle::make(s.first, s.second));
        paths.push_back(std::move(p));
    }
    std::pair<size_t, size_t> count_edge_directions() const {  // :1215-1227 (forward, backward) by node id
        size_t f = 0, b = 0;
        for (const auto& e : edges) {
            if (e.from.node_id() < e.to.node_id()) ++f;
            else if (e.from.node_id() > e.to.node_id()) ++b;
        }
        return {f, b};
    }

    /// `node_order` if non-empty, else the sorted live ids (sgd.rs:276-284, 659-668, 802-811).
    std::vector<size_t> sgd_node_ids() const {
        if (!node_order.empty()) return node_order;
        std::vector<size_t> ids;
        for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i]) ids.push_back(i);
        return ids;
    }

    /// What the host hands to the C ABI (gfasort_cuda.h "Node numbering"): dense idx over the LIVE nodes in
    /// sgd_node_ids() order (sgd.rs:286-293), step handles in dense space (idx == N: node missing from the
    /// graph, sgd.rs:52-54, 525-538), first-step table, node lengths, and the X init (cumulative length).
    struct Dense {
        std::vector<uint64_t> step_handles, path_first_step;
        std::vector<uint32_t> node_len;
        std::vector<size_t> node_id_of_idx;
        std::vector<double> x_init;
    };
    Dense dense() const {
        Dense d;
        const std::vector<size_t> ids = sgd_node_ids();
        std::vector<uint64_t> idx_of_id(nodes.size(), ~0ull);
        uint64_t len = 0;
        for (size_t id : ids) {
            if (id < nodes.size() && nodes[id] && idx_of_id[id] == ~0ull) {
                idx_of_id[id] = d.node_id_of_idx.size();
                d.node_id_of_idx.push_back(id);
                d.node_len.push_back((uint32_t)nodes[id]->sequence.size());
                d.x_init.push_back((double)len);
                len += nodes[id]->sequence.size();
            }
        }
        const uint64_t N = d.node_id_of_idx.size();
        d.path_first_step.push_back(0);
        for (const auto& p : paths) {
            for (Handle h : p.steps) {
                const size_t id = h.node_id();
                const uint64_t idx = (id < nodes.size() && idx_of_id[id] != ~0ull) ? idx_of_id[id] : N;
                d.step_handles.push_back((idx << 1) | (h.is_reverse() ? 1u : 0u));
            }
            d.path_first_step.push_back(d.step_handles.size());
        }
        return d;
    }

    /// The per-step half of every renumbering: one flat, multi-threaded pass of the library's gfs_remap_handles over
    /// each path (the reference probes a HashMap per step).  table[id] == UINT64_MAX / id >= table.size(): id kept.
    static std::vector<uint64_t> flat_table(const std::unordered_map<size_t, size_t>& mapping) {
        size_t top = 0;
        for (auto& kv : mapping) top = std::max(top, kv.first + 1);
        std::vector<uint64_t> t(top, ~0ull);
        for (auto& kv : mapping) t[kv.first] = kv.second;
        return t;
    }
    void remap_path_steps(const std::vector<uint64_t>& table, const std::vector<uint8_t>* flip) {
        static_assert(sizeof(Handle) == sizeof(uint64_t), "Handle is one u64");
        for (auto& p : paths)
            if (!p.steps.empty())
                check(gfs_remap_handles(reinterpret_cast<uint64_t*>(p.steps.data()), p.steps.size(), table.data(), table.size(),
                                        flip ? flip->data() : nullptr, flip ? flip->size() : 0));
    }

    /// graph_ops.rs:36-84.  Unmapped ids keep their id.
    void apply_node_id_mapping(const std::unordered_map<size_t, size_t>& mapping) {
        size_t max_new = 0;
        for (auto& kv : mapping) max_new = std::max(max_new, kv.second);
        std::vector<std::optional<BiNode>> nn(max_new + 1);
        auto map_id = [&](size_t id) { auto it = mapping.find(id); return it == mapping.end() ? id : it->second; };
        for (size_t old = 0; old < nodes.size(); ++old) {
            if (!nodes[old]) continue;
            const size_t nid = map_id(old);
            if (nid >= nn.size()) throw std::out_of_range("apply_node_id_mapping: unmapped node id beyond the new range");   // the reference panics here
            BiNode n = *nodes[old];
            n.id = nid;
            nn[nid] = std::move(n);
        }
        nodes = std::move(nn);
        std::unordered_set<BiEdge, BiEdgeHash> ne;
        for (const auto& e : edges)
            ne.insert(BiEdge(Handle::make(map_id(e.from.node_id()), e.from.is_reverse()), Handle::make(map_id(e.to.node_id()), e.to.is_reverse())));
        edges = std::move(ne);
        remap_path_steps(flat_table(mapping), nullptr);
    }

    /// graph_ops.rs:1939-2025: new id = rank + 1; edges with an unmapped end are dropped, steps on unmapped
    /// ids are left alone; `node_order` is NOT updated.
    void apply_ordering(const std::vector<Handle>& ordering, bool verbose = false) {
        if (ordering.empty()) return;
        std::unordered_map<size_t, size_t> old_to_new;
        for (size_t i = 0; i < ordering.size(); ++i) old_to_new[ordering[i].node_id()] = i + 1;
        size_t max_new = 0;
        for (auto& kv : old_to_new) max_new = std::max(max_new, kv.second);
        std::vector<std::optional<BiNode>> nn(max_new + 1);
        for (auto& kv : old_to_new) {
            if (kv.first < nodes.size() && nodes[kv.first]) {
                BiNode n = *nodes[kv.first];
                n.id = kv.second;
                n.rank = (uint64_t)(kv.second - 1);
                nn[kv.second] = std::move(n);
            }
        }
        nodes = std::move(nn);
        std::unordered_set<BiEdge, BiEdgeHash> ne;
        for (const auto& e : edges) {
            auto f = old_to_new.find(e.from.node_id()), t = old_to_new.find(e.to.node_id());
            if (f != old_to_new.end() && t != old_to_new.end())
                ne.insert(BiEdge(Handle::make(f->second, e.from.is_reverse()), Handle::make(t->second, e.to.is_reverse())));
        }
        edges = std::move(ne);
        remap_path_steps(flat_table(old_to_new), nullptr);
        if (verbose) std::cerr << "\n[apply_ordering] Applied ordering: renumbered " << old_to_new.size() << " nodes\n\n";
    }

    // ---- flat view for the library's host algorithms (gfs_find_head_nodes, gfs_groom_order, gfs_topological_order)
    struct Flat {
        std::vector<uint8_t> present;
        std::vector<uint64_t> edge_from, edge_to, steps, path_first;
    };
    Flat flat() const {
        Flat f;
        f.present.resize(nodes.size());
        for (size_t i = 0; i < nodes.size(); ++i) f.present[i] = nodes[i].has_value();
        f.edge_from.reserve(edges.size()); f.edge_to.reserve(edges.size());
        for (const auto& e : edges) { f.edge_from.push_back(e.from.value); f.edge_to.push_back(e.to.value); }
        f.path_first.push_back(0);
        for (const auto& p : paths) {
            for (Handle h : p.steps) f.steps.push_back(h.value);
            f.path_first.push_back(f.steps.size());
        }
        return f;
    }

    /// graph_ops.rs:1138-1183.
    std::vector<Handle> find_head_nodes() const {
        Flat f = flat();
        std::vector<uint64_t> out(node_count() + 1);
        uint64_t n = 0;
        check(gfs_find_head_nodes(f.present.data(), f.present.size(), f.edge_from.data(), f.edge_to.data(), f.edge_from.size(),
                                  f.steps.data(), f.path_first.data(), paths.size(), out.data(), &n));
        std::vector<Handle> h(n);
        for (uint64_t i = 0; i < n; ++i) h[i] = Handle::from_u64(out[i]);
        return h;
    }
    /// groom.rs:49-275 in BFS mode (the only mode the pipeline uses): all nodes in increasing id, as a
    /// reverse handle where the node must be flipped.
    std::vector<Handle> groom(bool use_bfs = true, bool verbose = false) const {
        if (!use_bfs) throw Error(GFS_ERR_INVALID, "groom: only the BFS mode of the Ygs pipeline is provided");
        Flat f = flat();
        std::vector<uint64_t> out(node_count());
        uint64_t flipped = 0;
        check(gfs_groom_order(f.present.data(), f.present.size(), f.edge_from.data(), f.edge_to.data(), f.edge_from.size(),
                              f.steps.data(), f.path_first.data(), paths.size(), out.data(), &flipped));
        if (verbose) std::cerr << "[groom] Flipped " << flipped << " nodes\n";
        std::vector<Handle> h(out.size());
        for (size_t i = 0; i < out.size(); ++i) h[i] = Handle::from_u64(out[i]);
        return h;
    }
    /// groom.rs:533-605.
    void apply_grooming_with_reorder(const std::vector<Handle>& groomed, bool reorder, bool verbose = false) {
        std::unordered_set<size_t> flips;
        for (Handle h : groomed) if (h.is_reverse()) flips.insert(h.node_id());
        if (verbose && !flips.empty()) std::cerr << "[apply_grooming] Flipping " << flips.size() << " nodes\n";
        for (size_t id : flips) if (id < nodes.size() && nodes[id]) nodes[id]->sequence = reverse_complement(nodes[id]->sequence);
        std::unordered_set<BiEdge, BiEdgeHash> ne;
        for (const auto& e : edges)
            ne.insert(BiEdge(flips.count(e.from.node_id()) ? e.from.flip() : e.from, flips.count(e.to.node_id()) ? e.to.flip() : e.to));
        edges = std::move(ne);
        {
            size_t top = 0;
            for (size_t id : flips) top = std::max(top, id + 1);
            std::vector<uint8_t> flip(top, 0);
            for (size_t id : flips) flip[id] = 1;
            remap_path_steps({}, &flip);
        }
        if (reorder) {
            std::unordered_map<size_t, size_t> m;
            for (size_t i = 0; i < groomed.size(); ++i) m[groomed[i].node_id()] = i + 1;
            apply_node_id_mapping(m);
        }
    }
    /// graph_ops.rs:1232-1485 with (use_heads, use_tails) = (true, false), the pipeline's `s`.
    std::vector<Handle> exact_odgi_topological_order(bool use_heads = true, bool use_tails = false, bool /*verbose*/ = false) const {
        if (!use_heads || use_tails) throw Error(GFS_ERR_INVALID, "exact_odgi_topological_order: only (use_heads, !use_tails) is provided");
        Flat f = flat();
        std::vector<uint64_t> out(node_count());
        uint64_t n = 0;
        check(gfs_topological_order(f.present.data(), f.present.size(), f.edge_from.data(), f.edge_to.data(), f.edge_from.size(),
                                    f.steps.data(), f.path_first.data(), paths.size(), out.data(), &n));
        std::vector<Handle> h(n);
        for (uint64_t i = 0; i < n; ++i) h[i] = Handle::from_u64(out[i]);
        return h;
    }

    /// graph_ops.rs:693-738 (S lines by id, L lines, P lines); L-line order is the set's iteration order,
    /// as in the reference (compare GFAs as sets of lines).
    void write_gfa(std::ostream& w) const {
        w << "H\tVN:Z:1.0\n";
        for (size_t id = 0; id < nodes.size(); ++id)
            if (const auto& n = nodes[id]) w << "S\t" << id << "\t" << std::string(n->sequence.begin(), n->sequence.end()) << "\n";
        for (const auto& e : edges)
            w << "L\t" << e.from.node_id() << "\t" << e.from.orientation_char() << "\t" << e.to.node_id() << "\t" << e.to.orientation_char() << "\t0M\n";
        for (const auto& p : paths) {
            w << "P\t" << p.name << "\t";
            for (size_t i = 0; i < p.steps.size(); ++i) w << (i ? "," : "") << p.steps[i].to_string();
            w << "\t*\n";
        }
    }
};

// ------------------------------------------------------------------------------------------------
// gfa_parser.rs
// ------------------------------------------------------------------------------------------------
namespace gfa_parser {
/// gfa_parser.rs:9-170: segment names -> ids 1.. in order of first appearance; links naming an unknown
/// segment are an error; path steps on unknown segments are skipped.
inline BidirectedGraph load_gfa(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Failed to open file: " + path);
    BidirectedGraph g;
    std::unordered_map<std::string, size_t> id_of;
    size_t next_id = 1;
    std::vector<std::vector<std::string>> links;
    std::vector<std::pair<std::string, std::string>> pending_paths;
    auto split = [](const std::string& s, char c) {
        std::vector<std::string> out;
        size_t a = 0;
        for (;;) {
            size_t b = s.find(c, a);
            out.push_back(s.substr(a, b == std::string::npos ? b : b - a));
            if (b == std::string::npos) break;
            a = b + 1;
        }
        return out;
    };
    auto trim = [](std::string s) {
        const char* ws = " \t\r\n";
        const size_t a = s.find_first_not_of(ws);
        if (a == std::string::npos) return std::string();
        return s.substr(a, s.find_last_not_of(ws) - a + 1);
    };
    std::string line;
    while (std::getline(in, line)) {
        line = trim(line);
        if (line.empty() || line[0] == 'H') continue;
        auto f = split(line, '\t');
        if (f[0] == "S") {
            if (f.size() < 3) continue;
            auto it = id_of.find(f[1]);
            if (it == id_of.end()) it = id_of.emplace(f[1], next_id++).first;
            g.add_node(it->second, f[2]);
        } else if (f[0] == "L") {
            if (f.size() >= 5) links.push_back({f[1], f[2], f[3], f[4]});
        } else if (f[0] == "P") {
            if (f.size() >= 3) pending_paths.emplace_back(f[1], f[2]);
        }
    }
    for (auto& l : links) {
        auto a = id_of.find(l[0]), b = id_of.find(l[2]);
        if (a == id_of.end()) throw std::runtime_error("Unknown node in link: " + l[0]);
        if (b == id_of.end()) throw std::runtime_error("Unknown node in link: " + l[2]);
        g.add_edge(Handle::make(a->second, l[1] != "+"), Handle::make(b->second, l[3] != "+"));
    }
    for (auto& pp : pending_paths) {
        BiPath p(pp.first);
        for (auto& raw : split(pp.second, ',')) {
            std::string s = trim(raw);
            if (s.empty()) continue;
            const char o = s.back();
            if (o != '+' && o != '-') continue;
            auto it = id_of.find(s.substr(0, s.size() - 1));
            if (it != id_of.end()) p.add_step(Handle::make(it->second, o == '-'));
        }
        if (!p.steps.empty()) g.paths.push_back(std::move(p));           // gfa_parser.rs:128-130
    }
    return g;
}
/// gfa_parser.rs:136-185: like BidirectedGraph::write_gfa but with the L lines sorted and "0M" overlaps on P lines.
inline void write_gfa(const BidirectedGraph& g, const std::string& path) {
    std::ofstream out(path);
    if (!out) throw std::runtime_error("Failed to create file: " + path);
    out << "H\tVN:Z:1.0\n";
    for (size_t id = 0; id < g.nodes.size(); ++id)
        if (g.nodes[id]) out << "S\t" << id << "\t" << std::string(g.nodes[id]->sequence.begin(), g.nodes[id]->sequence.end()) << "\n";
    std::vector<BiEdge> es(g.edges.begin(), g.edges.end());
    std::sort(es.begin(), es.end(), [](const BiEdge& a, const BiEdge& b) { return a.from != b.from ? a.from < b.from : a.to < b.to; });
    for (const auto& e : es)
        out << "L\t" << e.from.node_id() << "\t" << e.from.orientation_char() << "\t" << e.to.node_id() << "\t" << e.to.orientation_char() << "\t0M\n";
    for (const auto& p : g.paths) {
        out << "P\t" << p.name << "\t";
        for (size_t i = 0; i < p.steps.size(); ++i) out << (i ? "," : "") << p.steps[i].to_string();
        out << "\t";
        for (size_t i = 0; i + 1 < p.steps.size(); ++i) out << (i ? "," : "") << "0M";
        out << "\n";
    }
}
}  // namespace gfa_parser

// ------------------------------------------------------------------------------------------------
// sgd.rs — PathIndex
// ------------------------------------------------------------------------------------------------
/// sgd.rs:14-107.  The per-step offsets are computed and kept on the GPU (one 16-byte record per step);
/// step_to_path / step_to_rank / first_step / step_count are functions of the first-step table and live
/// on the host; accessors that need offsets export them once, lazily.
class PathIndex {
public:
    PathIndex() = default;
    PathIndex(const PathIndex&) = delete;
    PathIndex& operator=(const PathIndex&) = delete;
    PathIndex(PathIndex&& o) noexcept { *this = std::move(o); }
    PathIndex& operator=(PathIndex&& o) noexcept {
        if (this != &o) {
            reset();
            ix_ = o.ix_; o.ix_ = nullptr;
            first_ = std::move(o.first_); handle_of_step_ = std::move(o.handle_of_step_);
            pos_ = std::move(o.pos_); len_ = std::move(o.len_); have_pos_ = o.have_pos_; have_len_ = o.have_len_;
        }
        return *this;
    }
    ~PathIndex() { reset(); }

    static PathIndex from_graph(const BidirectedGraph& graph) {
        return from_dense(graph, graph.dense());
    }
    static PathIndex from_dense(const BidirectedGraph& graph, const BidirectedGraph::Dense& d) {
        PathIndex ix;
        ix.first_ = d.path_first_step;
        for (const auto& p : graph.paths) for (Handle h : p.steps) ix.handle_of_step_.push_back(h);
        check(gfs_index_build(d.step_handles.data(), d.path_first_step.data(), d.node_len.data(), d.step_handles.size(),
                              graph.paths.size(), d.node_len.size(), &ix.ix_));
        return ix;
    }

    size_t get_total_steps() const { return (size_t)first_.back(); }
    Handle get_handle_of_step(size_t step_idx) const { return handle_of_step_.at(step_idx); }
    size_t get_position_of_step(size_t step_idx) const { export_pos(); return (size_t)pos_.at(step_idx); }
    size_t get_path_of_step(size_t step_idx) const {
        return (size_t)(std::upper_bound(first_.begin(), first_.end(), (uint64_t)step_idx) - first_.begin()) - 1;
    }
    size_t get_rank_of_step(size_t step_idx) const { return step_idx - (size_t)first_[get_path_of_step(step_idx)]; }
    size_t get_path_step_count(size_t path_idx) const { return (size_t)(first_.at(path_idx + 1) - first_.at(path_idx)); }
    size_t get_step_at_path_position(size_t path_idx, size_t rank) const { return (size_t)first_.at(path_idx) + rank; }
    size_t num_paths() const { return first_.size() - 1; }
    size_t get_path_length(size_t path_idx) const { export_len(); return (size_t)len_.at(path_idx); }

    const gfs_index* handle() const { return ix_; }

private:
    void reset() { if (ix_) gfs_index_free(ix_); ix_ = nullptr; }
    void export_pos() const {
        if (have_pos_) return;
        pos_.resize(get_total_steps());
        check(gfs_index_export(ix_, pos_.data(), nullptr));
        have_pos_ = true;
    }
    void export_len() const {
        if (have_len_) return;
        len_.resize(num_paths());
        check(gfs_index_export(ix_, nullptr, len_.data()));
        have_len_ = true;
    }
    gfs_index* ix_ = nullptr;
    std::vector<uint64_t> first_{0};
    std::vector<Handle> handle_of_step_;
    mutable std::vector<uint64_t> pos_, len_;
    mutable bool have_pos_ = false, have_len_ = false;
};

// ------------------------------------------------------------------------------------------------
// sgd.rs — parameters
// ------------------------------------------------------------------------------------------------
/// sgd.rs:196-234 (field for field; defaults = the reference's `Default`).
struct PathSGDParams {
    uint64_t iter_max = 100;
    uint64_t iter_with_max_learning_rate = 0;
    uint64_t min_term_updates = 100;
    double delta = 0.0;
    double eps = 0.01;
    double eta_max = 100.0;
    double theta = 0.99;
    uint64_t space = 100;
    uint64_t space_max = 100;
    uint64_t space_quantization_step = 100;
    double cooling_start = 0.5;
    size_t nthreads = 1;
    bool progress = false;
    uint64_t seed = 9399220;

    gfs_sgd_params c() const {
        return gfs_sgd_params{iter_max, iter_with_max_learning_rate, min_term_updates, delta, eps, eta_max, theta, space,
                              space_max, space_quantization_step, cooling_start, (uint64_t)nthreads, progress ? 1u : 0u, seed};
    }
};

/// sgd.rs:676-763.
struct LayoutSGDParams {
    size_t dimensions = 2;
    uint64_t iter_max = 30;
    uint64_t iter_with_max_learning_rate = 0;
    uint64_t min_term_updates = 100;
    double delta = 0.0;
    double eps = 0.01;
    double eta_max = 100.0;
    double theta = 0.99;
    uint64_t space = 100;
    uint64_t space_max = 1000;
    uint64_t space_quantization_step = 100;
    double cooling_start = 0.5;
    size_t nthreads = 1;
    bool progress = false;
    uint64_t seed = 9399220;

    static LayoutSGDParams from_graph(const BidirectedGraph& graph, size_t dimensions, size_t nthreads) {
        // sgd.rs:733-763 builds a PathIndex only for the step counts; those need no device
        uint64_t sum = 0, mx = 0;
        for (const auto& p : graph.paths) { sum += p.steps.size(); mx = std::max<uint64_t>(mx, p.steps.size()); }
        LayoutSGDParams q;
        q.dimensions = dimensions;
        q.min_term_updates = 10 * sum;
        q.eta_max = (double)(mx * mx);
        q.space = mx;
        q.nthreads = nthreads;
        return q;
    }
    gfs_sgd_params c() const {
        return gfs_sgd_params{iter_max, iter_with_max_learning_rate, min_term_updates, delta, eps, eta_max, theta, space,
                              space_max, space_quantization_step, cooling_start, (uint64_t)nthreads, progress ? 1u : 0u, seed};
    }
};

// ------------------------------------------------------------------------------------------------
// layout.rs
// ------------------------------------------------------------------------------------------------
/// layout.rs:17-245: coords[node * 2 * dimensions + end * dimensions + dim], end 0 = '+', 1 = '-'.
struct Layout {
    size_t dimensions = 0, num_nodes = 0;
    std::vector<double> coords;

    static Layout make(size_t dimensions, size_t num_nodes) {            // Layout::new, layout.rs:28-35
        Layout l;
        l.dimensions = dimensions; l.num_nodes = num_nodes;
        l.coords.assign(num_nodes * 2 * dimensions, 0.0);
        return l;
    }
    static Layout from_vectors(const std::vector<std::vector<double>>& v) {   // layout.rs:39-69: [dim][2*node+end]
        if (v.empty()) throw std::invalid_argument("Must have at least 1 dimension");
        const size_t entries = v[0].size();
        if (entries % 2) throw std::invalid_argument("Must have even number of entries (2 per node)");
        for (auto& x : v) if (x.size() != entries) throw std::invalid_argument("All dimension vectors must have same length");
        Layout l = make(v.size(), entries / 2);
        for (size_t node = 0; node < l.num_nodes; ++node)
            for (size_t end = 0; end < 2; ++end)
                for (size_t dim = 0; dim < l.dimensions; ++dim) l.coords[l.index(node, end, dim)] = v[dim][node * 2 + end];
        return l;
    }
    size_t index(size_t node, size_t end, size_t dim) const { return node * 2 * dimensions + end * dimensions + dim; }
    double get(size_t node, size_t end, size_t dim) const { return coords.at(index(node, end, dim)); }
    void set(size_t node, size_t end, size_t dim, double value) { coords.at(index(node, end, dim)) = value; }
    const double* get_coords(size_t node, size_t end) const { return coords.data() + index(node, end, 0); }
    double x_plus(size_t node) const { return get(node, 0, 0); }
    double y_plus(size_t node) const { return dimensions > 1 ? get(node, 0, 1) : 0.0; }
    double x_minus(size_t node) const { return get(node, 1, 0); }
    double y_minus(size_t node) const { return dimensions > 1 ? get(node, 1, 1) : 0.0; }
    double distance(size_t node_a, size_t end_a, size_t node_b, size_t end_b) const {   // layout.rs:126-133
        double s = 0.0;
        for (size_t d = 0; d < dimensions; ++d) { const double x = get(node_a, end_a, d) - get(node_b, end_b, d); s += x * x; }
        return std::sqrt(s);
    }
    /// layout.rs:138-163, byte for byte (Rust `{}` float formatting), through the library's buffered writer.
    void write_tsv(const std::string& path) const {
        uint64_t bytes = 0;
        check(gfs_layout_write_tsv(coords.data(), num_nodes, (uint32_t)dimensions, path.c_str(), &bytes));
    }
    /// layout.rs:166-218.
    static Layout read_tsv(std::istream& in) {
        std::string header;
        if (!std::getline(in, header)) throw std::runtime_error("Empty file");
        const size_t cols = (size_t)std::count(header.begin(), header.end(), '\t') + 1;
        if (cols < 3 || (cols - 1) % 2) throw std::runtime_error("Invalid header format");
        const size_t dims = (cols - 1) / 2;
        std::vector<std::vector<double>> rows;
        std::string line;
        while (std::getline(in, line)) {
            if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
            std::stringstream ss(line);
            std::string f;
            std::vector<double> r;
            size_t k = 0;
            while (std::getline(ss, f, '\t')) { if (k++) r.push_back(std::stod(f)); }
            if (k != cols) throw std::runtime_error("Row has " + std::to_string(k) + " columns, expected " + std::to_string(cols));
            rows.push_back(std::move(r));
        }
        Layout l = make(dims, rows.size());
        for (size_t n = 0; n < rows.size(); ++n)
            for (size_t d = 0; d < dims; ++d) { l.set(n, 0, d, rows[n][d]); l.set(n, 1, d, rows[n][dims + d]); }
        return l;
    }
    /// layout.rs:224-245.
    double calculate_stress(const std::vector<std::tuple<size_t, size_t, size_t, size_t, double>>& targets) const {
        double ws = 0.0, wt = 0.0;
        for (auto& t : targets) {
            const double target = std::get<4>(t);
            if (target == 0.0) continue;
            const double w = 1.0 / (target * target);
            const double err = distance(std::get<0>(t), std::get<1>(t), std::get<2>(t), std::get<3>(t)) - target;
            ws += err * err * w; wt += w;
        }
        return wt > 0.0 ? std::sqrt(ws / wt) : 0.0;
    }
};

// ------------------------------------------------------------------------------------------------
// sgd.rs — the hot path
// ------------------------------------------------------------------------------------------------
/// Statistics of the last SGD call on this thread (replaces the reference's stderr progress lines).
inline gfs_stats& last_stats() { static thread_local gfs_stats st{}; return st; }

/// sgd.rs:237-614.  Key = dense idx in node_order; empty when the graph has no nodes or no path has more
/// than one step (:242-244, :258-261).
inline std::unordered_map<size_t, double> path_linear_sgd(const BidirectedGraph& graph, const PathSGDParams& params) {
    std::unordered_map<size_t, double> positions;
    if (graph.node_count() == 0) return positions;
    BidirectedGraph::Dense d = graph.dense();
    PathIndex ix = PathIndex::from_dense(graph, d);
    gfs_sgd_params cp = params.c();
    std::vector<double> x = d.x_init;
    const int rc = gfs_sgd_1d(ix.handle(), &cp, x.data(), &last_stats());
    if (rc == GFS_ERR_NO_VALID_PATH) { std::cerr << "[path_sgd] No paths with multiple steps found\n"; return positions; }
    check(rc);
    if (params.progress) std::cerr << "[path_sgd] Complete: " << last_stats().applied_updates << " term updates\n";
    positions.reserve(x.size());
    for (size_t i = 0; i < x.size(); ++i) positions.emplace(i, x[i]);
    return positions;
}

/// sgd.rs:641-672: forward handles of all nodes ordered by final position.  SGD and sort both run on the
/// device (gfs_sgd_sort_1d); ties by dense idx.
inline std::vector<Handle> path_sgd_sort(const BidirectedGraph& graph, const PathSGDParams& params) {
    std::vector<Handle> out;
    if (graph.node_count() == 0) return out;
    BidirectedGraph::Dense d = graph.dense();
    PathIndex ix = PathIndex::from_dense(graph, d);
    gfs_sgd_params cp = params.c();
    std::vector<double> x = d.x_init;
    std::vector<uint32_t> order(x.size());
    const int rc = gfs_sgd_sort_1d(ix.handle(), &cp, x.data(), order.data(), &last_stats());
    if (rc == GFS_ERR_NO_VALID_PATH) { std::cerr << "[path_sgd] No paths with multiple steps found\n"; return out; }
    check(rc);
    out.reserve(order.size());
    for (uint32_t idx : order) out.push_back(Handle::forward(d.node_id_of_idx.at(idx)));
    return out;
}

namespace detail {
// xoshiro256+ seeded through SplitMix64 (rand_xoshiro's seed_from_u64), N(0,1) by Marsaglia's polar method.
struct InitRng {
    uint64_t s[4];
    bool has_spare = false; double spare = 0.0;
    explicit InitRng(uint64_t seed) {
        uint64_t x = seed;
        for (auto& w : s) { x += 0x9e3779b97f4a7c15ull; uint64_t z = x; z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; w = z ^ (z >> 31); }
    }
    uint64_t next_u64() {
        const uint64_t r = s[0] + s[3], t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = (s[3] << 45) | (s[3] >> 19);
        return r;
    }
    double unit() { return (double)(next_u64() >> 11) * (1.0 / 9007199254740992.0); }
    double normal() {
        if (has_spare) { has_spare = false; return spare; }
        double u, v, q;
        do { u = 2.0 * unit() - 1.0; v = 2.0 * unit() - 1.0; q = u * u + v * v; } while (q >= 1.0 || q == 0.0);
        const double f = std::sqrt(-2.0 * std::log(q) / q);
        spare = v * f; has_spare = true;
        return u * f;
    }
};
}  // namespace detail

/// The coordinate init of sgd.rs:816-854 in Layout order: dim 0 = cumulative length (+ end) and
/// + node length (- end); dims >= 1 = N(0,1) * sqrt(2N), drawn node-major, + end's dims then - end's.
inline std::vector<double> initial_layout(const BidirectedGraph::Dense& d, size_t num_nodes, size_t dims, uint64_t seed) {
    const size_t n = d.node_len.size();
    std::vector<double> c(n * 2 * dims, 0.0);
    detail::InitRng rng(seed);
    const double sqrt_n = std::sqrt((double)num_nodes * 2.0);
    for (size_t i = 0; i < n; ++i) {
        c[i * 2 * dims] = d.x_init[i];
        for (size_t k = 1; k < dims; ++k) c[i * 2 * dims + k] = rng.normal() * sqrt_n;
        c[i * 2 * dims + dims] = d.x_init[i] + (double)d.node_len[i];
        for (size_t k = 1; k < dims; ++k) c[i * 2 * dims + dims + k] = rng.normal() * sqrt_n;
    }
    return c;
}

/// sgd.rs:773-1188.
inline Layout path_linear_sgd_layout(const BidirectedGraph& graph, const LayoutSGDParams& params) {
    const size_t num_nodes = graph.node_count();
    if (num_nodes == 0) return Layout::make(params.dimensions, 0);
    BidirectedGraph::Dense d = graph.dense();
    PathIndex ix = PathIndex::from_dense(graph, d);
    gfs_sgd_params cp = params.c();
    Layout l;
    l.dimensions = params.dimensions; l.num_nodes = d.node_len.size();
    l.coords = initial_layout(d, num_nodes, params.dimensions, params.seed);
    const int rc = gfs_sgd_nd(ix.handle(), &cp, (uint32_t)params.dimensions, l.coords.data(), &last_stats());
    if (rc == GFS_ERR_NO_VALID_PATH) {
        std::cerr << "[path_sgd_layout] No paths with multiple steps found\n";
        return Layout::make(params.dimensions, num_nodes);
    }
    check(rc);
    if (params.progress) std::cerr << "[path_sgd_layout] Complete\n";
    return l;
}

/// sgd.rs:1196-1283: sqrt(mean((d_layout - d_path)^2 / d_path^2)) over `sample_count` seeded draws
/// (uniform step, uniform partner on the same path, + ends).
inline double calculate_layout_stress(const BidirectedGraph& graph, const Layout& layout, size_t sample_count) {
    if (graph.node_count() == 0 || layout.num_nodes == 0) return 0.0;
    PathIndex ix = PathIndex::from_graph(graph);
    double rms = 0.0, mean_abs = 0.0;
    uint64_t counted = 0;
    check(gfs_stress(ix.handle(), (uint32_t)layout.dimensions, 1, layout.coords.data(), sample_count, 12345, &rms, &mean_abs, &counted));
    return rms;
}

// ------------------------------------------------------------------------------------------------
// ygs.rs
// ------------------------------------------------------------------------------------------------
/// ygs.rs:16-92.
struct YgsParams {
    PathSGDParams path_sgd;
    uint8_t verbose = 0;

    static YgsParams make_default() {                                     // Default::default, ygs.rs:23-46
        YgsParams p;
        p.path_sgd.min_term_updates = 0;
        p.path_sgd.eta_max = 0.0;
        p.path_sgd.space = 0;
        return p;
    }
    /// ygs.rs:50-92: min_term_updates = sum of path step counts, eta_max = (max step count)^2, space = longest
    /// path in bp — which is what needs the path index, built on the GPU.
    static YgsParams from_graph(const BidirectedGraph& graph, uint8_t verbose, size_t nthreads) {
        YgsParams p = make_default();
        p.verbose = verbose;
        p.path_sgd.nthreads = nthreads;
        p.path_sgd.progress = verbose >= 2;
        PathIndex ix = PathIndex::from_graph(graph);
        uint64_t sum = 0; size_t mx = 0, max_len = 0;
        for (size_t i = 0; i < ix.num_paths(); ++i) {
            sum += ix.get_path_step_count(i);
            mx = std::max(mx, ix.get_path_step_count(i));
            max_len = std::max(max_len, ix.get_path_length(i));
        }
        p.path_sgd.min_term_updates = sum;
        p.path_sgd.eta_max = (double)(mx * mx);
        p.path_sgd.space = max_len;
        if (verbose >= 2) {
            std::cerr << "[ygs_sort] Calculated parameters:\n  sum_path_step_count: " << sum << "\n  max_path_step_count: " << mx
                      << "\n  max_path_length: " << max_len << "\n  min_term_updates: " << p.path_sgd.min_term_updates
                      << "\n  eta_max: " << p.path_sgd.eta_max << "\n  space: " << p.path_sgd.space << "\n";
        }
        return p;
    }
};

/// ygs.rs:195-206.
inline void sgd_sort_only(BidirectedGraph& graph, const PathSGDParams& params, uint8_t verbose) {
    if (verbose >= 2) std::cerr << "[path_sgd] Starting path-guided SGD\n";
    graph.apply_ordering(path_sgd_sort(graph, params), false);
    if (verbose >= 2) std::cerr << "[path_sgd] Complete\n";
}
/// ygs.rs:180-192.
inline void groom_only(BidirectedGraph& graph, uint8_t verbose) {
    if (verbose >= 2) std::cerr << "[groom] Starting grooming\n";
    graph.apply_grooming_with_reorder(graph.groom(true, verbose >= 2), true, verbose >= 2);
    if (verbose >= 2) std::cerr << "[groom] Complete\n";
}
/// ygs.rs:147-159.
inline void topological_sort_only(BidirectedGraph& graph, uint8_t verbose) {
    if (verbose >= 2) std::cerr << "[topological_sort] Starting topological sort (heads only)\n";
    graph.apply_ordering(graph.exact_odgi_topological_order(true, false, verbose >= 2), false);
    if (verbose >= 2) std::cerr << "[topological_sort] Complete\n";
}
/// ygs.rs:97-143: Y (GPU) -> g -> s, renumbering after every step.
inline void ygs_sort(BidirectedGraph& graph, const YgsParams& params) {
    if (params.verbose >= 1) std::cerr << "[ygs_sort] Starting Ygs pipeline (Y=SGD, g=groom, s=topological_sort)\n";
    graph.apply_ordering(path_sgd_sort(graph, params.path_sgd), false);
    graph.apply_grooming_with_reorder(graph.groom(true, params.verbose >= 2), true, params.verbose >= 2);
    graph.apply_ordering(graph.exact_odgi_topological_order(true, false, params.verbose >= 2), false);
    if (params.verbose >= 1) std::cerr << "[ygs_sort] Ygs pipeline complete\n";
}

}  // namespace gfasort
#endif  // GFASORT_HPP
