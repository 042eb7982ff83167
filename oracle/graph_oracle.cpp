// graph_oracle.cpp — CPU ORACLE for the two host steps that consume `Y`'s order: `g` (groom) and `s`
// (heads-first topological sort).  TEST INFRASTRUCTURE, like gfs_oracle.cpp: only tests/ load it.
//
// A deliberately literal restatement of the reference, including its cost: find_head_nodes scans every
// edge per node (src/graph_ops.rs:1138-1183), groom_bfs_majority scans every edge per dequeued handle
// (src/groom.rs:202-275), exact_odgi_topological_order clones and sorts the whole edge set per processed
// handle (src/graph_ops.rs:1232-1485).  The product's linear-time versions
// (gfasort_b200/csrc/gfs_host_graph.cpp) must emit exactly these orders.
// Containers: HashSet -> std::set / std::unordered_set (only membership and deterministic min are used;
// wherever the reference iterates a HashSet it sorts first, or the iteration order cannot matter).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <set>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <vector>

namespace {
using Handle = uint64_t;                       // id << 1 | is_reverse (src/graph.rs:9-19)
inline uint64_t node_id(Handle h) { return h >> 1; }
inline bool is_rev(Handle h) { return h & 1; }
inline Handle flip(Handle h) { return h ^ 1; }
struct Edge { Handle from, to; bool operator<(const Edge& o) const { return from != o.from ? from < o.from : to < o.to; }
              bool operator==(const Edge& o) const { return from == o.from && to == o.to; } };
struct G {
    const uint8_t* present; uint64_t nodes_len;
    std::vector<Edge> edges;                   // the HashSet<BiEdge>, in the caller's order
    const uint64_t* steps; const uint64_t* path_first; uint64_t P;
};

// graph_ops.rs:1111-1125
std::unordered_map<uint64_t, uint64_t> build_path_position_map(const G& g) {
    std::unordered_map<uint64_t, uint64_t> m;
    for (uint64_t p = 0; p < g.P; ++p)
        for (uint64_t s = g.path_first[p]; s < g.path_first[p + 1]; ++s) {
            uint64_t id = node_id(g.steps[s]), pos = s - g.path_first[p];
            auto it = m.find(id);
            if (it == m.end()) m[id] = pos; else it->second = std::min(it->second, pos);
        }
    return m;
}
// graph_ops.rs:1138-1183
std::vector<Handle> find_head_nodes(const G& g) {
    std::vector<Handle> heads;
    for (uint64_t id = 0; id < g.nodes_len; ++id) {
        if (!g.present[id]) continue;
        Handle fwd = id << 1, rev = fwd | 1;
        bool has_left_incoming = false;
        for (const Edge& e : g.edges) {
            if (e.to == fwd) { has_left_incoming = true; break; }
            if (e.from == rev) { has_left_incoming = true; break; }
        }
        if (!has_left_incoming) heads.push_back(fwd);
    }
    auto pos = build_path_position_map(g);
    auto key = [&](Handle h) { auto it = pos.find(node_id(h)); return std::make_pair(it == pos.end() ? ~0ull : it->second, node_id(h)); };
    std::stable_sort(heads.begin(), heads.end(), [&](Handle a, Handle b) { return key(a) < key(b); });
    return heads;
}
// groom.rs:202-275
void groom_bfs_majority(const G& g, const std::vector<Handle>& seeds, std::unordered_set<uint64_t>& visited,
                        std::unordered_set<uint64_t>& flipped) {
    std::deque<Handle> queue;
    for (Handle seed : seeds)
        if (!visited.count(node_id(seed))) {
            queue.push_back(seed); visited.insert(node_id(seed));
            if (is_rev(seed)) flipped.insert(node_id(seed));
        }
    while (!queue.empty()) {
        Handle current = queue.front(); queue.pop_front();
        std::vector<Handle> next_handles;
        for (const Edge& e : g.edges) {
            if (e.from == current) next_handles.push_back(e.to);
            else if (flip(e.to) == current) next_handles.push_back(flip(e.from));
        }
        std::stable_sort(next_handles.begin(), next_handles.end(), [](Handle a, Handle b) {
            return std::make_pair(node_id(a), is_rev(a)) < std::make_pair(node_id(b), is_rev(b)); });
        for (Handle next : next_handles)
            if (!visited.count(node_id(next))) {
                visited.insert(node_id(next));
                if (is_rev(next)) flipped.insert(node_id(next));
                queue.push_back(next);
            }
    }
}
}  // namespace

extern "C" {

uint64_t oracle_find_head_nodes(const uint8_t* present, uint64_t nodes_len, const uint64_t* ef, const uint64_t* et, uint64_t E,
                                const uint64_t* steps, const uint64_t* path_first, uint64_t P, uint64_t* out) {
    G g{present, nodes_len, {}, steps, path_first, P};
    for (uint64_t e = 0; e < E; ++e) g.edges.push_back({ef[e], et[e]});
    auto h = find_head_nodes(g);
    std::memcpy(out, h.data(), h.size() * 8);
    return h.size();
}

// groom(use_bfs = true, use_coverage_dfs = false) — groom.rs:49-199.  Returns the number of flipped nodes.
uint64_t oracle_groom(const uint8_t* present, uint64_t nodes_len, const uint64_t* ef, const uint64_t* et, uint64_t E,
                      const uint64_t* steps, const uint64_t* path_first, uint64_t P, uint64_t* order_out) {
    G g{present, nodes_len, {}, steps, path_first, P};
    for (uint64_t e = 0; e < E; ++e) g.edges.push_back({ef[e], et[e]});
    std::vector<Handle> seeds = find_head_nodes(g);
    std::unordered_set<uint64_t> visited, flipped;
    std::vector<Handle> current_seeds;
    if (seeds.empty()) {
        for (uint64_t id = 0; id < nodes_len; ++id) if (present[id]) { current_seeds.push_back(id << 1); break; }
    } else current_seeds = seeds;
    while (visited.size() < nodes_len) {                                  // :135 (nodes.len() counts the None slots)
        if (current_seeds.empty()) {
            for (uint64_t id = 0; id < nodes_len; ++id) {
                if (!present[id]) continue;
                if (!visited.count(id)) { current_seeds.push_back(id << 1); break; }
            }
            if (current_seeds.empty()) break;
        }
        groom_bfs_majority(g, current_seeds, visited, flipped);
        current_seeds.clear();
    }
    uint64_t k = 0;
    for (uint64_t id = 0; id < nodes_len; ++id)                           // :171-186 (sorted ids)
        if (present[id]) order_out[k++] = (id << 1) | (flipped.count(id) ? 1 : 0);
    return flipped.size();
}

// exact_odgi_topological_order(use_heads = true, use_tails = false) — graph_ops.rs:1232-1485.
uint64_t oracle_topological_order(const uint8_t* present, uint64_t nodes_len, const uint64_t* ef, const uint64_t* et, uint64_t E,
                                  const uint64_t* steps, const uint64_t* path_first, uint64_t P, uint64_t* order_out) {
    G g{present, nodes_len, {}, steps, path_first, P};
    for (uint64_t e = 0; e < E; ++e) g.edges.push_back({ef[e], et[e]});
    std::vector<Handle> sorted;
    if (nodes_len == 0) return 0;
    std::set<Handle> s;                                                   // BTreeSet<Handle>
    std::unordered_set<uint64_t> visited_nodes;
    std::set<Handle> unvisited;
    for (uint64_t id = 0; id < nodes_len; ++id) if (present[id]) { unvisited.insert(id << 1); unvisited.insert((id << 1) | 1); }
    std::vector<Handle> seeds;
    std::set<Edge> masked_edges;
    for (Handle head : find_head_nodes(g)) { s.insert(head); unvisited.erase(head); unvisited.erase(flip(head)); }
    while (!unvisited.empty() || !s.empty()) {
        if (s.empty()) {
            while (!seeds.empty() && s.empty()) {
                std::stable_sort(seeds.begin(), seeds.end(), [](Handle a, Handle b) {
                    return std::make_pair(node_id(a), is_rev(a)) < std::make_pair(node_id(b), is_rev(b)); });
                Handle handle = seeds.front(); seeds.erase(seeds.begin());
                if (unvisited.count(handle)) { s.insert(handle); unvisited.erase(handle); unvisited.erase(flip(handle)); }
            }
            if (s.empty() && !unvisited.empty()) {
                Handle min_handle = *std::min_element(unvisited.begin(), unvisited.end(), [](Handle a, Handle b) {
                    return std::make_pair(node_id(a), is_rev(a)) < std::make_pair(node_id(b), is_rev(b)); });
                s.insert(min_handle); unvisited.erase(min_handle); unvisited.erase(flip(min_handle));
            }
        }
        while (!s.empty()) {
            Handle handle = *s.begin(); s.erase(s.begin());
            Handle forward_handle = node_id(handle) << 1;
            if (visited_nodes.insert(node_id(handle)).second) sorted.push_back(forward_handle);
            std::vector<Edge> edges_vec(g.edges.begin(), g.edges.end());                 // :1365-1366
            std::sort(edges_vec.begin(), edges_vec.end());
            auto edge_goes_to = [](const Edge& e, Handle h) { return e.to == h || e.from == flip(h); };
            auto edge_goes_from = [](const Edge& e, Handle h) { return e.from == h || e.to == flip(h); };
            auto get_next_handle = [](const Edge& e, Handle h) { return e.from == h ? e.to : flip(e.from); };
            for (const Edge& e : edges_vec) {
                if (edge_goes_to(e, forward_handle) && !masked_edges.count(e)) {
                    uint64_t source_node_id = e.to == forward_handle ? node_id(e.from) : node_id(flip(e.to));
                    Handle source_forward = source_node_id << 1;
                    if (!unvisited.count(source_forward) && !unvisited.count(flip(source_forward))) masked_edges.insert(e);
                }
            }
            for (const Edge& e : edges_vec) {
                if (edge_goes_from(e, forward_handle) && !masked_edges.count(e)) {
                    masked_edges.insert(e);
                    Handle next_handle = get_next_handle(e, forward_handle);
                    if (unvisited.count(next_handle)) {
                        bool has_unmasked_incoming = false;
                        for (const Edge& o : edges_vec)
                            if (edge_goes_to(o, next_handle) && !masked_edges.count(o)) { has_unmasked_incoming = true; break; }
                        if (!has_unmasked_incoming) { s.insert(next_handle); unvisited.erase(next_handle); unvisited.erase(flip(next_handle)); }
                        else if (std::find(seeds.begin(), seeds.end(), next_handle) == seeds.end()) seeds.push_back(next_handle);
                    }
                }
            }
        }
    }
    std::memcpy(order_out, sorted.data(), sorted.size() * 8);
    return sorted.size();
}

}  // extern "C"
