// gfs_io.cpp — flat GFA ingest and buffered writers.  Host code, no device.
//
//   gfs_gfa_parse*   one pass over the GFA text straight into the C ABI's flat arrays (node table, unique
//                    edges, concatenated path steps) — the CLI's parse_gfa (src/bin/gfasort.rs:88-167) makes
//                    three passes, splits every line into a Vec<&str>, and goes through
//                    Vec<Option<BiNode>> / HashSet<BiEdge>; PathIndex is then rebuilt from that three times.
//   gfs_layout_write_tsv   Layout::write_tsv (src/layout.rs:138-163) into one buffer, one write per MB —
//                    the reference issues one write! per field on an unbuffered File (5 syscalls per node).
//   gfs_gfa_write    BidirectedGraph::write_gfa (src/graph_ops.rs:693-738), buffered.
// Formats are byte-for-byte the reference's: header `idx\tx+\ty+\tx-\ty-`, Rust `{}` for f64 (shortest
// digits that round-trip, never an exponent, "NaN"/"inf"/"-inf"), `H\tVN:Z:1.0`, `L ... 0M`, `P ... *`.
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <unordered_set>
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

struct gfs_gfa {
    std::string text;                          // the file; sequences and names point into it
    std::vector<uint8_t> present;              // nodes: Vec<Option<BiNode>>
    std::vector<uint64_t> seq_off, seq_len;    // per node id: where its sequence sits in `text`
    std::vector<uint64_t> node_order;          // add_node insertion order (graph_ops.rs:613-623)
    std::vector<uint64_t> edge_from, edge_to;  // unique per {edge, complement} (add_edge, graph_ops.rs:626-638)
    std::vector<uint64_t> steps, path_first;   // concatenated path steps
    std::vector<uint64_t> name_off, name_len;  // per path
};

namespace {

struct PairHash { size_t operator()(const std::pair<uint64_t, uint64_t>& p) const {
    uint64_t x = p.first * 0x9e3779b97f4a7c15ULL ^ (p.second + 0x7f4a7c15ULL + (p.first << 6));
    x ^= x >> 31; x *= 0xbf58476d1ce4e5b9ULL; return (size_t)(x ^ (x >> 29)); } };

// Rust's str::parse::<usize>: optional leading '+', then one or more ASCII digits, nothing else.
bool parse_usize(const char* b, const char* e, uint64_t& out) {
    if (b < e && *b == '+') ++b;
    if (b >= e) return false;
    uint64_t v = 0;
    for (; b < e; ++b) {
        if (*b < '0' || *b > '9') return false;
        const uint64_t nv = v * 10 + (uint64_t)(*b - '0');
        if (nv < v) return false;
        v = nv;
    }
    out = v;
    return true;
}

// next tab-separated field of [p, end); returns false when the line has no more fields
inline bool next_field(const char*& p, const char* end, const char*& fb, const char*& fe, bool& more) {
    if (!more) return false;
    fb = p;
    const char* t = (const char*)memchr(p, '\t', (size_t)(end - p));
    if (t) { fe = t; p = t + 1; } else { fe = end; p = end; more = false; }
    return true;
}

int parse_text(gfs_gfa* g) {
    const char* base = g->text.data();
    const char* end = base + g->text.size();
    std::unordered_set<std::pair<uint64_t, uint64_t>, PairHash> seen;
    g->path_first.push_back(0);
    // the reference makes three passes (S, then L, then P) so that line order does not matter; S and P are
    // order-independent of the other kinds here too, and L only needs the dedupe set — one pass suffices
    for (const char* ls = base; ls < end;) {
        const char* nl = (const char*)memchr(ls, '\n', (size_t)(end - ls));
        const char* le = nl ? nl : end;
        const char* next = nl ? nl + 1 : end;
        if (le > ls && le[-1] == '\r') --le;                      // str::lines() strips "\r\n"
        if (le > ls && (*ls == 'S' || *ls == 'L' || *ls == 'P')) {
            const char kind = *ls;
            const char* p = ls; bool more = true;
            const char *f[5][2]; int nf = 0;
            const int want = kind == 'S' ? 3 : kind == 'L' ? 5 : 3;
            while (nf < want && next_field(p, le, f[nf][0], f[nf][1], more)) ++nf;
            if (nf >= want) {
                if (kind == 'S') {
                    uint64_t id;
                    if (!parse_usize(f[1][0], f[1][1], id)) { gfs::set_error("Failed to parse node ID"); return GFS_ERR_INVALID; }
                    if (id >= g->present.size()) { g->present.resize(id + 1, 0); g->seq_off.resize(id + 1, 0); g->seq_len.resize(id + 1, 0); }
                    if (!g->present[id]) g->node_order.push_back(id);
                    g->present[id] = 1;
                    g->seq_off[id] = (uint64_t)(f[2][0] - base);
                    g->seq_len[id] = (uint64_t)(f[2][1] - f[2][0]);
                } else if (kind == 'L') {
                    uint64_t a, b;
                    if (!parse_usize(f[1][0], f[1][1], a)) { gfs::set_error("Failed to parse from ID"); return GFS_ERR_INVALID; }
                    if (!parse_usize(f[3][0], f[3][1], b)) { gfs::set_error("Failed to parse to ID"); return GFS_ERR_INVALID; }
                    const bool fa = (f[2][1] - f[2][0] == 1) && *f[2][0] == '+';      // anything but "+" is reverse
                    const bool fb = (f[4][1] - f[4][0] == 1) && *f[4][0] == '+';
                    const uint64_t from = (a << 1) | (fa ? 0 : 1), to = (b << 1) | (fb ? 0 : 1);
                    if (!seen.count({from, to}) && !seen.count({to ^ 1, from ^ 1})) {
                        seen.insert({from, to});
                        g->edge_from.push_back(from); g->edge_to.push_back(to);
                    }
                } else {
                    g->name_off.push_back((uint64_t)(f[1][0] - base));
                    g->name_len.push_back((uint64_t)(f[1][1] - f[1][0]));
                    for (const char* s = f[2][0]; s <= f[2][1];) {
                        const char* c = (const char*)memchr(s, ',', (size_t)(f[2][1] - s));
                        const char* se = c ? c : f[2][1];
                        const char* sb = s;
                        while (sb < se && (*sb == ' ' || (*sb >= 9 && *sb <= 13))) ++sb;       // trim()
                        const char* st = se;
                        while (st > sb && (st[-1] == ' ' || (st[-1] >= 9 && st[-1] <= 13))) --st;
                        if (st > sb) {
                            uint64_t id;
                            if (!parse_usize(sb, st - 1, id)) { gfs::set_error("Failed to parse path node ID"); return GFS_ERR_INVALID; }
                            g->steps.push_back((id << 1) | (st[-1] == '+' ? 0 : 1));
                        }
                        if (!c) break;
                        s = c + 1;
                    }
                    g->path_first.push_back(g->steps.size());
                }
            }
        }
        ls = next;
    }
    return GFS_OK;
}

// Rust `{}` for f64: the shortest digit string that round-trips, laid out positionally (never an exponent):
// digits, then zeros up to the decimal point, or "0." and leading zeros.  to_chars(scientific) yields exactly
// those shortest digits and the decimal exponent.
inline char* fmt_f64(char* p, char* end, double v) {
    if (std::isnan(v)) { std::memcpy(p, "NaN", 3); return p + 3; }
    if (std::isinf(v)) { if (v < 0) { std::memcpy(p, "-inf", 4); return p + 4; } std::memcpy(p, "inf", 3); return p + 3; }
    char tmp[40];
    auto r = std::to_chars(tmp, tmp + sizeof tmp, v, std::chars_format::scientific);      // [-]d[.ddd]e[+-]XX
    const char* t = tmp;
    if (*t == '-') { *p++ = '-'; ++t; }
    char digits[24]; int nd = 0;
    for (; t < r.ptr && *t != 'e'; ++t) if (*t != '.') digits[nd++] = *t;
    int exp10 = 0;
    { ++t; const bool neg = *t == '-'; if (*t == '+' || *t == '-') ++t; for (; t < r.ptr; ++t) exp10 = exp10 * 10 + (*t - '0'); if (neg) exp10 = -exp10; }
    while (nd > 1 && digits[nd - 1] == '0') --nd;                       // "1.50e0" cannot occur, but be safe
    if (nd == 1 && digits[0] == '0') { *p++ = '0'; return p; }          // 0 and -0 print "0" / "-0"
    const int point = exp10 + 1;                                        // digits before the decimal point
    (void)end;
    if (point <= 0) {                                                   // 0.000ddd
        *p++ = '0'; *p++ = '.';
        for (int k = 0; k < -point; ++k) *p++ = '0';
        std::memcpy(p, digits, (size_t)nd); p += nd;
    } else if (point >= nd) {                                           // ddd000
        std::memcpy(p, digits, (size_t)nd); p += nd;
        for (int k = nd; k < point; ++k) *p++ = '0';
    } else {                                                            // dd.ddd
        std::memcpy(p, digits, (size_t)point); p += point;
        *p++ = '.';
        std::memcpy(p, digits + point, (size_t)(nd - point)); p += nd - point;
    }
    return p;
}
inline char* fmt_u64(char* p, char* end, uint64_t v) { return std::to_chars(p, end, v).ptr; }

struct BufWriter {
    FILE* f; std::vector<char> buf; size_t n = 0; uint64_t total = 0; bool ok = true;
    explicit BufWriter(FILE* f_) : f(f_), buf(1 << 20) {}
    char* room(size_t need) { if (n + need > buf.size()) flush(); if (need > buf.size()) buf.resize(need); return buf.data() + n; }
    void advance(char* p) { n = (size_t)(p - buf.data()); }
    void put(const char* s, size_t len) { char* p = room(len); std::memcpy(p, s, len); n += len; }
    void flush() { if (n) { if (fwrite(buf.data(), 1, n, f) != n) ok = false; total += n; n = 0; } }
};

}  // namespace

extern "C" int gfs_gfa_parse_text(const char* text, uint64_t len, gfs_gfa** out) try {
    if (!out || (len && !text)) { gfs::set_error("gfs_gfa_parse_text: null argument"); return GFS_ERR_INVALID; }
    *out = nullptr;
    std::unique_ptr<gfs_gfa> g(new gfs_gfa());
    g->text.assign(text, len);
    int rc = parse_text(g.get());
    if (rc) return rc;
    *out = g.release();
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_gfa_parse_text"); }

extern "C" int gfs_gfa_parse_file(const char* path, gfs_gfa** out) try {
    if (!out || !path) { gfs::set_error("gfs_gfa_parse_file: null argument"); return GFS_ERR_INVALID; }
    *out = nullptr;
    std::unique_ptr<FILE, int (*)(FILE*)> f(fopen(path, "rb"), &fclose);
    if (!f) { gfs::set_error(std::string("cannot open ") + path); return GFS_ERR_INVALID; }
    std::unique_ptr<gfs_gfa> g(new gfs_gfa());
    fseek(f.get(), 0, SEEK_END);
    const long sz = ftell(f.get());
    fseek(f.get(), 0, SEEK_SET);
    g->text.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = sz > 0 ? fread(&g->text[0], 1, (size_t)sz, f.get()) : 0;
    f.reset();
    if (got != g->text.size()) { gfs::set_error(std::string("short read on ") + path); return GFS_ERR_INVALID; }
    int rc = parse_text(g.get());
    if (rc) return rc;
    *out = g.release();
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_gfa_parse_file"); }

extern "C" int gfs_gfa_dims(const gfs_gfa* g, uint64_t* nodes_len, uint64_t* n_nodes, uint64_t* n_edges, uint64_t* n_steps,
                            uint64_t* n_paths) {
    if (!g) { gfs::set_error("gfs_gfa_dims: null"); return GFS_ERR_INVALID; }
    if (nodes_len) *nodes_len = g->present.size();
    if (n_nodes) *n_nodes = g->node_order.size();
    if (n_edges) *n_edges = g->edge_from.size();
    if (n_steps) *n_steps = g->steps.size();
    if (n_paths) *n_paths = g->path_first.size() - 1;
    return GFS_OK;
}

extern "C" int gfs_gfa_arrays(const gfs_gfa* g, const uint8_t** present, const uint64_t** seq_len, const uint64_t** node_order,
                              const uint64_t** edge_from, const uint64_t** edge_to, const uint64_t** steps,
                              const uint64_t** path_first) {
    if (!g) { gfs::set_error("gfs_gfa_arrays: null"); return GFS_ERR_INVALID; }
    if (present) *present = g->present.data();
    if (seq_len) *seq_len = g->seq_len.data();
    if (node_order) *node_order = g->node_order.data();
    if (edge_from) *edge_from = g->edge_from.data();
    if (edge_to) *edge_to = g->edge_to.data();
    if (steps) *steps = g->steps.data();
    if (path_first) *path_first = g->path_first.data();
    return GFS_OK;
}

extern "C" int gfs_gfa_text(const gfs_gfa* g, const char** text, const uint64_t** seq_off, const uint64_t** name_off,
                            const uint64_t** name_len) {
    if (!g) { gfs::set_error("gfs_gfa_text: null"); return GFS_ERR_INVALID; }
    if (text) *text = g->text.data();
    if (seq_off) *seq_off = g->seq_off.data();
    if (name_off) *name_off = g->name_off.data();
    if (name_len) *name_len = g->name_len.data();
    return GFS_OK;
}

extern "C" void gfs_gfa_free(gfs_gfa* g) { delete g; }

// Layout::write_tsv (src/layout.rs:138-163).  coords in Layout order [node][end][dim].
extern "C" int gfs_layout_write_tsv(const double* coords, uint64_t num_nodes, uint32_t dims, const char* path,
                                    uint64_t* bytes_written) try {
    if (!path || (num_nodes && !coords) || dims == 0) { gfs::set_error("gfs_layout_write_tsv: bad argument"); return GFS_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { gfs::set_error(std::string("cannot create ") + path); return GFS_ERR_INVALID; }
    BufWriter w(f);
    auto dim_name = [](uint32_t d) -> char { return d < 4 ? "xyzw"[d] : 'd'; };     // layout.rs:248-256
    w.put("idx", 3);
    for (int end = 0; end < 2; ++end)
        for (uint32_t d = 0; d < dims; ++d) { const char t[3] = {'\t', dim_name(d), end ? '-' : '+'}; w.put(t, 3); }
    w.put("\n", 1);
    const size_t per_row = 24 + (size_t)2 * dims * 400;          // fixed notation of 1e308 needs 309 digits
    for (uint64_t node = 0; node < num_nodes; ++node) {
        char* p = w.room(per_row);
        char* e = p + per_row;
        p = fmt_u64(p, e, node);
        const double* c = coords + node * 2 * dims;
        for (uint32_t k = 0; k < 2 * dims; ++k) { *p++ = '\t'; p = fmt_f64(p, e, c[k]); }
        *p++ = '\n';
        w.advance(p);
    }
    w.flush();
    const bool closed = fclose(f) == 0;
    const bool ok = w.ok && closed;
    if (bytes_written) *bytes_written = w.total;
    if (!ok) { gfs::set_error(std::string("write failed on ") + path); return GFS_ERR_INVALID; }
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_layout_write_tsv"); }

// BidirectedGraph::write_gfa (src/graph_ops.rs:693-738): header, S lines by increasing id, L lines in the
// stored edge order (the reference iterates a HashSet: compare as sets), P lines.
extern "C" int gfs_gfa_write(const char* path, const uint8_t* present, uint64_t nodes_len, const char* seq_blob,
                             const uint64_t* seq_off, const uint64_t* seq_len, const uint64_t* edge_from,
                             const uint64_t* edge_to, uint64_t E, const uint64_t* steps, const uint64_t* path_first, uint64_t P,
                             const char* name_blob, const uint64_t* name_off, const uint64_t* name_len, uint64_t* bytes_written) try {
    if (!path) { gfs::set_error("gfs_gfa_write: null path"); return GFS_ERR_INVALID; }
    FILE* f = fopen(path, "wb");
    if (!f) { gfs::set_error(std::string("cannot create ") + path); return GFS_ERR_INVALID; }
    BufWriter w(f);
    w.put("H\tVN:Z:1.0\n", 11);
    for (uint64_t id = 0; id < nodes_len; ++id)
        if (present[id]) {
            char* p = w.room(32); char* e = p + 32;
            *p++ = 'S'; *p++ = '\t'; p = fmt_u64(p, e, id); *p++ = '\t';
            w.advance(p);
            w.put(seq_blob + seq_off[id], seq_len[id]);
            w.put("\n", 1);
        }
    for (uint64_t k = 0; k < E; ++k) {
        char* p = w.room(64); char* e = p + 64;
        *p++ = 'L'; *p++ = '\t'; p = fmt_u64(p, e, edge_from[k] >> 1); *p++ = '\t'; *p++ = (edge_from[k] & 1) ? '-' : '+'; *p++ = '\t';
        p = fmt_u64(p, e, edge_to[k] >> 1); *p++ = '\t'; *p++ = (edge_to[k] & 1) ? '-' : '+';
        std::memcpy(p, "\t0M\n", 4); p += 4;
        w.advance(p);
    }
    for (uint64_t pi = 0; pi < P; ++pi) {
        w.put("P\t", 2);
        w.put(name_blob + name_off[pi], name_len[pi]);
        w.put("\t", 1);
        for (uint64_t s = path_first[pi]; s < path_first[pi + 1]; ++s) {
            char* p = w.room(24); char* e = p + 24;
            if (s > path_first[pi]) *p++ = ',';
            p = fmt_u64(p, e, steps[s] >> 1); *p++ = (steps[s] & 1) ? '-' : '+';
            w.advance(p);
        }
        w.put("\t*\n", 3);
    }
    w.flush();
    const bool closed = fclose(f) == 0;
    const bool ok = w.ok && closed;
    if (bytes_written) *bytes_written = w.total;
    if (!ok) { gfs::set_error(std::string("write failed on ") + path); return GFS_ERR_INVALID; }
    return GFS_OK;
} catch (...) { return gfs_host_exception("gfs_gfa_write"); }
