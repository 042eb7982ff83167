/* gfasort_cuda.h — C ABI of libgfasort_cuda.so: the B200 (sm_100a) implementation of gfasort's
 * path-guided SGD hot path.  This header is the drop-in boundary: the (otherwise unchanged) Rust
 * host binds exactly these symbols (see INTEGRATION.md for the `extern "C"` block and build.rs).
 *
 * The reference (pangenome/gfasort v0.1.0) has no FFI; the path is three Rust functions whose
 * bodies delegate here.  Each entry point cites the reference code it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function is blocking unless it says "asynchronous"; returns 0 on success, non-zero on
 *     error; gfs_last_error() returns a thread-local message for the last failure.
 *   - the caller owns every host buffer; the library keeps no host pointer after a call returns.
 *   - the library owns device memory behind opaque handles.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails loudly.
 *
 * Node numbering: the host passes dense node indices 0..N-1 in the order of
 * `graph.node_order` restricted to live nodes (src/sgd.rs:276-294).  A step handle is
 * (dense_idx << 1) | is_reverse, mirroring Handle (src/graph.rs:9-19).  dense_idx >= N marks a
 * step on a node missing from the graph: it contributes length 0 to the path offsets
 * (src/sgd.rs:52-54) and terms touching it are skipped (src/sgd.rs:525-538).
 */
#ifndef GFASORT_CUDA_H
#define GFASORT_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFS_OK 0
#define GFS_ERR_INVALID 1      /* bad argument */
#define GFS_ERR_CUDA 2         /* CUDA runtime / launch failure (message has the CUDA error string) */
#define GFS_ERR_NO_DEVICE 3    /* no usable sm_100 device */
#define GFS_ERR_NO_VALID_PATH 4 /* no path with more than one step (src/sgd.rs:250-261, 786-798):
                                   positions are returned unchanged */

/* Opaque device-resident path index.  Replaces `PathIndex` (src/sgd.rs:14-31). */
typedef struct gfs_index gfs_index;
/* Opaque device-resident SGD run (positions + schedule + RNG counters). */
typedef struct gfs_sgd_session gfs_sgd_session;

/* Mirrors PathSGDParams (src/sgd.rs:196-212) and LayoutSGDParams (src/sgd.rs:676-707) field for
 * field (LayoutSGDParams.dimensions travels as the separate `dims` argument).  `nthreads`,
 * `delta` and `progress` are accepted and inert: the GPU picks its own thread count, `delta` is
 * never read by the reference either (src/sgd.rs:554-567 maintains delta_max, nothing consumes it),
 * and progress lines are the host's business.  `seed` keys the Philox4x32-10 stream the way
 * `seed + tid` keys xoshiro256+ in the reference (src/sgd.rs:431-432). */
typedef struct gfs_sgd_params {
    uint64_t iter_max;
    uint64_t iter_with_max_learning_rate;
    uint64_t min_term_updates;
    double delta;
    double eps;
    double eta_max;
    double theta;
    uint64_t space;
    uint64_t space_max;
    uint64_t space_quantization_step;
    double cooling_start;
    uint64_t nthreads;
    uint64_t progress;
    uint64_t seed;
} gfs_sgd_params;

/* Optional launch configuration (NULL = defaults / environment).  Environment variables, because
 * the reference CLI is frozen: GFASORT_DEVICE, GFASORT_THREADS (total GPU threads, 0 = auto),
 * GFASORT_AGGREGATE (0/1 warp-level duplicate-node aggregation, default 1),
 * GFASORT_LAYOUT_F64 (0/1, nD coordinates in double instead of float, default 0),
 * GFASORT_RELABEL (0/1 internal first-appearance node order, default 1),
 * GFASORT_WINDOW (sampling window in steps: 0 = every step ~ U[0,S) exactly like the reference,
 * -1 = auto: 2^20 for graphs whose records exceed 64 MB, else 0), GFASORT_CHUNK (updates per claimed
 * chunk, default 256), GFASORT_COHERENT (0 = off, 1 = default: whole warps sample 32 consecutive steps in
 * window mode, 2..32 = group size), GFASORT_INFLIGHT (terms in flight per thread and pipeline stage, 1 or 2),
 * GFASORT_INDEX_CHUNK (steps per host->device chunk of the index build).  DESIGN.md §4. */
typedef struct gfs_launch_cfg {
    int32_t device;            /* CUDA device ordinal; -1 = current */
    uint32_t total_threads;    /* 0 = auto (full occupancy, capped by the work available) */
    int32_t aggregate;         /* -1 = default */
    int32_t layout_f64;        /* -1 = default */
    uint64_t rng_thread_base;  /* added to the thread id in the Philox counter: rank r of a
                                  multi-GPU run passes r * 2^24 so streams never overlap */
    void* stream;              /* cudaStream_t to launch on (NULL = library-owned stream) */
    void* device_positions;    /* optional caller-owned device buffer for the positions (1D: N doubles;
                                  nD: N*2*dims_stride coordinates), in the library's internal node
                                  order (gfs_index_export_relabel); NULL = library-allocated */
    uint64_t sample_begin;     /* sampled steps are drawn from [sample_begin, sample_end) of the index */
    uint64_t sample_end;       /* (partners still range over whole paths); 0,0 = every step */
} gfs_launch_cfg;

/* Per-run statistics (replaces the reference's stderr progress lines, src/sgd.rs:377-385, 609-611). */
typedef struct gfs_stats {
    uint64_t applied_updates;  /* terms applied (only these count toward min_term_updates, sgd.rs:579) */
    uint64_t attempts;         /* terms sampled (applied + skipped) */
    uint64_t epochs;           /* iter_max + 1 */
    uint64_t launches;         /* kernels launched by the call */
    double kernel_seconds;     /* CUDA-event time of the SGD kernel(s) */
    double h2d_seconds;        /* host->device copies */
    double d2h_seconds;        /* device->host copies */
    double total_seconds;      /* wall time of the call */
    uint32_t grid, block;      /* launch shape used */
    uint32_t coord_bytes;      /* 8 = f64 positions, 4 = f32 */
    uint32_t n_devices;        /* GPUs the call ran on (GFASORT_GPUS) */
    uint64_t window_steps;     /* sampling schedule used: 0 = every step ~ U[0,S) (the reference's), else the
                                  sliding window's length in steps */
    uint32_t coherent;         /* lanes per group that sampled consecutive steps (window mode only; 0 = none, 32 = whole warps) */
    uint32_t syncs_per_epoch;  /* replica reconciles per epoch (n_devices > 1) */
} gfs_stats;

const char* gfs_last_error(void);
/* Library / device facts as a JSON string (static storage). */
const char* gfs_device_info(void);

/* ---- path index -----------------------------------------------------------------------------
 * Replaces PathIndex::from_graph (src/sgd.rs:34-71).  step_handles[S]: all paths' steps
 * concatenated; path_first_step[P+1]: first step of each path, last entry = S; node_len[N]:
 * sequence length by dense idx.  Builds on the GPU, with a segmented exclusive scan, one 16-byte
 * record per step {node<<1|rev, node_len, offset}.  Each path must have < 2^32 steps. */
int gfs_index_build(const uint64_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                    uint64_t S, uint64_t P, uint64_t N, gfs_index** out);
/* Same with 32-bit step handles ((dense_idx << 1) | is_reverse still fits: N < 2^31): half the host->device
 * copy, which is what the index build costs end to end.  The host flattens Vec<Handle> once either way. */
int gfs_index_build32(const uint32_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                      uint64_t S, uint64_t P, uint64_t N, gfs_index** out);
/* Multi-GPU: under GFASORT_GPUS=G (G > 1; the reference CLI is frozen) gfs_index_build /
 * gfs_index_build32 build one shard per device 0..G-1 (concurrently, each device pulling its own steps), with one
 * node order for all shards, and gfs_sgd_1d / gfs_sgd_nd / gfs_sgd_sort_1d / gfs_stress / gfs_index_export on the
 * returned index drive all G GPUs from this one process: replicated positions, terms sharded by step slice,
 * replicas reconciled GFASORT_SYNCS times per epoch (default: once per S applied updates, gfs_default_syncs_per_epoch)
 * over NVLink peer memory.
 * The step array is streamed in chunks (GFASORT_INDEX_CHUNK steps, default 2^24) with the copy of chunk c+1 under
 * the kernel of chunk c; a pageable source goes through pinned bounce buffers filled by GFASORT_COPY_THREADS
 * host threads. */
/* Page-locked host memory for the step array (cudaHostAlloc, portable across devices): flattening Vec<Handle> straight
 * into it lets the copy engine stream it at PCIe rate; a pageable array is staged through bounce buffers by host threads
 * (config 3, 32-bit handles, B200: 0.08 s vs 0.25 s for the whole index build).  Optional. */
int gfs_host_alloc(uint64_t bytes, void** out);
void gfs_host_free(void* p);
/* How the last build went: wall seconds of the whole call, of the streamed copy + K1 phase, K1 kernel seconds
 * (CUDA events; max over shards), of the allocations before and the relabelling after; kernels launched, devices used.
 * Any pointer may be NULL. */
int gfs_index_build_info(const gfs_index* ix, double* build_seconds, double* copy_seconds, double* kernel_seconds,
                         double* alloc_seconds, double* relabel_seconds, uint64_t* launches, uint32_t* n_devices);
/* As gfs_index_build, on `device` and on a sub-range of paths [path_begin, path_end): the shard one GPU of a
 * multi-GPU run owns.  step_handles/path_first_step still describe the whole graph. */
int gfs_index_build_shard(const uint64_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                          uint64_t S, uint64_t P, uint64_t N, uint64_t path_begin, uint64_t path_end,
                          int32_t device, int32_t relabel_mode, const uint32_t* new_of_old, gfs_index** out);
int gfs_index_build_shard32(const uint32_t* step_handles, const uint64_t* path_first_step, const uint32_t* node_len,
                            uint64_t S, uint64_t P, uint64_t N, uint64_t path_begin, uint64_t path_end,
                            int32_t device, int32_t relabel_mode, const uint32_t* new_of_old, gfs_index** out);
/* Internal node numbering.  The library stores positions in the order in which nodes first appear
 * along the paths (path-adjacent nodes then share cache lines); uploads and downloads permute, so
 * callers never see it.  relabel_mode: 0 = keep the caller's dense order, 1 = first-appearance
 * order of this index's own steps (gfs_index_build's default; GFASORT_RELABEL=0 disables), 2 = the
 * permutation new_of_old[N] supplied by the caller — ranks of a multi-GPU run must share one
 * permutation so that their position replicas can be all-reduced element-wise. */
int gfs_index_export_relabel(const gfs_index* ix, uint32_t* new_of_old /*N*/);
/* Relabels an index that was built with relabel_mode 0 (so all ranks can run K1 at the same time and adopt
 * rank 0's order afterwards). */
int gfs_index_apply_relabel(gfs_index* ix, const uint32_t* new_of_old /*N*/);
/* Copies back what PathIndex holds: step_to_position (src/sgd.rs:18) and PathInfo.length (:29).
 * Either pointer may be NULL.  step_to_path / step_to_rank / first_step / step_count are functions
 * of path_first_step alone and stay on the host. */
int gfs_index_export(const gfs_index* ix, uint64_t* step_pos /*S*/, uint64_t* path_len /*P*/);
/* Device-side accessors, for tests: handle (dense_idx<<1|rev) and node length stored per step. */
int gfs_index_export_records(const gfs_index* ix, uint64_t* step_handle /*S*/, uint32_t* step_node_len /*S*/);
int gfs_index_dims(const gfs_index* ix, uint64_t* S, uint64_t* P, uint64_t* N, uint64_t* max_path_steps);
void gfs_index_free(gfs_index* ix);

/* ---- one-call SGD ---------------------------------------------------------------------------
 * gfs_sgd_1d replaces the body of path_linear_sgd after the X init (src/sgd.rs:296-601):
 * x_inout[N] holds the initial positions on entry (src/sgd.rs:286-293) and the final ones on exit.
 * Runs iter_max+1 epochs of exactly min_term_updates applied updates each with etas[0..=iter_max]
 * (src/sgd.rs:617-638), cooling when epoch > floor(cooling_start*iter_max) (src/sgd.rs:393-396). */
int gfs_sgd_1d(const gfs_index* ix, const gfs_sgd_params* params, double* x_inout, gfs_stats* stats);
/* gfs_sgd_nd replaces the body of path_linear_sgd_layout after the coordinate init
 * (src/sgd.rs:856-1172).  coords_inout is in Layout order coords[node*2*dims + end*dims + dim]
 * (src/layout.rs:14-24, 52-61), N*2*dims doubles.  1 <= dims <= 8. */
int gfs_sgd_nd(const gfs_index* ix, const gfs_sgd_params* params, uint32_t dims, double* coords_inout,
               gfs_stats* stats);
/* Same with an explicit launch configuration. */
int gfs_sgd_1d_cfg(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg,
                   double* x_inout, gfs_stats* stats);
int gfs_sgd_nd_cfg(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg, uint32_t dims,
                   double* coords_inout, gfs_stats* stats);

/* ---- sampled stress -------------------------------------------------------------------------
 * Replaces calculate_layout_stress (src/sgd.rs:1196-1283) on a fixed Philox(seed) sample:
 * sample k draws step_a uniformly, a uniform partner rank on the same path, skips equal ranks and
 * zero path distance, measures the Euclidean distance between the + ends.
 * rms_rel = sqrt(mean((dl-dp)^2/dp^2)) (the reference's value); mean_abs_rel = mean(|dl-dp|/dp)
 * (BASELINE.json's form).  dims == 1 with coords = x (N doubles, one per node) measures a 1D sort;
 * dims >= 1 with layout_order != 0 takes a Layout-order array (N*2*dims). */
int gfs_stress(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords, uint64_t samples,
               uint64_t seed, double* rms_rel, double* mean_abs_rel, uint64_t* counted);

/* One shard's share of the same measurement (one process per GPU): sample k draws its step over the WHOLE graph
 * (total_steps); this index — whose local step 0 is global step step_offset — evaluates the samples that land in
 * [step_begin, step_end) and ADDS {sum (dl-dp)^2/dp^2, sum |dl-dp|/dp, count} to sums3.  Summing the ranks' triples
 * gives exactly the single-GPU sample over all paths. */
int gfs_stress_partial(const gfs_index* ix, uint32_t dims, int32_t layout_order, const double* coords, uint64_t samples,
                       uint64_t seed, uint64_t total_steps, uint64_t step_offset, uint64_t step_begin, uint64_t step_end,
                       double* sums3);

/* ---- session API (device-resident runs, multi-GPU) -------------------------------------------
 * A session holds the positions on the device between calls so that the host can run the schedule
 * in slices and reconcile replicas between slices (NCCL all-reduce on gfs_sgd_session_positions). */
int gfs_sgd_session_create(const gfs_index* ix, const gfs_sgd_params* params, uint32_t dims /*0 = 1D Y*/,
                           const gfs_launch_cfg* cfg, gfs_sgd_session** out);
int gfs_sgd_session_upload(gfs_sgd_session* s, const double* positions);     /* host -> device (converts for f32) */
int gfs_sgd_session_download(gfs_sgd_session* s, double* positions);         /* device -> host */
/* Asynchronous: enqueue epochs [epoch_begin, epoch_end) on the session's stream; within each epoch
 * run only slice `slice` of `n_slices` equal parts of min_term_updates (n_slices = 1: whole epochs). */
int gfs_sgd_session_run(gfs_sgd_session* s, uint64_t epoch_begin, uint64_t epoch_end, uint32_t slice,
                        uint32_t n_slices);
int gfs_sgd_session_sync(gfs_sgd_session* s);
/* Asynchronous device-side snapshot / restore of the positions (rerun a schedule from the same start). */
int gfs_sgd_session_save(gfs_sgd_session* s);
int gfs_sgd_session_restore(gfs_sgd_session* s);
/* Device pointer + element count + element size (8 or 4) of the position buffer. */
int gfs_sgd_session_positions(gfs_sgd_session* s, void** dev_ptr, uint64_t* n_elems, uint32_t* elem_bytes);
int gfs_sgd_session_stats(gfs_sgd_session* s, gfs_stats* stats);             /* synchronises first */
void gfs_sgd_session_destroy(gfs_sgd_session* s);

/* ---- order by position -----------------------------------------------------
 * The host side of path_sgd_sort (src/sgd.rs:659-671) on the device: order[k] = dense idx of the node
 * with the k-th smallest position; stable, ties by dense idx (the reference's stable sort starts from
 * HashMap iteration order, so its tie order is unspecified); -0.0 == +0.0; NaN last.  n < 2^32. */
int gfs_sort_positions(const double* x /*host, n*/, uint64_t n, uint32_t* order /*host, n*/);
/* Same on a 1D session's current positions (no position download, no host sort). */
int gfs_sgd_session_sort(gfs_sgd_session* s, uint32_t* order /*host, N*/);
/* path_sgd_sort in one call: gfs_sgd_1d followed by the sort; x_inout as in gfs_sgd_1d. */
int gfs_sgd_sort_1d(const gfs_index* ix, const gfs_sgd_params* params, double* x_inout, uint32_t* order_out /*N*/,
                    gfs_stats* stats);

/* ---- host steps downstream of `Y` (CPU code, no device needed) -----------------
 * Linear-time versions of the reference's O(N*E) grooming and heads-first topological sort, emitting the
 * same orders.  Graph as flat arrays: present[nodes_len] (1 = node id exists), E unique edges as
 * (edge_from[e], edge_to[e]) handles (id << 1 | is_reverse), paths as steps[] + path_first[P+1].
 * Outputs hold one handle per present node. */
/* find_head_nodes (src/graph_ops.rs:1138-1183): forward handles with no incoming edge, by earliest path rank. */
int gfs_find_head_nodes(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from, const uint64_t* edge_to,
                        uint64_t E, const uint64_t* steps, const uint64_t* path_first, uint64_t P,
                        uint64_t* heads_out, uint64_t* n_heads);
/* groom(use_bfs = true) (src/groom.rs:49-275): nodes in increasing id, as a reverse handle when flipped. */
int gfs_groom_order(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from, const uint64_t* edge_to,
                    uint64_t E, const uint64_t* steps, const uint64_t* path_first, uint64_t P,
                    uint64_t* order_out, uint64_t* n_flipped);
/* exact_odgi_topological_order(use_heads = true, use_tails = false) (src/graph_ops.rs:1232-1485). */
int gfs_topological_order(const uint8_t* present, uint64_t nodes_len, const uint64_t* edge_from, const uint64_t* edge_to,
                          uint64_t E, const uint64_t* steps, const uint64_t* path_first, uint64_t P,
                          uint64_t* order_out, uint64_t* n_out);

/* Flat form of the handle rewrites after every pipeline step — apply_ordering (src/graph_ops.rs:1939-2025),
 * apply_node_id_mapping (:36-84), the flips of apply_grooming_with_reorder (src/groom.rs:533-605) — in place over
 * n handles (path steps or edge ends): orientation ^= flip[id] (flip may be NULL; indexed by the old id), then
 * id -> new_id[id] where id < table_len and new_id[id] != UINT64_MAX (other handles keep their id).  Multi-threaded. */
int gfs_remap_handles(uint64_t* handles, uint64_t n, const uint64_t* new_id, uint64_t table_len, const uint8_t* flip,
                      uint64_t flip_len);

/* The edge set of a graph given only by its paths (synthetic inputs): every pair of consecutive steps, one edge per
 * {edge, complement} class as add_edge keeps them (src/graph_ops.rs:626-638), in order of first occurrence and in the form
 * a path first walks it.  Pointers from gfs_edge_list_get stay valid until gfs_edge_list_free. */
typedef struct gfs_edge_list gfs_edge_list;
int gfs_edges_from_paths(const uint64_t* steps, const uint64_t* path_first, uint64_t P, gfs_edge_list** out);
int gfs_edge_list_get(const gfs_edge_list* el, const uint64_t** edge_from, const uint64_t** edge_to, uint64_t* n_edges);
void gfs_edge_list_free(gfs_edge_list* el);

/* ---- flat ingest and buffered writers (CPU code) ------------------------------
 * gfs_gfa_parse_*: the CLI's parse_gfa (src/bin/gfasort.rs:88-167) in one pass, straight into flat arrays:
 * present / seq_len indexed by node id, node_order (add_node order, src/graph_ops.rs:613-623), edges unique
 * per {edge, complement} (add_edge, :626-638), concatenated path steps.  Sequences and path names are
 * (offset, length) pairs into the kept text.  Pointers stay valid until gfs_gfa_free. */
typedef struct gfs_gfa gfs_gfa;
int gfs_gfa_parse_file(const char* path, gfs_gfa** out);
int gfs_gfa_parse_text(const char* text, uint64_t len, gfs_gfa** out);
int gfs_gfa_dims(const gfs_gfa* g, uint64_t* nodes_len, uint64_t* n_nodes, uint64_t* n_edges, uint64_t* n_steps,
                 uint64_t* n_paths);
int gfs_gfa_arrays(const gfs_gfa* g, const uint8_t** present, const uint64_t** seq_len, const uint64_t** node_order,
                   const uint64_t** edge_from, const uint64_t** edge_to, const uint64_t** steps, const uint64_t** path_first);
int gfs_gfa_text(const gfs_gfa* g, const char** text, const uint64_t** seq_off, const uint64_t** name_off,
                 const uint64_t** name_len);
void gfs_gfa_free(gfs_gfa* g);
/* Layout::write_tsv (src/layout.rs:138-163), byte for byte, through one buffer. coords: Layout order. */
int gfs_layout_write_tsv(const double* coords, uint64_t num_nodes, uint32_t dims, const char* path, uint64_t* bytes_written);
/* BidirectedGraph::write_gfa (src/graph_ops.rs:693-738), buffered; L lines in the given edge order. */
int gfs_gfa_write(const char* path, const uint8_t* present, uint64_t nodes_len, const char* seq_blob, const uint64_t* seq_off,
                  const uint64_t* seq_len, const uint64_t* edge_from, const uint64_t* edge_to, uint64_t E, const uint64_t* steps,
                  const uint64_t* path_first, uint64_t P, const char* name_blob, const uint64_t* name_off,
                  const uint64_t* name_len, uint64_t* bytes_written);

/* ---- replica reconcile (multi-GPU) ---------------------------------------------
 * The exchange step of a replicated run is one all-reduce(sum) of `buf` (2n floats) that the host
 * issues (NCCL) between these two asynchronous kernels, all on `stream`:
 *   pack:  buf[i] = x[i] - x_sync[i];  buf[n+i] = (x[i] != x_sync[i])
 *   apply: x[i] = x_sync[i] + buf[i] / max(buf[n+i], 1);  x_sync[i] = x[i]
 * = the mean of the displacements over the replicas that moved the element since the last sync
 * (equal to the plain replica mean wherever every replica moved it).  x / x_sync: device pointers to
 * n elements of elem_bytes (8 = f64 positions, 4 = f32 coordinates). */
int gfs_reconcile_pack(const void* x, const void* x_sync, uint64_t n, uint32_t elem_bytes, float* buf, void* stream);
int gfs_reconcile_apply(void* x, void* x_sync, uint64_t n, uint32_t elem_bytes, const float* buf, void* stream);

/* ---- replica reconcile over peer memory (multi-GPU; opt-in) -------
 * The same exchange as pack -> all-reduce -> apply, as ONE kernel per rank over NVLink peer memory: every rank
 * maps every other rank's replica, reduces its 1/G slice of the elements across the G replicas and stores the
 * result into all of them, between two in-kernel barriers (flags in peer memory, bounded spins).
 * A region holds one rank's replica x[n], the common base x_sync[n] of the last reconcile (elem_bytes = 8: f64 positions,
 * 4: f32 coordinates; after the first reconcile only the rank's OWN slice [n r/G, n (r+1)/G) of x_sync is kept up to
 * date — nobody else ever reads it), a second snapshot for the overlapped form and the barrier flags in one allocation; give gfs_p2p_region_ptrs()'s x to the session as
 * gfs_launch_cfg.device_positions.  One process per GPU: exchange gfs_p2p_region_ipc_handle() blobs (all-gather)
 * and call gfs_p2p_region_connect_ipc; one process driving several GPUs (or several replicas on one GPU):
 * gfs_p2p_region_connect_local.  max_blocks = 0: one block per SM. */
#define GFS_P2P_MAX_RANKS 16
#define GFS_P2P_HANDLE_BYTES 80    /* cudaIpcMemHandle_t + {n, elem_bytes, blocks}: peers must agree on the shape */
typedef struct gfs_p2p_region gfs_p2p_region;
int gfs_p2p_region_create(int32_t device, uint64_t n, uint32_t elem_bytes, uint32_t max_blocks, gfs_p2p_region** out);
int gfs_p2p_region_ptrs(gfs_p2p_region* r, void** x, void** x_sync, uint64_t* region_bytes);
int gfs_p2p_region_ipc_handle(gfs_p2p_region* r, uint8_t* handle /*GFS_P2P_HANDLE_BYTES*/);
int gfs_p2p_region_connect_ipc(gfs_p2p_region* r, const uint8_t* handles /*world x GFS_P2P_HANDLE_BYTES, rank order*/,
                               uint32_t world, uint32_t rank);
int gfs_p2p_region_connect_local(gfs_p2p_region* const* regions /*world, rank order*/, uint32_t world);
/* Asynchronous on `stream`: x_sync <- x.  Once, after the initial positions were uploaded into x (gfs_sgd_session_upload
 * with the region's x as gfs_launch_cfg.device_positions): all replicas start from equal snapshots. */
int gfs_p2p_region_snapshot(gfs_p2p_region* r, void* stream);
/* Asynchronous on `stream`: x <- x_sync + (sum of the replicas' displacements) / (#replicas that moved the element)
 * on every replica; the owner of a slice stores the same value into its x_sync.  Every rank calls it once per
 * reconcile, in the same order. */
int gfs_p2p_reconcile(gfs_p2p_region* r, void* stream);
/* Replicas that share ONE device (tests): all `world` (<= 8) ranks as one cooperative launch, block group g playing
 * rank g — kernels of one GPU that wait on one another must not be separate launches. */
int gfs_p2p_reconcile_local(gfs_p2p_region* const* regions /*world, rank order, one device*/, uint32_t world, void* stream);
/* Overlapped form (gfs_replica_run with GFASORT_OVERLAP=1/2; off by default): gfs_p2p_region_snapshot_x copies the replica into the region's
 * snapshot on the SGD's stream; gfs_p2p_reconcile_async, on a second stream that waits for that copy, exchanges the
 * SNAPSHOTS and adds (new common base - own snapshot) to every rank's live replica with red.add, so the next SGD slice runs
 * during the exchange; the next snapshot must wait for it.  The _local form is the same kernel for replicas sharing a device. */
int gfs_p2p_region_snap_ptr(gfs_p2p_region* r, void** x_snap);
int gfs_p2p_region_snapshot_x(gfs_p2p_region* r, void* stream);
int gfs_p2p_reconcile_async(gfs_p2p_region* r, void* stream);
int gfs_p2p_reconcile_async_local(gfs_p2p_region* const* regions, uint32_t world, void* stream);
/* Blocking: GFS_ERR_CUDA if a barrier of an earlier reconcile timed out on ANY rank (a rank missing, kernels not
 * co-resident).  After a timeout the replicas are undefined and every later reconcile returns at once: the run failed. */
int gfs_p2p_region_check(gfs_p2p_region* r);
void gfs_p2p_region_free(gfs_p2p_region* r);

/* ---- replicated multi-GPU runs -------------------------------------------------
 * Terms shard, positions do not.  Rank r of G samples the steps of its slice [S r/G, S (r+1)/G) of the concatenated
 * step array and needs the records of just the paths that slice overlaps; it applies its share of every epoch's
 * min_term_updates (exact in sum over ranks) to its own full replica of the positions; replicas are reconciled
 * (moved-replica mean, gfs_p2p_*) syncs_per_epoch times per epoch. */
typedef struct gfs_shard_plan {
    uint64_t sample_begin, sample_end;   /* global step range this rank samples from */
    uint64_t path_begin, path_end;       /* paths whose records it needs (gfs_index_build_shard) */
    uint64_t first_step;                 /* global step index of path_begin's first step */
} gfs_shard_plan;
int gfs_shard_plan_make(const uint64_t* path_first_step /*P+1*/, uint64_t P, uint32_t rank, uint32_t world, gfs_shard_plan* out);
uint64_t gfs_shard_epoch_quota(uint64_t min_term_updates, const gfs_shard_plan* plan, uint64_t total_steps);
/* Reconciles per epoch used when a caller passes syncs_per_epoch = 0 (and by GFASORT_GPUS runs unless GFASORT_SYNCS says
 * otherwise): one per total_steps applied updates of the whole run — 1 for `Y` (min_term_updates = S), 10 for `L` (10 S). */
uint32_t gfs_default_syncs_per_epoch(uint64_t min_term_updates, uint64_t total_steps);
/* One rank: a session on `shard` (built for plan->path_begin..path_end) whose positions live in a peer region.
 * `params` are the WHOLE run's (the quota is derived here); cfg may be NULL (total_threads, aggregate, layout_f64 are
 * honoured).  One process per GPU: create, exchange gfs_replica_ipc_handle blobs (GFS_P2P_HANDLE_BYTES each, rank
 * order), gfs_replica_connect_ipc.  One process, G devices: gfs_replica_connect_local — or simply GFASORT_GPUS. */
typedef struct gfs_replica gfs_replica;
int gfs_replica_create(const gfs_index* shard, const gfs_sgd_params* params, uint32_t dims, const gfs_launch_cfg* cfg,
                       const gfs_shard_plan* plan, uint64_t total_steps, uint32_t rank, uint32_t world,
                       uint32_t syncs_per_epoch /*0 = gfs_default_syncs_per_epoch*/, gfs_replica** out);
int gfs_replica_ipc_handle(gfs_replica* r, uint8_t* blob /*GFS_P2P_HANDLE_BYTES*/);
int gfs_replica_connect_ipc(gfs_replica* r, const uint8_t* blobs /*world x GFS_P2P_HANDLE_BYTES*/, uint32_t world, uint32_t rank);
int gfs_replica_connect_local(gfs_replica* const* replicas /*world, rank order, distinct devices*/, uint32_t world);
int gfs_replica_upload(gfs_replica* r, const double* positions);            /* every rank uploads the same positions */
/* Asynchronous: epochs [epoch_begin, epoch_end), syncs_per_epoch (SGD slice, reconcile) pairs each. */
int gfs_replica_run(gfs_replica* r, uint64_t epoch_begin, uint64_t epoch_end);
/* Asynchronous: the rank's stream waits for its last overlapped reconcile (before timing events / reads on that stream).
 * GFASORT_OVERLAP=0 selects the stop-the-world reconcile instead (one kernel on the rank's stream after every slice). */
int gfs_replica_flush(gfs_replica* r);
int gfs_replica_sync(gfs_replica* r);                                        /* blocking; reports reconcile time-outs */
int gfs_replica_download(gfs_replica* r, double* positions);
int gfs_replica_stats(gfs_replica* r, gfs_stats* stats);                     /* this rank's share; synchronises first */
/* The stream the rank's kernels run on and its replica (device pointer, element count, element bytes): for timing
 * with CUDA events on the launching stream and for checks.  Any pointer may be NULL. */
int gfs_replica_stream(gfs_replica* r, void** stream, void** dev_positions, uint64_t* n_elems, uint32_t* elem_bytes);
void gfs_replica_destroy(gfs_replica* r);

/* ---- synthetic pangenome graphs (bench / tests input) -------------------------
 * Seeded bubble-chain generator writing the C-ABI's own flat inputs.  Two calls: sizes, then fill. */
typedef struct gfs_synth_spec {
    uint64_t num_nodes;   /* N (exact) */
    uint64_t num_paths;   /* P */
    uint64_t seed;
    uint32_t permute_ids; /* 1 = randomly permute node ids (scrambled initial order) */
    uint32_t pinned;      /* 1 = allocate the step array in page-locked host memory (cudaHostAlloc) */
} gfs_synth_spec;
typedef struct gfs_synth_graph gfs_synth_graph;
int gfs_synth_create(const gfs_synth_spec* spec, gfs_synth_graph** out);
/* Only paths [path_begin, path_end) are materialised (a rank's shard); node_len covers all N nodes. */
int gfs_synth_create_range(const gfs_synth_spec* spec, uint64_t path_begin, uint64_t path_end,
                           gfs_synth_graph** out);
/* Step count of each of the P paths, without materialising steps. */
int gfs_synth_path_counts(const gfs_synth_spec* spec, uint64_t* counts /*P*/);
int gfs_synth_dims(const gfs_synth_graph* g, uint64_t* S, uint64_t* P, uint64_t* N);
/* Pointers into the generator's own storage (valid until gfs_synth_free). */
int gfs_synth_arrays(const gfs_synth_graph* g, const uint64_t** step_handles, const uint64_t** path_first_step,
                     const uint32_t** node_len);
void gfs_synth_free(gfs_synth_graph* g);

/* ---- debug / parity hooks (used by tests/ only) ----------------------------------------------
 * Device evaluation of the reference's scalar helpers, for bit-exact checks against the oracle. */
int gfs_debug_fast_precise_pow(const double* a, const double* b, double* out, uint64_t n);
int gfs_debug_dirty_zipf(const uint64_t* zmax, const double* theta, const double* zeta, const double* u,
                         uint64_t* out, uint64_t n);    /* min = 1, zeta2theta = 1 + fpp(0.5, theta) */
int gfs_debug_philox(const uint32_t* ctr4, const uint32_t* key2, uint32_t* out4, uint64_t n);
/* Sampled terms of thread `tid`, attempts [attempt0, attempt0+count): valid flag, step_a, step_b,
 * flags (bit0 other_end_a, bit1 other_end_b), term distance.  `epoch` selects eta/theta/cooling. */
int gfs_debug_trace_terms(const gfs_index* ix, const gfs_sgd_params* params, int32_t nd, uint64_t epoch,
                          uint32_t tid, uint64_t attempt0, uint64_t count, uint8_t* valid, uint64_t* step_a,
                          uint64_t* step_b, uint8_t* flags, double* dist);
/* Host-computed schedule and zeta table the kernels use (etas: iter_max+1; zetas: *n entries). */
int gfs_debug_schedule(const gfs_sgd_params* params, double* etas);
int gfs_debug_zetas(const gfs_index* ix, const gfs_sgd_params* params, double* zetas, uint64_t cap, uint64_t* n);
/* the same table without an index or a device (host arithmetic only): for a longest path of max_path_steps steps */
int gfs_debug_zetas_host(const gfs_sgd_params* params, uint64_t max_path_steps, double* zetas, uint64_t cap, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* GFASORT_CUDA_H */
