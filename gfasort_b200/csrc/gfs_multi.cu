// gfs_multi.cu — replicated multi-GPU runs behind the C ABI.
//
// Terms shard, positions do not: every term lives inside one path (reference src/sgd.rs:445, 502-503), any path can
// touch any node.  So rank r of G samples only the steps of its slice [S r/G, S (r+1)/G) of the concatenated step
// array, holds the records of just the paths that slice overlaps (partners never leave them), runs its share
// min_term_updates |slice| / S of every epoch (exact in sum; the global sampling distribution stays uniform over
// steps, src/sgd.rs:435, 444), and keeps a full replica of the positions.  Replicas are reconciled `syncs` times per
// epoch by the peer-memory kernel of gfs_p2p.cu (moved-replica mean over NVLink).
//
//   gfs_shard_plan_make / gfs_shard_epoch_quota   who samples what (host arithmetic, no device)
//   gfs_replica_*                                  one rank: session + peer region + the epoch loop, for hosts that run
//                                                  one process per GPU (exchange gfs_replica_ipc_handle blobs, connect)
//   gfs_multi_run_whole                            all ranks driven from ONE process: what gfs_sgd_1d / gfs_sgd_nd do on
//                                                  an index built under GFASORT_GPUS=G (the frozen reference CLI has no
//                                                  flag for it, so the GPU count is an environment variable)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gfs_internal.h"

using namespace gfs;

// ---------------------------------------------------------------------------------------------
// shard plan
// ---------------------------------------------------------------------------------------------
extern "C" int gfs_shard_plan_make(const uint64_t* path_first_step, uint64_t P, uint32_t rank, uint32_t world, gfs_shard_plan* out) {
    if (!path_first_step || !out || world == 0 || rank >= world) { set_error("gfs_shard_plan_make: bad argument"); return GFS_ERR_INVALID; }
    std::memset(out, 0, sizeof *out);
    const uint64_t S = path_first_step[P];
    const uint64_t b = (uint64_t)(((unsigned __int128)S * rank) / world);
    const uint64_t e = (uint64_t)(((unsigned __int128)S * (rank + 1)) / world);
    out->sample_begin = out->sample_end = b;
    if (e <= b) return GFS_OK;                                   // more ranks than steps: an empty shard
    // path of step b and of step e-1: the last p with first_step[p] <= step
    const uint64_t* fs = path_first_step;
    const uint64_t pb = (uint64_t)(std::upper_bound(fs, fs + P + 1, b) - fs) - 1;
    const uint64_t pe = (uint64_t)(std::upper_bound(fs, fs + P + 1, e - 1) - fs);
    out->sample_end = e;
    out->path_begin = pb; out->path_end = pe;
    out->first_step = fs[pb];
    return GFS_OK;
}

// Reconciles per epoch when the caller does not say (syncs_per_epoch = 0): one per S applied updates of the whole run,
// S = the graph's step count — once per epoch for `Y` (min_term_updates = S, ygs.rs:60-70), ten times for `L`
// (min_term_updates = 10 S, sgd.rs:736-745).  Replicas that run 10 S updates apart drift: the 2D layout of config 4 on
// 8 GPUs ended at 6.7x the one-GPU stress with one reconcile per epoch (profiles/r2_bench.md).
extern "C" uint32_t gfs_default_syncs_per_epoch(uint64_t min_term_updates, uint64_t total_steps) {
    if (total_steps == 0) return 1;
    const uint64_t k = (min_term_updates + total_steps / 2) / total_steps;
    return (uint32_t)std::min<uint64_t>(std::max<uint64_t>(k, 1), 64);
}

extern "C" uint64_t gfs_shard_epoch_quota(uint64_t min_term_updates, const gfs_shard_plan* plan, uint64_t total_steps) {
    if (!plan || total_steps == 0) return 0;
    const unsigned __int128 M = min_term_updates;
    return (uint64_t)(M * plan->sample_end / total_steps) - (uint64_t)(M * plan->sample_begin / total_steps);
}

// ---------------------------------------------------------------------------------------------
// one rank of a replicated run
// ---------------------------------------------------------------------------------------------
struct gfs_replica {
    gfs_sgd_session* s = nullptr;
    gfs_p2p_region* region = nullptr;
    uint32_t rank = 0, world = 1, syncs = 1;
    uint64_t reconciles = 0;
    uint64_t n_epochs = 0;
    // overlapped reconcile (GFASORT_OVERLAP, gfs_p2p.cu): the exchange runs on `side` over a snapshot while the next
    // SGD slice runs on the session's stream
    bool overlap = false, rc_pending = false;
    bool join_at_once = false;                 // GFASORT_OVERLAP=3: the overlapped arithmetic without the overlap (an A/B switch)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_rc = nullptr;
    std::vector<cudaEvent_t> trace;            // GFASORT_RC_TRACE=1: (start, end) of every overlapped exchange on `side`
};

extern "C" void gfs_replica_destroy(gfs_replica* r) {
    if (!r) return;
    if (r->s) cudaSetDevice(r->s->device);
    if (r->side) { cudaStreamSynchronize(r->side); cudaStreamDestroy(r->side); }
    if (r->ev_snap) cudaEventDestroy(r->ev_snap);
    if (r->ev_rc) cudaEventDestroy(r->ev_rc);
    for (cudaEvent_t e : r->trace) cudaEventDestroy(e);
    if (r->s) gfs_sgd_session_destroy(r->s);
    if (r->region) gfs_p2p_region_free(r->region);
    delete r;
}

extern "C" int gfs_replica_create(const gfs_index* shard, const gfs_sgd_params* params, uint32_t dims, const gfs_launch_cfg* cfg,
                                  const gfs_shard_plan* plan, uint64_t total_steps, uint32_t rank, uint32_t world,
                                  uint32_t syncs_per_epoch, gfs_replica** out) {
    if (!out) { set_error("gfs_replica_create: out is null"); return GFS_ERR_INVALID; }
    *out = nullptr;
    if (!shard || !params || !plan || world == 0 || rank >= world || world > GFS_P2P_MAX_RANKS) { set_error("gfs_replica_create: bad argument"); return GFS_ERR_INVALID; }
    if (!shard->shards.empty()) { set_error("gfs_replica_create: pass one shard, not a multi-GPU index"); return GFS_ERR_INVALID; }
    if (dims > 8) { set_error("gfs_replica_create: dims must be <= 8"); return GFS_ERR_INVALID; }
    if (plan->sample_end < plan->sample_begin || plan->sample_begin < plan->first_step ||
        plan->sample_end - plan->first_step > shard->S) { set_error("gfs_replica_create: the plan's step slice is outside the shard"); return GFS_ERR_INVALID; }
    gfs_replica* r = new gfs_replica();
    r->rank = rank; r->world = world;
    r->syncs = syncs_per_epoch ? syncs_per_epoch : gfs_default_syncs_per_epoch(params->min_term_updates, total_steps);
    auto fail = [&](int code) { gfs_replica_destroy(r); return code; };
    const int f64 = dims == 0 ? 1 : (cfg && cfg->layout_f64 >= 0 ? cfg->layout_f64 : (int)env_long("GFASORT_LAYOUT_F64", 0));
    const uint64_t n_elems = dims == 0 ? shard->N : shard->N * 2 * coord_stride(dims);
    int rc = gfs_p2p_region_create(shard->device, n_elems, f64 ? 8 : 4, 0, &r->region);
    if (rc) return fail(rc);
    void* x = nullptr;
    rc = gfs_p2p_region_ptrs(r->region, &x, nullptr, nullptr);
    if (rc) return fail(rc);
    gfs_sgd_params p = *params;
    p.min_term_updates = gfs_shard_epoch_quota(params->min_term_updates, plan, total_steps);
    gfs_launch_cfg c{};
    c.device = shard->device; c.total_threads = cfg ? cfg->total_threads : 0;
    c.aggregate = cfg ? cfg->aggregate : -1; c.layout_f64 = f64;
    c.rng_thread_base = (uint64_t)rank << 24;                       // disjoint Philox streams: the analogue of seed + tid (sgd.rs:431)
    c.stream = nullptr;
    c.device_positions = x;
    c.sample_begin = plan->sample_begin - plan->first_step;        // index-local step range
    c.sample_end = plan->sample_end - plan->first_step;
    if (c.sample_end <= c.sample_begin) { set_error("gfs_replica_create: rank " + std::to_string(rank) + " has no steps to sample (more GPUs than steps)"); return fail(GFS_ERR_INVALID); }
    rc = gfs_sgd_session_create(shard, &p, dims, &c, &r->s);
    if (rc) return fail(rc);
    r->n_epochs = params->iter_max + 1;
    {   // GFASORT_OVERLAP: 0 (default) = stop-the-world reconcile; 1 = overlapped when the exchange is short against the SGD
        // slice it runs beside (>= GFASORT_OVERLAP_MIN_RATIO, default 40, updates per rank and slice for every 16-byte
        // vector of the replica); 2 = always overlapped; 3 = the overlapped arithmetic, joined at once (A/B switch).
        // Why it is off by default (profiles/r2_experiments.md §7): corrections that arrive late are applied on top of what
        // the replica has meanwhile done about the same deviation by itself; with the capped step size (mu = 1) of most of
        // the schedule that double correction does not decay.  Measured on config 3: exchange done within ~10-20 % of the
        // slice (2 and 4 GPUs) -> same stress, +1 %; ~half of the slice (8 GPUs) -> +16 % stress; a whole slice late -> x2.
        const long ov = env_long("GFASORT_OVERLAP", 0);
        const uint64_t per_slice = p.min_term_updates / std::max(1u, r->syncs);
        const uint64_t nvec = (n_elems * (f64 ? 8 : 4) + 15) / 16;
        r->overlap = world > 1 && (ov >= 2 || (ov == 1 && per_slice >= (uint64_t)env_long("GFASORT_OVERLAP_MIN_RATIO", 40) * nvec));
        r->join_at_once = ov == 3;
    }
    if (r->overlap) {
        // highest priority: the exchange's few blocks must become resident the moment an SM has room, not behind the
        // next SGD launch's 444 blocks
        int pr_lo = 0, pr_hi = 0;
        cudaDeviceGetStreamPriorityRange(&pr_lo, &pr_hi);
        cudaError_t e = cudaStreamCreateWithPriority(&r->side, cudaStreamNonBlocking, env_long("GFASORT_RC_PRIORITY", 1) ? pr_hi : pr_lo);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_snap, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&r->ev_rc, cudaEventDisableTiming);
        if (e != cudaSuccess) { set_error(std::string("gfs_replica_create: ") + cudaGetErrorString(e)); return fail(GFS_ERR_CUDA); }
    }
    *out = r;
    return GFS_OK;
}

extern "C" int gfs_replica_ipc_handle(gfs_replica* r, uint8_t* blob) {
    if (!r) { set_error("gfs_replica_ipc_handle: null replica"); return GFS_ERR_INVALID; }
    return gfs_p2p_region_ipc_handle(r->region, blob);
}
extern "C" int gfs_replica_connect_ipc(gfs_replica* r, const uint8_t* blobs, uint32_t world, uint32_t rank) {
    if (!r) { set_error("gfs_replica_connect_ipc: null replica"); return GFS_ERR_INVALID; }
    if (world != r->world || rank != r->rank) { set_error("gfs_replica_connect_ipc: world / rank differ from gfs_replica_create"); return GFS_ERR_INVALID; }
    return gfs_p2p_region_connect_ipc(r->region, blobs, world, rank);
}
extern "C" int gfs_replica_connect_local(gfs_replica* const* replicas, uint32_t world) {
    if (!replicas || world == 0 || world > GFS_P2P_MAX_RANKS) { set_error("gfs_replica_connect_local: bad argument"); return GFS_ERR_INVALID; }
    std::vector<gfs_p2p_region*> regs(world);
    for (uint32_t g = 0; g < world; ++g) {
        if (!replicas[g] || replicas[g]->world != world || replicas[g]->rank != g) { set_error("gfs_replica_connect_local: replicas must come in rank order"); return GFS_ERR_INVALID; }
        regs[g] = replicas[g]->region;
        for (uint32_t h = 0; h < g; ++h)
            if (replicas[h]->s->device == replicas[g]->s->device) {
                set_error("gfs_replica_connect_local: two replicas share a device (their reconcile kernels would wait on one another inside one GPU)");
                return GFS_ERR_INVALID;
            }
    }
    return gfs_p2p_region_connect_local(regs.data(), world);
}

extern "C" int gfs_replica_upload(gfs_replica* r, const double* positions) {
    if (!r) { set_error("gfs_replica_upload: null replica"); return GFS_ERR_INVALID; }
    int rc = gfs_sgd_session_upload(r->s, positions);
    // the snapshot is taken on the session's stream, behind the upload: a snapshot racing the first SGD launch would
    // leave the ranks with different x_sync, which this reconcile (it never re-reads absolute positions) cannot repair
    if (!rc) rc = gfs_p2p_region_snapshot(r->region, r->s->stream);
    if (!rc) rc = gfs_sgd_session_sync(r->s);
    return rc;
}

// the session's stream waits (on the device) for the reconcile still in flight on the side stream
static int replica_join_reconcile(gfs_replica* r) {
    if (r->rc_pending) {
        GFS_CUDA(cudaStreamWaitEvent(r->s->stream, r->ev_rc, 0));
        r->rc_pending = false;
    }
    return GFS_OK;
}

// Slice k of epoch e and the reconcile behind it (asynchronous).
// Overlapped (GFASORT_OVERLAP=1/2; off by default, see gfs_replica_create): after a slice the replica is copied to a
// snapshot and the next slice starts at once; the exchange over the snapshots runs beside it on a second stream and adds
// its corrections to the live replicas (gfs_p2p.cu, rc_p2p_async).  The next snapshot waits for it.
static int replica_run_slice(gfs_replica* r, uint64_t e, uint32_t k) {
    int rc = gfs_sgd_session_run(r->s, e, e + 1, k, r->syncs);
    if (rc) return rc;
    if (r->world <= 1) return GFS_OK;
    if (r->overlap) {
        rc = replica_join_reconcile(r);                                       // the previous round's corrections are in
        if (!rc) rc = gfs_p2p_region_snapshot_x(r->region, r->s->stream);
        if (rc) return rc;
        GFS_CUDA(cudaEventRecord(r->ev_snap, r->s->stream));
        GFS_CUDA(cudaStreamWaitEvent(r->side, r->ev_snap, 0));
        const bool trace = env_long("GFASORT_RC_TRACE", 0) != 0 && r->trace.size() < 4096;
        if (trace) {
            cudaEvent_t t0, t1;
            GFS_CUDA(cudaEventCreate(&t0)); GFS_CUDA(cudaEventCreate(&t1));
            r->trace.push_back(t0); r->trace.push_back(t1);
            GFS_CUDA(cudaEventRecord(t0, r->side));
        }
        rc = gfs_p2p_reconcile_async(r->region, r->side);
        if (rc) return rc;
        if (trace) GFS_CUDA(cudaEventRecord(r->trace.back(), r->side));
        GFS_CUDA(cudaEventRecord(r->ev_rc, r->side));
        r->rc_pending = true;
        if (r->join_at_once) { rc = replica_join_reconcile(r); if (rc) return rc; }
    } else {
        rc = gfs_p2p_reconcile(r->region, r->s->stream);
        if (rc) return rc;
    }
    r->reconciles += 1;
    return GFS_OK;
}
// Asynchronous: epochs [epoch_begin, epoch_end) of this rank's share, `syncs` slices per epoch, replicas reconciled
// after every slice.  Every rank must enqueue the same epochs in the same order.
extern "C" int gfs_replica_run(gfs_replica* r, uint64_t epoch_begin, uint64_t epoch_end) {
    if (!r) { set_error("gfs_replica_run: null replica"); return GFS_ERR_INVALID; }
    for (uint64_t e = epoch_begin; e < epoch_end; ++e)
        for (uint32_t k = 0; k < r->syncs; ++k) {
            const int rc = replica_run_slice(r, e, k);
            if (rc) return rc;
        }
    return GFS_OK;
}

// Asynchronous: makes the session's stream wait for the last overlapped reconcile (call it before timing events or before
// reading the replica on that stream).
extern "C" int gfs_replica_flush(gfs_replica* r) {
    if (!r) { set_error("gfs_replica_flush: null replica"); return GFS_ERR_INVALID; }
    GFS_CUDA(cudaSetDevice(r->s->device));
    return replica_join_reconcile(r);
}

// Blocking: waits for everything enqueued; a reconcile barrier that timed out on ANY rank is an error here.
extern "C" int gfs_replica_sync(gfs_replica* r) {
    if (!r) { set_error("gfs_replica_sync: null replica"); return GFS_ERR_INVALID; }
    int rc = gfs_replica_flush(r);
    if (!rc) rc = gfs_sgd_session_sync(r->s);
    if (!rc) rc = gfs_p2p_region_check(r->region);
    if (!rc && !r->trace.empty()) {             // GFASORT_RC_TRACE: barrier wait + exchange, as the side stream saw them
        double sum = 0, mx = 0;
        const size_t n = r->trace.size() / 2;
        for (size_t k = 0; k < n; ++k) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, r->trace[2 * k], r->trace[2 * k + 1]) == cudaSuccess) { sum += ms; mx = std::max<double>(mx, ms); }
        }
        std::fprintf(stderr, "[gfasort rc trace] rank %u of %u: %zu overlapped exchanges, mean %.3f ms, max %.3f ms (barrier wait included)\n",
                     r->rank, r->world, n, sum / (double)n, mx);
        for (cudaEvent_t e : r->trace) cudaEventDestroy(e);
        r->trace.clear();
    }
    return rc;
}
extern "C" int gfs_replica_download(gfs_replica* r, double* positions) {
    if (!r) { set_error("gfs_replica_download: null replica"); return GFS_ERR_INVALID; }
    int rc = gfs_replica_sync(r);
    if (!rc) rc = gfs_sgd_session_download(r->s, positions);
    return rc;
}
extern "C" int gfs_replica_stats(gfs_replica* r, gfs_stats* st) {
    if (!r || !st) { set_error("gfs_replica_stats: null argument"); return GFS_ERR_INVALID; }
    int rc = gfs_replica_sync(r);
    if (!rc) rc = gfs_sgd_session_stats(r->s, st);
    if (!rc) { st->launches += r->reconciles; st->n_devices = r->world; st->syncs_per_epoch = r->syncs; }
    return rc;
}
extern "C" int gfs_replica_stream(gfs_replica* r, void** stream, void** dev_positions, uint64_t* n_elems, uint32_t* elem_bytes) {
    if (!r) { set_error("gfs_replica_stream: null replica"); return GFS_ERR_INVALID; }
    if (stream) *stream = r->s->stream;
    return gfs_sgd_session_positions(r->s, dev_positions, n_elems, elem_bytes);
}

// ---------------------------------------------------------------------------------------------
// all ranks from one process
// ---------------------------------------------------------------------------------------------
int gfs_multi_run_whole(const gfs_index* ix, const gfs_sgd_params* params, const gfs_launch_cfg* cfg, uint32_t dims,
                        double* pos_inout, gfs_stats* stats) {
    const double t0 = now_s();
    const uint32_t G = (uint32_t)ix->shards.size();
    if (!params) { set_error("params is null"); return GFS_ERR_INVALID; }
    if (!ix->any_multi_step) { set_error("no paths with multiple steps found"); return GFS_ERR_NO_VALID_PATH; }
    const uint32_t syncs = (uint32_t)std::max<long>(0, env_long("GFASORT_SYNCS", 0));       // 0 = gfs_default_syncs_per_epoch
    std::vector<gfs_replica*> reps(G, nullptr);
    auto cleanup = [&]() { for (gfs_replica* r : reps) gfs_replica_destroy(r); };
    // per-device setup runs on one host thread per device (allocations, table uploads, the 80 MB position upload):
    // done one after the other it costs more than the whole schedule at 8 GPUs
    std::vector<int> rcs(G, GFS_OK);
    std::vector<std::string> errs(G);
    auto on_all = [&](auto fn) {
        std::vector<std::thread> pool;
        for (uint32_t g = 0; g < G; ++g)
            pool.emplace_back([&, g] {
                rcs[g] = fn(g);
                if (rcs[g]) errs[g] = gfs_last_error();
            });
        for (auto& th : pool) th.join();
        for (uint32_t g = 0; g < G; ++g)
            if (rcs[g]) { set_error("device " + std::to_string(g) + ": " + errs[g]); return rcs[g]; }
        return (int)GFS_OK;
    };
    int rc = on_all([&](uint32_t g) { return gfs_replica_create(ix->shards[g], params, dims, cfg, &ix->plans[g], ix->S, g, G, syncs, &reps[g]); });
    if (!rc) rc = gfs_replica_connect_local(reps.data(), G);
    if (!rc) rc = on_all([&](uint32_t g) { return gfs_replica_upload(reps[g], pos_inout); });
    // one SLICE at a time over all devices.  Not an epoch at a time: a device's reconcile waits (on the device) for its
    // peers' reconcile of the same slice, and the host blocks once a session's ring of 16 timing events is full — with more
    // than 16 slices per epoch the one host thread would wait for device 0 while device 1 had not been given its work yet
    // (found with GFASORT_SYNCS=40: the barrier's bounded spin turned it into an error, as designed, not into a hang)
    const uint64_t n_epochs = params->iter_max + 1;
    const uint32_t n_slices = reps[0] ? reps[0]->syncs : 1;
    for (uint64_t e = 0; e < n_epochs && !rc; ++e)
        for (uint32_t k = 0; k < n_slices && !rc; ++k)
            for (uint32_t g = 0; g < G && !rc; ++g) rc = replica_run_slice(reps[g], e, k);
    for (uint32_t g = 0; g < G && !rc; ++g) rc = gfs_replica_sync(reps[g]);
    if (!rc) rc = gfs_replica_download(reps[0], pos_inout);        // the replicas are identical after the last reconcile
    gfs_stats tot{};
    for (uint32_t g = 0; g < G && !rc; ++g) {
        gfs_stats st{};
        rc = gfs_replica_stats(reps[g], &st);
        if (rc) break;
        tot.applied_updates += st.applied_updates; tot.attempts += st.attempts; tot.launches += st.launches;
        tot.kernel_seconds = std::max(tot.kernel_seconds, st.kernel_seconds);
        tot.h2d_seconds += st.h2d_seconds; tot.d2h_seconds += st.d2h_seconds;
        tot.epochs = st.epochs; tot.grid = st.grid; tot.block = st.block; tot.coord_bytes = st.coord_bytes;
        tot.window_steps = st.window_steps; tot.coherent = st.coherent;
    }
    tot.n_devices = G; tot.syncs_per_epoch = reps[0] ? reps[0]->syncs : syncs;
    tot.total_seconds = now_s() - t0;
    if (!rc && stats) *stats = tot;
    cleanup();
    return rc;
}
