// Lookahead session: host-side control flow around the GPU cost engine.
//
// The frame-type decision logic of [x264] encoder/slicetype.c (x264_slicetype_decide,
// x264_slicetype_analyse, scenecut, slicetype_path(_cost), macroblock_tree) and the queueing
// of [x264] encoder/lookahead.c are scalar, branchy and tiny, so they stay on the host -- but
// every slicetype_frame_cost(p0,p1,b) they ask for is a handful of kernel launches on this
// session's stream.  Results are memoised per frame exactly like upstream (MVs per
// (list,distance), costs per (b-p0,p1-b)), and the evaluation ORDER is kept identical because
// it is observable (weights are analysed only on the first P evaluation of a pair; the bidir
// direct-MV seed exists only once the P search of the later reference has run).
//
// Only values the host must branch on force a stream synchronisation (scenecut tests, path
// cost comparisons, weight scores).  The mb-tree walk and the rate-control precompute only
// enqueue work; their scalar results are collected at the next synchronisation.
#include "common.cuh"
#include "csp_kernels.h"
#include "la_kernels.h"
#include "frontend.h"
#include "rgb_math.cuh"
#include "../../include/x264vfw_cuda.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <chrono>
#include <deque>
#include <string>
#include <thread>
#include <sched.h>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <vector>
#include <array>

namespace xv {

const RgbCoef &rgb_coef(int colmatrix, int fullrange);
int convert_device_public(cudaStream_t st, int out_csp, int colmatrix, int fullrange, int ext,
                          const x264vfw_cuda_image_t *dst, const x264vfw_cuda_image_t *src,
                          int w, int h, size_t sfb, size_t dfb, int n_frames);

#define BMAX X264VFW_CUDA_BFRAME_MAX
#define LMAX X264VFW_CUDA_LOOKAHEAD_MAX
#define T_AUTO X264VFW_CUDA_TYPE_AUTO
#define T_IDR X264VFW_CUDA_TYPE_IDR
#define T_I X264VFW_CUDA_TYPE_I
#define T_P X264VFW_CUDA_TYPE_P
#define T_BREF X264VFW_CUDA_TYPE_BREF
#define T_B X264VFW_CUDA_TYPE_B
#define T_KEYFRAME 6
#define IS_B(t) ((t) == T_B || (t) == T_BREF)
#define IS_I(t) ((t) == T_I || (t) == T_IDR || (t) == T_KEYFRAME)
#define AUTO_OR_I(t) ((t) == T_AUTO || IS_I(t))
#define AUTO_OR_B(t) ((t) == T_AUTO || IS_B(t))
#define COST_MAX64 (1ULL << 60)
#define WEIGHTP_FAKE (-1)      // X264_WEIGHTP_FAKE of x264.h
#define PENDING (-2)

#define ME_SIDE 2
#define ME_EVENTS 64
#define TREE_RING 4
#define TREE_MAX_STEPS (2 * LMAX + 16)

struct Frame {
    // host mirror of the x264_frame_t fields the decision logic reads
    int i_frame = 0, i_type = T_AUTO, i_forced_type = T_AUTO, b_scenecut = 1, b_keyframe = 0, i_bframes = 0;
    float f_duration = 0;
    int cost_est[BMAX + 2][BMAX + 2], cost_est_aq[BMAX + 2][BMAX + 2], intra_mbs[BMAX + 2];
    bool searched[2][BMAX + 1];            // logical: what upstream's 0x7FFF sentinel would say
    bool spec[2][BMAX + 1];                // device arrays already hold the (unweighted) search result
    int spec_eng[2][BMAX + 1];             // ... produced by this ME engine
    uint64_t spec_seq[2][BMAX + 1];        // ... in this launch of it
    uint64_t touch_seq[1 + ME_SIDE] = {0}; // last launch of each side engine that reads/writes this frame
    bool b_intra_calculated = false;
    bool intra_dev = false;                    // intra_cost[] already computed on the device (at arrival)
    bool stats_ready = false;
    unsigned long long pixel_sum[3], pixel_ssd[3];
    WeightDev weight = {0, 1, 0, 0};
    float weighted_cost_delta[BMAX + 2];
    int rc_d0 = -1, rc_d1 = -1;
    bool in_use = false;
    bool ready = false;                        // lowres planes / AQ arrays have been enqueued
    // device side
    uint8_t *lowres = nullptr;                 // 4 padded planes (+ slack)
    uint16_t *intra_cost = nullptr, *inv_qscale = nullptr;
    int *propagate = nullptr;
    float *qp_offset = nullptr, *qp_offset_aq = nullptr;
    int *mvs[2][BMAX + 1], *mv_costs[2][BMAX + 1];
    uint16_t *lowres_costs = nullptr;          // [(B+2)*(B+2)][mb]
    int *row_satds = nullptr;                  // [(B+2)*(B+2)][mb_h]  ([x264] i_row_satds)
    unsigned long long *stats = nullptr;       // device [6]
    unsigned long long *h_stats = nullptr;     // pinned [6]
    cudaEvent_t ev_stats = nullptr;            // the async copy of the statistics has landed
    uint8_t *arena = nullptr;
};

struct PendingResult { Frame *f; int d0, d1; int slot; bool is_b; bool intra; };

// Optional per-kernel-class device timing: CUDA events recorded on the session's stream around
// each launch, resolved at the next synchronisation (bench.py reads the totals).
enum KClass { K_CSP, K_AQ, K_LOWRES, K_INTRA, K_ME, K_FINALIZE, K_WEIGHT, K_TREE, K_ME_PASS, K_FRONTEND, K_N };
struct ProfRec { int cls; cudaEvent_t a, b; };
struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> pool;
    std::vector<ProfRec> recs;
    double ms[K_N] = {0};
    uint64_t n[K_N] = {0};
};

struct Decision {
    x264vfw_cuda_la_decision d;
    Frame *f;
    float *h_qp, *h_qp_aq;                     // pinned staging, valid after the decide's sync
};

struct La {
    x264vfw_cuda_la_params p;
    LaGeom g;
    int device;
    cudaStream_t st;
    int in_csp, out_csp, colmatrix, fullrange, keep_frames;
    int speculate = 1;
    int ext = X264VFW_CUDA_EXT_NONE;   // packed 4:2:2 -> I444 uses the documented extension conversion
    int me_rows = 0;         // warps per search in the wavefront kernel
    int me_variant = 1;      // 0: plain wavefront, 1: speculative parallel passes + verification wavefront
    int me_passes = 3;       // parallel (Jacobi) passes of the speculative search: pass 0 = every MB from the guessed field
    int me_relax = 0;        // row-relaxation passes after them (X264VFW_CUDA_ME_RELAX; measured slower than Jacobi passes: profiles/README.md)
    int me_force_miss = 0;   // diagnostics (X264VFW_CUDA_ME_FORCE_MISS): the verification keeps nothing = cost at a 0 % hit rate
    uint64_t n_tree_steps = 0, n_tree_walks = 0;
    int *d_me_stats = nullptr;
    int stats_verbose = 0; std::string dbg_jobs; int dbg_prev[8] = {0};
    // searches the decision logic asked for during the current decision, and the ones predicted
    // for the next decision (same pattern, shifted by the mini-GOP just emitted): {frame, list, dist}
    std::vector<std::array<int, 3>> asked_now, wanted;
    int predict = 1;
    double spec_threshold = 0.7;   // speculate a (list,distance) pair when at least this share of frames asked for it
    int decide_lag = 1;      // run the decision due at put(n) during put(n+lag): same decisions, searches overlap
    bool flushing = false;
    int la_me_hex, la_subpel_refine, la_satd, do_edges;
    int slicetype_length, i_last_keyframe;
    // tables
    uint16_t *d_cost_mv = nullptr; int cost_mv_half = 0;
    float *d_log2_lut = nullptr; uint8_t *d_exp2_lut = nullptr;
    // scratch
    uint8_t *d_src = nullptr; size_t d_src_bytes = 0;      // packed input staging
    uint8_t *d_planes = nullptr; size_t d_planes_bytes = 0; // converted planes (tight)
    x264vfw_cuda_image_t planes_img;
    uint8_t *d_weight_buf = nullptr;
    // ME engines: [0] runs on the main stream (on-demand searches), [1..ME_SIDE] on side streams
    // (speculative searches), each with its own inter-row record / ticket scratch.
    int2 *d_rec[1 + ME_SIDE] = {nullptr}; int *d_ticket[1 + ME_SIDE] = {nullptr}; int4 *d_assumed[1 + ME_SIDE] = {nullptr};
    cudaStream_t st_me[1 + ME_SIDE] = {nullptr};
    cudaEvent_t ev_me[1 + ME_SIDE][ME_EVENTS];              // done-events of the side launches (ring)
    uint64_t me_seq[1 + ME_SIDE] = {0};                     // launches issued per engine
    uint64_t me_waited[1 + ME_SIDE] = {0};                  // highest launch the main stream already waits on
    cudaEvent_t ev_ready = nullptr;                         // main stream -> side stream hand-off
    cudaEvent_t ev_io = nullptr;                            // caller's buffers are free again (H2D / D2H of this put done)
    cudaEvent_t ev_h2d = nullptr, ev_csp = nullptr, ev_planes_free = nullptr;
    cudaEvent_t io_ev[5] = {nullptr}; double io_ms[4] = {0}; uint64_t io_n = 0;   // diagnostics (X264VFW_CUDA_STATS)
    bool planes_busy = false;                               // ev_planes_free has been recorded at least once
    cudaStream_t st_io = nullptr;                           // host <-> device copies of the borrowed buffers
    int me_guess = 1;
    int me_epoch = 0, me_rr = 0, me_side = ME_SIDE;   // side engines in use (1..ME_SIDE)
    int *d_results = nullptr; int *h_results = nullptr;     // ring of 4-int slots
    int result_head = 0;
    unsigned *d_wscore = nullptr; unsigned *h_wscore = nullptr;
    // mb-tree walks: the step list of a walk is built on the host, copied once and executed by
    // ONE cluster kernel (ring of staging buffers; a slot is reused only after its copy landed)
    TreeStep *h_tree[TREE_RING] = {nullptr}, *d_tree[TREE_RING] = {nullptr};
    cudaEvent_t ev_tree[TREE_RING] = {nullptr};
    int tree_head = 0, tree_chain = 1;
    std::vector<TreeStep> tree_zero, tree_steps;
    bool tree_building = false;
    std::vector<PendingResult> pending;
    // frames
    std::vector<Frame *> all;            // by display index when keep_frames, else pool
    std::vector<Frame *> pool;
    std::vector<Frame *> by_index;       // display index -> frame (null when recycled)
    std::deque<Frame *> next;
    Frame *last_nonb = nullptr;
    std::deque<Decision> outq;           // produced by the decision logic (worker side)
    std::deque<Decision> pubq;           // handed over to the caller (x264vfw_cuda_la_get_decision)
    std::vector<float *> qp_free; std::mutex qp_mu;
    // The decision that becomes due at put(n) runs on a per-session worker thread and is collected
    // at put(n+1): the caller's wait for its own H2D / conversion / D2H (frames on which no
    // decision is due would expose it) then overlaps the decision of the previous mini-GOP.
    // Same decisions, one put later -- deterministic, like decide_lag.
    std::thread worker; std::mutex mu; std::condition_variable cv;
    bool wstop = false;
    std::deque<std::array<long, 3>> jobs;       // {frame number, planes ring slot, fused front end ran} submitted, not yet taken
    long n_put = 0;                             // frames submitted by the caller
    long processed = 0;                         // frames the worker is done with (incl. the decisions they made due)
    std::deque<std::pair<long, Decision>> doneq; // decisions tagged with the frame whose arrival produced them
    int werr = 0; std::string werr_msg;
    int async = 1;
    // host waits: spin (lowest latency; one busy core per waiting thread) or block on an interrupt
    // (X264VFW_CUDA_SYNC=block: for hosts with fewer cores than session threads)
    bool blocking = false, yielding = false, notify = false; int spin_us = 60; cudaEvent_t ev_sync = nullptr;   // default: spin; hybrid: spin for spin_us, then block
    int io_depth = 2;                           // planes ring: frames the caller may run ahead of the worker
    uint8_t *d_planes_ring[4] = {nullptr}; x264vfw_cuda_image_t planes_ring[4];
    cudaEvent_t ev_csp_ring[4] = {nullptr}, ev_free_ring[4] = {nullptr};
    // Fused front end (frontend_kernel.cu: csp + AQ + lowres in one pass over the packed rows).  It runs on the I/O
    // stream inside put_frame and writes straight into the frame's arrays, so the frame slot of put n is chosen
    // ahead of time by the worker (when it is done with frame n - io_depth) and handed over through ring_frame[];
    // ev_frame_ready orders the kernel after the slot's reset on the main stream.
    int fused = 1;
    int caller_block = -1;   // the caller sleeps on the interrupt while its copies run: -1 = in yield mode only, 0 never, 1 always
                             // (X264VFW_CUDA_CALLER_BLOCK; for nodes where the workers fit the cores but workers + callers do not)
    Frame *ring_frame[4] = {nullptr};
    cudaEvent_t ev_frame_ready[4] = {nullptr};
    std::mutex prof_mu;
    int n_input = 0;
    uint64_t n_frame_cost = 0, n_mb_search = 0, n_sync = 0;
    std::atomic<uint64_t> n_launch{0};
    bool fail = false;   // set when a device call fails inside the value-returning helpers
    double t_put = 0, t_decide = 0, t_sync = 0, t_io = 0, t_csp_wait = 0;
    // what the host was waiting for at each synchronisation (diagnostics): [0] only light kernels,
    // [1] a speculative search batch still in flight, [2] an on-demand search batch
    int sync_kind = 0; double t_sync_kind[3] = {0, 0, 0}; uint64_t n_sync_kind[3] = {0, 0, 0};   // host wall-clock seconds (diagnostics)
    uint64_t n_ondemand = 0, n_ondemand_jobs = 0, n_spec_jobs = 0;
    uint64_t n_logical[2][BMAX + 1] = {{0}};      // searches upstream's control flow actually asked for, by list/distance
    uint64_t n_launched[2][BMAX + 1] = {{0}};     // searches launched (speculative + on demand), by list/distance
    Prof prof;
};

static const int RESULT_SLOTS = 1024;

#define LA_CUDA(expr) XV_CUDA_OK(expr)

static int la_sync(La *la);
static int wait_engine(La *la, int eng, uint64_t seq);

static inline double now_s();

// ---- one polling thread per process (X264VFW_CUDA_SYNC=notify) -------------------------------------------------
// On a node with fewer cores than session threads every polling waiter steals the core of a thread that has work to
// do.  In this mode a waiter registers its event and SLEEPS on a condition variable; a single notifier thread polls
// all registered events and wakes the owners (a few microseconds of wake-up latency instead of a scheduler
// quantum).  At 8 ranks x 8 streams that leaves 8 pollers instead of 128.
struct Notifier {
    struct Item { cudaEvent_t ev; int device; std::condition_variable *cv; bool *done; cudaError_t *err; };
    std::mutex mu;
    std::condition_variable cv_work;
    std::vector<Item> items;
    bool started = false;
    void run()
    {
        int cur = -1;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu);
            cv_work.wait(lk, [&] { return !items.empty(); });
            for (size_t i = 0; i < items.size();) {
                Item &it = items[i];
                if (it.device != cur) { cudaSetDevice(it.device); cur = it.device; }
                const cudaError_t q = cudaEventQuery(it.ev);
                if (q == cudaErrorNotReady) { i++; continue; }
                *it.err = q; *it.done = true;
                it.cv->notify_one();
                items[i] = items.back(); items.pop_back();
            }
            lk.unlock();
            sched_yield();
        }
    }
    int wait(int device, cudaEvent_t ev)
    {
        std::condition_variable cv;
        bool done = false;
        cudaError_t err = cudaSuccess;
        std::unique_lock<std::mutex> lk(mu);
        if (!started) { started = true; std::thread([this] { run(); }).detach(); }
        items.push_back(Item{ev, device, &cv, &done, &err});
        cv_work.notify_one();
        cv.wait(lk, [&] { return done; });
        if (err != cudaSuccess) { set_error("cudaEventQuery failed: %s", cudaGetErrorString(err)); return -1; }
        return 0;
    }
};
// heap-allocated and never destroyed: its thread is detached and may be waiting on the condition variable at exit
static Notifier &g_notifier = *new Notifier();

// Host wait for an event: poll for a short while (most waits are for a few light kernels), then
// block on the interrupt so that long waits (a search batch, a frame's copies) do not hold a core
// -- a node runs 2 threads per stream, usually more than it has cores once several GPUs are used.
static int wait_event(La *la, cudaEvent_t ev)
{
    if (la->notify) {
        if (cudaEventQuery(ev) == cudaSuccess) return 0;       // already there: no round trip through the notifier
        return g_notifier.wait(la->device, ev);
    }
    if (la->yielding) {
        // poll, but give the core away between polls: for nodes with more session threads than cores
        for (;;) {
            const cudaError_t q = cudaEventQuery(ev);
            if (q == cudaSuccess) return 0;
            if (q != cudaErrorNotReady) { set_error("cudaEventQuery failed: %s", cudaGetErrorString(q)); return -1; }
            sched_yield();
        }
    }
    if (la->blocking && la->spin_us > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
            const cudaError_t q = cudaEventQuery(ev);
            if (q == cudaSuccess) return 0;
            if (q != cudaErrorNotReady) { set_error("cudaEventQuery failed: %s", cudaGetErrorString(q)); return -1; }
            if (std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() > la->spin_us) break;
        }
    }
    LA_CUDA(cudaEventSynchronize(ev));
    return 0;
}

// The CALLER's wait for its borrowed buffers (a whole H2D -> conversion -> D2H chain, ~1 ms with several streams on
// the link): on a node with fewer cores than session threads (yield mode) it sleeps on the interrupt instead of
// polling -- the session worker, whose waits are short and latency-critical, keeps the core.
static int wait_event_caller(La *la, cudaEvent_t ev)
{
    if (la->caller_block > 0 || (la->caller_block < 0 && la->yielding)) { LA_CUDA(cudaEventSynchronize(ev)); return 0; }
    // notify mode: wait_event already sleeps
    return wait_event(la, ev);
}

static cudaEvent_t prof_event(La *la)
{
    if (!la->prof.pool.empty()) { cudaEvent_t e = la->prof.pool.back(); la->prof.pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {
    La *la; int cls; cudaStream_t st; cudaEvent_t a = nullptr;
    ProfScope(La *l, int c, cudaStream_t s = nullptr) : la(l), cls(c), st(s ? s : l->st) { if (la->prof.on && c >= 0) { std::lock_guard<std::mutex> lk(la->prof_mu); a = prof_event(la); cudaEventRecord(a, st); } }
    ~ProfScope() { if (a) { std::lock_guard<std::mutex> lk(la->prof_mu); cudaEvent_t b = prof_event(la); cudaEventRecord(b, st); la->prof.recs.push_back(ProfRec{cls, a, b}); } }
};
static void prof_resolve(La *la)
{
    std::lock_guard<std::mutex> lk(la->prof_mu);
    // records of side-stream launches may still be in flight: keep those for the next round
    std::vector<ProfRec> keep;
    for (const ProfRec &r : la->prof.recs) {
        if (cudaEventQuery(r.b) == cudaErrorNotReady) { keep.push_back(r); continue; }
        float ms = 0;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { la->prof.ms[r.cls] += ms; la->prof.n[r.cls]++; }
        la->prof.pool.push_back(r.a); la->prof.pool.push_back(r.b);
    }
    la->prof.recs.swap(keep);
}

// ------------------------------------------------------------------------------------------
// frame slots
// ------------------------------------------------------------------------------------------
static Frame *frame_alloc(La *la)
{
    Frame *f = new Frame();
    const int n = la->g.mb_count, B = la->p.bframes;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_lowres = take((size_t)4 * la->g.lplane + 64);
    const size_t o_intra = take(n * 2), o_invq = take(n * 2), o_prop = take(n * 4), o_qp = take(n * 4), o_qpaq = take(n * 4);
    size_t o_mvs[2][BMAX + 1], o_mvc[2][BMAX + 1];
    for (int l = 0; l < 2; l++) for (int d = 0; d <= B; d++) { o_mvs[l][d] = take(n * 4); o_mvc[l][d] = take(n * 4); }
    const size_t o_lc = take((size_t)(B + 2) * (B + 2) * n * 2);
    const size_t o_rs = take((size_t)(B + 2) * (B + 2) * la->g.mb_h * 4);
    const size_t o_stats = take(64);
    if (cudaMalloc((void **)&f->arena, off) != cudaSuccess) { set_error("cudaMalloc(%zu) failed for a lookahead frame", off); delete f; return nullptr; }
    if (cudaMallocHost((void **)&f->h_stats, 64) != cudaSuccess || cudaEventCreateWithFlags(&f->ev_stats, cudaEventDisableTiming | (la->blocking ? cudaEventBlockingSync : 0)) != cudaSuccess) { set_error("cudaMallocHost failed"); cudaFree(f->arena); delete f; return nullptr; }
    f->lowres = f->arena + o_lowres;
    f->intra_cost = (uint16_t *)(f->arena + o_intra); f->inv_qscale = (uint16_t *)(f->arena + o_invq);
    f->propagate = (int *)(f->arena + o_prop);
    f->qp_offset = (float *)(f->arena + o_qp); f->qp_offset_aq = (float *)(f->arena + o_qpaq);
    for (int l = 0; l < 2; l++) for (int d = 0; d <= B; d++) { f->mvs[l][d] = (int *)(f->arena + o_mvs[l][d]); f->mv_costs[l][d] = (int *)(f->arena + o_mvc[l][d]); }
    f->lowres_costs = (uint16_t *)(f->arena + o_lc);
    f->row_satds = (int *)(f->arena + o_rs);
    f->stats = (unsigned long long *)(f->arena + o_stats);
    return f;
}

static void frame_free(Frame *f)
{
    if (!f) return;
    cudaFree(f->arena);
    cudaFreeHost(f->h_stats);
    if (f->ev_stats) cudaEventDestroy(f->ev_stats);
    delete f;
}

static int frame_reset(La *la, Frame *f, int i_frame)
{
    const int n = la->g.mb_count, B = la->p.bframes;
    f->i_frame = i_frame; f->i_type = f->i_forced_type = T_AUTO; f->b_scenecut = 1; f->b_keyframe = 0; f->i_bframes = 0;
    f->f_duration = 0;
    for (int a = 0; a < BMAX + 2; a++) { f->intra_mbs[a] = 0; f->weighted_cost_delta[a] = 0; for (int b = 0; b < BMAX + 2; b++) f->cost_est[a][b] = f->cost_est_aq[a][b] = -1; }
    memset(f->searched, 0, sizeof(f->searched));
    memset(f->spec, 0, sizeof(f->spec));
    memset(f->spec_eng, 0, sizeof(f->spec_eng));
    memset(f->spec_seq, 0, sizeof(f->spec_seq));
    f->b_intra_calculated = false; f->intra_dev = false; f->stats_ready = false;
    f->weight = WeightDev{0, 1, 0, 0};
    f->rc_d0 = f->rc_d1 = -1;
    f->in_use = true;
    f->ready = false;
    // a recycled slot may still be in use by speculative searches on the side streams
    for (int e = 1; e <= ME_SIDE; e++) {
        if (f->touch_seq[e] && wait_engine(la, e, f->touch_seq[e]) < 0) return -1;
        f->touch_seq[e] = 0;
    }
    // zero the memo arrays the kernels may read before writing (MV predictors of unscanned
    // MBs, intra cost of unscanned edge MBs), and the stats accumulators
    const size_t zero_from = (uint8_t *)f->intra_cost - f->arena;
    const size_t zero_to = ((uint8_t *)f->stats - f->arena) + 64;
    (void)n; (void)B;
    LA_CUDA(cudaMemsetAsync(f->arena + zero_from, 0, zero_to - zero_from, la->st));
    return 0;
}

static Frame *frame_get(La *la, int i_frame)
{
    Frame *f = nullptr;
    if (!la->keep_frames) {
        for (Frame *c : la->pool) if (!c->in_use) { f = c; break; }
    }
    if (!f) {
        f = frame_alloc(la);
        if (!f) return nullptr;
        la->pool.push_back(f);
    }
    if (frame_reset(la, f, i_frame) < 0) return nullptr;
    if ((int)la->by_index.size() <= i_frame) la->by_index.resize(i_frame + 1, nullptr);
    la->by_index[i_frame] = f;
    return f;
}

static void frame_release(La *la, Frame *f)
{
    if (la->keep_frames || !f) return;
    f->in_use = false;
    if (f->i_frame < (int)la->by_index.size() && la->by_index[f->i_frame] == f) la->by_index[f->i_frame] = nullptr;
}

static inline const uint8_t *plane_org(const La *la, const Frame *f, int k) { return f->lowres + (size_t)k * la->g.lplane + la->g.lorigin; }
static inline int *rs_ptr(const La *la, const Frame *f, int d0, int d1) { return f->row_satds + ((size_t)d0 * (la->p.bframes + 2) + d1) * la->g.mb_h; }
static inline uint16_t *lc_ptr(const La *la, const Frame *f, int d0, int d1) { return f->lowres_costs + ((size_t)d0 * (la->p.bframes + 2) + d1) * la->g.mb_count; }

// ------------------------------------------------------------------------------------------
// results ring + synchronisation
// ------------------------------------------------------------------------------------------
static int result_slot(La *la)
{
    if ((int)la->pending.size() >= RESULT_SLOTS - 2) { if (la_sync(la) < 0) return -1; }
    int s = la->result_head;
    la->result_head = (la->result_head + 1) % RESULT_SLOTS;
    return s;
}

static inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int la_sync(La *la)
{
    const double t0 = now_s();
    struct Acc { La *l; double t; ~Acc() { l->t_sync += now_s() - t; } } acc{la, t0};
    if (!la->pending.empty())
        LA_CUDA(cudaMemcpyAsync(la->h_results, la->d_results, RESULT_SLOTS * 4 * sizeof(int), cudaMemcpyDeviceToHost, la->st));
    if (la->blocking || la->yielding || la->notify) { LA_CUDA(cudaEventRecord(la->ev_sync, la->st)); if (wait_event(la, la->ev_sync) < 0) return -1; }
    else LA_CUDA(cudaStreamSynchronize(la->st));
    la->n_sync++;
    la->t_sync_kind[la->sync_kind] += now_s() - t0; la->n_sync_kind[la->sync_kind]++; la->sync_kind = 0;
    if (!la->prof.recs.empty()) prof_resolve(la);
    for (const PendingResult &r : la->pending) {
        const int *v = la->h_results + r.slot * 4;
        Frame *f = r.f;
        if (r.intra) {
            f->cost_est[0][0] = v[0]; f->cost_est_aq[0][0] = v[1];
        } else {
            int score = v[0];
            if (r.is_b) score = (int)((uint64_t)score * 100 / (120 + la->p.b_bias));   // [x264] slicetype_frame_cost
            else f->intra_mbs[r.d0] = v[2];
            f->cost_est[r.d0][r.d1] = score;
            f->cost_est_aq[r.d0][r.d1] = v[1];
        }
    }
    la->pending.clear();
    return 0;
}

static int ensure_stats(La *la, Frame *f)
{
    if (f->stats_ready) return 0;
    // the AQ kernel + async copy of this frame were enqueued at put time: wait for that copy only
    if (wait_event(la, f->ev_stats) < 0) return -1;
    // [x264] x264_adaptive_quant_frame: "Remove mean from SSD calculation"
    const int cf = la->p.chroma_format;
    for (int i = 0; i < 3; i++) {
        unsigned long long sum = f->h_stats[i], ssd = f->h_stats[3 + i];
        unsigned long long pw = (16 * la->g.mb_w) >> (i && cf != 3), ph = (16 * la->g.mb_h) >> (i && cf == 1);
        f->pixel_sum[i] = sum;
        f->pixel_ssd[i] = ssd - (sum * sum + pw * ph / 2) / (pw * ph);
    }
    f->stats_ready = true;
    return 0;
}

// ------------------------------------------------------------------------------------------
// [x264] slicetype_frame_cost
// ------------------------------------------------------------------------------------------
static int launch_intra_for(La *la, Frame *fenc)
{
    IntraJob ij;
    ij.plane0 = plane_org(la, fenc, 0); ij.intra_cost = fenc->intra_cost; ij.full = la->p.subme > 1; ij.satd = la->la_satd;
    if (!fenc->intra_dev) {   // normally done when the frame arrived (process_frame): the costs do not depend on when
        ProfScope ps(la, K_INTRA);
        if (launch_intra(la->st, la->g, ij, la->do_edges) < 0) return -1;
        fenc->intra_dev = true;
    }
    const int slot = result_slot(la);
    if (slot < 0) return -1;
    LA_CUDA(cudaMemsetAsync(la->d_results + slot * 4, 0, 4 * sizeof(int), la->st));
    IntraSumJob sj;
    sj.intra_cost = fenc->intra_cost; sj.inv_qscale = fenc->inv_qscale; sj.aq_on = la->p.aq_mode != 0;
    sj.result = la->d_results + slot * 4; sj.row_satd = rs_ptr(la, fenc, 0, 0);
    { ProfScope ps(la, K_INTRA); if (launch_intra_sum(la->st, la->g, sj) < 0) return -1; }
    la->n_launch += 2;
    la->pending.push_back(PendingResult{fenc, 0, 0, slot, false, true});
    fenc->cost_est[0][0] = PENDING; fenc->cost_est_aq[0][0] = PENDING;
    fenc->b_intra_calculated = true;
    return 0;
}

static int ue_size(unsigned v) { v += 1; int n = 0; while (v >> (n + 1)) n++; return 2 * n + 1; }
static int se_size(int v) { int t = 1 - v * 2; if (t < 0) t = v * 2; int n = 0; while (t >> (n + 1)) n++; return 2 * n + 1; }

static int frame_cost(La *la, Frame **frames, int p0, int p1, int b, bool need_value);

// weight_cost_luma for the unweighted reference and for one candidate weight: two kernels, one
// synchronisation (the candidate only depends on the pixel statistics, not on the first score).
static int weight_scores(La *la, Frame *fenc, Frame *ref, const WeightDev &w, unsigned *orig, unsigned *cand)
{
    LA_CUDA(cudaMemsetAsync(la->d_wscore, 0, 2 * sizeof(unsigned), la->st));
    WeightCostJob j;
    j.fenc = plane_org(la, fenc, 0); j.ref = plane_org(la, ref, 0); j.intra_cost = fenc->intra_cost;
    j.satd = la->la_satd;
    j.w = WeightDev{0, 1, 0, 0}; j.result = la->d_wscore;
    { ProfScope ps(la, K_WEIGHT); if (launch_weight_cost(la->st, la->g, j) < 0) return -1; }
    j.w = w; j.result = la->d_wscore + 1;
    { ProfScope ps(la, K_WEIGHT); if (launch_weight_cost(la->st, la->g, j) < 0) return -1; }
    la->n_launch += 2;
    LA_CUDA(cudaMemcpyAsync(la->h_wscore, la->d_wscore, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, la->st));
    if (la_sync(la) < 0) return -1;
    *orig = la->h_wscore[0];
    // + weight_slice_header_cost
    *cand = la->h_wscore[1] + 1 * 1 * (10 + ue_size(w.denom) * 2 + 2 * (se_size(w.scale) + se_size(w.offset)));
    return 0;
}

// [x264] x264_weights_analyse( h, fenc, ref, b_lookahead = 1 ): luma only, one candidate
static int weights_analyse(La *la, Frame *fenc, Frame *ref)
{
    const float epsilon = 1.f / 128.f;
    fenc->weight = WeightDev{0, 1, 0, 0};
    if (ensure_stats(la, fenc) < 0 || ensure_stats(la, ref) < 0) return -1;
    const int zero_bias = !ref->pixel_ssd[0];
    const float fenc_var = fenc->pixel_ssd[0] + zero_bias;
    const float ref_var = ref->pixel_ssd[0] + zero_bias;
    const float guess_scale = sqrtf(fenc_var / ref_var);
    const float fenc_mean = (float)(fenc->pixel_sum[0] + zero_bias) / (la->g.luma_h * la->g.luma_w) / 1;
    const float ref_mean = (float)(ref->pixel_sum[0] + zero_bias) / (la->g.luma_h * la->g.luma_w) / 1;
    if (fabsf(ref_mean - fenc_mean) < 0.5f && fabsf(1.f - guess_scale) < epsilon) return 0;

    int scale = (int)round(guess_scale * 128), denom = 7;        // x264_weight_get_h264
    while (denom > 0 && scale > 127) { denom--; scale >>= 1; }
    if (scale > 127) scale = 127;
    int found = 0, mindenom = denom, minscale = scale, minoff = 0;

    if (!fenc->b_intra_calculated) {
        // upstream nests slicetype_frame_cost( fenc, 0, 0, 0 ) here: an intra-only evaluation
        Frame *one[1] = {fenc};
        if (frame_cost(la, one, 0, 0, 0, false) < 0) return -1;
    }
    unsigned origscore, minscore, candscore;
    {
        int cur_scale = minscale;
        int cur_offset = (int)(fenc_mean - ref_mean * cur_scale / (1 << mindenom) + 0.5f * 1);
        if (cur_offset < -128 || cur_offset > 127) {
            cur_offset = cur_offset < -128 ? -128 : 127;
            float cs = (1 << mindenom) * (fenc_mean - cur_offset) / ref_mean + 0.5f;
            cur_scale = (int)(cs < 0 ? 0 : cs > 127 ? 127 : cs);
        }
        const int i_off = cur_offset < -128 ? -128 : cur_offset > 127 ? 127 : cur_offset;
        if (weight_scores(la, fenc, ref, WeightDev{1, cur_scale, mindenom, i_off}, &origscore, &candscore) < 0) return -1;
        minscore = origscore;
        if (!minscore) return 0;
        if (candscore < minscore) { minscore = candscore; minscale = cur_scale; minoff = i_off; found = 1; }
    }
    while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }
    if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f) return 0;
    fenc->weight = WeightDev{1, minscale, mindenom, minoff};
    // X264_WEIGHTP_FAKE: what the weight gained is kept for macroblock_tree_finish
    if (la->p.weightp == WEIGHTP_FAKE) fenc->weighted_cost_delta[fenc->i_frame - ref->i_frame - 1] = (float)minscore / origscore;
    // x264_weight_scale_plane: the whole padded plane 0 of the reference
    { ProfScope ps(la, K_WEIGHT); if (launch_weight_plane(la->st, la->g, la->d_weight_buf, ref->lowres, fenc->weight) < 0) return -1; }
    la->n_launch++;
    return 0;
}

static void me_params_init(La *la, MeParams &mp)
{
    memset(&mp, 0, sizeof(mp));
    mp.bands = la->p.lookahead_threads > 0 ? la->p.lookahead_threads : 1;
    mp.do_edges = la->do_edges; mp.mv_range2 = 2 * la->p.mv_range; mp.me_hex = la->la_me_hex;
    mp.subpel_refine = la->la_subpel_refine; mp.satd = la->la_satd; mp.me_range = la->p.me_range;
    mp.cost_mv = la->d_cost_mv + la->cost_mv_half;
    mp.rows_in_flight = la->me_rows;
    mp.variant = la->me_variant; mp.npasses = la->me_passes; mp.nrelax = la->me_relax; mp.stats = la->d_me_stats;
    mp.force_miss = la->me_force_miss;
}

static void me_add_job(La *la, MeParams &mp, int eng, Frame *fenc, Frame *ref, int list, int dist, const WeightDev *w)
{
    MeJob &j = mp.job[mp.njobs];
    j.fenc = plane_org(la, fenc, 0);
    for (int k = 0; k < 4; k++) j.fref[k] = plane_org(la, ref, k);
    j.fref_w = j.fref[0];
    j.w = WeightDev{0, 1, 0, 0};
    if (w) { j.w = *w; j.fref_w = la->d_weight_buf + la->g.lorigin; }
    j.mvs = fenc->mvs[list][dist - 1];
    j.mv_costs = fenc->mv_costs[list][dist - 1];
    j.rec = la->d_rec[eng] + (size_t)mp.njobs * la->g.mb_count;
    j.ticket = la->d_ticket[eng] + mp.njobs;
    j.assumed = la->d_assumed[eng] + (size_t)mp.njobs * la->g.mb_count;
    // first guess of the speculative search: the same (list, distance) field of the previous
    // frame.  Only a hint: whatever it holds (even a field still being written) is recorded as
    // "assumed" and checked against the final MVs, so it needs no ordering.
    j.guess = nullptr; j.guess_num = j.guess_den = 1;
    if (la->me_guess) {
        const int pi = fenc->i_frame - 1;
        Frame *pf = (pi >= 0 && pi < (int)la->by_index.size()) ? la->by_index[pi] : nullptr;
        if (pf == fenc) pf = nullptr;
        auto have = [](const Frame *f, int l, int d) { return f->spec[l][d - 1] || f->searched[l][d - 1]; };
        const int dmax = list ? la->p.bframes : la->p.bframes + 1;
        if (w && have(fenc, list, dist)) j.guess = j.mvs;                      // weighted re-search: start from the unweighted result (same array)
        else if (pf && have(pf, list, dist)) j.guess = pf->mvs[list][dist - 1];
        else {
            // this frame's (or the previous frame's) field of another distance, rescaled
            for (int k = 1; k <= dmax && !j.guess; k++)
                for (int sgn = -1; sgn <= 1 && !j.guess; sgn += 2) {
                    const int d2 = dist + sgn * k;
                    if (d2 < 1 || d2 > dmax) continue;
                    if (have(fenc, list, d2)) { j.guess = fenc->mvs[list][d2 - 1]; j.guess_num = dist; j.guess_den = d2; }
                    else if (pf && have(pf, list, d2)) { j.guess = pf->mvs[list][d2 - 1]; j.guess_num = dist; j.guess_den = d2; }
                }
        }
    }
    mp.njobs++;
    la->n_mb_search += la->g.mb_count;
    la->n_launched[list][dist - 1]++;
    if (la->stats_verbose) { char b[96]; snprintf(b, sizeof(b), " [f%d l%d d%d%s%s]", fenc->i_frame, list, dist, w ? " W" : "", j.guess ? "" : " noguess"); la->dbg_jobs += b; }
}

// Make the main stream wait for launch `seq` of side engine `eng` (and everything before it).
static int wait_engine(La *la, int eng, uint64_t seq)
{
    if (eng == 0 || seq <= la->me_waited[eng]) return 0;
    // ring slot of an old launch may have been re-recorded by a newer one: waiting on the newer
    // one is a superset
    const uint64_t use = (la->me_seq[eng] - seq >= ME_EVENTS) ? la->me_seq[eng] : seq;
    LA_CUDA(cudaStreamWaitEvent(la->st, la->ev_me[eng][use % ME_EVENTS], 0));
    la->me_waited[eng] = use;
    if (la->d_me_stats && cudaEventQuery(la->ev_me[eng][use % ME_EVENTS]) != cudaSuccess && la->sync_kind < 1) la->sync_kind = 1;
    return 0;
}

// One batch of searches: speculative parallel passes + verification wavefront, or the plain
// wavefront (me_variant 0).
static int me_kernels(La *la, cudaStream_t st, const MeParams &mp, bool prof)
{
    if (!mp.variant) { ProfScope ps(la, prof ? K_ME : -1, st); return launch_me(st, la->g, mp); }
    {
        ProfScope ps(la, prof ? K_ME_PASS : -1, st);
        for (int pass = 0; pass < mp.npasses; pass++)
            if (launch_me_pass(st, la->g, mp, pass) < 0) return -1;
        for (int r = 0; r < mp.nrelax; r++)
            if (launch_me_verify(st, la->g, mp, mp.npasses + r) < 0) return -1;
    }
    la->n_launch += mp.npasses + mp.nrelax;
    ProfScope ps(la, prof ? K_ME : -1, st);
    return launch_me_verify(st, la->g, mp);
}

static int me_launch(La *la, MeParams &mp, int eng)
{
    if (!mp.njobs) return 0;
    cudaStream_t st = la->st_me[eng];
    LA_CUDA(cudaMemsetAsync(la->d_ticket[eng], 0, XV_ME_MAX_JOBS * sizeof(int), st));
    mp.epoch = ++la->me_epoch;
    if (la->stats_verbose && la->d_me_stats) {
        // diagnostics only: serialises the launch to attribute time and hit rates to it
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaStreamSynchronize(la->st); cudaStreamSynchronize(st);
        cudaEventRecord(a, st);
        if (me_kernels(la, st, mp, false) < 0) return -1;
        cudaEventRecord(b, st); cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        int v[8]; cudaMemcpy(v, la->d_me_stats, sizeof(v), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[me eng%d] %.3f ms kept %d miss %d pass0 %d pass1 %d pass2 %d |%s\n", eng, ms, v[0] - la->dbg_prev[0], v[1] - la->dbg_prev[1],
                v[2] - la->dbg_prev[2], v[3] - la->dbg_prev[3], v[4] - la->dbg_prev[4], la->dbg_jobs.c_str());
        memcpy(la->dbg_prev, v, sizeof(v)); la->dbg_jobs.clear();
        cudaEventDestroy(a); cudaEventDestroy(b);
    } else
    { if (me_kernels(la, st, mp, true) < 0) return -1; }
    la->n_launch++;
    if (eng) la->n_spec_jobs += mp.njobs; else { la->n_ondemand++; la->n_ondemand_jobs += mp.njobs; la->sync_kind = 2; }
    if (eng) {
        const uint64_t seq = ++la->me_seq[eng];
        LA_CUDA(cudaEventRecord(la->ev_me[eng][seq % ME_EVENTS], st));
    }
    return 0;
}

// Speculative searches at put time: frame n just arrived, so every (frame, list, distance)
// pair whose reference is n (list 1 of n-1..n-B) or whose frame is n (list 0 towards
// n-1..n-B-1) can be searched now, all in ONE launch.  Search results do not depend on when
// they are computed (each (list,distance) array only reads itself); whether and when upstream
// would have run them stays tracked by the logical `searched` flags.
static int speculate_searches(La *la, Frame *fn)
{
    if (!la->speculate) return 0;
    const int n = fn->i_frame, B = la->p.bframes;
    const int eng = 1 + (la->me_rr++ % la->me_side);
    // hand-off: the side stream may start once everything queued so far on the main stream
    // (this frame's lowres planes, the zeroing of recycled arrays) is done
    LA_CUDA(cudaEventRecord(la->ev_ready, la->st));
    LA_CUDA(cudaStreamWaitEvent(la->st_me[eng], la->ev_ready, 0));
    MeParams mp;
    me_params_init(la, mp);
    auto alive = [&](int i) -> Frame * { return (i >= 0 && i < (int)la->by_index.size()) ? la->by_index[i] : nullptr; };
    auto add = [&](Frame *fenc, Frame *ref, int list, int d) -> int {
        me_add_job(la, mp, eng, fenc, ref, list, d, nullptr);
        fenc->spec[list][d - 1] = true;
        fenc->spec_eng[list][d - 1] = eng;
        fenc->spec_seq[list][d - 1] = la->me_seq[eng] + 1;        // the launch this job will be part of
        fenc->touch_seq[eng] = ref->touch_seq[eng] = la->me_seq[eng] + 1;
        if (mp.njobs == XV_ME_MAX_JOBS) { if (me_launch(la, mp, eng) < 0) return -1; me_params_init(la, mp); }
        return 0;
    };
    // Only speculate the (list, distance) pairs the decision logic usually asks for: measured on
    // this session so far (distance 1 always).  Everything else is searched on demand; results
    // are identical either way, only the time at which they are computed changes.
    auto likely = [&](int list, int d) -> bool {
        if (la->speculate >= 2) return true;
        if (d == 1) return true;
        if (la->n_input < 24) return false;
        return (double)la->n_logical[list][d - 1] >= la->spec_threshold * (double)la->n_input;
    };
    for (int d = 1; d <= B + 1; d++) {
        Frame *ref = alive(n - d);
        if (ref && !fn->spec[0][d - 1] && likely(0, d) && add(fn, ref, 0, d) < 0) return -1;
    }
    for (int d = 1; d <= B; d++) {
        Frame *b = alive(n - d);
        if (b && !b->spec[1][d - 1] && likely(1, d) && add(b, fn, 1, d) < 0) return -1;
    }
    // predicted searches whose two frames are both on the device by now
    for (size_t k = 0; k < la->wanted.size();) {
        const int i = la->wanted[k][0], l = la->wanted[k][1], d = la->wanted[k][2];
        Frame *fe = alive(i), *rf = alive(l ? i + d : i - d);
        const bool gone = i < n - la->slicetype_length - 8 || (fe && (fe->spec[l][d - 1] || fe->searched[l][d - 1]));
        if (!gone && fe && rf && fe->ready && rf->ready) {
            if (add(fe, rf, l, d) < 0) return -1;
            la->wanted[k] = la->wanted.back(); la->wanted.pop_back();
        } else if (gone) { la->wanted[k] = la->wanted.back(); la->wanted.pop_back(); }
        else k++;
    }
    return me_launch(la, mp, eng);
}

// Enqueues everything slicetype_frame_cost(p0,p1,b) computes.  need_value: synchronise and
// return the score; otherwise return 0 and leave the result pending.
static int frame_cost(La *la, Frame **frames, int p0, int p1, int b, bool need_value)
{
    Frame *fenc = frames[b];
    const int d0 = b - p0, d1 = p1 - b;
    if (fenc->cost_est[d0][d1] >= 0) return fenc->cost_est[d0][d1];
    if (fenc->cost_est[d0][d1] == PENDING) {
        if (!need_value) return 0;
        if (la_sync(la) < 0) return -1;
        return fenc->cost_est[d0][d1];
    }
    la->n_frame_cost++;
    if (p0 == p1) {
        // intra only
        if (!fenc->b_intra_calculated) {
            if (launch_intra_for(la, fenc) < 0) return -1;
            // every MB of an intra-only evaluation counts as intra: INTRA_MBS = the scored MBs
            const bool tiny = la->g.mb_w <= 2 || la->g.mb_h <= 2;
            fenc->intra_mbs[0] = tiny ? la->g.mb_count : (la->g.mb_w - 2) * (la->g.mb_h - 2);
        } else { fenc->cost_est[0][0] = 0; fenc->cost_est_aq[0][0] = 0; }    // unreachable through the memo
        if (!need_value) return 0;
        if (la_sync(la) < 0) return -1;
        return fenc->cost_est[0][0];
    }
    Frame *fref0 = frames[p0], *fref1 = frames[p1];
    bool do_search[2];
    WeightDev w = {0, 1, 0, 0};
    do_search[0] = b != p0 && !fenc->searched[0][d0 - 1];
    do_search[1] = b != p1 && !fenc->searched[1][d1 - 1];
    if (do_search[0]) {
        if (la->p.weightp && b == p1) {
            if (weights_analyse(la, fenc, fref0) < 0) return -1;
            w = fenc->weight;
        }
        fenc->searched[0][d0 - 1] = true;
    }
    if (do_search[1]) fenc->searched[1][d1 - 1] = true;
    if (do_search[0]) { la->n_logical[0][d0 - 1]++; la->asked_now.push_back({fenc->i_frame, 0, d0}); }
    if (do_search[1]) { la->n_logical[1][d1 - 1]++; la->asked_now.push_back({fenc->i_frame, 1, d1}); }
    const int dist_scale_factor = (((b - p0) << 8) + ((p1 - p0) >> 1)) / (p1 - p0);

    if (!fenc->b_intra_calculated && launch_intra_for(la, fenc) < 0) return -1;

    // ---- searches (wavefront kernel; both lists of a B evaluation share the launch) ----
    // A list whose unweighted result was already produced speculatively at put time is not
    // searched again; a weighted P search always runs (it differs from the speculative one).
    {
        MeParams mp;
        me_params_init(la, mp);
        for (int l = 0; l < 2; l++) {
            if (!do_search[l]) continue;
            const int dist = l ? d1 : d0;
            const bool weighted = l == 0 && w.on;
            if (fenc->spec[l][dist - 1]) {
                // produced (or being produced) by a side engine: order the main stream after it
                if (wait_engine(la, fenc->spec_eng[l][dist - 1], fenc->spec_seq[l][dist - 1]) < 0) return -1;
                if (!weighted) continue;
            }
            me_add_job(la, mp, 0, fenc, l ? fref1 : fref0, l, dist, weighted ? &w : nullptr);
            if (!weighted) { fenc->spec[l][dist - 1] = true; fenc->spec_eng[l][dist - 1] = 0; }
            // the decision logic usually asks next for the neighbouring frame against the SAME
            // reference (path "B..BP" after "B..PP"): search it in the same launch
            if (la->speculate && mp.njobs < XV_ME_MAX_JOBS) {
                const int nd = dist + 1;
                const int ni = l ? fenc->i_frame - 1 : fenc->i_frame + 1;      // display index of the neighbour
                Frame *nf = (ni >= 0 && ni < (int)la->by_index.size()) ? la->by_index[ni] : nullptr;
                const int nd_max = l ? la->p.bframes : la->p.bframes + 1;
                // ... unless the decision logic (almost) never asks for that (list, distance) on this content
                const bool asked_sometimes = la->n_input < 24 || (double)la->n_logical[l][nd - 1 < BMAX ? nd - 1 : BMAX] >= 0.1 * (double)la->n_input;
                if (nf && nf->ready && nf != (l ? fref1 : fref0) && nd <= nd_max && asked_sometimes && !nf->spec[l][nd - 1] && !nf->searched[l][nd - 1]) {
                    me_add_job(la, mp, 0, nf, l ? fref1 : fref0, l, nd, nullptr);
                    nf->spec[l][nd - 1] = true; nf->spec_eng[l][nd - 1] = 0;
                }
            }
        }
        if (me_launch(la, mp, 0) < 0) return -1;
        // memoised lists are read by the selection kernel as well: same ordering requirement
        for (int l = 0; l < 2; l++) {
            const int dist = l ? d1 : d0;
            if (dist && !do_search[l] && fenc->spec[l][dist - 1])
                if (wait_engine(la, fenc->spec_eng[l][dist - 1], fenc->spec_seq[l][dist - 1]) < 0) return -1;
        }
        if (b < p1 && fref1->searched[0][p1 - p0 - 1] && fref1->spec[0][p1 - p0 - 1])
            if (wait_engine(la, fref1->spec_eng[0][p1 - p0 - 1], fref1->spec_seq[0][p1 - p0 - 1]) < 0) return -1;
    }

    // ---- per-MB selection + accumulators ----
    const int slot = result_slot(la);
    if (slot < 0) return -1;
    LA_CUDA(cudaMemsetAsync(la->d_results + slot * 4, 0, 4 * sizeof(int), la->st));
    FinalizeJob fj;
    memset(&fj, 0, sizeof(fj));
    fj.fenc = plane_org(la, fenc, 0);
    for (int k = 0; k < 4; k++) { fj.fref0[k] = plane_org(la, fref0, k); fj.fref1[k] = plane_org(la, fref1, k); }
    fj.mvs0 = b != p0 ? fenc->mvs[0][d0 - 1] : nullptr; fj.mv_costs0 = b != p0 ? fenc->mv_costs[0][d0 - 1] : nullptr;
    fj.mvs1 = b != p1 ? fenc->mvs[1][d1 - 1] : nullptr; fj.mv_costs1 = b != p1 ? fenc->mv_costs[1][d1 - 1] : nullptr;
    fj.ref1_mvs = (b < p1 && fref1->searched[0][p1 - p0 - 1]) ? fref1->mvs[0][p1 - p0 - 1] : nullptr;
    fj.intra_cost = fenc->intra_cost; fj.inv_qscale = fenc->inv_qscale;
    fj.lowres_costs = lc_ptr(la, fenc, d0, d1);
    fj.row_satd = rs_ptr(la, fenc, d0, d1); fj.result = la->d_results + slot * 4;
    fj.b_bidir = b < p1; fj.b_p = b == p1;
    fj.dist_scale_factor = dist_scale_factor;
    fj.bipred_weight = la->p.weightb ? 64 - (dist_scale_factor >> 2) : 32;
    fj.aq_on = la->p.aq_mode != 0; fj.subme_gt1 = la->p.subme > 1; fj.satd = la->la_satd;
    fj.mv_range2 = 2 * la->p.mv_range; fj.do_edges = la->do_edges;
    { ProfScope ps(la, K_FINALIZE); if (launch_finalize(la->st, la->g, fj) < 0) return -1; }
    la->n_launch++;
    la->pending.push_back(PendingResult{fenc, d0, d1, slot, b != p1, false});
    fenc->cost_est[d0][d1] = PENDING; fenc->cost_est_aq[d0][d1] = PENDING;
    if (!need_value) return 0;
    if (la_sync(la) < 0) return -1;
    return fenc->cost_est[d0][d1];
}

// ------------------------------------------------------------------------------------------
// [x264] macroblock_tree / macroblock_tree_propagate / macroblock_tree_finish
// ------------------------------------------------------------------------------------------
static inline float clip_duration(float f) { return f < 0.01f ? 0.01f : f > 1.00f ? 1.00f : f; }
#define MBTREE_PRECISION 0.5f

static int tree_finish(La *la, Frame *frame, float average_duration, int ref0_distance)
{
    TreeFinishJob j;
    j.fps_factor = (int)round(clip_duration(average_duration) / clip_duration(frame->f_duration) * 256 / MBTREE_PRECISION);
    float weightdelta = 0.0;
    if (ref0_distance && frame->weighted_cost_delta[ref0_distance - 1] > 0)
        weightdelta = (1.0 - frame->weighted_cost_delta[ref0_distance - 1]);
    j.weightdelta = weightdelta;
    j.strength = 5.0f * (1.0f - la->p.qcompress);
    j.propagate = frame->propagate; j.intra_cost = frame->intra_cost; j.inv_qscale = frame->inv_qscale;
    j.qp_offset_aq = frame->qp_offset_aq; j.qp_offset = frame->qp_offset; j.log2_lut = la->d_log2_lut;
    if (la->tree_building) {
        TreeStep s;
        memset(&s, 0, sizeof(s));
        s.op = 2; s.sync = 0;       // the finishing steps of a walk touch different frames
        s.fin_propagate = j.propagate; s.fin_intra = j.intra_cost; s.fin_invq = j.inv_qscale;
        s.fin_qp_aq = j.qp_offset_aq; s.fin_qp = j.qp_offset;
        s.fin_fps_factor = j.fps_factor; s.fin_weightdelta = j.weightdelta; s.fin_strength = j.strength;
        la->tree_steps.push_back(s);
        return 0;
    }
    la->n_launch++;
    ProfScope ps(la, K_TREE);
    return launch_tree_finish(la->st, la->g, j);
}

static int tree_zero(La *la, int *p, int n)
{
    if (la->tree_building) {
        // Every accumulator of a walk is cleared exactly once and before anything is added to it,
        // so all clears can run first, in one phase.
        TreeStep s;
        memset(&s, 0, sizeof(s));
        s.op = 0; s.n = n; s.zero = p; s.sync = 0;
        la->tree_zero.push_back(s);
        return 0;
    }
    LA_CUDA(cudaMemsetAsync(p, 0, (size_t)n * sizeof(int), la->st));
    return 0;
}

static int tree_propagate(La *la, Frame **frames, float average_duration, int p0, int p1, int b, int referenced)
{
    PropagateJob j;
    const int dist_scale_factor = (((b - p0) << 8) + ((p1 - p0) >> 1)) / (p1 - p0);
    j.bipred_weight = la->p.weightb ? 64 - (dist_scale_factor >> 2) : 32;
    j.fps_factor = clip_duration(frames[b]->f_duration) / (clip_duration(average_duration) * 256.0f) * MBTREE_PRECISION;
    // upstream zeroes one row of the unreferenced frame's own array and reads that as its input
    if (!referenced && tree_zero(la, frames[b]->propagate, la->g.mb_w) < 0) return -1;
    j.propagate_in = referenced ? frames[b]->propagate : nullptr;
    j.intra_cost = frames[b]->intra_cost; j.inv_qscale = frames[b]->inv_qscale;
    j.lowres_costs = lc_ptr(la, frames[b], b - p0, p1 - b);
    j.mvs0 = b != p0 ? frames[b]->mvs[0][b - p0 - 1] : nullptr;
    j.mvs1 = b != p1 ? frames[b]->mvs[1][p1 - b - 1] : nullptr;
    j.ref0_cost = frames[p0]->propagate; j.ref1_cost = frames[p1]->propagate;
    j.b_bidir = b != p1;
    if (la->tree_building) {
        TreeStep s;
        memset(&s, 0, sizeof(s));
        s.op = 1; s.prop = j;
        // Steps only ADD (integer atomics, order-free) into accumulators of frames earlier in
        // time, and only a referenced frame READS an accumulator (its own, complete once every
        // earlier step has landed): a barrier is needed before a referenced step, nowhere else.
        s.sync = 0;
        if (referenced && !la->tree_steps.empty()) la->tree_steps.back().sync = 1;
        la->tree_steps.push_back(s);
        return 0;
    }
    la->n_launch++;
    ProfScope ps(la, K_TREE);
    return launch_propagate(la->st, la->g, j);
}

static int zero_propagate(La *la, Frame *f)
{
    return tree_zero(la, f->propagate, la->g.mb_count);
}

// Copy the step list of the walk that was just described and run it as one cluster kernel.
static int tree_run(La *la)
{
    la->tree_building = false;
    std::vector<TreeStep> &z = la->tree_zero, &t = la->tree_steps;
    if (t.empty()) { z.clear(); return 0; }
    if (!z.empty()) z.back().sync = 1;
    // the finishing steps read accumulators the last propagate steps add to
    for (size_t k = 0; k < t.size(); k++) if (t[k].op == 2 && k > 0) { t[k - 1].sync = 1; break; }
    const size_t n = z.size() + t.size();
    if (n > TREE_MAX_STEPS) { set_error("mb-tree walk too long (%zu steps)", n); return -1; }
    const int slot = la->tree_head;
    la->tree_head = (la->tree_head + 1) % TREE_RING;
    LA_CUDA(cudaEventSynchronize(la->ev_tree[slot]));
    memcpy(la->h_tree[slot], z.data(), z.size() * sizeof(TreeStep));
    memcpy(la->h_tree[slot] + z.size(), t.data(), t.size() * sizeof(TreeStep));
    LA_CUDA(cudaMemcpyAsync(la->d_tree[slot], la->h_tree[slot], n * sizeof(TreeStep), cudaMemcpyHostToDevice, la->st));
    LA_CUDA(cudaEventRecord(la->ev_tree[slot], la->st));
    z.clear(); t.clear();
    la->n_launch++;
    la->n_tree_steps += n; la->n_tree_walks++;
    ProfScope ps(la, K_TREE);
    return launch_tree_chain(la->st, la->g, la->d_tree[slot], (int)n, la->d_log2_lut);
}

static int macroblock_tree_walk(La *la, Frame **frames, int num_frames, int b_intra);

// The frame costs a walk asks for are enqueued as they come (they never read an accumulator);
// the propagate / finish steps are collected and run afterwards in one launch.
static int macroblock_tree(La *la, Frame **frames, int num_frames, int b_intra)
{
    la->tree_building = la->tree_chain != 0;
    la->tree_zero.clear(); la->tree_steps.clear();
    const int r = macroblock_tree_walk(la, frames, num_frames, b_intra);
    if (!la->tree_building) return r;
    if (r < 0) { la->tree_building = false; return r; }
    return tree_run(la);
}

static int macroblock_tree_walk(La *la, Frame **frames, int num_frames, int b_intra)
{
    int idx = !b_intra;
    int last_nonb, cur_nonb = 1, bframes = 0;
    float total_duration = 0.0;
    for (int j = 0; j <= num_frames; j++) total_duration += frames[j]->f_duration;
    const float average_duration = total_duration / (num_frames + 1);
    int i = num_frames;

    if (b_intra && frame_cost(la, frames, 0, 0, 0, false) < 0) return -1;
    while (i > 0 && IS_B(frames[i]->i_type)) i--;
    last_nonb = i;
    if (last_nonb < idx) return 0;
    if (zero_propagate(la, frames[last_nonb]) < 0) return -1;

    while (i-- > idx) {
        cur_nonb = i;
        while (IS_B(frames[cur_nonb]->i_type) && cur_nonb > 0) cur_nonb--;
        if (cur_nonb < idx) break;
        if (frame_cost(la, frames, cur_nonb, last_nonb, last_nonb, false) < 0) return -1;
        if (zero_propagate(la, frames[cur_nonb]) < 0) return -1;
        bframes = last_nonb - cur_nonb - 1;
        if (la->p.b_pyramid && bframes > 1) {
            const int middle = (bframes + 1) / 2 + cur_nonb;
            if (frame_cost(la, frames, cur_nonb, last_nonb, middle, false) < 0) return -1;
            if (zero_propagate(la, frames[middle]) < 0) return -1;
            while (i > cur_nonb) {
                const int p0 = i > middle ? middle : cur_nonb;
                const int p1 = i < middle ? middle : last_nonb;
                if (i != middle) {
                    if (frame_cost(la, frames, p0, p1, i, false) < 0) return -1;
                    if (tree_propagate(la, frames, average_duration, p0, p1, i, 0) < 0) return -1;
                }
                i--;
            }
            if (tree_propagate(la, frames, average_duration, cur_nonb, last_nonb, middle, 1) < 0) return -1;
        } else {
            while (i > cur_nonb) {
                if (frame_cost(la, frames, cur_nonb, last_nonb, i, false) < 0) return -1;
                if (tree_propagate(la, frames, average_duration, cur_nonb, last_nonb, i, 0) < 0) return -1;
                i--;
            }
        }
        if (tree_propagate(la, frames, average_duration, cur_nonb, last_nonb, last_nonb, 1) < 0) return -1;
        last_nonb = cur_nonb;
    }
    if (tree_finish(la, frames[last_nonb], average_duration, last_nonb) < 0) return -1;
    if (la->p.b_pyramid && bframes > 1)
        if (tree_finish(la, frames[last_nonb + (bframes + 1) / 2], average_duration, 0) < 0) return -1;
    return 0;
}

// ------------------------------------------------------------------------------------------
// [x264] slicetype_path_cost / slicetype_path / scenecut / x264_slicetype_analyse
// ------------------------------------------------------------------------------------------

static uint64_t path_cost(La *la, Frame **frames, const char *path, uint64_t threshold)
{
    uint64_t cost = 0;
    int loc = 1, cur_nonb = 0;
    path--;
    if (threshold == COST_MAX64) {
        // no early termination possible: every evaluation of this path is unconditional, so
        // enqueue all of them first (same order) and pay one synchronisation instead of one each
        int l2 = 1, cn = 0;
        while (path[l2]) {
            int nn = l2;
            while (path[nn] == 'B') nn++;
            if (path[nn] == 'P') frame_cost(la, frames, cn, nn, nn, false);
            else frame_cost(la, frames, nn, nn, nn, false);
            if (la->p.b_pyramid && nn - cn > 2) {
                const int middle = cn + (nn - cn) / 2;
                frame_cost(la, frames, cn, nn, middle, false);
                for (int nb = l2; nb < middle; nb++) frame_cost(la, frames, cn, middle, nb, false);
                for (int nb = middle + 1; nb < nn; nb++) frame_cost(la, frames, middle, nn, nb, false);
            } else
                for (int nb = l2; nb < nn; nb++) frame_cost(la, frames, cn, nn, nb, false);
            l2 = nn + 1;
            cn = nn;
        }
    }
    auto fc = [&](int p0, int p1, int b) -> uint64_t { int v = frame_cost(la, frames, p0, p1, b, true); if (v < 0) { la->fail = true; return 0; } return (uint64_t)v; };
    while (path[loc]) {
        int next_nonb = loc;
        while (path[next_nonb] == 'B') next_nonb++;
        if (path[next_nonb] == 'P') cost += fc(cur_nonb, next_nonb, next_nonb);
        else cost += fc(next_nonb, next_nonb, next_nonb);
        if (cost > threshold) break;
        if (la->p.b_pyramid && next_nonb - cur_nonb > 2) {
            const int middle = cur_nonb + (next_nonb - cur_nonb) / 2;
            cost += fc(cur_nonb, next_nonb, middle);
            for (int next_b = loc; next_b < middle && cost < threshold; next_b++) cost += fc(cur_nonb, middle, next_b);
            for (int next_b = middle + 1; next_b < next_nonb && cost < threshold; next_b++) cost += fc(middle, next_nonb, next_b);
        } else
            for (int next_b = loc; next_b < next_nonb && cost < threshold; next_b++) cost += fc(cur_nonb, next_nonb, next_b);
        loc = next_nonb + 1;
        cur_nonb = next_nonb;
    }
    return cost;
}

static void slicetype_path(La *la, Frame **frames, int length, char (*best_paths)[LMAX + 1])
{
    char paths[2][LMAX + 1];
    const int num_paths = la->p.bframes + 1 < length ? la->p.bframes + 1 : length;
    uint64_t best_cost = COST_MAX64;
    int best_possible = 0, idx = 0;
    for (int path = 0; path < num_paths; path++) {
        const int len = length - (path + 1);
        memcpy(paths[idx], best_paths[len % (BMAX + 1)], len);
        memset(paths[idx] + len, 'B', path);
        strcpy(paths[idx] + len + path, "P");
        int possible = 1;
        for (int i = 1; i <= length; i++) {
            const int t = frames[i]->i_type;
            if (t == T_AUTO) continue;
            if (IS_B(t)) possible = possible && (i < len || i == length || paths[idx][i - 1] == 'B');
            else {
                possible = possible && (i < len || paths[idx][i - 1] != 'B');
                paths[idx][i - 1] = IS_I(t) ? 'I' : 'P';
            }
        }
        if (possible || !best_possible) {
            if (possible && !best_possible) best_cost = COST_MAX64;
            const uint64_t cost = path_cost(la, frames, paths[idx], best_cost);
            if (cost < best_cost) { best_cost = cost; best_possible = possible; idx ^= 1; }
        }
    }
    memcpy(best_paths[length % (BMAX + 1)], paths[idx ^ 1], length);
}

static int scenecut_internal(La *la, Frame **frames, int p0, int p1)
{
    Frame *frame = frames[p1];
    if (frame_cost(la, frames, p0, p1, p1, true) < 0) { la->fail = true; return 0; }
    if (frame->cost_est[0][0] == PENDING && la_sync(la) < 0) { la->fail = true; return 0; }
    const int icost = frame->cost_est[0][0];
    const int pcost = frame->cost_est[p1 - p0][0];
    float f_bias;
    const int i_gop_size = frame->i_frame - la->i_last_keyframe;
    const float f_thresh_max = la->p.scenecut / 100.0;
    float f_thresh_min = f_thresh_max * 0.25;
    if (la->p.keyint_min == la->p.keyint_max) f_thresh_min = f_thresh_max;
    if (i_gop_size <= la->p.keyint_min / 4) f_bias = f_thresh_min / 4;
    else if (i_gop_size <= la->p.keyint_min) f_bias = f_thresh_min * i_gop_size / la->p.keyint_min;
    else f_bias = f_thresh_min + (f_thresh_max - f_thresh_min) * (i_gop_size - la->p.keyint_min) / (la->p.keyint_max - la->p.keyint_min);
    return pcost >= (1.0 - f_bias) * icost;
}

static int scenecut(La *la, Frame **frames, int p0, int p1, int real_scenecut, int num_frames, int i_max_search)
{
    if (real_scenecut && la->p.bframes) {
        int origmaxp1 = p0 + 1;
        if (la->p.b_adapt == 2) origmaxp1 += la->p.bframes;
        else origmaxp1++;
        const int maxp1 = origmaxp1 < num_frames ? origmaxp1 : num_frames;
        // none of the evaluations below depends on another one's value: enqueue them all first
        // (same order as the loops), synchronise once
        for (int curp1 = p1; curp1 <= maxp1; curp1++) frame_cost(la, frames, p0, curp1, curp1, false);
        if (!(origmaxp1 > i_max_search))
            for (int curp0 = p0; curp0 < maxp1; curp0++) frame_cost(la, frames, curp0, maxp1, maxp1, false);
        for (int curp1 = p1; curp1 <= maxp1; curp1++)
            if (!scenecut_internal(la, frames, p0, curp1))
                for (int i = curp1; i > p0; i--) frames[i]->b_scenecut = 0;
        for (int curp0 = p0; curp0 <= maxp1; curp0++)
            if (origmaxp1 > i_max_search || (curp0 < maxp1 && scenecut_internal(la, frames, curp0, maxp1)))
                frames[curp0]->b_scenecut = 0;
    }
    if (!frames[p1]->b_scenecut) return 0;
    return scenecut_internal(la, frames, p0, p1);
}

static int slicetype_analyse(La *la, int intra_minigop)
{
    const x264vfw_cuda_la_params *p = &la->p;
    Frame *frames[LMAX + 3] = {nullptr};
    int num_frames, orig_num_frames, keyint_limit, framecnt;
    int i_max_search = (int)la->next.size() < LMAX ? (int)la->next.size() : LMAX;
    if (i_max_search > la->slicetype_length + 1 - intra_minigop) i_max_search = la->slicetype_length + 1 - intra_minigop;   // b_deterministic
    const int keyframe = !!intra_minigop;

    if (!la->last_nonb) return 0;
    frames[0] = la->last_nonb;
    for (framecnt = 0; framecnt < i_max_search; framecnt++) frames[framecnt + 1] = la->next[framecnt];

    if (!framecnt) {
        if (p->b_mbtree) return macroblock_tree(la, frames, 0, keyframe);
        return 0;
    }
    keyint_limit = p->keyint_max - frames[0]->i_frame + la->i_last_keyframe - 1;
    orig_num_frames = num_frames = framecnt < keyint_limit ? framecnt : keyint_limit;
    if (p->b_psy && p->b_mbtree) num_frames = framecnt;
    else if (p->open_gop && num_frames < framecnt) num_frames++;
    else if (num_frames == 0) { frames[1]->i_type = T_I; return 0; }

    if (AUTO_OR_I(frames[1]->i_type) && p->scenecut && scenecut(la, frames, 0, 1, 1, orig_num_frames, i_max_search)) {
        if (frames[1]->i_type == T_AUTO) frames[1]->i_type = T_I;
        return la->fail ? -1 : 0;
    }
    for (int j = 1; j <= num_frames; j++)
        if (frames[j]->i_type == T_KEYFRAME) frames[j]->i_type = p->open_gop ? T_I : T_IDR;
    for (int j = 2; j <= num_frames; j++)
        if (frames[j]->i_type == T_IDR && AUTO_OR_B(frames[j - 1]->i_type)) frames[j - 1]->i_type = T_P;

    int num_analysed_frames = num_frames;
    int reset_start;
    if (p->bframes) {
        if (p->b_adapt == 2) {
            if (num_frames > 1) {
                char best_paths[BMAX + 1][LMAX + 1];
                memset(best_paths, 0, sizeof(best_paths));
                strcpy(best_paths[1], "P");
                const int best_path_index = num_frames % (BMAX + 1);
                for (int j = 2; j <= num_frames; j++) slicetype_path(la, frames, j, best_paths);
                for (int j = 1; j < num_frames; j++) {
                    if (best_paths[best_path_index][j - 1] != 'B') {
                        if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = T_P;
                    } else if (frames[j]->i_type == T_AUTO) frames[j]->i_type = T_B;
                }
            }
        } else if (p->b_adapt == 1) {
            int last_nonb = 0, num_bframes = p->bframes;
            char path[LMAX + 1];
            for (int j = 1; j < num_frames; j++) {
                if (j - 1 > 0 && IS_B(frames[j - 1]->i_type)) num_bframes--;
                else { last_nonb = j - 1; num_bframes = p->bframes; }
                if (!num_bframes) {
                    if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = T_P;
                    continue;
                }
                if (frames[j]->i_type != T_AUTO) continue;
                if (IS_B(frames[j + 1]->i_type)) { frames[j]->i_type = T_P; continue; }
                const int bframes = j - last_nonb - 1;
                memset(path, 'B', bframes);
                strcpy(path + bframes, "PP");
                const uint64_t cost_p = path_cost(la, frames + last_nonb, path, COST_MAX64);
                strcpy(path + bframes, "BP");
                const uint64_t cost_b = path_cost(la, frames + last_nonb, path, cost_p);
                frames[j]->i_type = cost_b < cost_p ? T_B : T_P;
            }
        } else {
            int num_bframes = p->bframes;
            for (int j = 1; j < num_frames; j++) {
                if (!num_bframes) {
                    if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = T_P;
                } else if (frames[j]->i_type == T_AUTO) {
                    if (IS_B(frames[j + 1]->i_type)) frames[j]->i_type = T_P;
                    else frames[j]->i_type = T_B;
                }
                if (IS_B(frames[j]->i_type)) num_bframes--;
                else num_bframes = p->bframes;
            }
        }
        if (AUTO_OR_B(frames[num_frames]->i_type)) frames[num_frames]->i_type = T_P;

        int num_bframes = 0;
        while (num_bframes < num_frames && IS_B(frames[num_bframes + 1]->i_type)) num_bframes++;
        for (int j = 1; j < num_bframes + 1; j++) {
            if (frames[j]->i_forced_type == T_AUTO && AUTO_OR_I(frames[j + 1]->i_forced_type) &&
                p->scenecut && scenecut(la, frames, j, j + 1, 0, orig_num_frames, i_max_search)) {
                frames[j]->i_type = T_P;
                num_analysed_frames = j;
                break;
            }
        }
        reset_start = keyframe ? 1 : (num_bframes + 2 < num_analysed_frames + 1 ? num_bframes + 2 : num_analysed_frames + 1);
    } else {
        for (int j = 1; j <= num_frames; j++)
            if (AUTO_OR_B(frames[j]->i_type)) frames[j]->i_type = T_P;
        reset_start = !keyframe + 1;
    }
    if (la->fail) return -1;

    if (p->b_mbtree && macroblock_tree(la, frames, num_frames < p->keyint_max ? num_frames : p->keyint_max, keyframe) < 0) return -1;

    {   // enforce keyframe limit
        int last_keyframe = la->i_last_keyframe, last_possible = 0;
        for (int j = 1; j <= num_frames; j++) {
            Frame *frm = frames[j];
            int keyframe_dist = frm->i_frame - last_keyframe;
            if (AUTO_OR_I(frm->i_forced_type)) {
                if (p->open_gop || !IS_B(frames[j - 1]->i_forced_type)) last_possible = j;
            }
            if (keyframe_dist >= p->keyint_max) {
                if (last_possible != 0 && last_possible != j) {
                    j = last_possible;
                    frm = frames[j];
                    keyframe_dist = frm->i_frame - last_keyframe;
                }
                last_possible = 0;
                if (frm->i_type != T_IDR) frm->i_type = p->open_gop ? T_I : T_IDR;
            }
            if (frm->i_type == T_I && keyframe_dist >= p->keyint_min) {
                if (p->open_gop) last_keyframe = frm->i_frame;
                else if (frm->i_forced_type != T_I) frm->i_type = T_IDR;
            }
            if (frm->i_type == T_IDR) {
                last_keyframe = frm->i_frame;
                if (j > 1 && IS_B(frames[j - 1]->i_type)) frames[j - 1]->i_type = T_P;
            }
        }
    }
    for (int j = reset_start; j <= num_frames; j++) frames[j]->i_type = frames[j]->i_forced_type;
    return 0;
}

// ------------------------------------------------------------------------------------------
// [x264] x264_slicetype_decide + lookahead_slicetype_decide
// ------------------------------------------------------------------------------------------
static float *qp_staging(La *la)
{
    {
        std::lock_guard<std::mutex> lk(la->qp_mu);
        if (!la->qp_free.empty()) { float *p = la->qp_free.back(); la->qp_free.pop_back(); return p; }
    }
    float *p = nullptr;
    if (cudaMallocHost((void **)&p, (size_t)2 * la->g.mb_count * sizeof(float)) != cudaSuccess) { set_error("cudaMallocHost failed"); return nullptr; }
    return p;
}

static int decide_and_shift(La *la)
{
    const x264vfw_cuda_la_params *p = &la->p;
    Frame *frames[BMAX + 3];
    Frame *frm;
    int bframes, brefs;
    if (la->next.empty()) return 0;
    la->fail = false;
    la->asked_now.clear();

    for (Frame *f : la->next) f->f_duration = (float)((double)2 * p->fps_den / ((double)p->fps_num * 2));

    if ((p->bframes && p->b_adapt) || p->scenecut || p->b_mbtree)
        if (slicetype_analyse(la, 0) < 0) return -1;

    // frames queued beyond the decision's own window (decide_lag) are invisible to it
    int n_next = (int)la->next.size();
    if (!la->flushing && n_next > la->slicetype_length + 1) n_next = la->slicetype_length + 1;
    for (bframes = 0, brefs = 0;; bframes++) {
        frm = la->next[bframes];
        if (frm->i_type == T_BREF && p->b_pyramid < 2 && brefs == p->b_pyramid) frm->i_type = T_B;
        else if (frm->i_type == T_BREF && p->b_pyramid == 2 && brefs && p->frame_reference <= (brefs + 3)) frm->i_type = T_B;
        if (frm->i_type == T_KEYFRAME) frm->i_type = p->open_gop ? T_I : T_IDR;
        if (frm->i_frame - la->i_last_keyframe >= p->keyint_max) {
            if (frm->i_type == T_AUTO || frm->i_type == T_I)
                frm->i_type = p->open_gop && la->i_last_keyframe >= 0 ? T_I : T_IDR;
            int warn = frm->i_type != T_IDR;
            if (warn && p->open_gop) warn &= frm->i_type != T_I;
            if (warn) frm->i_type = p->open_gop && la->i_last_keyframe >= 0 ? T_I : T_IDR;
        }
        if (frm->i_type == T_I && frm->i_frame - la->i_last_keyframe >= p->keyint_min) {
            if (p->open_gop) { la->i_last_keyframe = frm->i_frame; frm->b_keyframe = 1; }
            else frm->i_type = T_IDR;
        }
        if (frm->i_type == T_IDR) {
            la->i_last_keyframe = frm->i_frame;
            frm->b_keyframe = 1;
            if (bframes > 0) { bframes--; la->next[bframes]->i_type = T_P; }
        }
        if (bframes == p->bframes || bframes + 1 >= n_next) {
            if (frm->i_type == T_AUTO || IS_B(frm->i_type)) frm->i_type = T_P;
        }
        if (frm->i_type == T_BREF) brefs++;
        if (frm->i_type == T_AUTO) frm->i_type = T_B;
        else if (!IS_B(frm->i_type)) break;
    }
    la->next[bframes]->i_bframes = bframes;
    if (p->b_pyramid && bframes > 1 && !brefs) { la->next[(bframes - 1) / 2]->i_type = T_BREF; brefs++; }

    {   // frame cost ratecontrol will ask for (x264_rc_analyse_slice)
        int p0, p1, b;
        p1 = b = bframes + 1;
        frames[0] = la->last_nonb;
        for (int i = 0; i <= bframes; i++) frames[1 + i] = la->next[i];
        if (IS_I(la->next[bframes]->i_type)) p0 = bframes + 1;
        else p0 = 0;
        if (frame_cost(la, frames, p0, p1, b, false) < 0) return -1;
        la->next[bframes]->rc_d0 = b - p0; la->next[bframes]->rc_d1 = p1 - b;
    }

    // coded order: non-B, then BREF, then B
    std::vector<Frame *> coded;
    coded.push_back(la->next[bframes]);
    for (int i = 0; i < bframes; i++) if (la->next[i]->i_type == T_BREF) coded.push_back(la->next[i]);
    for (int i = 0; i < bframes; i++) if (la->next[i]->i_type != T_BREF) coded.push_back(la->next[i]);

    Frame *old_nonb = la->last_nonb;
    la->last_nonb = la->next[bframes];
    const int shift = bframes + 1;
    for (int i = 0; i < shift; i++) la->next.pop_front();

    if (p->b_mbtree && IS_I(la->last_nonb->i_type))
        if (slicetype_analyse(la, shift) < 0) return -1;

    // stage the per-MB offsets of the shifted frames, then one synchronisation for the decide
    for (Frame *f : coded) {
        Decision d;
        memset(&d.d, 0, sizeof(d.d));
        d.f = f;
        d.h_qp = qp_staging(la);
        if (!d.h_qp) return -1;
        d.h_qp_aq = d.h_qp + la->g.mb_count;
        LA_CUDA(cudaMemcpyAsync(d.h_qp, f->qp_offset, la->g.mb_count * sizeof(float), cudaMemcpyDeviceToHost, la->st));
        LA_CUDA(cudaMemcpyAsync(d.h_qp_aq, f->qp_offset_aq, la->g.mb_count * sizeof(float), cudaMemcpyDeviceToHost, la->st));
        la->outq.push_back(d);
    }
    if (la_sync(la) < 0) return -1;
    for (size_t k = la->outq.size() - coded.size(); k < la->outq.size(); k++) {
        Decision &d = la->outq[k];
        Frame *f = d.f;
        d.d.i_frame = f->i_frame; d.d.i_type = f->i_type; d.d.b_keyframe = f->b_keyframe; d.d.i_bframes = f->i_bframes;
        d.d.mb_count = la->g.mb_count;
        d.d.i_cost_est = d.d.i_cost_est_aq = d.d.i_intra_mbs = -1;
        if (f->rc_d0 >= 0) {
            d.d.i_cost_est = f->cost_est[f->rc_d0][f->rc_d1];
            d.d.i_cost_est_aq = f->cost_est_aq[f->rc_d0][f->rc_d1];
            d.d.i_intra_mbs = f->intra_mbs[f->rc_d0];
        }
        d.f = nullptr;
    }
    // next decision: most likely the same evaluation pattern, shifted by the mini-GOP just emitted
    if (la->predict && !la->flushing)
        for (const auto &a : la->asked_now) la->wanted.push_back({a[0] + shift, a[1], a[2]});
    // recycle: B-frames are dead once shifted; the previous last_nonb is dead now
    for (Frame *f : coded) if (f != la->last_nonb) frame_release(la, f);
    if (old_nonb && old_nonb != la->last_nonb) frame_release(la, old_nonb);
    return la->fail ? -1 : 0;
}

// The fused front end for frame f: packed BGRA rows at `src` (device) -> converted planes + AQ arrays + lowres
// planes of f, on stream st.  Returns 1 when it ran, 0 when the frame is not eligible (the separate kernels run
// instead), -1 on error.
static int fused_eligible(const La *la, const x264vfw_cuda_image_t &planes, const uint8_t *src, long long stride)
{
    if (!la->fused || (la->in_csp & X264VFW_CUDA_CSP_MASK) != X264VFW_CUDA_CSP_BGRA || la->out_csp != X264VFW_CUDA_OUT_I420) return 0;
    if (la->p.chroma_format != 1 || planes.i_stride[1] != planes.i_stride[2]) return 0;
    return frontend_eligible(src, stride, 0, la->p.width, la->p.height, 1, planes.plane[0], planes.i_stride[0], planes.plane[1], planes.plane[2],
                             planes.i_stride[1], 0) ? 1 : 0;
}

static int launch_fused(La *la, cudaStream_t st, Frame *f, const x264vfw_cuda_image_t &planes, const uint8_t *src, long long stride)
{
    const bool aq_on = la->p.aq_mode != 0 && la->p.aq_strength != 0;
    FrontendJob j;
    memset(&j, 0, sizeof(j));
    j.dst_y = planes.plane[0]; j.dst_u = planes.plane[1]; j.dst_v = planes.plane[2];
    j.y_stride = planes.i_stride[0]; j.c_stride = planes.i_stride[1]; j.dst_frame_bytes = 0;
    j.lowres = f->lowres; j.lowres_frame_bytes = 0; j.lw = la->g.lw; j.lh = la->g.lh; j.lstride = la->g.lstride;
    j.lplane_bytes = la->g.lplane; j.lorigin = la->g.lorigin;
    j.qp_offset = f->qp_offset; j.qp_offset_aq = f->qp_offset_aq; j.inv_qscale = f->inv_qscale; j.stats = f->stats; j.mb_frame_stride = 0;
    j.aq_on = aq_on; j.aq_mode = la->p.aq_mode; j.strength = la->p.aq_strength * 1.0397f;
    j.log2_lut = la->d_log2_lut; j.exp2_lut = la->d_exp2_lut;
    j.w = la->p.width; j.h = la->p.height; j.mb_w = la->g.mb_w; j.mb_h = la->g.mb_h; j.luma_h = la->g.luma_h;
    j.flip = (la->in_csp & X264VFW_CUDA_CSP_VFLIP) != 0;
    j.k = make_rgb_kernel_coef(rgb_coef(la->colmatrix, la->fullrange));
    { ProfScope ps(la, K_FRONTEND, st); if (launch_frontend(st, j, src, stride, 0, 1) < 0) return -1; }
    la->n_launch++;
    if (aq_on && la->p.aq_mode >= 2) {
        AqJob aq;
        memset(&aq, 0, sizeof(aq));
        aq.aq_on = 1; aq.aq_mode = la->p.aq_mode; aq.aq_strength = la->p.aq_strength;
        aq.qp_offset = f->qp_offset; aq.qp_offset_aq = f->qp_offset_aq; aq.inv_qscale = f->inv_qscale; aq.exp2_lut = la->d_exp2_lut;
        ProfScope ps(la, K_AQ, st);
        if (launch_aq_auto(st, la->g, aq) < 0) return -1;
        la->n_launch++;
    }
    XV_CUDA_OK(cudaMemcpyAsync(f->h_stats, f->stats, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    XV_CUDA_OK(cudaEventRecord(f->ev_stats, st));
    return 1;
}

// [x264] x264_adaptive_quant_frame + x264_frame_init_lowres of frame f from the converted planes
// (device, encoder csp).  NV12 sessions hand the interleaved chroma plane to the AQ kernel, which
// reads U and V out of it: libx264 computes the same 4:2:0 energies on its internal NV12 frame.
static int frame_prep(La *la, Frame *f, const x264vfw_cuda_image_t &planes)
{
    const int w = la->p.width, hgt = la->p.height;
    // ---- [x264] x264_adaptive_quant_frame ----
    const bool planar_yuv = la->out_csp == X264VFW_CUDA_OUT_I420 || la->out_csp == X264VFW_CUDA_OUT_I422 || la->out_csp == X264VFW_CUDA_OUT_I444;
    const bool nv12 = la->out_csp == X264VFW_CUDA_OUT_NV12;
    const bool aq_on = la->p.aq_mode != 0 && la->p.aq_strength != 0;
    AqJob aq;
    aq.y = planes.plane[0]; aq.y_stride = planes.i_stride[0];
    aq.u = planar_yuv || nv12 ? planes.plane[1] : nullptr; aq.v = planar_yuv ? planes.plane[2] : nv12 ? planes.plane[1] + 1 : nullptr;
    aq.c_stride = planes.i_stride[1]; aq.c_step = nv12 ? 2 : 1;
    aq.chroma_format = planar_yuv ? la->p.chroma_format : nv12 ? 1 : 0;
    aq.aq_on = aq_on; aq.strength = la->p.aq_strength * 1.0397f;
    aq.aq_mode = la->p.aq_mode; aq.aq_strength = la->p.aq_strength;
    aq.qp_offset = f->qp_offset; aq.qp_offset_aq = f->qp_offset_aq; aq.inv_qscale = f->inv_qscale;
    aq.stats = f->stats; aq.log2_lut = la->d_log2_lut; aq.exp2_lut = la->d_exp2_lut;
    { ProfScope ps(la, K_AQ); if (launch_aq(la->st, la->g, aq) < 0) return -1; }
    XV_CUDA_OK(cudaMemcpyAsync(f->h_stats, f->stats, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, la->st));
    XV_CUDA_OK(cudaEventRecord(f->ev_stats, la->st));
    // ---- [x264] x264_frame_init_lowres ----
    LowresJob lj;
    lj.y = planes.plane[0]; lj.y_stride = planes.i_stride[0]; lj.w = w; lj.h = hgt; lj.dst = f->lowres;
    lj.luma_w = la->g.luma_w; lj.luma_h = la->g.luma_h; lj.lw = la->g.lw; lj.lh = la->g.lh;
    lj.lstride = la->g.lstride; lj.lplane_bytes = la->g.lplane; lj.lorigin = la->g.lorigin;
    lj.src_frame_bytes = 0; lj.dst_frame_bytes = 0;
    { ProfScope ps(la, K_LOWRES); if (launch_lowres_init(la->st, lj, 1) < 0) return -1; }
    la->n_launch += 2;
    return 0;
}

// ------------------------------------------------------------------------------------------
// Session worker.  put_frame only moves the caller's buffers (H2D, conversion, D2H on the I/O
// stream) and queues the frame; everything that touches the lookahead state -- frame slot, AQ,
// lowres planes, speculative searches, the decisions that become due -- runs here, in arrival
// order.  The caller may run io_depth frames ahead (ring of converted-plane buffers), so its
// wait for its own copies overlaps the decision of an earlier mini-GOP instead of adding to
// it.  Decisions are published deterministically: put(n) returns what frames <= n - io_depth
// made due, like [x264]'s own lookahead thread behind sync-lookahead.
// ------------------------------------------------------------------------------------------
// Frame f has been converted into `planes` (signalled by ev_csp on the I/O stream): [x264]
// x264_adaptive_quant_frame, x264_frame_init_lowres, then the searches this frame enables.
static int process_frame(La *la, Frame *f, const x264vfw_cuda_image_t &planes, cudaEvent_t ev_csp, cudaEvent_t ev_free, bool fused)
{
    const int w = la->p.width, hgt = la->p.height;
    if (ev_csp) {
        // Wait for the conversion HERE, on the host, not in the main stream: a stream-side wait
        // would make the next decision's synchronisation wait for this frame's copies as well.
        if (cudaEventQuery(ev_csp) != cudaSuccess) { const double t0 = now_s(); if (wait_event(la, ev_csp) < 0) return -1; la->t_csp_wait += now_s() - t0; }
        XV_CUDA_OK(cudaStreamWaitEvent(la->st, ev_csp, 0));
    }
    if (!fused && frame_prep(la, f, planes) < 0) return -1;   // the fused front end already produced AQ arrays and lowres planes
    XV_CUDA_OK(cudaEventRecord(ev_free, la->st));           // the planes may be overwritten again
    {   // every frame's intra costs are needed sooner or later and only depend on its lowres plane
        IntraJob ij;
        ij.plane0 = plane_org(la, f, 0); ij.intra_cost = f->intra_cost; ij.full = la->p.subme > 1; ij.satd = la->la_satd;
        ProfScope ps(la, K_INTRA);
        if (launch_intra(la->st, la->g, ij, la->do_edges) < 0) return -1;
        f->intra_dev = true;
        la->n_launch++;
    }
    f->ready = true;
    return speculate_searches(la, f);
}

static int run_due_decisions(La *la)
{
    const double t0 = now_s();
    int rc = 0;
    while ((int)la->next.size() > la->slicetype_length + la->decide_lag)
        if (decide_and_shift(la) < 0) { rc = -1; break; }
    la->t_decide += now_s() - t0;
    return rc;
}

static void worker_main(La *la)
{
    cudaSetDevice(la->device);
    for (;;) {
        std::array<long, 3> job;
        {
            std::unique_lock<std::mutex> lk(la->mu);
            la->cv.wait(lk, [&] { return la->wstop || !la->jobs.empty(); });
            if (la->jobs.empty()) return;
            job = la->jobs.front(); la->jobs.pop_front();
        }
        const int slot = (int)job[1];
        // The decision that becomes due with frame n only looks at frames < n (decide_lag >= 1):
        // run it first, and enqueue this frame's AQ / lowres / searches afterwards -- by then its
        // conversion on the I/O stream has finished, so the main stream never waits for a copy.
        // (When the conversion is already done -- device-resident sources -- the frame goes first,
        // so that its searches start as early as possible.)
        int rc = 0;
        const bool fused = job[2] != 0;
        Frame *f = la->ring_frame[slot];                       // chosen io_depth frames ago (see La::ring_frame)
        if (!f || f->i_frame != (int)job[0]) { set_error("lookahead worker: frame slot hand-over out of step"); rc = -1; }
        const bool converted = rc == 0 && cudaEventQuery(la->ev_csp_ring[slot]) == cudaSuccess;
        if (rc == 0 && converted) rc = process_frame(la, f, la->planes_ring[slot], la->ev_csp_ring[slot], la->ev_free_ring[slot], fused);
        if (rc == 0) { la->n_input = (int)job[0] + 1; la->next.push_back(f); rc = run_due_decisions(la); }
        if (rc == 0 && !converted) rc = process_frame(la, f, la->planes_ring[slot], la->ev_csp_ring[slot], la->ev_free_ring[slot], fused);
        if (rc == 0) {
            // the slot of put (n + io_depth), which re-uses this ring position
            Frame *nf = frame_get(la, (int)job[0] + la->io_depth);
            if (!nf) rc = -1;
            else if (cudaEventRecord(la->ev_frame_ready[slot], la->st) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = -1; }
            la->ring_frame[slot] = nf;
        }
        std::lock_guard<std::mutex> lk(la->mu);
        if (rc < 0 && !la->werr) { la->werr = -1; la->werr_msg = x264vfw_cuda_last_error(); }
        while (!la->outq.empty()) { la->doneq.emplace_back(job[0], la->outq.front()); la->outq.pop_front(); }
        la->processed = job[0] + 1;
        la->cv.notify_all();
    }
}

// Wait until the worker is done with `count` frames; then publish what frames <= upto made due.
static int worker_wait(La *la, long count, long upto)
{
    if (!la->worker.joinable()) return 0;
    std::unique_lock<std::mutex> lk(la->mu);
    la->cv.wait(lk, [&] { return la->processed >= count; });
    while (!la->doneq.empty() && la->doneq.front().first <= upto) { la->pubq.push_back(la->doneq.front().second); la->doneq.pop_front(); }
    if (la->werr) { set_error("%s", la->werr_msg.c_str()); return -1; }
    return 0;
}
// everything submitted so far is processed and published (white-box calls, flush, close)
static int worker_join(La *la) { return worker_wait(la, la->n_put, la->n_put); }

} // namespace xv

using namespace xv;

extern "C" {

int x264vfw_cuda_la_params_preset(x264vfw_cuda_la_params *p, const char *preset, int width, int height)
{
    // x264 defaults (x264_param_default) + the preset deltas x264vfw documents (config.c:1460-1498)
    if (!p || !preset) return -1;
    memset(p, 0, sizeof(*p));
    p->width = width; p->height = height; p->chroma_format = 1;
    p->bframes = 3; p->b_adapt = 1; p->b_pyramid = 2; p->b_bias = 0;
    p->rc_lookahead = 40; p->b_mbtree = 1; p->scenecut = 40;
    p->keyint_max = 250; p->keyint_min = 25; p->open_gop = 0;
    p->weightp = 2; p->weightb = 1; p->subme = 7; p->me_method = 1; p->me_range = 16; p->mv_range = 512;
    p->aq_mode = 1; p->aq_strength = 1.0f; p->qcompress = 0.6f; p->frame_reference = 3;
    p->lookahead_threads = 1; p->fps_num = 25; p->fps_den = 1; p->b_psy = 1;
    if (!strcmp(preset, "ultrafast")) {
        p->frame_reference = 1; p->scenecut = 0; p->bframes = 0; p->b_adapt = 0; p->me_method = 0; p->subme = 0;
        p->aq_mode = 0; p->b_mbtree = 0; p->rc_lookahead = 0; p->weightp = 0; p->weightb = 0;
    } else if (!strcmp(preset, "superfast")) {
        p->me_method = 0; p->subme = 1; p->frame_reference = 1; p->b_mbtree = 0; p->rc_lookahead = 0; p->weightp = 1;
    } else if (!strcmp(preset, "veryfast")) {
        p->subme = 2; p->frame_reference = 1; p->weightp = 1; p->rc_lookahead = 10;
    } else if (!strcmp(preset, "faster")) {
        p->frame_reference = 2; p->subme = 4; p->weightp = 1; p->rc_lookahead = 20;
    } else if (!strcmp(preset, "fast")) {
        p->frame_reference = 2; p->subme = 6; p->weightp = 1; p->rc_lookahead = 30;
    } else if (!strcmp(preset, "medium")) {
    } else if (!strcmp(preset, "slow")) {
        p->subme = 8; p->frame_reference = 5; p->rc_lookahead = 50;
    } else if (!strcmp(preset, "slower")) {
        p->me_method = 2; p->subme = 9; p->frame_reference = 8; p->b_adapt = 2; p->rc_lookahead = 60;
    } else if (!strcmp(preset, "veryslow")) {
        p->me_method = 2; p->subme = 10; p->me_range = 24; p->frame_reference = 16; p->b_adapt = 2; p->bframes = 8; p->rc_lookahead = 60;
    } else if (!strcmp(preset, "placebo")) {
        // config.c:1493 / [x264] x264_param_apply_preset: bframes 16, b-adapt 2, me tesa, merange 24, ref 16, subme 11, rc-lookahead 60
        p->me_method = 4; p->subme = 11; p->me_range = 24; p->frame_reference = 16; p->b_adapt = 2; p->bframes = 16; p->rc_lookahead = 60;
    } else { set_error("unknown preset %s", preset); return -1; }
    return 0;
}

static int apply_one_tune(x264vfw_cuda_la_params *p, const char *tune)
{
    if (!strcmp(tune, "film") || !strcmp(tune, "none") || !*tune) {
    } else if (!strcmp(tune, "animation")) {
        p->frame_reference = p->frame_reference > 1 ? p->frame_reference * 2 : 1;
        p->aq_strength = 0.6f; p->bframes += 2;
    } else if (!strcmp(tune, "grain")) {
        p->aq_strength = 0.5f; p->qcompress = 0.8f;
    } else if (!strcmp(tune, "stillimage")) {
        p->aq_strength = 1.2f;
    } else if (!strcmp(tune, "psnr")) {
        p->aq_mode = 0; p->b_psy = 0;
    } else if (!strcmp(tune, "ssim")) {
        p->aq_mode = 2; p->b_psy = 0;
    } else if (!strcmp(tune, "fastdecode")) {
        p->weightb = 0; p->weightp = 0;
    } else if (!strcmp(tune, "zerolatency")) {
        p->rc_lookahead = 0; p->bframes = 0; p->b_mbtree = 0;
    } else if (!strcmp(tune, "touhou")) {
        p->frame_reference = p->frame_reference > 1 ? p->frame_reference * 2 : 1;
        p->aq_strength = 1.3f;
    } else { set_error("unknown tune %s", tune); return -1; }
    return 0;
}

int x264vfw_cuda_la_params_tune(x264vfw_cuda_la_params *p, const char *tune)
{
    // [x264] x264_param_apply_tune reduced to the fields the lookahead reads (the tunings x264vfw
    // offers in its configuration dialog, config.c "Tuning"); apply after the preset.  Like upstream
    // the string may hold several names separated by ',', '.', '/', '-', '+' or blanks -- the
    // reference passes e.g. "film,fastdecode,zerolatency" (codec.c:1430-1445).
    if (!p || !tune) return -1;
    char buf[128];
    size_t n = strlen(tune);
    if (n >= sizeof(buf)) { set_error("tune string too long"); return -1; }
    memcpy(buf, tune, n + 1);
    char *save = nullptr;
    int seen = 0;
    for (char *t = strtok_r(buf, ",./-+ ", &save); t; t = strtok_r(nullptr, ",./-+ ", &save)) {
        if (apply_one_tune(p, t) < 0) return -1;
        seen++;
    }
    if (!seen && apply_one_tune(p, "") < 0) return -1;
    if (p->bframes > BMAX) p->bframes = BMAX;
    return 0;
}

int x264vfw_cuda_la_open(x264vfw_cuda_la **pla, const x264vfw_cuda_la_params *params, int device,
                         int in_csp, int out_csp, int colmatrix, int fullrange, int keep_frames)
{
    if (!pla || !params) { set_error("null argument"); return -1; }
    *pla = nullptr;
    int ndev = 0;
    XV_CUDA_OK(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) { set_error("no CUDA device: this library has no CPU fallback"); return -1; }
    if (device < 0) XV_CUDA_OK(cudaGetDevice(&device));
    XV_CUDA_OK(cudaSetDevice(device));
    if (params->width <= 0 || params->height <= 0 || (params->width & 1) || (params->height & 1)) { set_error("width/height must be positive and even"); return -1; }
    La *la = new La();
    memset(la->ev_me, 0, sizeof(la->ev_me));
    la->p = *params;
    x264vfw_cuda_la_params &p = la->p;
    if (p.bframes > BMAX) p.bframes = BMAX;
    if (p.bframes < 0) p.bframes = 0;
    if (p.rc_lookahead > LMAX) p.rc_lookahead = LMAX;
    if (p.keyint_min <= 0) { int fps = p.fps_num / (p.fps_den > 0 ? p.fps_den : 1); p.keyint_min = p.keyint_max / 10 < fps ? p.keyint_max / 10 : fps; }
    if (p.chroma_format < 0 || p.chroma_format > 3) { set_error("bad chroma_format"); delete la; return -1; }
    // [x264] validate_parameters: weightp off + mb-tree + psy => X264_WEIGHTP_FAKE: the lookahead still
    // analyses (and uses) luma weights and records f_weighted_cost_delta for macroblock_tree_finish
    if (!p.weightp && p.b_mbtree && p.b_psy) p.weightp = WEIGHTP_FAKE;
    // [x264] macroblock_tree has a separate extrapolation branch for rc-lookahead 0 (no BASELINE
    // configuration uses it; the presets that set rc-lookahead 0 also switch mb-tree off): not
    // restated, so the combination is refused instead of answered with different qp offsets
    if (p.b_mbtree && p.rc_lookahead <= 0) { set_error("mb-tree with rc-lookahead 0 (lookaheadless mb-tree) is not supported: set rc_lookahead >= 1 or b_mbtree = 0"); delete la; return -1; }
    if (p.aq_mode < 0 || p.aq_mode > 3) { set_error("bad aq_mode %d (0 off, 1 variance, 2 auto-variance, 3 auto-variance biased)", p.aq_mode); delete la; return -1; }
    la->device = device; la->in_csp = in_csp; la->out_csp = out_csp; la->colmatrix = colmatrix; la->fullrange = fullrange;
    la->keep_frames = keep_frames;
    if (((in_csp & X264VFW_CUDA_CSP_MASK) == X264VFW_CUDA_CSP_YUYV || (in_csp & X264VFW_CUDA_CSP_MASK) == X264VFW_CUDA_CSP_UYVY) && out_csp == X264VFW_CUDA_OUT_I444)
        la->ext = X264VFW_CUDA_EXT_422_TO_I444;
    // "RGB24/32 -> NV12" (BASELINE config 3) is the documented extension conversion as well (csp.c:490-492 has no such pair)
    if (((in_csp & X264VFW_CUDA_CSP_MASK) == X264VFW_CUDA_CSP_BGR || (in_csp & X264VFW_CUDA_CSP_MASK) == X264VFW_CUDA_CSP_BGRA) && out_csp == X264VFW_CUDA_OUT_NV12)
        la->ext = X264VFW_CUDA_EXT_RGB_TO_NV12;
    // Kept-RGB sessions (codec.c:293-297 -> X264_CSP_BGR/BGRA): libx264 analyses the G plane of its internal
    // planar GBR frame with 4:4:4 AQ energies; that frame is not built here, so refuse instead of analysing the
    // packed bytes as if they were luma.
    if (out_csp == X264VFW_CUDA_OUT_BGR || out_csp == X264VFW_CUDA_OUT_BGRA) {
        set_error("lookahead sessions on kept-RGB encoder formats (X264_CSP_BGR/BGRA) are not supported: convert to a YUV encoder csp");
        delete la; return -1;
    }
    if (const char *e = getenv("X264VFW_CUDA_DECIDE_LAG")) { la->decide_lag = atoi(e); if (la->decide_lag < 0) la->decide_lag = 0; if (la->decide_lag > 4) la->decide_lag = 4; }
    if (const char *e = getenv("X264VFW_CUDA_SPECULATE")) la->speculate = atoi(e) != 0;
    if (const char *e = getenv("X264VFW_CUDA_SPEC_THRESHOLD")) la->spec_threshold = atof(e);
    if (const char *e = getenv("X264VFW_CUDA_PREDICT")) la->predict = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_TREE_CHAIN")) la->tree_chain = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_ME_SIDE")) { la->me_side = atoi(e); if (la->me_side < 1) la->me_side = 1; if (la->me_side > ME_SIDE) la->me_side = ME_SIDE; }
    if (const char *e = getenv("X264VFW_CUDA_ME_VARIANT")) la->me_variant = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_ME_PASSES")) { la->me_passes = atoi(e); if (la->me_passes < 1) la->me_passes = 1; if (la->me_passes > 4) la->me_passes = 4; }
    if (const char *e = getenv("X264VFW_CUDA_ME_RELAX")) { la->me_relax = atoi(e); if (la->me_relax < 0) la->me_relax = 0; if (la->me_relax > 3) la->me_relax = 3; }
    if (la->me_passes + la->me_relax > 4) la->me_relax = 4 - la->me_passes;
    if (const char *e = getenv("X264VFW_CUDA_ME_GUESS")) la->me_guess = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_ME_FORCE_MISS")) la->me_force_miss = atoi(e) != 0;
    if (const char *e = getenv("X264VFW_CUDA_ASYNC")) la->async = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_FUSED")) la->fused = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_CALLER_BLOCK")) la->caller_block = atoi(e);
    if (const char *e = getenv("X264VFW_CUDA_SYNC")) {      // spin (default) | yield | notify | block | hybrid | hybrid:<microseconds>
        la->yielding = !strcmp(e, "yield");
        la->notify = !strcmp(e, "notify");
        la->blocking = strcmp(e, "spin") != 0 && !la->yielding && !la->notify;
        if (!strcmp(e, "block")) la->spin_us = 0;
        else if (!strncmp(e, "hybrid:", 7)) la->spin_us = atoi(e + 7);
    }
    if (const char *e = getenv("X264VFW_CUDA_IO_DEPTH")) { la->io_depth = atoi(e); if (la->io_depth < 1) la->io_depth = 1; if (la->io_depth > 4) la->io_depth = 4; }
    if (const char *e = getenv("X264VFW_CUDA_ME_ROWS")) la->me_rows = atoi(e);
    else la->me_rows = -1;   // resolved below once the geometry is known
    x264vfw_cuda_lowres_geom lg;
    x264vfw_cuda_lowres_geometry(&lg, p.width, p.height);
    LaGeom &g = la->g;
    g.width = p.width; g.height = p.height; g.mb_w = lg.mb_w; g.mb_h = lg.mb_h; g.mb_count = lg.mb_w * lg.mb_h;
    g.luma_w = lg.luma_w; g.luma_h = lg.luma_h; g.lw = lg.lw; g.lh = lg.lh; g.lstride = lg.lstride;
    g.lplane = lg.lplane_bytes; g.lorigin = lg.lorigin;
    if (la->me_rows < 0) {
        // at most mb_w/2 rows of a search can be busy at once (a row is mb_w steps long and rows
        // start 2 steps apart); 2/3 of that keeps the pipeline full with far fewer idle warps
        const int busy = g.mb_h < (g.mb_w + 1) / 2 ? g.mb_h : (g.mb_w + 1) / 2;
        la->me_rows = busy * 2 / 3 > 4 ? busy * 2 / 3 : 4;
    }
    // [x264] lowres_context_init
    if (p.subme > 1) { la->la_me_hex = p.me_method >= 1; la->la_subpel_refine = 4; }
    else { la->la_me_hex = 0; la->la_subpel_refine = 2; }
    la->la_satd = p.subme > 1;
    la->do_edges = p.b_mbtree || g.mb_w <= 2 || g.mb_h <= 2;
    la->slicetype_length = p.bframes > p.rc_lookahead ? p.bframes : p.rc_lookahead;
    la->i_last_keyframe = -p.keyint_max;
    // the main stream carries the latency-critical chain (the host blocks on it); speculative
    // searches on the side streams yield to it
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&la->st, cudaStreamNonBlocking, prio_hi) != cudaSuccess) { set_error("cudaStreamCreate failed"); delete la; return -1; }

    // tables: [x264] x264_analyse_init_costs for X264_LOOKAHEAD_QP (lambda 1), x264_log2_lut, x264_exp2_lut
    {
        const int n = 2 * 4 * p.mv_range;
        std::vector<uint16_t> tab(2 * n + 1);
        for (int i = 0; i <= n; i++) {
            float lg2 = i == 0 ? 0.718f : log2f((float)(i + 1)) * 2.0f + 1.718f;
            int c = (int)(1 * lg2 + .5f);
            tab[n - i] = tab[n + i] = (uint16_t)(c < 65535 ? c : 65535);
        }
        la->cost_mv_half = n;
        float l2[128]; uint8_t e2[64];
        for (int i = 0; i < 128; i++) l2[i] = (float)(round(log2(1.0 + i / 128.0) * 100000.0) / 100000.0);
        for (int i = 0; i < 64; i++) e2[i] = (uint8_t)lround(256.0 * (pow(2.0, i / 64.0) - 1.0));
        bool ok = cudaMalloc((void **)&la->d_cost_mv, tab.size() * 2) == cudaSuccess &&
                  cudaMalloc((void **)&la->d_log2_lut, sizeof(l2)) == cudaSuccess &&
                  cudaMalloc((void **)&la->d_exp2_lut, 64) == cudaSuccess &&
                  cudaMemcpy(la->d_cost_mv, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(la->d_log2_lut, l2, sizeof(l2), cudaMemcpyHostToDevice) == cudaSuccess &&
                  cudaMemcpy(la->d_exp2_lut, e2, 64, cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMalloc((void **)&la->d_weight_buf, g.lplane + 64) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_ready, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_io, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_sync, cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_h2d, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_csp, cudaEventDisableTiming) == cudaSuccess &&
             cudaEventCreateWithFlags(&la->ev_planes_free, cudaEventDisableTiming) == cudaSuccess &&
             cudaMalloc((void **)&la->d_results, RESULT_SLOTS * 4 * sizeof(int)) == cudaSuccess &&
             cudaMallocHost((void **)&la->h_results, RESULT_SLOTS * 4 * sizeof(int)) == cudaSuccess &&
             cudaMalloc((void **)&la->d_wscore, 64) == cudaSuccess &&
             cudaMallocHost((void **)&la->h_wscore, 64) == cudaSuccess;
        for (int k = 0; k < TREE_RING && ok; k++)
            ok = ok && cudaMallocHost((void **)&la->h_tree[k], TREE_MAX_STEPS * sizeof(TreeStep)) == cudaSuccess &&
                 cudaMalloc((void **)&la->d_tree[k], TREE_MAX_STEPS * sizeof(TreeStep)) == cudaSuccess &&
                 cudaEventCreateWithFlags(&la->ev_tree[k], cudaEventDisableTiming) == cudaSuccess;
        if (const char *e = getenv("X264VFW_CUDA_STATS")) la->stats_verbose = atoi(e) >= 2;
        if (getenv("X264VFW_CUDA_STATS"))
            ok = ok && cudaMalloc((void **)&la->d_me_stats, 8 * sizeof(int)) == cudaSuccess && cudaMemset(la->d_me_stats, 0, 8 * sizeof(int)) == cudaSuccess;
        la->st_me[0] = la->st;
        for (int e = 0; e <= la->me_side && ok; e++) {
            ok = ok && cudaMalloc((void **)&la->d_rec[e], (size_t)XV_ME_MAX_JOBS * g.mb_count * sizeof(int2)) == cudaSuccess &&
                 cudaMemset(la->d_rec[e], 0, (size_t)XV_ME_MAX_JOBS * g.mb_count * sizeof(int2)) == cudaSuccess &&
                 cudaMalloc((void **)&la->d_ticket[e], XV_ME_MAX_JOBS * sizeof(int)) == cudaSuccess &&
                 cudaMalloc((void **)&la->d_assumed[e], (size_t)XV_ME_MAX_JOBS * g.mb_count * sizeof(int4)) == cudaSuccess;
            if (e) ok = ok && cudaStreamCreateWithPriority(&la->st_me[e], cudaStreamNonBlocking, prio_lo) == cudaSuccess;
            for (int k = 0; k < ME_EVENTS && ok; k++)
                ok = ok && cudaEventCreateWithFlags(&la->ev_me[e][k], cudaEventDisableTiming) == cudaSuccess;
        }
        int64_t pbytes = x264vfw_cuda_picture_layout(&la->planes_img, nullptr, out_csp, p.width, p.height);
        if (pbytes < 0) { set_error("bad encoder csp %d", out_csp); ok = false; }
        else {
            la->d_planes_bytes = (size_t)pbytes;
            ok = ok && cudaMalloc((void **)&la->d_planes, (size_t)pbytes + 256) == cudaSuccess;
            if (ok) x264vfw_cuda_picture_layout(&la->planes_img, la->d_planes, out_csp, p.width, p.height);
        }
        if (!ok) { if (!*x264vfw_cuda_last_error()) set_error("lookahead allocation failed"); x264vfw_cuda_la_close((x264vfw_cuda_la *)la); return -1; }
    }
    // ev_frame_ready[0] for the synchronous path (keep_frames / X264VFW_CUDA_ASYNC=0); the ring creates its own below
    if (!(la->async && la->decide_lag >= 1 && !keep_frames) && cudaEventCreateWithFlags(&la->ev_frame_ready[0], cudaEventDisableTiming) != cudaSuccess) {
        set_error("cudaEventCreate failed"); x264vfw_cuda_la_close((x264vfw_cuda_la *)la); return -1;
    }
    // pre-allocate the frame pool: allocation is slow and serialises across sessions
    if (!keep_frames) {
        const int want = la->slicetype_length + p.bframes + 6 + la->decide_lag + la->io_depth;
        for (int i = 0; i < want; i++) {
            Frame *f = frame_alloc(la);
            if (!f) { x264vfw_cuda_la_close((x264vfw_cuda_la *)la); return -1; }
            la->pool.push_back(f);
        }
        for (int i = 0; i < 4; i++) { float *q = qp_staging(la); if (q) la->qp_free.push_back(q); }
    }
    if (la->async && la->decide_lag >= 1 && !keep_frames) {
        bool ok = cudaStreamCreateWithPriority(&la->st_io, cudaStreamNonBlocking, prio_hi) == cudaSuccess;
        for (int k = 0; k < la->io_depth && ok; k++) {
            ok = cudaEventCreateWithFlags(&la->ev_frame_ready[k], cudaEventDisableTiming) == cudaSuccess &&
                 cudaMalloc((void **)&la->d_planes_ring[k], la->d_planes_bytes + 256) == cudaSuccess &&
                 cudaEventCreateWithFlags(&la->ev_csp_ring[k], cudaEventDisableTiming | (la->blocking ? cudaEventBlockingSync : 0)) == cudaSuccess &&
                 cudaEventCreateWithFlags(&la->ev_free_ring[k], cudaEventDisableTiming) == cudaSuccess;
            if (ok) x264vfw_cuda_picture_layout(&la->planes_ring[k], la->d_planes_ring[k], out_csp, p.width, p.height);
        }
        for (int k = 0; k < la->io_depth && ok; k++) {
            la->ring_frame[k] = frame_get(la, k);
            ok = la->ring_frame[k] && cudaEventRecord(la->ev_frame_ready[k], la->st) == cudaSuccess;
        }
        if (!ok) { if (!*x264vfw_cuda_last_error()) set_error("lookahead allocation failed"); x264vfw_cuda_la_close((x264vfw_cuda_la *)la); return -1; }
        la->worker = std::thread(worker_main, la);
    }
    *pla = (x264vfw_cuda_la *)la;
    return 0;
}

void x264vfw_cuda_la_close(x264vfw_cuda_la *h)
{
    La *la = (La *)h;
    if (!la) return;
    cudaSetDevice(la->device);
    if (la->worker.joinable()) {
        worker_join(la);
        { std::lock_guard<std::mutex> lk(la->mu); la->wstop = true; la->cv.notify_all(); }
        la->worker.join();
    }
    if (getenv("X264VFW_CUDA_STATS")) {
        fprintf(stderr, "[x264vfw_cuda] frames %d searches asked for by (list,dist):", la->n_input);
        for (int l = 0; l < 2; l++) for (int d = 0; d <= la->p.bframes; d++) fprintf(stderr, " l%d/d%d=%llu", l, d + 1, (unsigned long long)la->n_logical[l][d]);
        fprintf(stderr, "  | launched:");
        for (int l = 0; l < 2; l++) for (int d = 0; d <= la->p.bframes; d++) fprintf(stderr, " l%d/d%d=%llu", l, d + 1, (unsigned long long)la->n_launched[l][d]);
        fprintf(stderr, "  | speculative jobs %llu, on-demand launches %llu (%llu jobs)\n", (unsigned long long)la->n_spec_jobs,
                (unsigned long long)la->n_ondemand, (unsigned long long)la->n_ondemand_jobs);
        if (la->io_n) fprintf(stderr, "[x264vfw_cuda] I/O stream us per frame: wait for planes %.0f, H2D %.0f, conversion %.0f, D2H %.0f\n",
                              1e3 * la->io_ms[0] / la->io_n, 1e3 * la->io_ms[1] / la->io_n, 1e3 * la->io_ms[2] / la->io_n, 1e3 * la->io_ms[3] / la->io_n);
        fprintf(stderr, "[x264vfw_cuda] waits: light kernels only %llu x %.0f us, speculative search in flight %llu x %.0f us, on-demand search %llu x %.0f us\n",
                (unsigned long long)la->n_sync_kind[0], 1e6 * la->t_sync_kind[0] / (la->n_sync_kind[0] ? la->n_sync_kind[0] : 1),
                (unsigned long long)la->n_sync_kind[1], 1e6 * la->t_sync_kind[1] / (la->n_sync_kind[1] ? la->n_sync_kind[1] : 1),
                (unsigned long long)la->n_sync_kind[2], 1e6 * la->t_sync_kind[2] / (la->n_sync_kind[2] ? la->n_sync_kind[2] : 1));
        fprintf(stderr, "[x264vfw_cuda] host us per frame: put %.0f decide %.0f (of which waiting %.0f) final wait for the borrowed buffers %.0f, worker waiting for a conversion %.0f\n",
                1e6 * la->t_put / (la->n_input ? la->n_input : 1), 1e6 * la->t_decide / (la->n_input ? la->n_input : 1),
                1e6 * la->t_sync / (la->n_input ? la->n_input : 1), 1e6 * la->t_io / (la->n_input ? la->n_input : 1), 1e6 * la->t_csp_wait / (la->n_input ? la->n_input : 1));
    }
    if (la->st) cudaStreamSynchronize(la->st);
    for (int e = 1; e <= ME_SIDE; e++) if (la->st_me[e]) cudaStreamSynchronize(la->st_me[e]);
    if (la->d_me_stats && getenv("X264VFW_CUDA_STATS")) {
        int v[8] = {0};
        cudaMemcpy(v, la->d_me_stats, sizeof(v), cudaMemcpyDeviceToHost);
        fprintf(stderr, "[x264vfw_cuda] speculative search: kept %d, re-searched in order %d (%.2f%%); searched per pass: %d %d %d %d\n",
                v[0], v[1], 100.0 * v[1] / (v[0] + v[1] > 0 ? v[0] + v[1] : 1), v[2], v[3], v[4], v[5]);
    }
    cudaFree(la->d_me_stats);
    prof_resolve(la);
    for (cudaEvent_t e : la->prof.pool) cudaEventDestroy(e);
    for (Frame *f : la->pool) frame_free(f);
    for (Decision &d : la->outq) if (d.h_qp) cudaFreeHost(d.h_qp);
    for (Decision &d : la->pubq) if (d.h_qp) cudaFreeHost(d.h_qp);
    for (auto &pd : la->doneq) if (pd.second.h_qp) cudaFreeHost(pd.second.h_qp);
    for (int k = 0; k < 4; k++) { if (la->ev_frame_ready[k]) cudaEventDestroy(la->ev_frame_ready[k]); cudaFree(la->d_planes_ring[k]); if (la->ev_csp_ring[k]) cudaEventDestroy(la->ev_csp_ring[k]); if (la->ev_free_ring[k]) cudaEventDestroy(la->ev_free_ring[k]); }
    for (float *q : la->qp_free) cudaFreeHost(q);
    cudaFree(la->d_cost_mv); cudaFree(la->d_log2_lut); cudaFree(la->d_exp2_lut); cudaFree(la->d_weight_buf);
    for (int e = 0; e <= ME_SIDE; e++) {
        cudaFree(la->d_rec[e]); cudaFree(la->d_ticket[e]); cudaFree(la->d_assumed[e]);
        for (int k = 0; k < ME_EVENTS; k++) if (la->ev_me[e][k]) cudaEventDestroy(la->ev_me[e][k]);
        if (e && la->st_me[e]) cudaStreamDestroy(la->st_me[e]);
    }
    for (int k = 0; k < TREE_RING; k++) {
        if (la->h_tree[k]) cudaFreeHost(la->h_tree[k]);
        cudaFree(la->d_tree[k]);
        if (la->ev_tree[k]) cudaEventDestroy(la->ev_tree[k]);
    }
    if (la->ev_ready) cudaEventDestroy(la->ev_ready);
    if (la->ev_io) cudaEventDestroy(la->ev_io);
    if (la->ev_sync) cudaEventDestroy(la->ev_sync);
    if (la->ev_h2d) cudaEventDestroy(la->ev_h2d);
    if (la->ev_csp) cudaEventDestroy(la->ev_csp);
    if (la->ev_planes_free) cudaEventDestroy(la->ev_planes_free);
    if (la->st_io) { cudaStreamSynchronize(la->st_io); cudaStreamDestroy(la->st_io); } cudaFree(la->d_results); cudaFree(la->d_wscore); cudaFree(la->d_planes); cudaFree(la->d_src);
    if (la->h_results) cudaFreeHost(la->h_results);
    if (la->h_wscore) cudaFreeHost(la->h_wscore);
    if (la->st) cudaStreamDestroy(la->st);
    delete la;
}

int x264vfw_cuda_la_put_frame(x264vfw_cuda_la *h, const x264vfw_cuda_image_t *src, int src_on_device, x264vfw_cuda_image_t *conv_pic)
{
    La *la = (La *)h;
    if (!la || !src) { set_error("null argument"); return -1; }
    const double t_begin = now_s();
    XV_CUDA_OK(cudaSetDevice(la->device));
    const int w = la->p.width, hgt = la->p.height;
    const int in = la->in_csp & X264VFW_CUDA_CSP_MASK;
    x264vfw_cuda_image_t planes = la->planes_img;           // device, tight, encoder csp
    const bool host_src = !src_on_device && in != X264VFW_CUDA_CSP_NONE;
    const bool borrowed = conv_pic || !src_on_device;
    // A device source is borrowed for the call like a host one (src_on_device == 1: its last reader, the
    // conversion or the plane copy, has finished when the call returns) unless the caller declares it
    // RESIDENT (src_on_device == 2: it stays unmodified until x264vfw_cuda_la_flush / _close, e.g. a clip
    // that lives in HBM) -- then nothing waits.
    const bool wait_dev_src = src_on_device == 1 && !conv_pic;
    auto chroma_rows = [&](int csp_is_420, int i) { return (i && csp_is_420) ? hgt / 2 : hgt; };

    if (borrowed && !la->st_io) {
        // host <-> device copies get their own stream (created on first use: device-resident
        // sessions keep one hardware queue less busy)
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        XV_CUDA_OK(cudaStreamCreateWithPriority(&la->st_io, cudaStreamNonBlocking, prio_hi));
    }
    const bool async = la->worker.joinable();
    const long n = la->n_put;
    const int slot = async ? (int)(n % la->io_depth) : 0;
    if (async) {
        planes = la->planes_ring[slot];
        // the worker must be done with the frame that used this slot (its AQ / lowres kernels are
        // then at least enqueued, ev_free_ring orders the conversion after them on the device)
        if (n >= la->io_depth && worker_wait(la, n - la->io_depth + 1, n - la->io_depth) < 0) return -1;
    }
    Frame *f = nullptr;
    if (!async) {
        f = frame_get(la, la->n_input);
        if (!f) return -1;
        la->n_input++;
        if (la->ev_frame_ready[0]) XV_CUDA_OK(cudaEventRecord(la->ev_frame_ready[0], la->st));   // the slot's reset is queued on the main stream
    }
    Frame *ff = async ? la->ring_frame[slot] : f;           // the frame slot this put fills (fused front end writes into it)
    bool did_fuse = false;

    const int out420 = la->out_csp == X264VFW_CUDA_OUT_I420 || la->out_csp == X264VFW_CUDA_OUT_NV12;
    x264vfw_cuda_image_t geo_out;
    x264vfw_cuda_picture_layout(&geo_out, nullptr, la->out_csp, w, hgt);
    // Stage 1 of a frame whose buffers are borrowed (host source and/or conv_pic) runs entirely
    // on the I/O stream -- H2D, conversion, D2H of the converted planes -- so that all of it
    // overlaps the decision logic below, which keeps the main stream to itself.
    cudaStream_t st1 = (borrowed || async) ? la->st_io : la->st;
    cudaEvent_t ev_free = async ? la->ev_free_ring[slot] : la->ev_planes_free, ev_csp = async ? la->ev_csp_ring[slot] : la->ev_csp;
    const bool planes_in_use = async ? n >= la->io_depth : la->planes_busy;
    x264vfw_cuda_image_t dsrc = *src;
    dsrc.i_csp = la->in_csp;
    auto stage1 = [&]() -> int {
        const bool iop = borrowed && la->d_me_stats;
        if (iop && !la->io_ev[0]) for (int k = 0; k < 5; k++) cudaEventCreate(&la->io_ev[k]);
        if (iop) cudaEventRecord(la->io_ev[0], la->st_io);
        // the previous frame's AQ / lowres kernels read the planes this conversion overwrites
        if ((borrowed || async) && planes_in_use) XV_CUDA_OK(cudaStreamWaitEvent(la->st_io, ev_free, 0));
        if (iop) cudaEventRecord(la->io_ev[1], la->st_io);
        if (host_src) {
            x264vfw_cuda_image_t geo;
            if (x264vfw_cuda_img_fill(&geo, nullptr, in, w, hgt) < 0) return -1;
            const int in420 = in == X264VFW_CUDA_CSP_I420 || in == X264VFW_CUDA_CSP_YV12 || in == X264VFW_CUDA_CSP_NV12;
            size_t need = 0, off[4];
            for (int i = 0; i < geo.i_plane; i++) { off[i] = need; need += ((size_t)src->i_stride[i] * chroma_rows(in420, i) + 255) & ~(size_t)255; }
            if (la->d_src_bytes < need) {
                if (la->d_src) { XV_CUDA_OK(cudaStreamSynchronize(la->st)); XV_CUDA_OK(cudaStreamSynchronize(la->st_io)); cudaFree(la->d_src); la->d_src = nullptr; }
                XV_CUDA_OK(cudaMalloc((void **)&la->d_src, need + 256));
                la->d_src_bytes = need;
            }
            for (int i = 0; i < geo.i_plane; i++) {
                dsrc.plane[i] = la->d_src + off[i];
                XV_CUDA_OK(cudaMemcpyAsync(dsrc.plane[i], src->plane[i], (size_t)src->i_stride[i] * chroma_rows(in420, i), cudaMemcpyHostToDevice, la->st_io));
            }
        }
        if (iop) cudaEventRecord(la->io_ev[2], la->st_io);
        if (in == X264VFW_CUDA_CSP_NONE) {
            // planar frame already in the encoder csp: copy rows into the tight device planes
            for (int i = 0; i < geo_out.i_plane; i++)
                XV_CUDA_OK(cudaMemcpy2DAsync(planes.plane[i], planes.i_stride[i], src->plane[i], src->i_stride[i], geo_out.i_stride[i], chroma_rows(out420, i),
                                             src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st1));
        } else if (ff && fused_eligible(la, planes, dsrc.plane[0], dsrc.i_stride[0])) {
            // one kernel: converted planes + AQ arrays + lowres planes of the frame slot
            if (st1 != la->st) XV_CUDA_OK(cudaStreamWaitEvent(st1, la->ev_frame_ready[slot], 0));
            if (launch_fused(la, st1, ff, planes, dsrc.plane[0], dsrc.i_stride[0]) < 0) return -1;
            did_fuse = true;
        } else {
            { ProfScope ps(la, K_CSP, st1); if (convert_device_public(st1, la->out_csp, la->colmatrix, la->fullrange, la->ext, &planes, &dsrc, w, hgt, 0, 0, 1) < 0) return -1; }
            la->n_launch++;
        }
        if (!(borrowed || async)) XV_CUDA_OK(cudaEventRecord(ev_csp, st1));      // device source on the main stream
        if (borrowed || async) {
            XV_CUDA_OK(cudaEventRecord(ev_csp, la->st_io));
            if (iop) cudaEventRecord(la->io_ev[3], la->st_io);
            if (conv_pic) {
                // tight planes that follow each other on both sides (the usual conv_pic from
                // x264_picture_alloc / picture_layout) go back as ONE linear copy
                bool linear = true;
                size_t total = 0;
                for (int i = 0; i < geo_out.i_plane; i++) {
                    const size_t bytes = (size_t)geo_out.i_stride[i] * chroma_rows(out420, i);
                    linear = linear && conv_pic->i_stride[i] == geo_out.i_stride[i] && planes.i_stride[i] == geo_out.i_stride[i] &&
                             conv_pic->plane[i] == conv_pic->plane[0] + total && planes.plane[i] == planes.plane[0] + total;
                    total += bytes;
                }
                if (linear) XV_CUDA_OK(cudaMemcpyAsync(conv_pic->plane[0], planes.plane[0], total, cudaMemcpyDeviceToHost, la->st_io));
                else
                    for (int i = 0; i < geo_out.i_plane; i++)
                        XV_CUDA_OK(cudaMemcpy2DAsync(conv_pic->plane[i], conv_pic->i_stride[i], planes.plane[i], planes.i_stride[i], geo_out.i_stride[i], chroma_rows(out420, i),
                                                     cudaMemcpyDeviceToHost, la->st_io));
            }
            if (iop) cudaEventRecord(la->io_ev[4], la->st_io);
            XV_CUDA_OK(cudaEventRecord(la->ev_io, la->st_io));
        }
        return 0;
    };
    // ---- 1. borrowed buffers: stage 1 goes first, on the I/O stream ----
    if ((borrowed || async) && stage1() < 0) return -1;
    if (async) {
        // hand the frame to the session worker; wait for the borrowed buffers only
        {
            std::lock_guard<std::mutex> lk(la->mu);
            la->jobs.push_back({n, (long)slot, (long)did_fuse});
            la->cv.notify_all();
        }
        la->n_put++;
        la->t_put += now_s() - t_begin;
        if (borrowed) { const double t0 = now_s(); if (wait_event_caller(la, la->ev_io) < 0) return -1; la->t_io += now_s() - t0; }
        else if (wait_dev_src) { const double t0 = now_s(); if (wait_event(la, ev_csp) < 0) return -1; la->t_io += now_s() - t0; }
        if (borrowed && la->d_me_stats && la->io_ev[0]) {
            for (int k = 0; k < 4; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, la->io_ev[k], la->io_ev[k + 1]) == cudaSuccess) la->io_ms[k] += ms; }
            la->io_n++;
        }
        return (int)la->pubq.size();
    }

    // ---- 2. the decision that became due (deferred by decide_lag frames) ----
    double t_dec = 0;
    if (la->decide_lag > 0) {
        la->next.push_back(f);      // [x264] x264_lookahead_put_frame (frame n is outside the window being decided)
        const double t0 = now_s();
        while ((int)la->next.size() > la->slicetype_length + la->decide_lag)
            if (decide_and_shift(la) < 0) return -1;
        t_dec = now_s() - t0;
    }

    // ---- 3. stage 1 on the main stream (device-resident frame), or its hand-over to it ----
    if (!borrowed) { if (stage1() < 0) return -1; }
    else XV_CUDA_OK(cudaStreamWaitEvent(la->st, la->ev_csp, 0));

    // ---- 4./5. [x264] x264_adaptive_quant_frame, x264_frame_init_lowres (unless the fused front end produced them) ----
    if (!did_fuse && frame_prep(la, f, planes) < 0) return -1;
    XV_CUDA_OK(cudaEventRecord(la->ev_planes_free, la->st));
    la->planes_busy = true;

    f->ready = true;
    if (speculate_searches(la, f) < 0) return -1;
    const double t_mid = now_s();
    if (la->decide_lag == 0) {
        la->next.push_back(f);
        while ((int)la->next.size() > la->slicetype_length)
            if (decide_and_shift(la) < 0) return -1;
        t_dec = now_s() - t_mid;
    }
    la->t_decide += t_dec;
    la->t_put += (la->decide_lag == 0 ? t_mid - t_begin : now_s() - t_begin - t_dec);
    if (borrowed) { const double t0 = now_s(); if (wait_event_caller(la, la->ev_io) < 0) return -1; la->t_io += now_s() - t0; }
    else if (wait_dev_src) { const double t0 = now_s(); if (wait_event(la, ev_csp) < 0) return -1; la->t_io += now_s() - t0; }
    if (borrowed && la->d_me_stats && la->io_ev[0]) {
        for (int k = 0; k < 4; k++) { float ms = 0; if (cudaEventElapsedTime(&ms, la->io_ev[k], la->io_ev[k + 1]) == cudaSuccess) la->io_ms[k] += ms; }
        la->io_n++;
    }   // caller's buffers are borrowed for the call only
    la->n_put++;
    while (!la->outq.empty()) { la->pubq.push_back(la->outq.front()); la->outq.pop_front(); }
    return (int)la->pubq.size();
}

int x264vfw_cuda_la_flush(x264vfw_cuda_la *h)
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    // decisions deferred by decide_lag still see exactly the frames upstream would have had
    while ((int)la->next.size() > la->slicetype_length)
        if (decide_and_shift(la) < 0) return -1;
    la->flushing = true;
    while (!la->next.empty())
        if (decide_and_shift(la) < 0) { la->flushing = false; return -1; }
    la->flushing = false;
    while (!la->outq.empty()) { la->pubq.push_back(la->outq.front()); la->outq.pop_front(); }
    return (int)la->pubq.size();
}

int x264vfw_cuda_la_get_decision(x264vfw_cuda_la *h, x264vfw_cuda_la_decision *d, float *qp_offset, float *qp_offset_aq)
{
    La *la = (La *)h;
    if (!la || !d) return -1;
    if (la->pubq.empty()) return 0;
    Decision &q = la->pubq.front();
    *d = q.d;
    if (qp_offset) memcpy(qp_offset, q.h_qp, la->g.mb_count * sizeof(float));
    if (qp_offset_aq) memcpy(qp_offset_aq, q.h_qp_aq, la->g.mb_count * sizeof(float));
    { std::lock_guard<std::mutex> lk(la->qp_mu); la->qp_free.push_back(q.h_qp); }
    la->pubq.pop_front();
    return 1;
}

int x264vfw_cuda_la_frame_cost(x264vfw_cuda_la *h, int p0, int p1, int b)
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    if (p0 < 0 || p1 >= (int)la->by_index.size() || b < p0 || b > p1 || b - p0 > la->p.bframes + 1 || p1 - b > la->p.bframes + 1) { set_error("frame_cost: bad indices"); return -1; }
    for (int i = p0; i <= p1; i++) if (!la->by_index[i]) { set_error("frame %d was recycled (open with keep_frames)", i); return -1; }
    return frame_cost(la, la->by_index.data(), p0, p1, b, true);
}

int x264vfw_cuda_la_mbtree(x264vfw_cuda_la *h, const int *frame_idx, const int *types, int num_frames, int b_intra)
{
    La *la = (La *)h;
    if (!la || num_frames > LMAX) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    Frame *frames[LMAX + 3];
    for (int i = 0; i <= num_frames; i++) {
        if (frame_idx[i] < 0 || frame_idx[i] >= (int)la->by_index.size() || !la->by_index[frame_idx[i]]) { set_error("mbtree: bad frame"); return -1; }
        frames[i] = la->by_index[frame_idx[i]];
        frames[i]->i_type = types[i];
        frames[i]->f_duration = (float)((double)la->p.fps_den / la->p.fps_num);
    }
    if (macroblock_tree(la, frames, num_frames, b_intra) < 0) return -1;
    return la_sync(la);
}

int64_t x264vfw_cuda_la_read(x264vfw_cuda_la *h, int frame, int what, int a, int b, void *dst, size_t cap)
{
    La *la = (La *)h;
    if (!la || !dst) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    if (la_sync(la) < 0) return -1;
    for (int e = 1; e <= la->me_side; e++) XV_CUDA_OK(cudaStreamSynchronize(la->st_me[e]));
    const int n = la->g.mb_count, B = la->p.bframes;
    if (what == X264VFW_CUDA_LA_CONV_PLANES) {
        if (cap < la->d_planes_bytes) { set_error("read: buffer too small"); return -1; }
        const uint8_t *last = la->worker.joinable() ? la->d_planes_ring[(la->n_put + la->io_depth - 1) % la->io_depth] : la->d_planes;
        if (la->st_io) XV_CUDA_OK(cudaStreamSynchronize(la->st_io));
        XV_CUDA_OK(cudaMemcpy(dst, last, la->d_planes_bytes, cudaMemcpyDeviceToHost));
        return (int64_t)la->d_planes_bytes;
    }
    if (frame < 0 || frame >= (int)la->by_index.size() || !la->by_index[frame]) { set_error("read: frame %d not resident", frame); return -1; }
    Frame *f = la->by_index[frame];
    const void *src = nullptr; size_t bytes = 0; bool host = false;
    int tmp[4]; unsigned long long st[6];
    switch (what) {
    case X264VFW_CUDA_LA_LOWRES: src = f->lowres; bytes = (size_t)4 * la->g.lplane; break;
    case X264VFW_CUDA_LA_INTRA_COST: src = f->intra_cost; bytes = n * 2; break;
    case X264VFW_CUDA_LA_INV_QSCALE: src = f->inv_qscale; bytes = n * 2; break;
    case X264VFW_CUDA_LA_PROPAGATE: src = f->propagate; bytes = n * 4; break;
    case X264VFW_CUDA_LA_QP_OFFSET: src = f->qp_offset; bytes = n * 4; break;
    case X264VFW_CUDA_LA_QP_OFFSET_AQ: src = f->qp_offset_aq; bytes = n * 4; break;
    case X264VFW_CUDA_LA_MVS: if (a < 0 || a > 1 || b < 1 || b > B + 1) return -1; src = f->mvs[a][b - 1]; bytes = n * 4; break;
    case X264VFW_CUDA_LA_MV_COSTS: if (a < 0 || a > 1 || b < 1 || b > B + 1) return -1; src = f->mv_costs[a][b - 1]; bytes = n * 4; break;
    case X264VFW_CUDA_LA_LOWRES_COSTS: if (a < 0 || a > B + 1 || b < 0 || b > B + 1) return -1; src = lc_ptr(la, f, a, b); bytes = n * 2; break;
    case X264VFW_CUDA_LA_ROW_SATDS: if (a < 0 || a > B + 1 || b < 0 || b > B + 1) return -1; src = rs_ptr(la, f, a, b); bytes = la->g.mb_h * 4; break;
    case X264VFW_CUDA_LA_COST_EST:
        if (a < 0 || a > B + 1 || b < 0 || b > B + 1) return -1;
        tmp[0] = f->cost_est[a][b]; tmp[1] = f->cost_est_aq[a][b]; tmp[2] = f->intra_mbs[a]; src = tmp; bytes = 12; host = true; break;
    case X264VFW_CUDA_LA_PIXEL_STATS:
        if (ensure_stats(la, f) < 0) return -1;
        for (int i = 0; i < 3; i++) { st[i] = f->pixel_sum[i]; st[3 + i] = f->pixel_ssd[i]; }
        src = st; bytes = 48; host = true; break;
    case X264VFW_CUDA_LA_WEIGHT:
        tmp[0] = f->weight.scale; tmp[1] = f->weight.denom; tmp[2] = f->weight.offset; tmp[3] = f->weight.on; src = tmp; bytes = 16; host = true; break;
    default: set_error("read: unknown selector %d", what); return -1;
    }
    if (cap < bytes) { set_error("read: buffer too small"); return -1; }
    if (host) memcpy(dst, src, bytes);
    else XV_CUDA_OK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return (int64_t)bytes;
}

int x264vfw_cuda_la_profile(x264vfw_cuda_la *h, int enable, double ms[16], uint64_t count[16])
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    if (la_sync(la) < 0) return -1;
    for (int e = 1; e <= la->me_side; e++) XV_CUDA_OK(cudaStreamSynchronize(la->st_me[e]));
    prof_resolve(la);
    for (int i = 0; i < 16; i++) { if (ms) ms[i] = i < K_N ? la->prof.ms[i] : 0; if (count) count[i] = i < K_N ? la->prof.n[i] : 0; }
    if (enable >= 0) {
        la->prof.on = enable != 0;
        for (int i = 0; i < K_N; i++) { la->prof.ms[i] = 0; la->prof.n[i] = 0; }
        // the search kernels count their work while profiling is on (x264vfw_cuda_la_stats)
        if (enable && !la->d_me_stats) {
            XV_CUDA_OK(cudaMalloc((void **)&la->d_me_stats, 8 * sizeof(int)));
        }
        if (enable) { XV_CUDA_OK(cudaMemset(la->d_me_stats, 0, 8 * sizeof(int))); la->n_tree_steps = la->n_tree_walks = 0; }
    }
    return 0;
}

// [x264] x264_weights_analyse(h, fenc, ref, 0) on a session that kept its frames (la_weights_full.cu)
int x264vfw_cuda_la_weights_analyse(x264vfw_cuda_la *h, int fenc_i, int ref_i, const uint8_t *fenc_uv, const uint8_t *ref_uv,
                                    int uv_stride, int32_t out[3][4], float *cost_delta)
{
    La *la = (La *)h;
    if (!la || !fenc_uv || !ref_uv || !out) { set_error("null argument"); return -1; }
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    if (la->p.chroma_format != 1) { set_error("weights_analyse: 4:2:0 only"); return -1; }
    const int n = (int)la->by_index.size();
    if (fenc_i < 0 || fenc_i >= n || ref_i < 0 || ref_i >= fenc_i || fenc_i - ref_i > la->p.bframes + 1) { set_error("weights_analyse: bad indices"); return -1; }
    Frame *fenc = la->by_index[fenc_i], *ref = la->by_index[ref_i];
    if (!fenc || !ref) { set_error("weights_analyse: frame was recycled (open with keep_frames)"); return -1; }
    if (la_sync(la) < 0) return -1;
    for (int e = 1; e <= la->me_side; e++) XV_CUDA_OK(cudaStreamSynchronize(la->st_me[e]));
    if (ensure_stats(la, fenc) < 0 || ensure_stats(la, ref) < 0) return -1;
    if (!fenc->b_intra_calculated) {                         // "if( !fenc->b_intra_calculated ) slicetype_frame_cost( h, &a, &fenc, 0, 0, 0 )"
        Frame *one[1] = {fenc};
        if (frame_cost(la, one, 0, 0, 0, true) < 0) return -1;
    }
    const int dist = fenc_i - ref_i - 1;
    x264vfw_cuda_weights_in in;
    memset(&in, 0, sizeof(in));
    in.width = la->g.width; in.height = la->g.height;
    in.fenc_lowres = fenc->lowres; in.ref_lowres = ref->lowres;
    in.lowres_mvs = fenc->searched[0][dist] ? (const int16_t *)fenc->mvs[0][dist] : nullptr;
    in.intra_cost = fenc->intra_cost;
    in.fenc_uv = fenc_uv; in.ref_uv = ref_uv; in.uv_stride = uv_stride;
    for (int i = 0; i < 3; i++) {
        in.fenc_sum[i] = fenc->pixel_sum[i]; in.fenc_ssd[i] = fenc->pixel_ssd[i];
        in.ref_sum[i] = ref->pixel_sum[i]; in.ref_ssd[i] = ref->pixel_ssd[i];
    }
    in.subme = la->p.subme; in.weightp = la->p.weightp;
    unsigned *d_res = nullptr, *h_res = nullptr;
    XV_CUDA_OK(cudaMalloc((void **)&d_res, sizeof(unsigned) * 3 * 48));
    if (cudaMallocHost((void **)&h_res, sizeof(unsigned) * 3 * 48) != cudaSuccess) { cudaFree(d_res); set_error("pinned allocation failed"); return -1; }
    memset(h_res, 0, sizeof(unsigned) * 3 * 48);
    const int rc = weights_analyse_full(la->st, la->g, &in, out, cost_delta, d_res, h_res);
    cudaFree(d_res); cudaFreeHost(h_res);
    return rc;
}

int x264vfw_cuda_la_stats(x264vfw_cuda_la *h, uint64_t out[16])
{
    La *la = (La *)h;
    if (!la || !out) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (worker_join(la) < 0) return -1;
    if (la_sync(la) < 0) return -1;
    for (int e = 1; e <= la->me_side; e++) XV_CUDA_OK(cudaStreamSynchronize(la->st_me[e]));
    for (int i = 0; i < 16; i++) out[i] = 0;
    if (la->d_me_stats) {
        int v[8];
        XV_CUDA_OK(cudaMemcpy(v, la->d_me_stats, sizeof(v), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 8; i++) out[i] = (uint64_t)(unsigned)v[i];
    }
    out[8] = la->n_tree_steps; out[9] = la->n_tree_walks;
    out[10] = la->n_spec_jobs; out[11] = la->n_ondemand_jobs; out[12] = la->n_ondemand;
    for (int l = 0; l < 2; l++) for (int d = 0; d <= BMAX; d++) out[13] += la->n_logical[l][d];
    out[14] = (uint64_t)la->n_input;
    return 0;
}

// ---- [x264] encoder/slicetype-cl.c hook names on a cost-engine session --------------------------
static int hook_frames(La *la, int a, int b)
{
    if (!la->keep_frames) { set_error("opencl-style hooks need a session opened with keep_frames"); return -1; }
    for (int i = a; i <= b; i++)
        if (i < 0 || i >= (int)la->by_index.size() || !la->by_index[i] || !la->by_index[i]->ready) { set_error("hook: frame %d is not resident", i); return -1; }
    return 0;
}

// enqueue the unweighted search of frame b towards `ref` unless its result exists already (mp collects a batch)
static int hook_add_search(La *la, MeParams &mp, int b, int ref, int list)
{
    const int dist = list ? ref - b : b - ref;
    if (dist < 1 || dist > la->p.bframes + 1) { set_error("hook: distance %d outside 1..%d", dist, la->p.bframes + 1); return -1; }
    Frame *fenc = la->by_index[b], *fref = la->by_index[ref];
    if (fenc->spec[list][dist - 1] || fenc->searched[list][dist - 1]) return 0;
    me_add_job(la, mp, 0, fenc, fref, list, dist, nullptr);
    fenc->spec[list][dist - 1] = true; fenc->spec_eng[list][dist - 1] = 0;
    if (mp.njobs == XV_ME_MAX_JOBS) { if (me_launch(la, mp, 0) < 0) return -1; me_params_init(la, mp); }
    return 0;
}

int x264vfw_cuda_opencl_lowres_init(x264vfw_cuda_la *h, const x264vfw_cuda_image_t *src, int src_on_device)
{
    La *la = (La *)h;
    if (!la || !la->keep_frames) { set_error("opencl-style hooks need a session opened with keep_frames"); return -1; }
    if (x264vfw_cuda_la_put_frame(h, src, src_on_device, nullptr) < 0) return -1;
    return la->n_input - 1;
}

int x264vfw_cuda_opencl_motionsearch(x264vfw_cuda_la *h, int b, int ref, int b_islist1)
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (hook_frames(la, b < ref ? b : ref, b < ref ? ref : b) < 0) return -1;
    MeParams mp;
    me_params_init(la, mp);
    if (hook_add_search(la, mp, b, ref, b_islist1 != 0) < 0) return -1;
    return me_launch(la, mp, 0);
}

int x264vfw_cuda_opencl_finalize_cost(x264vfw_cuda_la *h, int p0, int p1, int b, int cost_out[3])
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (p0 > b || b > p1 || hook_frames(la, p0, p1) < 0) { if (p0 > b || b > p1) set_error("finalize_cost: need p0 <= b <= p1"); return -1; }
    if (b - p0 > la->p.bframes + 1 || p1 - b > la->p.bframes + 1) { set_error("finalize_cost: distance too large"); return -1; }
    const int score = frame_cost(la, la->by_index.data(), p0, p1, b, true);
    if (score < 0) return -1;
    Frame *f = la->by_index[b];
    if (cost_out) { cost_out[0] = f->cost_est[b - p0][p1 - b]; cost_out[1] = f->cost_est_aq[b - p0][p1 - b]; cost_out[2] = f->intra_mbs[b - p0]; }
    return score;
}

int x264vfw_cuda_opencl_flush(x264vfw_cuda_la *h)
{
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (la_sync(la) < 0) return -1;
    for (int e = 1; e <= la->me_side; e++) XV_CUDA_OK(cudaStreamSynchronize(la->st_me[e]));
    return 0;
}

int x264vfw_cuda_opencl_slicetype_prep(x264vfw_cuda_la *h, int first, int num_frames)
{
    // upstream precomputes, for trellis B-adapt, the searches of every frame of the window towards the frames up to
    // bframes away in both directions; here that is one or a few batched launches for any b-adapt
    La *la = (La *)h;
    if (!la) return -1;
    XV_CUDA_OK(cudaSetDevice(la->device));
    if (hook_frames(la, first, first + num_frames) < 0) return -1;
    MeParams mp;
    me_params_init(la, mp);
    for (int b = first; b <= first + num_frames; b++)
        for (int j = 1; j <= la->p.bframes; j++) {
            if (b - j >= first && hook_add_search(la, mp, b, b - j, 0) < 0) return -1;
            if (b + j <= first + num_frames && hook_add_search(la, mp, b, b + j, 1) < 0) return -1;
        }
    return me_launch(la, mp, 0);
}

int x264vfw_cuda_opencl_slicetype_end(x264vfw_cuda_la *h) { return x264vfw_cuda_opencl_flush(h); }

void x264vfw_cuda_la_counters(x264vfw_cuda_la *h, uint64_t out[8])
{
    La *la = (La *)h;
    if (!la) return;
    out[0] = la->n_frame_cost; out[1] = la->n_mb_search; out[2] = la->n_launch; out[3] = la->n_sync;
    out[4] = (uint64_t)(la->t_put * 1e6); out[5] = (uint64_t)(la->t_decide * 1e6); out[6] = (uint64_t)(la->t_sync * 1e6); out[7] = (uint64_t)la->n_input;
}

} // extern "C"
