// Host side of the C ABI (include/oi_b200.h): device buffers, the lockstep batch scheduler and
// the entry points.  Replaces the reference's per-rank cell loop (GPR_CS2S3.py:258-261, :317-319):
// instead of one cell at a time per MPI rank, every lockstep iteration evaluates SMLII (or the
// posterior) for ALL active cells at once, each cell at the point its own optimiser asked for.
#include <cuda_runtime.h>
#include <algorithm>
#include <deque>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include "../../include/oi_b200.h"
#include "oi_types.h"
#include "oi_launch.h"
#include "cg_scipy.h"
#include "lbfgs_fast.h"
#include "oi_shared_queue.h"

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(e_ == cudaErrorMemoryAllocation ? OI_ERR_NOMEM : OI_ERR_CUDA,              \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
    } while (0)

struct oi_handle {
    int device = 0;
    cudaStream_t st = nullptr, own_st = nullptr;
    // observations
    double *ox = nullptr, *oy = nullptr, *ot = nullptr, *oz = nullptr;
    int64_t n_obs = 0, obs_cap = 0;
    std::vector<double> h_t;            // host copy of t: the day window's index ranges are found on the host
    double t_lo = -INFINITY, t_hi = INFINITY, t_shift = 0.0;      // day window of the gather (oi_set_time_window)
    // cells
    double* X = nullptr;
    int64_t n_cells = 0, cell_cap = 0;
    // neighbours (CSR)
    int* counts = nullptr; long long* offsets = nullptr; int* indices = nullptr;
    long long total = 0, idx_cap = 0;
    bool have_nbr = false;
    std::vector<int> h_counts;
    std::vector<long long> h_offsets;
    // packed per-cell coordinates
    double *px = nullptr, *py = nullptr, *pt = nullptr, *pr = nullptr;
    // per-cell state
    OiCellArrays ca{};
    // lockstep batch
    char* arena = nullptr; size_t arena_bytes = 0;
    OiSlot* d_slots = nullptr; OiSlot* h_slots = nullptr;
    int *d_slot_phase = nullptr, *h_slot_phase = nullptr, *d_fail = nullptr;
    int slot_cap = 0;
    cudaEvent_t ev[10]{};
    std::vector<struct OiGroup*> groups;
    oi_stats stats{};
    bool have_results = false;
    double* d_dbg = nullptr; int* d_dbg_count = nullptr; int dbg_cell = -1, dbg_cap = 0;
    OiSharedQueue queue;                 // work list shared with the other GPU processes of the box (oi_set_shared_queue)
    std::vector<uint8_t> owned;          // cells this handle computed in the last oi_run
};

struct OiGroup;
static void free_groups(oi_handle* h);
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t slot_bytes(int n) {
    size_t N = (n + OI_NB - 1) / OI_NB, npad = N * OI_NB;
    size_t b = align_up(npad * npad * 8, 256);
    b += align_up(N * OI_TILE * 8, 256);
    b += align_up(3 * npad * 8, 256);
    b += align_up((N + 8 + 5 * N * (N + 1) / 2) * 8, 256);
    b += align_up(N * (N + 1) / 2 * 2 * OI_TILE * 8, 256);      // Q and exp(-Q) tiles
    return b;
}

extern "C" int oi_version(void) { return 120; }
extern "C" int oi_sizeof_params(void) { return (int)sizeof(oi_params); }
extern "C" int oi_sizeof_stats(void) { return (int)sizeof(oi_stats); }
extern "C" const char* oi_last_error(void) { return g_err.c_str(); }

extern "C" int oi_create(int device, oi_handle** out) {
    if (!out) return fail(OI_ERR_ARG, "oi_create: out is NULL");
    // (The stream groups like their own hardware queues: CUDA_DEVICE_MAX_CONNECTIONS=32 in the host's environment before
    // CUDA initialises is worth +2 %; the library does not touch the environment, see INTEGRATION.md.)
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(OI_ERR_CUDA, std::string("oi_create: no CUDA device (") + cudaGetErrorString(e) +
                                     "); this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(OI_ERR_ARG, "oi_create: bad device index");
    CK(cudaSetDevice(device));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
        return fail(OI_ERR_CUDA, "oi_create: device is sm_" + std::to_string(p.major * 10 + p.minor) +
                                     ", kernels are built for sm_100a only");
    // opt-in shared-memory sizes are per device context: set them for THIS device (one handle per GPU, several per process)
    if (int ae = oi_set_kernel_attributes())
        return fail(OI_ERR_CUDA, std::string("oi_create: cudaFuncSetAttribute: ") + cudaGetErrorString((cudaError_t)ae));
    oi_handle* h = new oi_handle();
    h->device = device;
    cudaError_t ce = cudaStreamCreate(&h->own_st);
    for (auto& ev : h->ev) if (ce == cudaSuccess) ce = cudaEventCreate(&ev);
    if (ce != cudaSuccess) {
        oi_destroy(h);
        return fail(ce == cudaErrorMemoryAllocation ? OI_ERR_NOMEM : OI_ERR_CUDA, std::string("oi_create: ") + cudaGetErrorString(ce));
    }
    h->st = h->own_st;
    *out = h;
    return OI_OK;
}

extern "C" int oi_set_stream(oi_handle* h, void* cuda_stream) {
    if (!h) return fail(OI_ERR_ARG, "oi_set_stream: NULL handle");
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->st));
    h->st = cuda_stream ? (cudaStream_t)cuda_stream : h->own_st;
    return OI_OK;
}

static void free_cells(oi_handle* h) {
    cudaFree(h->X); cudaFree(h->counts); cudaFree(h->offsets);
    cudaFree(h->ca.hyp); cudaFree(h->ca.phase); cudaFree(h->ca.out); cudaFree(h->ca.nfev); cudaFree(h->ca.status);
    cudaFree(h->ca.evf); cudaFree(h->ca.evg); cudaFree(h->ca.cg);
    h->X = nullptr; h->counts = nullptr; h->offsets = nullptr; h->ca = OiCellArrays{}; h->cell_cap = 0;
}

extern "C" void oi_destroy(oi_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->ox); cudaFree(h->oy); cudaFree(h->ot); cudaFree(h->oz);
    free_cells(h);
    cudaFree(h->indices); cudaFree(h->px); cudaFree(h->py); cudaFree(h->pt); cudaFree(h->pr);
    cudaFree(h->d_dbg); cudaFree(h->d_dbg_count);
    h->queue.detach();
    cudaFree(h->arena); cudaFree(h->d_slots); cudaFree(h->d_slot_phase); cudaFree(h->d_fail);
    cudaFreeHost(h->h_slots); cudaFreeHost(h->h_slot_phase);
    free_groups(h);
    for (auto& ev : h->ev) if (ev) cudaEventDestroy(ev);
    if (h->own_st) cudaStreamDestroy(h->own_st);
    delete h;
}

extern "C" int oi_set_observations(oi_handle* h, const double* x, const double* y, const double* t, const double* z,
                                   int64_t n_obs) {
    if (!h || !x || !y || !t || !z || n_obs <= 0 || n_obs > 0x7fffffff) return fail(OI_ERR_ARG, "oi_set_observations: bad argument");
    CK(cudaSetDevice(h->device));
    if (n_obs > h->obs_cap) {
        cudaFree(h->ox); cudaFree(h->oy); cudaFree(h->ot); cudaFree(h->oz);
        CK(cudaMalloc(&h->ox, n_obs * 8)); CK(cudaMalloc(&h->oy, n_obs * 8));
        CK(cudaMalloc(&h->ot, n_obs * 8)); CK(cudaMalloc(&h->oz, n_obs * 8));
        h->obs_cap = n_obs;
    }
    h->n_obs = n_obs;
    h->h_t.assign(t, t + n_obs);
    CK(cudaMemcpyAsync(h->ox, x, n_obs * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->oy, y, n_obs * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->ot, t, n_obs * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->oz, z, n_obs * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->have_nbr = false; h->have_results = false;
    return OI_OK;
}

extern "C" int oi_set_time_window(oi_handle* h, double t_lo, double t_hi) {
    if (!h || !(t_lo <= t_hi)) return fail(OI_ERR_ARG, "oi_set_time_window: need t_lo <= t_hi");
    h->t_lo = t_lo; h->t_hi = t_hi; h->t_shift = std::isfinite(t_lo) ? t_lo : 0.0;
    h->have_nbr = false; h->have_results = false;
    return OI_OK;
}

extern "C" int oi_set_cells(oi_handle* h, const double* X, int64_t n_cells) {
    if (!h || !X || n_cells <= 0 || n_cells > 0x3fffffff) return fail(OI_ERR_ARG, "oi_set_cells: bad argument");
    CK(cudaSetDevice(h->device));
    if (n_cells > h->cell_cap) {
        free_cells(h);
        size_t c = (size_t)n_cells;
        CK(cudaMalloc(&h->X, c * 16)); CK(cudaMalloc(&h->counts, c * 4)); CK(cudaMalloc(&h->offsets, (c + 1) * 8));
        CK(cudaMalloc(&h->ca.hyp, c * 40)); CK(cudaMalloc(&h->ca.phase, c * 4)); CK(cudaMalloc(&h->ca.out, c * 64));
        CK(cudaMalloc(&h->ca.nfev, c * 4)); CK(cudaMalloc(&h->ca.status, c * 4)); CK(cudaMalloc(&h->ca.evf, c * 8));
        CK(cudaMalloc(&h->ca.evg, c * 8 * OI_MAXH)); CK(cudaMalloc(&h->ca.cg, c * OI_OPT_STATE_BYTES));
        h->cell_cap = n_cells;
    }
    h->ca.X = h->X;
    h->n_cells = n_cells;
    CK(cudaMemcpyAsync(h->X, X, n_cells * 16, cudaMemcpyHostToDevice, h->st));
    CK(cudaStreamSynchronize(h->st));
    h->have_nbr = false; h->have_results = false;
    return OI_OK;
}

extern "C" int oi_gather_neighbours(oi_handle* h, double radius_m, int32_t* counts_out) {
    if (!h || !(radius_m >= 0)) return fail(OI_ERR_ARG, "oi_gather_neighbours: bad argument");
    if (h->n_obs <= 0 || h->n_cells <= 0) return fail(OI_ERR_STATE, "oi_gather_neighbours: set observations and cells first");
    CK(cudaSetDevice(h->device));
    const double r2 = radius_m * radius_m;
    const int nc = (int)h->n_cells, no = (int)h->n_obs;
    // Maximal runs of observations inside the day window (a resident season is stream-major / day-major: one run per
    // stream; 20x less to scan than the season for a 180-day season).  More than OI_MAX_RANGES runs: scan everything.
    OiRanges rg; rg.n = 1; rg.lo[0] = 0; rg.hi[0] = no;
    if (std::isfinite(h->t_lo) || std::isfinite(h->t_hi)) {
        OiRanges w; w.n = 0;
        bool ok = true;
        for (int i = 0; i < no && ok;) {
            if (!(h->h_t[i] >= h->t_lo && h->h_t[i] <= h->t_hi)) { i++; continue; }
            int j = i;
            while (j < no && h->h_t[j] >= h->t_lo && h->h_t[j] <= h->t_hi) j++;
            if (w.n == OI_MAX_RANGES) { ok = false; break; }
            w.lo[w.n] = i; w.hi[w.n] = j; w.n++;
            i = j;
        }
        if (ok) rg = w;
    }
    CK(cudaEventRecord(h->ev[8], h->st));
    oi_launch_count(h->ox, h->oy, h->ot, no, h->X, nc, r2, h->t_lo, h->t_hi, rg, h->counts, h->st);
    oi_launch_scan(h->counts, nc, h->offsets, h->st);
    CK(cudaGetLastError());
    h->h_counts.resize(nc);
    h->h_offsets.resize((size_t)nc + 1);
    CK(cudaMemcpyAsync(h->h_counts.data(), h->counts, (size_t)nc * 4, cudaMemcpyDeviceToHost, h->st));
    CK(cudaMemcpyAsync(h->h_offsets.data(), h->offsets, ((size_t)nc + 1) * 8, cudaMemcpyDeviceToHost, h->st));   // the device scan's result
    CK(cudaStreamSynchronize(h->st));
    const long long total = h->h_offsets[(size_t)nc];
    if (total > h->idx_cap) {
        cudaFree(h->indices); cudaFree(h->px); cudaFree(h->py); cudaFree(h->pt); cudaFree(h->pr);
        size_t c = (size_t)std::max<long long>(total, 1);
        CK(cudaMalloc(&h->indices, c * 4));
        CK(cudaMalloc(&h->px, c * 8)); CK(cudaMalloc(&h->py, c * 8)); CK(cudaMalloc(&h->pt, c * 8)); CK(cudaMalloc(&h->pr, c * 8));
        h->idx_cap = total;
    }
    h->total = total;
    if (total > 0) oi_launch_fill(h->ox, h->oy, h->ot, no, h->X, nc, r2, h->t_lo, h->t_hi, rg, h->offsets, h->indices, h->st);
    CK(cudaGetLastError());
    CK(cudaEventRecord(h->ev[9], h->st));
    CK(cudaStreamSynchronize(h->st));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]);
    h->stats.ms_gather = ms;
    h->stats.sum_n = total;
    h->have_nbr = true; h->have_results = false;
    if (counts_out) std::memcpy(counts_out, h->h_counts.data(), (size_t)nc * 4);
    return OI_OK;
}

extern "C" int oi_get_neighbours(oi_handle* h, int64_t* offsets, int32_t* indices) {
    if (!h) return fail(OI_ERR_ARG, "oi_get_neighbours: NULL handle");
    if (!h->have_nbr) return fail(OI_ERR_STATE, "oi_get_neighbours: call oi_gather_neighbours first");
    CK(cudaSetDevice(h->device));
    if (offsets) CK(cudaMemcpy(offsets, h->offsets, (size_t)(h->n_cells + 1) * 8, cudaMemcpyDeviceToHost));
    if (indices && h->total > 0) CK(cudaMemcpy(indices, h->indices, (size_t)h->total * 4, cudaMemcpyDeviceToHost));
    return OI_OK;
}

// ------------------------------------------------------------------------------------------
// the lockstep batch scheduler
//
// The active cells are split into G independent GROUPS, each with its own CUDA stream, slot table and
// share of the scratch arena.  A group runs lockstep iterations (build -> Cholesky -> ... -> finalize ->
// D2H of the new phases) on its stream; the host services the groups round-robin, so while it retires
// and refills group g the other G-1 groups' kernel chains are already queued.  Kernels of different
// groups overlap on the device: the low-parallelism launches of one group (last block columns of the
// factorisation, large block distances of the inverse, the optimiser's tail with a handful of cells)
// fill with tiles of the others.  A cell's numbers do not depend on its group or its batch
// (fixed-order reductions), so results are bit-identical for every G.
// ------------------------------------------------------------------------------------------
struct OiGroup {
    cudaStream_t st = nullptr;
    cudaEvent_t ev[8]{};            // family boundaries of the iteration in flight
    cudaEvent_t done = nullptr;
    OiSlot *d_slots = nullptr, *h_slots = nullptr;
    int *d_slot_phase = nullptr, *h_slot_phase = nullptr, *d_fail = nullptr;
    char* arena = nullptr; size_t arena_bytes = 0, used = 0;
    long long tiles = 0;            // sum of N(N+1)/2 over the active cells (work admitted to the group)
    bool high_prio = false;
    // CUDA graph of one iteration, replayed while the (small) batch keeps its composition
    cudaGraphExec_t gexec = nullptr; std::vector<int> gcells, last_cells; int same_count = 0; bool graph_mode = false;
    int slot_cap = 0;
    std::vector<int> active, cnt_gt;
    bool in_flight = false;
    int A = 0, Nmax = 0;
    double fl = 0, flf = 0, flf_chol = 0, flf_fit = 0; int64_t nev = 0;
};

static int ensure_batch_buffers(oi_handle* h, size_t want_arena, int want_slots, int G) {
    if (want_arena > h->arena_bytes) {
        cudaFree(h->arena); h->arena = nullptr; h->arena_bytes = 0;
        CK(cudaMalloc(&h->arena, want_arena));
        h->arena_bytes = want_arena;
    }
    if (want_slots > h->slot_cap) {
        cudaFree(h->d_slots); cudaFree(h->d_slot_phase); cudaFree(h->d_fail);
        cudaFreeHost(h->h_slots); cudaFreeHost(h->h_slot_phase);
        CK(cudaMalloc(&h->d_slots, (size_t)want_slots * sizeof(OiSlot)));
        CK(cudaMalloc(&h->d_slot_phase, (size_t)want_slots * 4));
        CK(cudaMalloc(&h->d_fail, (size_t)want_slots * 4));
        CK(cudaMallocHost(&h->h_slots, (size_t)want_slots * sizeof(OiSlot)));
        CK(cudaMallocHost(&h->h_slot_phase, (size_t)want_slots * 4));
        h->slot_cap = want_slots;
    }
    while ((int)h->groups.size() < G) {
        OiGroup* g = new OiGroup();
        h->groups.push_back(g);                       // owned by the handle from here on (freed in free_groups)
        CK(cudaStreamCreateWithFlags(&g->st, cudaStreamNonBlocking));
        for (auto& e : g->ev) CK(cudaEventCreate(&e));
        CK(cudaEventCreateWithFlags(&g->done, cudaEventDisableTiming));
    }
    return OI_OK;
}

static void free_groups(oi_handle* h) {
    for (OiGroup* g : h->groups) {
        for (auto& e : g->ev) if (e) cudaEventDestroy(e);
        if (g->done) cudaEventDestroy(g->done);
        if (g->gexec) cudaGraphExecDestroy(g->gexec);
        if (g->st) cudaStreamDestroy(g->st);
        delete g;
    }
    h->groups.clear();
}

struct LockstepRun {
    oi_handle* h; std::vector<int>& h_phase; const OiRunConst& rc; double t_pred;
    // pending cells sorted by descending size, admitted from BOTH ends: most bulk groups take the largest cells first
    // (cost-sorted batches), the last n_small bulk groups take the smallest first.  Evaluation counts above ~500 occur
    // almost only in small cells (measured on the 25 km day: 10 % of the cells with n < 400, 2.5 % for n in 400...700,
    // < 0.2 % above, none above 1300), and small cells are cheap: started at t = 0 next to the big ones, their long
    // optimiser runs finish underneath the bulk work instead of forming a tail of a few latency-bound cells after it.
    std::vector<int> pending; size_t next = 0, back = 0;
    int n_small = 0;
    bool takes_small(int gi) const { return !is_express(gi) && gi >= G - n_express - n_small; }
    // With a shared queue (several GPU processes of one box) the two cursors live in shared memory and every rank claims
    // from the same list (oi_shared_queue.h); without, they are the local next/back.
    OiSharedQueue* q = nullptr;
    long peek(bool small_end) const {
        if (q) return small_end ? q->peek_back() : q->peek_front();
        return next < back ? (long)(small_end ? back - 1 : next) : -1;
    }
    bool claim(bool small_end, long idx) {
        if (q) return small_end ? q->claim_back(idx) : q->claim_front(idx);
        if (small_end) back--; else next++;
        return true;
    }
    bool list_empty() const { return peek(false) < 0; }
    OiPacked pk; FILE* trace = nullptr;
    double ms_factor = 0;
    // express lanes: the last n_express groups only take cells that already spent express_after iterations in a
    // bulk group.  Such cells (line searches that keep failing, maxiter) need up to ~2000 evaluations; in a small
    // batch an iteration takes ~1 ms instead of the ~10 ms of a full bulk batch, so their long dependent chains
    // finish underneath the bulk work instead of forming a serial tail after it.
    int G = 1, n_express = 0, express_after = 0, express_cap = 0;
    std::deque<int> express_pending;
    std::vector<int> iters;

    long long tile_budget = 0;      // per bulk group
    bool is_express(int gi) const { return gi >= G - n_express; }
    long long cell_tiles(int c) const { long long N = (h->h_counts[c] + OI_NB - 1) / OI_NB; return N * (N + 1) / 2; }

    int graph_max_A = 0;            // batches up to this many cells replay a captured graph

    // the kernel chain of one lockstep iteration of group g (timed: with the family boundary events)
    int launch_chain(OiGroup& g, int A, int Nmax, const int* cg, cudaStream_t st, bool timed) {
        if (timed) CK(cudaEventRecord(g.ev[0], st));
        oi_launch_build(g.d_slots, A, Nmax, cg, h->ca, pk, st);
        if (timed) CK(cudaEventRecord(g.ev[1], st));
        for (int k = 0; k < Nmax; k++) {
            oi_launch_chol_update(g.d_slots, A, Nmax, cg, k, st);
            oi_launch_chol_panel(g.d_slots, A, Nmax, cg, k, st);      // panel of column k + row scaling of row k
        }
        if (timed) CK(cudaEventRecord(g.ev[2], st));
        oi_launch_fwd(g.d_slots, A, h->ca, pk, t_pred, st);
        if (timed) CK(cudaEventRecord(g.ev[3], st));
        for (int d = 1; d < Nmax; d++) oi_launch_trtri(g.d_slots, A, Nmax, cg, d, h->ca.phase, st);
        if (timed) CK(cudaEventRecord(g.ev[4], st));
        oi_launch_alpha(g.d_slots, A, Nmax, h->ca.phase, st);
        if (timed) CK(cudaEventRecord(g.ev[5], st));
        oi_launch_lauum_trace(g.d_slots, A, Nmax, cg, h->ca, pk, st);
        if (timed) CK(cudaEventRecord(g.ev[6], st));
        oi_launch_finalize(g.d_slots, A, h->ca, rc, g.d_slot_phase, st);
        if (timed) CK(cudaEventRecord(g.ev[7], st));
        return OI_OK;
    }

    // admit pending cells (largest first) into group g, pack its slot table and queue one iteration
    int issue(OiGroup& g, int gi) {
        const int cap = is_express(gi) ? std::min(express_cap, g.slot_cap) : g.slot_cap;
        if (!is_express(gi)) {
            const bool small_end = takes_small(gi);
            // From a list shared with other ranks a group claims at most 64 cells per iteration: batches then grow over a few
            // (short) iterations instead of one rank taking a sixth of the list in its first call, and a rank whose batches
            // are large iterates -- and therefore claims -- more slowly: the ranks' shares balance themselves.
            int claimed = 0;
            const int claim_cap = q ? 64 : 0x7fffffff;
            while ((int)g.active.size() < cap && claimed < claim_cap) {
                const long idx = peek(small_end);
                if (idx < 0) break;
                const int c = pending[(size_t)idx];
                size_t need = slot_bytes(h->h_counts[c]);
                if (g.used + need > g.arena_bytes) break;
                if (!g.active.empty() && g.tiles + cell_tiles(c) > tile_budget) break;   // enough work to fill the GPU share
                if (!claim(small_end, idx)) continue;                                     // another rank was faster: look again
                g.used += need; g.tiles += cell_tiles(c); g.active.push_back(c); claimed++;
                h->owned[(size_t)c] = 1;
            }
        }
        // express cells go to the express lanes; once the bulk list is exhausted, bulk groups whose own batch has
        // become small help out
        if (is_express(gi) || (list_empty() && (int)g.active.size() < express_cap)) {
            const int xcap = std::min(cap, express_cap);
            while (!express_pending.empty() && (int)g.active.size() < xcap) {
                size_t need = slot_bytes(h->h_counts[express_pending.front()]);
                if (g.used + need > g.arena_bytes) break;
                g.used += need; g.tiles += cell_tiles(express_pending.front());
                g.active.push_back(express_pending.front()); express_pending.pop_front();
            }
        }
        g.in_flight = false;
        if (g.active.empty()) return OI_OK;
        const int A = (int)g.active.size();
        size_t off = 0;
        int Nmax = 0;
        g.fl = g.flf = g.flf_chol = g.flf_fit = 0; g.nev = 0;
        // slots are kept in descending size so that the cells owning block row x are a prefix
        std::stable_sort(g.active.begin(), g.active.end(), [&](int a, int b) { return h->h_counts[a] > h->h_counts[b]; });
        for (int a = 0; a < A; a++) {
            int c = g.active[a], n = h->h_counts[c];
            int N = (n + OI_NB - 1) / OI_NB, npad = N * OI_NB;
            OiSlot& s = g.h_slots[a];
            s.M = (double*)(g.arena + off); off += align_up((size_t)npad * npad * 8, 256);
            s.Dinv = (double*)(g.arena + off); off += align_up((size_t)N * OI_TILE * 8, 256);
            s.vec = (double*)(g.arena + off); off += align_up((size_t)3 * npad * 8, 256);
            s.part = (double*)(g.arena + off); off += align_up((size_t)(N + 8 + 5 * N * (N + 1) / 2) * 8, 256);
            s.QE = (double*)(g.arena + off); off += align_up((size_t)N * (N + 1) / 2 * 2 * OI_TILE * 8, 256);
            s.fail = g.d_fail + a;
            s.pt_off = h->h_offsets[c];
            s.cell = c; s.n = n; s.npad = npad; s.N = N; s.n16 = (n + 15) / 16 * 16; s.pad_ = 0;
            Nmax = std::max(Nmax, N);
            double dn = n;
            g.flf_chol += dn * dn * dn / 3;
            if (h_phase[c] == OI_PH_PREDICT) { g.fl += dn * dn * dn / 3 + 19 * dn * dn; g.flf += dn * dn * dn / 3; }
            else { g.fl += dn * dn * dn + 22 * dn * dn; g.flf += dn * dn * dn; g.flf_fit += dn * dn * dn; g.nev++; }
        }
        // cnt_gt[x] = number of (size-sorted) slots with more than x blocks
        g.cnt_gt.assign((size_t)Nmax + 2, 0);
        for (int a = 0; a < A; a++) g.cnt_gt[g.h_slots[a].N]++;                      // histogram of N
        for (int x = Nmax; x >= 0; x--) g.cnt_gt[x] = g.cnt_gt[x + 1] + g.cnt_gt[x];  // -> #slots with N >= x
        for (int x = 0; x <= Nmax; x++) g.cnt_gt[x] = g.cnt_gt[x + 1];                // -> #slots with N > x
        const int* cg = g.cnt_gt.data();
        cudaStream_t st = g.st;
        CK(cudaMemcpyAsync(g.d_slots, g.h_slots, (size_t)A * sizeof(OiSlot), cudaMemcpyHostToDevice, st));
        // Small batches (the optimiser's tail) are bound by the host's launch rate and the gaps between ~4N
        // dependent launches: once a batch has kept its composition for a few iterations, the chain is captured
        // into a CUDA graph and replayed with one call until a cell leaves.
        bool use_graph = false;
        if (graph_max_A > 0 && A <= graph_max_A) {
            if (g.gexec && g.gcells == g.active) use_graph = true;
            else {
                if (g.last_cells == g.active) g.same_count++;
                else { g.same_count = 0; g.last_cells = g.active; }
                if (g.same_count >= 2) {
                    cudaGraph_t graph = nullptr;
                    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                    const int lr = launch_chain(g, A, Nmax, cg, st, false);
                    cudaError_t ce = cudaStreamEndCapture(st, &graph);      // always leave capture mode
                    if (lr) { if (graph) cudaGraphDestroy(graph); return lr; }
                    CK(ce);
                    if (g.gexec) { cudaGraphExecDestroy(g.gexec); g.gexec = nullptr; }
                    CK(cudaGraphInstantiate(&g.gexec, graph, 0));
                    cudaGraphDestroy(graph);
                    g.gcells = g.active;
                    use_graph = true;
                    h->stats.n_graph_captures++;
                }
            }
        }
        g.graph_mode = use_graph;
        if (use_graph) {
            CK(cudaEventRecord(g.ev[0], st));
            CK(cudaGraphLaunch(g.gexec, st));
            CK(cudaEventRecord(g.ev[7], st));
        } else {
            int r = launch_chain(g, A, Nmax, cg, st, true);
            if (r) return r;
        }
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(g.h_slot_phase, g.d_slot_phase, (size_t)A * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(g.done, st));
        g.A = A; g.Nmax = Nmax; g.in_flight = true;
        (void)gi;
        return OI_OK;
    }

    // wait for group g's iteration, account for it, retire the finished cells
    int retire(OiGroup& g, int gi) {
        CK(cudaEventSynchronize(g.done));
        float m[7] = {0, 0, 0, 0, 0, 0, 0};
        oi_stats& S = h->stats;
        if (g.graph_mode) {
            float t = 0; cudaEventElapsedTime(&t, g.ev[0], g.ev[7]);
            S.ms_graph += t; ms_factor += t; S.n_graph_launches++;
        } else
            for (int q = 0; q < 7; q++) cudaEventElapsedTime(&m[q], g.ev[q], g.ev[q + 1]);
        S.ms_build += m[0]; S.ms_chol += m[1]; S.ms_fwd += m[2]; S.ms_trtri += m[3];
        S.ms_alpha += m[4]; S.ms_lauum += m[5]; S.ms_finalize += m[6];
        ms_factor += m[1] + m[3] + m[5];
        const int A = g.A, Nmax = g.Nmax;
        if (trace) {
            float since = 0; cudaEventElapsedTime(&since, h->ev[0], g.ev[7]);     // device time since the fork
            std::fprintf(trace, "%lld,%d,%d,%d,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.6g,%.3f\n", (long long)S.n_iterations, gi, A, Nmax,
                         m[0], m[1], m[2], m[3], m[4], m[5], m[6], g.flf, since);
        }
        S.flops += g.fl; S.flops_factor += g.flf; S.n_evals += g.nev;
        S.flops_chol += g.flf_chol; S.flops_trtri += g.flf_fit / 3; S.flops_lauum += g.flf_fit / 3;
        const bool roww = A >= OI_ROWWISE_MIN_SLOTS_HOST;
        const int chol_l = 2 * Nmax - 1;
        const int scale_l = Nmax > 1 ? 1 : 0;       // 2N launches per factorisation in all
        S.launches_chol += chol_l + scale_l;
        S.launches_trtri += std::max(0, Nmax - 1); S.launches_lauum += roww ? Nmax : 1;
        S.n_iterations++;
        S.n_launches += (roww ? Nmax : 1) + chol_l + scale_l + 1 + std::max(0, Nmax - 1) + 1 +
                        (roww ? Nmax : 1) + 1;
        size_t w = 0;
        const bool bulk = !is_express(gi) && n_express > 0 && A > express_cap;    // only big (slow) batches hand over
        for (int a = 0; a < A; a++) {
            int c = g.active[a];
            h_phase[c] = g.h_slot_phase[a];
            if (h_phase[c] == OI_PH_DONE) { g.used -= slot_bytes(h->h_counts[c]); g.tiles -= cell_tiles(c); }
            else if (bulk && ++iters[c] >= express_after && h_phase[c] == OI_PH_FIT) {
                g.used -= slot_bytes(h->h_counts[c]); g.tiles -= cell_tiles(c);   // hand the long-runner over to an express lane
                express_pending.push_back(c);
                S.n_express_cells++;
            } else g.active[w++] = c;
        }
        g.active.resize(w);
        g.in_flight = false;
        return OI_OK;
    }
};

// Runs lockstep iterations until every cell with observations reaches OI_PH_DONE.
// h_phase: host copy of the initial per-cell phase (already uploaded to ca.phase).
static int run_lockstep(oi_handle* h, std::vector<int>& h_phase, const OiRunConst& rc, double t_pred,
                        double scratch_gib, int max_active, int n_groups) {
    const int nc = (int)h->n_cells;
    LockstepRun R{h, h_phase, rc, t_pred};
    std::vector<int>& pending = R.pending;
    pending.reserve(nc);
    size_t biggest = 0;
    for (int c = 0; c < nc; c++)
        if (h->h_counts[c] > 0 && h_phase[c] != OI_PH_DONE) { pending.push_back(c); biggest = std::max(biggest, slot_bytes(h->h_counts[c])); }
    // largest cells first: cost-sorted ragged batches (cost ~ n^3)
    std::stable_sort(pending.begin(), pending.end(), [&](int a, int b) { return h->h_counts[a] > h->h_counts[b]; });
    if (h->queue.attached()) {
        // every rank must see the same list: it is built from the same observations and cells in the same order
        if (pending.size() >= (1u << 24)) return fail(OI_ERR_ARG, "run_lockstep: shared queue holds at most 2^24 - 1 cells");
        h->queue.begin_run((uint32_t)pending.size());
        R.q = &h->queue;
    }
    if (pending.empty()) return OI_OK;
    R.back = pending.size();
    if (max_active <= 0) max_active = 8192;
    max_active = std::min<int>(max_active, 65535);
    if (const char* e = std::getenv("OI_GROUPS")) n_groups = std::atoi(e);
    if (n_groups <= 0) n_groups = OI_DEFAULT_GROUPS;
    int G = std::max(1, std::min(n_groups, 16));
    size_t want;
    if (scratch_gib > 0) want = (size_t)(scratch_gib * 1073741824.0);
    else {
        size_t fr = 0, tot = 0;
        CK(cudaMemGetInfo(&fr, &tot));
        fr += h->arena_bytes;
        want = std::min<size_t>((size_t)(fr * 0.8), (size_t)64 << 30);
    }
    size_t all = 0;
    for (int c : pending) all += slot_bytes(h->h_counts[c]);
    while (G > 1 && want / G < biggest) G--;            // every group must be able to hold the largest cell
    // express lanes: the last n_express groups (see LockstepRun); they hold at most express_cap cells each
    R.G = G;
    R.n_express = G >= 4 ? std::max(1, G / 4) : 0;
    R.express_after = 224; R.express_cap = 24;
    if (const char* e = std::getenv("OI_EXPRESS")) R.n_express = std::min(std::max(std::atoi(e), 0), G - 1);
    if (const char* e = std::getenv("OI_EXPRESS_AFTER")) R.express_after = std::max(std::atoi(e), 1);
    if (const char* e = std::getenv("OI_EXPRESS_CAP")) R.express_cap = std::max(std::atoi(e), 1);
    const int nbulk = G - R.n_express;
    R.n_small = nbulk >= 5 ? 2 : (nbulk >= 3 ? 1 : 0);      // measured on the 2390-cell step, 6 bulk groups: 0 -> 15.5 s, 1 -> 15.2 s, 2 -> 14.3-14.5 s, 3 -> 14.6 s
    if (const char* e = std::getenv("OI_SMALL_GROUPS")) R.n_small = std::min(std::max(std::atoi(e), 0), nbulk - 1);
    // arena: the express lanes get room for express_cap of the largest cells, the bulk groups share the rest equally;
    // never more than everything resident
    const size_t xshare = R.n_express > 0 ? align_up(std::min<size_t>((size_t)R.express_cap * biggest, want / (2 * G)) + biggest, 256) : 0;
    want = std::min(want, align_up(all / nbulk + all / (3 * nbulk) + biggest, 256) * nbulk + xshare * R.n_express);   // 1.33x: the first groups take the largest cells
    want = std::max(want, (biggest + 256) * nbulk + xshare * R.n_express);
    const int total_slots = std::min<int>(max_active, (int)pending.size());
    const int per = std::max((total_slots + nbulk - 1) / nbulk, std::min(R.express_cap, total_slots));
    int rcode = ensure_batch_buffers(h, want, per * G, G);
    if (rcode) return rcode;
    const size_t share = ((h->arena_bytes - xshare * R.n_express) / nbulk) & ~(size_t)255;
    for (int gi = 0; gi < G; gi++) {
        OiGroup& g = *h->groups[gi];
        const bool x = gi >= nbulk;
        g.arena = h->arena + (x ? share * nbulk + xshare * (size_t)(gi - nbulk) : share * (size_t)gi);
        g.arena_bytes = x ? xshare : share; g.used = 0;
        g.d_slots = h->d_slots + (size_t)gi * per; g.h_slots = h->h_slots + (size_t)gi * per;
        g.d_slot_phase = h->d_slot_phase + (size_t)gi * per; g.h_slot_phase = h->h_slot_phase + (size_t)gi * per;
        g.d_fail = h->d_fail + (size_t)gi * per;
        g.slot_cap = per; g.active.clear(); g.in_flight = false;
    }
    R.pk = OiPacked{h->px, h->py, h->pt, h->pr};
    // OI_TRACE=<file>: one CSV line per group iteration (active cells, Nmax, stream-ms per kernel family)
    struct TraceGuard { FILE*& f; ~TraceGuard() { if (f) { std::fclose(f); f = nullptr; } } } trace_guard{R.trace};   // closes on every return path
    if (const char* tp = std::getenv("OI_TRACE")) {
        R.trace = std::fopen(tp, "a");
        if (R.trace) std::fprintf(R.trace, "iter,group,active,Nmax,ms_build,ms_chol,ms_fwd,ms_trtri,ms_alpha,ms_lauum,ms_finalize,flops_factor,t_ms\n");
    }
    R.iters.assign((size_t)nc, 0);
    R.graph_max_A = 32;
    if (const char* e = std::getenv("OI_GRAPH_MAX")) R.graph_max_A = std::max(std::atoi(e), 0);
    for (int gi = 0; gi < G; gi++) { OiGroup& g = *h->groups[gi]; g.gcells.clear(); g.last_cells.clear(); g.same_count = 0; }
    // work admitted per bulk group: enough tiles in flight to fill the GPU, few enough that an iteration stays short
    // (a cell's iteration latency = active work / throughput; long optimiser runs are only recognised by their
    // iteration count, so latency decides how early they reach an express lane)
    long long tile_total = 80000;       // measured with two small-first groups (2390-cell step, two boxes): 40 k 14.31 s, 60 k 14.25, 80 k 14.00-14.20, 110 k 14.62, 150 k 15.06
    if (const char* e = std::getenv("OI_TILE_BUDGET")) tile_total = std::max(1LL, std::atoll(e));
    R.tile_budget = std::max(1LL, tile_total / std::max(1, G - R.n_express));
    {
        int least = 0, greatest = 0;
        CK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for (int gi = 0; gi < G; gi++) {
            OiGroup& g = *h->groups[gi];
            const bool want = R.is_express(gi) && greatest != least;
            if (g.high_prio != want) {               // express lanes run on high-priority streams
                CK(cudaStreamSynchronize(g.st));
                cudaStream_t old_st = g.st;
                g.st = nullptr;                       // never leave a destroyed stream in the group (free_groups skips NULL)
                CK(cudaStreamDestroy(old_st));
                CK(cudaStreamCreateWithPriority(&g.st, cudaStreamNonBlocking, want ? greatest : least));
                g.high_prio = want;
            }
            g.tiles = 0;
        }
    }
    // fork: the group streams start after everything queued on the handle's stream.  From here on errors do not
    // return early: they fall through to the join below so that the handle's stream is ordered after every group.
    int rc2 = OI_OK;
    {
        cudaError_t fe = cudaEventRecord(h->ev[0], h->st);
        for (int gi = 0; gi < G && fe == cudaSuccess; gi++) fe = cudaStreamWaitEvent(h->groups[gi]->st, h->ev[0], 0);
        if (fe != cudaSuccess) rc2 = fail(OI_ERR_CUDA, std::string("run_lockstep: fork: ") + cudaGetErrorString(fe));
    }
    for (int gi = 0; gi < G && !rc2; gi++) rc2 = R.issue(*h->groups[gi], gi);
    // event loop: poll the groups (non-blocking), express lanes first and again after every serviced bulk group,
    // so that their short iterations are never held up behind the host work of a bulk group
    auto service = [&](int gi, bool& progressed) -> int {
        OiGroup& g = *h->groups[gi];
        if (g.in_flight) {
            cudaError_t q = cudaEventQuery(g.done);
            if (q == cudaErrorNotReady) return OI_OK;
            if (q != cudaSuccess) return fail(OI_ERR_CUDA, std::string("run_lockstep: ") + cudaGetErrorString(q));
            int r = R.retire(g, gi);
            if (r) return r;
            progressed = true;
        }
        int r = R.issue(g, gi);
        if (g.in_flight) progressed = true;
        return r;
    };
    while (!rc2) {
        bool progressed = false;
        for (int xi = G - R.n_express; xi < G && !rc2; xi++) rc2 = service(xi, progressed);
        for (int gi = 0; gi < G - R.n_express && !rc2; gi++) {
            bool p = false;
            rc2 = service(gi, p);
            if (p) {
                progressed = true;
                for (int xi = G - R.n_express; xi < G && !rc2; xi++) rc2 = service(xi, progressed);
            }
        }
        bool any = false;
        for (int gi = 0; gi < G; gi++) any = any || h->groups[gi]->in_flight;
        if (!any) break;
        if (!progressed) std::this_thread::yield();
    }
    if (!rc2 && (!R.list_empty() || !R.express_pending.empty()))
        rc2 = fail(OI_ERR_NOMEM, "run_lockstep: scratch arena too small for one cell");
    // join: the handle's stream continues after every group
    for (int gi = 0; gi < G; gi++) {
        cudaEventRecord(h->groups[gi]->done, h->groups[gi]->st);
        cudaStreamWaitEvent(h->st, h->groups[gi]->done, 0);
    }
    if (rc2) cudaDeviceSynchronize();
    h->stats.ms_factor += R.ms_factor;
    h->stats.n_groups = G;
    return rc2;
}

static int pack_points(oi_handle* h, double mean) {
    oi_launch_pack(h->indices, h->total, h->ox, h->oy, h->ot, h->oz, mean, h->t_shift, h->px, h->py, h->pt, h->pr, h->st);
    CK(cudaGetLastError());
    return OI_OK;
}

static void reset_stats(oi_handle* h) {
    double g = h->stats.ms_gather; int64_t sn = h->stats.sum_n;
    h->stats = oi_stats{};
    h->stats.ms_gather = g; h->stats.sum_n = sn;
}

extern "C" int oi_nlml_grad(oi_handle* h, const double* hypers, int32_t n_hyp, double prior_mean, int32_t grad_convention,
                            double* nlz_out, double* grad_out) {
    if (!h || !hypers || !nlz_out || !grad_out || n_hyp < 5 || n_hyp > OI_MAXH) return fail(OI_ERR_ARG, "oi_nlml_grad: bad argument");
    if (!h->have_nbr) return fail(OI_ERR_STATE, "oi_nlml_grad: call oi_gather_neighbours first");
    CK(cudaSetDevice(h->device));
    const int nc = (int)h->n_cells;
    std::vector<double> hyp((size_t)nc * 5);
    std::vector<int> phase(nc);
    h->owned.assign((size_t)nc, 0);
    for (int c = 0; c < nc; c++) {
        for (int q = 0; q < 5; q++) hyp[(size_t)c * 5 + q] = std::exp(hypers[(size_t)c * n_hyp + q]);   // GPR_CS2S3.py:120-122
        phase[c] = h->h_counts[c] > 0 ? OI_PH_EVAL : OI_PH_DONE;
    }
    CK(cudaMemcpyAsync(h->ca.hyp, hyp.data(), hyp.size() * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->ca.phase, phase.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, h->st));
    int r = pack_points(h, prior_mean);
    if (r) return r;
    CK(cudaStreamSynchronize(h->st));
    OiRunConst rc{};
    rc.mean = prior_mean; rc.gtol = 1e-5; rc.n_hyp = n_hyp; rc.grad_convention = grad_convention; rc.maxiter = 0;
    reset_stats(h);
    r = run_lockstep(h, phase, rc, 0.0, 0.0, 0, 0);
    if (r) return r;
    std::vector<double> f(nc), g((size_t)nc * OI_MAXH);
    CK(cudaMemcpy(f.data(), h->ca.evf, (size_t)nc * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(g.data(), h->ca.evg, g.size() * 8, cudaMemcpyDeviceToHost));
    for (int c = 0; c < nc; c++) {
        bool none = h->h_counts[c] <= 0;
        nlz_out[c] = none ? NAN : f[c];
        for (int q = 0; q < n_hyp; q++) grad_out[(size_t)c * n_hyp + q] = none ? NAN : g[(size_t)c * OI_MAXH + q];
    }
    return OI_OK;
}

extern "C" int oi_run(oi_handle* h, const oi_params* p, const double* hypers_in) {
    if (!h || !p) return fail(OI_ERR_ARG, "oi_run: NULL argument");
    if (p->mode != OI_MODE_FIT && p->mode != OI_MODE_PREDICT) return fail(OI_ERR_ARG, "oi_run: bad mode");
    if (p->mode == OI_MODE_PREDICT && !hypers_in) return fail(OI_ERR_ARG, "oi_run: predict mode needs hypers_in");
    if (p->mode == OI_MODE_FIT && (p->n_hyp < 5 || p->n_hyp > OI_MAXH)) return fail(OI_ERR_ARG, "oi_run: n_hyp must be 5 or 6");
    if (p->engine != OI_ENGINE_LOCKSTEP)
        return fail(OI_ERR_ARG, "oi_run: only OI_ENGINE_LOCKSTEP exists (the experimental persistent engine was removed in v1.2)");
    if (p->optimiser != OI_OPT_CG && p->optimiser != OI_OPT_LBFGS) return fail(OI_ERR_ARG, "oi_run: bad optimiser");
    if (!h->have_nbr) return fail(OI_ERR_STATE, "oi_run: call oi_gather_neighbours first");
    CK(cudaSetDevice(h->device));
    const int nc = (int)h->n_cells;
    reset_stats(h);
    CK(cudaEventRecord(h->ev[8], h->st));
    int r = pack_points(h, p->prior_mean);
    if (r) return r;
    OiRunConst rc{};
    rc.mean = p->prior_mean; rc.gtol = p->gtol > 0 ? p->gtol : 1e-5; rc.n_hyp = p->mode == OI_MODE_FIT ? p->n_hyp : 5;
    rc.grad_convention = p->grad_convention; rc.maxiter = p->maxiter;
    rc.optimiser = p->optimiser == OI_OPT_LBFGS ? 1 : 0;
    for (int q = 0; q < OI_MAXH; q++) rc.x0[q] = p->x0[q];
    std::vector<int> phase(nc);
    h->owned.assign((size_t)nc, 0);
    for (int c = 0; c < nc; c++) if (h->h_counts[c] <= 0) h->owned[(size_t)c] = 1;
    // cells without observations: NaN tuple, status NO_OBS (the reference would raise inside pdist)
    std::vector<double> out0((size_t)nc * 8, NAN);
    std::vector<int> st0(nc), nf0(nc, 0);
    for (int c = 0; c < nc; c++) {
        bool none = h->h_counts[c] <= 0;
        phase[c] = none ? OI_PH_DONE : (p->mode == OI_MODE_FIT ? OI_PH_FIT : OI_PH_PREDICT);
        st0[c] = none ? OI_CELL_NO_OBS : OI_CELL_OK;
    }
    CK(cudaMemcpyAsync(h->ca.out, out0.data(), out0.size() * 8, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->ca.status, st0.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->ca.nfev, nf0.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, h->st));
    CK(cudaMemcpyAsync(h->ca.phase, phase.data(), (size_t)nc * 4, cudaMemcpyHostToDevice, h->st));
    h->ca.dbg = h->d_dbg; h->ca.dbg_cell = h->dbg_cell; h->ca.dbg_cap = h->dbg_cap; h->ca.dbg_count = h->d_dbg_count;
    if (h->d_dbg_count) CK(cudaMemsetAsync(h->d_dbg_count, 0, 4, h->st));
    if (p->mode == OI_MODE_FIT) {
        oi_launch_cg_init(h->ca, nc, rc, h->st);
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(h->ca.hyp, hypers_in, (size_t)nc * 40, cudaMemcpyHostToDevice, h->st));
    }
    CK(cudaStreamSynchronize(h->st));   // staging vectors go out of scope below
    r = run_lockstep(h, phase, rc, p->t_pred, p->scratch_gib, p->max_active, p->n_groups);
    if (r) return r;
    CK(cudaEventRecord(h->ev[9], h->st));
    CK(cudaStreamSynchronize(h->st));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[8], h->ev[9]);
    h->stats.ms_total = ms;
    h->have_results = true;
    return OI_OK;
}

extern "C" int oi_get_results(oi_handle* h, double* out, int32_t* n_out, int32_t* nfev_out, int32_t* status_out) {
    if (!h || !out) return fail(OI_ERR_ARG, "oi_get_results: NULL argument");
    if (!h->have_results) return fail(OI_ERR_STATE, "oi_get_results: call oi_run first");
    CK(cudaSetDevice(h->device));
    const size_t nc = (size_t)h->n_cells;
    CK(cudaMemcpy(out, h->ca.out, nc * 64, cudaMemcpyDeviceToHost));
    if (nfev_out) CK(cudaMemcpy(nfev_out, h->ca.nfev, nc * 4, cudaMemcpyDeviceToHost));
    if (status_out) CK(cudaMemcpy(status_out, h->ca.status, nc * 4, cudaMemcpyDeviceToHost));
    if (n_out) std::memcpy(n_out, h->h_counts.data(), nc * 4);
    return OI_OK;
}

// Several GPU processes of one box share ONE cost-sorted work list (oi_shared_queue.h).  Every process calls this with the
// same (fresh, unique) POSIX shared-memory name before its first oi_run, and then the same sequence of oi_run calls on the
// same observations and cells; each cell is computed by exactly one of them.  NULL detaches.
extern "C" int oi_set_shared_queue(oi_handle* h, const char* shm_name) {
    if (!h) return fail(OI_ERR_ARG, "oi_set_shared_queue: NULL handle");
    if (!shm_name) { h->queue.detach(); return OI_OK; }
    if (std::strlen(shm_name) >= sizeof(h->queue.name) || shm_name[0] != '/') return fail(OI_ERR_ARG, "oi_set_shared_queue: name must start with '/' and be shorter than 128 characters");
    int e = h->queue.attach(shm_name);
    if (e) return fail(OI_ERR_STATE, std::string("oi_set_shared_queue: shm_open/mmap failed: ") + std::strerror(e));
    return OI_OK;
}
extern "C" int oi_unlink_shared_queue(const char* shm_name) {
    if (!shm_name) return fail(OI_ERR_ARG, "oi_unlink_shared_queue: NULL name");
    OiSharedQueue::unlink_name(shm_name);
    return OI_OK;
}
// owned[n_cells]: 1 where this handle computed the cell in the last oi_run (all cells without a shared queue).
extern "C" int oi_get_owned(oi_handle* h, uint8_t* owned) {
    if (!h || !owned) return fail(OI_ERR_ARG, "oi_get_owned: NULL argument");
    if (!h->have_results) return fail(OI_ERR_STATE, "oi_get_owned: call oi_run first");
    std::memcpy(owned, h->owned.data(), h->owned.size());
    return OI_OK;
}

// Diagnostic: record every objective evaluation (natural-unit hyperparameters, value, gradient) the optimiser of ONE cell
// sees during the next oi_run(OI_MODE_FIT) calls.  cell < 0 switches it off.
extern "C" int oi_debug_trace(oi_handle* h, int64_t cell, int32_t capacity) {
    if (!h) return fail(OI_ERR_ARG, "oi_debug_trace: NULL handle");
    CK(cudaSetDevice(h->device));
    cudaFree(h->d_dbg); cudaFree(h->d_dbg_count); h->d_dbg = nullptr; h->d_dbg_count = nullptr; h->dbg_cell = -1; h->dbg_cap = 0;
    if (cell < 0 || capacity <= 0) return OI_OK;
    CK(cudaMalloc(&h->d_dbg, (size_t)capacity * 12 * 8));
    CK(cudaMalloc(&h->d_dbg_count, 4));
    CK(cudaMemset(h->d_dbg_count, 0, 4));
    h->dbg_cell = (int)cell; h->dbg_cap = capacity;
    return OI_OK;
}
extern "C" int oi_get_debug_trace(oi_handle* h, double* rows, int32_t* n_rows) {
    if (!h || !rows || !n_rows) return fail(OI_ERR_ARG, "oi_get_debug_trace: NULL argument");
    if (!h->d_dbg) return fail(OI_ERR_STATE, "oi_get_debug_trace: call oi_debug_trace first");
    CK(cudaSetDevice(h->device));
    int cnt = 0;
    CK(cudaMemcpy(&cnt, h->d_dbg_count, 4, cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, h->dbg_cap);
    if (cnt > 0) CK(cudaMemcpy(rows, h->d_dbg, (size_t)cnt * 12 * 8, cudaMemcpyDeviceToHost));
    *n_rows = cnt;
    return OI_OK;
}

extern "C" int oi_get_stats(oi_handle* h, oi_stats* s) {
    if (!h || !s) return fail(OI_ERR_ARG, "oi_get_stats: NULL argument");
    *s = h->stats;
    return OI_OK;
}

extern "C" int oi_gpr_day(oi_handle* h, const double* x, const double* y, const double* t, const double* z, int64_t n_obs,
                          const double* X, int64_t n_cells, const oi_params* p, const double* hypers_in,
                          double* out, int32_t* n_out, int32_t* nfev_out, int32_t* status_out) {
    if (!p) return fail(OI_ERR_ARG, "oi_gpr_day: NULL params");
    int r;
    if ((r = oi_set_observations(h, x, y, t, z, n_obs))) return r;
    if ((r = oi_set_cells(h, X, n_cells))) return r;
    if ((r = oi_gather_neighbours(h, p->radius_m, nullptr))) return r;
    if ((r = oi_run(h, p, hypers_in))) return r;
    return oi_get_results(h, out, n_out, nfev_out, status_out);
}
