// Hand-written sm_100a kernels of the per-cell GP hot path: the neighbour gather, the two
// execution engines over the tile functions of oi_tiles.cuh, and their launch wrappers.
//
//   k_gather / k_scan_counts / k_pack   neighbour gather        (GPR_CS2S3.py:159-164)
//   k_gp_persistent                     persistent group engine: a group of CTAs takes a cell from a queue and
//                                       runs whole NLML+gradient evaluations and optimiser steps on it
//   k_build ... k_finalize              lockstep engine: one launch per algorithmic step over all active cells
//
// Compiled with -fmad=false (see oi_optim.cuh); hot scalar loops use explicit fma().
#ifndef OI_PIPE_MINB
#define OI_PIPE_MINB 4      // resident CTAs per SM the pipeline-only kernels are compiled for (48 KB of shared memory each)
#endif
#include "oi_tiles.cuh"
#include "oi_optim.cuh"

// ------------------------------------------------------------------------------------------
// kernel (1): neighbour gather.  One warp per cell, observations staged through shared memory
// in chunks shared by the 8 cells of the CTA; ordered compaction by ballot + popc prefix.
// ------------------------------------------------------------------------------------------
#define G_CHUNK 1024
template <bool FILL>
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ ox, const double* __restrict__ oy,
                                                const double* __restrict__ ot, int n_obs,
                                                const double* __restrict__ X, int n_cells, double r2, double t_lo, double t_hi,
                                                int* __restrict__ counts, const long long* __restrict__ offsets,
                                                int* __restrict__ indices) {
    __shared__ double sx[G_CHUNK], sy[G_CHUNK], st[G_CHUNK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell = blockIdx.x * 8 + warp;
    const bool live = cell < n_cells;
    double cx = 0, cy = 0;
    if (live) { cx = X[2 * cell]; cy = X[2 * cell + 1]; }
    long long base = 0;
    if (FILL && live) base = offsets[cell];
    int cnt = 0;
    for (int c0 = 0; c0 < n_obs; c0 += G_CHUNK) {
        int m = min(G_CHUNK, n_obs - c0);
        __syncthreads();
        for (int q = threadIdx.x; q < m; q += 256) { sx[q] = ox[c0 + q]; sy[q] = oy[c0 + q]; st[q] = ot[c0 + q]; }
        __syncthreads();
        if (live) {
            for (int q0 = 0; q0 < m; q0 += 32) {
                int q = q0 + lane;
                bool in = false;
                if (q < m) {
                    double dx = sx[q] - cx, dy = sy[q] - cy;
                    // inclusive boundary, no FMA: ties on the 25 km lattice resolve as in the reference (GPR_CS2S3.py:159);
                    // the day window [t_lo, t_hi] (inclusive; +-inf = off) is the reference's obs[..., day:day+T] (:213)
                    in = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2 && st[q] >= t_lo && st[q] <= t_hi;
                }
                unsigned bal = __ballot_sync(0xffffffffu, in);
                if (FILL && in) indices[base + cnt + __popc(bal & ((1u << lane) - 1u))] = c0 + q;
                cnt += __popc(bal);
            }
        }
    }
    if (!FILL && live && lane == 0) counts[cell] = cnt;
}

__global__ void k_scan_counts(const int* __restrict__ counts, int n, long long* __restrict__ offsets) {
    // single-block exclusive scan (n_cells ~ 2e4): 1024 threads, sequential chunks
    __shared__ long long part[1024];
    int per = (n + 1023) / 1024;
    int lo = threadIdx.x * per, hi = min(n, lo + per);
    long long s = 0;
    for (int i = lo; i < hi; i++) s += counts[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int i = 0; i < 1024; i++) { long long v = part[i]; part[i] = run; run += v; }
        offsets[n] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (int i = lo; i < hi; i++) { offsets[i] = run; run += counts[i]; }
}

__global__ void k_pack(const int* __restrict__ indices, long long total, const double* __restrict__ ox,
                       const double* __restrict__ oy, const double* __restrict__ ot, const double* __restrict__ oz,
                       double mean, double t_shift, double* __restrict__ px, double* __restrict__ py, double* __restrict__ pt,
                       double* __restrict__ pr) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int id = indices[i];
    // t counts from the start of the day window (the reference's t_train is the day index inside the window, :227-235);
    // outputs - mX (GPR_CS2S3.py:127)
    px[i] = ox[id]; py[i] = oy[id]; pt[i] = ot[id] - t_shift; pr[i] = oz[id] - mean;
}


// ------------------------------------------------------------------------------------------
// lockstep engine: thin kernels around the tile functions, one launch per step
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(OI_THREADS) k_build(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, int row) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }      // one launch per block row (exact grids for big batches)
    else tile_ij(blockIdx.x, i, j);
    if (i >= s.N) return;
    tile_build(s, ca, pk, i, j, ca.phase[s.cell] != OI_PH_PREDICT, smem);
}

__global__ void __launch_bounds__(OI_THREADS, OI_PIPE_MINB) k_chol_update(const OiSlot* __restrict__ slots, int k) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    const int i = k + blockIdx.x;
    if (i >= s.N) return;
    if (k == 0 && i != 0) return;          // nothing to subtract from the first block column
    if (OI_FAILED(s)) return;
    tile_chol_update(s, i, k, smem);
}

// Everything that only waits for Dinv[k]: the panel tiles (i, k), i > k, and -- row k of L being final once block
// column k has been updated -- the row scaling Ls_kj = L_kk^-1 L_kj of row k (j < k).  One launch, no extra pass.
__global__ void __launch_bounds__(OI_THREADS, 3) k_chol_panel(const OiSlot* __restrict__ slots, int k, int n_panel) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    if (k >= s.N) return;
    if (OI_FAILED(s)) return;
    if ((int)blockIdx.x < n_panel) {
        const int i = k + 1 + blockIdx.x;
        if (i < s.N) tile_chol_panel(s, i, k, smem);
    } else {
        const int j = blockIdx.x - n_panel;
        if (j < k) tile_scale(s, k, j, smem);
    }
}


// ------------------------------------------------------------------------------------------
// fused Cholesky: ONE launch for all block columns of all cells.  CTAs draw a ticket (atomic counter, so the
// ticket order is the order in which CTAs start, whatever the hardware's dispatch order) and the tickets are laid
// out column by column: [diagonal tiles of column k, all cells][off-diagonal tiles of column k, all cells] ...
// A tile only ever waits for tiles with smaller tickets (which have started and never wait for larger ones), so
// the chain always makes progress:
//   tile (i,k), k > 0 : waits until the previous column of ITS cell is complete (col_done[k-1] == N-k)
//   diagonal (k,k)    : update + factor + inverse, then publishes diag_done[k]
//   off-diagonal (i,k): update, waits for diag_done[k], panel product L_ik = A_ik L_kk^-T, then col_done[k] += 1
// There are no launch boundaries between the 2N dependent steps: a cell's next column starts as soon as that
// cell is ready, the serial 64x64 factorisations hide behind the other cells' tiles, and small batches save the
// launch gaps.  A failed factorisation still publishes its flags (the tiles behind it skip their work).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_flag_ge(const int* p, int target) {     // thread 0 spins, the CTA follows
    if (threadIdx.x == 0) {
        while (ld_acquire_s32(p) < target) __nanosleep(64);
    }
    __syncthreads();
}
__global__ void __launch_bounds__(OI_THREADS, 3) k_chol_fused(const OiSlot* __restrict__ slots, OiCholPlan plan, int* __restrict__ ticket) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_t[4];
    if (threadIdx.x == 0) {
        const int t = atomicAdd(ticket, 1);
        int seg = 0;
        while (seg + 1 < 2 * plan.Nmax && plan.off[seg + 1] <= t) seg++;
        const int k = seg >> 1, local = t - plan.off[seg];
        int slot, i;
        if ((seg & 1) == 0) { slot = local; i = k; }
        else { const int tpc = plan.Nmax - k - 1; slot = local / tpc; i = k + 1 + local % tpc; }
        s_t[0] = k; s_t[1] = slot; s_t[2] = i;
    }
    __syncthreads();
    const int k = s_t[0], i = s_t[2];
    const OiSlot s = slots[s_t[1]];
    if (i >= s.N) return;
    int* diag_done = s.flags;
    int* col_done = s.flags + s.N;
    if (k > 0) wait_flag_ge(&col_done[k - 1], s.N - k);          // all N-k off-diagonal tiles of column k-1 are final
    if (!OI_FAILED(s)) tile_chol_update(s, i, k, smem);
    __syncthreads();
    if (i == k) {
        if (threadIdx.x == 0) { __threadfence(); atomicExch(&diag_done[k], 1); }
        return;
    }
    wait_flag_ge(&diag_done[k], 1);
    if (!OI_FAILED(s)) tile_chol_panel(s, i, k, smem);
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(&col_done[k], 1); }
}

__global__ void __launch_bounds__(256) k_fwd(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, double t_pred) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.x];
    if (OI_FAILED(s)) return;
    cell_fwd(s, ca, pk, t_pred, ca.phase[s.cell] == OI_PH_PREDICT, smem);
}

__global__ void __launch_bounds__(OI_THREADS, 3) k_scale_rows(const OiSlot* __restrict__ slots, int row) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }
    else { tile_ij(blockIdx.x, i, j); i += 1; }   // strictly lower tiles: (i, j), 1 <= i < N, j < i
    if (i >= s.N) return;
    if (OI_FAILED(s)) return;
    tile_scale(s, i, j, smem);
}

__global__ void __launch_bounds__(OI_THREADS, OI_PIPE_MINB) k_trtri(const OiSlot* __restrict__ slots, const int* __restrict__ phase, int d) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    const int kb = blockIdx.x;
    if (kb + d >= s.N) return;
    if (phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    tile_trtri(s, kb, d, smem);
}

__global__ void __launch_bounds__(OI_THREADS) k_alpha(const OiSlot* __restrict__ slots, const int* __restrict__ phase) {
    const OiSlot s = slots[blockIdx.y];
    const int rb = blockIdx.x;
    if (rb >= s.N) return;
    if (phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    rows_alpha(s, rb);
}

__global__ void __launch_bounds__(OI_THREADS, OI_PIPE_MINB) k_lauum_trace(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, int row) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }
    else tile_ij(blockIdx.x, i, j);
    if (i >= s.N) return;
    if (ca.phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    tile_lauum_trace(s, ca, pk, i, j, smem);
}

__global__ void __launch_bounds__(128) k_finalize(const OiSlot* __restrict__ slots, int A, OiCellArrays ca, OiRunConst rc,
                                                  int* __restrict__ slot_phase) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= A) return;
    const OiSlot s = slots[warp];
    int np = warp_finalize(s, ca, rc, ca.phase[s.cell], lane);
    if (lane == 0) slot_phase[warp] = np;
}

// ------------------------------------------------------------------------------------------
// persistent group engine
//
// The grid is n_groups x gs CTAs, all co-resident (3 per SM).  A group takes the next cell of the
// (size-sorted) work list, its gs CTAs walk the cell through build -> Cholesky -> substitution ->
// inverse -> trace terms with a static tile-to-rank assignment, synchronising through a monotonic
// counter in global memory, and rank 0 then runs the optimiser step.  The group keeps evaluating the
// same cell (up to evals_cap evaluations per launch) before it takes the next one: cells advance
// independently of each other, so there is no launch boundary or lockstep between them and the
// low-parallelism steps of one cell overlap with the bulk steps of the cells sharing its SMs.
// Scratch (matrix, Dinv, vectors, partial sums) belongs to the group, not to the cell.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all gs CTAs of the group arrive; epoch is the running arrival target (uniform over the group)
__device__ __forceinline__ void group_barrier(OiGroupCtl* ctl, unsigned& epoch, int gs) {
    epoch += (unsigned)gs;
    if (gs == 1) { __syncthreads(); return; }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&ctl->count, 1u);
        while ((int)(ld_acquire_u32(&ctl->count) - epoch) < 0) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
}

enum { PT_BUILD = 0, PT_CHOL, PT_SCALE, PT_FWD_TRTRI, PT_ALPHA, PT_LAUUM, PT_FINAL, PT_IDLE, PT_N };   // gs == 1: fwd is booked under PT_SCALE

__global__ void __launch_bounds__(OI_THREADS, 3) k_gp_persistent(OiPersist P, OiCellArrays ca, OiPacked pk, OiRunConst rc, double t_pred) {
    extern __shared__ __align__(16) double smem[];
    const int gs = P.gs, grp = blockIdx.x / gs, r = blockIdx.x % gs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    OiGroupCtl* ctl = P.ctl + grp;
    char* scratch = P.scratch + (size_t)grp * P.scratch_stride;
    unsigned epoch = 0;
    // per-CTA accounting lives in shared memory (thread 0 only): cycles per phase, algorithmic flops, counts
    __shared__ long long tacc[PT_N + 1];
    __shared__ double facc[3];
    __shared__ long long nacc[2];
    if (tid == 0) {
        for (int q = 0; q < PT_N; q++) tacc[q] = 0;
        tacc[PT_N] = clock64();
        facc[0] = facc[1] = facc[2] = 0.0; nacc[0] = nacc[1] = 0;
    }
#define PT_MARK(ph) do { if (tid == 0) { long long now_ = clock64(); tacc[ph] += now_ - tacc[PT_N]; tacc[PT_N] = now_; } } while (0)

    for (int round = 0;; round++) {
        if (r == 0 && tid == 0) {
            int idx = atomicAdd(P.queue_head, 1);
            ctl->cur[round & 1] = idx < P.n_work ? idx : -1;
        }
        group_barrier(ctl, epoch, gs);
        const int widx = *(volatile int*)&ctl->cur[round & 1];
        PT_MARK(PT_IDLE);
        if (widx < 0) break;
        const OiWork w = P.work[widx];
        OiSlot s;
        {
            const int N = (w.n + NB - 1) / NB, npad = N * NB;
            size_t off = 0;
            s.M = (double*)(scratch + off); off += ((size_t)npad * npad * 8 + 255) & ~(size_t)255;
            s.Dinv = (double*)(scratch + off); off += ((size_t)N * OI_TILE * 8 + 255) & ~(size_t)255;
            s.vec = (double*)(scratch + off); off += ((size_t)3 * npad * 8 + 255) & ~(size_t)255;
            s.part = (double*)(scratch + off); off += ((size_t)(N + 8 + 5 * N * (N + 1) / 2) * 8 + 255) & ~(size_t)255;
            s.QE = (double*)(scratch + off); off += ((size_t)N * (N + 1) / 2 * 2 * OI_TILE * 8 + 255) & ~(size_t)255;
            s.flags = (int*)(scratch + off);
            s.fail = P.fail + grp;
            s.pt_off = w.pt_off; s.cell = w.cell; s.n = w.n; s.npad = npad; s.N = N; s.n16 = (w.n + 15) / 16 * 16; s.pad_ = 0;
        }
        const int N = s.N, ntl = N * (N + 1) / 2;
        const double dn = (double)s.n;
        for (int ev = 0; ev < P.evals_cap; ev++) {
            const int phase = __ldcg(&ca.phase[s.cell]);
            if (phase == OI_PH_DONE) break;
            const bool pred = phase == OI_PH_PREDICT;
            // ---- covariance ----
            for (int t = r; t < ntl; t += gs) { int i, j; tile_ij(t, i, j); tile_build(s, ca, pk, i, j, !pred, smem); }
            group_barrier(ctl, epoch, gs);
            PT_MARK(PT_BUILD);
            // ---- blocked left-looking Cholesky ----
            bool failed = false;
            for (int k = 0; k < N; k++) {
                if (k == 0) { if (r == 0) tile_chol_update(s, 0, 0, smem); }
                else for (int q = r; q < N - k; q += gs) tile_chol_update(s, k + q, k, smem);
                group_barrier(ctl, epoch, gs);
                failed = OI_FAILED(s);
                if (failed) break;
                if (k + 1 < N) {
                    for (int q = r; q < N - k - 1; q += gs) tile_chol_panel(s, k + 1 + q, k, smem);
                    group_barrier(ctl, epoch, gs);
                }
            }
            PT_MARK(PT_CHOL);
            if (!failed) {
                if (N > 1) {
                    for (int t = r; t < ntl - N; t += gs) { int i, j; tile_ij(t, i, j); tile_scale(s, i + 1, j, smem); }
                    group_barrier(ctl, epoch, gs);
                }
                PT_MARK(PT_SCALE);
                // ---- substitution (last rank) next to the first block distance of the inverse (other ranks) ----
                if (r == gs - 1) cell_fwd(s, ca, pk, t_pred, pred, smem);
                if (gs == 1) PT_MARK(PT_SCALE);
                if (!pred) {
                    const int wk = gs > 1 ? gs - 1 : 1;
                    if (r < wk) for (int kb = r; kb + 1 < N; kb += wk) tile_trtri(s, kb, 1, smem);
                    group_barrier(ctl, epoch, gs);
                    for (int d = 2; d < N; d++) {
                        for (int kb = r; kb + d < N; kb += gs) tile_trtri(s, kb, d, smem);
                        group_barrier(ctl, epoch, gs);
                    }
                    PT_MARK(PT_FWD_TRTRI);
                    for (int rb = r; rb < N; rb += gs) rows_alpha(s, rb);
                    group_barrier(ctl, epoch, gs);
                    PT_MARK(PT_ALPHA);
                    for (int t = r; t < ntl; t += gs) { int i, j; tile_ij(t, i, j); tile_lauum_trace(s, ca, pk, i, j, smem); }
                }
                group_barrier(ctl, epoch, gs);
                PT_MARK(PT_LAUUM);
            }
            // ---- NLML, gradient, optimiser step ----
            if (r == 0 && warp == 0) warp_finalize(s, ca, rc, phase, lane);
            group_barrier(ctl, epoch, gs);
            PT_MARK(PT_FINAL);
            if (r == 0 && tid == 0) {
                facc[2] += dn * dn * dn / 3;
                if (pred) { facc[0] += dn * dn * dn / 3 + 19 * dn * dn; facc[1] += dn * dn * dn / 3; nacc[1]++; }
                else { facc[0] += dn * dn * dn + 22 * dn * dn; facc[1] += dn * dn * dn; nacc[0]++; }
            }
        }
    }
    if (tid == 0) {
        for (int q = 0; q < PT_N; q++) atomicAdd((unsigned long long*)&P.acc->cycles[q], (unsigned long long)tacc[q]);
        if (r == 0) {
            atomicAdd(&P.acc->flops, facc[0]); atomicAdd(&P.acc->flops_factor, facc[1]); atomicAdd(&P.acc->flops_chol, facc[2]);
            atomicAdd((unsigned long long*)&P.acc->n_evals, (unsigned long long)nacc[0]);
            atomicAdd((unsigned long long*)&P.acc->n_pred, (unsigned long long)nacc[1]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
static bool g_attr_done = false;
static void set_attrs() {
    if (g_attr_done) return;
    cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_PIPE);
    cudaFuncSetAttribute(k_chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_BYTES);
    cudaFuncSetAttribute(k_chol_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_BYTES);
    cudaFuncSetAttribute(k_trtri, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_PIPE);
    cudaFuncSetAttribute(k_scale_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_BYTES);
    cudaFuncSetAttribute(k_lauum_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_PIPE);
    cudaFuncSetAttribute(k_gp_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_BYTES);
    g_attr_done = true;
}
static_assert(OI_SMEM_PIPE == PIPE_BYTES, "pipeline size");
static_assert(OI_SMEM_BYTES >= PIPE_BYTES && OI_SMEM_BYTES >= 2 * NB * TS * 8 + 4 * 64 * 8 + 16, "shared memory budget");
static_assert(PIPE_BYTES >= NB * TS * 8 + 4 * 64 * 8 + 16, "packed diagonal block must fit the pipeline buffers");

void oi_launch_count(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                     double t_lo, double t_hi, int* counts, cudaStream_t st) {
    k_gather<false><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, ot, n_obs, X, n_cells, r2, t_lo, t_hi, counts, nullptr, nullptr);
}
void oi_launch_scan(const int* counts, int n, long long* offsets, cudaStream_t st) {
    k_scan_counts<<<1, 1024, 0, st>>>(counts, n, offsets);
}
void oi_launch_fill(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                    double t_lo, double t_hi, const long long* offsets, int* indices, cudaStream_t st) {
    k_gather<true><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, ot, n_obs, X, n_cells, r2, t_lo, t_hi, nullptr, offsets, indices);
}
void oi_launch_pack(const int* indices, long long total, const double* ox, const double* oy, const double* ot,
                    const double* oz, double mean, double t_shift, double* px, double* py, double* pt, double* pr, cudaStream_t st) {
    if (total <= 0) return;
    k_pack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(indices, total, ox, oy, ot, oz, mean, t_shift, px, py, pt, pr);
}
// Slots are sorted by descending size, so the cells that own block row/column x are the prefix cnt_gt[x]
// (number of slots with N > x): every grid below is exact in the slot dimension.  Big batches launch
// the tile-parallel kernels one block row at a time (exact in both dimensions); small batches (the
// optimiser's tail) use one 2-D launch to save launch latency.
#define OI_ROWWISE_MIN_SLOTS OI_ROWWISE_MIN_SLOTS_HOST
#define SMALL_SMEM (8 * NB * 8)
void oi_launch_build(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st) {
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 0; i < Nmax; i++) k_build<<<dim3(i + 1, cnt_gt[i]), OI_THREADS, SMALL_SMEM, st>>>(slots, ca, pk, i);
    } else k_build<<<dim3(Nmax * (Nmax + 1) / 2, A), OI_THREADS, SMALL_SMEM, st>>>(slots, ca, pk, -1);
}
void oi_launch_chol_update(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st) {
    set_attrs();
    k_chol_update<<<dim3(k == 0 ? 1 : Nmax - k, cnt_gt[k]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, k);
}
void oi_launch_chol_panel(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st) {
    set_attrs();
    const int n_panel = cnt_gt[k + 1] > 0 ? Nmax - k - 1 : 0;     // panel tiles exist for cells with N > k+1, row k for N > k
    if (n_panel + k <= 0 || cnt_gt[k] <= 0) return;
    k_chol_panel<<<dim3(n_panel + k, cnt_gt[k]), OI_THREADS, OI_SMEM_BYTES, st>>>(slots, k, n_panel);
}
// all block columns in one launch; ticket must be zero (the caller resets it on the stream before the launch)
void oi_launch_chol_fused(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int* ticket, cudaStream_t st) {
    set_attrs();
    OiCholPlan plan;
    plan.Nmax = Nmax;
    int off = 0;
    for (int k = 0; k < Nmax; k++) {
        plan.off[2 * k] = off; off += cnt_gt[k];                              // diagonal tiles: cells with N > k
        plan.off[2 * k + 1] = off; off += cnt_gt[k + 1] * (Nmax - k - 1);     // off-diagonal: cells with N > k+1
    }
    plan.off[2 * Nmax] = off;
    if (off > 0) k_chol_fused<<<off, OI_THREADS, OI_SMEM_BYTES, st>>>(slots, plan, ticket);
}
void oi_launch_fwd(const OiSlot* slots, int A, OiCellArrays ca, OiPacked pk, double t_pred, cudaStream_t st) {
    k_fwd<<<A, 256, SMALL_SMEM, st>>>(slots, ca, pk, t_pred);
}
void oi_launch_trtri(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int d, const int* phase, cudaStream_t st) {
    set_attrs();
    if (Nmax - d <= 0 || cnt_gt[d] <= 0) return;
    k_trtri<<<dim3(Nmax - d, cnt_gt[d]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, phase, d);
}
void oi_launch_scale_rows(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, cudaStream_t st) {
    set_attrs();
    if (Nmax < 2) return;
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 1; i < Nmax; i++) k_scale_rows<<<dim3(i, cnt_gt[i]), OI_THREADS, OI_SMEM_BYTES, st>>>(slots, i);
    } else k_scale_rows<<<dim3(Nmax * (Nmax - 1) / 2, A), OI_THREADS, OI_SMEM_BYTES, st>>>(slots, -1);
}
void oi_launch_alpha(const OiSlot* slots, int A, int Nmax, const int* phase, cudaStream_t st) {
    k_alpha<<<dim3(Nmax, A), OI_THREADS, 0, st>>>(slots, phase);
}
void oi_launch_lauum_trace(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st) {
    set_attrs();
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 0; i < Nmax; i++)
            k_lauum_trace<<<dim3(i + 1, cnt_gt[i]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, ca, pk, i);
    } else k_lauum_trace<<<dim3(Nmax * (Nmax + 1) / 2, A), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, ca, pk, -1);
}
void oi_launch_cg_init(OiCellArrays ca, int n_cells, OiRunConst rc, cudaStream_t st) {
    k_cg_init<<<(n_cells + 127) / 128, 128, 0, st>>>(ca, n_cells, rc);
}
void oi_launch_finalize(const OiSlot* slots, int A, OiCellArrays ca, OiRunConst rc, int* slot_phase, cudaStream_t st) {
    k_finalize<<<(A * 32 + 127) / 128, 128, 0, st>>>(slots, A, ca, rc, slot_phase);
}
int oi_persistent_capacity() {
    set_attrs();
    int dev = 0, sms = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gp_persistent, OI_THREADS, OI_SMEM_BYTES);
    return sms * occ;
}
void oi_launch_persistent(const OiPersist& P, int n_groups, OiCellArrays ca, OiPacked pk, OiRunConst rc, double t_pred, cudaStream_t st) {
    set_attrs();
    k_gp_persistent<<<n_groups * P.gs, OI_THREADS, OI_SMEM_BYTES, st>>>(P, ca, pk, rc, t_pred);
}
