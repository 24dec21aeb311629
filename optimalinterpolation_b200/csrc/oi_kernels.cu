// Hand-written sm_100a kernels of the per-cell GP hot path: the neighbour gather, the
// lockstep kernels over the tile functions of oi_tiles.cuh, and their launch wrappers.
//
//   k_gather / k_scan_counts / k_pack   neighbour gather        (GPR_CS2S3.py:159-164)
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
                                                OiRanges rg, int* __restrict__ counts, const long long* __restrict__ offsets,
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
    // Only the index ranges that can hold observations of the day window are scanned (ascending, so the neighbour order
    // is unchanged): a resident season is stream-major / day-major, i.e. one contiguous range per stream.
    for (int ri = 0; ri < rg.n; ri++)
    for (int c0 = rg.lo[ri]; c0 < rg.hi[ri]; c0 += G_CHUNK) {
        int m = min(G_CHUNK, rg.hi[ri] - c0);
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
__global__ void __launch_bounds__(OI_THREADS, OI_PIPE_MINB) k_chol_panel(const OiSlot* __restrict__ slots, int k, int n_panel) {
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


__global__ void __launch_bounds__(256) k_fwd(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, double t_pred) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.x];
    if (OI_FAILED(s)) return;
    cell_fwd(s, ca, pk, t_pred, ca.phase[s.cell] == OI_PH_PREDICT, smem);
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
// launch wrappers
// ------------------------------------------------------------------------------------------
// cudaFuncSetAttribute applies to the current device's context only: called once per device from oi_create (which holds
// the device current), so a process that opens handles on several GPUs gets the > 48 KB opt-in on each of them.
int oi_set_kernel_attributes() {
    cudaError_t e = cudaSuccess;
    auto set = [&](const void* f, int bytes) {
        cudaError_t r = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
        if (e == cudaSuccess) e = r;
    };
    set((const void*)k_chol_update, OI_SMEM_PIPE);
    set((const void*)k_chol_panel, OI_SMEM_PIPE);
    set((const void*)k_trtri, OI_SMEM_PIPE);
    set((const void*)k_lauum_trace, OI_SMEM_PIPE);
    return (int)e;
}
static_assert(OI_SMEM_PIPE == PIPE_BYTES, "pipeline size");
static_assert(PIPE_BYTES >= NB * TS * 8 + 4 * 64 * 8 + 16, "packed diagonal block must fit the pipeline buffers");

void oi_launch_count(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                     double t_lo, double t_hi, const OiRanges& rg, int* counts, cudaStream_t st) {
    k_gather<false><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, ot, n_obs, X, n_cells, r2, t_lo, t_hi, rg, counts, nullptr, nullptr);
}
void oi_launch_scan(const int* counts, int n, long long* offsets, cudaStream_t st) {
    k_scan_counts<<<1, 1024, 0, st>>>(counts, n, offsets);
}
void oi_launch_fill(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                    double t_lo, double t_hi, const OiRanges& rg, const long long* offsets, int* indices, cudaStream_t st) {
    k_gather<true><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, ot, n_obs, X, n_cells, r2, t_lo, t_hi, rg, nullptr, offsets, indices);
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
    k_chol_update<<<dim3(k == 0 ? 1 : Nmax - k, cnt_gt[k]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, k);
}
void oi_launch_chol_panel(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st) {
    const int n_panel = cnt_gt[k + 1] > 0 ? Nmax - k - 1 : 0;     // panel tiles exist for cells with N > k+1, row k for N > k
    if (n_panel + k <= 0 || cnt_gt[k] <= 0) return;
    k_chol_panel<<<dim3(n_panel + k, cnt_gt[k]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, k, n_panel);
}
void oi_launch_fwd(const OiSlot* slots, int A, OiCellArrays ca, OiPacked pk, double t_pred, cudaStream_t st) {
    k_fwd<<<A, 256, SMALL_SMEM, st>>>(slots, ca, pk, t_pred);
}
void oi_launch_trtri(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int d, const int* phase, cudaStream_t st) {
    if (Nmax - d <= 0 || cnt_gt[d] <= 0) return;
    k_trtri<<<dim3(Nmax - d, cnt_gt[d]), OI_THREADS, OI_SMEM_PIPE, st>>>(slots, phase, d);
}
void oi_launch_alpha(const OiSlot* slots, int A, int Nmax, const int* phase, cudaStream_t st) {
    k_alpha<<<dim3(Nmax, A), OI_THREADS, 0, st>>>(slots, phase);
}
void oi_launch_lauum_trace(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st) {
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
