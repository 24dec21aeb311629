// Hand-written sm_100a kernels of the per-cell GP hot path.
//
//   k_count / k_fill / k_pack   neighbour gather            (GPR_CS2S3.py:159-164)
//   k_build                     Matern-3/2 ARD covariance   (GPR_CS2S3.py:78-105, :126)
//   k_chol_update / k_chol_panel  blocked left-looking Cholesky, FP64 DMMA tiles (np.linalg.cholesky, :126/:177)
//   k_fwd                       t = L^-1 (y - m), v = L^-1 k*   (:127, :178-180)
//   k_trtri                     U = L^-T by block distance, FP64 DMMA tiles      (explicit inverse of :130)
//   k_alpha                     alpha = U t
//   k_lauum_trace               K^-1 tiles = U U^T fused with the five trace terms of :131-138,
//                               dK/dtheta recomputed in registers, K^-1 never written
//
// All dense contractions are NT GEMM tiles (both operands K-contiguous) on mma.sync.m8n8k4.f64
// (SASS DMMA.8x8x4), fed by a 3-stage cp.async pipeline; tcgen05 has no FP64 kind.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "oi_types.h"
#include "oi_launch.h"

#define NB OI_NB
#define KT 16
#define LDS_ (KT + 4)          // smem row stride (doubles) of a streamed operand chunk: conflict-free DMMA fragment loads
#define STAGES 3
#define GEMM_THREADS 128
#define TS 68                  // smem row stride of a resident 64x64 tile
#define STAGE_DOUBLES (2 * NB * LDS_)
#define PIPE_BYTES (STAGES * STAGE_DOUBLES * 8)

#define ROOT3 1.7320508075688772   // np.sqrt(3.)
#ifdef OI_EXP_SAMEBLOCK
#define OI_FAILED(s) false
#else
#define OI_FAILED(s) (*(volatile int*)(s).fail != 0)
#endif

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

__device__ __forceinline__ void tile_ij(int t, int& i, int& j) {
    // lower-triangular tile enumeration t -> (i, j), j <= i, row by row
    int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((r + 1) * (r + 2) / 2 <= t) r++;
    while (r * (r + 1) / 2 > t) r--;
    i = r; j = t - r * (r + 1) / 2;
}

// Matern-3/2 pair quantities exactly in the reference's operation order (no FMA contraction):
// Q = sqrt(((dx*dx) + dy*dy) + dt*dt) of pre-scaled coordinates (scipy pdist 'euclidean').
__device__ __forceinline__ double pair_Q(double dx, double dy, double dt) {
    double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dt, dt));
    return sqrt(s);
}

// ------------------------------------------------------------------------------------------
// kernel (1): neighbour gather.  One warp per cell, observations staged through shared memory
// in chunks shared by the 8 cells of the CTA; ordered compaction by ballot + popc prefix.
// ------------------------------------------------------------------------------------------
#define G_CHUNK 2048
template <bool FILL>
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ ox, const double* __restrict__ oy, int n_obs,
                                                const double* __restrict__ X, int n_cells, double r2,
                                                int* __restrict__ counts, const long long* __restrict__ offsets,
                                                int* __restrict__ indices) {
    __shared__ double sx[G_CHUNK], sy[G_CHUNK];
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
        for (int q = threadIdx.x; q < m; q += 256) { sx[q] = ox[c0 + q]; sy[q] = oy[c0 + q]; }
        __syncthreads();
        if (live) {
            for (int q0 = 0; q0 < m; q0 += 32) {
                int q = q0 + lane;
                bool in = false;
                if (q < m) {
                    double dx = sx[q] - cx, dy = sy[q] - cy;
                    // inclusive boundary, no FMA: ties on the 25 km lattice resolve as in the reference (GPR_CS2S3.py:159)
                    in = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2;
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
                       double mean, double* __restrict__ px, double* __restrict__ py, double* __restrict__ pt,
                       double* __restrict__ pr) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int id = indices[i];
    px[i] = ox[id]; py[i] = oy[id]; pt[i] = ot[id]; pr[i] = oz[id] - mean;   // outputs - mX (GPR_CS2S3.py:127)
}

// ------------------------------------------------------------------------------------------
// kernel (2): covariance tiles.  K = sf2*(1+Q)exp(-Q) + sn2*I on the lower block triangle
// (GPR_CS2S3.py:93-94, :126); padding rows/cols are identity so every later tile is full.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_build(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, int row) {
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }      // one launch per block row (exact grids for big batches)
    else tile_ij(blockIdx.x, i, j);
    if (i >= s.N) return;
    if (i == 0 && threadIdx.x == 0) *s.fail = 0;
    __shared__ double ru[3][NB], cu[3][NB];
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    const double sf2 = h[3], sn2 = h[4];
    if (threadIdx.x < 2 * NB) {
        int which = threadIdx.x / NB, q = threadIdx.x % NB;
        int g = (which ? j : i) * NB + q;
        double ux = 0, uy = 0, ut = 0;
        if (g < s.n) {
            // np.sqrt(3.)*x/ell  (GPR_CS2S3.py:93): multiply, then divide
            ux = (ROOT3 * pk.x[s.pt_off + g]) / h[0];
            uy = (ROOT3 * pk.y[s.pt_off + g]) / h[1];
            ut = (ROOT3 * pk.t[s.pt_off + g]) / h[2];
        }
        double(*dst)[NB] = which ? cu : ru;
        dst[0][q] = ux; dst[1][q] = uy; dst[2][q] = ut;
    }
    __syncthreads();
    const long long ld = s.npad;
#pragma unroll 4
    for (int e = 0; e < TILE_PER_THREAD_256; e++) {
        int idx = threadIdx.x + e * 256;
        int r = idx / NB, c = idx % NB;
        int gi = i * NB + r, gj = j * NB + c;
        double val;
        if (gi >= s.n || gj >= s.n) val = (gi == gj) ? 1.0 : 0.0;
        else if (gi == gj) val = sf2 + sn2;
        else {
            double Q = pair_Q(ru[0][r] - cu[0][c], ru[1][r] - cu[1][c], ru[2][r] - cu[2][c]);
            // + np.eye(n)*sn2 off the diagonal is +0*sn2: NaN when sn2 overflowed (GPR_CS2S3.py:126)
            val = sf2 * ((1.0 + Q) * exp(-Q)) + 0.0 * sn2;
        }
        s.M[(long long)gi * ld + gj] = val;
    }
}

// ------------------------------------------------------------------------------------------
// FP64 DMMA tile core: acc(64x64) += A(64 x [k0,k1)) * B(64 x [k0,k1))^T, both K-contiguous.
// 4 warps (2x2), warp tile 32x32 = 4x4 m8n8k4 tiles, 3-stage cp.async pipeline of 16-wide chunks.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_stage(double* st, const double* __restrict__ A, long long lda,
                                           const double* __restrict__ B, long long ldb, int kk, int tid) {
    double* As = st;
    double* Bs = st + NB * LDS_;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        int c = tid + it * GEMM_THREADS;      // 0..511
        int row = c >> 3, col = (c & 7) * 2;
#ifdef OI_EXP_SAMEBLOCK
        // experiment: every chunk re-reads the first chunk of the panel (L1/L2 resident) -> compute-only bound
        cp_async16(&As[row * LDS_ + col], &A[(long long)row * lda + (kk & 0) + col]);
        cp_async16(&Bs[row * LDS_ + col], &B[(long long)row * ldb + (kk & 0) + col]);
#else
        cp_async16(&As[row * LDS_ + col], &A[(long long)row * lda + kk + col]);
        cp_async16(&Bs[row * LDS_ + col], &B[(long long)row * ldb + kk + col]);
#endif
    }
}

// Sub-tile ranges: a warp computes the 8x8 sub-tiles mb in [mlo, mhi) x nb in [nlo, nhi) of its 32x32
// warp tile for one K chunk.  All bounds are in {0, 2, 4} (structure comes in multiples of 16), so each
// combination is its own fully unrolled code path; ranges are warp-uniform.  They skip structural zeros
// (triangular diagonal blocks), the unused half of diagonal tiles and the rows/cols beyond the cell's real
// size in its last block.
struct SubRange { int mlo, mhi, nlo, nhi; };
__device__ __forceinline__ int clamp024(int v) { return v <= 0 ? 0 : (v >= 32 ? 4 : (v >= 16 ? 2 : 0)); }
// sub-tiles whose first row (col) is < limit / whose last row (col) is >= limit, limit a multiple of 16
__device__ __forceinline__ int hi_lt(int w, int limit) { return clamp024(limit - w * 32); }
__device__ __forceinline__ int lo_ge(int w, int limit) { return clamp024(limit - w * 32); }
#define SR_ALL SubRange{0, 4, 0, 4}

template <int MLO, int MHI, int NLO, int NHI>
__device__ __forceinline__ void mma_chunk_t(double (&acc)[4][4][2], const double* As, const double* Bs, int lda_s, int ldb_s,
                                            int wm, int wn, int lane, int kofs, int ksteps) {
    const int fr = lane >> 2, fc = lane & 3;
#pragma unroll
    for (int ks = 0; ks < ksteps; ks++) {
        double a[4], b[4];
#pragma unroll
        for (int mb = MLO; mb < MHI; mb++) a[mb] = As[(wm * 32 + mb * 8 + fr) * lda_s + kofs + ks * 4 + fc];
#pragma unroll
        for (int nb = NLO; nb < NHI; nb++) b[nb] = Bs[(wn * 32 + nb * 8 + fr) * ldb_s + kofs + ks * 4 + fc];
#pragma unroll
        for (int mb = MLO; mb < MHI; mb++)
#pragma unroll
            for (int nb = NLO; nb < NHI; nb++) dmma(acc[mb][nb], a[mb], b[nb]);
    }
}
template <int MLO, int MHI>
__device__ __forceinline__ void mma_chunk_n(double (&acc)[4][4][2], const double* As, const double* Bs, int lda_s, int ldb_s,
                                            int wm, int wn, int lane, int kofs, int ksteps, int nlo, int nhi) {
    if (nlo == 0 && nhi == 4) mma_chunk_t<MLO, MHI, 0, 4>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps);
    else if (nlo == 0 && nhi == 2) mma_chunk_t<MLO, MHI, 0, 2>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps);
    else if (nlo == 2 && nhi == 4) mma_chunk_t<MLO, MHI, 2, 4>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps);
}
__device__ __forceinline__ void mma_chunk(double (&acc)[4][4][2], const double* As, const double* Bs, int lda_s, int ldb_s,
                                          int wm, int wn, int lane, int kofs, int ksteps, SubRange r) {
    if (r.mlo == 0 && r.mhi == 4) mma_chunk_n<0, 4>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps, r.nlo, r.nhi);
    else if (r.mlo == 0 && r.mhi == 2) mma_chunk_n<0, 2>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps, r.nlo, r.nhi);
    else if (r.mlo == 2 && r.mhi == 4) mma_chunk_n<2, 4>(acc, As, Bs, lda_s, ldb_s, wm, wn, lane, kofs, ksteps, r.nlo, r.nhi);
}

template <class MaskFn>
__device__ __forceinline__ void gemm_nt_stream(double (&acc)[4][4][2], const double* __restrict__ A, long long lda,
                                               const double* __restrict__ B, long long ldb, int k0, int k1,
                                               double* smem, MaskFn maskfn) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int nk = (k1 - k0) / KT;
#pragma unroll
    for (int s = 0; s < STAGES - 1; s++) {
        if (s < nk) load_stage(smem + s * STAGE_DOUBLES, A, lda, B, ldb, k0 + s * KT, tid);
        cp_async_commit();
    }
    for (int it = 0; it < nk; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        int nx = it + STAGES - 1;
        if (nx < nk) load_stage(smem + (nx % STAGES) * STAGE_DOUBLES, A, lda, B, ldb, k0 + nx * KT, tid);
        cp_async_commit();
        const double* st = smem + (it % STAGES) * STAGE_DOUBLES;
        mma_chunk(acc, st, st + NB * LDS_, LDS_, LDS_, wm, wn, lane, 0, KT / 4, maskfn(k0 + it * KT));
    }
    cp_async_wait<0>();
    __syncthreads();
}

#define ACC_ZERO(acc)                                                        \
    _Pragma("unroll") for (int mb_ = 0; mb_ < 4; mb_++)                      \
    _Pragma("unroll") for (int nb_ = 0; nb_ < 4; nb_++) { acc[mb_][nb_][0] = 0.0; acc[mb_][nb_][1] = 0.0; }

// fragment element (mb, nb, e) of this thread sits at tile row/col:
#define FRAG_ROW(wm, mb, lane) ((wm) * 32 + (mb) * 8 + ((lane) >> 2))
#define FRAG_COL(wn, nb, lane) ((wn) * 32 + (nb) * 8 + (((lane) & 3) << 1))

// ------------------------------------------------------------------------------------------
// 64x64 diagonal block: Cholesky factor (lower, in place in T) and its inverse (W), both in shared
// memory, on 8x8 sub-blocks: the 8x8 pivot block is factored + inverted by one warp in registers
// (dpotf2 order of operations; a pivot <= 0 sets *s_bad, a NaN pivot propagates -- OpenBLAS potf2
// semantics, which is what np.linalg.cholesky runs), every other sub-block operation (panel solve,
// trailing update, inverse by block distance) is one or two DMMA m8n8k4 per 8x8 block.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void diag_factor_invert(double* T, double* W, double* sc, int* s_bad, int tid) {
    const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fc = lane & 3;
    for (int j = 0; j < 8; j++) {
        const int jb = j * 8;
        if (warp == 0) {
            const int r = lane & 7;
            double a[8];
#pragma unroll
            for (int c = 0; c < 8; c++) a[c] = T[(jb + r) * TS + jb + c];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 8; c++) {
#pragma unroll
                for (int l = 0; l < c; l++) {
                    double acl = __shfl_sync(0xffffffffu, a[l], c, 8);      // L[c][l]
                    if (r >= c) a[c] -= a[l] * acl;
                }
                double piv = __shfl_sync(0xffffffffu, a[c], c, 8);
                if (piv <= 0.0) bad = true;
                double sq = sqrt(piv), inv = 1.0 / sq;
                if (r == c) a[c] = sq;
                else if (r > c) a[c] *= inv;
            }
            if (bad) {
                if (lane == 0) *s_bad = 1;
            } else {
                if (lane < 8) {
#pragma unroll
                    for (int c = 0; c < 8; c++) if (c <= r) T[(jb + r) * TS + jb + c] = a[c];
                }
                __syncwarp();
                // X = L8^-1, lane b owns column b:  x[rr] = -(sum_{l<rr} L[rr][l] x[l]) / L[rr][rr]
                const int b = r;
                double x[8];
#pragma unroll
                for (int rr = 0; rr < 8; rr++) {
                    double sacc = 0.0;
#pragma unroll
                    for (int l = 0; l < rr; l++) sacc += T[(jb + rr) * TS + jb + l] * x[l];
                    double dinv = 1.0 / T[(jb + rr) * TS + jb + rr];
                    x[rr] = (rr < b) ? 0.0 : ((rr == b) ? dinv : -sacc * dinv);
                }
                if (lane < 8) {
#pragma unroll
                    for (int rr = 0; rr < 8; rr++) if (rr >= b) W[(jb + rr) * TS + jb + b] = x[rr];
                }
            }
        }
        __syncthreads();
#ifndef OI_EXP_SAMEBLOCK
        if (*s_bad) return;
#endif
        // panel: L_ij = A_ij * X_jj^T  (i > j)
        for (int i = j + 1 + warp; i < 8; i += 4) {
            double c2[2] = {0.0, 0.0};
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
                dmma(c2, T[(i * 8 + fr) * TS + jb + ks * 4 + fc], W[(jb + fr) * TS + jb + ks * 4 + fc]);
            __syncwarp();
            T[(i * 8 + fr) * TS + jb + fc * 2] = c2[0];
            T[(i * 8 + fr) * TS + jb + fc * 2 + 1] = c2[1];
        }
        __syncthreads();
        // trailing update: A_il -= L_ij L_lj^T  (j < l <= i)
        const int m = 7 - j, cnt = m * (m + 1) / 2;
        for (int q = warp; q < cnt; q += 4) {
            int ii, ll;
            tile_ij(q, ii, ll);
            const int i = j + 1 + ii, l = j + 1 + ll;
            double c2[2];
            c2[0] = T[(i * 8 + fr) * TS + l * 8 + fc * 2];
            c2[1] = T[(i * 8 + fr) * TS + l * 8 + fc * 2 + 1];
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
                dmma(c2, -T[(i * 8 + fr) * TS + jb + ks * 4 + fc], T[(l * 8 + fr) * TS + jb + ks * 4 + fc]);
            T[(i * 8 + fr) * TS + l * 8 + fc * 2] = c2[0];
            T[(i * 8 + fr) * TS + l * 8 + fc * 2 + 1] = c2[1];
        }
        __syncthreads();
    }
    // inverse by 8x8 block distance: W_ik = -X_ii * sum_{j=k}^{i-1} L_ij W_jk
    for (int d = 1; d < 8; d++) {
        for (int kb = warp; kb + d < 8; kb += 4) {
            const int i = kb + d;
            double c1[2] = {0.0, 0.0};
            for (int jj = kb; jj < i; jj++) {
#pragma unroll
                for (int ks = 0; ks < 2; ks++)
                    dmma(c1, T[(i * 8 + fr) * TS + jj * 8 + ks * 4 + fc], W[(jj * 8 + ks * 4 + fc) * TS + kb * 8 + fr]);
            }
            sc[fr * 8 + fc * 2] = c1[0];
            sc[fr * 8 + fc * 2 + 1] = c1[1];
            __syncwarp();
            double c2[2] = {0.0, 0.0};
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
                dmma(c2, W[(i * 8 + fr) * TS + i * 8 + ks * 4 + fc], sc[(ks * 4 + fc) * 8 + fr]);
            W[(i * 8 + fr) * TS + kb * 8 + fc * 2] = -c2[0];
            W[(i * 8 + fr) * TS + kb * 8 + fc * 2 + 1] = -c2[1];
            __syncwarp();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// kernel (3a): left-looking block-column update  A_ik -= sum_{j<k} L_ij L_kj^T  (i >= k);
// the CTA of the diagonal tile then factors it in shared memory (dpotrf semantics: a pivot
// <= 0 or NaN raises the cell's fail flag), inverts the 64x64 factor and stores
//   Dinv[k] = L_kk^-1 (row-major)   and   M(k,k) = U_kk = L_kk^-T (upper, zeros below).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS) k_chol_update(const OiSlot* __restrict__ slots, int k) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    const int i = k + blockIdx.x;
    if (i >= s.N) return;
    if (k == 0 && i != 0) return;          // nothing to subtract from the first block column
    if (OI_FAILED(s)) return;
    const long long ld = s.npad;
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    // the tile being updated is fetched up front so its latency hides behind the K loop
    double2 cin[4][4];
    {
        const double* Cr = s.M + (long long)i * NB * ld + (long long)k * NB;
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int nb = 0; nb < 4; nb++)
                cin[mb][nb] = *(const double2*)&Cr[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)];
    }
    {
        // rows of block i / cols of block k beyond the cell's size are padding; of the diagonal tile only
        // the lower triangle is needed (the warp above the diagonal idles)
        SubRange sr{0, hi_lt(wm, s.n16 - i * NB), 0, hi_lt(wn, s.n16 - k * NB)};
        if (i == k && wm < wn) sr.mhi = 0;
        gemm_nt_stream(acc, s.M + (long long)i * NB * ld, ld, s.M + (long long)k * NB * ld, ld, 0, k * NB, smem,
                       [sr](int) { return sr; });
    }
    double* Cg = s.M + (long long)i * NB * ld + (long long)k * NB;
    if (i != k) {
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                double2 v = cin[mb][nb];
                v.x -= acc[mb][nb][0]; v.y -= acc[mb][nb][1];
                *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)] = v;
            }
        return;
    }
    // ---- diagonal tile: T = A_kk - acc, factor + invert in shared memory ----
    double* T = smem;                    // [64][TS]  A_kk -> L_kk (lower)
    double* W = smem + NB * TS;          // [64][TS]  L_kk^-1 (lower, zeros above)
    double* sc = smem + 2 * NB * TS + warp * 64;   // per-warp 8x8 scratch
    __shared__ int s_bad;
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            int r = FRAG_ROW(wm, mb, lane), c = FRAG_COL(wn, nb, lane);
            T[r * TS + c] = cin[mb][nb].x - acc[mb][nb][0];
            T[r * TS + c + 1] = cin[mb][nb].y - acc[mb][nb][1];
        }
    for (int idx = tid; idx < NB * TS; idx += GEMM_THREADS) W[idx] = 0.0;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    diag_factor_invert(T, W, sc, &s_bad, tid);
#ifndef OI_EXP_SAMEBLOCK
    if (s_bad) {
        if (tid == 0) *s.fail = 1;
        return;
    }
#endif
    if (warp == 0) {
        // log-determinant part: sum_i log L_ii of this block (GPR_CS2S3.py:128), fixed order
        double v = log(T[lane * TS + lane]) + log(T[(lane + 32) * TS + lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s.part[k] = v;
    }
    double* Dk = s.Dinv + (long long)k * OI_TILE;
    for (int idx = tid; idx < OI_TILE; idx += GEMM_THREADS) {
        int r = idx >> 6, c = idx & 63;
        Dk[idx] = W[r * TS + c];
        Cg[(long long)r * ld + c] = (c >= r) ? W[c * TS + r] : 0.0;
    }
}

// kernel (3b): panel  L_ik = A_ik * L_kk^-T  (i > k), as an NT tile against Dinv[k]
__global__ void __launch_bounds__(GEMM_THREADS) k_chol_panel(const OiSlot* __restrict__ slots, int k) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    const int i = k + 1 + blockIdx.x;
    if (i >= s.N) return;
    if (OI_FAILED(s)) return;
    const long long ld = s.npad;
    double acc[4][4][2];
    ACC_ZERO(acc);
    double* Cg = s.M + (long long)i * NB * ld + (long long)k * NB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    {
        // Dinv[k][nn][kk] is lower triangular: output column nn only needs kk <= nn
        const int mhi = hi_lt(wm, s.n16 - i * NB);
        gemm_nt_stream(acc, Cg, ld, s.Dinv + (long long)k * OI_TILE, NB, 0, NB, smem,
                       [mhi, wn](int kk) { return SubRange{0, mhi, lo_ge(wn, kk), 4}; });
    }
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 v; v.x = acc[mb][nb][0]; v.y = acc[mb][nb][1];
            *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)] = v;
        }
}

// ------------------------------------------------------------------------------------------
// kernel (3c): forward substitution with the factor, one CTA per cell:
//   t = L^-1 (y - m)            (GPR_CS2S3.py:127 inner solve)
//   v = L^-1 k*   (predict)     (GPR_CS2S3.py:180)
// scalars: t.t (=> (y-m)^T alpha), v.t (=> k*^T alpha), v.v
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fwd(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, double t_pred) {
    const OiSlot s = slots[blockIdx.x];
    if (OI_FAILED(s)) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool pred = ca.phase[s.cell] == OI_PH_PREDICT;
    const int nrhs = pred ? 2 : 1;
    const long long ld = s.npad;
    double* tv = s.vec;              // t
    double* vv = s.vec + s.npad;     // v
    __shared__ double sb[2][NB];
    __shared__ double red[3][8];
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    // right-hand sides
    for (int g = tid; g < s.npad; g += 256) {
        double r = 0.0, ks = 0.0;
        if (g < s.n) {
            r = pk.r[s.pt_off + g];
            if (pred) {
                // cdist(sqrt(3)*x/ell, sqrt(3)*xs/ell) (GPR_CS2S3.py:100-101)
                double dx = (ROOT3 * pk.x[s.pt_off + g]) / h[0] - (ROOT3 * ca.X[2 * (size_t)s.cell]) / h[0];
                double dy = (ROOT3 * pk.y[s.pt_off + g]) / h[1] - (ROOT3 * ca.X[2 * (size_t)s.cell + 1]) / h[1];
                double dt = (ROOT3 * pk.t[s.pt_off + g]) / h[2] - (ROOT3 * t_pred) / h[2];
                double Q = pair_Q(dx, dy, dt);
                ks = h[3] * ((1.0 + Q) * exp(-Q));
            }
        }
        tv[g] = r; vv[g] = ks;
    }
    __syncthreads();
    for (int k = 0; k < s.N; k++) {
        const int kc = k * NB;
        // s[r] = b[kc+r] - sum_{c<kc} L[kc+r][c] * x[c]; warp w owns rows w*8 .. w*8+7
        double a0[8], a1[8];
#pragma unroll
        for (int q = 0; q < 8; q++) { a0[q] = 0.0; a1[q] = 0.0; }
        const double* Lrow = s.M + (long long)(kc + warp * 8) * ld;
        for (int c = lane; c < kc; c += 32) {
            double x0 = tv[c], x1 = pred ? vv[c] : 0.0;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                double l = Lrow[(long long)q * ld + c];
                a0[q] += l * x0; a1[q] += l * x1;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0[q] += __shfl_down_sync(0xffffffffu, a0[q], o);
                a1[q] += __shfl_down_sync(0xffffffffu, a1[q], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 8; q++) { sb[0][warp * 8 + q] = a0[q]; sb[1][warp * 8 + q] = a1[q]; }
        }
        // x_k = Dinv[k] * b_k - sum_{c<kc} Ls[kc+r][c] x[c]   (lower-triangular 64x64 mat-vec), thread (rhs, row)
        double dsum = 0.0;
        if (tid < NB * nrhs) {
            int rh = tid / NB, r = tid % NB;
            const double* D = s.Dinv + (long long)k * OI_TILE + r * NB;
            const double* bsrc = (rh ? vv : tv) + kc;
            for (int c = 0; c <= r; c++) dsum += D[c] * bsrc[c];
        }
        __syncthreads();
        if (tid < NB * nrhs) {
            int rh = tid / NB, r = tid % NB;
            (rh ? vv : tv)[kc + r] = dsum - sb[rh][r];
        }
        __syncthreads();
    }
    // scalars, fixed summation order
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    for (int g = tid; g < s.npad; g += 256) {
        double a = tv[g], b = pred ? vv[g] : 0.0;
        q0 += a * a; q1 += a * b; q2 += b * b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        q0 += __shfl_down_sync(0xffffffffu, q0, o);
        q1 += __shfl_down_sync(0xffffffffu, q1, o);
        q2 += __shfl_down_sync(0xffffffffu, q2, o);
    }
    if (lane == 0) { red[0][warp] = q0; red[1][warp] = q1; red[2][warp] = q2; }
    __syncthreads();
    if (tid < 3) {
        double a = 0.0;
        for (int w = 0; w < 8; w++) a += red[tid][w];
        s.part[s.N + tid] = a;
    }
}

// ------------------------------------------------------------------------------------------
// kernel (3c'): row scaling  Ls_ij = L_ii^-1 * L_ij  (i > j), in place, one launch for all tiles.
// With it the forward substitution and the inverse need no per-step triangular solve:
//   t_i = L_ii^-1 r_i - sum_{j<i} Ls_ij t_j            W_ik = -sum_{j=k}^{i-1} Ls_ij W_jk
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS) k_scale_rows(const OiSlot* __restrict__ slots, int row) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }
    else { tile_ij(blockIdx.x, i, j); i += 1; }   // strictly lower tiles: (i, j), 1 <= i < N, j < i
    if (i >= s.N) return;
    if (OI_FAILED(s)) return;
    const long long ld = s.npad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    double* TA = smem;             // Dinv_i [m][kk]
    double* TB = smem + NB * TS;   // L_ij   [kk][n]
    const double* Di = s.Dinv + (long long)i * OI_TILE;
    double* Lg = s.M + (long long)i * NB * ld + (long long)j * NB;
    for (int idx = tid; idx < OI_TILE / 2; idx += GEMM_THREADS) {
        int r = idx >> 5, c = (idx & 31) * 2;
        cp_async16(&TA[r * TS + c], &Di[r * NB + c]);
        cp_async16(&TB[r * TS + c], &Lg[(long long)r * ld + c]);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int vi = s.n16 - i * NB;              // valid rows of block i
    const int fr = lane >> 2, fc = lane & 3;
    // out[m][n] = sum_kk Dinv_i[m][kk] L_ij[kk][n], Dinv_i lower triangular: row m needs kk <= m
    for (int c = 0; c < NB && c < vi; c += KT) {
        const int mlo = lo_ge(wm, c), mhi = hi_lt(wm, vi);
#pragma unroll
        for (int ks = 0; ks < KT / 4; ks++) {
            double a[4], b[4];
#pragma unroll
            for (int mb = 0; mb < 4; mb++) a[mb] = TA[(wm * 32 + mb * 8 + fr) * TS + c + ks * 4 + fc];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) b[nb] = TB[(c + ks * 4 + fc) * TS + wn * 32 + nb * 8 + fr];
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
                if (mb >= mlo && mb < mhi) {
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) dmma(acc[mb][nb], a[mb], b[nb]);
                }
        }
    }
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 v; v.x = acc[mb][nb][0]; v.y = acc[mb][nb][1];
            *(double2*)&Lg[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)] = v;
        }
}

// ------------------------------------------------------------------------------------------
// kernel (3d): U = L^-T by block distance d:  W_ik = -sum_{j=k}^{i-1} Ls_ij W_jk, i = k+d,
// stored transposed (U[k-block][i-block] = W_ik^T) so every later contraction stays NT.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS) k_trtri(const OiSlot* __restrict__ slots, const int* __restrict__ phase, int d) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    const int kb = blockIdx.x, i = kb + d;
    if (i >= s.N) return;
    if (phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    const long long ld = s.npad;
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    {
        // first K block is U_kk (upper triangular): column nn of the output only needs kk' >= nn;
        // rows of block i beyond the cell's size are padding
        const int mhi = hi_lt(wm, s.n16 - i * NB);
        const int kfirst = kb * NB;
        gemm_nt_stream(acc, s.M + (long long)i * NB * ld, ld, s.M + (long long)kb * NB * ld, ld, kb * NB, i * NB, smem,
                       [mhi, wn, kfirst](int kk) {
                           int c = kk - kfirst;
                           return SubRange{0, mhi, 0, c < NB ? hi_lt(wn, c + KT) : 4};
                       });
    }
    // transpose through shared memory, then coalesced stores: U[kb*64+nn][i*64+m] = -acc[m][nn]
    double* TA = smem;             // [nn][m]
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            int r = FRAG_ROW(wm, mb, lane), c = FRAG_COL(wn, nb, lane);
            TA[c * TS + r] = -acc[mb][nb][0];
            TA[(c + 1) * TS + r] = -acc[mb][nb][1];
        }
    __syncthreads();
    double* Ug = s.M + (long long)kb * NB * ld + (long long)i * NB;
    for (int idx = tid; idx < OI_TILE / 2; idx += GEMM_THREADS) {
        int r = idx >> 5, c = (idx & 31) * 2;
        *(double2*)&Ug[(long long)r * ld + c] = *(const double2*)&TA[r * TS + c];
    }
}

// kernel (3e): alpha = K^-1 (y-m) = U t   (rows of U dotted with t), 64 rows per CTA
__global__ void __launch_bounds__(256) k_alpha(const OiSlot* __restrict__ slots, const int* __restrict__ phase) {
    const OiSlot s = slots[blockIdx.y];
    const int rb = blockIdx.x;
    if (rb >= s.N) return;
    if (phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ld = s.npad;
    const double* tv = s.vec;
    double* al = s.vec + 2 * (long long)s.npad;
    const int r0 = rb * NB + warp * 8;
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = 0.0;
    const double* Urow = s.M + (long long)r0 * ld;
    for (int c = rb * NB + lane; c < s.npad; c += 32) {
        double x = tv[c];
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] += Urow[(long long)q * ld + c] * x;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_down_sync(0xffffffffu, a[q], o);
        if (lane == 0) al[r0 + q] = a[q];
    }
}

// ------------------------------------------------------------------------------------------
// kernel (4): K^-1 tile (i,j) = sum_{m >= i} U_im U_jm^T on DMMA, fused with the trace terms of
// GPR_CS2S3.py:130-138:  Qm = K^-1 - alpha alpha^T,
//   S_theta = sum Qm * q_theta^2 exp(-Q)   (theta = x, y, t)      S_3 = sum Qm * (1+Q) exp(-Q)
//   S_4 = tr(Qm)
// dK/dtheta is recomputed from the coordinates in registers; K^-1 is never stored.
// Each tile writes five partial sums; off-diagonal tiles count twice (symmetry).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS) k_lauum_trace(const OiSlot* __restrict__ slots, OiCellArrays ca, OiPacked pk, int row) {
    extern __shared__ __align__(16) double smem[];
    const OiSlot s = slots[blockIdx.y];
    int i, j;
    if (row >= 0) { i = row; j = blockIdx.x; }
    else tile_ij(blockIdx.x, i, j);
    if (i >= s.N) return;
    if (ca.phase[s.cell] == OI_PH_PREDICT) return;
    if (OI_FAILED(s)) return;
    const long long ld = s.npad;
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    double praw[4] = {0, 0, 0, 0};              // this thread's point (row or col of the tile): x, y, t, alpha
    {
        int g = ((tid / NB) ? j : i) * NB + tid % NB;
        if (g < s.n) {
            praw[0] = pk.x[s.pt_off + g]; praw[1] = pk.y[s.pt_off + g]; praw[2] = pk.t[s.pt_off + g];
            praw[3] = s.vec[2 * (long long)s.npad + g];
        }
    }
    {
        // K range ends at the cell's real size (rounded to 16); the first K block is U_ii (upper triangular):
        // row m only needs kk' >= m (for the diagonal tile likewise column n, and the warp above the diagonal idles)
        const int mhi0 = (i == j && wm < wn) ? 0 : hi_lt(wm, s.n16 - i * NB), nhi0 = hi_lt(wn, s.n16 - j * NB);
        const int kfirst = i * NB;
        const bool diag = (i == j);
        gemm_nt_stream(acc, s.M + (long long)i * NB * ld, ld, s.M + (long long)j * NB * ld, ld, i * NB, s.n16, smem,
                       [mhi0, nhi0, wm, wn, kfirst, diag](int kk) {
                           int c = kk - kfirst;
                           if (c >= NB) return SubRange{0, mhi0, 0, nhi0};
                           int mh = min(mhi0, hi_lt(wm, c + KT));
                           int nh = diag ? min(nhi0, hi_lt(wn, c + KT)) : nhi0;
                           return SubRange{0, mh, 0, nh};
                       });
    }
    // per-point data of the 64 rows and 64 cols: u (3), v (3), alpha (raw values were fetched before the K loop)
    double(*P)[7][NB] = (double(*)[7][NB])smem;    // P[0]=rows, P[1]=cols
    {
        int which = tid / NB, q = tid % NB;        // 128 threads: rows then cols
        double x = praw[0], y = praw[1], t = praw[2], a = praw[3];
        P[which][0][q] = (ROOT3 * x) / h[0]; P[which][1][q] = (ROOT3 * y) / h[1]; P[which][2][q] = (ROOT3 * t) / h[2];
        // np.sqrt(3.)*(x[:,theta]/ell[theta])  (GPR_CS2S3.py:97): divide, then multiply
        P[which][3][q] = ROOT3 * (x / h[0]); P[which][4][q] = ROOT3 * (y / h[1]); P[which][5][q] = ROOT3 * (t / h[2]);
        P[which][6][q] = a;
    }
    __syncthreads();
    double S[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        const int r = FRAG_ROW(wm, mb, lane), gi = i * NB + r;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = FRAG_COL(wn, nb, lane) + e, gj = j * NB + c;
                if (gi < s.n && gj < s.n && gj <= gi) {
                    double Qm = acc[mb][nb][e] - P[0][6][r] * P[1][6][c];
                    if (gi == gj) {
                        // Q = 0: dK_theta = 0, K = sf2
                        S[3] += Qm; S[4] += Qm;
                    } else {
                        Qm *= 2.0;          // (gi, gj) and (gj, gi): K^-1, alpha alpha^T and dK are symmetric
                        double Q = pair_Q(P[0][0][r] - P[1][0][c], P[0][1][r] - P[1][1][c], P[0][2][r] - P[1][2][c]);
                        double E = exp(-Q);
                        double qx = P[0][3][r] - P[1][3][c], qy = P[0][4][r] - P[1][4][c], qt = P[0][5][r] - P[1][5][c];
                        S[0] += Qm * (qx * qx * E);
                        S[1] += Qm * (qy * qy * E);
                        S[2] += Qm * (qt * qt * E);
                        S[3] += Qm * ((1.0 + Q) * E);
                    }
                }
            }
        }
    }
    __shared__ double red[5][4];
#pragma unroll
    for (int q = 0; q < 5; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) S[q] += __shfl_down_sync(0xffffffffu, S[q], o);
        if (lane == 0) red[q][warp] = S[q];
    }
    __syncthreads();
    if (tid < 5) {
        double v = ((red[tid][0] + red[tid][1]) + red[tid][2]) + red[tid][3];
        s.part[s.N + 8 + 5 * (long long)(i * (i + 1) / 2 + j) + tid] = v;
    }
}

// ------------------------------------------------------------------------------------------
// launch wrappers
// ------------------------------------------------------------------------------------------
static bool g_attr_done = false;
static void set_attrs() {
    if (g_attr_done) return;
    cudaFuncSetAttribute(k_chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_CHOL);
    cudaFuncSetAttribute(k_chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES);
    cudaFuncSetAttribute(k_trtri, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES);
    cudaFuncSetAttribute(k_scale_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, OI_SMEM_TRTRI);
    cudaFuncSetAttribute(k_lauum_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES);
    g_attr_done = true;
}

void oi_launch_count(const double* ox, const double* oy, int n_obs, const double* X, int n_cells, double r2, int* counts,
                     cudaStream_t st) {
    k_gather<false><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, n_obs, X, n_cells, r2, counts, nullptr, nullptr);
}
void oi_launch_scan(const int* counts, int n, long long* offsets, cudaStream_t st) {
    k_scan_counts<<<1, 1024, 0, st>>>(counts, n, offsets);
}
void oi_launch_fill(const double* ox, const double* oy, int n_obs, const double* X, int n_cells, double r2,
                    const long long* offsets, int* indices, cudaStream_t st) {
    k_gather<true><<<(n_cells + 7) / 8, 256, 0, st>>>(ox, oy, n_obs, X, n_cells, r2, nullptr, offsets, indices);
}
void oi_launch_pack(const int* indices, long long total, const double* ox, const double* oy, const double* ot,
                    const double* oz, double mean, double* px, double* py, double* pt, double* pr, cudaStream_t st) {
    if (total <= 0) return;
    k_pack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(indices, total, ox, oy, ot, oz, mean, px, py, pt, pr);
}
// Slots are sorted by descending size, so the cells that own block row/column x are the prefix cnt_gt[x]
// (number of slots with N > x): every grid below is exact in the slot dimension.  Big batches launch
// the tile-parallel kernels one block row at a time (exact in both dimensions); small batches (the
// optimiser's tail) use one 2-D launch to save launch latency.
#define OI_ROWWISE_MIN_SLOTS OI_ROWWISE_MIN_SLOTS_HOST
void oi_launch_build(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st) {
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 0; i < Nmax; i++) k_build<<<dim3(i + 1, cnt_gt[i]), 256, 0, st>>>(slots, ca, pk, i);
    } else k_build<<<dim3(Nmax * (Nmax + 1) / 2, A), 256, 0, st>>>(slots, ca, pk, -1);
}
void oi_launch_chol_update(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st) {
    set_attrs();
    k_chol_update<<<dim3(k == 0 ? 1 : Nmax - k, cnt_gt[k]), GEMM_THREADS, OI_SMEM_CHOL, st>>>(slots, k);
}
void oi_launch_chol_panel(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st) {
    set_attrs();
    if (Nmax - k - 1 <= 0 || cnt_gt[k + 1] <= 0) return;
    k_chol_panel<<<dim3(Nmax - k - 1, cnt_gt[k + 1]), GEMM_THREADS, PIPE_BYTES, st>>>(slots, k);
}
void oi_launch_fwd(const OiSlot* slots, int A, OiCellArrays ca, OiPacked pk, double t_pred, cudaStream_t st) {
    k_fwd<<<A, 256, 0, st>>>(slots, ca, pk, t_pred);
}
void oi_launch_trtri(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int d, const int* phase, cudaStream_t st) {
    set_attrs();
    if (Nmax - d <= 0 || cnt_gt[d] <= 0) return;
    k_trtri<<<dim3(Nmax - d, cnt_gt[d]), GEMM_THREADS, PIPE_BYTES, st>>>(slots, phase, d);
}
void oi_launch_scale_rows(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, cudaStream_t st) {
    set_attrs();
    if (Nmax < 2) return;
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 1; i < Nmax; i++) k_scale_rows<<<dim3(i, cnt_gt[i]), GEMM_THREADS, OI_SMEM_TRTRI, st>>>(slots, i);
    } else k_scale_rows<<<dim3(Nmax * (Nmax - 1) / 2, A), GEMM_THREADS, OI_SMEM_TRTRI, st>>>(slots, -1);
}
void oi_launch_alpha(const OiSlot* slots, int A, int Nmax, const int* phase, cudaStream_t st) {
    k_alpha<<<dim3(Nmax, A), 256, 0, st>>>(slots, phase);
}
void oi_launch_lauum_trace(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st) {
    set_attrs();
    if (A >= OI_ROWWISE_MIN_SLOTS) {
        for (int i = 0; i < Nmax; i++)
            k_lauum_trace<<<dim3(i + 1, cnt_gt[i]), GEMM_THREADS, PIPE_BYTES, st>>>(slots, ca, pk, i);
    } else k_lauum_trace<<<dim3(Nmax * (Nmax + 1) / 2, A), GEMM_THREADS, PIPE_BYTES, st>>>(slots, ca, pk, -1);
}
