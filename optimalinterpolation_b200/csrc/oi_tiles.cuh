// Tile-level device functions of the per-cell GP hot path.  Every function is executed
// by ONE CTA of OI_THREADS (128) threads on one 64x64 tile (or one cell for the substitutions) and is shared
// by the two execution engines:
//   * the persistent group kernel  (oi_kernels.cu: k_gp_persistent): a group of CTAs walks one cell through a whole
//     NLML+gradient evaluation, synchronising through a global-memory group barrier;
//   * the lockstep launch-per-step kernels (oi_kernels.cu): one launch per algorithmic step over all cells.
// Both engines therefore produce bit-identical numbers (fixed reduction orders, no atomics on data).
// The translation unit is compiled with -fmad=false (the optimiser restatement needs NumPy's rounding);
// the multiply-adds of the hot scalar loops are therefore written as explicit fma().
//
//   tile_build                  Matern-3/2 ARD covariance   (GPR_CS2S3.py:78-105, :126)
//   tile_chol_update / _panel   blocked left-looking Cholesky, FP64 DMMA tiles (np.linalg.cholesky, :126/:177)
//   tile_scale                  Ls_ij = L_ii^-1 L_ij
//   cell_fwd                    t = L^-1 (y - m), v = L^-1 k*   (:127, :178-180)
//   tile_trtri                  U = L^-T by block distance, FP64 DMMA tiles      (explicit inverse of :130)
//   rows_alpha                  alpha = U t
//   tile_lauum_trace            K^-1 tiles = U U^T fused with the five trace terms of :131-138,
//                               dK/dtheta recomputed in registers, K^-1 never written
//
// All dense contractions are NT GEMM tiles (both operands K-contiguous) on mma.sync.m8n8k4.f64
// (SASS DMMA.8x8x4), fed by a 3-stage cp.async pipeline; tcgen05 has no FP64 kind.
//
// Memory-model rule (persistent engine): data written by one CTA and read by another inside the same
// launch must not be served from a stale L1 line, so every load of MUTABLE global data goes through
// L2: cp.async.cg, __ldcg or volatile.  Immutable inputs (packed coordinates, slot table) use plain loads.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "oi_types.h"
#include "oi_launch.h"

#define NB OI_NB
#define KT 16
#ifndef STAGES
#define STAGES 3
#endif
#define GEMM_THREADS OI_THREADS
#define TS 68                  // smem row stride of a resident 64x64 tile
#define STAGE_DOUBLES (2 * NB * KT)   // operand chunks are stored unpadded (64 rows x 128 B) with an XOR swizzle
#define PIPE_BYTES (STAGES * STAGE_DOUBLES * 8)

#define ROOT3 1.7320508075688772   // np.sqrt(3.)
#define OI_FAILED(s) (*(volatile int*)(s).fail != 0)

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
// kernel (2): covariance tile.  K = sf2*(1+Q)exp(-Q) + sn2*I on the lower block triangle
// (GPR_CS2S3.py:93-94, :126); padding rows/cols are identity so every later tile is full.
// smem: 6*NB doubles.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_build(const OiSlot& s, const OiCellArrays& ca, const OiPacked& pk, int i, int j, bool want_qe,
                                           double* smem) {
    const int tid = threadIdx.x;
    if (i == 0 && tid == 0) *s.fail = 0;
    double(*ru)[NB] = (double(*)[NB])smem;
    double(*cu)[NB] = ru + 3;
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    const double h0 = __ldcg(h + 0), h1 = __ldcg(h + 1), h2 = __ldcg(h + 2), sf2 = __ldcg(h + 3), sn2 = __ldcg(h + 4);
    __syncthreads();
    {
        int which = tid / NB, q = tid % NB;
        int g = (which ? j : i) * NB + q;
        double ux = 0, uy = 0, ut = 0;
        if (g < s.n) {
            // np.sqrt(3.)*x/ell  (GPR_CS2S3.py:93): multiply, then divide
            ux = (ROOT3 * pk.x[s.pt_off + g]) / h0;
            uy = (ROOT3 * pk.y[s.pt_off + g]) / h1;
            ut = (ROOT3 * pk.t[s.pt_off + g]) / h2;
        }
        double(*dst)[NB] = which ? cu : ru;
        dst[0][q] = ux; dst[1][q] = uy; dst[2][q] = ut;
    }
    __syncthreads();
    const long long ld = s.npad;
    double* qe = s.QE + 2 * (long long)OI_TILE * (i * (i + 1) / 2 + j);
#pragma unroll 4
    for (int e = 0; e < OI_TILE / OI_THREADS; e++) {
        int idx = tid + e * OI_THREADS;
        int r = idx / NB, c = idx % NB;
        int gi = i * NB + r, gj = j * NB + c;
        double val, Q = 0.0, E = 1.0;
        if (gi >= s.n || gj >= s.n) val = (gi == gj) ? 1.0 : 0.0;
        else if (gi == gj) val = sf2 + sn2;
        else {
            Q = pair_Q(ru[0][r] - cu[0][c], ru[1][r] - cu[1][c], ru[2][r] - cu[2][c]);
            E = exp(-Q);
            // + np.eye(n)*sn2 off the diagonal is +0*sn2: NaN when sn2 overflowed (GPR_CS2S3.py:126)
            val = sf2 * ((1.0 + Q) * E) + 0.0 * sn2;
        }
        s.M[(long long)gi * ld + gj] = val;
        if (want_qe) { qe[idx] = Q; qe[OI_TILE + idx] = E; }
    }
}

// ------------------------------------------------------------------------------------------
// FP64 DMMA tile core: acc(64x64) += A(64 x [k0,k1)) * B(64 x [k0,k1))^T, both K-contiguous.
// 4 warps (2x2), warp tile 32x32 = 4x4 m8n8k4 tiles, STAGES-deep cp.async pipeline of 16-wide chunks.
// ------------------------------------------------------------------------------------------
// Shared-memory layout of a streamed 64x16 operand chunk: row r occupies 128 B; its eight 16-byte pieces are
// stored at piece index (p ^ (r & 7)).  cp.async writes stay 16-byte aligned and the DMMA fragment loads
// (8 rows x 4 consecutive doubles per instruction) touch every bank pair exactly twice = the two wavefronts
// a 64-bit warp load needs anyway: conflict-free without padding, so a stage is 16 KB instead of 20 KB.
__device__ __forceinline__ void load_stage(double* st, const double* __restrict__ A, long long lda,
                                           const double* __restrict__ B, long long ldb, int kk, int tid) {
    double* As = st;
    double* Bs = st + NB * KT;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        int c = tid + it * GEMM_THREADS;      // 0..511
        int row = c >> 3, p = c & 7;
        int dst = row * KT + ((p ^ (row & 7)) << 1);
        cp_async16(&As[dst], &A[(long long)row * lda + kk + p * 2]);
        cp_async16(&Bs[dst], &B[(long long)row * ldb + kk + p * 2]);
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

// (mma.sync m16n8k16.f64 was tried: ptxas lowers it to eight DMMA.8x8x4 on sm_100a -- the hardware shape --
// so it brings no instruction-count saving; the m8n8k4 form is kept.)
template <int MLO, int MHI, int NLO, int NHI>
__device__ __forceinline__ void mma_chunk_t(double (&acc)[4][4][2], const double* As, const double* Bs, int wm, int wn, int lane) {
    const int fr = lane >> 2, fc = lane & 3;
    const double* Ar = As + (wm * 32 + fr) * KT;
    const double* Br = Bs + (wn * 32 + fr) * KT;
    // column ks*4+fc of a row with (row & 7) == fr sits at swizzled offset sw[ks]
    int sw[KT / 4];
#pragma unroll
    for (int ks = 0; ks < KT / 4; ks++) sw[ks] = (((ks * 2 + (fc >> 1)) ^ fr) << 1) + (fc & 1);
#pragma unroll
    for (int ks = 0; ks < KT / 4; ks++) {
        double a[4], b[4];
#pragma unroll
        for (int mb = MLO; mb < MHI; mb++) a[mb] = Ar[mb * 8 * KT + sw[ks]];
#pragma unroll
        for (int nb = NLO; nb < NHI; nb++) b[nb] = Br[nb * 8 * KT + sw[ks]];
#pragma unroll
        for (int mb = MLO; mb < MHI; mb++)
#pragma unroll
            for (int nb = NLO; nb < NHI; nb++) dmma(acc[mb][nb], a[mb], b[nb]);
    }
}
template <int MLO, int MHI>
__device__ __forceinline__ void mma_chunk_n(double (&acc)[4][4][2], const double* As, const double* Bs, int wm, int wn, int lane,
                                            int nlo, int nhi) {
    if (nlo == 0 && nhi == 4) mma_chunk_t<MLO, MHI, 0, 4>(acc, As, Bs, wm, wn, lane);
    else if (nlo == 0 && nhi == 2) mma_chunk_t<MLO, MHI, 0, 2>(acc, As, Bs, wm, wn, lane);
    else if (nlo == 2 && nhi == 4) mma_chunk_t<MLO, MHI, 2, 4>(acc, As, Bs, wm, wn, lane);
}
__device__ __forceinline__ void mma_chunk(double (&acc)[4][4][2], const double* As, const double* Bs, int wm, int wn, int lane, SubRange r) {
    if (r.mlo == 0 && r.mhi == 4) mma_chunk_n<0, 4>(acc, As, Bs, wm, wn, lane, r.nlo, r.nhi);
    else if (r.mlo == 0 && r.mhi == 2) mma_chunk_n<0, 2>(acc, As, Bs, wm, wn, lane, r.nlo, r.nhi);
    else if (r.mlo == 2 && r.mhi == 4) mma_chunk_n<2, 4>(acc, As, Bs, wm, wn, lane, r.nlo, r.nhi);
}

// The pipeline starts with a CTA barrier (the shared-memory buffers may still be in use by the previous tile
// of a persistent CTA) and ends with one (they are free again on return).
template <class MaskFn>
__device__ __forceinline__ void gemm_nt_stream(double (&acc)[4][4][2], const double* __restrict__ A, long long lda,
                                               const double* __restrict__ B, long long ldb, int k0, int k1,
                                               double* smem, MaskFn maskfn) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 1, wn = warp & 1;
    const int nk = (k1 - k0) / KT;
    __syncthreads();
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
        mma_chunk(acc, st, st + NB * KT, wm, wn, lane, maskfn(k0 + it * KT));
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
// The diagonal 64x64 block lives in ONE shared-memory array P[64][TS]: the Cholesky factor L (lower triangle incl.
// the diagonal) at P[r][c], c <= r, and its inverse W = L^-1 (also lower triangular) transposed into the strict upper
// part shifted by one column: W(r, c) = P[c][r + 1], r >= c.  37 KB instead of two 35 KB arrays, so the kernel's
// shared memory is the 48 KB operand pipeline and 4 CTAs fit an SM (the CTAs of diagonal tiles keep one warp busy
// for ~20 us; more co-resident CTAs keep the DMMA pipe fed meanwhile).
__device__ __forceinline__ double& Wat(double* P, int r, int c) { return P[c * TS + r + 1]; }                  // r >= c
__device__ __forceinline__ double Wval(const double* P, int r, int c) { return c <= r ? P[c * TS + r + 1] : 0.0; }

// 8x8 pivot block (jb, jb): Cholesky factor in place (lower) and its inverse, by one warp in registers
// (lane r = row r; the four groups of 8 lanes do the same work).  dpotf2 order of operations.
__device__ __forceinline__ void pivot_factor_invert(double* P, int jb, int lane, int* s_bad) {
    const int r = lane & 7;
    double a[8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[c] = P[(jb + r) * TS + jb + c];         // entries c > r are loaded but never used
    bool bad = false;
    double dinv[8];                                                  // 1 / L[c][c], known to every lane
#pragma unroll
    for (int c = 0; c < 8; c++) {
#pragma unroll
        for (int l = 0; l < c; l++) {
            double acl = __shfl_sync(0xffffffffu, a[l], c, 8);      // L[c][l]
            if (r >= c) a[c] = fma(-a[l], acl, a[c]);
        }
        double piv = __shfl_sync(0xffffffffu, a[c], c, 8);
        if (piv <= 0.0) bad = true;
        // 1/sqrt(piv) first, then sqrt(piv) = piv * (1/sqrt(piv)): the 64 pivots of a diagonal block are one dependent chain
        // (the critical path of the optimiser's tail, where a handful of small cells iterate at kernel latency), and
        // sqrt followed by a division is the longest link of it.  rsqrt is within 1 ulp, the product adds half an ulp: far inside
        // the 1e-9 parity gate (measured 1e-14), not IEEE sqrt.  piv = +inf gives NaN here and inf in LAPACK; both end in
        // a NaN objective.
        double inv = rsqrt(piv), sq = piv * inv;
        dinv[c] = inv;
        if (r == c) a[c] = sq;
        else if (r > c) a[c] *= inv;
    }
    if (bad) {
        if (lane == 0) *s_bad = 1;
        return;
    }
    if (lane < 8) {
#pragma unroll
        for (int c = 0; c < 8; c++) if (c <= r) P[(jb + r) * TS + jb + c] = a[c];
    }
    __syncwarp();
    // X = L8^-1, lane b owns column b:  x[rr] = -(sum_{l<rr} L[rr][l] x[l]) / L[rr][rr]
    const int b = r;
    double x[8];
#pragma unroll
    for (int rr = 0; rr < 8; rr++) {
        double sacc = 0.0;
#pragma unroll
        for (int l = 0; l < rr; l++) sacc = fma(P[(jb + rr) * TS + jb + l], x[l], sacc);
        x[rr] = (rr < b) ? 0.0 : ((rr == b) ? dinv[rr] : -sacc * dinv[rr]);
    }
    if (lane < 8) {
#pragma unroll
        for (int rr = 0; rr < 8; rr++) if (rr >= b) Wat(P, jb + rr, jb + b) = x[rr];
    }
}

// trailing update of one 8x8 block of the diagonal tile: A_il -= L_ij L_lj^T
__device__ __forceinline__ void diag_trailing_block(double* T, int i, int l, int jb, int fr, int fc) {
    double c2[2];
    c2[0] = T[(i * 8 + fr) * TS + l * 8 + fc * 2];
    c2[1] = T[(i * 8 + fr) * TS + l * 8 + fc * 2 + 1];
#pragma unroll
    for (int ks = 0; ks < 2; ks++)
        dmma(c2, -T[(i * 8 + fr) * TS + jb + ks * 4 + fc], T[(l * 8 + fr) * TS + jb + ks * 4 + fc]);
    T[(i * 8 + fr) * TS + l * 8 + fc * 2] = c2[0];
    T[(i * 8 + fr) * TS + l * 8 + fc * 2 + 1] = c2[1];
}

// The 64 pivots are one dependent chain run by warp 0 (the critical path of the kernel's diagonal CTAs), so the
// chain is overlapped with the DMMA work: after the panel of step j, warp 0 updates only the next pivot block and
// factors it (look-ahead) while warps 1-3 apply the rest of step j's trailing update.  (The trailing update of an
// 8x8 diagonal sub-block (l == i, i > j+1) also rewrites that sub-block's strict upper part: it holds no W yet --
// X_ii is written when block i becomes the pivot -- and nothing reads it.)
__device__ __forceinline__ void diag_factor_invert(double* P, double* sc, int* s_bad, int tid) {
    const int warp = tid >> 5, lane = tid & 31, fr = lane >> 2, fc = lane & 3;
    if (warp == 0) pivot_factor_invert(P, 0, lane, s_bad);
    __syncthreads();
    if (*s_bad) return;
    for (int j = 0; j < 8; j++) {
        const int jb = j * 8;
        // panel: L_ij = A_ij * X_jj^T  (i > j); X_jj[n][k] = W(jb+n, jb+k), zero for k > n
        for (int i = j + 1 + warp; i < 8; i += 4) {
            double c2[2] = {0.0, 0.0};
#pragma unroll
            for (int ks = 0; ks < 2; ks++)
                dmma(c2, P[(i * 8 + fr) * TS + jb + ks * 4 + fc], Wval(P, jb + fr, jb + ks * 4 + fc));
            __syncwarp();
            P[(i * 8 + fr) * TS + jb + fc * 2] = c2[0];
            P[(i * 8 + fr) * TS + jb + fc * 2 + 1] = c2[1];
        }
        __syncthreads();
        if (j == 7) break;
        // trailing update: A_il -= L_ij L_lj^T  (j < l <= i); block q = 0 is the next pivot block (j+1, j+1)
        const int m = 7 - j, cnt = m * (m + 1) / 2;
        if (warp == 0) {
            diag_trailing_block(P, j + 1, j + 1, jb, fr, fc);
            __syncwarp();
            pivot_factor_invert(P, jb + 8, lane, s_bad);
        } else {
            for (int q = warp; q < cnt; q += 3) {
                int ii = 0;                                   // q -> (ii, ll), ll <= ii, integer only (cnt <= 28)
                while ((ii + 1) * (ii + 2) / 2 <= q) ii++;
                const int ll = q - ii * (ii + 1) / 2;
                diag_trailing_block(P, j + 1 + ii, j + 1 + ll, jb, fr, fc);
            }
        }
        __syncthreads();
        if (*s_bad) return;
    }
    // inverse by 8x8 block distance: W_ik = -X_ii * sum_{j=k}^{i-1} L_ij W_jk
    for (int d = 1; d < 8; d++) {
        for (int kb = warp; kb + d < 8; kb += 4) {
            const int i = kb + d;
            double c1[2] = {0.0, 0.0};
            for (int jj = kb; jj < i; jj++) {
#pragma unroll
                for (int ks = 0; ks < 2; ks++)         // B[k][n] = W(jj*8+k, kb*8+n): lower triangular when jj == kb
                    dmma(c1, P[(i * 8 + fr) * TS + jj * 8 + ks * 4 + fc], Wval(P, jj * 8 + ks * 4 + fc, kb * 8 + fr));
            }
            sc[fr * 8 + fc * 2] = c1[0];
            sc[fr * 8 + fc * 2 + 1] = c1[1];
            __syncwarp();
            double c2[2] = {0.0, 0.0};
#pragma unroll
            for (int ks = 0; ks < 2; ks++)             // A[m][k] = X_ii[m][k] = W(i*8+m, i*8+k), zero for k > m
                dmma(c2, Wval(P, i * 8 + fr, i * 8 + ks * 4 + fc), sc[(ks * 4 + fc) * 8 + fr]);
            Wat(P, i * 8 + fr, kb * 8 + fc * 2) = -c2[0];
            Wat(P, i * 8 + fr, kb * 8 + fc * 2 + 1) = -c2[1];
            __syncwarp();
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// kernel (3a): left-looking block-column update  A_ik -= sum_{j<k} L_ij L_kj^T  (i >= k);
// the CTA of the diagonal tile then factors it in shared memory (dpotrf semantics: a pivot
// <= 0 raises the cell's fail flag), inverts the 64x64 factor and stores
//   Dinv[k] = L_kk^-1 (row-major)   and   M(k,k) = U_kk = L_kk^-T (upper, zeros below).
// smem: PIPE_BYTES (the packed diagonal block needs 37 KB of it).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_chol_update(const OiSlot& s, int i, int k, double* smem) {
    const long long ld = s.npad;
    double acc[4][4][2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    double* Cg = s.M + (long long)i * NB * ld + (long long)k * NB;
    // The accumulators start at -A_ik (the DMMA fragment layout is the tile's own row/col layout), so the K loop leaves
    // -(A_ik - sum L_ij L_kj^T) in them: no second copy of the tile in registers (it cost 64 registers and spills at
    // 4 CTAs/SM), and the loads overlap the pipeline prologue.
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const double2 c = __ldcg((const double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)]);
            acc[mb][nb][0] = -c.x; acc[mb][nb][1] = -c.y;
        }
    {
        // rows of block i / cols of block k beyond the cell's size are padding; of the diagonal tile only
        // the lower triangle is needed (the warp above the diagonal idles)
        SubRange sr{0, hi_lt(wm, s.n16 - i * NB), 0, hi_lt(wn, s.n16 - k * NB)};
        if (i == k && wm < wn) sr.mhi = 0;
        gemm_nt_stream(acc, s.M + (long long)i * NB * ld, ld, s.M + (long long)k * NB * ld, ld, 0, k * NB, smem,
                       [sr](int) { return sr; });
    }
    if (i != k) {
#pragma unroll
        for (int mb = 0; mb < 4; mb++)
#pragma unroll
            for (int nb = 0; nb < 4; nb++) {
                double2 v;
                v.x = -acc[mb][nb][0]; v.y = -acc[mb][nb][1];
                *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * ld + FRAG_COL(wn, nb, lane)] = v;
            }
        return;
    }
    // ---- diagonal tile: P = A_kk - sum = -acc, factor + invert in shared memory (packed L / W layout, see above) ----
    double* P = smem;                          // [64][TS]
    double* sc = smem + NB * TS + warp * 64;   // per-warp 8x8 scratch
    int* s_bad = (int*)(smem + NB * TS + 4 * 64);
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            int r = FRAG_ROW(wm, mb, lane), c = FRAG_COL(wn, nb, lane);
            P[r * TS + c] = -acc[mb][nb][0];
            P[r * TS + c + 1] = -acc[mb][nb][1];
        }
    if (tid == 0) *s_bad = 0;
    __syncthreads();
    diag_factor_invert(P, sc, s_bad, tid);
    if (*s_bad) {
        if (tid == 0) *s.fail = 1;
        return;
    }
    if (warp == 0) {
        // log-determinant part: sum_i log L_ii of this block (GPR_CS2S3.py:128), fixed order
        double v = log(P[lane * TS + lane]) + log(P[(lane + 32) * TS + lane + 32]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) s.part[k] = v;
    }
    double* Dk = s.Dinv + (long long)k * OI_TILE;
    for (int idx = tid; idx < OI_TILE; idx += GEMM_THREADS) {
        int r = idx >> 6, c = idx & 63;
        Dk[idx] = Wval(P, r, c);                                   // Dinv[k] = L_kk^-1, row-major, zeros above the diagonal
        Cg[(long long)r * ld + c] = Wval(P, c, r);                 // U_kk = L_kk^-T: upper triangular, zeros below
    }
}

// kernel (3b): panel  L_ik = A_ik * L_kk^-T  (i > k), as an NT tile against Dinv[k].  smem: PIPE_BYTES.
__device__ __forceinline__ void tile_chol_panel(const OiSlot& s, int i, int k, double* smem) {
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
// Runs with 128 or 256 threads: with 128, thread tid also plays virtual thread tid+128 (8 virtual warps),
// so every sum has the same order either way.  smem: 2*NB + 24 doubles.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void cell_fwd(const OiSlot& s, const OiCellArrays& ca, const OiPacked& pk, double t_pred, bool pred,
                                         double* smem) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nthr = blockDim.x, nwarp = nthr >> 5, passes = 8 / nwarp;     // 128 threads: 2 passes, 256 threads: 1
    const int nrhs = pred ? 2 : 1;
    const long long ld = s.npad;
    double* tv = s.vec;              // t
    double* vv = s.vec + s.npad;     // v
    double(*sb)[NB] = (double(*)[NB])smem;             // [2][NB]
    double(*red)[8] = (double(*)[8])(smem + 2 * NB);   // [3][8]
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    const double h0 = __ldcg(h + 0), h1 = __ldcg(h + 1), h2 = __ldcg(h + 2), h3 = __ldcg(h + 3);
    __syncthreads();
    // right-hand sides
    for (int g = tid; g < s.npad; g += nthr) {
        double r = 0.0, ks = 0.0;
        if (g < s.n) {
            r = pk.r[s.pt_off + g];
            if (pred) {
                // cdist(sqrt(3)*x/ell, sqrt(3)*xs/ell) (GPR_CS2S3.py:100-101)
                double dx = (ROOT3 * pk.x[s.pt_off + g]) / h0 - (ROOT3 * ca.X[2 * (size_t)s.cell]) / h0;
                double dy = (ROOT3 * pk.y[s.pt_off + g]) / h1 - (ROOT3 * ca.X[2 * (size_t)s.cell + 1]) / h1;
                double dt = (ROOT3 * pk.t[s.pt_off + g]) / h2 - (ROOT3 * t_pred) / h2;
                double Q = pair_Q(dx, dy, dt);
                ks = h3 * ((1.0 + Q) * exp(-Q));
            }
        }
        tv[g] = r; vv[g] = ks;
    }
    __syncthreads();
    for (int k = 0; k < s.N; k++) {
        const int kc = k * NB;
        // s[r] = sum_{c<kc} Ls[kc+r][c] * x[c]; virtual warp vw owns rows vw*8 .. vw*8+7
#pragma unroll 1
        for (int half = 0; half < passes; half++) {
            const int vw = warp + nwarp * half;
            double a0[8], a1[8];
#pragma unroll
            for (int q = 0; q < 8; q++) { a0[q] = 0.0; a1[q] = 0.0; }
            const double* Lrow = s.M + (long long)(kc + vw * 8) * ld;
            // unrolled so that several iterations' loads (8 rows + x per iteration) are in flight; the FMAs keep their order
#pragma unroll 4
            for (int c = lane; c < kc; c += 32) {
                double x0 = __ldcg(&tv[c]), x1 = pred ? __ldcg(&vv[c]) : 0.0;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    double l = __ldcg(&Lrow[(long long)q * ld + c]);
                    a0[q] = fma(l, x0, a0[q]); a1[q] = fma(l, x1, a1[q]);
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
                for (int q = 0; q < 8; q++) { sb[0][vw * 8 + q] = a0[q]; sb[1][vw * 8 + q] = a1[q]; }
            }
        }
        // x_k = Dinv[k] * b_k - sum_{c<kc} Ls[kc+r][c] x[c]   (lower-triangular 64x64 mat-vec), thread (rhs, row)
        double dsum = 0.0;
        if (tid < NB * nrhs) {
            int rh = tid / NB, r = tid % NB;
            const double* D = s.Dinv + (long long)k * OI_TILE + r * NB;
            const double* bsrc = (rh ? vv : tv) + kc;
            for (int c = 0; c <= r; c++) dsum = fma(__ldcg(&D[c]), __ldcg(&bsrc[c]), dsum);
        }
        __syncthreads();
        if (tid < NB * nrhs) {
            int rh = tid / NB, r = tid % NB;
            (rh ? vv : tv)[kc + r] = dsum - sb[rh][r];
        }
        __syncthreads();
    }
    // scalars, fixed summation order (256 virtual threads, 8 virtual warps)
#pragma unroll 1
    for (int half = 0; half < passes; half++) {
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        for (int g = tid + nthr * half; g < s.npad; g += 256) {
            double a = __ldcg(&tv[g]), b = pred ? __ldcg(&vv[g]) : 0.0;
            q0 = fma(a, a, q0); q1 = fma(a, b, q1); q2 = fma(b, b, q2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            q0 += __shfl_down_sync(0xffffffffu, q0, o);
            q1 += __shfl_down_sync(0xffffffffu, q1, o);
            q2 += __shfl_down_sync(0xffffffffu, q2, o);
        }
        if (lane == 0) { red[0][warp + nwarp * half] = q0; red[1][warp + nwarp * half] = q1; red[2][warp + nwarp * half] = q2; }
    }
    __syncthreads();
    if (tid < 3) {
        double a = 0.0;
        for (int w = 0; w < 8; w++) a += red[tid][w];
        s.part[s.N + tid] = a;
    }
}

// ------------------------------------------------------------------------------------------
// kernel (3c'): row scaling  Ls_ij = L_ii^-1 * L_ij  (i > j), in place.
// With it the forward substitution and the inverse need no per-step triangular solve:
//   t_i = L_ii^-1 r_i - sum_{j<i} Ls_ij t_j            W_ik = -sum_{j=k}^{i-1} Ls_ij W_jk
// An "NN" product out[m][n] = sum_kk Dinv_i[m][kk] L_ij[kk][n] (the second operand is not K-contiguous), streamed through
// the same STAGES-deep cp.async pipeline as the NT tiles (smem: PIPE_BYTES, so the panel kernel keeps 4 CTAs/SM):
//   A chunk = Dinv_i[:, c:c+16]   64 rows x 16, the usual XOR-swizzled layout
//   B chunk = L_ij[c:c+16, :]     16 rows (kk) x 64 cols (n); element (kk, n) is stored at kk*64 + (n ^ ((kk & 3) << 2)):
//                                 the DMMA B fragment (4 consecutive kk x 8 consecutive n per instruction) then touches 16
//                                 distinct double-banks per half-warp, and 16-byte cp.async pieces stay whole.
// Dinv_i is lower triangular: output row m only needs kk <= m; rows of L_ij beyond the cell's size are zero.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_stage_nn(double* st, const double* __restrict__ A, const double* __restrict__ B, long long ldb,
                                              int kk0, int tid) {
    double* As = st;
    double* Bs = st + NB * KT;
#pragma unroll
    for (int it = 0; it < 4; it++) {
        const int c = tid + it * GEMM_THREADS;      // 0..511
        {
            const int row = c >> 3, p = c & 7;      // A: 64 rows x 8 pieces
            cp_async16(&As[row * KT + ((p ^ (row & 7)) << 1)], &A[row * NB + kk0 + p * 2]);
        }
        {
            const int kk = c >> 5, p = c & 31;      // B: 16 rows x 32 pieces
            cp_async16(&Bs[kk * NB + ((p * 2) ^ ((kk & 3) << 2))], &B[(long long)(kk0 + kk) * ldb + p * 2]);
        }
    }
}

__device__ __forceinline__ void tile_scale(const OiSlot& s, int i, int j, double* smem) {
    const long long ld = s.npad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const int fr = lane >> 2, fc = lane & 3;
    const double* Di = s.Dinv + (long long)i * OI_TILE;
    double* Lg = s.M + (long long)i * NB * ld + (long long)j * NB;
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int vi = min(NB, s.n16 - i * NB);     // valid rows of block i (a multiple of 16)
    const int nk = vi / KT;
    const int mhi = hi_lt(wm, vi);
    __syncthreads();
#pragma unroll
    for (int q = 0; q < STAGES - 1; q++) {
        if (q < nk) load_stage_nn(smem + q * STAGE_DOUBLES, Di, Lg, ld, q * KT, tid);
        cp_async_commit();
    }
    for (int it = 0; it < nk; it++) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        const int nx = it + STAGES - 1;
        if (nx < nk) load_stage_nn(smem + (nx % STAGES) * STAGE_DOUBLES, Di, Lg, ld, nx * KT, tid);
        cp_async_commit();
        const double* As = smem + (it % STAGES) * STAGE_DOUBLES;
        const double* Bs = As + NB * KT;
        const int mlo = lo_ge(wm, it * KT);
#pragma unroll
        for (int ks = 0; ks < KT / 4; ks++) {
            double a[4], b[4];
            const int sw = (((ks * 2 + (fc >> 1)) ^ fr) << 1) + (fc & 1);
#pragma unroll
            for (int mb = 0; mb < 4; mb++) a[mb] = As[(wm * 32 + mb * 8 + fr) * KT + sw];
#pragma unroll
            for (int nb = 0; nb < 4; nb++) b[nb] = Bs[(ks * 4 + fc) * NB + ((wn * 32 + nb * 8 + fr) ^ (fc << 2))];
#pragma unroll
            for (int mb = 0; mb < 4; mb++)
                if (mb >= mlo && mb < mhi) {
#pragma unroll
                    for (int nb = 0; nb < 4; nb++) dmma(acc[mb][nb], a[mb], b[nb]);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
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
// smem: PIPE_BYTES.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_trtri(const OiSlot& s, int kb, int d, double* smem) {
    const int i = kb + d;
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

// kernel (3e): alpha = K^-1 (y-m) = U t   (rows of U dotted with t), 64 rows per call
__device__ __forceinline__ void rows_alpha(const OiSlot& s, int rb) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ld = s.npad;
    const double* tv = s.vec;
    double* al = s.vec + 2 * (long long)s.npad;
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        const int r0 = rb * NB + (warp + 4 * half) * 8;
        double a[8];
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = 0.0;
        const double* Urow = s.M + (long long)r0 * ld;
#pragma unroll 4
        for (int c = rb * NB + lane; c < s.npad; c += 32) {
            double x = __ldcg(&tv[c]);
#pragma unroll
            for (int q = 0; q < 8; q++) a[q] = fma(__ldcg(&Urow[(long long)q * ld + c]), x, a[q]);
        }
#pragma unroll
        for (int q = 0; q < 8; q++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_down_sync(0xffffffffu, a[q], o);
            if (lane == 0) al[r0 + q] = a[q];
        }
    }
}

// ------------------------------------------------------------------------------------------
// kernel (4): K^-1 tile (i,j) = sum_{m >= i} U_im U_jm^T on DMMA, fused with the trace terms of
// GPR_CS2S3.py:130-138:  Qm = K^-1 - alpha alpha^T,
//   S_theta = sum Qm * q_theta^2 exp(-Q)   (theta = x, y, t)      S_3 = sum Qm * (1+Q) exp(-Q)
//   S_4 = tr(Qm)
// dK/dtheta is rebuilt in registers from the coordinates and the stored Q, exp(-Q) tiles; K^-1 is never stored.
// Each tile writes five partial sums; off-diagonal tiles count twice (symmetry).  smem: PIPE_BYTES.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_lauum_trace(const OiSlot& s, const OiCellArrays& ca, const OiPacked& pk, int i, int j, double* smem) {
    const long long ld = s.npad;
    double acc[4][4][2];
    ACC_ZERO(acc);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const double* h = ca.hyp + 5 * (size_t)s.cell;
    const double h0 = __ldcg(h + 0), h1 = __ldcg(h + 1), h2 = __ldcg(h + 2);
    double praw[4] = {0, 0, 0, 0};              // this thread's point (row or col of the tile): x, y, t, alpha
    {
        int g = ((tid / NB) ? j : i) * NB + tid % NB;
        if (g < s.n) {
            praw[0] = pk.x[s.pt_off + g]; praw[1] = pk.y[s.pt_off + g]; praw[2] = pk.t[s.pt_off + g];
            praw[3] = __ldcg(&s.vec[2 * (long long)s.npad + g]);
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
    // per-point data of the 64 rows and 64 cols: scaled coordinates for q_theta (3) and alpha (raw values were fetched
    // before the K loop); Q and exp(-Q) of every pair come from the tile the covariance build left in s.QE
    double(*P)[4][NB] = (double(*)[4][NB])smem;    // P[0]=rows, P[1]=cols
    double(*red)[4] = (double(*)[4])(smem + 2 * 4 * NB);   // [5][4]
    {
        int which = tid / NB, q = tid % NB;        // 128 threads: rows then cols
        double x = praw[0], y = praw[1], t = praw[2], a = praw[3];
        // np.sqrt(3.)*(x[:,theta]/ell[theta])  (GPR_CS2S3.py:97): divide, then multiply
        P[which][0][q] = ROOT3 * (x / h0); P[which][1][q] = ROOT3 * (y / h1); P[which][2][q] = ROOT3 * (t / h2);
        P[which][3][q] = a;
    }
    __syncthreads();
    const double* qe = s.QE + 2 * (long long)OI_TILE * (i * (i + 1) / 2 + j);
    double S[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int mb = 0; mb < 4; mb++) {
        const int r = FRAG_ROW(wm, mb, lane), gi = i * NB + r;
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            const int c0 = FRAG_COL(wn, nb, lane);
            if (gi >= s.n || j * NB + c0 > gi) continue;     // padding row, or both columns above the diagonal
            const double2 Q2 = __ldcg((const double2*)&qe[r * NB + c0]);
            const double2 E2 = __ldcg((const double2*)&qe[OI_TILE + r * NB + c0]);
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int c = c0 + e, gj = j * NB + c;
                if (gj < s.n && gj <= gi) {
                    double Qm = fma(-P[0][3][r], P[1][3][c], acc[mb][nb][e]);
                    if (gi == gj) {
                        // Q = 0: dK_theta = 0, K = sf2
                        S[3] += Qm; S[4] += Qm;
                    } else {
                        Qm *= 2.0;          // (gi, gj) and (gj, gi): K^-1, alpha alpha^T and dK are symmetric
                        const double Q = e ? Q2.y : Q2.x, E = e ? E2.y : E2.x;
                        double qx = P[0][0][r] - P[1][0][c], qy = P[0][1][r] - P[1][1][c], qt = P[0][2][r] - P[1][2][c];
                        S[0] = fma(Qm, qx * qx * E, S[0]);
                        S[1] = fma(Qm, qy * qy * E, S[1]);
                        S[2] = fma(Qm, qt * qt * E, S[2]);
                        S[3] = fma(Qm, (1.0 + Q) * E, S[3]);
                    }
                }
            }
        }
    }
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
