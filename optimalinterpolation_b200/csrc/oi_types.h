// Shared device/host types of the lockstep batched GP evaluation.
#pragma once
#include <stdint.h>

#define OI_NB 64            // algorithmic block size = DMMA CTA tile edge
#define OI_TILE (OI_NB * OI_NB)

// per-cell phase inside one oi_run
enum { OI_PH_FIT = 0, OI_PH_PREDICT = 1, OI_PH_DONE = 2, OI_PH_EVAL = 3 };

// One entry per cell that takes part in the current lockstep iteration.  All scratch is only
// live within the iteration, so the host re-packs it into the arena every iteration.
struct OiSlot {
    double* M;        // npad x npad row-major: lower = K -> L (off-diagonal blocks), upper = U = L^-T
    double* Dinv;     // N x (64x64) row-major inverses of the diagonal Cholesky blocks
    double* vec;      // 3*npad: t = L^-1 r | v = L^-1 k* | alpha = K^-1 r
    double* part;     // [0,N) logdet parts | [N,N+8) scalars (t.t, v.t, v.v) | [N+8, ..) 5 trace partials per tile
    double* QE;       // per lower tile (i,j): 64x64 Q = scaled distances, then 64x64 E = exp(-Q), written by the covariance
                      // build and re-read by the trace epilogue (trades FP64-pipe work, shared with DMMA, for idle HBM bandwidth)
    int* fail;        // set when a Cholesky pivot is <= 0 or NaN (np.linalg.LinAlgError in the reference)
    int* flags;       // [2N] dependency flags of the fused Cholesky: diag_done[k] | col_done[k] (finished off-diagonal tiles)
    long long pt_off; // offset of this cell's points in the packed (CSR-ordered) coordinate arrays
    int cell, n, npad, N;
    int n16, pad_;    // n rounded up to the DMMA K chunk (16): K loops and edge sub-tiles stop here
};

struct OiCellArrays {
    const double* X;      // [n_cells][2] target coordinates
    double* hyp;          // [n_cells][5] natural-unit hyperparameters of the NEXT evaluation
    int* phase;           // [n_cells]
    double* out;          // [n_cells][8]
    int* nfev;            // [n_cells]
    int* status;          // [n_cells]
    double* evf;          // [n_cells]    (OI_PH_EVAL)
    double* evg;          // [n_cells][6] (OI_PH_EVAL)
    void* cg;             // [n_cells] OiCgState
};

struct OiPacked {
    const double* x; const double* y; const double* t; const double* r;   // r = z - prior mean
};

// ---- persistent group engine (oi_kernels.cu: k_gp_persistent) ----
struct OiWork { long long pt_off; int cell, n; };            // one unfinished cell of the work list (sorted by descending n)
struct OiGroupCtl { unsigned count; int cur[2]; int pad_[29]; };   // 128 B per group: barrier counter + current work index
struct OiPersistAcc {
    unsigned long long cycles[8];      // CTA clock cycles per phase (build, chol, scale, fwd+trtri, alpha, lauum, finalize, idle)
    double flops, flops_factor, flops_chol;
    unsigned long long n_evals, n_pred;
};
struct OiPersist {
    const OiWork* work; int n_work;
    int* queue_head;                   // next work index
    OiGroupCtl* ctl;                   // [n_groups]
    char* scratch; size_t scratch_stride;   // per-group scratch (sized for the largest cell of the work list)
    int* fail;                         // [n_groups]
    int gs;                            // CTAs per group
    int evals_cap;                     // evaluations a group spends on one cell before it takes the next
    OiPersistAcc* acc;
};

// ---- fused Cholesky (oi_kernels.cu: k_chol_fused): CTA tickets are laid out column by column,
// segment 2k = diagonal tiles (k,k) of all cells with N > k, segment 2k+1 = off-diagonal tiles of column k
#define OI_MAX_NB 128
struct OiCholPlan { int Nmax; int off[2 * OI_MAX_NB + 1]; };
