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
    void* cg;             // [n_cells] optimiser state (OiCgState or OiLbfgsState)
    double* dbg;          // optional evaluation trace of ONE cell (oi_debug_trace): [cap][12] = hyp(5) | f | g(6)
    int dbg_cell, dbg_cap;
    int* dbg_count;
};

struct OiPacked {
    const double* x; const double* y; const double* t; const double* r;   // r = z - prior mean
};
