// Finalisation of one evaluation + the per-cell optimiser step, executed by ONE WARP per cell (the
// translation unit is compiled with -fmad=false so the optimiser's arithmetic rounds like the
// NumPy/Python code it restates, cg_scipy.h).
//
//   FIT / EVAL : nlZ and gradient as SMLII returns them (GPR_CS2S3.py:128-140), then one resume of the
//                scipy-CG state machine (GPR_CS2S3.py:166) -> next trial hyperparameters or "converged"
//   PREDICT    : fs, sfs2, lZ and the hyperparameters as GPR3D returns them (GPR_CS2S3.py:179-191)
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "oi_types.h"
#include "oi_launch.h"
#include "cg_scipy.h"
#include "lbfgs_fast.h"

#define LOG_2PI 1.8378770664093453   // np.log(2*np.pi)

__global__ void k_cg_init(OiCellArrays ca, int n_cells, OiRunConst rc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    double dummy[OI_MAXH] = {0, 0, 0, 0, 0, 0};
    if (rc.optimiser == 1) {
        OiLbfgsState* S = (OiLbfgsState*)((char*)ca.cg + (size_t)c * OI_OPT_STATE_BYTES);
        oi_lbfgs_init(*S, rc.x0, rc.n_hyp, rc.maxiter, rc.gtol);
        oi_lbfgs_resume(*S, 0.0, dummy);
        for (int q = 0; q < 5; q++) ca.hyp[5 * (size_t)c + q] = exp(S->req_x[q]);
        return;
    }
    OiCgState* S = (OiCgState*)((char*)ca.cg + (size_t)c * OI_OPT_STATE_BYTES);
    OiCgState L;
    oi_cg_init(L, rc.x0, rc.n_hyp, rc.maxiter, rc.gtol);
    oi_cg_resume(L, 0.0, dummy);          // first resume only publishes x0 as the first request
    *S = L;
    for (int q = 0; q < 5; q++) ca.hyp[5 * (size_t)c + q] = exp(L.req_x[q]);
}

// Returns the cell's new phase (valid in lane 0; broadcast by the caller if needed).  All loads of data
// produced by other CTAs go through L2 (__ldcg), see oi_tiles.cuh.
__device__ __noinline__ int warp_finalize(const OiSlot s, const OiCellArrays ca, const OiRunConst rc, int phase, int lane) {
    const bool failed = *(volatile int*)s.fail != 0;
    double S[5] = {0, 0, 0, 0, 0};
    if (!failed && phase != OI_PH_PREDICT) {
        const int nt = s.N * (s.N + 1) / 2;
        const double* tp = s.part + s.N + 8;
        for (int t = lane; t < nt; t += 32)
            for (int q = 0; q < 5; q++) S[q] += __ldcg(&tp[5 * (size_t)t + q]);
        for (int q = 0; q < 5; q++)
            for (int o = 16; o > 0; o >>= 1) S[q] += __shfl_down_sync(0xffffffffu, S[q], o);
    }
    if (lane != 0) return OI_PH_DONE;
    double logdet = 0.0;
    if (!failed) for (int k = 0; k < s.N; k++) logdet += __ldcg(&s.part[k]);
    const double quad = __ldcg(&s.part[s.N + 0]), vt = __ldcg(&s.part[s.N + 1]), vv = __ldcg(&s.part[s.N + 2]);
    double* h = ca.hyp + 5 * (size_t)s.cell;
    const double sf2 = __ldcg(&h[3]), sn2 = __ldcg(&h[4]);
    int newphase = OI_PH_DONE;
    if (phase == OI_PH_PREDICT) {
        double* o = ca.out + 8 * (size_t)s.cell;
        if (failed) {
            for (int q = 0; q < 8; q++) o[q] = nan("");
            ca.status[s.cell] = 3;
        } else {
            o[0] = rc.mean + vt;                                   // fs  = mean + k*^T A            (:181)
            o[1] = sqrt(sf2 - vv);                                 // sfs2 = sqrt(k** - v^T v)       (:182)
            o[2] = -quad / 2 - logdet - s.n * LOG_2PI / 2;         // lZ                             (:179)
            for (int q = 0; q < 5; q++) o[3 + q] = __ldcg(&h[q]);
        }
    } else {
        double f, g[OI_MAXH];
        if (failed) {
            f = INFINITY;
            for (int q = 0; q < OI_MAXH; q++) g[q] = INFINITY;      // :139-140
        } else {
            f = quad / 2 + logdet + s.n * LOG_2PI / 2;              // :128
            g[0] = (sf2 * S[0]) / 2; g[1] = (sf2 * S[1]) / 2; g[2] = (sf2 * S[2]) / 2;   // :134
            g[3] = sf2 * S[3];                                      // (Q*(2*Kx)).sum()/2   :136
            g[4] = sn2 * S[4];                                      // sn2*trace(Q)         :138
            if (rc.grad_convention == 1) { g[3] /= 2; g[4] /= 2; }  // true derivatives
            g[5] = 0.0;
        }
        if (ca.dbg && s.cell == ca.dbg_cell && phase == OI_PH_FIT) {
            const int at = atomicAdd(ca.dbg_count, 1);
            if (at < ca.dbg_cap) {
                double* d = ca.dbg + 12 * (size_t)at;
                for (int q = 0; q < 5; q++) d[q] = __ldcg(&h[q]);
                d[5] = f;
                for (int q = 0; q < OI_MAXH; q++) d[6 + q] = g[q];
            }
        }
        if (phase == OI_PH_EVAL) {
            ca.evf[s.cell] = f;
            for (int q = 0; q < OI_MAXH; q++) ca.evg[OI_MAXH * (size_t)s.cell + q] = g[q];
        } else if (rc.optimiser == 1) {
            // fast mode: exact-gradient L-BFGS (lbfgs_fast.h); the state stays in global memory (1.2 KB per cell)
            OiLbfgsState* Sg = (OiLbfgsState*)((char*)ca.cg + (size_t)s.cell * OI_OPT_STATE_BYTES);
            int r = oi_lbfgs_resume(*Sg, f, g);
            if (r == OI_CG_NEED_EVAL) {
                for (int q = 0; q < 5; q++) h[q] = exp(Sg->req_x[q]);
                newphase = OI_PH_FIT;
            } else {
                for (int q = 0; q < 5; q++) h[q] = exp(Sg->xk[q]);
                ca.nfev[s.cell] = Sg->nfev;
                ca.status[s.cell] = Sg->status == 3 ? 5 : Sg->status;
                newphase = OI_PH_PREDICT;
            }
        } else {
            OiCgState* Sg = (OiCgState*)((char*)ca.cg + (size_t)s.cell * OI_OPT_STATE_BYTES);
            OiCgState L = *Sg;
            int r = oi_cg_resume(L, f, g);
            *Sg = L;
            if (r == OI_CG_NEED_EVAL) {
                for (int q = 0; q < 5; q++) h[q] = exp(L.req_x[q]);
                newphase = OI_PH_FIT;
            } else {
                for (int q = 0; q < 5; q++) h[q] = exp(L.xk[q]);    // np.exp(res.x), status ignored (:166)
                ca.nfev[s.cell] = L.nfev;
                ca.status[s.cell] = L.status == 3 ? 5 : L.status;
                newphase = OI_PH_PREDICT;
            }
        }
    }
    ca.phase[s.cell] = newphase;
    return newphase;
}
