// Fast-mode optimiser (NOT the parity mode): resumable limited-memory BFGS on the LOG hyperparameters with the TRUE
// gradient (OI_GRAD_EXACT), one objective+gradient evaluation per resume, so that a lockstep batch of cells can each
// run their own optimiser on the device next to the reference's CG restatement (cg_scipy.h).
//
// BASELINE.json's north_star (5) asks for "a lockstep batched L-BFGS-B on device with per-cell convergence masks and the
// reference's bounds and parameterisation".  The reference (GPR_CS2S3.py:166) fits log-hyperparameters without bounds, so
// the box is (-inf, +inf) and L-BFGS-B reduces to L-BFGS:
//   * direction: two-loop recursion over the last OI_LBFGS_M (s, y) pairs, H0 = (s.y / y.y) I          (Nocedal 1980)
//   * line search: More'-Thuente DCSRCH with L-BFGS-B's constants ftol = 1e-3, gtol = 0.9, xtol = 0.1, first step
//     min(1/|d|, stpmax) when there is no history, 1 otherwise, at most 20 evaluations per search (lnsrlb);
//     the dcstep function is the restatement in cg_scipy.h
//   * stops like scipy's L-BFGS-B defaults: max|g| <= pgtol (1e-5) or (f_k - f_{k+1}) <= factr*eps * max(|f_k|,|f_{k+1}|,1)
//     (2.2e-9), or maxiter iterations / 4*maxiter evaluations
//   * an evaluation that returns inf/NaN (Cholesky failure, GPR_CS2S3.py:139-140) shortens the step to a quarter of the
//     distance to the best point of the search and brackets it from above.
// SURVEY.md Appendix C.5 measured 28-72 evaluations per cell with scipy's L-BFGS-B against the CG's 60-156.
#pragma once
#include "cg_scipy.h"

#define OI_LBFGS_M 8

struct OiLbfgsState {
    int pc, dim, status, k, nfev, maxiter;      // status: 0 converged, 1 iteration/evaluation limit, 2 line search failed, 3 NaN at x0
    double pgtol, ftol;
    double xk[OI_MAXH], gk[OI_MAXH], d[OI_MAXH], fk;
    double req_x[OI_MAXH];
    // history ring: slot (hist_head - 1 - j) mod M holds the j-th newest pair
    int hist_n, hist_head;
    double S[OI_LBFGS_M][OI_MAXH], Y[OI_LBFGS_M][OI_MAXH], rho[OI_LBFGS_M], gamma;
    // line search
    int ls_evals, task, brackt, stage;
    double stp, ginit, gtest, gx, gy, finit, fx, fy, stx, sty, stmin, stmax, width, width1, stpmax;
    double fbest, stbest, gbest[OI_MAXH];       // lowest finite value seen in this search
    double fnew, gnew[OI_MAXH];
};

// per-cell optimiser state slot in device memory: big enough for either optimiser
#define OI_OPT_STATE_BYTES ((sizeof(OiCgState) > sizeof(OiLbfgsState) ? sizeof(OiCgState) : sizeof(OiLbfgsState)))

namespace oilbfgs {
using oicg::T_START; using oicg::T_FG; using oicg::T_CONV; using oicg::T_WARN; using oicg::T_ERROR;
const double FTOL = 1e-3, GTOL = 0.9, XTOL = 0.1, STPMIN = 0.0, BIG = 1e10;

// _dcsrch.py DCSRCH._iterate (the same algorithm as oicg::dcsrch_iterate, other constants, state in OiLbfgsState)
OI_HD void dcsrch(OiLbfgsState& S, double& stp, double f, double g, int& task) {
    using namespace oicg;
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (task == T_START) {
        if (stp < STPMIN || stp > S.stpmax || g >= 0) { task = T_ERROR; return; }
        S.brackt = 0; S.stage = 1; S.finit = f; S.ginit = g; S.gtest = FTOL * S.ginit;
        S.width = S.stpmax - STPMIN; S.width1 = S.width / p5;
        S.stx = 0.0; S.fx = S.finit; S.gx = S.ginit;
        S.sty = 0.0; S.fy = S.finit; S.gy = S.ginit;
        S.stmin = 0; S.stmax = stp + xtrapu * stp;
        task = T_FG;
        return;
    }
    double ftest = S.finit + stp * S.gtest;
    if (S.stage == 1 && f <= ftest && g >= 0) S.stage = 2;
    if (S.brackt && (stp <= S.stmin || stp >= S.stmax)) task = T_WARN;
    if (S.brackt && S.stmax - S.stmin <= XTOL * S.stmax) task = T_WARN;
    if (stp == S.stpmax && f <= ftest && g <= S.gtest) task = T_WARN;
    if (stp == STPMIN && (f > ftest || g >= S.gtest)) task = T_WARN;
    if (f <= ftest && fabs(g) <= GTOL * -S.ginit) task = T_CONV;
    if (task == T_WARN || task == T_CONV) return;
    if (S.stage == 1 && f <= S.fx && f > ftest) {
        double fm = f - stp * S.gtest, fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
        double gm = g - S.gtest, gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
        S.fx = fxm + S.stx * S.gtest; S.fy = fym + S.sty * S.gtest;
        S.gx = gxm + S.gtest; S.gy = gym + S.gtest;
    } else {
        dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, g, S.brackt, S.stmin, S.stmax);
    }
    if (S.brackt) {
        if (fabs(S.sty - S.stx) >= p66 * S.width1) stp = S.stx + p5 * (S.sty - S.stx);
        S.width1 = S.width;
        S.width = fabs(S.sty - S.stx);
    }
    if (S.brackt) { S.stmin = pymin(S.stx, S.sty); S.stmax = pymax(S.stx, S.sty); }
    else { S.stmin = stp + xtrapl * (stp - S.stx); S.stmax = stp + xtrapu * (stp - S.stx); }
    stp = npclip(stp, STPMIN, S.stpmax);
    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) ||
        (S.brackt && S.stmax - S.stmin <= XTOL * S.stmax)) stp = S.stx;
    task = T_FG;
}

OI_HD bool all_finite(double f, const double* g, int n) {
    bool ok = oicg::finite_(f);
    for (int i = 0; i < n; i++) ok = ok && oicg::finite_(g[i]);
    return ok;
}

// d = -H g by the two-loop recursion
OI_HD void direction(OiLbfgsState& S) {
    const int n = S.dim;
    double q[OI_MAXH], a[OI_LBFGS_M];
    for (int i = 0; i < n; i++) q[i] = S.gk[i];
    for (int j = 0; j < S.hist_n; j++) {
        const int s = (S.hist_head - 1 - j + 2 * OI_LBFGS_M) % OI_LBFGS_M;
        a[j] = S.rho[s] * oicg::dot(S.S[s], q, n);
        for (int i = 0; i < n; i++) q[i] -= a[j] * S.Y[s][i];
    }
    const double g0 = S.hist_n > 0 ? S.gamma : 1.0;
    for (int i = 0; i < n; i++) q[i] *= g0;
    for (int j = S.hist_n - 1; j >= 0; j--) {
        const int s = (S.hist_head - 1 - j + 2 * OI_LBFGS_M) % OI_LBFGS_M;
        const double b = S.rho[s] * oicg::dot(S.Y[s], q, n);
        for (int i = 0; i < n; i++) q[i] += S.S[s][i] * (a[j] - b);
    }
    for (int i = 0; i < n; i++) S.d[i] = -q[i];
}
}  // namespace oilbfgs

OI_HD void oi_lbfgs_init(OiLbfgsState& S, const double* x0, int dim, int maxiter, double pgtol) {
    S.pc = 0; S.dim = dim; S.status = -1; S.k = 0; S.nfev = 0;
    S.maxiter = maxiter > 0 ? maxiter : 500;
    S.pgtol = pgtol; S.ftol = 1e7 * 2.220446049250313e-16;       // scipy: factr = ftol / eps with ftol = 2.22e-9
    for (int i = 0; i < OI_MAXH; i++) { S.xk[i] = i < dim ? x0[i] : 0.0; S.gk[i] = 0; S.d[i] = 0; S.req_x[i] = S.xk[i]; }
    S.hist_n = 0; S.hist_head = 0; S.gamma = 1.0; S.fk = 0;
}

// First call: pc == 0, (f_in, g_in) ignored; returns NEED_EVAL with S.req_x = x0.  Every later call passes the
// objective and its true gradient at S.req_x.
OI_HD int oi_lbfgs_resume(OiLbfgsState& S, double f_in, const double* g_in) {
    using namespace oilbfgs;
    const int n = S.dim;
    switch (S.pc) {
    case 0:
        for (int i = 0; i < n; i++) S.req_x[i] = S.xk[i];
        S.pc = 1; return OI_CG_NEED_EVAL;
    case 1:
        S.nfev++;
        S.fk = f_in;
        for (int i = 0; i < n; i++) S.gk[i] = g_in[i];
        if (!all_finite(f_in, g_in, n)) { S.status = 3; S.pc = -1; return OI_CG_DONE; }
        S.status = 0;
        while (true) {
            if (oicg::amax_abs(S.gk, n) <= S.pgtol) { S.status = 0; break; }
            if (S.k >= S.maxiter || S.nfev >= 4 * S.maxiter) { S.status = 1; break; }
            direction(S);
            if (!(oicg::dot(S.gk, S.d, n) < 0)) {                 // not a descent direction: drop the history
                S.hist_n = 0;
                for (int i = 0; i < n; i++) S.d[i] = -S.gk[i];
            }
            // ---------------- line search (lnsrlb + dcsrch) ----------------
            S.stpmax = BIG;
            S.stp = S.hist_n == 0 ? oicg::pymin(1.0 / sqrt(oicg::dot(S.d, S.d, n)), S.stpmax) : 1.0;
            S.ls_evals = 0; S.task = T_START;
            S.fbest = S.fk; S.stbest = 0.0;
            dcsrch(S, S.stp, S.fk, oicg::dot(S.gk, S.d, n), S.task);
            while (S.task == T_FG) {
                if (S.ls_evals >= 20) { S.task = T_WARN; break; }
                for (int i = 0; i < n; i++) S.req_x[i] = S.xk[i] + S.stp * S.d[i];
                S.pc = 2; return OI_CG_NEED_EVAL;
    case 2:
                S.nfev++; S.ls_evals++;
                if (!all_finite(f_in, g_in, n)) {
                    // infeasible trial point: bracket from above, retreat towards the best point of the search
                    S.sty = S.stp; S.fy = fabs(S.fx) * 2 + 1e3; S.gy = fabs(S.ginit);
                    S.brackt = 1;
                    S.stmin = oicg::pymin(S.stx, S.sty); S.stmax = oicg::pymax(S.stx, S.sty);
                    S.width = fabs(S.sty - S.stx); S.width1 = 2 * S.width;
                    S.stp = S.stx + 0.25 * (S.sty - S.stx);
                    if (!(S.stp > 0) || S.stp == S.stx) { S.task = T_WARN; break; }
                    continue;
                }
                S.fnew = f_in;
                for (int i = 0; i < n; i++) S.gnew[i] = g_in[i];
                if (f_in < S.fbest) {
                    S.fbest = f_in; S.stbest = S.stp;
                    for (int i = 0; i < n; i++) S.gbest[i] = g_in[i];
                }
                {
                    double stp_eval = S.stp;
                    dcsrch(S, S.stp, f_in, oicg::dot(g_in, S.d, n), S.task);
                    if (S.task == T_CONV) S.stp = stp_eval;
                }
            }
            if (S.task != T_CONV) {
                // search failed: take the best point it saw if that is a decrease, else restart from steepest descent once
                if (S.stbest > 0 && S.fbest < S.fk) {
                    S.stp = S.stbest; S.fnew = S.fbest;
                    for (int i = 0; i < n; i++) S.gnew[i] = S.gbest[i];
                } else if (S.hist_n > 0) {
                    S.hist_n = 0;
                    continue;
                } else { S.status = 2; break; }
            }
            // ---------------- accept the step, update the history ----------------
            {
                double s[OI_MAXH], y[OI_MAXH];
                for (int i = 0; i < n; i++) { s[i] = S.stp * S.d[i]; y[i] = S.gnew[i] - S.gk[i]; }
                const double sy = oicg::dot(s, y, n), yy = oicg::dot(y, y, n);
                if (sy > 2.220446049250313e-16 * yy && yy > 0) {
                    const int slot = S.hist_head;
                    for (int i = 0; i < n; i++) { S.S[slot][i] = s[i]; S.Y[slot][i] = y[i]; }
                    S.rho[slot] = 1.0 / sy; S.gamma = sy / yy;
                    S.hist_head = (S.hist_head + 1) % OI_LBFGS_M;
                    if (S.hist_n < OI_LBFGS_M) S.hist_n++;
                }
                const double fold = S.fk;
                for (int i = 0; i < n; i++) { S.xk[i] += s[i]; S.gk[i] = S.gnew[i]; }
                S.fk = S.fnew;
                S.k++;
                if (fold - S.fk <= S.ftol * oicg::pymax3(fabs(fold), fabs(S.fk), 1.0)) {
                    S.status = 0;                                 // L-BFGS-B: CONVERGENCE: REL_REDUCTION_OF_F <= FACTR*EPSMCH
                    break;
                }
            }
        }
        S.pc = -1;
        return OI_CG_DONE;
    default:
        return OI_CG_DONE;
    }
}
