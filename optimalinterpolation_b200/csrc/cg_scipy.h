// Resumable (one evaluation per resume) restatement of the optimiser the reference calls at
// /root/reference/2021_paper_production/GPR_CS2S3.py:166:
//     scipy.optimize.minimize(SMLII, x0, args, method='CG', jac=True)
// i.e. scipy 1.18.1  optimize/_optimize.py:1719-1878 (_minimize_cg, Polak-Ribiere+),
//      _optimize.py:1156-1198 (_line_search_wolfe12), _linesearch.py:37-187 (wolfe1 -> DCSRCH),
//      _linesearch.py:192-560 (wolfe2, _zoom, _cubicmin, _quadmin), _dcsrch.py (dcsrch, dcstep).
// scipy is a third-party dependency of the reference (unpinned there); the algorithm is
// restated here from its published source with the same constants, the same branch order and
// Python's min/max/NaN comparison semantics, so that a lockstep batch of cells can each run
// their own line search on the device: the state machine asks for ONE objective+gradient
// evaluation at a time (S.req_x) and is resumed with the result.
//
// Compiled for the device (one thread per cell, oi_kernels.cu) and for the host (tests only,
// tests/cg_host.cpp) from this same header.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define OI_HD __host__ __device__ __forceinline__
#else
#define OI_HD inline
#endif

#define OI_MAXH 6

enum { OI_CG_NEED_EVAL = 0, OI_CG_DONE = 1 };
// status mirrors scipy's warnflag: 0 success, 1 maxiter, 2 precision loss (line search failed), 3 NaN
struct OiCgState {
    int pc, dim, status, k, nfev, maxiter;
    double gtol;
    double xk[OI_MAXH], pk[OI_MAXH], gfk[OI_MAXH];
    double old_fval, old_old_fval, gnorm, deltak;
    // cached_step of _minimize_cg
    int cached_valid; double cached_alpha, c_x[OI_MAXH], c_p[OI_MAXH], c_g[OI_MAXH], c_gnorm;
    // evaluation request / last evaluation in this line search
    double req_x[OI_MAXH];
    int ev_valid; double ev_alpha, ev_f, ev_g[OI_MAXH];
    // MemoizeJac (_optimize.py:61-85): the last point actually evaluated and its value/gradient
    int memo_valid; double memo_x[OI_MAXH], memo_f, memo_g[OI_MAXH];
    // line-search common
    double derphi0, phi0, alpha_k, ls_fval; int ls_ok, have_gfkp1; double gfkp1[OI_MAXH];
    // DCSRCH
    int d_task, d_i, d_brackt, d_stage;
    double d_stp, d_phi1, d_derphi1, d_alpha1;
    double d_ginit, d_gtest, d_gx, d_gy, d_finit, d_fx, d_fy, d_stx, d_sty, d_stmin, d_stmax, d_width, d_width1;
    // wolfe2
    int w_i, w_zoom, w_fail;
    double w_alpha0, w_alpha1, w_phi_a0, w_phi_a1, w_derphi_a0, w_derphi_a1;
    double w_alpha_star, w_phi_star; int w_have_star, w_have_dstar;
    // zoom
    int z_i;
    double z_alo, z_ahi, z_philo, z_phihi, z_dlo, z_arec, z_phirec, z_aj, z_phiaj, z_daj;
};

namespace oicg {
enum { T_START = 0, T_FG = 1, T_CONV = 2, T_WARN = 3, T_ERROR = 4 };
const double C1 = 1e-4, C2 = 0.4, AMIN = 1e-100, AMAX = 1e100, XTOL = 1e-14, SIGMA3 = 0.01;

OI_HD double pymax(double a, double b) { return (b > a) ? b : a; }   // Python max(a, b)
OI_HD double pymin(double a, double b) { return (b < a) ? b : a; }   // Python min(a, b)
OI_HD double pymax3(double a, double b, double c) { return pymax(pymax(a, b), c); }
OI_HD double npclip(double x, double lo, double hi) {                // np.clip (NaN propagates)
    if (x != x) return x;
    double m = (x < lo) ? lo : x;
    return (m > hi) ? hi : m;
}
OI_HD double npsign(double x) { return (x != x) ? x : ((x > 0) ? 1.0 : ((x < 0) ? -1.0 : 0.0)); }
OI_HD bool finite_(double x) { return (x - x) == 0.0; }
// np.dot on short vectors = OpenBLAS ddot scalar tail loop `dot += y[i]*x[i]`, which its x86-64
// kernels compile with FMA contraction; sequential fma reproduces it bit for bit (tests/test_cg.py).
OI_HD double dot(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; i++) s = fma(b[i], a[i], s); return s; }
OI_HD double amax_abs(const double* a, int n) {                       // np.amax(np.abs(a)), NaN propagates
    double m = fabs(a[0]);
    for (int i = 1; i < n; i++) { double v = fabs(a[i]); if (v != v || m != m) m = v + m; else if (v > m) m = v; }
    return m;
}

// _dcsrch.py dcstep
OI_HD void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy, double& dy, double& stp,
                  double fp, double dp, int& brackt, double stpmin, double stpmax) {
    double sgnd = npsign(dp) * npsign(dx);
    double stpf, stpc, stpq, theta, s, gamma, p, q, r;
    if (fp > fx) {
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp;
        s = pymax3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp < stx) gamma *= -1;
        p = (gamma - dx) + theta;
        q = ((gamma - dx) + gamma) + dp;
        r = p / q;
        stpc = stx + r * (stp - stx);
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx);
        if (fabs(stpc - stx) <= fabs(stpq - stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = pymax3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s));
        if (stp > stx) gamma *= -1;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + dx;
        r = p / q;
        stpc = stp + r * (stx - stp);
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
        else stpf = stpq;
        brackt = 1;
    } else if (fabs(dp) < fabs(dx)) {
        theta = 3 * (fx - fp) / (stp - stx) + dx + dp;
        s = pymax3(fabs(theta), fabs(dx), fabs(dp));
        gamma = s * sqrt(pymax(0, (theta / s) * (theta / s) - (dx / s) * (dp / s)));
        if (stp > stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (dx - dp)) + gamma;
        r = p / q;
        if (r < 0 && gamma != 0) stpc = stp + r * (stx - stp);
        else if (stp > stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = stp + (dp / (dp - dx)) * (stx - stp);
        if (brackt) {
            if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            if (stp > stx) stpf = pymin(stp + 0.66 * (sty - stp), stpf);
            else stpf = pymax(stp + 0.66 * (sty - stp), stpf);
        } else {
            if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
            else stpf = stpq;
            stpf = npclip(stpf, stpmin, stpmax);
        }
    } else {
        if (brackt) {
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp;
            s = pymax3(fabs(theta), fabs(dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s));
            if (stp > sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + dy;
            r = p / q;
            stpc = stp + r * (sty - stp);
            stpf = stpc;
        } else if (stp > stx) stpf = stpmax;
        else stpf = stpmin;
    }
    if (fp > fx) { sty = stp; fy = fp; dy = dp; }
    else {
        if (sgnd < 0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// _dcsrch.py DCSRCH._iterate with ftol=C1, gtol=C2, xtol=XTOL, stpmin=AMIN, stpmax=AMAX
OI_HD void dcsrch_iterate(OiCgState& S, double& stp, double f, double g, int& task) {
    const double p5 = 0.5, p66 = 0.66, xtrapl = 1.1, xtrapu = 4.0;
    if (task == T_START) {
        if (stp < AMIN) task = T_ERROR;
        if (stp > AMAX) task = T_ERROR;
        if (g >= 0) task = T_ERROR;
        if (task == T_ERROR) return;
        S.d_brackt = 0; S.d_stage = 1; S.d_finit = f; S.d_ginit = g; S.d_gtest = C1 * S.d_ginit;
        S.d_width = AMAX - AMIN; S.d_width1 = S.d_width / p5;
        S.d_stx = 0.0; S.d_fx = S.d_finit; S.d_gx = S.d_ginit;
        S.d_sty = 0.0; S.d_fy = S.d_finit; S.d_gy = S.d_ginit;
        S.d_stmin = 0; S.d_stmax = stp + xtrapu * stp;
        task = T_FG;
        return;
    }
    double ftest = S.d_finit + stp * S.d_gtest;
    if (S.d_stage == 1 && f <= ftest && g >= 0) S.d_stage = 2;
    if (S.d_brackt && (stp <= S.d_stmin || stp >= S.d_stmax)) task = T_WARN;
    if (S.d_brackt && S.d_stmax - S.d_stmin <= XTOL * S.d_stmax) task = T_WARN;
    if (stp == AMAX && f <= ftest && g <= S.d_gtest) task = T_WARN;
    if (stp == AMIN && (f > ftest || g >= S.d_gtest)) task = T_WARN;
    if (f <= ftest && fabs(g) <= C2 * -S.d_ginit) task = T_CONV;
    if (task == T_WARN || task == T_CONV) return;
    if (S.d_stage == 1 && f <= S.d_fx && f > ftest) {
        double fm = f - stp * S.d_gtest, fxm = S.d_fx - S.d_stx * S.d_gtest, fym = S.d_fy - S.d_sty * S.d_gtest;
        double gm = g - S.d_gtest, gxm = S.d_gx - S.d_gtest, gym = S.d_gy - S.d_gtest;
        dcstep(S.d_stx, fxm, gxm, S.d_sty, fym, gym, stp, fm, gm, S.d_brackt, S.d_stmin, S.d_stmax);
        S.d_fx = fxm + S.d_stx * S.d_gtest; S.d_fy = fym + S.d_sty * S.d_gtest;
        S.d_gx = gxm + S.d_gtest; S.d_gy = gym + S.d_gtest;
    } else {
        dcstep(S.d_stx, S.d_fx, S.d_gx, S.d_sty, S.d_fy, S.d_gy, stp, f, g, S.d_brackt, S.d_stmin, S.d_stmax);
    }
    if (S.d_brackt) {
        if (fabs(S.d_sty - S.d_stx) >= p66 * S.d_width1) stp = S.d_stx + p5 * (S.d_sty - S.d_stx);
        S.d_width1 = S.d_width;
        S.d_width = fabs(S.d_sty - S.d_stx);
    }
    if (S.d_brackt) { S.d_stmin = pymin(S.d_stx, S.d_sty); S.d_stmax = pymax(S.d_stx, S.d_sty); }
    else { S.d_stmin = stp + xtrapl * (stp - S.d_stx); S.d_stmax = stp + xtrapu * (stp - S.d_stx); }
    stp = npclip(stp, AMIN, AMAX);
    if ((S.d_brackt && (stp <= S.d_stmin || stp >= S.d_stmax)) ||
        (S.d_brackt && S.d_stmax - S.d_stmin <= XTOL * S.d_stmax)) stp = S.d_stx;
    task = T_FG;
}

// x**3 rounded once (what pow(x, 3) returns): double-double product, then one rounding
OI_HD double cube(double x) {
    double p = x * x, e = fma(x, x, -p);
    double r = p * x, e2 = fma(p, x, -r);
    return r + (e2 + e * x);
}
// _linesearch.py _cubicmin / _quadmin; return false for Python's None
OI_HD bool cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double& xmin) {
    double C = fpa, db = b - a, dc = c - a;
    double denom = (db * dc) * (db * dc) * (db - dc);
    double v0 = fb - fa - C * db, v1 = fc - fa - C * dc;
    // d1 = [[dc**2, -db**2], [-dc**3, db**3]]; `**3` and `**2` on numpy scalars are C pow(): the restatement rounds the
    // exact power once (glibc's pow differs from that in ~0.1 % of the arguments, by one ulp; not reproducible here).
    // np.dot(d1, v) (2x2 dgemv, OpenBLAS 0.3.30 Haswell/SkylakeX dgemv_t): row r = fma(d1[r][0], v0, d1[r][1]*v1) --
    // measured on 20000 random 2x2 systems with 1 and 8 BLAS threads, 100 % agreement with this form, 67 % with
    // fma(d1[r][1], v1, d1[r][0]*v0) (round 1's form, which only the rarely taken cubic branch of _zoom exposed).
    double A = fma(dc * dc, v0, (-(db * db)) * v1);
    double B = fma(-cube(dc), v0, cube(db) * v1);
    if (denom == 0.0 || denom != denom) return false;
    A /= denom; B /= denom;
    double radical = B * B - 3 * A * C;
    if (!(radical >= 0)) return false;
    if (3 * A == 0.0) return false;
    xmin = a + (-B + sqrt(radical)) / (3 * A);
    return finite_(xmin);
}
OI_HD bool quadmin(double a, double fa, double fpa, double b, double fb, double& xmin) {
    double D = fa, C = fpa, db = b - a * 1.0;
    if (db * db == 0.0) return false;
    double B = (fb - D - C * db) / (db * db);
    if (2.0 * B == 0.0) return false;
    xmin = a - C / (2.0 * B);
    return finite_(xmin);
}

// polak_ribiere_powell_step into the cached_step slots (_optimize.py:1808-1815)
OI_HD void pr_step(OiCgState& S, double alpha, const double* g1) {
    int n = S.dim;
    double yk[OI_MAXH];
    for (int i = 0; i < n; i++) { S.c_x[i] = S.xk[i] + alpha * S.pk[i]; yk[i] = g1[i] - S.gfk[i]; }
    double beta = pymax(0, dot(yk, g1, n) / S.deltak);
    for (int i = 0; i < n; i++) { S.c_p[i] = -g1[i] + beta * S.pk[i]; S.c_g[i] = g1[i]; }
    S.c_gnorm = amax_abs(g1, n);
    S.cached_alpha = alpha; S.cached_valid = 1;
}
// descent_condition (_optimize.py:1817-1832)
OI_HD bool descent_condition(OiCgState& S, double alpha, const double* g1) {
    pr_step(S, alpha, g1);
    if (S.c_gnorm <= S.gtol) return true;
    return dot(S.c_p, S.c_g, S.dim) <= -SIGMA3 * dot(S.c_g, S.c_g, S.dim);
}
// MemoizeJac: a step so small that x + alpha*p == x costs no new evaluation
OI_HD bool memo_hit(const OiCgState& S) {
    if (!S.memo_valid) return false;
    for (int i = 0; i < S.dim; i++) if (!(S.req_x[i] == S.memo_x[i])) return false;
    return true;
}
OI_HD void memo_store(OiCgState& S, double f, const double* g) {
    for (int i = 0; i < S.dim; i++) { S.memo_x[i] = S.req_x[i]; S.memo_g[i] = g[i]; }
    S.memo_f = f; S.memo_valid = 1; S.nfev++;
}
}  // namespace oicg

OI_HD void oi_cg_init(OiCgState& S, const double* x0, int dim, int maxiter, double gtol) {
    S.pc = 0; S.dim = dim; S.status = -1; S.k = 0; S.nfev = 0;
    S.maxiter = maxiter > 0 ? maxiter : 200 * dim;      // _optimize.py:1773-1774
    S.gtol = gtol;
    for (int i = 0; i < OI_MAXH; i++) { S.xk[i] = i < dim ? x0[i] : 0.0; S.pk[i] = 0; S.gfk[i] = 0; S.req_x[i] = S.xk[i]; }
    S.ev_valid = 0; S.cached_valid = 0; S.memo_valid = 0;
}

#define OICG_YIELD_EVAL(alpha_)                                                              \
    do {                                                                                     \
        S.ev_alpha = (alpha_);                                                               \
        for (int i_ = 0; i_ < S.dim; i_++) S.req_x[i_] = S.xk[i_] + S.ev_alpha * S.pk[i_];   \
        if (!memo_hit(S)) {                                                                  \
            S.pc = __LINE__; return OI_CG_NEED_EVAL; case __LINE__:;                         \
            memo_store(S, f_in, g_in);                                                       \
        }                                                                                    \
        S.ev_f = S.memo_f; for (int i_ = 0; i_ < S.dim; i_++) S.ev_g[i_] = S.memo_g[i_];     \
        S.ev_valid = 1;                                                                      \
    } while (0)
#define OICG_ENSURE_EVAL(alpha_)                                                             \
    do { if (!(S.ev_valid && S.ev_alpha == (alpha_))) OICG_YIELD_EVAL(alpha_); } while (0)

// Resume the optimiser.  First call: pc==0, (f_in, g_in) ignored; it returns NEED_EVAL with
// S.req_x = x0.  Every later call passes the objective and gradient at S.req_x.
OI_HD int oi_cg_resume(OiCgState& S, double f_in, const double* g_in) {
    using namespace oicg;
    const int n = S.dim;
    switch (S.pc) {
    case 0:
        for (int i = 0; i < n; i++) S.req_x[i] = S.xk[i];
        S.pc = 1; return OI_CG_NEED_EVAL;
    case 1:
        memo_store(S, f_in, g_in);
        S.old_fval = f_in;
        for (int i = 0; i < n; i++) { S.gfk[i] = g_in[i]; S.pk[i] = -g_in[i]; }
        S.k = 0;
        S.old_old_fval = S.old_fval + sqrt(dot(S.gfk, S.gfk, n)) / 2;     // :1788
        S.gnorm = amax_abs(S.gfk, n);
        S.status = 0;
        while ((S.gnorm > S.gtol) && (S.k < S.maxiter)) {
            S.deltak = dot(S.gfk, S.gfk, n);
            S.cached_valid = 0;
            S.ev_valid = 0;
            S.derphi0 = dot(S.gfk, S.pk, n);
            S.phi0 = S.old_fval;
            // ---------------- line_search_wolfe1 / scalar_search_wolfe1 ----------------
            for (int i = 0; i < n; i++) S.gfkp1[i] = S.gfk[i];   // gval = [gfk]
            if (S.derphi0 != 0) {
                S.d_alpha1 = pymin(1.0, 1.01 * 2 * (S.phi0 - S.old_old_fval) / S.derphi0);
                if (S.d_alpha1 < 0) S.d_alpha1 = 1.0;
            } else S.d_alpha1 = 1.0;
            S.d_phi1 = S.phi0; S.d_derphi1 = S.derphi0; S.d_task = T_START; S.ls_ok = 0;
            S.d_stp = S.d_alpha1;
            for (S.d_i = 0; S.d_i < 100; S.d_i++) {
                S.d_stp = S.d_alpha1;
                dcsrch_iterate(S, S.d_stp, S.d_phi1, S.d_derphi1, S.d_task);
                if (!finite_(S.d_stp)) { S.d_task = T_WARN; break; }
                if (S.d_task == T_FG) {
                    S.d_alpha1 = S.d_stp;
                    OICG_YIELD_EVAL(S.d_stp);
                    S.d_phi1 = S.ev_f;
                    for (int i = 0; i < n; i++) S.gfkp1[i] = S.ev_g[i];
                    S.d_derphi1 = dot(S.ev_g, S.pk, n);
                } else break;
            }
            if (S.d_i >= 100) S.d_task = T_WARN;            // for/else: did not converge
            S.ls_ok = (S.d_task == T_CONV);
            if (S.ls_ok) {
                // extra_condition (descent_condition) on the wolfe1 result (_optimize.py:1176-1180)
                if (!descent_condition(S, S.d_stp, S.gfkp1)) S.ls_ok = 0;
            }
            if (S.ls_ok) {
                S.alpha_k = S.d_stp; S.ls_fval = S.d_phi1; S.have_gfkp1 = 1;
            } else {
                // ---------------- line_search_wolfe2 / scalar_search_wolfe2 ----------------
                S.w_alpha0 = 0;
                if (S.derphi0 != 0) S.w_alpha1 = pymin(1.0, 1.01 * 2 * (S.phi0 - S.old_old_fval) / S.derphi0);
                else S.w_alpha1 = 1.0;
                if (S.w_alpha1 < 0) S.w_alpha1 = 1.0;
                S.w_alpha1 = pymin(S.w_alpha1, AMAX);
                OICG_ENSURE_EVAL(S.w_alpha1);
                S.w_phi_a1 = S.ev_f;
                S.w_phi_a0 = S.phi0; S.w_derphi_a0 = S.derphi0;
                S.w_zoom = 0; S.w_fail = 0; S.w_have_star = 0; S.w_have_dstar = 0;
                for (S.w_i = 0; S.w_i < 10; S.w_i++) {
                    if (S.w_alpha1 == 0 || S.w_alpha0 > AMAX) { S.w_fail = 1; break; }
                    if ((S.w_phi_a1 > S.phi0 + C1 * S.w_alpha1 * S.derphi0) ||
                        ((S.w_phi_a1 >= S.w_phi_a0) && (S.w_i > 0))) {
                        S.z_alo = S.w_alpha0; S.z_ahi = S.w_alpha1; S.z_philo = S.w_phi_a0; S.z_phihi = S.w_phi_a1;
                        S.z_dlo = S.w_derphi_a0; S.w_zoom = 1; break;
                    }
                    OICG_ENSURE_EVAL(S.w_alpha1);
                    S.w_derphi_a1 = dot(S.ev_g, S.pk, n);
                    if (fabs(S.w_derphi_a1) <= -C2 * S.derphi0) {
                        if (descent_condition(S, S.w_alpha1, S.ev_g)) {
                            S.w_alpha_star = S.w_alpha1; S.w_phi_star = S.w_phi_a1;
                            S.w_have_star = 1; S.w_have_dstar = 1;
                            for (int i = 0; i < n; i++) S.gfkp1[i] = S.ev_g[i];
                            break;
                        }
                    }
                    if (S.w_derphi_a1 >= 0) {
                        S.z_alo = S.w_alpha1; S.z_ahi = S.w_alpha0; S.z_philo = S.w_phi_a1; S.z_phihi = S.w_phi_a0;
                        S.z_dlo = S.w_derphi_a1; S.w_zoom = 1; break;
                    }
                    S.w_alpha0 = S.w_alpha1;
                    S.w_alpha1 = pymin(2 * S.w_alpha1, AMAX);
                    S.w_phi_a0 = S.w_phi_a1;
                    OICG_YIELD_EVAL(S.w_alpha1);
                    S.w_phi_a1 = S.ev_f;
                    S.w_derphi_a0 = S.w_derphi_a1;
                }
                if (!S.w_zoom && !S.w_fail && !S.w_have_star) {
                    // for/else: maxiter reached -> alpha_star = alpha1, derphi_star = None
                    S.w_alpha_star = S.w_alpha1; S.w_phi_star = S.w_phi_a1; S.w_have_star = 1; S.w_have_dstar = 0;
                }
                if (S.w_zoom) {
                    // ---------------- _zoom ----------------
                    S.z_i = 0; S.z_phirec = S.phi0; S.z_arec = 0;
                    while (true) {
                        {
                            double dalpha = S.z_ahi - S.z_alo, a, b, cchk = 0, qchk, aj = 0;
                            bool have = false;
                            if (dalpha < 0) { a = S.z_ahi; b = S.z_alo; } else { a = S.z_alo; b = S.z_ahi; }
                            if (S.z_i > 0) {
                                cchk = 0.2 * dalpha;
                                have = cubicmin(S.z_alo, S.z_philo, S.z_dlo, S.z_ahi, S.z_phihi, S.z_arec, S.z_phirec, aj);
                            }
                            if ((S.z_i == 0) || !have || (aj > b - cchk) || (aj < a + cchk)) {
                                qchk = 0.1 * dalpha;
                                have = quadmin(S.z_alo, S.z_philo, S.z_dlo, S.z_ahi, S.z_phihi, aj);
                                if (!have || (aj > b - qchk) || (aj < a + qchk)) aj = S.z_alo + 0.5 * dalpha;
                            }
                            S.z_aj = aj;
                        }
                        OICG_YIELD_EVAL(S.z_aj);
                        S.z_phiaj = S.ev_f;
                        if ((S.z_phiaj > S.phi0 + C1 * S.z_aj * S.derphi0) || (S.z_phiaj >= S.z_philo)) {
                            S.z_phirec = S.z_phihi; S.z_arec = S.z_ahi; S.z_ahi = S.z_aj; S.z_phihi = S.z_phiaj;
                        } else {
                            S.z_daj = dot(S.ev_g, S.pk, n);
                            if (fabs(S.z_daj) <= -C2 * S.derphi0 && descent_condition(S, S.z_aj, S.ev_g)) {
                                S.w_alpha_star = S.z_aj; S.w_phi_star = S.z_phiaj; S.w_have_star = 1; S.w_have_dstar = 1;
                                for (int i = 0; i < n; i++) S.gfkp1[i] = S.ev_g[i];
                                break;
                            }
                            if (S.z_daj * (S.z_ahi - S.z_alo) >= 0) {
                                S.z_phirec = S.z_phihi; S.z_arec = S.z_ahi; S.z_ahi = S.z_alo; S.z_phihi = S.z_philo;
                            } else { S.z_phirec = S.z_philo; S.z_arec = S.z_alo; }
                            S.z_alo = S.z_aj; S.z_philo = S.z_phiaj; S.z_dlo = S.z_daj;
                        }
                        S.z_i++;
                        if (S.z_i > 10) break;     // a_star = None
                    }
                }
                if (!S.w_have_star) { S.status = 2; break; }      // _LineSearchError -> warnflag 2
                S.alpha_k = S.w_alpha_star; S.ls_fval = S.w_phi_star; S.have_gfkp1 = S.w_have_dstar;
            }
            // old_fval, old_old_fval = ret[3], ret[4]
            S.old_old_fval = S.phi0;
            S.old_fval = S.ls_fval;
            if (!(S.cached_valid && S.alpha_k == S.cached_alpha)) {
                if (!S.have_gfkp1) {
                    OICG_ENSURE_EVAL(S.alpha_k);
                    for (int i = 0; i < n; i++) S.gfkp1[i] = S.ev_g[i];
                }
                pr_step(S, S.alpha_k, S.gfkp1);
            }
            for (int i = 0; i < n; i++) { S.xk[i] = S.c_x[i]; S.pk[i] = S.c_p[i]; S.gfk[i] = S.c_g[i]; }
            S.gnorm = S.c_gnorm;
            S.k++;
        }
        if (S.status != 2) {
            if (S.k >= S.maxiter) S.status = 1;
            else {
                bool anynan = (S.gnorm != S.gnorm) || (S.old_fval != S.old_fval);
                for (int i = 0; i < n; i++) anynan = anynan || (S.xk[i] != S.xk[i]);
                S.status = anynan ? 3 : 0;
            }
        }
        S.pc = -1;
        return OI_CG_DONE;
    default:
        return OI_CG_DONE;
    }
}
