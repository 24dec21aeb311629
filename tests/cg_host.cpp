// Host build of the optimiser state machine (optimalinterpolation_b200/csrc/cg_scipy.h) for CPU tests:
// tests drive it with a Python objective and compare against scipy.optimize.minimize(method='CG').
#include "../optimalinterpolation_b200/csrc/cg_scipy.h"
#include "../optimalinterpolation_b200/csrc/lbfgs_fast.h"
#include <cstdlib>
extern "C" {
OiCgState* cgh_new() { return (OiCgState*)calloc(1, sizeof(OiCgState)); }
void cgh_free(OiCgState* s) { free(s); }
void cgh_init(OiCgState* s, const double* x0, int dim, int maxiter, double gtol) { oi_cg_init(*s, x0, dim, maxiter, gtol); }
int cgh_resume(OiCgState* s, double f, const double* g) { return oi_cg_resume(*s, f, g); }
const double* cgh_req_x(OiCgState* s) { return s->req_x; }
const double* cgh_x(OiCgState* s) { return s->xk; }
double cgh_fval(OiCgState* s) { return s->old_fval; }
int cgh_status(OiCgState* s) { return s->status; }
int cgh_nit(OiCgState* s) { return s->k; }
int cgh_nfev(OiCgState* s) { return s->nfev; }
// fast-mode optimiser (lbfgs_fast.h)
OiLbfgsState* lbh_new() { return (OiLbfgsState*)calloc(1, sizeof(OiLbfgsState)); }
void lbh_free(OiLbfgsState* s) { free(s); }
void lbh_init(OiLbfgsState* s, const double* x0, int dim, int maxiter, double pgtol) { oi_lbfgs_init(*s, x0, dim, maxiter, pgtol); }
int lbh_resume(OiLbfgsState* s, double f, const double* g) { return oi_lbfgs_resume(*s, f, g); }
const double* lbh_req_x(OiLbfgsState* s) { return s->req_x; }
const double* lbh_x(OiLbfgsState* s) { return s->xk; }
double lbh_fval(OiLbfgsState* s) { return s->fk; }
int lbh_status(OiLbfgsState* s) { return s->status; }
int lbh_nit(OiLbfgsState* s) { return s->k; }
int lbh_nfev(OiLbfgsState* s) { return s->nfev; }
}
// unit access to the interpolation helpers (rounding checks against scipy's Python, tests/test_cg.py)
extern "C" {
int cgh_cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc, double* x) { return oicg::cubicmin(a, fa, fpa, b, fb, c, fc, *x) ? 1 : 0; }
int cgh_quadmin(double a, double fa, double fpa, double b, double fb, double* x) { return oicg::quadmin(a, fa, fpa, b, fb, *x) ? 1 : 0; }
}
extern "C" double cgh_fma(double a, double b, double c) { return fma(a, b, c); }
extern "C" double cgh_cube(double x) { return oicg::cube(x); }
extern "C" void cgh_dcstep(double* v, int* brackt, double fp, double dp, double stpmin, double stpmax) {
    // v = {stx, fx, dx, sty, fy, dy, stp}
    oicg::dcstep(v[0], v[1], v[2], v[3], v[4], v[5], v[6], fp, dp, *brackt, stpmin, stpmax);
}
// the work list shared by the GPU processes of a box (csrc/oi_shared_queue.h): host access for the multi-process CPU test
#include "../optimalinterpolation_b200/csrc/oi_shared_queue.h"
extern "C" {
OiSharedQueue* sq_new() { return new OiSharedQueue(); }
void sq_free(OiSharedQueue* q) { q->detach(); delete q; }
int sq_attach(OiSharedQueue* q, const char* name) { return q->attach(name); }
void sq_unlink(const char* name) { OiSharedQueue::unlink_name(name); }
void sq_begin(OiSharedQueue* q, unsigned n) { q->begin_run(n); }
long sq_take(OiSharedQueue* q, int small_end) {       // peek + claim until it succeeds or the list is empty
    for (;;) {
        long i = small_end ? q->peek_back() : q->peek_front();
        if (i < 0) return -1;
        if (small_end ? q->claim_back(i) : q->claim_front(i)) return i;
    }
}
}
