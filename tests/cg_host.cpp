// Host build of the optimiser state machine (optimalinterpolation_b200/csrc/cg_scipy.h) for CPU tests:
// tests drive it with a Python objective and compare against scipy.optimize.minimize(method='CG').
#include "../optimalinterpolation_b200/csrc/cg_scipy.h"
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
}
