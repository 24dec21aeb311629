/* oi_b200.h — C ABI of the B200-native per-grid-cell GP regression hot path.
 *
 * Drop-in boundary for /root/reference/2021_paper_production/GPR_CS2S3.py.  The reference has no
 * FFI/plugin interface; its seam is the Python function GPR3D(index, opt) (GPR_CS2S3.py:143-191)
 * called once per cell from the per-rank loops (:258-261 and :317-319) and reading the module
 * globals set up at :201-246.  A whole day is passed in one call because batching is the point:
 *
 *   reference                                          this ABI
 *   ---------                                          --------
 *   x_train,y_train,t_train,z  (:238-241)              oi_set_observations()
 *   sat = obs[:,:,:,day:day+T] (:213)                  oi_set_time_window() on a resident season (optional)
 *   X = ice-cell coordinates   (:244)                  oi_set_cells()
 *   X_tree.query_ball_point    (:159, :245-246)        oi_gather_neighbours() / oi_get_neighbours()
 *   SMLII(hypers, x, y, mX)    (:107-141)              oi_nlml_grad()
 *   GPR3D(index, opt=True)     (:143-168, :173-184)    oi_run(OI_MODE_FIT) / oi_gpr_day()
 *   GPR3D(index, opt=False)    (:169-172, :185-186)    oi_run(OI_MODE_PREDICT) with hypers_in
 *   results tuple (:184)                               out[n_cells][8] = fs, sfs2(std), lZ, lx, ly, lt, sf2, sn2
 *   scipy.optimize.minimize(method='CG') (:166)        oi_params.optimiser = OI_OPT_CG (restated evaluation by evaluation);
 *                                                      OI_OPT_LBFGS = the fast mode (not the reference's stopping points)
 *   split(container, count) + COMM.scatter (:18-23,    one process per GPU: a static split on the host (shard.py) or
 *     :250-256), COMM.gather (:262)                    oi_set_shared_queue() / oi_get_owned(): one cost-sorted work list
 *                                                      shared by the processes of a box; the gather stays with the host (NCCL)
 *
 * Conventions: all pointers are HOST pointers to C-contiguous arrays; the library owns every
 * device allocation behind the opaque handle.  One handle per GPU; a handle is not thread-safe;
 * calls are synchronous.  Return value 0 = ok, negative = error (oi_last_error() gives the text).
 * A numerical failure inside one cell (non-positive Cholesky pivot) is DATA, not an error: that
 * cell gets NaN outputs (GPR_CS2S3.py:187-191) and status 3, the call still returns 0.
 * There is no CPU fallback: every entry point fails with OI_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef OI_B200_H
#define OI_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oi_handle oi_handle;

enum { OI_OK = 0, OI_ERR_ARG = -1, OI_ERR_CUDA = -2, OI_ERR_STATE = -3, OI_ERR_NOMEM = -4 };

/* mode */
enum { OI_MODE_FIT = 0,      /* GPR3D(index, opt=True):  fit by CG then predict  */
       OI_MODE_PREDICT = 1   /* GPR3D(index, opt=False): predict with hypers_in   */ };
/* execution engine.  LOCKSTEP: one launch per algorithmic step over all active cells, n_groups independent groups on
 * their own streams.  The experimental PERSISTENT engine of v1.1 (one resident kernel with a global-memory barrier;
 * measured slower, never race-checked) was removed in v1.2: oi_run refuses it. */
enum { OI_ENGINE_LOCKSTEP = 0, OI_ENGINE_PERSISTENT = 1 };
/* gradient convention of SMLII: the reference's components 3 and 4 are twice the true derivative
 * (GPR_CS2S3.py:135-138).  REFERENCE reproduces that; EXACT gives the true gradient. */
enum { OI_GRAD_REFERENCE = 0, OI_GRAD_EXACT = 1 };
/* optimiser of OI_MODE_FIT.  CG restates scipy.optimize.minimize(method='CG') evaluation by evaluation (GPR_CS2S3.py:166)
 * and is the parity mode.  LBFGS is the fast mode BASELINE.json's north_star (5) asks for: limited-memory BFGS (m = 8,
 * More'-Thuente line search, scipy L-BFGS-B's stopping rules, no bounds as in the reference) on the same log
 * hyperparameters; use it with OI_GRAD_EXACT.  It reaches the same or a lower NLML in ~4x fewer evaluations but does
 * not reproduce the reference's stopping points. */
enum { OI_OPT_CG = 0, OI_OPT_LBFGS = 1 };
/* per-cell status (scipy's warnflag where it applies) */
enum { OI_CELL_OK = 0, OI_CELL_MAXITER = 1, OI_CELL_LINESEARCH = 2, OI_CELL_CHOL_FAIL = 3, OI_CELL_NO_OBS = 4,
       OI_CELL_NAN = 5 };

typedef struct oi_params {
    double radius_m;        /* search radius in metres            (radius*1000, GPR_CS2S3.py:159, :208) */
    double t_pred;          /* prediction time coordinate          (T_mid, :207, :164)                  */
    double prior_mean;      /* constant prior mean                 (mean, :212, :163)                   */
    int32_t n_hyp;          /* len(x0): 5 or 6                     (:217 has 6, the last one is dead)    */
    int32_t mode;           /* OI_MODE_*                                                                */
    int32_t grad_convention;/* OI_GRAD_*                                                                */
    int32_t maxiter;        /* 0 => 200*n_hyp (scipy default)                                           */
    double x0[6];           /* initial LOG hyperparameters         (:217)                               */
    double gtol;            /* 0 => 1e-5 (scipy default)                                                */
    double scratch_gib;     /* device scratch budget for the lockstep batch, 0 => automatic             */
    int32_t max_active;     /* maximum cells evaluated per lockstep iteration, 0 => automatic           */
    int32_t n_groups;       /* lockstep engine: independent groups (one CUDA stream each) whose kernels overlap, 0 => automatic */
    int32_t engine;         /* must be OI_ENGINE_LOCKSTEP (0)                                              */
    int32_t group_size;     /* unused (was: persistent engine, removed in v1.2); kept for the struct layout */
    int32_t evals_per_launch; /* unused (same)                                                              */
    int32_t optimiser;      /* OI_OPT_*: 0 = the reference's scipy CG (parity mode), 1 = exact-gradient L-BFGS (fast mode) */
} oi_params;

typedef struct oi_stats {
    double ms_total;        /* device time of the last oi_run (CUDA events)                              */
    double ms_gather;       /* neighbour gather part                                                     */
    double flops;           /* algorithmic FP64 flops of the last oi_run (SURVEY.md §8d formula)          */
    int64_t n_evals;        /* total NLML+gradient evaluations                                           */
    int64_t n_launches;     /* kernels launched                                                          */
    int64_t n_iterations;   /* lockstep iterations                                                       */
    int64_t sum_n;          /* sum of neighbour counts                                                   */
    double ms_factor;       /* device time inside the Cholesky/TRTRI/LAUUM kernels (CUDA events)          */
    double flops_factor;    /* algorithmic flops of those kernels                                        */
    double ms_build, ms_chol, ms_fwd, ms_trtri, ms_alpha, ms_lauum, ms_finalize;  /* per kernel family   */
    double flops_chol, flops_trtri, flops_lauum;   /* n^3/3 each per evaluation            */
    int64_t launches_chol, launches_trtri, launches_lauum;
    int64_t n_groups;       /* groups used (last launch); lockstep engine with > 1: the per-family ms_* are per-stream times that overlap */
    int64_t group_size;     /* unused since v1.2 (layout kept)                                            */
    int64_t launches_persistent; /* unused since v1.2                                                     */
    int64_t n_graph_captures, n_graph_launches; /* lockstep engine: small-batch iterations replayed as CUDA graphs */
    double ms_graph;        /* device time of those iterations (not split by kernel family)              */
    int64_t n_express_cells; /* lockstep engine: cells handed to the express lanes (long optimiser runs)           */
    double ms_persistent;   /* unused since v1.2                                                         */
    double cycles_phase[8]; /* unused since v1.2                                                         */
} oi_stats;

int  oi_version(void);
int  oi_sizeof_params(void);   /* sizeof(oi_params) / sizeof(oi_stats) of the built library (binding self-check) */
int  oi_sizeof_stats(void);
const char* oi_last_error(void);

int  oi_create(int device, oi_handle** out);
/* Launch on a caller-owned CUDA stream (cudaStream_t passed as void*), e.g. a torch.cuda.Stream, so that the caller's
 * CUDA events bracket the kernels.  NULL (0) restores the handle's own stream: the legacy default stream (whose handle is
 * 0) cannot be selected -- pass a stream you created (bench.py does). */
int  oi_set_stream(oi_handle* h, void* cuda_stream);
void oi_destroy(oi_handle* h);

/* Observations of the day window: x_train, y_train, t_train, z (GPR_CS2S3.py:238-241). */
int oi_set_observations(oi_handle* h, const double* x, const double* y, const double* t, const double* z,
                        int64_t n_obs);
/* Day window of the neighbour gather: only observations with t_lo <= t <= t_hi take part, and their time coordinate
 * is counted from t_lo (the reference slices obs[:, :, :, day:day+T] and numbers the days of the window 0..T-1,
 * GPR_CS2S3.py:213, :227-235).  With it a whole season of flattened observations (stream-major, day-major, row-major,
 * t = absolute day index) stays resident and every day only moves the window: oi_set_time_window(h, day, day+T-1).
 * Default: no window (-inf, +inf), t used as given. */
int oi_set_time_window(oi_handle* h, double t_lo, double t_hi);
/* Target cells X[n_cells][2] (GPR_CS2S3.py:244). */
int oi_set_cells(oi_handle* h, const double* X, int64_t n_cells);

/* Kernel (1): for every cell the set {i : dx*dx + dy*dy <= r*r} (inclusive, :159), as CSR in
 * ascending observation order.  counts_out[n_cells] may be NULL. */
int oi_gather_neighbours(oi_handle* h, double radius_m, int32_t* counts_out);
/* Copy the CSR back: offsets[n_cells+1], indices[sum n] (either may be NULL). */
int oi_get_neighbours(oi_handle* h, int64_t* offsets, int32_t* indices);

/* SMLII (GPR_CS2S3.py:107-141) for every cell at its own LOG hyperparameters hypers[n_cells][n_hyp]:
 * nlz_out[n_cells], grad_out[n_cells][n_hyp].  Needs oi_gather_neighbours first. */
int oi_nlml_grad(oi_handle* h, const double* hypers, int32_t n_hyp, double prior_mean, int32_t grad_convention,
                 double* nlz_out, double* grad_out);

/* Fit+predict or predict-only on the resident day (needs observations, cells and neighbours).
 * hypers_in[n_cells][5] (natural units lx,ly,lt,sf2,sn2) is read only in OI_MODE_PREDICT.
 * Results stay on the device until oi_get_results. */
int oi_run(oi_handle* h, const oi_params* p, const double* hypers_in);
/* out[n_cells][8]; nfev_out/status_out/n_out[n_cells] may be NULL. */
int oi_get_results(oi_handle* h, double* out, int32_t* n_out, int32_t* nfev_out, int32_t* status_out);
int oi_get_stats(oi_handle* h, oi_stats* s);
/* Several GPU processes of ONE box (one per GPU) share one cost-sorted work list instead of a static split of the cells
 * (reference: split(container, count), GPR_CS2S3.py:18-23, :250-256): a 64-bit word in POSIX shared memory holds the two
 * cursors of the list, the largest cells are claimed from the front, the smallest from the back, each cell is computed by
 * exactly one process (csrc/oi_shared_queue.h).  Every process calls oi_set_shared_queue with the same fresh, unique name
 * ("/something") before its first oi_run and then runs the SAME calls on the SAME observations and cells; oi_get_owned
 * tells which rows of oi_get_results are this process's (the others are NaN).  NULL detaches; oi_unlink_shared_queue removes
 * the segment (one process, at the end). */
int oi_set_shared_queue(oi_handle* h, const char* shm_name);
int oi_unlink_shared_queue(const char* shm_name);
int oi_get_owned(oi_handle* h, uint8_t* owned);
/* Diagnostic (tools/cg_diagnose.py): record every objective evaluation the optimiser of ONE cell sees during the next
 * OI_MODE_FIT runs -- rows of 12 doubles: natural-unit hyperparameters (5), value, gradient (6).  cell < 0: off. */
int oi_debug_trace(oi_handle* h, int64_t cell, int32_t capacity);
int oi_get_debug_trace(oi_handle* h, double* rows, int32_t* n_rows);

/* The whole day in one call: set_observations + set_cells + gather + run + get_results. */
int oi_gpr_day(oi_handle* h, const double* x, const double* y, const double* t, const double* z, int64_t n_obs,
               const double* X, int64_t n_cells, const oi_params* p, const double* hypers_in,
               double* out, int32_t* n_out, int32_t* nfev_out, int32_t* status_out);

#ifdef __cplusplus
}
#endif
#endif /* OI_B200_H */
