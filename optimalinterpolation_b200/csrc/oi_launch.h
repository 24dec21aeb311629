// Launch wrappers shared between the kernel translation units and the host API (oi_api.cu).
#pragma once
#include <cuda_runtime.h>
#include "oi_types.h"

#define OI_THREADS 128                 // threads of every tile CTA (4 warps, 2x2 warp tiles of 32x32)
#define OI_ROWWISE_MIN_SLOTS_HOST 96   // batches at least this big launch the tile-parallel kernels one block row at a time
#define OI_DEFAULT_GROUPS 8
#ifndef STAGES
#define STAGES 3
#endif
#define OI_SMEM_PIPE (STAGES * 2 * OI_NB * 16 * 8)              // cp.async pipeline stages of two swizzled 64x16 operand chunks

struct OiRunConst {
    double mean, gtol;
    int n_hyp, grad_convention, maxiter, optimiser;
    double x0[6];
};
// index ranges [lo, hi) of the observation arrays that the neighbour gather scans (ascending, disjoint)
#define OI_MAX_RANGES 16
struct OiRanges { int n; int lo[OI_MAX_RANGES], hi[OI_MAX_RANGES]; };
void oi_launch_count(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                     double t_lo, double t_hi, const OiRanges& rg, int* counts, cudaStream_t st);
void oi_launch_scan(const int* counts, int n, long long* offsets, cudaStream_t st);
void oi_launch_fill(const double* ox, const double* oy, const double* ot, int n_obs, const double* X, int n_cells, double r2,
                    double t_lo, double t_hi, const OiRanges& rg, const long long* offsets, int* indices, cudaStream_t st);
void oi_launch_pack(const int* indices, long long total, const double* ox, const double* oy, const double* ot,
                    const double* oz, double mean, double t_shift, double* px, double* py, double* pt, double* pr, cudaStream_t st);
void oi_launch_build(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st);
void oi_launch_chol_update(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st);
void oi_launch_chol_panel(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int k, cudaStream_t st);
void oi_launch_fwd(const OiSlot* slots, int A, OiCellArrays ca, OiPacked pk, double t_pred, cudaStream_t st);
void oi_launch_trtri(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, int d, const int* phase, cudaStream_t st);
void oi_launch_alpha(const OiSlot* slots, int A, int Nmax, const int* phase, cudaStream_t st);
void oi_launch_lauum_trace(const OiSlot* slots, int A, int Nmax, const int* cnt_gt, OiCellArrays ca, OiPacked pk, cudaStream_t st);

int oi_set_kernel_attributes();   // > 48 KB dynamic shared memory opt-in for the CURRENT device; returns a cudaError_t
void oi_launch_cg_init(OiCellArrays ca, int n_cells, OiRunConst rc, cudaStream_t st);
void oi_launch_finalize(const OiSlot* slots, int A, OiCellArrays ca, OiRunConst rc, int* slot_phase, cudaStream_t st);
