// TMA probe (measurement tool, not part of the product library): the NT DMMA K loop of csrc/oi_tiles.cuh with its operand
// chunks staged (a) by the product's cp.async pipeline (gemm_nt_stream, SASS LDGSTS) and (b) by TMA
// (cp.async.bulk.tensor.2d + mbarrier complete_tx, SASS UTMALDG) into the SAME shared-memory layout -- the product's
// XOR swizzle of 16-byte pieces by (row & 7) is TMA's SWIZZLE_128B for 128-byte rows -- so both feed the identical
// mma_chunk code and must give bit-identical tiles.  BASELINE.json's north_star asks for TMA staging; DESIGN.md §5 argues
// the K loops are DMMA-bound and TMA would buy little: this measures it.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I optimalinterpolation_b200/csrc -fmad=false \
//        -o tools/tma_probe tools/tma_probe.cu          (no -lcuda: the driver entry point is fetched at run time)
//   tools/tma_probe [n=1536] [reps=5]
//
// Workload: C(i,j) = A_i A_j^T over the full K = n for every 64x64 tile (i,j) of an n x n row-major FP64 matrix
// (N^2 CTAs, N = n/64), i.e. the shape of k_trtri / k_lauum_trace tiles at their longest K.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "oi_tiles.cuh"

#define CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

__global__ void __launch_bounds__(GEMM_THREADS, 4) k_nt_cpasync(const double* __restrict__ M, double* __restrict__ C, int n) {
    extern __shared__ __align__(16) double smem[];
    const int i = blockIdx.y, j = blockIdx.x;
    double acc[4][4][2];
    ACC_ZERO(acc);
    gemm_nt_stream(acc, M + (long long)i * NB * n, n, M + (long long)j * NB * n, n, 0, n, smem,
                   [](int) { return SubRange{0, 4, 0, 4}; });
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    double* Cg = C + (long long)i * NB * n + (long long)j * NB;
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 v; v.x = acc[mb][nb][0]; v.y = acc[mb][nb][1];
            *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * n + FRAG_COL(wn, nb, lane)] = v;
        }
}

// ---- TMA + mbarrier helpers (raw PTX) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// returns false if the barrier did not complete within the spin limit (a probe must never hang the GPU)
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    for (int spin = 0; spin < (1 << 20); spin++) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return true;
    }
    return false;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int x, int y, unsigned long long* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(GEMM_THREADS, 4) k_nt_tma(const __grid_constant__ CUtensorMap map, double* __restrict__ C, int n, int* __restrict__ err) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    // SWIZZLE_128B needs 1024-byte aligned destinations: align the dynamic window by hand (1 KB of slack is requested)
    double* smem = (double*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    const int i = blockIdx.y, j = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const int nk = n / KT;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < nk; s++) {
            mbar_expect_tx(&full[s], STAGE_DOUBLES * 8);
            tma_load_2d(smem + s * STAGE_DOUBLES, &map, s * KT, i * NB, &full[s]);
            tma_load_2d(smem + s * STAGE_DOUBLES + NB * KT, &map, s * KT, j * NB, &full[s]);
        }
    }
    double acc[4][4][2];
    ACC_ZERO(acc);
    bool ok = true;
    for (int it = 0; it < nk; it++) {
        const int s = it % STAGES;
        ok = mbar_wait(&full[s], (it / STAGES) & 1);
        if (!__syncthreads_and(ok)) { ok = false; break; }      // a barrier timed out: give up on this tile (reported)
        const double* st = smem + s * STAGE_DOUBLES;
        mma_chunk_t<0, 4, 0, 4>(acc, st, st + NB * KT, wm, wn, lane);
        __syncthreads();                                   // every warp is done with stage s: it may be refilled
        if (tid == 0 && it + STAGES < nk) {
            mbar_expect_tx(&full[s], STAGE_DOUBLES * 8);
            tma_load_2d(smem + s * STAGE_DOUBLES, &map, (it + STAGES) * KT, i * NB, &full[s]);
            tma_load_2d(smem + s * STAGE_DOUBLES + NB * KT, &map, (it + STAGES) * KT, j * NB, &full[s]);
        }
    }
    if (!ok && lane == 0) atomicAdd(err, 1);
    double* Cg = C + (long long)i * NB * n + (long long)j * NB;
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 v; v.x = acc[mb][nb][0]; v.y = acc[mb][nb][1];
            *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * n + FRAG_COL(wn, nb, lane)] = v;
        }
}

// Variant 2: no CTA barrier in the K loop.  "full" barriers as above, plus "empty" barriers (one arrival per warp): a warp
// that is done with a stage says so and moves on; only the issuing thread waits for all four before it refills the stage.
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__global__ void __launch_bounds__(GEMM_THREADS, 4) k_nt_tma2(const __grid_constant__ CUtensorMap map, double* __restrict__ C, int n, int* __restrict__ err) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long full[STAGES], empty[STAGES];
    double* smem = (double*)(((size_t)smem_raw + 1023) & ~(size_t)1023);
    const int i = blockIdx.y, j = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wm = warp >> 1, wn = warp & 1;
    const int nk = n / KT;
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], GEMM_THREADS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < STAGES && s < nk; s++) {
            mbar_expect_tx(&full[s], STAGE_DOUBLES * 8);
            tma_load_2d(smem + s * STAGE_DOUBLES, &map, s * KT, i * NB, &full[s]);
            tma_load_2d(smem + s * STAGE_DOUBLES + NB * KT, &map, s * KT, j * NB, &full[s]);
        }
    }
    double acc[4][4][2];
    ACC_ZERO(acc);
    bool ok = true;
    for (int it = 0; it < nk && ok; it++) {
        const int s = it % STAGES;
        const unsigned par = (it / STAGES) & 1;
        ok = __all_sync(0xffffffffu, mbar_wait(&full[s], par));      // warp-uniform
        if (!ok) break;
        const double* st = smem + s * STAGE_DOUBLES;
        mma_chunk_t<0, 4, 0, 4>(acc, st, st + NB * KT, wm, wn, lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (tid == 0 && it + STAGES < nk) {
            ok = mbar_wait(&empty[s], par);
            if (ok) {
                mbar_expect_tx(&full[s], STAGE_DOUBLES * 8);
                tma_load_2d(smem + s * STAGE_DOUBLES, &map, (it + STAGES) * KT, i * NB, &full[s]);
                tma_load_2d(smem + s * STAGE_DOUBLES + NB * KT, &map, (it + STAGES) * KT, j * NB, &full[s]);
            }
        }
        ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) != 0;      // thread 0's verdict for its own warp; other warps keep theirs
    }
    if (!ok && lane == 0) atomicAdd(err, 1);
    double* Cg = C + (long long)i * NB * n + (long long)j * NB;
#pragma unroll
    for (int mb = 0; mb < 4; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) {
            double2 v; v.x = acc[mb][nb][0]; v.y = acc[mb][nb][1];
            *(double2*)&Cg[(long long)FRAG_ROW(wm, mb, lane) * n + FRAG_COL(wn, nb, lane)] = v;
        }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 1536, reps = argc > 2 ? std::atoi(argv[2]) : 5;
    if (n % NB) { std::printf("n must be a multiple of %d\n", NB); return 1; }
    const int N = n / NB;
    std::vector<double> h((size_t)n * n);
    unsigned long long sd = 12345;
    for (auto& v : h) { sd = sd * 6364136223846793005ULL + 1442695040888963407ULL; v = (double)((sd >> 11) & 0xfffff) / 1048576.0 - 0.5; }
    double *M, *C1, *C2, *C3; int* err;
    CHECK(cudaMalloc(&M, (size_t)n * n * 8)); CHECK(cudaMalloc(&C1, (size_t)n * n * 8)); CHECK(cudaMalloc(&C2, (size_t)n * n * 8)); CHECK(cudaMalloc(&C3, (size_t)n * n * 8));
    CHECK(cudaMalloc(&err, 4)); CHECK(cudaMemset(err, 0, 4));
    CHECK(cudaMemcpy(M, h.data(), (size_t)n * n * 8, cudaMemcpyHostToDevice));
    CHECK(cudaMemset(C1, 0, (size_t)n * n * 8)); CHECK(cudaMemset(C2, 0xff, (size_t)n * n * 8)); CHECK(cudaMemset(C3, 0xff, (size_t)n * n * 8));

    void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
    CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { std::printf("cuTensorMapEncodeTiled not available\n"); return 3; }
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)n};                 // innermost (columns) first
    cuuint64_t strides[1] = {(cuuint64_t)n * 8};                          // bytes between rows
    cuuint32_t box[2] = {(cuuint32_t)KT, (cuuint32_t)NB};                 // 16 doubles = 128 B wide, 64 rows
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = ((EncodeTiledFn)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, M, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { std::printf("cuTensorMapEncodeTiled failed: %d\n", (int)cr); return 3; }

    CHECK(cudaFuncSetAttribute(k_nt_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES));
    CHECK(cudaFuncSetAttribute(k_nt_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES + 1024));
    CHECK(cudaFuncSetAttribute(k_nt_tma2, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_BYTES + 1024));
    dim3 grid(N, N);
    cudaEvent_t e0, e1; CHECK(cudaEventCreate(&e0)); CHECK(cudaEventCreate(&e1));
    const double flops = 2.0 * (double)n * n * n;
    float best[3] = {1e30f, 1e30f, 1e30f};
    for (int r = 0; r < reps + 1; r++) {
        CHECK(cudaEventRecord(e0)); k_nt_cpasync<<<grid, GEMM_THREADS, PIPE_BYTES>>>(M, C1, n); CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1)); float ms; CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best[0] = ms < best[0] ? ms : best[0];
        CHECK(cudaEventRecord(e0)); k_nt_tma<<<grid, GEMM_THREADS, PIPE_BYTES + 1024>>>(map, C2, n, err); CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1)); CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best[1] = ms < best[1] ? ms : best[1];
        CHECK(cudaEventRecord(e0)); k_nt_tma2<<<grid, GEMM_THREADS, PIPE_BYTES + 1024>>>(map, C3, n, err); CHECK(cudaEventRecord(e1));
        CHECK(cudaEventSynchronize(e1)); CHECK(cudaEventElapsedTime(&ms, e0, e1)); if (r) best[2] = ms < best[2] ? ms : best[2];
        CHECK(cudaGetLastError());
    }
    int herr = 0; CHECK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
    std::vector<double> c1((size_t)n * n), c2((size_t)n * n), c3((size_t)n * n);
    CHECK(cudaMemcpy(c3.data(), C3, (size_t)n * n * 8, cudaMemcpyDeviceToHost));
    CHECK(cudaMemcpy(c1.data(), C1, (size_t)n * n * 8, cudaMemcpyDeviceToHost)); CHECK(cudaMemcpy(c2.data(), C2, (size_t)n * n * 8, cudaMemcpyDeviceToHost));
    double maxd = 0, ref = 0; size_t nbad = 0, nbad3 = 0;
    for (size_t q = 0; q < c1.size(); q++) if (!(c1[q] - c3[q] == 0.0)) nbad3++;
    for (size_t q = 0; q < c1.size(); q++) { double dlt = c1[q] - c2[q]; if (!(dlt == 0.0)) nbad++; if (dlt < 0) dlt = -dlt; if (dlt > maxd || dlt != dlt) maxd = dlt; if (c1[q] > ref) ref = c1[q]; }
    // spot check of the cp.async result against a host dot product
    double hs = 0; for (int k = 0; k < n; k++) hs += h[(size_t)5 * n + k] * h[(size_t)70 * n + k];
    std::printf("n %d (N %d, %d CTAs, K loop %d chunks), %d reps, best of\n", n, N, N * N, n / KT, reps);
    std::printf("  cp.async pipeline (product code, LDGSTS): %8.3f ms  %6.2f TFLOP/s\n", best[0], flops / best[0] * 1e-9);
    std::printf("  TMA + mbarrier      (UTMALDG)           : %8.3f ms  %6.2f TFLOP/s   (%+.1f %% time)\n", best[1], flops / best[1] * 1e-9, 100.0 * (best[1] / best[0] - 1.0));
    std::printf("  TMA, no CTA barrier (full/empty mbarriers)  : %8.3f ms  %6.2f TFLOP/s   (%+.1f %% time), %zu elements differ\n", best[2], flops / best[2] * 1e-9, 100.0 * (best[2] / best[0] - 1.0), nbad3);
    std::printf("  tiles differing between the two paths: %zu elements, max |diff| %.3e (max value %.3e); barrier time-outs %d; host check C[5][70] %.12g vs %.12g\n",
                nbad, maxd, ref, herr, hs, c1[(size_t)5 * n + 70]);
    return (nbad == 0 && nbad3 == 0 && herr == 0) ? 0 : 4;
}
