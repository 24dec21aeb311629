// FP64 peak microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync m8n8k4 / m16n8k16 f64).
// Decides which pipe the batched Cholesky/TRTRI/LAUUM tiles are written for.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__global__ void dmma884_kernel(double* out, int iters, double a, double b) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__global__ void dmma16816_kernel(double* out, int iters, double av, double bv) {
    double c[8][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = av + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = bv + i;
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + i + j;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F> float time_it(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s sms %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    double* out; cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024);
    int iters = 20000;
    for (int threads : {128, 256, 512, 1024}) {
        for (int bps : {1, 2}) {
            if (threads * bps > 2048) continue;
            int blocks = p.multiProcessorCount * bps;
            float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            double fl = 2.0 * 16 * iters * (double)threads * blocks;
            printf("DFMA      threads %4d x%d blocks/SM: %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            ms = time_it([&] { dmma884_kernel<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
            fl = 2.0 * 8 * 8 * 4 * 16 * iters * (double)(threads / 32) * blocks;
            printf("DMMA884   threads %4d x%d blocks/SM: %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
            ms = time_it([&] { dmma16816_kernel<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
            fl = 2.0 * 16 * 8 * 16 * 8 * (iters / 4) * (double)(threads / 32) * blocks;
            printf("DMMA16816 threads %4d x%d blocks/SM: %8.3f ms  %7.2f TFLOP/s\n", threads, bps, ms, fl / ms * 1e-9);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
