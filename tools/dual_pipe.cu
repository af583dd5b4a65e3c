// dual_pipe.cu — do DFMA (FP64 FMA pipe) and DMMA (FP64 tensor pipe) overlap on B200?  Even warps run a DMMA
// loop, odd warps a DFMA loop; compare with each loop alone.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void __launch_bounds__(512) k(double* out, int iters, int mode) {   // mode 0: all DMMA, 1: all DFMA, 2: alternate warps
    const int warp = threadIdx.x >> 5;
    const bool mma = mode == 0 || (mode == 2 && (warp & 1) == 0);
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; i++) { c[i][0] = i; c[i][1] = -i; }
    if (mma) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
        }
    } else {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) { c[i][0] = fma(c[i][0], a, b); c[i][1] = fma(c[i][1], b, a); }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 2 * 512 * sizeof(double));
    const int iters = 1 << 14;
    for (int mode = 0; mode < 3; mode++) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<<<p.multiProcessorCount * 2, 512>>>(out, iters, mode); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<p.multiProcessorCount * 2, 512>>>(out, iters, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)p.multiProcessorCount * 2 * 16;
        double tf_mma = 0, tf_fma = 0;
        if (mode == 0) tf_mma = warps * iters * 8 * 512.0 / ms * 1e-9;
        if (mode == 1) tf_fma = warps * 32 * iters * 16 * 2.0 / ms * 1e-9;
        if (mode == 2) { tf_mma = warps / 2 * iters * 8 * 512.0 / ms * 1e-9; tf_fma = warps / 2 * 32 * iters * 16 * 2.0 / ms * 1e-9; }
        printf("mode %d: %.3f ms  DMMA %.2f TFLOP/s  DFMA %.2f TFLOP/s  total %.2f\n", mode, ms, tf_mma, tf_fma, tf_mma + tf_fma);
    }
    return 0;
}
