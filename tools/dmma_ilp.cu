// dmma_ilp.cu — DMMA.8x8x4 throughput vs independent accumulators per warp and warps per SM (B200 exploration).
#include <cuda_runtime.h>
#include <cstdio>
template <int NACC>
__global__ void __launch_bounds__(1024) k(double* out, int iters) {
    double a[NACC], b[NACC], c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { a[i] = 1.0 + (threadIdx.x + i) * 1e-9; b[i] = 1.0 - (threadIdx.x + 3 * i) * 1e-9; c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16 / NACC; r++)
#pragma unroll
            for (int i = 0; i < NACC; i++)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[i]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
void run(double* out, int sms) {
    for (int warps : {4, 8, 16, 32}) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 4096;
        k<NACC><<<sms, warps * 32>>>(out, iters); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<NACC><<<sms, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("nacc=%d warps/SM=%d  %.2f TFLOP/s\n", NACC, warps, (double)sms * warps * iters * 16 * 512.0 / ms * 1e-9);
    }
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, (size_t)p.multiProcessorCount * 1024 * sizeof(double));
    run<1>(out, p.multiProcessorCount); run<2>(out, p.multiProcessorCount); run<4>(out, p.multiProcessorCount);
    run<8>(out, p.multiProcessorCount); run<16>(out, p.multiProcessorCount);
    return 0;
}
