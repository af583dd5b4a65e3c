// peaks.cu — micro-benchmarks for the roofline denominators MEASURED_PEAKS.json does not carry:
//   * FP64 tensor pipe (DMMA.8x8x4 via mma.sync m8n8k4 f64) sustained throughput
//   * FP64 FMA pipe (DFMA) throughput
//   * HBM copy / read-modify-write bandwidth with 128-bit accesses (the sweep kernel's access width)
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_build/peaks tools/peaks.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void __launch_bounds__(1024) dmma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// same loop with a distinct (a, b) register pair per accumulator chain, as a GEMM inner loop has
__global__ void __launch_bounds__(1024) dmma_kernel_distinct(double* out, int iters) {
    double a[16], b[16], c[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { a[i] = 1.0 + (threadIdx.x + i) * 1e-9; b[i] = 1.0 - (threadIdx.x + 3 * i) * 1e-9; c[i][0] = 0; c[i][1] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i]), "d"(b[i]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// NACC independent accumulator chains, each with its own operand registers (tools/dmma_ilp.cu): the shape that
// reaches the pipe's peak on B200 from 8 warps per SM on
template <int NACC>
__global__ void __launch_bounds__(1024) dmma_kernel_ilp(double* out, int iters) {
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

__global__ void __launch_bounds__(1024) dfma_kernel(double* out, int iters) {
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__global__ void rmw_kernel(double2* __restrict__ buf, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 v = buf[i];
        v.x = v.x * 0.999 - v.y * 0.01; v.y = v.y * 0.999 + v.x * 0.01;
        buf[i] = v;
    }
}

template <typename F>
static float best_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaGetLastError()); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, (size_t)sms * 8 * 1024 * sizeof(double)));
    double dmma_best = 0, dfma_best = 0, dmma2_best = 0, dmma_ilp_best = 0; int dmma_cfg = 0, dfma_cfg = 0, ilp_cfg = 0;
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int iters = 4096;
        for (int ctas = 1; ctas <= 2; ctas++) {
            if (warps * ctas > 32) continue;
            float ms = best_ms([&] { dmma_kernel<<<sms * ctas, warps * 32>>>(out, iters); }, 5);
            double fl = (double)sms * ctas * warps * iters * 16 * 512.0;   // m8n8k4 = 2*8*8*4 flop
            double tf = fl / ms * 1e-9;
            if (tf > dmma_best) { dmma_best = tf; dmma_cfg = warps * 100 + ctas; }
            ms = best_ms([&] { dmma_kernel_distinct<<<sms * ctas, warps * 32>>>(out, iters); }, 5);
            tf = (double)sms * ctas * warps * iters * 16 * 512.0 / ms * 1e-9;
            if (tf > dmma2_best) dmma2_best = tf;
            ms = best_ms([&] { dmma_kernel_ilp<4><<<sms * ctas, warps * 32>>>(out, iters); }, 5);
            tf = (double)sms * ctas * warps * iters * 16 * 512.0 / ms * 1e-9;
            if (tf > dmma_ilp_best) { dmma_ilp_best = tf; ilp_cfg = warps * 100 + ctas; }
            ms = best_ms([&] { dfma_kernel<<<sms * ctas, warps * 32>>>(out, iters); }, 5);
            fl = (double)sms * ctas * warps * 32 * iters * 16 * 2.0;
            tf = fl / ms * 1e-9;
            if (tf > dfma_best) { dfma_best = tf; dfma_cfg = warps * 100 + ctas; }
        }
    }
    // sustained DMMA over ~2 s
    double dmma_sus = 0;
    {
        const int warps = ilp_cfg / 100, ctas = ilp_cfg % 100, iters = 1 << 16;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0; float ms = 0;
        do {
            dmma_kernel_ilp<4><<<sms * ctas, warps * 32>>>(out, iters); n++;
            CK(cudaGetLastError());
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        } while (ms < 2000.f);
        dmma_sus = (double)n * sms * ctas * warps * iters * 16 * 512.0 / ms * 1e-9;
    }
    const size_t n = (size_t)1 << 28;   // 4 GiB per buffer
    double2 *a, *b; CK(cudaMalloc(&a, n * sizeof(double2))); CK(cudaMalloc(&b, n * sizeof(double2)));
    CK(cudaMemset(a, 1, n * sizeof(double2))); CK(cudaMemset(b, 0, n * sizeof(double2)));
    float ms_copy = best_ms([&] { copy_kernel<<<sms * 16, 256>>>(a, b, n); }, 10);
    float ms_rmw = best_ms([&] { rmw_kernel<<<sms * 16, 256>>>(a, n); }, 10);
    float ms_memcpy = best_ms([&] { cudaMemcpyAsync(b, a, n * sizeof(double2), cudaMemcpyDeviceToDevice); }, 10);
    const double dmma_peak = dmma_ilp_best > dmma2_best ? (dmma_ilp_best > dmma_best ? dmma_ilp_best : dmma_best) : (dmma2_best > dmma_best ? dmma2_best : dmma_best);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dmma_tflops\": %.2f, \"dmma_tflops_sustained\": %.2f, \"dmma_same_operand_tflops\": %.2f, \"dmma_distinct_operands_tflops\": %.2f, \"dmma_ilp4_tflops\": %.2f, \"dmma_cfg_warps_ctas\": %d, "
           "\"dfma_tflops\": %.2f, \"dfma_cfg_warps_ctas\": %d, \"copy128_gbs\": %.1f, \"rmw128_gbs\": %.1f, \"memcpy_d2d_gbs\": %.1f}\n",
           p.name, sms, dmma_peak, dmma_sus, dmma_best, dmma2_best, dmma_ilp_best, ilp_cfg, dfma_best, dfma_cfg,
           2.0 * n * 16 / ms_copy * 1e-6, 2.0 * n * 16 / ms_rmw * 1e-6, 2.0 * n * 16 / ms_memcpy * 1e-6);
    return 0;
}
