// lds_pattern.cu — which lane -> 16-byte-slot patterns make LDS.128 / STS.128 conflict-free on sm_100a?
// Lane bit b contributes XOR mask m[b] to the slot index (the sweep kernel's addressing is exactly of
// this form).  Each mask's low 3 bits pick the 16-byte bank group.  For every assignment of bank-group
// bits to the 5 lane bits the kernel times a long stream of 128-bit shared loads (or stores).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/lds_pattern tools/lds_pattern.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Masks { uint32_t m[5]; };

template <bool STORE>
__global__ void __launch_bounds__(256) k(Masks mk, int iters, double* out) {
    __shared__ double2 tile[2048];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = make_double2(i, -i);
    __syncthreads();
    uint32_t slot = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) if ((lane >> b) & 1) slot ^= mk.m[b];
    slot ^= warp << 8;
    const unsigned base = (unsigned)__cvta_generic_to_shared(tile);
    double acc = 0.0;
    for (int it = 0; it < iters; ++it) {
        const unsigned a = base + ((slot ^ ((it & 7) << 3) ^ ((it & 1) << 7)) << 4);
        if (STORE) {
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" :: "r"(a), "d"(acc), "d"(acc) : "memory");
        } else {
            double x, y;
            asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) : "memory");
            acc += x + y;
        }
    }
    if (acc == 12345.678) out[0] = acc;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, 8);
    // candidate masks per bank-group bit (distinct index bits so slots stay distinct)
    const uint32_t cand[3][4] = {{1u, 8u ^ 1u, 64u ^ 1u, 0}, {2u, 16u ^ 2u, 128u ^ 2u, 0}, {4u, 32u ^ 4u, 0, 0}};
    const int ncand[3] = {3, 3, 2};
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int store = 0; store < 2; ++store) {
        printf("# %s.128: assignment of bank-group bit to lane bits 0..4, ms, relative to best\n", store ? "STS" : "LDS");
        float best = 1e30f; float res[243]; int valid[243];
        for (int code = 0; code < 243; ++code) {
            int asg[5], c = code, used[3] = {0, 0, 0}; bool ok = true; Masks mk;
            for (int b = 0; b < 5; ++b) { asg[b] = c % 3; c /= 3; }
            for (int b = 0; b < 5; ++b) { if (used[asg[b]] >= ncand[asg[b]]) { ok = false; break; } mk.m[b] = cand[asg[b]][used[asg[b]]++]; }
            valid[code] = ok; if (!ok) continue;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (store) k<true><<<p.multiProcessorCount * 4, 256>>>(mk, iters, d); else k<false><<<p.multiProcessorCount * 4, 256>>>(mk, iters, d);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            cudaEventElapsedTime(&res[code], e0, e1);
            if (res[code] < best) best = res[code];
        }
        for (int code = 0; code < 243; ++code) {
            if (!valid[code]) continue;
            int c = code; printf("%s ", store ? "STS" : "LDS");
            for (int b = 0; b < 5; ++b) { printf("%d", c % 3); c /= 3; }
            printf("  %.3f ms  x%.2f\n", res[code], res[code] / best);
        }
    }
    return 0;
}
