// subpass_bench.cu — how fast can the FP64 tensor pipe run the sweep kernel's sub-pass body?
// A CTA of 8 warps owns a 2^11-amplitude tile in shared memory and repeats: per warp, 32 vectors of 8
// amplitudes times one 8x8 complex matrix (A fragments fixed, B operands loaded from the tile, results
// stored back), barrier.  No global traffic: the number printed is the compute ceiling of a sweep.
// Variants differ in DMMA issue order / grouping only.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/subpass_bench tools/subpass_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct cplx { double x, y; };

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// slot masks of one synthetic sub-pass (matrix qubits at tile positions 4,5,6; conflict-free by the
// quarter-warp rule measured with tools/lds_pattern.cu)
__device__ __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 3) & 7u) ^ ((i >> 6) & 7u) ^ ((i >> 9) & 7u); }

struct FastSub { uint64_t vm0, vm1; uint32_t mat_off, stage, sr2, st0, gx1, gx2, simple, pad; };
struct FastWarp { uint64_t g; uint32_t s; uint32_t pad; };

template <int V>
__global__ void __launch_bounds__(256, 4) k(int iters, const cplx* mats, double* sink) {
    extern __shared__ __align__(16) unsigned char raw[];
    cplx* tile = reinterpret_cast<cplx*>(raw);
    cplx* pool = tile + 2048;
    FastSub* fast = reinterpret_cast<FastSub*>(pool + 264);
    FastWarp* fwarp = reinterpret_cast<FastWarp*>(fast + 4);
    uint32_t* flane = reinterpret_cast<uint32_t*>(fwarp + 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, q = lane >> 2, kk = lane & 3;
    for (int i = tid; i < 2048; i += 256) { tile[i].x = 1e-3 * i; tile[i].y = -1e-3 * i; }
    for (int i = tid; i < 264; i += 256) pool[i] = mats[i & 255];
    __syncthreads();
    // thread bits: 0->pos0, 1->pos2, 2->pos3, 3->pos1, 4->pos7, warp bits -> pos 8,9,10; matrix bits -> pos 4,5,6
    const uint32_t st0 = swz(1u << 0), st1 = swz(1u << 2), st2 = swz(1u << 3), st3 = swz(1u << 1), st4 = swz(1u << 7);
    const uint32_t sr0 = swz(1u << 4), sr1 = swz(1u << 5), sr2 = swz(1u << 6);
    const uint32_t swarp = swz((uint32_t)warp << 8);
    const uint32_t baseB = swarp ^ ((q & 1) ? st0 : 0u) ^ ((q & 2) ? st1 : 0u) ^ ((q & 4) ? st2 : 0u) ^ ((kk & 1) ? sr0 : 0u) ^ ((kk & 2) ? sr1 : 0u);
    const uint32_t baseC = swarp ^ ((kk & 1) ? st1 : 0u) ^ ((kk & 2) ? st2 : 0u) ^ ((q & 1) ? sr0 : 0u) ^ ((q & 2) ? sr1 : 0u) ^ ((q & 4) ? sr2 : 0u);
    if (V == 5) {
        if (tid < 4) { FastSub f; f.vm0 = 1ull << (12 + tid); f.vm1 = 0; f.mat_off = 0; f.stage = tid; f.sr2 = sr2; f.st0 = st0; f.gx1 = st3; f.gx2 = st4; f.simple = 1; f.pad = 0; fast[tid] = f; }
        if (tid < 32) { FastWarp w; w.g = (uint64_t)(tid & 7) << 20; w.s = swz((uint32_t)(tid & 7) << 8); w.pad = 0; fwarp[tid] = w; }
        if (tid < 128) {
            const int l = tid & 31, q2 = l >> 2, k2 = l & 3;
            const uint32_t b = ((q2 & 1) ? st0 : 0u) ^ ((q2 & 2) ? st1 : 0u) ^ ((q2 & 4) ? st2 : 0u) ^ ((k2 & 1) ? sr0 : 0u) ^ ((k2 & 2) ? sr1 : 0u);
            const uint32_t c = ((k2 & 1) ? st1 : 0u) ^ ((k2 & 2) ? st2 : 0u) ^ ((q2 & 1) ? sr0 : 0u) ^ ((q2 & 2) ? sr1 : 0u) ^ ((q2 & 4) ? sr2 : 0u);
            flane[tid] = b | (c << 16);
        }
        __syncthreads();
    }
    for (int it = 0; it < iters; ++it) {
        if (V == 5) {
            const int sidx = it & 3;
            const FastSub& f = fast[sidx];
            if (f.simple) {
                const uint32_t lt = flane[sidx * 32 + lane];
                const FastWarp fw = fwarp[sidx * 8 + warp];
                const uint32_t bB = fw.s ^ (lt & 0xffffu), bC = fw.s ^ (lt >> 16);
                const uint64_t gw = ((uint64_t)blockIdx.x << 11) | fw.g;
                const int off = ((int)f.stage == iters) ? 200 : (int)f.mat_off;
                const int var = ((gw & f.vm0) != 0 ? 1 : 0) | ((gw & f.vm1) != 0 ? 2 : 0);
                const cplx* M = pool + off + var * 65;
                const cplx m0 = M[lane], m1 = M[32 + lane];
                const double nm0y = -m0.y, nm1y = -m1.y;
                const uint32_t gx1 = f.gx1, gx2 = f.gx2, xr2 = f.sr2, xt0 = f.st0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    cplx v0[2], v1[2];
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                        v0[g] = tile[bB ^ gx]; v1[g] = tile[bB ^ gx ^ xr2];
                    }
                    __syncwarp();
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        const uint32_t gx = (g ? gx1 : 0u) ^ (h ? gx2 : 0u);
                        double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
                        dmma(cr0, cr1, m0.x, v0[g].x); dmma(ci0, ci1, m0.x, v0[g].y);
                        dmma(cr0, cr1, m1.x, v1[g].x); dmma(ci0, ci1, m1.x, v1[g].y);
                        dmma(cr0, cr1, nm0y, v0[g].y); dmma(ci0, ci1, m0.y, v0[g].x);
                        dmma(cr0, cr1, nm1y, v1[g].y); dmma(ci0, ci1, m1.y, v1[g].x);
                        cplx o0, o1; o0.x = cr0; o0.y = ci0; o1.x = cr1; o1.y = ci1;
                        tile[bC ^ gx] = o0; tile[bC ^ gx ^ xt0] = o1;
                    }
                }
            }
            __syncthreads();
            continue;
        }
        const cplx* M = pool + (it & 3) * 64;
        const cplx m0 = M[lane], m1 = M[32 + lane];
        const double nm0y = -m0.y, nm1y = -m1.y;
        if (V == 0 || V == 1 || V == 2) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                cplx v0[2], v1[2];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t gx = (g ? st3 : 0u) ^ (h ? st4 : 0u);
                    v0[g] = tile[baseB ^ gx]; v1[g] = tile[baseB ^ gx ^ sr2];
                }
                __syncwarp();
                double cr[2][2] = {{0, 0}, {0, 0}}, ci[2][2] = {{0, 0}, {0, 0}};
                if (V == 0) {            // group-major, re/im interleaved (what the sweep kernel does)
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        dmma(cr[g][0], cr[g][1], m0.x, v0[g].x); dmma(ci[g][0], ci[g][1], m0.x, v0[g].y);
                        dmma(cr[g][0], cr[g][1], m1.x, v1[g].x); dmma(ci[g][0], ci[g][1], m1.x, v1[g].y);
                        dmma(cr[g][0], cr[g][1], nm0y, v0[g].y); dmma(ci[g][0], ci[g][1], m0.y, v0[g].x);
                        dmma(cr[g][0], cr[g][1], nm1y, v1[g].y); dmma(ci[g][0], ci[g][1], m1.y, v1[g].x);
                    }
                } else if (V == 1) {     // A-major: four independent accumulators between dependent DMMAs
#pragma unroll
                    for (int g = 0; g < 2; ++g) { dmma(cr[g][0], cr[g][1], m0.x, v0[g].x); dmma(ci[g][0], ci[g][1], m0.x, v0[g].y); }
#pragma unroll
                    for (int g = 0; g < 2; ++g) { dmma(cr[g][0], cr[g][1], m1.x, v1[g].x); dmma(ci[g][0], ci[g][1], m1.x, v1[g].y); }
#pragma unroll
                    for (int g = 0; g < 2; ++g) { dmma(cr[g][0], cr[g][1], nm0y, v0[g].y); dmma(ci[g][0], ci[g][1], m0.y, v0[g].x); }
#pragma unroll
                    for (int g = 0; g < 2; ++g) { dmma(cr[g][0], cr[g][1], nm1y, v1[g].y); dmma(ci[g][0], ci[g][1], m1.y, v1[g].x); }
                } else {                 // chain-major: a whole dependent chain at a time
#pragma unroll
                    for (int g = 0; g < 2; ++g) {
                        dmma(cr[g][0], cr[g][1], m0.x, v0[g].x); dmma(cr[g][0], cr[g][1], m1.x, v1[g].x);
                        dmma(cr[g][0], cr[g][1], nm0y, v0[g].y); dmma(cr[g][0], cr[g][1], nm1y, v1[g].y);
                        dmma(ci[g][0], ci[g][1], m0.x, v0[g].y); dmma(ci[g][0], ci[g][1], m1.x, v1[g].y);
                        dmma(ci[g][0], ci[g][1], m0.y, v0[g].x); dmma(ci[g][0], ci[g][1], m1.y, v1[g].x);
                    }
                }
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const uint32_t gx = (g ? st3 : 0u) ^ (h ? st4 : 0u);
                    cplx o0, o1; o0.x = cr[g][0]; o0.y = ci[g][0]; o1.x = cr[g][1]; o1.y = ci[g][1];
                    tile[baseC ^ gx] = o0; tile[baseC ^ gx ^ st0] = o1;
                }
            }
        } else if (V == 3) {             // all four groups in flight, A-major
            cplx v0[4], v1[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t gx = ((g & 1) ? st3 : 0u) ^ ((g & 2) ? st4 : 0u);
                v0[g] = tile[baseB ^ gx]; v1[g] = tile[baseB ^ gx ^ sr2];
            }
            __syncwarp();
            double cr[4][2], ci[4][2];
#pragma unroll
            for (int g = 0; g < 4; ++g) { cr[g][0] = cr[g][1] = ci[g][0] = ci[g][1] = 0.0; }
#pragma unroll
            for (int g = 0; g < 4; ++g) { dmma(cr[g][0], cr[g][1], m0.x, v0[g].x); dmma(ci[g][0], ci[g][1], m0.x, v0[g].y); }
#pragma unroll
            for (int g = 0; g < 4; ++g) { dmma(cr[g][0], cr[g][1], m1.x, v1[g].x); dmma(ci[g][0], ci[g][1], m1.x, v1[g].y); }
#pragma unroll
            for (int g = 0; g < 4; ++g) { dmma(cr[g][0], cr[g][1], nm0y, v0[g].y); dmma(ci[g][0], ci[g][1], m0.y, v0[g].x); }
#pragma unroll
            for (int g = 0; g < 4; ++g) { dmma(cr[g][0], cr[g][1], nm1y, v1[g].y); dmma(ci[g][0], ci[g][1], m1.y, v1[g].x); }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t gx = ((g & 1) ? st3 : 0u) ^ ((g & 2) ? st4 : 0u);
                cplx o0, o1; o0.x = cr[g][0]; o0.y = ci[g][0]; o1.x = cr[g][1]; o1.y = ci[g][1];
                tile[baseC ^ gx] = o0; tile[baseC ^ gx ^ st0] = o1;
            }
        }
        __syncthreads();
    }
    if (tile[tid].x == 1.2345) sink[0] = tile[tid].y;
}

template <int V>
static void run(const char* name, int sms, const cplx* mats, double* sink) {
    const int iters = 3000;
    const size_t smem = (2048 + 264) * sizeof(cplx) + 4 * sizeof(FastSub) + 32 * sizeof(FastWarp) + 128 * 4;
    cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<V><<<sms * 4, 256, smem>>>(iters, mats, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double flops = (double)sms * 4 * iters * 8 * 32 * 512.0;      // 8 warps x 32 DMMA x 512 flop
    printf("variant %d (%s): %.3f ms  %.2f TFLOP/s  %s\n", V, name, best, flops / best * 1e-9, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    cplx* mats; double* sink;
    cudaMalloc(&mats, 256 * sizeof(cplx)); cudaMalloc(&sink, 8);
    cplx h[256];
    for (int i = 0; i < 256; ++i) { h[i].x = (i % 9 == 0) ? 0.7 : 0.01; h[i].y = (i % 7 == 0) ? -0.7 : 0.02; }
    cudaMemcpy(mats, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>("group-major interleaved", p.multiProcessorCount, mats, sink);
    run<1>("A-major, 2 groups", p.multiProcessorCount, mats, sink);
    run<2>("chain-major", p.multiProcessorCount, mats, sink);
    run<3>("A-major, 4 groups", p.multiProcessorCount, mats, sink);
    run<5>("table-driven like the sweep kernel", p.multiProcessorCount, mats, sink);
    return 0;
}
