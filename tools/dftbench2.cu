// dftbench2.cu -- development microbenchmark: one "iteration" of the 4096 kernel's register work
// (unpack 16 samples, 3 x radix-16, 2 x 15 twiddles, power) with optional extras, no exchange.
//   MODE 0: FP only (inputs recycled)        MODE 1: + 16 LDS.U16 + 32 PRMT unpack per iteration
//   MODE 2: MODE 1 + 6 mbarrier try_wait on an already-completed barrier per 2 iterations (3 per iteration)
//   MODE 3: MODE 1 + __syncthreads x2 per iteration
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_math.cuh"
#include "stage_prims.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k(const float2 *in, float2 *out, int iters, const float2 *tw)
{
    __shared__ __align__(16) uint8_t rawbuf[8192];
    __shared__ uint64_t bar;
    c32 v[16];
    float2 w1[15], w2[15];
    for (int i = 0; i < 16; i++) v[i] = in[threadIdx.x + 256 * i];
    for (int i = 0; i < 15; i++) { w1[i] = tw[threadIdx.x + 256 * i]; w2[i] = tw[threadIdx.x + 256 * i + 1]; }
    for (int i = threadIdx.x; i < 8192; i += 256) rawbuf[i] = (uint8_t)(i * 7 + 3);
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); }
    __syncthreads();
    if (threadIdx.x == 0) mbar_arrive(smem_u32(&bar));     // phase 0 complete forever after
    __syncthreads();
    float acc[16] = {0};
    const uint32_t rb = smem_u32(rawbuf) + 2 * threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (MODE >= 1) {
#pragma unroll
            for (int n = 0; n < 16; n++) {
                uint32_t r;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(rb + n * 512));
                v[n] = unpack_biased(r + (uint32_t)it);
            }
        }
        if (MODE == 2) mbar_wait(smem_u32(&bar), 0);
        Dft<16>::run(v);
        if (MODE >= 1) v[0] = __fadd2_rn(v[0], make_float2(-16.0f * kBias, -16.0f * kBias));
#pragma unroll
        for (int i = 1; i < 16; i++) v[i] = cmul(v[i], w1[i - 1]);
        if (MODE == 3) __syncthreads();
        if (MODE == 2) mbar_wait(smem_u32(&bar), 0);
        Dft<16>::run(v);
#pragma unroll
        for (int i = 1; i < 16; i++) v[i] = cmul(v[i], w2[i - 1]);
        if (MODE == 3) __syncthreads();
        if (MODE == 2) mbar_wait(smem_u32(&bar), 0);
        Dft<16>::run(v);
#pragma unroll
        for (int i = 0; i < 16; i++) acc[i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, acc[i]));
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) { v[i].x *= 2.4e-4f; v[i].y *= 2.4e-4f; }
        }
    }
    for (int i = 0; i < 16; i++) out[threadIdx.x + 256 * i + blockIdx.x * 4096] = make_float2(v[i].x + acc[i], v[i].y);
}

template <int MODE> void run(int blocksPerSM, int iters, float2 *in, float2 *out, float2 *tw)
{
    int grid = 148 * blocksPerSM;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(in, out, iters, tw);
    cudaEventRecord(e0);
    k<MODE><<<grid, 256>>>(in, out, iters, tw);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double laneops = 3 * 160 + 60 + 32 + (MODE == 0 ? 32 : 2);
    double total = laneops * 256.0 * grid * iters;
    double peak = 148.0 * 128 * 1.965e9;
    double cyc = ms * 1e-3 * 1.965e9 / iters / blocksPerSM;      // SM cycles per CTA-iteration (4096 samples)
    printf("mode %d warps/SM %2d: %.3f ms  FP32 %.1f%%  %.0f cycles per spectrum-iteration (FMA floor 1268)\n", MODE,
           blocksPerSM * 8, ms, 100.0 * total / (ms * 1e-3) / peak, cyc);
}

int main()
{
    float2 *in, *out, *tw;
    cudaMalloc(&in, 4096 * 8); cudaMalloc(&tw, 4100 * 8); cudaMalloc(&out, 4096 * 8 * 148 * 8);
    cudaMemset(in, 0, 4096 * 8); cudaMemset(tw, 0, 4100 * 8);
    for (int b : {1, 2}) { run<0>(b, 1000, in, out, tw); run<1>(b, 1000, in, out, tw); run<2>(b, 1000, in, out, tw); run<3>(b, 1000, in, out, tw); }
    return 0;
}
