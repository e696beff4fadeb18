// dftbench.cu -- development microbenchmark: how close can the radix-16 butterfly + twiddle
// instruction mix get to the FP32 pipe peak, as a function of resident warps?  (No memory.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../rtl-power-fftw_b200/csrc dftbench.cu -o dftbench
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_math.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k(const float2 *in, float2 *out, int iters, const float2 *tw)
{
    c32 v[16];
    float2 w[15];
    for (int i = 0; i < 16; i++) v[i] = in[threadIdx.x + 256 * i];
    for (int i = 0; i < 15; i++) w[i] = tw[threadIdx.x + 256 * i];
    float acc[16] = {0};
    for (int it = 0; it < iters; it++) {
        Dft<16>::run(v);
        if (MODE >= 1) {
#pragma unroll
            for (int i = 1; i < 16; i++) v[i] = cmul(v[i], w[i - 1]);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, acc[i]));
        }
#pragma unroll
        for (int i = 0; i < 16; i++) { v[i].x *= 0.0625f; v[i].y *= 0.0625f; }   // keep values bounded (counted below)
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
    // lane-ops per thread-iteration: DFT16 = 80 packed (160), cmul 15*2 packed (60), power 32, scale 16 packed?(FMUL x2 = 32 scalar or 16 packed)
    double laneops = 160 + (MODE >= 1 ? 60 : 0) + (MODE >= 2 ? 32 : 0) + 32;
    double total = laneops * 256.0 * grid * iters;
    double peak = 148.0 * 128 * 1.965e9;
    printf("mode %d blocks/SM %d (warps/SM %2d): %.3f ms  %.1f%% of FP32 lane peak (at 1.965 GHz)\n", MODE, blocksPerSM,
           blocksPerSM * 8, ms, 100.0 * total / (ms * 1e-3) / peak);
}

int main()
{
    float2 *in, *out, *tw;
    cudaMalloc(&in, 4096 * 8); cudaMalloc(&tw, 4096 * 8); cudaMalloc(&out, 4096 * 8 * 148 * 8);
    cudaMemset(in, 0, 4096 * 8); cudaMemset(tw, 0, 4096 * 8);
    for (int b : {1, 2, 3, 4}) { run<0>(b, 2000, in, out, tw); run<1>(b, 2000, in, out, tw); run<2>(b, 2000, in, out, tw); }
    return 0;
}
