// issuebench.cu -- does a packed FP32x2 op cost one issue slot or two?  Interleave K "other" ops per
// packed op and see whether the time follows max(2*P, P+K*P) [other ops issue in the shadow] or 2*P + K*P.
#include <cstdio>
#include <cuda_runtime.h>
template <int KIND, int K>   // KIND 0: none, 1: LOP3/IADD (alu), 2: PRMT, 3: LDS.64 (conflict-free), 4: scalar FFMA
__global__ void __launch_bounds__(256) k(float2 *out, int iters, float2 a0, unsigned seed)
{
    __shared__ float2 sbuf[1024];
    float2 x[8];
    unsigned u[8];
    float f[8];
    for (int i = 0; i < 8; i++) { x[i] = make_float2(a0.x + i, a0.y - i); u[i] = seed + i * 77u + threadIdx.x; f[i] = a0.x * i; }
    sbuf[threadIdx.x] = a0; sbuf[threadIdx.x + 256] = a0; sbuf[threadIdx.x + 512] = a0; sbuf[threadIdx.x + 768] = a0;
    __syncthreads();
    const float2 m = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                x[i] = __ffma2_rn(x[i], m, c);
#pragma unroll
                for (int q = 0; q < K; q++) {
                    if (KIND == 1) u[i] = (u[i] ^ (u[(i + 1) & 7] + 0x9e37u));
                    if (KIND == 2) u[i] = __byte_perm(u[i], u[(i + 3) & 7], 0x7604);
                    if (KIND == 3) { float2 t = sbuf[(threadIdx.x + 32 * (i + 8 * q + r)) & 1023]; f[i] += t.x; }
                    if (KIND == 4) f[i] = fmaf(f[i], 1.0001f, 1e-3f);
                }
            }
        }
    }
    float s = 0; unsigned us = 0;
    for (int i = 0; i < 8; i++) { s += x[i].x + x[i].y + f[i]; us += u[i]; }
    out[blockIdx.x * 256 + threadIdx.x] = make_float2(s, __uint_as_float(us));
}
template <int KIND, int K> void run(const char *name, float2 *out)
{
    const int iters = 2000, grid = 148 * 2;   // 16 warps / SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<KIND, K><<<grid, 256>>>(out, iters, make_float2(1.f, 2.f), 5u);
    cudaEventRecord(e0);
    k<KIND, K><<<grid, 256>>>(out, iters, make_float2(1.f, 2.f), 5u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double packed_per_sched = 64.0 * iters * 4;      // 4 warps per scheduler x 64 packed per warp-iteration
    double cyc = ms * 1e-3 * 1.965e9;
    printf("%-22s K=%d: %.3f ms  %.2f cycles per packed op per scheduler\n", name, K, ms, cyc / packed_per_sched);
}
int main()
{
    float2 *out; cudaMalloc(&out, 148 * 2 * 256 * 8);
    run<0, 0>("packed only", out);
    run<1, 1>("+ LOP3/IADD3", out); run<1, 2>("+ LOP3/IADD3", out); run<1, 3>("+ LOP3/IADD3", out);
    run<2, 1>("+ PRMT", out); run<2, 2>("+ PRMT", out);
    run<3, 1>("+ LDS.64", out); run<3, 2>("+ LDS.64", out);
    run<4, 1>("+ scalar FFMA", out); run<4, 2>("+ scalar FFMA", out);
    return 0;
}
