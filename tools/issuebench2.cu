// issuebench2.cu -- which instructions can issue in the shadow of a packed FP32x2 op?
//
// One CTA per SM, W warps per scheduler, every instruction in the timed loop is inline PTX in a fixed
// order:  [packed FFMA2, K x OTHER] x 64 per iteration, eight independent packed chains.  Time is
// read with clock64 inside the kernel, so the answer is in SM cycles whatever the clock does.
// Output: cycles per packed op per warp (1 warp/scheduler: latency-exposed; 2..4: throughput).
// Development tool; not part of the product.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { NONE, LOP3, IADD3, PRMT, LDS32, LDS64, LDS128, STS64, STS128, FFMA_R, FFMA_I, HFMA2_R, FADD_R, IMAD_R, LDSU16 };
static const char *kName[] = {"packed only", "LOP3", "IADD3", "PRMT", "LDS.32", "LDS.64", "LDS.128", "STS.64", "STS.128",
                              "FFMA reg", "FFMA imm", "HFMA2", "FADD", "IMAD", "LDS.U16"};

template <int KIND>
__device__ __forceinline__ void other(uint32_t &u, uint32_t &v, float &f, float &g, uint32_t sa, unsigned long long &sink)
{
    if (KIND == LOP3) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(u) : "r"(v));
    if (KIND == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(u) : "r"(v));
    if (KIND == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(u) : "r"(v));
    if (KIND == LDS32) asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(sa));
    if (KIND == LDSU16) asm volatile("ld.volatile.shared.u16 %0, [%1];" : "=r"(v) : "r"(sa));
    if (KIND == LDS64) { uint32_t t; asm volatile("ld.volatile.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "r"(sa)); }
    if (KIND == LDS128) { uint32_t t0, t1, t2; asm volatile("ld.volatile.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v), "=r"(t0), "=r"(t1), "=r"(t2) : "r"(sa)); }
    if (KIND == STS64) asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %1};" ::"r"(sa), "r"(u));
    if (KIND == STS128) asm volatile("st.volatile.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(sa), "r"(u));
    if (KIND == FFMA_R) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f) : "f"(g));
    if (KIND == FFMA_I) asm volatile("fma.rn.f32 %0, %0, %1, 0f3A83126F;" : "+f"(f) : "f"(g));
    if (KIND == FADD_R) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f) : "f"(g));
    if (KIND == HFMA2_R) asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(u) : "r"(v));
    if (KIND == IMAD_R) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(u) : "r"(v));
}

// PACKED: 0 = no packed op at all (OTHER alone), 1 = FFMA2, 2 = FADD2
template <int KIND, int K, int PACKED, int P>
__global__ void __launch_bounds__(512, 1) kern(float *out, long long *cycles, int iters, float a0)
{
    __shared__ __align__(16) uint32_t sbuf[512 * 4 + 64];
    unsigned long long x[8], m, c, sink = 0;
    uint32_t u[8], v[8];
    float f[8], g[8];
    for (int i = 0; i < 8; i++) {
        float lo = a0 + i, hi = a0 - i;
        asm("mov.b64 %0, {%1, %2};" : "=l"(x[i]) : "f"(lo), "f"(hi));
        u[i] = threadIdx.x * 77u + i; v[i] = threadIdx.x + 3u * i + 1u; f[i] = a0 * i; g[i] = 1e-3f * (i + 1);
    }
    { float ml = 1.0001f, mh = 0.9999f, cl = 1e-3f, ch = -1e-3f;
      asm("mov.b64 %0, {%1, %2};" : "=l"(m) : "f"(ml), "f"(mh)); asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(cl), "f"(ch)); }
    for (int i = threadIdx.x; i < 512 * 4 + 64; i += blockDim.x) sbuf[i] = i;
    __shared__ long long tmin, tmax;
    if (threadIdx.x == 0) { tmin = 0x7fffffffffffffffLL; tmax = 0; }
    __syncthreads();
    // conflict-free addresses for every access width: lane-contiguous
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sbuf);
    const uint32_t width = (KIND == LDS64 || KIND == STS64) ? 8 : (KIND == LDS128 || KIND == STS128) ? 16 : (KIND == LDSU16 ? 2 : 4);
    const uint32_t sa = base + (threadIdx.x & 31) * width + (threadIdx.x >> 5) * 512 % 4096;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 64 / P; r++) {
#pragma unroll
            for (int i = 0; i < P; i++) {
                if (PACKED == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[(r * P + i) & 7]) : "l"(m), "l"(c));
                if (PACKED == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(x[(r * P + i) & 7]) : "l"(c));
            }
#pragma unroll
            for (int q = 0; q < K; q++) other<KIND>(u[(r + 4 * q) & 7], v[(r + 4 * q) & 7], f[(r + 4 * q) & 7], g[(r + 4 * q) & 7], sa, sink);
        }
    }
    long long t1 = clock64();
    float s = 0; uint32_t us = 0;
    for (int i = 0; i < 8; i++) {
        float lo, hi;
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(x[i]));
        s += lo + hi + f[i]; us += u[i] ^ v[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(us);
    // whole-CTA time: first start to last finish over all warps (one warp alone may be favoured by the arbiter)
    if ((threadIdx.x & 31) == 0) { atomicMin((unsigned long long *)&tmin, (unsigned long long)t0); atomicMax((unsigned long long *)&tmax, (unsigned long long)t1); }
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = tmax - tmin;
}

template <int KIND, int K, int PACKED, int P = 1>
void run(float *out, long long *cyc_d)
{
    const int iters = 2000, grid = 148;
    for (int W = 1; W <= 4; W++) {
        kern<KIND, K, PACKED, P><<<grid, 128 * W>>>(out, cyc_d, iters, 1.0f);
        kern<KIND, K, PACKED, P><<<grid, 128 * W>>>(out, cyc_d, iters, 1.0f);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc_d, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < grid; i++) avg += (double)h[i];
        avg /= grid;
        // per scheduler: W warps x 64 slots per iteration
        // per scheduler and iteration: W warps x (64 packed + (64/P) K other)
        printf("%-9s %d packed : %d other  W=%d: %8.1f cycles per warp-iteration-slot-group; %6.3f cycles per packed op per scheduler\n", kName[KIND],
               PACKED ? P : 0, K, W, avg / iters / W, avg / (64.0 * iters * W));
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
}

int main()
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    printf("64 packed ops per warp-iteration in groups of P, each group followed by K OTHER ops; whole-CTA time, one CTA of 4W warps per SM\n");
    run<NONE, 0, 1>(out, cyc);
    run<NONE, 0, 2>(out, cyc);
#define ALL(KIND) run<KIND, 1, 0>(out, cyc); run<KIND, 1, 1>(out, cyc); run<KIND, 2, 1>(out, cyc); run<KIND, 1, 1, 4>(out, cyc); run<KIND, 1, 1, 8>(out, cyc);
    ALL(LOP3) ALL(PRMT) ALL(IADD3) ALL(IMAD_R) ALL(LDSU16) ALL(LDS64) ALL(LDS128) ALL(STS64) ALL(STS128)
    ALL(FFMA_R) ALL(FFMA_I) ALL(FADD_R) ALL(HFMA2_R)
    return 0;
}
