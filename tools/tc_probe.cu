// tc_probe.cu -- development probe for the tensor-core pass-1 experiment (VERDICT r1 task 6).
//
// Checks, on one CTA, that a radix-16 DFT over exact 8-bit samples can be computed by tcgen05.mma:
//   A[row t][k = (n1, re|im)]  = fp16(byte - 127)            (K-major, no swizzle, 128 x 32)
//   B[n = (k1, re|im)][k]      = W16 matrix split hi + lo fp16 (K-major, no swizzle, 32 x 32 each)
//   D[t][n] (TMEM, fp32)       = A Bhi^T + A Blo^T            (4 MMAs of 128 x 32 x 16)
// and that tcgen05.ld 32x32b.x32 hands thread t its 16 complex outputs.  The descriptor convention
// (which of LBO/SBO is the K-direction stride) is a runtime flag so one run tests both.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tc_probe tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;                 // layout_type = 0 (no swizzle), base_offset = 0
}

__global__ void __launch_bounds__(128)
probe_kernel(const uint8_t *raw /* [16][128][2] */, const __half *Bhi, const __half *Blo, float *D /* [128][32] */,
             int swap, long long *cycles)
{
    __shared__ __align__(1024) uint8_t sA[128 * 64];
    __shared__ __align__(1024) uint8_t sB[2][32 * 64];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int t = threadIdx.x;

    // ---- A tile: row t, 16 samples -> 4 chunks of 16 B ----
    {
        uint32_t h[16];
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            const uint32_t r = *reinterpret_cast<const uint16_t *>(raw + (size_t)(n1 * 128 + t) * 2);
            // (1024 + I, 1024 + Q) as half2, then - 1151 -> (I - 127, Q - 127), exact
            const uint32_t v = __byte_perm(r, 0x64646464u, 0x7170);
            __half2 hv = *reinterpret_cast<const __half2 *>(&v);
            hv = __hadd2(hv, __floats2half2_rn(-1151.0f, -1151.0f));
            h[n1] = *reinterpret_cast<uint32_t *>(&hv);
        }
        const uint32_t base = smem_u32(sA) + (t >> 3) * 512 + (t & 7) * 16;
#pragma unroll
        for (int kc = 0; kc < 4; kc++)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + kc * 128), "r"(h[4 * kc]), "r"(h[4 * kc + 1]),
                         "r"(h[4 * kc + 2]), "r"(h[4 * kc + 3]) : "memory");
    }
    // ---- B tiles ----
    for (int i = t; i < 32 * 32; i += 128) {
        const int n = i >> 5, k = i & 31;
        const int off = ((n >> 3) * 4 + (k >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__half *>(sB[0] + off) = Bhi[i];
        *reinterpret_cast<__half *>(sB[1] + off) = Blo[i];
    }
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (t < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = tmem_base;

    const long long c0 = clock64();
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t kdir = 128, mdir = 512;   // byte strides of adjacent core matrices along K / along M,N
        const uint32_t lbo = swap ? mdir : kdir, sbo = swap ? kdir : mdir;
        int first = 1;
        for (int h = 0; h < 2; h++)
            for (int ks = 0; ks < 2; ks++) {
                const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * kdir, lbo, sbo);
                const uint64_t db = make_desc(smem_u32(sB[h]) + ks * 2 * kdir, lbo, sbo);
                const uint32_t acc = first ? 0u : 1u;
                first = 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        } while (!done);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[32];
    const uint32_t taddr = tm + ((uint32_t)((t >> 5) * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long c1 = clock64();
    if (t == 0) *cycles = c1 - c0;
#pragma unroll
    for (int i = 0; i < 32; i++) D[t * 32 + i] = __uint_as_float(v[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(32) : "memory");
}

int main()
{
    const double PI = 3.14159265358979323846;
    std::vector<uint8_t> raw(16 * 128 * 2);
    srand(7);
    for (auto &b : raw) b = (uint8_t)(rand() & 255);
    std::vector<__half> bhi(32 * 32), blo(32 * 32);
    std::vector<double> bex(32 * 32), bsplit(32 * 32);
    for (int k1 = 0; k1 < 16; k1++)
        for (int n1 = 0; n1 < 16; n1++) {
            const double th = 2 * PI * ((n1 * k1) % 16) / 16.0, c = cos(th), s = sin(th);
            // out_re = sum xr c + xi s ; out_im = sum xi c - xr s
            const double e[2][2] = {{c, s}, {-s, c}};
            for (int o = 0; o < 2; o++)
                for (int i = 0; i < 2; i++) {
                    const int idx = (2 * k1 + o) * 32 + (2 * n1 + i);
                    const __half hi = __float2half_rn((float)e[o][i]);
                    const __half lo = __float2half_rn((float)(e[o][i] - (double)__half2float(hi)));
                    bhi[idx] = hi; blo[idx] = lo; bex[idx] = e[o][i];
                    bsplit[idx] = (double)__half2float(hi) + (double)__half2float(lo);
                }
        }
    uint8_t *d_raw; __half *d_hi, *d_lo; float *d_D; long long *d_cyc;
    cudaMalloc(&d_raw, raw.size()); cudaMalloc(&d_hi, 2048); cudaMalloc(&d_lo, 2048); cudaMalloc(&d_D, 128 * 32 * 4);
    cudaMalloc(&d_cyc, 8);
    cudaMemcpy(d_raw, raw.data(), raw.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_hi, bhi.data(), 2048, cudaMemcpyHostToDevice);
    cudaMemcpy(d_lo, blo.data(), 2048, cudaMemcpyHostToDevice);
    for (int swap = 0; swap < 2; swap++) {
        cudaMemset(d_D, 0, 128 * 32 * 4);
        probe_kernel<<<1, 128>>>(d_raw, d_hi, d_lo, d_D, swap, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("swap=%d: CUDA error %s\n", swap, cudaGetErrorString(e)); return 1; }
        std::vector<float> D(128 * 32);
        long long cyc = 0;
        cudaMemcpy(D.data(), d_D, D.size() * 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        double emax = 0, esplit = 0, vmax = 0;
        for (int t = 0; t < 128; t++)
            for (int n = 0; n < 32; n++) {
                double ex = 0, sp = 0;
                for (int k = 0; k < 32; k++) {
                    const double x = (double)raw[(size_t)((k >> 1) * 128 + t) * 2 + (k & 1)] - 127.0;
                    ex += x * bex[n * 32 + k];
                    sp += x * bsplit[n * 32 + k];
                }
                emax = fmax(emax, fabs(D[t * 32 + n] - ex));
                esplit = fmax(esplit, fabs(D[t * 32 + n] - sp));
                vmax = fmax(vmax, fabs(ex));
            }
        printf("swap=%d (LBO=%s): max|D-exact|=%.3e  max|D-split|=%.3e  max|exact|=%.1f  rel=%.3e  cycles(mma+ld)=%lld  D[0][0..3]=%g %g %g %g\n",
               swap, swap ? "M-stride" : "K-stride", emax, esplit, vmax, emax / vmax, cyc, D[0], D[1], D[2], D[3]);
    }
    return 0;
}
