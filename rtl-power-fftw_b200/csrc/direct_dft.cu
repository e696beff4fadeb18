// direct_dft.cu -- even N that is NOT a power of two (the reference only requires N even,
// src/params.cxx:152-155; its man page uses `-b 500`).  FFTW handles such N with mixed-radix /
// Rader / Bluestein plans; here the fused path is an O(N^2) DFT per spectrum, which is plenty for
// these sizes (N <= 8190): at N = 500 one CTA needs 250 k complex MACs per spectrum and the GPU
// sustains > 10^3 x the dongle's sample rate.  Same contract as the FFT kernels:
//   x[n] = (u8 - 127) (* w[n]),  X[k] = sum_n x[n] W_N^{nk},  pwr[(k + N/2) mod N] += |X[k]|^2.
// One CTA per spectrum at a time; x and the twiddle table W_N^j live in shared memory; thread `tid`
// owns output bins tid, tid + 256, ...; the twiddle index n*k mod N is advanced incrementally in
// integers, so no precision is lost in the phase.  fp32 accumulation over N terms: ~sqrt(N) ulp.
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"

namespace {

constexpr int kDdBlock = 256;
constexpr int kDdMaxBinsPerThread = 32;     // N <= 8192
constexpr int kDdFlush = 64;

__global__ void __launch_bounds__(kDdBlock)
direct_dft_kernel(int N, const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
                  const float *__restrict__ window, double *__restrict__ rows)
{
    extern __shared__ __align__(16) uint8_t dsm[];
    c32 *xs = reinterpret_cast<c32 *>(dsm);          // [N]
    c32 *ws = xs + N;                                // [N]  W_N^j
    const int tid = threadIdx.x;
    const int64_t s0 = (nspec * (int64_t)blockIdx.x) / gridDim.x;
    const int64_t s1 = (nspec * (int64_t)(blockIdx.x + 1)) / gridDim.x;
    if (s0 >= s1) return;
    for (int j = tid; j < N; j += kDdBlock) ws[j] = twiddle[j];
    double *row = rows + (size_t)blockIdx.x * N;
    float acc[kDdMaxBinsPerThread];
#pragma unroll
    for (int b = 0; b < kDdMaxBinsPerThread; b++) acc[b] = 0.0f;
    const int nb = (N - tid + kDdBlock - 1) / kDdBlock;     // bins owned by this thread
    int since = 0;
    auto flush = [&]() {
#pragma unroll
        for (int b = 0; b < kDdMaxBinsPerThread; b++) {
            if (b < nb) {
                int out = tid + b * kDdBlock + N / 2;           // fftshift == (-1)^n of src/datastore.cxx:73
                if (out >= N) out -= N;
                row[out] += (double)acc[b];
                acc[b] = 0.0f;
            }
        }
    };
    for (int64_t s = s0; s < s1; s++) {
        const uint8_t *raw = bytes + (size_t)s * 2 * N;
        __syncthreads();                                        // previous spectrum fully consumed
        for (int n = tid; n < N; n += kDdBlock) {
            const uint32_t r = *reinterpret_cast<const uint16_t *>(raw + 2 * n);
            float re = (float)(r & 255u) - 127.0f, im = (float)(r >> 8) - 127.0f;   // src/datastore.cxx:74-75
            if (window) { const float w = window[n]; re *= w; im *= w; }            // :76-77
            xs[n] = make_float2(re, im);
        }
        __syncthreads();
#pragma unroll 1
        for (int b = 0; b < nb; b++) {
            const int k = tid + b * kDdBlock;
            float sr = 0.0f, si = 0.0f;
            int idx = 0;                                         // n*k mod N
            for (int n = 0; n < N; n++) {
                const c32 x = xs[n], w = ws[idx];
                sr = fmaf(x.x, w.x, fmaf(-x.y, w.y, sr));
                si = fmaf(x.x, w.y, fmaf(x.y, w.x, si));
                idx += k; if (idx >= N) idx -= N;
            }
            // acc[] is indexed with a compile-time-unrollable switch to stay in registers
#pragma unroll
            for (int q = 0; q < kDdMaxBinsPerThread; q++)
                if (q == b) acc[q] = fmaf(sr, sr, fmaf(si, si, acc[q]));
        }
        if (++since == kDdFlush) { flush(); since = 0; }
    }
    flush();
}

}  // namespace

size_t rpf_direct_dft_smem(int N) { return (size_t)N * 2 * sizeof(float2); }

cudaError_t rpf_direct_dft_configure()
{
    return cudaFuncSetAttribute(direct_dft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rpf_direct_dft_smem(8192));
}

cudaError_t rpf_launch_direct_dft(int N, const RpfLaunch &L)
{
    if (L.nspec <= 0) return cudaSuccess;
    const int grid = (int)(L.nspec < L.grid ? L.nspec : L.grid);
    direct_dft_kernel<<<grid, kDdBlock, rpf_direct_dft_smem(N), L.stream>>>(N, L.bytes, L.nspec, L.twiddle, L.window, L.rows);
    return cudaGetLastError();
}
