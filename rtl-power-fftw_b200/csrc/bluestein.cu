// bluestein.cu -- every even N the fused kernels do not cover, in O(N log N):
//   * N even, not a power of two (the reference only forces N even, src/params.cxx:152-155; FFTW handles such
//     N with mixed-radix / Rader / Bluestein plans): chirp-z (Bluestein) over M = 2^m >= 2N - 1,
//         a[n] = x[n] c[n],  c[n] = exp(-i pi n^2 / N)  (n^2 reduced mod 2N in integers: no phase error),
//         A = FFT_M(a),  C = conj(A . Bhat),  y = FFT_M(C),  |X[k]|^2 = |y[k]|^2   (k < N)
//     with Bhat = FFT_M(conj chirp, wrapped) / M tabulated in double on the host.  The inverse transform is a
//     forward one on the conjugate, and the final chirp multiplication drops out of |.|^2 (|c| = 1).
//   * N = 2^m > 65536: one plain forward FFT_N.
// The FFT is a batched global-memory Stockham autosort, radix 4 (one radix-2 pass when log2 M is odd); every pass
// streams the batch once.  Correctness path for sizes the reference's CLI allows -- ~10 Gsamples/s, three orders of
// magnitude above the dongle's rate -- not a tuned kernel.  Same contract as the fused kernels:
//   x[n] = (u8 - 127) (* w[n]),  pwr[(k + N/2) mod N] += |X[k]|^2, fp32 partial sums over <= 128 spectra, then float64.
// Replaces reference src/datastore.cxx:68-87 for those N.
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"

namespace {

constexpr int kBlock = 256;

// bytes -> a[b][n] = (x[n] - 127(1+j)) (w[n]) c[n], zero-padded to M            (src/datastore.cxx:73-77)
__global__ void __launch_bounds__(kBlock)
blue_prep_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, int N, int M, const float2 *__restrict__ chirp,
                 const float *__restrict__ window, float2 *__restrict__ a)
{
    const int64_t total = nspec * (int64_t)M;
    for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kBlock) {
        const int64_t b = idx / M;
        const int n = (int)(idx - b * M);
        c32 v = make_float2(0.0f, 0.0f);
        if (n < N) {
            const uint32_t r = *reinterpret_cast<const uint16_t *>(bytes + ((size_t)b * N + n) * 2);
            float re = (float)(r & 255u) - 127.0f, im = (float)(r >> 8) - 127.0f;
            if (window) { const float w = window[n]; re *= w; im *= w; }
            v = make_float2(re, im);
            if (chirp) v = cmul(v, chirp[n]);
        }
        a[idx] = v;
    }
}

// One Stockham pass of radix R over a batch of M-point rows: before the pass the row holds Ns-point transforms.
template <int R>
__global__ void __launch_bounds__(kBlock)
gfft_pass_kernel(const float2 *__restrict__ in, float2 *__restrict__ out, int M, int Ns, const float2 *__restrict__ tw,
                 int64_t total /* rows * M / R */)
{
    const int q = M / R;
    for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kBlock) {
        const int64_t b = idx / q;
        const int j = (int)(idx - b * q);
        const int k = j & (Ns - 1);
        const float2 *src = in + (size_t)b * M;
        c32 v[R];
#pragma unroll
        for (int r = 0; r < R; r++) v[r] = src[j + r * q];
        if (Ns > 1) {
            const int tstep = M / (Ns * R);                      // W_{Ns R}^{r k} = W_M^{r k tstep}
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmul(v[r], tw[(r * k * tstep) & (M - 1)]);
        }
        Dft<R>::run(v);
        float2 *dst = out + (size_t)b * M + (size_t)(j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; r++) dst[r * Ns] = v[r];
    }
}

// C = conj(A . Bhat), in place
__global__ void __launch_bounds__(kBlock)
blue_mul_kernel(float2 *__restrict__ a, const float2 *__restrict__ bhat, int M, int64_t total)
{
    for (int64_t idx = (int64_t)blockIdx.x * kBlock + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * kBlock) {
        const c32 p = cmul(a[idx], bhat[idx & (M - 1)]);
        a[idx] = make_float2(p.x, -p.y);
    }
}

// rows[slice][(k + N/2) mod N] += sum over the spectra b = slice, slice + S, ... of |y[b][k]|^2, k < N
__global__ void __launch_bounds__(kBlock)
blue_acc_kernel(const float2 *__restrict__ y, int64_t nspec, int N, int M, int slices, double *__restrict__ rows)
{
    const int k = blockIdx.x * kBlock + threadIdx.x;
    const int slice = blockIdx.y;
    if (k >= N) return;
    double total = 0.0;
    float acc = 0.0f;
    int n = 0;
    for (int64_t b = slice; b < nspec; b += slices) {
        const c32 v = y[(size_t)b * M + k];
        acc = fmaf(v.x, v.x, fmaf(v.y, v.y, acc));
        if (++n == 128) { total += (double)acc; acc = 0.0f; n = 0; }
    }
    total += (double)acc;
    int out = k + N / 2;
    if (out >= N) out -= N;
    rows[(size_t)slice * N + out] += total;
}

int grid_for(int64_t total, int sms)
{
    int64_t blocks = (total + kBlock - 1) / kBlock;
    const int64_t cap = (int64_t)sms * 16;
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

// forward FFT_M of `rows` rows, ping-pong between p and q; returns the buffer that holds the result
float2 *gfft_forward(float2 *p, float2 *q, int64_t rows, int M, const float2 *tw, int sms, cudaStream_t stream, int *launches)
{
    int lg = 0;
    while ((1 << lg) < M) lg++;
    int Ns = 1;
    if (lg & 1) {                                               // odd log: one radix-2 pass first
        const int64_t total = rows * (int64_t)(M / 2);
        gfft_pass_kernel<2><<<grid_for(total, sms), kBlock, 0, stream>>>(p, q, M, Ns, tw, total);
        Ns *= 2; (*launches)++;
        float2 *t = p; p = q; q = t;
    }
    while (Ns < M) {
        const int64_t total = rows * (int64_t)(M / 4);
        gfft_pass_kernel<4><<<grid_for(total, sms), kBlock, 0, stream>>>(p, q, M, Ns, tw, total);
        Ns *= 4; (*launches)++;
        float2 *t = p; p = q; q = t;
    }
    return p;
}

}  // namespace

int rpf_bluestein_M(int N)
{
    if ((N & (N - 1)) == 0) return N;                           // plain FFT
    int M = 1;
    while (M < 2 * N - 1) M <<= 1;
    return M;
}

cudaError_t rpf_launch_bluestein(const RpfBluestein &B, const uint8_t *bytes, int64_t nspec, int *launches)
{
    if (nspec <= 0) return cudaSuccess;
    const bool chirped = B.M != B.N;
    const int64_t totalM = nspec * (int64_t)B.M;
    blue_prep_kernel<<<grid_for(totalM, B.sms), kBlock, 0, B.stream>>>(bytes, nspec, B.N, B.M, chirped ? B.chirp : nullptr,
                                                                       B.window, B.buf0);
    (*launches)++;
    float2 *res = gfft_forward(B.buf0, B.buf1, nspec, B.M, B.twM, B.sms, B.stream, launches);
    if (chirped) {
        blue_mul_kernel<<<grid_for(totalM, B.sms), kBlock, 0, B.stream>>>(res, B.bhat, B.M, totalM);
        (*launches)++;
        res = gfft_forward(res, res == B.buf0 ? B.buf1 : B.buf0, nspec, B.M, B.twM, B.sms, B.stream, launches);
    }
    const int slices = (int)(nspec < B.slices ? nspec : B.slices);
    dim3 grid((unsigned)((B.N + kBlock - 1) / kBlock), (unsigned)slices);
    blue_acc_kernel<<<grid, kBlock, 0, B.stream>>>(res, nspec, B.N, B.M, slices, B.rows);
    (*launches)++;
    return cudaGetLastError();
}
