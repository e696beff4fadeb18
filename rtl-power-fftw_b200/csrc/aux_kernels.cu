// aux_kernels.cu -- everything around the fused kernel:
//   * small_kernel      : any power-of-two N in [2,256] (generic smem Stockham)
//   * fourstep_outer    : N = M*4096 (M = 2,4,8,16): radix-M over the stride-4096
//                         digit, twiddle W_N^{m k1}, complex64 scratch for the
//                         fused 4096-point kernel
//   * reduce_rows       : private float64 rows -> pwr[N], fixed order
//   * synth             : counter-based IQ byte source (device side)
//   * write_data        : Acquisition::write_data arithmetic
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "synth_gen.h"

namespace {

// ---------------------------------------------------------------------------
// Generic small N.  One spectrum per CTA at a time, 128 threads, radix-2
// Stockham (DIT): before a stage the array holds l-point DFTs of the stride-N/l
// subsequences as x[(j*2+q)*m + k]; the stage writes y[(q'*l + j)*m + k].
// ---------------------------------------------------------------------------
constexpr int kSmallBlock = 128;
constexpr int kSmallFlush = 128;

__global__ void __launch_bounds__(kSmallBlock)
small_kernel(int N, const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
             const float *__restrict__ window, double *__restrict__ rows)
{
    __shared__ c32 buf[2][256];
    const int tid = threadIdx.x;
    const int64_t s0 = (nspec * (int64_t)blockIdx.x) / gridDim.x;
    const int64_t s1 = (nspec * (int64_t)(blockIdx.x + 1)) / gridDim.x;
    double *row = rows + (size_t)blockIdx.x * N;
    float acc[2] = {0.0f, 0.0f};
    int since = 0;
    auto flush = [&]() {
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int bin = tid + h * kSmallBlock;
            if (bin < N) { row[(bin + N / 2) & (N - 1)] += (double)acc[h]; acc[h] = 0.0f; }
        }
    };
    for (int64_t s = s0; s < s1; s++) {
        const uint8_t *raw = bytes + (size_t)s * 2 * N;
        for (int n = tid; n < N; n += kSmallBlock) {
            const uint32_t r = *reinterpret_cast<const uint16_t *>(raw + 2 * n);
            float re = (float)(r & 255u) - 127.0f, im = (float)(r >> 8) - 127.0f;   // src/datastore.cxx:74-75
            if (window) { const float w = window[n]; re *= w; im *= w; }            // :76-77
            buf[0][n] = make_float2(re, im);
        }
        __syncthreads();
        int cur = 0;
        for (int l = 1; l < N; l <<= 1) {
            const int m = N / (2 * l);
            for (int b = tid; b < N / 2; b += kSmallBlock) {
                const int j = b / m, k = b - j * m;
                const c32 a = buf[cur][(2 * j) * m + k];
                const c32 bb = cmul(buf[cur][(2 * j + 1) * m + k], twiddle[j * m]);   // W_{2l}^j = W_N^{j m}
                buf[cur ^ 1][j * m + k] = cadd(a, bb);
                buf[cur ^ 1][(l + j) * m + k] = csub(a, bb);
            }
            __syncthreads();
            cur ^= 1;
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int bin = tid + h * kSmallBlock;
            if (bin < N) { const c32 y = buf[cur][bin]; acc[h] = fmaf(y.x, y.x, fmaf(y.y, y.y, acc[h])); }
        }
        __syncthreads();
        if (++since == kSmallFlush) { flush(); since = 0; }
    }
    flush();
}

// ---------------------------------------------------------------------------
// Four-step outer pass.  n = n1*4096 + m, k = k1 + M*k'.  Persistent threads: the grid is a whole
// number of 4096-thread "rows", so a thread keeps the same m for every spectrum it visits and its
// M-1 twiddles W_N^{m k1} (and window values) live in registers; per spectrum it reads M samples
// (stride 8 KiB, 64 contiguous bytes per warp), runs the radix-M butterfly and writes M coalesced
// complex64 values.
// ---------------------------------------------------------------------------
template <int M, bool WINDOW>
__global__ void __launch_bounds__(256)
fourstep_outer_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twbig,
                      const float *__restrict__ window, float2 *__restrict__ scratch)
{
    constexpr int NI = 4096;
    constexpr int NT = M * NI;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int m = (int)(gtid & (NI - 1));
    const int64_t spec0 = gtid / NI, stride = ((int64_t)gridDim.x * blockDim.x) / NI;   // whole rows only
    float2 tw[M];
    float win[WINDOW ? M : 1];
#pragma unroll
    for (int k1 = 1; k1 < M; k1++) tw[k1] = __ldg(&twbig[(m * k1) & (NT - 1)]);
    if (WINDOW) {
#pragma unroll
        for (int n1 = 0; n1 < M; n1++) win[n1] = window[n1 * NI + m];
    }
    for (int64_t spec = spec0; spec < nspec; spec += stride) {
        const uint8_t *raw = bytes + (size_t)spec * 2 * NT + 2 * m;
        uint32_t r[M];
#pragma unroll
        for (int n1 = 0; n1 < M; n1++) r[n1] = *reinterpret_cast<const uint16_t *>(raw + (size_t)n1 * 2 * NI);
        c32 v[M];
#pragma unroll
        for (int n1 = 0; n1 < M; n1++) {
            float re = (float)(r[n1] & 255u) - 127.0f, im = (float)(r[n1] >> 8) - 127.0f;   // src/datastore.cxx:74-75
            if (WINDOW) { re *= win[n1]; im *= win[n1]; }                                     // :76-77
            v[n1] = make_float2(re, im);
        }
        Dft<M>::run(v);
        float2 *dst = scratch + (size_t)spec * NT + m;
        dst[0] = v[0];
#pragma unroll
        for (int k1 = 1; k1 < M; k1++) dst[(size_t)k1 * NI] = cmul(v[k1], tw[k1]);
    }
}

// Private float64 rows -> pwr[], parallel over rows AND deterministic: a block covers 32 output bins; its 8 warps
// each sum the rows r = w, w + 8, ... (in that order, 4 loads in flight), and the 8 partial sums are then added in
// warp order.  The grouping depends only on `nrows`, so a given sequence of pushes always gives the same bits.
constexpr int kRedCols = 32, kRedLanes = 8;

__global__ void __launch_bounds__(kRedCols * kRedLanes)
reduce_rows_kernel(const double *__restrict__ rows, int nrows, int64_t rowlen, double *__restrict__ pwr)
{
    __shared__ double part[kRedLanes][kRedCols];
    const int col = threadIdx.x & (kRedCols - 1), lane = threadIdx.x / kRedCols;
    const int64_t i = (int64_t)blockIdx.x * kRedCols + col;
    double s = 0.0;
    if (i < rowlen) {
        int r = lane;
        for (; r + 3 * kRedLanes < nrows; r += 4 * kRedLanes) {
            const double a0 = rows[(size_t)r * rowlen + i], a1 = rows[(size_t)(r + kRedLanes) * rowlen + i];
            const double a2 = rows[(size_t)(r + 2 * kRedLanes) * rowlen + i], a3 = rows[(size_t)(r + 3 * kRedLanes) * rowlen + i];
            s += a0; s += a1; s += a2; s += a3;
        }
        for (; r < nrows; r += kRedLanes) s += rows[(size_t)r * rowlen + i];
    }
    part[lane][col] = s;
    __syncthreads();
    if (lane == 0 && i < rowlen) {
        double t = part[0][col];
#pragma unroll
        for (int l = 1; l < kRedLanes; l++) t += part[l][col];
        pwr[i] = t;
    }
}

// Four-step: rows[c][k] holds sum |X|^2 of FFT index k of the inner transform for outer bin ko = c % M.
// pwr[i] (reference order, DC at N/2) = sum over the CTAs bound to that ko; same fixed-order scheme as above.
__global__ void __launch_bounds__(kRedCols * kRedLanes)
reduce_rows_fourstep_kernel(const double *__restrict__ rows, int nrows, int M, int inner, double *__restrict__ pwr)
{
    __shared__ double part[kRedLanes][kRedCols];
    const int col = threadIdx.x & (kRedCols - 1), lane = threadIdx.x / kRedCols;
    const int64_t n = (int64_t)M * inner;
    const int64_t i = (int64_t)blockIdx.x * kRedCols + col;
    double s = 0.0;
    if (i < n) {
        const int64_t bin = (i + n / 2) & (n - 1);          // undo the fftshift: i = (bin + n/2) mod n
        const int ko = (int)(bin % M);
        const int64_t k = bin / M;
        for (int c = ko + lane * M; c < nrows; c += kRedLanes * M) s += rows[(size_t)c * inner + k];
    }
    part[lane][col] = s;
    __syncthreads();
    if (lane == 0 && i < n) {
        double t = part[0][col];
#pragma unroll
        for (int l = 1; l < kRedLanes; l++) t += part[l][col];
        pwr[i] = t;
    }
}

__global__ void synth_kernel(uint8_t *bytes, size_t len, uint64_t off, uint64_t seed, int mode)
{
    // word-aligned body; bytes[i] is stream byte off+i
    const uint64_t first_word = off >> 3, last_word = (off + len + 7) >> 3;
    for (uint64_t w = first_word + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < last_word;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t v = rpf_synth_word(seed, mode, w);
        const uint64_t base = w << 3;
        if (base >= off && base + 8 <= off + len && (((uintptr_t)(bytes + (base - off))) & 7) == 0) {
            *reinterpret_cast<uint64_t *>(bytes + (base - off)) = v;
        } else {
            for (int b = 0; b < 8; b++) {
                const uint64_t i = base + b;
                if (i >= off && i < off + len) bytes[i - off] = (uint8_t)(v >> (8 * b));
            }
        }
    }
}

// Acquisition::write_data, reference src/acquisition.cxx:377,390-399 (+ the float32 matrix row of :400-405).
__global__ void write_data_kernel(const double *__restrict__ pwr, int N, int64_t repeats_done, double tuned_freq,
                                  int samplerate, int linear, const double *__restrict__ baseline,
                                  double *__restrict__ freq_out, double *__restrict__ val_out, float *__restrict__ row_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    double p = pwr[i];
    // :377 reads pwr[N/2 + 1]: for N = 2 that is one past the end of the vector in the reference (undefined there);
    // here the index wraps, i.e. the DC bin becomes the mean of its only two neighbours pwr[0], pwr[0].
    if (i == N / 2) p = (pwr[N / 2 - 1] + pwr[(N / 2 + 1) % N]) / 2;
    if (freq_out) freq_out[i] = tuned_freq + (i - N / 2.0) * samplerate / N;   // :391
    double v = p / repeats_done / N / samplerate;                          // :393,397
    if (!linear) v = 10 * log10(v);
    v -= (baseline ? baseline[i] : 0);
    val_out[i] = v;
    if (row_out) row_out[i] = (float)v;                                    // :404 fpwrdb = pwrdb
}

}  // namespace

cudaError_t rpf_launch_small(int N, const RpfLaunch &L)
{
    if (L.nspec <= 0) return cudaSuccess;
    const int grid = (int)(L.nspec < L.grid ? L.nspec : L.grid);
    small_kernel<<<grid, kSmallBlock, 0, L.stream>>>(N, L.bytes, L.nspec, L.twiddle, L.window, L.rows);
    return cudaGetLastError();
}

template <int M>
static cudaError_t launch_outer(const uint8_t *bytes, int64_t nspec, const float2 *twbig, const float *window,
                                float2 *scratch, int sms, cudaStream_t stream)
{
    // whole 4096-thread rows: 16 blocks of 256 per row, about 16 rows' worth of blocks per wave
    int64_t rows = nspec < sms ? nspec : sms;
    const int blocks = (int)(rows * 16);
    if (window)
        fourstep_outer_kernel<M, true><<<blocks, 256, 0, stream>>>(bytes, nspec, twbig, window, scratch);
    else
        fourstep_outer_kernel<M, false><<<blocks, 256, 0, stream>>>(bytes, nspec, twbig, nullptr, scratch);
    return cudaGetLastError();
}

cudaError_t rpf_launch_fourstep_outer(int M, const uint8_t *bytes, int64_t nspec, const float2 *twbig,
                                      const float *window, float2 *scratch, int sms, cudaStream_t stream)
{
    if (nspec <= 0) return cudaSuccess;
    switch (M) {
        case 2: return launch_outer<2>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 4: return launch_outer<4>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 8: return launch_outer<8>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 16: return launch_outer<16>(bytes, nspec, twbig, window, scratch, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t rpf_launch_reduce_rows(const double *rows, int nrows, int64_t rowlen, double *pwr, cudaStream_t stream)
{
    reduce_rows_kernel<<<(unsigned)((rowlen + kRedCols - 1) / kRedCols), kRedCols * kRedLanes, 0, stream>>>(rows, nrows, rowlen, pwr);
    return cudaGetLastError();
}

cudaError_t rpf_launch_reduce_rows_fourstep(const double *rows, int nrows, int M, int inner, double *pwr, cudaStream_t stream)
{
    const int64_t n = (int64_t)M * inner;
    reduce_rows_fourstep_kernel<<<(unsigned)((n + kRedCols - 1) / kRedCols), kRedCols * kRedLanes, 0, stream>>>(rows, nrows, M, inner, pwr);
    return cudaGetLastError();
}

cudaError_t rpf_launch_synth(uint8_t *bytes, size_t len, uint64_t off, uint64_t seed, int mode, int sms, cudaStream_t stream)
{
    if (len == 0) return cudaSuccess;
    const uint64_t words = (len + 15) / 8;
    int blocks = (int)((words + 255) / 256);
    if (blocks > sms * 32) blocks = sms * 32;
    synth_kernel<<<blocks, 256, 0, stream>>>(bytes, len, off, seed, mode);
    return cudaGetLastError();
}

cudaError_t rpf_launch_write_data(const double *pwr, int N, int64_t repeats_done, double tuned_freq, int samplerate,
                                  int linear, const double *baseline, double *freq_out, double *val_out,
                                  float *row_out, cudaStream_t stream)
{
    write_data_kernel<<<(N + 255) / 256, 256, 0, stream>>>(pwr, N, repeats_done, tuned_freq, samplerate, linear,
                                                           baseline, freq_out, val_out, row_out);
    return cudaGetLastError();
}
