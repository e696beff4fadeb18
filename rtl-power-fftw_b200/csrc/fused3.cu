// fused3.cu -- the hot path: u8-IQ unpack (+window) -> N-point complex FFT ->
// |X|^2 accumulate, one fused persistent kernel, N in {512, 1024, 2048, 4096}.
//
// Replaces hot loop A (reference src/datastore.cxx:68-80), fftwf_execute
// (src/datastore.cxx:82) and hot loop B (src/datastore.cxx:83-85).
//
// Shape of the computation (per spectrum, N = R1*R2*R3, 16 values per thread,
// T = N/16 threads per spectrum, 256-thread CTAs => G = 256/T spectra in flight):
//   TMA bulk copy (cp.async.bulk + mbarrier) stages 8 KiB of raw bytes per CTA
//     iteration into a 4-deep shared-memory ring          [HBM: 2 B / sample]
//   pass 1: radix-R1 over the stride-N/R1 digit, straight from the raw bytes
//           (bias removed exactly, see unpack()), twiddle W_N^{t k1}   -> S1
//   pass 2: radix-R2, twiddle W_{R2 R3}^{n3 k2}                          -> S2
//   pass 3: radix-R3, |X|^2 accumulated in registers in pass-3 order
//   every kFlushEvery iterations: registers -> the CTA's private float64 row,
//     at index (k + N/2) mod N.  The (-1)^n multiply of src/datastore.cxx:73
//     is never executed: it is exactly that rotation of the output.
// Twiddles live in registers as scalars (FFMA2 broadcasts a scalar operand).
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kBlock = 256;
#ifndef RPF_STAGES
#define RPF_STAGES 3                // TMA ring depth.  3 x 8 KiB keeps two CTAs/SM under the 200 KiB shared
#endif                              // carve-out, so the L1 keeps ~28 KiB for the few spilled twiddles
constexpr int kStages = RPF_STAGES;
constexpr int kIterBytes = 8192;    // G * 2N for every supported N
constexpr int kFlushEvery = 128;    // CTA iterations between float64 flushes

// Development switches (A/B variants are built by `make variant`; the defaults are the product).
#ifndef RPF_MINBLOCKS
#define RPF_MINBLOCKS 2             // CTAs per SM the register allocator must leave room for
#endif
#ifndef RPF_EXCH_MODE
#define RPF_EXCH_MODE 1             // 1: exchange 2 warp-local + split mbarrier on S1; 0: two CTA barriers
#endif

template <int N> struct Radices;
template <> struct Radices<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8;  };
template <> struct Radices<1024> { static constexpr int R1 = 16, R2 = 8,  R3 = 8;  };
template <> struct Radices<2048> { static constexpr int R1 = 8,  R2 = 16, R3 = 16; };
template <> struct Radices<4096> { static constexpr int R1 = 16, R2 = 16, R3 = 16; };

template <int N>
struct Geo {
    static constexpr int R1 = Radices<N>::R1, R2 = Radices<N>::R2, R3 = Radices<N>::R3;
    static constexpr int VPT = 16;
    static constexpr int T = N / VPT;             // threads per spectrum
    static constexpr int G = kBlock / T;          // spectra per CTA iteration
    static constexpr int C1 = VPT / R1, C2 = VPT / R2, C3 = VPT / R3;   // columns per thread
    static constexpr int NC1 = N / R1;            // pass-1 columns (= stride of the first digit)
    static constexpr int P1 = NC1 + (R3 % 16);    // S1 row pitch (bank-conflict padding)
    static constexpr int S1_ELEMS = R1 * P1;      // per spectrum
    static constexpr int P2 = R3 + 1;             // S2 row pitch: rows are pass-3 columns u = k1*R2 + k2
    static constexpr int S2_ELEMS = (N / R3) * P2;
    // Exchange 2 (pass 2 -> pass 3) only moves data between threads that share k1.  When every
    // warp owns the same k1 set in both passes the exchange is warp-local: __syncwarp, no CTA barrier.
    static constexpr bool WARP_LOCAL2 = RPF_EXCH_MODE && ((T == 32) || (C2 == 1 && C3 == 1 && R2 == R3));
    static constexpr size_t SMEM = 64 /*mbarriers*/ + (size_t)kStages * kIterBytes +
                                   (size_t)G * (S1_ELEMS + S2_ELEMS) * sizeof(c32);
    static_assert(G * 2 * N == kIterBytes, "8 KiB of raw bytes per CTA iteration");
    static_assert(R1 * R2 * R3 == N, "radix schedule");
};

template <int N, bool WINDOW>
__global__ void __launch_bounds__(kBlock, RPF_MINBLOCKS)
fused3_kernel(const uint8_t *__restrict__ bytes, int64_t nspec,
              const float2 *__restrict__ twiddle, const float *__restrict__ window, double *__restrict__ rows)
{
    using g = Geo<N>;
    constexpr int R1 = g::R1, R2 = g::R2, R3 = g::R3, T = g::T, G = g::G;
    constexpr int C1 = g::C1, C2 = g::C2, C3 = g::C3, NC1 = g::NC1, P1 = g::P1;

    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *ring = smem + 64;
    c32 *S1 = reinterpret_cast<c32 *>(ring + (size_t)kStages * kIterBytes);
    c32 *S2 = S1 + (size_t)G * g::S1_ELEMS;

    const int tid = threadIdx.x;
    const int grp = tid / T;            // which of the G spectra of this iteration
    const int tau = tid % T;            // thread within the spectrum
    c32 *s1 = S1 + (size_t)grp * g::S1_ELEMS;
    c32 *s2 = S2 + (size_t)grp * g::S2_ELEMS;

    // ---- work range (CTA iterations; one iteration = G consecutive spectra) --
    const int64_t nitems = nspec;
    const int64_t niter = (nitems + G - 1) / G;
    const int64_t it0 = (niter * (int64_t)blockIdx.x) / gridDim.x;
    const int nit = (int)((niter * (int64_t)(blockIdx.x + 1)) / gridDim.x - it0);   // < 2^31 per CTA
    if (nit <= 0) return;               // nothing for this CTA (uniform per CTA)
    // group `grp` holds a real spectrum in local iterations j < jvalid (only the stream's last
    // iteration can be partial)
    const int64_t jv = (nitems - grp - it0 * G + G - 1) / G;
    const int jvalid = jv < nit ? (int)jv : nit;

    // ---- per-thread constants: twiddles (+ window) in registers --------------
    float2 tw1[C1][R1 - 1], tw2[C2][R2 - 1];
#pragma unroll
    for (int c = 0; c < C1; c++) {
        const int t = tau + T * c;
#pragma unroll
        for (int k1 = 1; k1 < R1; k1++) tw1[c][k1 - 1] = twiddle[(t * k1) & (N - 1)];
    }
#pragma unroll
    for (int c = 0; c < C2; c++) {
        const int n3 = (tau + T * c) % R3;
#pragma unroll
        for (int k2 = 1; k2 < R2; k2++) tw2[c][k2 - 1] = twiddle[(n3 * k2 * R1) & (N - 1)];
    }
    float win[WINDOW ? C1 : 1][WINDOW ? R1 : 1];
    if (WINDOW) {
#pragma unroll
        for (int c = 0; c < C1; c++)
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) win[c][n1] = window[n1 * NC1 + tau + T * c];
    }

    float acc[C3][R3];
#pragma unroll
    for (int c = 0; c < C3; c++)
#pragma unroll
        for (int k = 0; k < R3; k++) acc[c][k] = 0.0f;

    const int64_t rowlen = N;
    double *row = rows + ((size_t)blockIdx.x * G + grp) * (size_t)rowlen;

    auto flush = [&]() {
#pragma unroll
        for (int c = 0; c < C3; c++) {
            const int u = tau + T * c;
            const int k1 = u / R2, k2 = u % R2;
#pragma unroll
            for (int k3 = 0; k3 < R3; k3++) {
                const int k = k1 + R1 * k2 + R1 * R2 * k3;                 // FFT output index
                const int64_t out = (k + (rowlen >> 1)) & (rowlen - 1);    // fftshift == (-1)^n trick
                row[out] += (double)acc[c][k3];
                acc[c][k3] = 0.0f;
            }
        }
    };

    // ---- TMA ring ------------------------------------------------------------
    const uint8_t *cta_bytes = bytes + (size_t)it0 * kIterBytes;
    auto issue = [&](int j) {           // called by thread 0 only
        const int slot = j % kStages;
        const int64_t left = nitems - (it0 + j) * G;
        const uint32_t nbytes = (uint32_t)((left < G ? left : G) * 2 * N);
        const uint32_t bar = smem_u32(&bars[slot]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, nbytes);
        tma_load_1d(smem_u32(ring + (size_t)slot * kIterBytes), cta_bytes + (size_t)j * kIterBytes, nbytes, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) mbar_init(smem_u32(&bars[s]), 1);
        mbar_init(smem_u32(&bars[kStages]), kBlock / 32);   // "S1 has been read" -- one arrival per warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < kStages; s++)
            if (s < nit) issue(s);
    const uint32_t bar_s1free = smem_u32(&bars[kStages]);

    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const bool valid = (G == 1) ? true : (j < jvalid);
        c32 v[16];

        // ================= pass 1: load + radix-R1 over the high digit =========
        {
            const int slot = j % kStages;
            mbar_wait(smem_u32(&bars[slot]), (uint32_t)((j / kStages) & 1));
            const uint8_t *raw = ring + (size_t)slot * kIterBytes + (size_t)grp * 2 * N;
#pragma unroll
            for (int c = 0; c < C1; c++)
#pragma unroll
                for (int n1 = 0; n1 < R1; n1++) {
                    const uint32_t r = *reinterpret_cast<const uint16_t *>(raw + 2 * (n1 * NC1 + tau + T * c));
                    c32 x = unpack_biased(r);
                    if (WINDOW) {   // (u8 - 127.0f) * w, one rounding per rail: src/datastore.cxx:75-77
                        x = __fadd2_rn(x, make_float2(-kBias, -kBias));
                        x = __fmul2_rn(x, make_float2(win[c][n1], win[c][n1]));
                    }
                    v[c * R1 + n1] = x;
                }
        }
#pragma unroll
        for (int c = 0; c < C1; c++) {
            Dft<R1>::run(&v[c * R1]);
            if (!WINDOW)   // remove the R1 * (32768+127) left in output 0 -- exact
                v[c * R1] = __fadd2_rn(v[c * R1], make_float2(-(float)R1 * kBias, -(float)R1 * kBias));
            const int t = tau + T * c;
#pragma unroll
            for (int k1 = 1; k1 < R1; k1++) v[c * R1 + k1] = cmul(v[c * R1 + k1], tw1[c][k1 - 1]);
            // split barrier: every warp must have finished reading S1 in the previous iteration
            // (it arrived on bar_s1free right after its pass-2 loads, long ago)
            if (g::WARP_LOCAL2 && c == 0 && j > 0) mbar_wait(bar_s1free, (uint32_t)((j - 1) & 1));
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) s1[k1 * P1 + t] = v[c * R1 + k1];
        }
        __syncthreads();   // S1 complete; ring slot consumed
        if (tid == 0 && j + kStages < nit) issue(j + kStages);

        // ================= pass 2 ==============================================
#pragma unroll
        for (int c = 0; c < C2; c++) {
            const int q = tau + T * c;
            const int k1 = q / R3, n3 = q % R3;
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) v[c * R2 + n2] = s1[k1 * P1 + n2 * R3 + n3];
        }
        if (g::WARP_LOCAL2) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_s1free);   // this warp's S1 reads are issued (LSU is in order)
        }
#pragma unroll
        for (int c = 0; c < C2; c++) {
            const int q = tau + T * c;
            const int k1 = q / R3, n3 = q % R3;
            Dft<R2>::run(&v[c * R2]);
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) {
                const c32 y = (k2 == 0) ? v[c * R2] : cmul(v[c * R2 + k2], tw2[c][k2 - 1]);
                s2[(k1 * R2 + k2) * g::P2 + n3] = y;               // immediate offsets off one base register
            }
        }
        // S2 rows of this warp's k1 set complete.  (When not warp-local, the CTA barrier also keeps
        // the previous iteration's pass-3 reads ahead of these writes: see the hazard note above.)
        if (g::WARP_LOCAL2) __syncwarp(); else __syncthreads();

        // ================= pass 3 + |X|^2 ======================================
#pragma unroll
        for (int c = 0; c < C3; c++) {
            const int u = tau + T * c;                             // = k1*R2 + k2
#pragma unroll
            for (int n3 = 0; n3 < R3; n3++) v[c * R3 + n3] = s2[u * g::P2 + n3];
        }
#pragma unroll
        for (int c = 0; c < C3; c++) {
            Dft<R3>::run(&v[c * R3]);
            if (valid) {
#pragma unroll
                for (int k3 = 0; k3 < R3; k3++) {
                    const c32 y = v[c * R3 + k3];
                    acc[c][k3] = fmaf(y.x, y.x, fmaf(y.y, y.y, acc[c][k3]));
                }
            }
        }
        if (++since_flush == kFlushEvery) { flush(); since_flush = 0; }
    }
    flush();
}

template <int N>
cudaError_t launch_n(const RpfLaunch &L)
{
    using g = Geo<N>;
    const int64_t niter = (L.nspec + g::G - 1) / g::G;
    if (niter <= 0) return cudaSuccess;
    dim3 grid((unsigned)(niter < L.grid ? niter : L.grid)), block(kBlock);
    if (L.window)
        fused3_kernel<N, true><<<grid, block, g::SMEM, L.stream>>>(L.bytes, L.nspec, L.twiddle, L.window, L.rows);
    else
        fused3_kernel<N, false><<<grid, block, g::SMEM, L.stream>>>(L.bytes, L.nspec, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}

template <int N>
cudaError_t configure_n()
{
    using g = Geo<N>;
    cudaError_t e = cudaFuncSetAttribute(fused3_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fused3_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g::SMEM);
}

}  // namespace

bool rpf_fused3_supported(int N) { return N == 512 || N == 1024 || N == 2048 || N == 4096; }
int rpf_fused3_groups(int N) { return kBlock / (N / 16); }

cudaError_t rpf_fused3_configure()
{
    cudaError_t e;
    if ((e = configure_n<512>()) != cudaSuccess) return e;
    if ((e = configure_n<1024>()) != cudaSuccess) return e;
    if ((e = configure_n<2048>()) != cudaSuccess) return e;
    return configure_n<4096>();
}

cudaError_t rpf_launch_fused3(int N, const RpfLaunch &L)
{
    switch (N) {
        case 512: return launch_n<512>(L);
        case 1024: return launch_n<1024>(L);
        case 2048: return launch_n<2048>(L);
        case 4096: return launch_n<4096>(L);
        default: return cudaErrorInvalidValue;
    }
}
