// fused4096.cu -- the headline configuration (4096 bins) gets its own kernel.
//
// Same arithmetic as fused3_kernel<4096> (16 x 16 x 16, packed FP32, exact
// offset-binary unpack, register accumulation, float64 private rows), but the
// thread <-> data mapping is chosen so that
//   * pass-2 twiddles W_256^{n3 k2} are WARP-UNIFORM (a warp owns one n3 and the
//     32 (spectrum, k1) combinations of two spectra): they come out of constant
//     memory through the uniform datapath and cost no vector registers, so the
//     kernel fits 128 registers without spilling;
//   * pass-1 results are written as 8 x STS.128 per thread into a [t][k1] layout
//     (16-byte units XOR-swizzled by t & 7), pass 2 reads them back as whole
//     128-byte rows: conflict-free both ways;
//   * pass-3 threads are (k2, k1) with k1 fastest, so the S2 exchange needs no
//     padding and the float64 flush is fully coalesced.
// One CTA = 512 threads = two consecutive spectra per iteration (16 KiB of raw
// bytes per TMA bulk copy), one CTA per SM, 16 warps.
//
// Replaces reference src/datastore.cxx:68-87 for N = 4096 (and is the inner
// transform of the four-step path for N = M * 4096).
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;
constexpr int kT = 256;               // threads per spectrum
constexpr int kG = 2;                 // spectra per CTA iteration
constexpr int kBlock4 = kT * kG;      // 512
#ifndef RPF4096_STAGES
#define RPF4096_STAGES 3
#endif
constexpr int kStages4 = RPF4096_STAGES;
// Ablation switches for timing experiments only (results are wrong when set):
//   RPF_ABL & 1: skip the floating-point work; RPF_ABL & 2: skip the shared-memory exchange loads/stores
//   (barriers stay: the TMA ring protocol needs them).
#ifndef RPF_ABL
#define RPF_ABL 0
#endif
#define ABL_FP(x)   do { if (!(RPF_ABL & 1)) { x; } } while (0)
#define ABL_MEM(x)  do { if (!(RPF_ABL & 2)) { x; } } while (0)
constexpr int kSlotBytes = kG * 2 * kN;   // 16 KiB
constexpr int kFlushEvery4 = 128;         // iterations between float64 flushes (fp32 partial sums of <= 128 spectra)
constexpr int kHdr4 = 128;                // mbarriers; keeps S1 128-byte aligned for the XOR swizzle
constexpr size_t kSmem4 = kHdr4 + (size_t)kStages4 * kSlotBytes + (size_t)kG * 2 * kN * sizeof(c32);

// W_256^{n3 k2} = exp(-2 pi i n3 k2 / 256), [n3][k2]; filled by rpf_fused4096_configure().
__constant__ float2 c_tw2_4096[16][16];

template <bool WINDOW, bool SRC_C32>
__global__ void __launch_bounds__(kBlock4, 1)
fused4096_kernel(const uint8_t *__restrict__ bytes, const float2 *__restrict__ cplx, int64_t nspec, int M,
                 const float2 *__restrict__ twiddle, const float *__restrict__ window, double *__restrict__ rows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint8_t *ring = smem + kHdr4;
    uint8_t *S1 = ring + (size_t)kStages4 * kSlotBytes;     // [2][256 rows t][16 k1] c32, 128 B per row
    uint8_t *S2 = S1 + (size_t)kG * kN * sizeof(c32);       // [2][k2][n3][k1] c32

    const int tid = threadIdx.x;
    // pass 1 / pass 3 role: spectrum s, column index within the spectrum
    const int s13 = tid >> 8;
    const int t = tid & 255;                                // pass 1: column t = 16 b + c
    // pass 2 role: warp <-> n3 (uniform), lane <-> (spectrum, k1)
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // tell the compiler it is warp-uniform
    const int lane = tid & 31;
    const int n3 = warp;
    const int s2 = lane >> 4, k1p2 = lane & 15;

    // ---- work range ------------------------------------------------------------
    const int64_t nitems = SRC_C32 ? nspec * (int64_t)M : nspec;
    const int64_t niter = (nitems + kG - 1) / kG;
    const int64_t it0 = (niter * (int64_t)blockIdx.x) / gridDim.x;
    const int nit = (int)((niter * (int64_t)(blockIdx.x + 1)) / gridDim.x - it0);
    if (nit <= 0) return;
    const int64_t jv = (nitems - s13 - it0 * kG + kG - 1) / kG;
    const int jvalid = jv < nit ? (int)jv : nit;             // spectrum s13 is real for j < jvalid

    // ---- per-thread constants ----------------------------------------------------
    float2 tw1[15];
#pragma unroll
    for (int k1 = 1; k1 < 16; k1++) tw1[k1 - 1] = twiddle[(t * k1) & (kN - 1)];     // W_4096^{t k1}
    float win[WINDOW ? 16 : 1];
    if (WINDOW) {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) win[n1] = window[n1 * 256 + t];
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;

    // shared-memory addresses (32-bit shared-window offsets; all loop offsets are immediates)
    const uint32_t s1_st = smem_u32(S1) + s13 * (kN * 8) + t * 128 + ((t & 7) << 4);            // ^ (j << 4)
    const uint32_t s1_ld = smem_u32(S1) + s2 * (kN * 8) + n3 * 128 + ((((k1p2 >> 1) ^ (n3 & 7)) << 4) | ((k1p2 & 1) << 3));
    const uint32_t s2_st = smem_u32(S2) + s2 * (kN * 8) + n3 * 128 + k1p2 * 8;                   // + k2 * 2048
    const uint32_t s2_ld = smem_u32(S2) + s13 * (kN * 8) + (t >> 4) * 2048 + (t & 15) * 8;        // + n3 * 128
    // pass 3: this thread is (k2 = t >> 4, k1 = t & 15); output k = k1 + 16 k2 + 256 k3 = t + 256 k3

    const int64_t rowlen = (int64_t)kN * M;
    double *row = rows + ((size_t)blockIdx.x * kG + s13) * (size_t)rowlen;
    int kouter = 0;

    auto flush = [&]() {
        // out = (kouter + M * (t + 256 k3) + rowlen/2) mod rowlen: consecutive threads -> consecutive
        // (M = 1) doubles.  Loads first, then adds, then stores: 16 independent RMWs in flight.
        double cur[16];
        int64_t idx[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            const int64_t bin = (int64_t)kouter + (int64_t)M * (t + 256 * k3);
            idx[k3] = (bin + (rowlen >> 1)) & (rowlen - 1);          // fftshift == the (-1)^n trick
            cur[k3] = row[idx[k3]];
        }
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            row[idx[k3]] = cur[k3] + (double)acc[k3];
            acc[k3] = 0.0f;
        }
    };

    // ---- TMA ring ----------------------------------------------------------------
    const uint8_t *cta_bytes = SRC_C32 ? nullptr : bytes + (size_t)it0 * kSlotBytes;
    auto issue = [&](int j) {            // thread 0 only
        const int slot = j % kStages4;
        const int64_t left = nitems - (it0 + j) * kG;
        const uint32_t nbytes = (uint32_t)((left < kG ? left : kG) * 2 * kN);
        const uint32_t bar = smem_u32(&bars[slot]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar, nbytes);
        tma_load_1d(smem_u32(ring + (size_t)slot * kSlotBytes), cta_bytes + (size_t)j * kSlotBytes, nbytes, bar);
    };
    if (!SRC_C32) {
        if (tid == 0) {
#pragma unroll
            for (int s = 0; s < kStages4; s++) mbar_init(smem_u32(&bars[s]), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (tid == 0)
            for (int s = 0; s < kStages4; s++)
                if (s < nit) issue(s);
    }

    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        c32 v[16];
        // ================= pass 1: radix-16 over n1 (stride 256), straight from the bytes ======
        if (SRC_C32) {
            const int64_t item = (it0 + j) * kG + s13;
            const int64_t itc = item < nitems ? item : nitems - 1;       // clamp (result discarded if !valid)
            const int ko = (int)(itc / nspec);
            if (ko != kouter) { flush(); kouter = ko; since_flush = 0; }
            const float2 *src = cplx + ((size_t)(itc - (int64_t)ko * nspec) * M + ko) * (size_t)kN;
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) v[n1] = __ldg(&src[n1 * 256 + t]);
        } else {
            const int slot = j % kStages4;
            mbar_wait(smem_u32(&bars[slot]), (uint32_t)((j / kStages4) & 1));
            const uint32_t raw = smem_u32(ring) + slot * kSlotBytes + s13 * (2 * kN) + 2 * t;
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) {
                uint32_t r;
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(raw + n1 * 512));
                c32 x = unpack_biased(r);
                if (WINDOW) {        // (u8 - 127.0f) * w, one rounding per rail: src/datastore.cxx:75-77
                    x = __fadd2_rn(x, make_float2(-kBias, -kBias));
                    x = __fmul2_rn(x, make_float2(win[n1], win[n1]));
                }
                v[n1] = x;
            }
        }
        ABL_FP(Dft<16>::run(v));
        if (!SRC_C32 && !WINDOW)     // remove the 16 * (32768 + 127) left in output 0 -- exact
            v[0] = __fadd2_rn(v[0], make_float2(-16.0f * kBias, -16.0f * kBias));
#pragma unroll
        for (int k1 = 1; k1 < 16; k1++) ABL_FP(v[k1] = cmul(v[k1], tw1[k1 - 1]));
#pragma unroll
        for (int u = 0; u < 8; u++)  // row t of S1: 16-byte units (k1 = 2u, 2u+1), unit index XOR (t & 7)
            ABL_MEM(asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s1_st ^ (u << 4)), "f"(v[2 * u].x),
                         "f"(v[2 * u].y), "f"(v[2 * u + 1].x), "f"(v[2 * u + 1].y) : "memory"));
        __syncthreads();             // S1 complete; ring slot consumed; S2 of the previous iteration fully read
        if (!SRC_C32 && tid == 0 && j + kStages4 < nit) issue(j + kStages4);

        // ================= pass 2: radix-16 over n2; this warp owns n3 ===========================
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++)
            ABL_MEM(asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[n2].x), "=f"(v[n2].y) : "r"(s1_ld + n2 * 2048)));
        ABL_FP(Dft<16>::run(v));
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            c32 y = v[k2];
            if (k2 > 0) ABL_FP(y = cmul(y, c_tw2_4096[n3][k2]));          // warp-uniform operand
            ABL_MEM(asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s2_st + k2 * 2048), "f"(y.x), "f"(y.y) : "memory"));
            if (RPF_ABL & 2) v[k2] = y;
        }
        __syncthreads();             // S2 complete; S1 free for the next iteration

        // ================= pass 3: radix-16 over n3, |X|^2 ========================================
#pragma unroll
        for (int m = 0; m < 16; m++)
            ABL_MEM(asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[m].x), "=f"(v[m].y) : "r"(s2_ld + m * 128)));
        ABL_FP(Dft<16>::run(v));
        if (j < jvalid) {
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) acc[k3] = fmaf(v[k3].x, v[k3].x, fmaf(v[k3].y, v[k3].y, acc[k3]));
        }
        if (++since_flush == kFlushEvery4) { flush(); since_flush = 0; }
    }
    flush();
}

}  // namespace

cudaError_t rpf_fused4096_configure()
{
    float2 h[16][16];
    const double PI = 3.14159265358979323846264338327950288;
    for (int a = 0; a < 16; a++)
        for (int b = 0; b < 16; b++) {
            const double ang = -2.0 * PI * (double)((a * b) & 255) / 256.0;
            h[a][b] = make_float2((float)cos(ang), (float)sin(ang));
        }
    cudaError_t e = cudaMemcpyToSymbol(c_tw2_4096, h, sizeof(h));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused4096_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem4);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(fused4096_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem4);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fused4096_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem4);
}

int rpf_fused4096_groups() { return kG; }

cudaError_t rpf_launch_fused4096(const RpfLaunch &L)
{
    const bool c32src = L.cplx != nullptr;
    const bool win = L.window != nullptr && !c32src;
    const int64_t nitems = c32src ? L.nspec * L.M : L.nspec;
    const int64_t niter = (nitems + kG - 1) / kG;
    if (niter <= 0) return cudaSuccess;
    dim3 grid((unsigned)(niter < L.grid ? niter : L.grid)), block(kBlock4);
    if (c32src)
        fused4096_kernel<false, true><<<grid, block, kSmem4, L.stream>>>(nullptr, L.cplx, L.nspec, L.M, L.twiddle, nullptr, L.rows);
    else if (win)
        fused4096_kernel<true, false><<<grid, block, kSmem4, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, L.window, L.rows);
    else
        fused4096_kernel<false, false><<<grid, block, kSmem4, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}
