// pingpong4096.cu -- N = 4096, two warp-groups per CTA taking turns on the FP32 pipe.
//
// Measurements behind this kernel (tools/dftbench*.cu, ablations in DESIGN.md): the radix-16
// butterflies are FP32-pipe bound (~88 % of lane peak for two warps per scheduler), while everything
// else a spectrum needs -- staging, unpack, the two shared-memory exchanges, barriers -- is
// latency: waits that block the issuing warp.  A CTA whose warps all do the same thing at the same
// time pays both in series.  Here one CTA = 512 threads = two independent groups of 8 warps; each
// group owns one spectrum per iteration (its own ring slot, S1, S2 and named barriers) and the two
// groups alternate strictly on the math:
//
//     group g:  issue loads | bar.sync TURN[g] | butterflies + twiddles | bar.arrive TURN[1-g] |
//               stores | bar.sync GROUP[g] | issue loads | bar.sync TURN[g] | ...
//
// so while one group's 2 warps per scheduler keep the FP32 pipe busy, the other group's stores,
// barrier and loads complete underneath.  Same arithmetic and layouts as fused4096.cu
// (S1 rows [t][k1] XOR-swizzled, S2 [k2][n3][k1]); pass-2 twiddles come from constant memory
// (two addresses per warp), so the kernel fits 128 registers -> 16 warps per SM.
//
// Replaces reference src/datastore.cxx:68-87 for N = 4096.
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;
constexpr int kGroupThreads = 256;
constexpr int kBlockPP = 512;
#ifndef RPF_PP_STAGES
#define RPF_PP_STAGES 3
#endif
constexpr int kStagesPP = RPF_PP_STAGES;
constexpr int kSpecBytes = 2 * kN;
constexpr int kFlushEveryPP = 128;
constexpr int kHdrPP = 128;
constexpr int kRingOff = kHdrPP;                                   // [group][slot][8 KiB]
constexpr int kS1OffPP = kRingOff + 2 * kStagesPP * kSpecBytes;    // [group][256 rows][128 B]
constexpr int kS2OffPP = kS1OffPP + 2 * kN * 8;                    // [group][k2][n3][k1]
constexpr size_t kSmemPP = (size_t)kS2OffPP + 2 * kN * 8;

__constant__ float2 c_tw2_pp[16][16];                              // W_256^{n3 k2}

#ifdef RPF_PP_TIMING
// development instrumentation: cycles spent (lane 0 of every warp, summed) in
// [0] waiting for the turn, [1] math sections, [2] memory sections, [3] number of sections
__device__ unsigned long long pp_timing[4];
#define PP_T(var) const long long var = clock64()
#define PP_ADD(i, a, b) do { if ((tid & 31) == 0) atomicAdd(&pp_timing[i], (unsigned long long)((b) - (a))); } while (0)
#else
#define PP_T(var)
#define PP_ADD(i, a, b)
#endif

// named barriers: 1,2 = group-local exchange barriers; 3,4 = turn tokens
__device__ __forceinline__ void nb_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <bool WINDOW>
__global__ void __launch_bounds__(kBlockPP, 1)
pingpong4096_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
                    const float *__restrict__ window, double *__restrict__ rows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int g = tid >> 8;                         // group 0 / 1
    const int t = tid & 255;                        // pass 1 column; pass 3 (k2 = t >> 4, k1 = t & 15)
    const int lo = t & 15, hi = t >> 4;             // pass 2 (k1 = lo, n3 = hi)

    // ---- work: CTA takes pairs of spectra; group g takes spectrum 2 (p0 + j) + g ------------
    const int64_t npairs = (nspec + 1) / 2;
    const int64_t p0 = (npairs * (int64_t)blockIdx.x) / gridDim.x;
    const int nit = (int)((npairs * (int64_t)(blockIdx.x + 1)) / gridDim.x - p0);
    if (nit <= 0) return;
    const int nvalid = (g == 1 && (p0 + nit) * 2 > nspec) ? nit - 1 : nit;   // my spectrum exists for j < nvalid

    float2 tw1[15];
#pragma unroll
    for (int k = 1; k < 16; k++) tw1[k - 1] = twiddle[(t * k) & (kN - 1)];
    float win[WINDOW ? 16 : 1];
    if (WINDOW) {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) win[n1] = window[n1 * 256 + t];
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;

    const uint32_t ring = sm + kRingOff + g * (kStagesPP * kSpecBytes);
    const uint32_t S1 = sm + kS1OffPP + g * (kN * 8), S2 = sm + kS2OffPP + g * (kN * 8);
    const uint32_t s1_st = S1 + t * 128 + ((t & 7) << 4);
    const uint32_t s1_ld = S1 + hi * 128 + ((((lo >> 1) ^ (hi & 7)) << 4) | ((lo & 1) << 3));
    const uint32_t s2_st = S2 + hi * 128 + lo * 8;
    const uint32_t s2_ld = S2 + hi * 2048 + lo * 8;
    const uint32_t bars = sm + 8 * (g * kStagesPP);        // my FULL barriers
    const int GROUP_BAR = 1 + g, MY_TURN = 3 + g, OTHER_TURN = 4 - g;

    double *row = rows + ((size_t)blockIdx.x * 2 + g) * kN;
    auto flush = [&]() {
        double cur[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) cur[k3] = row[(t + 256 * k3 + kN / 2) & (kN - 1)];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            row[(t + 256 * k3 + kN / 2) & (kN - 1)] = cur[k3] + (double)acc[k3];
            acc[k3] = 0.0f;
        }
    };

    const uint8_t *my_bytes = bytes + ((size_t)p0 * 2 + g) * kSpecBytes;     // + j * 2 * kSpecBytes
    auto issue = [&](int j) {                        // thread 0 of the group
        if (j >= nvalid) return;
        const int slot = j % kStagesPP;
        mbar_expect_tx(bars + 8 * slot, kSpecBytes);
        tma_load_1d(ring + slot * kSpecBytes, my_bytes + (size_t)j * 2 * kSpecBytes, kSpecBytes, bars + 8 * slot);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2 * kStagesPP; i++) mbar_init(sm + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0)
        for (int j = 0; j < kStagesPP && j < nit; j++) issue(j);
    if (g == 1) nb_arrive(3, kBlockPP);             // group 0 moves first

    c32 v[16];
    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const int slot = j % kStagesPP;
        const bool valid = j < nvalid;
        uint32_t raw[16];
        // ---------------- memory phase 0: raw bytes --------------------------------------------------
        if (valid) {
            mbar_wait(bars + 8 * slot, (uint32_t)((j / kStagesPP) & 1));
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++)
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw[n1]) : "r"(ring + slot * kSpecBytes + 2 * t + n1 * 512));
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 16; n1++) raw[n1] = 0x7f7fu;
        }
        // ---------------- math phase 1 ---------------------------------------------------------------
        PP_T(c0);
        nb_sync(MY_TURN, kBlockPP);
        PP_T(c1);
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            c32 x = unpack_biased(raw[n1]);
            if (WINDOW) {        // (u8 - 127.0f) * w, one rounding per rail: src/datastore.cxx:75-77
                x = __fadd2_rn(x, make_float2(-kBias, -kBias));
                x = __fmul2_rn(x, make_float2(win[n1], win[n1]));
            }
            v[n1] = x;
        }
        Dft<16>::run(v);
        if (!WINDOW) v[0] = __fadd2_rn(v[0], make_float2(-16.0f * kBias, -16.0f * kBias));
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw1[k - 1]);
        pin16(v);                                    // math stays inside the turn
        nb_arrive(OTHER_TURN, kBlockPP);
        PP_T(c2);
        // ---------------- memory phase 1: exchange through S1 -------------------------------------
#pragma unroll
        for (int u = 0; u < 8; u++)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(s1_st ^ (u << 4)), "f"(v[2 * u].x), "f"(v[2 * u].y),
                         "f"(v[2 * u + 1].x), "f"(v[2 * u + 1].y) : "memory");
        nb_sync(GROUP_BAR, kGroupThreads);           // S1 complete; ring slot consumed
        if (t == 0 && j + kStagesPP < nit) issue(j + kStagesPP);
#pragma unroll
        for (int n2 = 0; n2 < 16; n2++)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[n2].x), "=f"(v[n2].y) : "r"(s1_ld + n2 * 2048));
        // ---------------- math phase 2 ---------------------------------------------------------------
        PP_T(c3);
        nb_sync(MY_TURN, kBlockPP);
        PP_T(c4);
        pin16(v);                                    // ... and does not start before the turn is ours
        Dft<16>::run(v);
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], c_tw2_pp[hi][k]);
        pin16(v);
        nb_arrive(OTHER_TURN, kBlockPP);
        PP_T(c5);
        PP_ADD(0, c0, c1); PP_ADD(1, c1, c2); PP_ADD(2, c2, c3); PP_ADD(0, c3, c4); PP_ADD(1, c4, c5);
        if ((tid & 31) == 0) { PP_ADD(3, 0, 2); }
        // ---------------- memory phase 2: exchange through S2 -------------------------------------
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++)
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s2_st + k2 * 2048), "f"(v[k2].x), "f"(v[k2].y) : "memory");
        nb_sync(GROUP_BAR, kGroupThreads);           // S2 complete; S1 free
#pragma unroll
        for (int m = 0; m < 16; m++)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[m].x), "=f"(v[m].y) : "r"(s2_ld + m * 128));
        // ---------------- math phase 3 ---------------------------------------------------------------
        nb_sync(MY_TURN, kBlockPP);
        pin16(v);
        Dft<16>::run(v);
        if (valid) {
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) acc[k3] = fmaf(v[k3].x, v[k3].x, fmaf(v[k3].y, v[k3].y, acc[k3]));
        }
        asm volatile("" : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3]), "+f"(acc[4]), "+f"(acc[5]), "+f"(acc[6]),
                          "+f"(acc[7]), "+f"(acc[8]), "+f"(acc[9]), "+f"(acc[10]), "+f"(acc[11]), "+f"(acc[12]), "+f"(acc[13]),
                          "+f"(acc[14]), "+f"(acc[15]) :: "memory");
        nb_arrive(OTHER_TURN, kBlockPP);
        // S2 reads of this iteration precede the next iteration's S2 writes by one GROUP_BAR (after S1 stores)
        if (++since_flush == kFlushEveryPP) { flush(); since_flush = 0; }
    }
    flush();
    if (g == 0) nb_sync(3, kBlockPP);                // absorb group 1's final token so the barrier ends balanced
}

}  // namespace

#ifdef RPF_PP_TIMING
extern "C" int rpf_debug_pp_timing(unsigned long long *out, int reset)
{
    if (cudaMemcpyFromSymbol(out, pp_timing, sizeof(unsigned long long) * 4) != cudaSuccess) return -1;
    if (reset) { unsigned long long z[4] = {0, 0, 0, 0}; cudaMemcpyToSymbol(pp_timing, z, sizeof(z)); }
    return 0;
}
#endif

cudaError_t rpf_pingpong4096_configure()
{
    float2 h[16][16];
    const double PI = 3.14159265358979323846264338327950288;
    for (int a = 0; a < 16; a++)
        for (int b = 0; b < 16; b++) {
            const double ang = -2.0 * PI * (double)((a * b) & 255) / 256.0;
            h[a][b] = make_float2((float)cos(ang), (float)sin(ang));
        }
    cudaError_t e = cudaMemcpyToSymbol(c_tw2_pp, h, sizeof(h));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pingpong4096_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPP);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pingpong4096_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemPP);
}

cudaError_t rpf_launch_pingpong4096(const RpfLaunch &L)
{
    const int64_t npairs = (L.nspec + 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    dim3 grid((unsigned)(npairs < L.grid ? npairs : L.grid)), block(kBlockPP);
    if (L.window)
        pingpong4096_kernel<true><<<grid, block, kSmemPP, L.stream>>>(L.bytes, L.nspec, L.twiddle, L.window, L.rows);
    else
        pingpong4096_kernel<false><<<grid, block, kSmemPP, L.stream>>>(L.bytes, L.nspec, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}
