// pipe4096_tc.cu -- N = 4096, the two-stream pipeline of pipe4096.cu with PASS 1 ON THE TENSOR CORES
// (experiment of VERDICT r1 task 6; selected with RPF_FLAG_TC_PASS1, raw-byte source without window only).
//
// Pass 1 of the 16 x 16 x 16 decomposition is, per spectrum, a 256 x 32 x 32 real GEMM on exact 8-bit
// integers: row t holds the 16 samples x[256 n1 + t] as (re, im) pairs, the 32 x 32 matrix is the DFT-16
// (cos / sin) and the 32 outputs of row t are thread t's 16 complex values.  Here
//   * the compute threads convert their raw bytes to fp16 (one PRMT builds 1024 + byte on both rails, one
//     HADD2 subtracts 1151: exact) and write row t of the A operand straight into TENSOR MEMORY (tcgen05.st,
//     16 columns of packed fp16 pairs per row; no shared-memory tile, hence no generic->async proxy fence),
//   * one lane of the copy warp group issues tcgen05.mma (kind::f16, A from TMEM, B from shared memory,
//     M = 128, N = 32, K = 16; fp32 accumulators in TMEM) -- the DFT-16 matrix is split hi + lo in fp16 (2 x 2 K-steps per 128 rows), so
//     the products carry ~22 bits -- and commits to an mbarrier,
//   * thread t reads its 32 accumulator columns back with tcgen05.ld (32x32b.x32) and continues with the
//     W_4096^{t k1} twiddles exactly as the FP32 kernel does.  Passes 2 and 3 are unchanged.
// That removes the 75 packed FP32 operations of the first radix-16 (of 286 per thread-spectrum).
//
// Replaces reference src/datastore.cxx:68-87 for N = 4096 (no window).
#include <cuda_fp16.h>
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;
constexpr int kThreads = 256;                 // compute threads (8 warps)
constexpr int kLaunchThreads = kThreads + 128;
constexpr int kStg = 2;                       // TMA ring depth per stream
constexpr int kSlot = 2 * kN;                 // 8 KiB of raw bytes per spectrum
constexpr int kFlushPairs = 64;
// shared-memory plan (bytes)
constexpr int kOffB = 1024;                   // DFT-16 matrix tiles: hi | lo, 2 KiB each
constexpr int kOffRing = kOffB + 4096;        // raw ring [stream][slot]
constexpr int kOffS1 = kOffRing + 2 * kStg * kSlot;
constexpr int kOffS2 = kOffS1 + 2 * kN * 8;
constexpr size_t kSmem = (size_t)kOffS2 + 2 * kN * 8;
// mbarriers (8 bytes each, at the start of shared memory)
__device__ __forceinline__ constexpr int bar_full(int s, int slot) { return s * kStg + slot; }          // 0..3
__device__ __forceinline__ constexpr int bar_rd(int s, int e) { return 4 + s * 2 + (e - 1); }           // 4..7
__device__ __forceinline__ constexpr int bar_afull(int s) { return 8 + s; }                             // 8, 9
__device__ __forceinline__ constexpr int bar_dready(int s) { return 10 + s; }                           // 10, 11
constexpr int kTmemSlotOff = 12 * 8;          // uint32: TMEM base address written by tcgen05.alloc
constexpr int kTmemCols = 256;                // D: [stream][half][32 columns] | A: 128 + [stream][half][16 columns]
constexpr int kTmemA = 128;

__constant__ float kCos16[16] = {1.0f, 0.923879532511286756f, 0.707106781186547524f, 0.382683432365089772f,
                                 0.0f, -0.382683432365089772f, -0.707106781186547524f, -0.923879532511286756f,
                                 -1.0f, -0.923879532511286756f, -0.707106781186547524f, -0.382683432365089772f,
                                 0.0f, 0.382683432365089772f, 0.707106781186547524f, 0.923879532511286756f};

template <int STRIDE>
__device__ __forceinline__ void lds64x16(uint32_t base, c32 (&v)[16])
{
#pragma unroll
    for (int i = 0; i < 16; i++)
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[i].x), "=f"(v[i].y) : "r"(base + i * STRIDE));
}
__device__ __forceinline__ void sts_s1(uint32_t base, const c32 (&v)[16])
{
#pragma unroll
    for (int u = 0; u < 8; u++)
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base ^ (u << 4)), "f"(v[2 * u].x), "f"(v[2 * u].y),
                     "f"(v[2 * u + 1].x), "f"(v[2 * u + 1].y) : "memory");
}
__device__ __forceinline__ void sts_s2(uint32_t base, const c32 (&v)[16])
{
#pragma unroll
    for (int k2 = 0; k2 < 16; k2++)
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(base + k2 * 2048), "f"(v[k2].x), "f"(v[k2].y) : "memory");
}
__device__ __forceinline__ void warp_arrive(uint32_t bar)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}

// K-major, no-swizzle UMMA shared-memory descriptor: core matrices of 8 rows x 16 bytes; `lbo` = byte
// distance of core matrices adjacent in K, `sbo` = adjacent in M/N (checked on hardware by tools/tc_probe.cu).
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t *r)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, c32 (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
          "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y),
          "=f"(v[8].x), "=f"(v[8].y), "=f"(v[9].x), "=f"(v[9].y), "=f"(v[10].x), "=f"(v[10].y), "=f"(v[11].x), "=f"(v[11].y),
          "=f"(v[12].x), "=f"(v[12].y), "=f"(v[13].x), "=f"(v[13].y), "=f"(v[14].x), "=f"(v[14].y), "=f"(v[15].x), "=f"(v[15].y)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__global__ void __launch_bounds__(kLaunchThreads, 1)
pipe4096_tc_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
                   double *__restrict__ rows)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int lo = tid & 15, hi = tid >> 4;       // pass 2: (k1 = lo, n3 = hi); pass 3: (k1 = lo, k2 = hi)

    // ---- work range: spectrum pairs ----------------------------------------------------
    const int64_t npairs = (nspec + 1) / 2;
    const int64_t p0 = (npairs * (int64_t)blockIdx.x) / gridDim.x;
    const int nit = (int)((npairs * (int64_t)(blockIdx.x + 1)) / gridDim.x - p0);
    const bool lastB_missing = (p0 + nit) * 2 > nspec;
    if (nit <= 0) return;

    const uint32_t ring = sm + kOffRing;
    const uint8_t *cta_bytes = bytes + (size_t)p0 * 2 * kSlot;
    // Ring slots are filled for generations 0..nit of BOTH streams (one past the end), so that every stage of
    // every iteration converts, multiplies and loads unconditionally; generations without a spectrum of their
    // own re-read the last valid one and their result is never accumulated.
    const int last_valid = 2 * nit - 1 - (lastB_missing ? 1 : 0);
    auto issue = [&](int s, int j) {                 // the async lane only
        const int slot = j % kStg;
        const uint32_t bar = sm + 8 * bar_full(s, slot);
        const int idx = min(2 * j + s, last_valid);
        mbar_expect_tx(bar, kSlot);
        tma_load_1d(ring + (s * kStg + slot) * kSlot, cta_bytes + (size_t)idx * kSlot, kSlot, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) mbar_init(sm + 8 * i, 1);                                   // full
#pragma unroll
        for (int i = 4; i < 8; i++) mbar_init(sm + 8 * i, kThreads / 32);                       // rd
        mbar_init(sm + 8 * bar_afull(0), kThreads / 32 + 1);                                    // + the async lane
        mbar_init(sm + 8 * bar_afull(1), kThreads / 32 + 1);
        mbar_init(sm + 8 * bar_dready(0), 1);                                                   // tcgen05.commit
        mbar_init(sm + 8 * bar_dready(1), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // DFT-16 as a real 32 x 32 matrix B[n = (k1, re|im)][k = (n1, re|im)], split hi + lo in fp16:
    //   out_re = sum xr c + xi s,  out_im = sum xi c - xr s,  c - i s = W_16^{n1 k1}
    for (int i = tid; i < 1024; i += kLaunchThreads) {
        const int n = i >> 5, k = i & 31;
        const int j16 = ((n >> 1) * (k >> 1)) & 15;
        const float c = kCos16[j16], s = kCos16[(j16 + 12) & 15];
        const float e = ((n ^ k) & 1) == 0 ? c : ((n & 1) ? -s : s);
        const __half eh = __float2half_rn(e);
        const __half el = __float2half_rn(e - __half2float(eh));
        const int off = ((n >> 3) * 4 + (k >> 3)) * 128 + (n & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__half *>(smem + kOffB + off) = eh;
        *reinterpret_cast<__half *>(smem + kOffB + 2048 + off) = el;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm + kTmemSlotOff), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + kTmemSlotOff);

    if (tid >= kThreads) {                             // the async warp group: one lane issues TMA and MMA
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (tid == kThreads) {
            for (int j = 0; j < kStg; j++) {           // generations 0..nit exist for both streams
                if (j <= nit) { issue(0, j); issue(1, j); }
            }
            constexpr uint32_t kIdesc = (1u << 4) | ((32u >> 3) << 17) | ((128u >> 4) << 24);   // f16 x f16 -> f32, M 128, N 32
            constexpr uint32_t kKdir = 128, kMdir = 512;
            for (int j = 0; j <= nit; j++) {
                const uint32_t ph = (uint32_t)(j & 1);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    mbar_arrive(sm + 8 * bar_afull(s));
                    mbar_wait(sm + 8 * bar_afull(s), ph);        // every warp has stored its rows of A(s, j) (and read D(s, j-1))
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const uint32_t a = tmem + kTmemA + s * 32 + h * 16;      // 8 columns = 16 fp16 of K per step
                        const uint32_t d = tmem + s * 64 + h * 32;
#pragma unroll
                        for (int q = 0; q < 4; q++) {            // (hi, k 0..15) (hi, k 16..31) (lo, ...) (lo, ...)
                            const uint64_t db = umma_desc(sm + kOffB + (q >> 1) * 2048 + (q & 1) * 2 * kKdir, kKdir, kMdir);
                            umma_f16_ts(d, a + (q & 1) * 8, db, kIdesc, q ? 1u : 0u);
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                                 ::"r"(sm + 8 * bar_dready(s)) : "memory");
                    if (j + kStg <= nit) issue(s, j + kStg);     // the raw slot of generation j has been consumed
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");

    // ---- per-thread constants ----------------------------------------------------------------
    float2 tw1[15], tw2[15];
#pragma unroll
    for (int k = 1; k < 16; k++) {
        tw1[k - 1] = twiddle[(tid * k) & (kN - 1)];             // W_4096^{t k1}
        tw2[k - 1] = twiddle[(hi * k * 16) & (kN - 1)];         // W_256^{n3 k2}
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;

    const uint32_t raw_off = 2 * tid;                            // + n1 * 512
    const uint32_t a_st = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + kTmemA + (tid >> 7) * 16;   // + s * 32 + part * 8
    const uint32_t d_ld = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (tid >> 7) * 32;           // + s * 64
    const uint32_t s1_st = sm + kOffS1 + tid * 128 + ((tid & 7) << 4);
    const uint32_t s1_ld = sm + kOffS1 + hi * 128 + ((((lo >> 1) ^ (hi & 7)) << 4) | ((lo & 1) << 3));   // + n2 * 2048
    const uint32_t s2_st = sm + kOffS2 + hi * 128 + lo * 8;                                              // + k2 * 2048
    const uint32_t s2_ld = sm + kOffS2 + hi * 2048 + lo * 8;                                             // + n3 * 128
    constexpr uint32_t kStreamS = kN * 8;

    double *row = rows + (size_t)blockIdx.x * kN;
    auto flush = [&]() {
        double cur[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) cur[k3] = row[(tid + 256 * k3 + kN / 2) & (kN - 1)];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            row[(tid + 256 * k3 + kN / 2) & (kN - 1)] = cur[k3] + (double)acc[k3];   // bin k = tid + 256 k3, fftshifted
            acc[k3] = 0.0f;
        }
    };
    c32 vA[16], vB[16];

    // Row t of A(s), generation g: raw bytes -> fp16 (byte - 127, exact) -> tensor memory (two tcgen05.st of 8
    // columns); conv_commit publishes it to the tensor core once the stores have completed -- issued a stage
    // later, so nobody waits for the store latency.
    auto conv_issue = [&](int s, int g) {
        const int slot = g % kStg;
        mbar_wait(sm + 8 * bar_full(s, slot), (uint32_t)((g / kStg) & 1));
        const uint32_t src = ring + (s * kStg + slot) * kSlot + raw_off;
        uint32_t r[16];
#pragma unroll
        for (int i = 0; i < 16; i++) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r[i]) : "r"(src + i * 512));
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const uint32_t p = __byte_perm(r[i], 0x64646464u, 0x7170);         // half2(1024 + I, 1024 + Q)
            asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r[i]) : "r"(p), "r"(0xE47FE47Fu));   // - 1151
        }
        tmem_st8(a_st + s * 32, r);
        tmem_st8(a_st + s * 32 + 8, r + 8);
    };
    auto conv_commit = [&](int s) {
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();                       // my tcgen05.st of A(s) and tcgen05.ld of D(s) precede the next MMA
        warp_arrive(sm + 8 * bar_afull(s));
    };
    auto post1 = [&](c32 (&v)[16], int s, auto &&late) {   // W_4096^{t k1}, store row t of S1[s], arrive
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw1[k - 1]);
        late();
        sts_s1(s1_st + s * kStreamS, v);
        warp_arrive(sm + 8 * bar_rd(s, 1));
    };
    auto post2 = [&](c32 (&v)[16], int s) {
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw2[k - 1]);
        sts_s2(s2_st + s * kStreamS, v);
        warp_arrive(sm + 8 * bar_rd(s, 2));
    };
    auto power = [&](const c32 (&v)[16]) {
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) acc[k3] = fmaf(v[k3].x, v[k3].x, fmaf(v[k3].y, v[k3].y, acc[k3]));
    };
    auto radix = [&](c32 (&v)[16], auto &&fetch) {
        c32 y[16];
        dft16_first(v);
        dft16_second_a(v, y);
        dft16_second_b(v, y);
        fetch();
    };

    // ---- prologue: generation 0 of both streams to the tensor core; D(A, 0) into registers -------
    conv_issue(0, 0); conv_commit(0);
    conv_issue(1, 0); conv_commit(1);
    mbar_wait(sm + 8 * bar_dready(0), 0);
    tc_fence_after();
    tmem_ld32(d_ld, vA);

    const uint32_t rd = sm + 8 * bar_rd(0, 1);           // bar_rd(s, e) = rd + 8 (2 s + e - 1)
    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const uint32_t ph = (uint32_t)(j & 1);
        const bool haveB = !(lastB_missing && j == nit - 1);

        // Stage order T1(A) C(A) T1(B) C(B) R2(A) R2(B) R3(A) R3(B): the conversion of generation j+1 follows the
        // twiddle stage of generation j directly (its operand and accumulator columns are free by then), so the
        // tensor core has passes 2 and 3 of both streams (~1100 issue cycles) to finish before D(j+1) is read; every
        // exchange barrier is waited for at least ~220 issue cycles after the thread's own arrive; tcgen05.st /
        // tcgen05.ld are issued a stage before their completion is needed.
        // ---- T1(A): D(A, j) is in vA | fetch D(B, j); C(A): raw A(j+1) -> tensor memory -----------------
        tmem_ld_wait();
        post1(vA, 0, [&]() {
            mbar_wait(sm + 8 * bar_dready(1), ph);
            tc_fence_after();
            tmem_ld32(d_ld + 64, vB);
        });
        conv_issue(0, j + 1);
        // ---- T1(B); C(B) | fetch S1[A] ---------------------------------------------------------
        tmem_ld_wait();
        post1(vB, 1, [&]() { conv_commit(0); });
        conv_issue(1, j + 1);
        mbar_wait(rd, ph);
        lds64x16<2048>(s1_ld, vA);
        // ---- R2(A) | fetch S1[B] ------------------------------------------------------------
        radix(vA, [&]() { conv_commit(1); mbar_wait(rd + 16, ph); lds64x16<2048>(s1_ld + kStreamS, vB); });
        post2(vA, 0);
        // ---- R2(B) | fetch S2[A] ------------------------------------------------------------
        radix(vB, [&]() { mbar_wait(rd + 8, ph); lds64x16<128>(s2_ld, vA); });
        post2(vB, 1);
        // ---- R3(A) | fetch S2[B]; fetch D(A, j+1) -----------------------------------------------
        radix(vA, [&]() { mbar_wait(rd + 24, ph); lds64x16<128>(s2_ld + kStreamS, vB); });
        power(vA);
        mbar_wait(sm + 8 * bar_dready(0), ph ^ 1u);
        tc_fence_after();
        tmem_ld32(d_ld, vA);
        // ---- R3(B) ---------------------------------------------------------------------------
        radix(vB, []() {});
        if (haveB) power(vB);

        if (++since_flush == kFlushPairs) { flush(); since_flush = 0; }
    }
    tmem_ld_wait();
    mbar_wait(sm + 8 * bar_dready(1), (uint32_t)(nit & 1));      // the last (unused) MMA into D(B) has finished
    flush();
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");  // the eight compute warps
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

cudaError_t rpf_pipe4096_tc_configure()
{
    return cudaFuncSetAttribute(pipe4096_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
}

cudaError_t rpf_launch_pipe4096_tc(const RpfLaunch &L)
{
    const int64_t npairs = (L.nspec + 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    dim3 grid((unsigned)(npairs < L.grid ? npairs : L.grid)), block(kLaunchThreads);
    pipe4096_tc_kernel<<<grid, block, kSmem, L.stream>>>(L.bytes, L.nspec, L.twiddle, L.rows);
    return cudaGetLastError();
}
