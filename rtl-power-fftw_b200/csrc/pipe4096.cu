// pipe4096.cu -- N = 4096, two-stream software pipeline (the headline kernel).
//
// 16 x 16 x 16 decomposition, 16 values per thread, packed FP32.  The butterflies keep the FP32 pipe ~95 %
// busy when nothing interferes (tools/dftbench.cu), and the staging + exchange traffic alone would run at
// > 28 % of the HBM roofline, but a CTA whose warps all move through load -> butterfly -> store -> barrier
// in lock-step serialises the two.  Here every compute thread carries TWO spectra (streams A and B, 16
// values each) and alternates between them:
//
//     stage(Y,p):  radix-16 on Y
//                  wait(mbarrier of X's next pass);  issue X's 16 loads      <- land under Y's tail
//                  twiddle;  store Y;  arrive(mbarrier of Y's pass)
//
// with the stage order (A,1)(B,1)(A,2)(B,2)(A,3)(B,3).  Barriers are split mbarrier arrive/wait pairs a
// whole butterfly apart in program order, so no warp ever sits in bar.sync; every stage of every
// iteration is unconditional (ring slots without a spectrum of their own re-read the last valid one),
// so the loop has no data-dependent branches.  One CTA per SM: 256 compute threads (both twiddle sets in
// registers, no spills) plus a third warp group, one lane of which issues the TMA bulk copies.
//
// Replaces reference src/datastore.cxx:68-87 for N = 4096.
#include <stdint.h>
#include <type_traits>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;
constexpr int kThreads = 256;            // compute threads (8 warps)
// RPF_TMAWARP = 1: a third warp group is launched only so that one of its lanes can issue the bulk copies; the loop of
// the compute warps then contains no divergent issue code at all.  The copy lane is a registered participant of the
// two pass-1 exchange barriers (it arrives for phase j as soon as it has seen phase j-1 complete, a whole iteration
// before the compute warps), so it can neither miss a phase nor hold anybody up.  Registers: launched at 384 x 168,
// the copy group drops to 24 and the compute groups take 208 (setmaxnreg).
#ifndef RPF_TMAWARP
#define RPF_TMAWARP 1
#endif
constexpr int kLaunchThreads = kThreads + (RPF_TMAWARP ? 128 : 0);
#ifndef RPF_FLUSH_PAIRS
#define RPF_FLUSH_PAIRS 64
#endif
#ifndef RPF_PIPE_STAGES
#define RPF_PIPE_STAGES 2             // measured (round-1 start): 2 -> 587, 3 -> 578 (modulo-3 slot arithmetic), 4 -> 582 Gsamples/s
#endif
constexpr int kStagesP = RPF_PIPE_STAGES;         // TMA ring depth per stream
constexpr int kSpecBytes = 2 * kN;                // 8 KiB
constexpr int kFlushPairs = RPF_FLUSH_PAIRS;      // flush every 64 iterations = 128 spectra
constexpr int kHdrP = 128;
// Shared-memory plan: header (mbarriers) | TMA ring [stream][slot] | S1 [stream][32 KiB] | S2 [stream][32 KiB].
// Raw-byte spectra: kStagesP slots of 8 KiB per stream.  Complex64 rows (four-step inner transform): one slot
// of 32 KiB per stream -- the slot is refilled as soon as every warp has stored its pass-1 rows, four
// stages (~1.2 us) before the next row is needed, which covers the HBM latency.
template <bool C32>
struct Plan {
    static constexpr int kStg = C32 ? 1 : kStagesP;            // ring depth per stream
    static constexpr int kSlot = C32 ? kN * 8 : kSpecBytes;     // bytes per ring slot
    static constexpr int kS1 = kHdrP + 2 * kStg * kSlot;
    static constexpr int kS2 = kS1 + 2 * kN * 8;
    static constexpr size_t kSmem = (size_t)kS2 + 2 * kN * 8;
    // barrier indices inside the header
    static __device__ __forceinline__ constexpr int bar_full(int s, int slot) { return s * kStg + slot; }
    static __device__ __forceinline__ constexpr int bar_rd(int s, int e) { return 2 * kStg + s * 2 + (e - 1); }
};

__device__ __forceinline__ void lds16x16(uint32_t base, uint32_t (&raw)[16])
{
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw[n1]) : "r"(base + n1 * 512));
}
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

// Stage scheduling knobs (measured on B200):
//   RPF_LOADPOS  where, inside stage (Y,p), the wait for X's barrier and X's 16 loads sit:
//                0 after the first half of Y's radix-16, 1 after three quarters, 2 after all of it,
//                3 after the twiddle multiplications (just before Y's stores)
#ifndef RPF_LOADPOS
#define RPF_LOADPOS 2
#endif
// per fetch kind (raw bytes / S1 rows / S2 rows); default: all the same
#ifndef RPF_LOADPOS_RAW
#define RPF_LOADPOS_RAW 3
#endif
#ifndef RPF_LOADPOS_S1
#define RPF_LOADPOS_S1 RPF_LOADPOS
#endif
#ifndef RPF_LOADPOS_S2
#define RPF_LOADPOS_S2 RPF_LOADPOS
#endif

// SRC_C32 = false: spectra come as raw bytes (TMA ring), bins are flushed fftshifted.
// SRC_C32 = true : the inner transform of the four-step path (N = M * 4096).  Input rows are complex64
//   scratch[spec][ko][4096]; CTA c is bound to ONE outer bin ko = c % M for its whole life and to the
//   spectra spec = part, part + P, ... (part = c / M, P = gridDim / M) of the batch, so its private
//   accumulator row is just 4096 doubles in natural order k (L2 resident, coalesced);
//   reduce_rows_fourstep() maps (ko, k) -> bin ko + M k.  The complex rows travel through the same TMA ring
//   as the raw bytes (one 32 KiB bulk copy per row), so the loop is the same for both sources; only the
//   pass-1 input differs (LDS.64 of complex values instead of LDS.U16 + unpack).
//
// Q > 0: N = 4096 >> Q bins (2048, 1024, 512).  A ring slot then holds SPS = 2^Q consecutive spectra, which the same
//   loop transforms together: with n = 256 a + 16 b + c, a = R1 p + a' (R1 = 16 >> Q), spectrum p of the slot is a
//   R1 x 16 x 16 transform -- pass 1 becomes SPS radix-R1 butterflies over a' per thread (twiddles W_N^{t k1'}, the same
//   for every p), row t of S1 holds (p, k1') in place of k1, passes 2 and 3 and both exchanges are UNCHANGED, and thread
//   (k1 = (p, k1'), k2) ends up with the bins k1' + R1 k2 + 16 R1 k3 of spectrum p: one private accumulator row per
//   (CTA, p).  `nspec` counts spectra of N bins; the stream's last slot may be partial (fewer bytes are fetched and the
//   threads of the missing p do not accumulate).  Q > 4 (N = 128 ... 16) repeats the step on the second digit: pass 1
//   is empty, b = R2 pb + b' (R2 = 16 >> (Q - 4)), pass 2 becomes 16 / R2 radix-R2 butterflies with twiddles
//   W_N^{c k2'}, and thread (k1, k2 = (pb, k2')) holds bins k2' + R2 k3 of spectrum (16 / R2) k1 + pb.
template <bool WINDOW, bool SRC_C32, int Q = 0>
__global__ void __launch_bounds__(kLaunchThreads, 1)
pipe4096_kernel(const uint8_t *__restrict__ bytes, const float2 *__restrict__ cplx, int64_t nspec, int M,
                const float2 *__restrict__ twiddle, const float *__restrict__ window, double *__restrict__ rows)
{
    static_assert(Q == 0 || !SRC_C32, "batched small transforms read raw bytes");
    static_assert(Q >= 0 && Q <= 8, "N = 4096 >> Q, 16 <= N <= 4096");
    constexpr int SPS = 1 << Q, NB = kN >> Q;                      // spectra per slot, bins
    constexpr int R1 = Q < 4 ? 16 >> Q : 1;                         // radix of pass 1 (over a')
    constexpr int R2 = Q <= 4 ? 16 : 16 >> (Q - 4);                 // radix of pass 2 (over b'); pass 3 is always radix 16
    constexpr int L1 = Q < 4 ? 4 - Q : 0, L2 = Q <= 4 ? 4 : 8 - Q;  // log2 R1, log2 R2
    using P = Plan<SRC_C32>;
    constexpr int kStg = P::kStg, kSlot = P::kSlot;
    auto bar_full = [](int s, int slot) { return P::bar_full(s, slot); };
    auto bar_rd = [](int s, int e) { return P::bar_rd(s, e); };
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int lo = tid & 15, hi = tid >> 4;       // pass 2: (k1 = lo, n3 = hi); pass 3: (k1 = lo, k2 = hi)

    // ---- work range: spectrum pairs ----------------------------------------------------
    int64_t p0 = 0;
    int nit;
    bool lastB_missing;
    const int64_t nslots = SRC_C32 ? nspec : (nspec + SPS - 1) / SPS;   // 4096-sample slots in the stream
    const int ko = SRC_C32 ? (int)(blockIdx.x % M) : 0;
    const int part = SRC_C32 ? (int)(blockIdx.x / M) : 0, nparts = SRC_C32 ? (int)(gridDim.x / M) : 1;
    if (SRC_C32) {
        // my spectra of the batch: part, part + nparts, ...  (count = mine)
        const int mine = (int)((nspec - part + nparts - 1) / nparts);
        nit = (mine + 1) / 2;
        lastB_missing = (mine & 1) != 0;
    } else {
        const int64_t npairs = (nslots + 1) / 2;
        p0 = (npairs * (int64_t)blockIdx.x) / gridDim.x;
        nit = (int)((npairs * (int64_t)(blockIdx.x + 1)) / gridDim.x - p0);
        lastB_missing = (p0 + nit) * 2 > nslots;     // stream B of the last pair of the stream may not exist
    }
    if (nit <= 0) return;
    // Q > 0: the stream's last slot may hold fewer than SPS spectra
    const int tail_count = (int)(nspec - (nslots - 1) * SPS);            // spectra in the stream's last slot (SPS if full)
    const bool tail_mine = Q > 0 && tail_count < SPS && (p0 + nit) * 2 >= nslots;   // ... and that slot is one of mine

    const uint32_t ring = sm + kHdrP;                            // [stream][slot]
    // ---- TMA: one bulk copy (8 KiB of bytes / 32 KiB complex row) per (stream, iteration) ------
    const uint8_t *cta_bytes = SRC_C32 ? nullptr : bytes + (size_t)p0 * 2 * kSpecBytes;
    // Ring slots are filled for A: iterations 0..nit (one past the end), B: 0..nit-1, so that every stage of
    // every iteration waits and loads unconditionally; slots that have no spectrum of their own (B of an odd
    // tail, A of iteration nit) re-read the last valid spectrum and their result is never accumulated.
    const int last_valid = 2 * nit - 1 - (lastB_missing ? 1 : 0);
    auto issue = [&](int s, int j) {                 // one lane only (the copy lane); my spectrum / row number 2 j + s
        const int slot = j % kStg;
        const uint32_t bar = sm + 8 * bar_full(s, slot);
        const int idx = min(2 * j + s, last_valid);
        const void *src = SRC_C32 ? (const void *)(cplx + ((size_t)((int64_t)part + (int64_t)nparts * idx) * M + ko) * (size_t)kN)
                                  : (const void *)(cta_bytes + (size_t)idx * kSpecBytes);
        // no proxy fence: the slot was only read through the generic proxy and those loads have completed
        const uint32_t nbytes = (tail_mine && idx == last_valid) ? (uint32_t)(tail_count * 2 * NB) : (uint32_t)kSlot;
        mbar_expect_tx(bar, nbytes);
        tma_load_1d(ring + (s * kStg + slot) * kSlot, src, nbytes, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2 * kStg; i++) mbar_init(sm + 8 * i, 1);
#pragma unroll
        for (int i = 0; i < 4; i++) mbar_init(sm + 8 * (2 * kStg + i), kThreads / 32 + ((RPF_TMAWARP && (i & 1) == 0) ? 1 : 0));   // index 2 s + e - 1
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
#if RPF_TMAWARP
    if (tid >= kThreads) {                             // the copy warp group: one lane works
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == kThreads) {
            for (int j = 0; j < kStg; j++) {
                if (j <= nit) issue(0, j);
                if (j < nit) issue(1, j);
            }
            for (int j = 0; j < nit; j++) {
                const uint32_t ph = (uint32_t)(j & 1);
#ifdef RPF_COPYLANE_HINT_NS
#define RPF_LANE_WAIT(bar, par) mbar_wait_parked((bar), (par), RPF_COPYLANE_HINT_NS)
#else
#define RPF_LANE_WAIT(bar, par) mbar_wait((bar), (par))
#endif
                mbar_arrive(sm + 8 * bar_rd(0, 1));
                RPF_LANE_WAIT(sm + 8 * bar_rd(0, 1), ph);              // every compute warp has stored its (A,1) rows
                if (j + kStg <= nit) issue(0, j + kStg);               // ... so A's ring slot is free
                mbar_arrive(sm + 8 * bar_rd(1, 1));
                RPF_LANE_WAIT(sm + 8 * bar_rd(1, 1), ph);
                if (j + kStg < nit) issue(1, j + kStg);
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#else
    if (tid == 0)
        for (int j = 0; j < kStg; j++) {
            if (j <= nit) issue(0, j);
            if (j < nit) issue(1, j);
        }
#endif

    // ---- per-thread constants ----------------------------------------------------------------
    float2 tw1[15], tw2[15];
#pragma unroll
    for (int k = 1; k < 16; k++) {
        tw1[k - 1] = twiddle[(tid * k * SPS) & (kN - 1)];       // W_N^{t k1'}, t = tid (only k < R1 is used when Q > 0)
        tw2[k - 1] = twiddle[(hi * k * (kN / (16 * R2))) & (kN - 1)];   // W_{16 R2}^{n3 k2'}, n3 = hi (W_256^{n3 k2} unless Q > 4)
    }
    float win[WINDOW ? 16 : 1];
    if (WINDOW) {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) win[n1] = window[(n1 * 256 + (tid & (kThreads - 1))) & (NB - 1)];   // w[n mod N]
    }
    // RPF_ACC_F64 (experiment): |X|^2 accumulated on the FP64 pipe (2 F2F + 2 DFMA per bin instead of 2 FFMA), no
    // periodic flush
#ifdef RPF_ACC_F64
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0;
#else
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;
#endif

    // shared-memory addresses (stream B = +kSpecBytes in the ring, +32 KiB in S1/S2)
    const uint32_t raw_off = 2 * tid;                            // + n1 * 512
    const uint32_t s1_st = sm + P::kS1 + tid * 128 + ((tid & 7) << 4);
    const uint32_t s1_ld = sm + P::kS1 + hi * 128 + ((((lo >> 1) ^ (hi & 7)) << 4) | ((lo & 1) << 3));   // + n2 * 2048
    const uint32_t s2_st = sm + P::kS2 + hi * 128 + lo * 8;                                              // + k2 * 2048
    const uint32_t s2_ld = sm + P::kS2 + hi * 2048 + lo * 8;                                             // + n3 * 128
    constexpr uint32_t kStreamS = kN * 8;                        // S1/S2 stream stride

    // my spectrum of the slot and my first bin: k = k1' + R1 (k2' + R2 k3)   (Q = 0: tid + 256 k3)
    const int myp = ((lo >> L1) << (4 - L2)) + (hi >> L2), kbase = (lo & (R1 - 1)) + R1 * (hi & (R2 - 1));
    double *row = rows + ((size_t)blockIdx.x * SPS + myp) * NB;
    constexpr int kShift = SRC_C32 ? 0 : NB / 2;     // bytes path: fftshift here; four-step: in the reduce kernel
    auto flush = [&]() {
        double cur[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) cur[k3] = row[(kbase + R2 * R1 * k3 + kShift) & (NB - 1)];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            row[(kbase + R2 * R1 * k3 + kShift) & (NB - 1)] = cur[k3] + (double)acc[k3];
            acc[k3] = 0;
        }
    };
    const bool tail_skip = tail_mine && myp >= tail_count;      // my spectrum of the stream's last slot does not exist
    c32 vA[16], vB[16];
    uint32_t raw[16];

    auto unpack = [&](c32 (&v)[16]) {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) {
            c32 x = unpack_biased(raw[n1]);
            // window: u8 - 127.0f here (exact), the multiplication by w (src/datastore.cxx:76-77) is fused into the
            // first radix-4 layer of pass 1 (dft16_first_windowed)
            if (WINDOW) x = __fadd2_rn(x, make_float2(-kBias, -kBias));
            if (WINDOW && Q > 0) x = __fmul2_rn(x, make_float2(win[n1], win[n1]));   // (Q = 0: fused into the first radix-4 layer)
            v[n1] = x;
        }
    };
    auto post1 = [&](c32 (&v)[16], int s, auto &&late) {   // bias fix, W_4096^{t k1}, store row t of S1[s], arrive
        if (!WINDOW && !SRC_C32) {
#pragma unroll
            for (int p = 0; p < 16 / R1; p++) v[p * R1] = __fadd2_rn(v[p * R1], make_float2(-(float)R1 * kBias, -(float)R1 * kBias));
        }
#pragma unroll
        for (int k = 1; k < 16; k++)
            if ((k & (R1 - 1)) != 0) v[k] = cmul(v[k], tw1[(k & (R1 - 1)) - 1]);
        late();
        sts_s1(s1_st + s * kStreamS, v);
        warp_arrive(sm + 8 * bar_rd(s, 1));
    };
    auto post2 = [&](c32 (&v)[16], int s, auto &&late) {   // W_256^{n3 k2}, store S2[s], arrive
#pragma unroll
        for (int k = 1; k < 16; k++)
            if ((k & (R2 - 1)) != 0) v[k] = cmul(v[k], tw2[(k & (R2 - 1)) - 1]);
        late();
        sts_s2(s2_st + s * kStreamS, v);
        warp_arrive(sm + 8 * bar_rd(s, 2));
    };
    auto power = [&](const c32 (&v)[16]) {
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
#ifdef RPF_ACC_F64
            const double x = (double)v[k3].x, y = (double)v[k3].y;
            acc[k3] = fma(x, x, fma(y, y, acc[k3]));
#else
            acc[k3] = fmaf(v[k3].x, v[k3].x, fmaf(v[k3].y, v[k3].y, acc[k3]));
#endif
        }
    };

    // pass-1 input of one spectrum from its ring slot: raw bytes (unpacked at the top of the consuming stage)
    // or the complex row straight into the idle stream's registers
    auto load_input = [&](uint32_t slot_addr, c32 (&v)[16]) {
        if (SRC_C32) lds64x16<2048>(slot_addr + tid * 8, v);
        else lds16x16(slot_addr + raw_off, raw);
    };
    // prologue: input of stream A, iteration 0
    mbar_wait(sm + 8 * bar_full(0, 0), 0);
    load_input(ring, vA);

    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const int slot = j % kStg;
        const uint32_t ph = (uint32_t)(j & 1);
        const bool haveB = !(lastB_missing && j == nit - 1);

        // stage (Y,p): radix-16 on Y; `fetch` = wait for X's barrier + X's 16 loads, placed by RPF_LOADPOS
        auto none = []() {};
        // RPF_TMAWARP = 0 only: thread 0 issues the bulk copies inside the loop (a warp-uniform branch with elect.sync
        // was measured 14 % slower)
        auto issuer = [&](bool cond) -> bool { return !RPF_TMAWARP && tid == 0 && cond; };
        auto radix = [&](c32 (&vY)[16], auto pos, auto &&fetch, int pass = 3) {
            constexpr int kPos = decltype(pos)::value;
            const bool pass1 = pass == 1;
            if (Q > 0 && pass == 1) {                // 16 / R1 independent radix-R1 butterflies over a'
#pragma unroll
                for (int p = 0; p < 16 / R1; p++) Dft<R1>::run(&vY[p * R1]);
                if (kPos <= 2) fetch();
                return;
            }
            if (Q > 4 && pass == 2) {                // 16 / R2 independent radix-R2 butterflies over b'
#pragma unroll
                for (int p = 0; p < 16 / R2; p++) Dft<R2>::run(&vY[p * R2]);
                if (kPos <= 2) fetch();
                return;
            }
            c32 y[16];
            if (WINDOW && !SRC_C32 && pass1) dft16_first_windowed(vY, win); else dft16_first(vY);
            if (kPos == 0) fetch();
            dft16_second_a(vY, y);
            if (kPos == 1) fetch();
            dft16_second_b(vY, y);
            if (kPos == 2) fetch();
        };
        using PosRaw = std::integral_constant<int, RPF_LOADPOS_RAW>;
        using PosS1 = std::integral_constant<int, RPF_LOADPOS_S1>;
        using PosS2 = std::integral_constant<int, RPF_LOADPOS_S2>;
        const uint32_t rd = sm + 8 * bar_rd(0, 1);       // bar_rd(s, e) = rd + 8 (2 s + e - 1)
        auto fetch_rawB = [&]() {
            const uint32_t bar = sm + 8 * bar_full(1, slot), par = (uint32_t)((j / kStg) & 1);
            mbar_wait(bar, par);
            load_input(ring + (kStg + slot) * kSlot, vB);
        };
        auto fetch_s1A = [&]() {
            mbar_wait(rd, ph);                               // every warp has stored its (A,1) rows ...
            if (issuer(j + kStg <= nit)) issue(0, j + kStg);         // ... so A's ring slot is free
            lds64x16<2048>(s1_ld, vA);
        };
        auto fetch_s1B = [&]() {
            mbar_wait(rd + 16, ph);
            if (issuer(j + kStg < nit)) issue(1, j + kStg);
            lds64x16<2048>(s1_ld + kStreamS, vB);
        };
        auto fetch_s2A = [&]() { mbar_wait(rd + 8, ph); lds64x16<128>(s2_ld, vA); };
        auto fetch_s2B = [&]() { mbar_wait(rd + 24, ph); lds64x16<128>(s2_ld + kStreamS, vB); };
        auto fetch_rawA = [&]() {
            const uint32_t bar = sm + 8 * bar_full(0, (j + 1) % kStg), par = (uint32_t)(((j + 1) / kStg) & 1);
            mbar_wait(bar, par);
            load_input(ring + ((j + 1) % kStg) * kSlot, vA);
        };

        // ---- (A,1)  | prefetch raw B ---------------------------------------------------------
        if (!SRC_C32) unpack(vA);
        radix(vA, PosRaw(), fetch_rawB, 1);
        if (PosRaw::value == 3) post1(vA, 0, fetch_rawB); else post1(vA, 0, none);

        // ---- (B,1)  | prefetch S1[A] ----------------------------------------------------------
        if (!SRC_C32) unpack(vB);
        radix(vB, PosS1(), fetch_s1A, 1);
        if (PosS1::value == 3) post1(vB, 1, fetch_s1A); else post1(vB, 1, none);

        // ---- (A,2)  | prefetch S1[B] ----------------------------------------------------------
        radix(vA, PosS1(), fetch_s1B, 2);
        if (PosS1::value == 3) post2(vA, 0, fetch_s1B); else post2(vA, 0, none);

        // ---- (B,2)  | prefetch S2[A] ----------------------------------------------------------
        radix(vB, PosS2(), fetch_s2A, 2);
        if (PosS2::value == 3) post2(vB, 1, fetch_s2A); else post2(vB, 1, none);

        // ---- (A,3)  | prefetch S2[B] ----------------------------------------------------------
        radix(vA, PosS2(), fetch_s2B);
        if (PosS2::value == 3) fetch_s2B();
        if (Q == 0 || !(tail_skip && j == nit - 1 && lastB_missing)) power(vA);      // (A is the stream's last slot only then)

        // ---- (B,3)  | prefetch raw A of the next iteration ------------------------------------
        radix(vB, PosRaw(), fetch_rawA);
        if (PosRaw::value == 3) fetch_rawA();
        if (haveB && (Q == 0 || !(tail_skip && j == nit - 1))) power(vB);

#ifndef RPF_ACC_F64
        if (++since_flush == kFlushPairs) { flush(); since_flush = 0; }
#endif
    }
    flush();
}

}  // namespace

cudaError_t rpf_pipe4096_configure()
{
    cudaError_t e = cudaFuncSetAttribute(pipe4096_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<false>::kSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pipe4096_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<false>::kSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(pipe4096_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<true>::kSmem);
    if (e != cudaSuccess) return e;
    const void *small[] = {(const void *)pipe4096_kernel<false, false, 1>, (const void *)pipe4096_kernel<true, false, 1>,
                           (const void *)pipe4096_kernel<false, false, 2>, (const void *)pipe4096_kernel<true, false, 2>,
                           (const void *)pipe4096_kernel<false, false, 3>, (const void *)pipe4096_kernel<true, false, 3>,
                           (const void *)pipe4096_kernel<false, false, 4>, (const void *)pipe4096_kernel<true, false, 4>,
                           (const void *)pipe4096_kernel<false, false, 5>, (const void *)pipe4096_kernel<true, false, 5>,
                           (const void *)pipe4096_kernel<false, false, 6>, (const void *)pipe4096_kernel<true, false, 6>,
                           (const void *)pipe4096_kernel<false, false, 7>, (const void *)pipe4096_kernel<true, false, 7>,
                           (const void *)pipe4096_kernel<false, false, 8>, (const void *)pipe4096_kernel<true, false, 8>};
    for (const void *f : small) {
        e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Plan<false>::kSmem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// N = 2048 ... 16 on the same loop: 2 ... 256 spectra per 4096-sample ring slot (template parameter Q).
template <int Q>
static cudaError_t launch_small(const RpfLaunch &L)
{
    const int64_t nslots = (L.nspec + (1 << Q) - 1) >> Q, npairs = (nslots + 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    dim3 grid((unsigned)(npairs < L.grid ? npairs : L.grid)), block(kLaunchThreads);
    if (L.window)
        pipe4096_kernel<true, false, Q><<<grid, block, Plan<false>::kSmem, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, L.window, L.rows);
    else
        pipe4096_kernel<false, false, Q><<<grid, block, Plan<false>::kSmem, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}

bool rpf_pipe4096_small_supported(int N) { return N >= 16 && N <= 2048 && (N & (N - 1)) == 0; }

cudaError_t rpf_launch_pipe4096_small(int N, const RpfLaunch &L)
{
    switch (N) {
        case 2048: return launch_small<1>(L);
        case 1024: return launch_small<2>(L);
        case 512: return launch_small<3>(L);
        case 256: return launch_small<4>(L);       // 16 x 16: pass 1 is empty (R1 = 1)
        case 128: return launch_small<5>(L);       // 8 x 16
        case 64: return launch_small<6>(L);        // 4 x 16
        case 32: return launch_small<7>(L);        // 2 x 16
        case 16: return launch_small<8>(L);        // pass 3 only
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t rpf_launch_pipe4096(const RpfLaunch &L)
{
    if (L.cplx) {   // four-step inner transform: the grid is fixed (multiple of M) so CTA <-> outer bin is stable
        if (L.nspec <= 0) return cudaSuccess;
        pipe4096_kernel<false, true><<<L.grid, kLaunchThreads, Plan<true>::kSmem, L.stream>>>(nullptr, L.cplx, L.nspec, L.M, L.twiddle, nullptr, L.rows);
        return cudaGetLastError();
    }
    const int64_t npairs = (L.nspec + 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    dim3 grid((unsigned)(npairs < L.grid ? npairs : L.grid)), block(kLaunchThreads);
    if (L.window)
        pipe4096_kernel<true, false><<<grid, block, Plan<false>::kSmem, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, L.window, L.rows);
    else
        pipe4096_kernel<false, false><<<grid, block, Plan<false>::kSmem, L.stream>>>(L.bytes, nullptr, L.nspec, 1, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}
