// pipen.cu -- N = 512, 1024, 2048: the two-stream mbarrier pipeline of pipe4096.cu on the generic 3-pass geometry
// of fused3.cu (VERDICT r1 task 5).
//
// N = R1 R2 R3 (8.8.8, 16.8.8, 8.16.16), 16 values per thread, T = N/16 threads per spectrum, so the 256 compute
// threads of a CTA carry G = 256/T spectra at a time (one "group iteration" = G consecutive spectra = 8 KiB of raw
// bytes for every N).  Every thread carries TWO group iterations (streams A and B) and alternates between them
//
//     stage(Y,p):  butterflies of Y's pass p
//                  wait(mbarrier of X's next pass);  issue X's 16 loads      <- land under Y's tail
//                  twiddle;  store Y;  arrive(mbarrier of Y's pass)
//
// in the order (A,1)(B,1)(A,2)(B,2)(A,3)(B,3): no bar.sync and no data-dependent branch in the loop, the exchange
// barriers are split arrive/wait pairs a whole butterfly apart, one lane of a third warp group issues the TMA bulk
// copies (setmaxnreg moves its registers to the compute warps, so the twiddle sets stay in registers: no spills,
// which the two-CTA-per-SM lock-step fused3 kernel could not avoid).  One CTA per SM.
//
// Replaces reference src/datastore.cxx:68-87 for these N.
#include <stdint.h>
#include <type_traits>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kThreads = 256;                 // compute threads (8 warps)
constexpr int kLaunchThreads = kThreads + 128;
#ifndef RPF_PIPEN_STAGES
#define RPF_PIPEN_STAGES 3             // measured (512 / 1024 / 2048 bins, Gsamples/s): 2 -> 751 / 620 / 547, 3 -> 774 / 620 / 574, 4 -> 762 / 612 / 564
#endif
constexpr int kStg = RPF_PIPEN_STAGES;        // TMA ring depth per stream
constexpr int kIterBytes = 8192;              // G * 2N for every supported N
constexpr int kFlushPairs = 64;               // flush every 64 loop iterations (= 128 G spectra per CTA)
constexpr int kHdr = 256;

template <int N> struct Radices;
template <> struct Radices<512>  { static constexpr int R1 = 8,  R2 = 8,  R3 = 8;  };
template <> struct Radices<1024> { static constexpr int R1 = 16, R2 = 8,  R3 = 8;  };
template <> struct Radices<2048> { static constexpr int R1 = 8,  R2 = 16, R3 = 16; };

template <int N>
struct Geo {
    static constexpr int R1 = Radices<N>::R1, R2 = Radices<N>::R2, R3 = Radices<N>::R3;
    static constexpr int T = N / 16;              // threads per spectrum
    static constexpr int G = kThreads / T;        // spectra per group iteration
    static constexpr int C1 = 16 / R1, C2 = 16 / R2, C3 = 16 / R3;   // columns per thread
    static constexpr int NC1 = N / R1;            // pass-1 columns (= stride of the first digit)
    static constexpr int P1 = NC1 + (R3 % 16);    // S1 row pitch (bank-conflict padding)
    static constexpr int S1_ELEMS = R1 * P1;      // per spectrum
    static constexpr int P2 = R3 + 1;             // S2 row pitch: rows are pass-3 columns u = k1*R2 + k2
    static constexpr int S2_ELEMS = (N / R3) * P2;
    static constexpr int kS1Bytes = G * S1_ELEMS * 8, kS2Bytes = G * S2_ELEMS * 8;   // per stream
    // shared memory: header | ring [stream][slot] | S1 [stream] | S2 [stream]
    static constexpr int kOffS1 = kHdr + 2 * kStg * kIterBytes;
    static constexpr int kOffS2 = kOffS1 + 2 * kS1Bytes;
    static constexpr size_t kSmem = (size_t)kOffS2 + 2 * kS2Bytes;
    static_assert(G * 2 * N == kIterBytes, "8 KiB of raw bytes per group iteration");
    static_assert(R1 * R2 * R3 == N, "radix schedule");
};
// mbarriers: full[s][slot] (TMA landed) | empty[s][slot] (raw slot consumed: all 8 warps, once per use of the slot) |
// rd[s][e][group] (exchange buffer e of stream s of one spectrum group stored: the T/32 warps of that group only --
// spectra are independent, so a group never waits for another one; with T = 32 the exchange is warp-local and
// needs no mbarrier at all)
__device__ __forceinline__ constexpr int bar_full(int s, int slot) { return s * kStg + slot; }
__device__ __forceinline__ constexpr int bar_empty(int s, int slot) { return 2 * kStg + s * kStg + slot; }
template <int G> __device__ __forceinline__ constexpr int bar_rd(int s, int e, int grp) { return 4 * kStg + ((s * 2 + (e - 1)) * G + grp); }
static_assert(4 * kStg + 4 * 4 <= kHdr / 8, "mbarriers fit the header (G <= 4 has exchange barriers)");

__device__ __forceinline__ void warp_arrive(uint32_t bar)
{
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar);
}
__device__ __forceinline__ c32 lds64(uint32_t addr)
{
    c32 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, c32 v)
{
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y) : "memory");
}

template <int N, bool WINDOW>
__global__ void __launch_bounds__(kLaunchThreads, 1)
pipen_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
             const float *__restrict__ window, double *__restrict__ rows)
{
    using g = Geo<N>;
    constexpr int R1 = g::R1, R2 = g::R2, R3 = g::R3, T = g::T, G = g::G;
    constexpr int C1 = g::C1, C2 = g::C2, C3 = g::C3, NC1 = g::NC1, P1 = g::P1, P2 = g::P2;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int grp = (tid & (kThreads - 1)) / T;     // which of the G spectra of a group iteration
    const int tau = tid % T;                        // thread within the spectrum

    // ---- work range: pairs of group iterations -------------------------------------------
    const int64_t niter = (nspec + G - 1) / G;      // group iterations in the stream; the last one may be partial
    const int64_t npairs = (niter + 1) / 2;
    const int64_t p0 = (npairs * (int64_t)blockIdx.x) / gridDim.x;
    const int nit = (int)((npairs * (int64_t)(blockIdx.x + 1)) / gridDim.x - p0);
    if (nit <= 0) return;
    const int64_t it0 = 2 * p0;                     // my first group iteration
    const bool lastB_missing = it0 + 2 * nit > niter;
    const int last_valid = 2 * nit - 1 - (lastB_missing ? 1 : 0);     // local index of my last existing iteration

    const uint32_t ring = sm + kHdr;
    // Ring slots are filled for A: local iterations 0, 2, .. 2 nit (one past the end), B: 1, 3, .. 2 nit - 1, so that
    // every stage of every loop iteration waits and loads unconditionally; slots without an iteration of their own
    // re-read my last existing one and their result is never accumulated.
    auto issue = [&](int s, int j) {                 // the copy lane only
        const int slot = j % kStg;
        const uint32_t bar = sm + 8 * bar_full(s, slot);
        const int64_t it = it0 + min(2 * j + s, last_valid);
        const int64_t left = nspec - it * G;
        const uint32_t nbytes = (uint32_t)((left < G ? left : G) * 2 * N);     // the stream's last iteration may be short
        mbar_expect_tx(bar, nbytes);
        tma_load_1d(ring + (s * kStg + slot) * kIterBytes, bytes + (size_t)it * kIterBytes, nbytes, bar);
    };
    constexpr bool kWarpLocal = (T == 32);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 2 * kStg; i++) mbar_init(sm + 8 * i, 1);
        for (int i = 0; i < 2 * kStg; i++) mbar_init(sm + 8 * (2 * kStg + i), kThreads / 32);     // empty: all eight warps
        if (!kWarpLocal)
            for (int i = 0; i < 4 * G; i++) mbar_init(sm + 8 * (4 * kStg + i), T / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= kThreads) {                             // the copy warp group: one lane works
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == kThreads) {
            for (int j = 0; j < kStg; j++) {
                if (j <= nit) issue(0, j);
                if (j < nit) issue(1, j);
            }
            // Spectrum groups are not coupled to each other by the exchange barriers (with T = 32 nothing couples the
            // warps at all), so a warp may run more than an iteration ahead of another one.  "Slot consumed" is therefore
            // one mbarrier PER RING SLOT on which every warp arrives once per use of that slot: a second arrival by the same
            // warp needs the slot refilled, and it is refilled only after this lane has seen the phase complete -- no
            // barrier can be completed by the fast warps alone, and the lane cannot miss a phase.
            for (int j = 0; j < nit; j++) {
                const int slot = j % kStg;
                const uint32_t par = (uint32_t)((j / kStg) & 1);
                mbar_wait(sm + 8 * bar_empty(0, slot), par);           // every warp has stored its (A,1) rows of iteration j
                if (j + kStg <= nit) issue(0, j + kStg);               // ... so A's ring slot is free
                mbar_wait(sm + 8 * bar_empty(1, slot), par);
                if (j + kStg < nit) issue(1, j + kStg);
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");

    // ---- per-thread constants: twiddles (+ window) in registers --------------------------------
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
    float win[WINDOW ? 16 : 1];
    if (WINDOW) {
#pragma unroll
        for (int c = 0; c < C1; c++)
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++) win[c * R1 + n1] = window[n1 * NC1 + tau + T * c];
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;

    // shared-memory byte addresses (stream B = + one stream's worth)
    const uint32_t raw_ofs = grp * 2 * N + 2 * tau;
    const uint32_t s1 = sm + g::kOffS1 + grp * g::S1_ELEMS * 8;
    const uint32_t s2 = sm + g::kOffS2 + grp * g::S2_ELEMS * 8;

    double *row = rows + ((size_t)blockIdx.x * G + grp) * (size_t)N;
    auto flush = [&]() {
#pragma unroll
        for (int c = 0; c < C3; c++) {
            const int u = tau + T * c;
            const int k1 = u / R2, k2 = u % R2;
#pragma unroll
            for (int k3 = 0; k3 < R3; k3++) {
                const int k = k1 + R1 * k2 + R1 * R2 * k3;             // FFT output index
                row[(k + N / 2) & (N - 1)] += (double)acc[c * R3 + k3];   // fftshift == the (-1)^n of src/datastore.cxx:73
                acc[c * R3 + k3] = 0.0f;
            }
        }
    };

    c32 vA[16], vB[16];
    uint32_t raw[16];

    auto load_raw = [&](uint32_t slot_addr) {
#pragma unroll
        for (int c = 0; c < C1; c++)
#pragma unroll
            for (int n1 = 0; n1 < R1; n1++)
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw[c * R1 + n1]) : "r"(slot_addr + raw_ofs + 2 * (n1 * NC1 + T * c)));
    };
    auto unpack = [&](c32 (&v)[16]) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            c32 x = unpack_biased(raw[i]);
            if (WINDOW) {   // (u8 - 127.0f) * w, one rounding per rail: src/datastore.cxx:75-77
                x = __fadd2_rn(x, make_float2(-kBias, -kBias));
                x = __fmul2_rn(x, make_float2(win[i], win[i]));
            }
            v[i] = x;
        }
    };
    auto pass1 = [&](c32 (&v)[16]) {
#pragma unroll
        for (int c = 0; c < C1; c++) Dft<R1>::run(&v[c * R1]);
    };
    auto post1 = [&](c32 (&v)[16], int s, int slot, auto &&late) {
#pragma unroll
        for (int c = 0; c < C1; c++) {
            if (!WINDOW)   // remove the R1 * (32768+127) left in output 0 -- exact
                v[c * R1] = __fadd2_rn(v[c * R1], make_float2(-(float)R1 * kBias, -(float)R1 * kBias));
#pragma unroll
            for (int k1 = 1; k1 < R1; k1++) v[c * R1 + k1] = cmul(v[c * R1 + k1], tw1[c][k1 - 1]);
        }
        late();
#pragma unroll
        for (int c = 0; c < C1; c++)
#pragma unroll
            for (int k1 = 0; k1 < R1; k1++) sts64(s1 + s * g::kS1Bytes + 8 * (k1 * P1 + tau + T * c), v[c * R1 + k1]);
        // the stores above carry values computed from this warp's raw loads, so the ring slot is consumed
        if (kWarpLocal) __syncwarp(); else warp_arrive(sm + 8 * bar_rd<G>(s, 1, grp));
        if ((tid & 31) == 0) mbar_arrive(sm + 8 * bar_empty(s, slot));
    };
    auto load_s1 = [&](c32 (&v)[16], int s) {
#pragma unroll
        for (int c = 0; c < C2; c++) {
            const int q = tau + T * c;
            const int k1 = q / R3, n3 = q % R3;
#pragma unroll
            for (int n2 = 0; n2 < R2; n2++) v[c * R2 + n2] = lds64(s1 + s * g::kS1Bytes + 8 * (k1 * P1 + n2 * R3 + n3));
        }
    };
    auto pass2 = [&](c32 (&v)[16]) {
#pragma unroll
        for (int c = 0; c < C2; c++) Dft<R2>::run(&v[c * R2]);
    };
    auto post2 = [&](c32 (&v)[16], int s, auto &&late) {
#pragma unroll
        for (int c = 0; c < C2; c++)
#pragma unroll
            for (int k2 = 1; k2 < R2; k2++) v[c * R2 + k2] = cmul(v[c * R2 + k2], tw2[c][k2 - 1]);
        late();
#pragma unroll
        for (int c = 0; c < C2; c++) {
            const int q = tau + T * c;
            const int k1 = q / R3, n3 = q % R3;
#pragma unroll
            for (int k2 = 0; k2 < R2; k2++) sts64(s2 + s * g::kS2Bytes + 8 * ((k1 * R2 + k2) * P2 + n3), v[c * R2 + k2]);
        }
        if (kWarpLocal) __syncwarp(); else warp_arrive(sm + 8 * bar_rd<G>(s, 2, grp));
    };
    auto load_s2 = [&](c32 (&v)[16], int s) {
#pragma unroll
        for (int c = 0; c < C3; c++) {
            const int u = tau + T * c;
#pragma unroll
            for (int n3 = 0; n3 < R3; n3++) v[c * R3 + n3] = lds64(s2 + s * g::kS2Bytes + 8 * (u * P2 + n3));
        }
    };
    auto pass3 = [&](c32 (&v)[16]) {
#pragma unroll
        for (int c = 0; c < C3; c++) Dft<R3>::run(&v[c * R3]);
    };
    auto power = [&](const c32 (&v)[16], bool on) {
        if (on) {
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fmaf(v[i].x, v[i].x, fmaf(v[i].y, v[i].y, acc[i]));
        }
    };
    // does local iteration `li` hold a spectrum for my group?
    auto valid = [&](int li) { return li <= last_valid && (it0 + li) * G + grp < nspec; };

    // prologue: raw input of stream A, iteration 0
    mbar_wait(sm + 8 * bar_full(0, 0), 0);
    load_raw(ring);

    auto wait_rd = [&](int s, int e, uint32_t ph) { if (!kWarpLocal) mbar_wait(sm + 8 * bar_rd<G>(s, e, grp), ph); };
    auto none = []() {};
    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const int slot = j % kStg;
        const uint32_t ph = (uint32_t)(j & 1);
        // ---- (A,1)  | prefetch raw B ---------------------------------------------------------
        unpack(vA);
        pass1(vA);
        post1(vA, 0, slot, [&]() {
            mbar_wait(sm + 8 * bar_full(1, slot), (uint32_t)((j / kStg) & 1));
            load_raw(ring + (kStg + slot) * kIterBytes);
        });
        // ---- (B,1)  | prefetch S1[A] ----------------------------------------------------------
        unpack(vB);
        pass1(vB);
        wait_rd(0, 1, ph);
        load_s1(vA, 0);
        post1(vB, 1, slot, none);
        // ---- (A,2)  | prefetch S1[B] ----------------------------------------------------------
        pass2(vA);
        wait_rd(1, 1, ph);
        load_s1(vB, 1);
        post2(vA, 0, none);
        // ---- (B,2)  | prefetch S2[A] ----------------------------------------------------------
        pass2(vB);
        wait_rd(0, 2, ph);
        load_s2(vA, 0);
        post2(vB, 1, none);
        // ---- (A,3)  | prefetch S2[B] ----------------------------------------------------------
        pass3(vA);
        wait_rd(1, 2, ph);
        load_s2(vB, 1);
        power(vA, valid(2 * j));
        // ---- (B,3)  | prefetch raw A of the next iteration ------------------------------------
        pass3(vB);
        mbar_wait(sm + 8 * bar_full(0, (j + 1) % kStg), (uint32_t)(((j + 1) / kStg) & 1));
        load_raw(ring + ((j + 1) % kStg) * kIterBytes);
        power(vB, valid(2 * j + 1));

        if (++since_flush == kFlushPairs) { flush(); since_flush = 0; }
    }
    flush();
}

template <int N>
cudaError_t launch_n(const RpfLaunch &L)
{
    using g = Geo<N>;
    const int64_t niter = (L.nspec + g::G - 1) / g::G;
    const int64_t npairs = (niter + 1) / 2;
    if (npairs <= 0) return cudaSuccess;
    dim3 grid((unsigned)(npairs < L.grid ? npairs : L.grid)), block(kLaunchThreads);
    if (L.window)
        pipen_kernel<N, true><<<grid, block, g::kSmem, L.stream>>>(L.bytes, L.nspec, L.twiddle, L.window, L.rows);
    else
        pipen_kernel<N, false><<<grid, block, g::kSmem, L.stream>>>(L.bytes, L.nspec, L.twiddle, nullptr, L.rows);
    return cudaGetLastError();
}

template <int N>
cudaError_t configure_n()
{
    using g = Geo<N>;
    cudaError_t e = cudaFuncSetAttribute(pipen_kernel<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g::kSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pipen_kernel<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g::kSmem);
}

}  // namespace

bool rpf_pipen_supported(int N) { return N == 512 || N == 1024 || N == 2048; }
int rpf_pipen_groups(int N) { return kThreads / (N / 16); }

cudaError_t rpf_pipen_configure()
{
    cudaError_t e;
    if ((e = configure_n<512>()) != cudaSuccess) return e;
    if ((e = configure_n<1024>()) != cudaSuccess) return e;
    return configure_n<2048>();
}

cudaError_t rpf_launch_pipen(int N, const RpfLaunch &L)
{
    switch (N) {
        case 512: return launch_n<512>(L);
        case 1024: return launch_n<1024>(L);
        case 2048: return launch_n<2048>(L);
        default: return cudaErrorInvalidValue;
    }
}
