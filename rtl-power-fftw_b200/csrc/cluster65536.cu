// cluster65536.cu -- N = 65536 in ONE launch with no scratch round trip through HBM.
//
// n = 4096 n1 + m, k = ko + 16 k'.  A cluster of 16 CTAs owns a spectrum pair (streams A and B); CTA `ko` of the
// cluster runs the 4096-point inner transform of outer bin ko for both spectra with exactly the two-stream
// pipeline of pipe4096.cu.  What the two-kernel path writes to and reads back from a complex64 scratch in HBM
// (16 B per sample) moves through distributed shared memory instead:
//
//   outer stage O(s'):  thread t of CTA c owns column m = 256 c + t of the NEXT spectrum of stream s: it reads its 16
//                       samples x[4096 n1 + m] from the CTA's raw ring (16 bulk copies of 512 B per spectrum),
//                       runs the radix-16 over n1 on the biased bytes (exact, as in pass 1 of pipe4096.cu),
//                       multiplies by W_65536^{m ko} (15 thread-constant twiddles) and sends output ko to CTA ko with
//                       st.async: 8 bytes into row slot s of that CTA at offset 8 m, complete_tx on that CTA's
//                       "row full" mbarrier.  A row slot is complete when 32 KiB have landed from the 16 CTAs.
//   "slot free":        the copy lane of CTA ko sees (local mbarrier) that all its warps have stored their pass-1
//                       rows of stream s -- the row slot has been read -- re-arms "row full" and arrives on the
//                       "free[s]" mbarrier of all 16 CTAs (remote arrive, release.cluster); producers wait on their
//                       own free[s] (acquire.cluster) before the stores of the next generation.
//
// Stage order per iteration:  (A,1)(B,1)(A,2)(B,2) O(A') O(B') (A,3)(B,3): both cluster-wide dependencies (slot free
// -> O stores; O stores of all 16 CTAs -> pass 1 of the next iteration) have about two stages (~1 us) of slack.
// No window (a windowed 65536-bin transform takes the two-kernel path).
//
// Replaces reference src/datastore.cxx:68-87 for N = 65536.
#include <stdint.h>
#include <type_traits>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;                      // inner transform
constexpr int kM = 16;                        // outer radix = cluster size
constexpr int kBig = kN * kM;
constexpr int kThreads = 256;                 // compute threads (8 warps)
constexpr int kLaunchThreads = kThreads + 128;
constexpr int kFlushPairs = 64;
constexpr int kRowBytes = kN * 8;             // one complex64 row of the inner transform
constexpr int kCols = kN / kM;                // 256 columns of the outer pass per CTA (one per thread)
constexpr int kChunk = kCols * 2;             // 512 raw bytes per (n1, CTA)
constexpr int kRawSlot = kM * kChunk;         // 8 KiB of raw bytes per (stream, generation)
constexpr int kRawStg = 2;
// shared memory: header | row slots [stream] | S1 [stream] | S2 [stream] | raw ring [stream][slot]
constexpr int kHdr = 256;
constexpr int kOffRow = kHdr;
constexpr int kOffS1 = kOffRow + 2 * kRowBytes;
constexpr int kOffS2 = kOffS1 + 2 * kRowBytes;
constexpr int kOffRaw = kOffS2 + 2 * kRowBytes;
constexpr size_t kSmem = (size_t)kOffRaw + 2 * kRawStg * kRawSlot;
static_assert(kSmem <= 232448, "shared memory per CTA");
// mbarriers (8 bytes each)
__device__ __forceinline__ constexpr int bar_rowfull(int s) { return s; }                       // 0, 1: tx 32 KiB
__device__ __forceinline__ constexpr int bar_free(int s) { return 2 + s; }                      // 2, 3: 16 remote arrives
__device__ __forceinline__ constexpr int bar_rd(int s, int e) { return 4 + 2 * s + (e - 1); }   // 4..7
__device__ __forceinline__ constexpr int bar_rawfull(int s, int slot) { return 8 + 2 * s + slot; }   // 8..11
__device__ __forceinline__ constexpr int bar_orc(int s) { return 12 + s; }                      // 12, 13: raw slot consumed
constexpr int kBarDone = 14;

__device__ __forceinline__ uint32_t cluster_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// A release.cluster arrive compiles to MEMBAR.ALL.CTA + MEMBAR.ALL.GPU + ERRBAR + CGAERRBAR + the arrive; sixteen of them
// per stream and iteration serialised the copy lane for several microseconds (the first version of this kernel ran at
// 104 Gsamples/s because of it).  One cluster-scope fence, then sixteen relaxed arrives.
__device__ __forceinline__ void fence_cluster() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ void remote_arrive_relaxed(uint32_t raddr)
{
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_async_c32(uint32_t raddr, c32 v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "f"(v.x), "f"(v.y), "r"(rbar) : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void lds16x16(uint32_t base, uint32_t (&raw)[16])
{
#pragma unroll
    for (int n1 = 0; n1 < 16; n1++) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw[n1]) : "r"(base + n1 * kChunk));
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

// Cluster q takes the spectra q, q + nclusters, ... of the launch, two per iteration (streams A and B); rows[] has
// one private 4096-double row per CTA (outer bin ko = blockIdx.x % 16, natural order k'), summed by
// reduce_rows_fourstep().
__global__ void __launch_bounds__(kLaunchThreads, 1)
cluster65536_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
                    const float2 *__restrict__ twbig, double *__restrict__ rows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int lo = tid & 15, hi = tid >> 4;       // pass 2: (k1 = lo, n3 = hi); pass 3: (k1 = lo, k2 = hi)
    const uint32_t rank = cluster_rank();         // = outer bin ko of this CTA
    const int q = (int)(blockIdx.x / kM), nclusters = (int)(gridDim.x / kM);

    // ---- work range of the cluster -------------------------------------------------------
    const int mine = (int)((nspec - q + nclusters - 1) / nclusters);       // >= 1: the host launches no idle cluster
    const int nit = (mine + 1) / 2;
    const bool lastB_missing = (mine & 1) != 0;
    const int last_valid = mine - 1;
    // generations 0..nit of both streams are produced (one past the end), so that every stage of every iteration
    // runs unconditionally; generations without a spectrum of their own re-read the last valid one and are never
    // accumulated
    auto spectrum_of = [&](int s, int g) { return (int64_t)q + (int64_t)nclusters * min(2 * g + s, last_valid); };

    auto issue_raw = [&](int s, int g) {             // the copy lane: my 256 columns of all 16 n1, 512 bytes each
        const int slot = g % kRawStg;
        const uint32_t bar = sm + 8 * bar_rawfull(s, slot);
        const uint8_t *src = bytes + (size_t)spectrum_of(s, g) * (2 * kBig) + (size_t)rank * kChunk;
        const uint32_t dst = sm + kOffRaw + (s * kRawStg + slot) * kRawSlot;
        mbar_expect_tx(bar, kRawSlot);
#pragma unroll
        for (int n1 = 0; n1 < kM; n1++) tma_load_1d(dst + n1 * kChunk, src + (size_t)n1 * 2 * kN, kChunk, bar);
    };
    if (tid == 0) {
        mbar_init(sm + 8 * bar_rowfull(0), 1);
        mbar_init(sm + 8 * bar_rowfull(1), 1);
        mbar_init(sm + 8 * bar_free(0), kM);
        mbar_init(sm + 8 * bar_free(1), kM);
#pragma unroll
        for (int s = 0; s < 2; s++) {
            mbar_init(sm + 8 * bar_rd(s, 1), kThreads / 32 + 1);           // + the copy lane
            mbar_init(sm + 8 * bar_rd(s, 2), kThreads / 32);
            mbar_init(sm + 8 * bar_rawfull(s, 0), 1);
            mbar_init(sm + 8 * bar_rawfull(s, 1), 1);
            mbar_init(sm + 8 * bar_orc(s), kThreads / 32);
        }
        mbar_init(sm + 8 * kBarDone, kM);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(sm + 8 * bar_rowfull(0), kRowBytes);               // generation 0
        mbar_expect_tx(sm + 8 * bar_rowfull(1), kRowBytes);
    }
    cluster_sync_all();                                // every CTA's barriers exist before anybody signals them

    if (tid >= kThreads) {                             // the copy warp group: one lane works
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (tid == kThreads) {
            for (int g = 0; g < kRawStg; g++)
                if (g <= nit) { issue_raw(0, g); issue_raw(1, g); }
            // the raw slots of generation 0 are consumed in the prologue (phase 0 of orc)
            for (int s = 0; s < 2; s++) {
                mbar_wait(sm + 8 * bar_orc(s), 0);
                if (kRawStg <= nit) issue_raw(s, kRawStg);
            }
            for (int j = 0; j < nit; j++) {
                const uint32_t ph = (uint32_t)(j & 1);
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    mbar_arrive(sm + 8 * bar_rd(s, 1));
                    mbar_wait(sm + 8 * bar_rd(s, 1), ph);              // every compute warp has stored its (s,1) rows:
                    mbar_expect_tx(sm + 8 * bar_rowfull(s), kRowBytes);   // row slot s is free; arm it for generation j + 1
                    const uint32_t fr = sm + 8 * bar_free(s);
                    fence_cluster();
#pragma unroll
                    for (int r = 0; r < kM; r++) remote_arrive_relaxed(mapa(fr, r));
                }
#pragma unroll
                for (int s = 0; s < 2; s++) {
                    mbar_wait(sm + 8 * bar_orc(s), ph ^ 1u);           // raw slot of generation j + 1 has been read
                    if (j + 1 + kRawStg <= nit) issue_raw(s, j + 1 + kRawStg);
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");

    // ---- per-thread constants ----------------------------------------------------------------
    float2 tw1[15], tw2[15], twO[15];
    const int m = (int)rank * kCols + tid;                       // my column of the outer pass
#pragma unroll
    for (int k = 1; k < 16; k++) {
        tw1[k - 1] = twiddle[(tid * k) & (kN - 1)];             // W_4096^{t k1}
        tw2[k - 1] = twiddle[(hi * k * 16) & (kN - 1)];         // W_256^{n3 k2}
        twO[k - 1] = twbig[(m * k) & (kBig - 1)];               // W_65536^{m ko}
    }
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = 0.0f;

    const uint32_t raw_off = sm + kOffRaw + 2 * tid;             // + (s * kRawStg + slot) * kRawSlot + n1 * 512
    const uint32_t row_ld = sm + kOffRow + tid * 8;              // + s * 32 KiB + a * 2048: x[256 a + t]
    const uint32_t row_st = ((sm + kOffRow) & 0x00FFFFFFu) + (uint32_t)m * 8;   // + s * 32 KiB, in CTA ko (mapa)
    const uint32_t s1_st = sm + kOffS1 + tid * 128 + ((tid & 7) << 4);
    const uint32_t s1_ld = sm + kOffS1 + hi * 128 + ((((lo >> 1) ^ (hi & 7)) << 4) | ((lo & 1) << 3));   // + n2 * 2048
    const uint32_t s2_st = sm + kOffS2 + hi * 128 + lo * 8;                                              // + k2 * 2048
    const uint32_t s2_ld = sm + kOffS2 + hi * 2048 + lo * 8;                                             // + n3 * 128
    constexpr uint32_t kStreamS = kRowBytes;

    double *row = rows + (size_t)blockIdx.x * kN;
    auto flush = [&]() {
        double cur[16];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) cur[k3] = row[tid + 256 * k3];
#pragma unroll
        for (int k3 = 0; k3 < 16; k3++) {
            row[tid + 256 * k3] = cur[k3] + (double)acc[k3];     // inner bin k' = tid + 256 k3
            acc[k3] = 0.0f;
        }
    };
    c32 vA[16], vB[16];
    uint32_t raw[16];

    auto post1 = [&](c32 (&v)[16], int s) {
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw1[k - 1]);
        sts_s1(s1_st + s * kStreamS, v);
        warp_arrive(sm + 8 * bar_rd(s, 1));
    };
    auto post2 = [&](c32 (&v)[16], int s, auto &&late) {
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], tw2[k - 1]);
        late();
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
    auto fetch_raw = [&](int s, int g) {               // raw bytes of generation g of stream s -> raw[]
        const int slot = g % kRawStg;
        mbar_wait(sm + 8 * bar_rawfull(s, slot), (uint32_t)((g / kRawStg) & 1));
        lds16x16(raw_off + (s * kRawStg + slot) * kRawSlot, raw);
    };
    // outer stage: raw[] holds my column of generation g of stream s.  `freepar` < 0: the row slots are known to be
    // free (generation 0).
    auto outer = [&](c32 (&v)[16], int s, int freepar, auto &&fetch) {
#pragma unroll
        for (int n1 = 0; n1 < 16; n1++) v[n1] = unpack_biased(raw[n1]);
        radix(v, fetch);
        warp_arrive(sm + 8 * bar_orc(s));                                 // the raw slot has been read by this warp
        v[0] = __fadd2_rn(v[0], make_float2(-16.0f * kBias, -16.0f * kBias));   // the bias survives in ko = 0 only (exact)
#pragma unroll
        for (int k = 1; k < 16; k++) v[k] = cmul(v[k], twO[k - 1]);
        if (freepar >= 0) mbar_wait_cluster(sm + 8 * bar_free(s), (uint32_t)freepar);
        // shared::cluster address = (cta_rank << 24) | offset on sm_100a -- what mapa computes (one PRMT; printed by
        // tools/cluster_probe.cu).  Written out here so that ko stays an immediate: with 32 mapa results per stage ptxas
        // keeps them in registers and the loop spills.
        const uint32_t dst = row_st + s * kStreamS, fb = (sm + 8 * bar_rowfull(s)) & 0x00FFFFFFu;
#pragma unroll
        for (int ko = 0; ko < 16; ko++) st_async_c32(dst | ((uint32_t)ko << 24), v[ko], fb | ((uint32_t)ko << 24));
    };

    // ---- prologue: generation 0 of both streams through the outer stage; row A into registers ------
    fetch_raw(0, 0);
    outer(vB, 0, -1, [&]() { fetch_raw(1, 0); });
    outer(vB, 1, -1, []() {});
    mbar_wait(sm + 8 * bar_rowfull(0), 0);
    lds64x16<2048>(row_ld, vA);

    const uint32_t rd = sm + 8 * bar_rd(0, 1);           // bar_rd(s, e) = rd + 8 (2 s + e - 1)
    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const uint32_t ph = (uint32_t)(j & 1);
        const bool haveB = !(lastB_missing && j == nit - 1);

        // ---- (A,1) | row B of generation j ------------------------------------------------------
        radix(vA, [&]() { mbar_wait(sm + 8 * bar_rowfull(1), ph); lds64x16<2048>(row_ld + kStreamS, vB); });
        post1(vA, 0);
        // ---- (B,1) | S1[A] ------------------------------------------------------------------------
        radix(vB, [&]() { mbar_wait(rd, ph); lds64x16<2048>(s1_ld, vA); });
        post1(vB, 1);
        // ---- (A,2) | S1[B] ------------------------------------------------------------------------
        radix(vA, [&]() { mbar_wait(rd + 16, ph); lds64x16<2048>(s1_ld + kStreamS, vB); });
        post2(vA, 0, []() {});
        // ---- (B,2) | S2[A]; raw bytes of A, generation j + 1 ----------------------------------------
        radix(vB, [&]() { mbar_wait(rd + 8, ph); lds64x16<128>(s2_ld, vA); });
        post2(vB, 1, [&]() { fetch_raw(0, j + 1); });
        // ---- O(A') | raw bytes of B, generation j + 1;  O(B') -----------------------------------------
        outer(vB, 0, (int)ph, [&]() { fetch_raw(1, j + 1); });
        outer(vB, 1, (int)ph, []() {});
        // ---- (A,3) | S2[B] ------------------------------------------------------------------------
        radix(vA, [&]() { mbar_wait(rd + 24, ph); lds64x16<128>(s2_ld + kStreamS, vB); });
        power(vA);
        // ---- (B,3) | row A of generation j + 1 -------------------------------------------------------
        radix(vB, [&]() { mbar_wait(sm + 8 * bar_rowfull(0), ph ^ 1u); lds64x16<2048>(row_ld, vA); });
        if (haveB) power(vB);

        if (++since_flush == kFlushPairs) { flush(); since_flush = 0; }
    }
    mbar_wait(sm + 8 * bar_rowfull(1), (uint32_t)(nit & 1));     // the last (unused) generation has landed here too
    flush();
    // nobody leaves while a peer may still store into it or signal it: all 16 CTAs meet on `done`
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    if (tid == 0) {
        const uint32_t dn = sm + 8 * kBarDone;
        fence_cluster();
#pragma unroll
        for (int r = 0; r < kM; r++) remote_arrive_relaxed(mapa(dn, r));
        mbar_wait_cluster(dn, 0);
    }
}

}  // namespace

static cudaLaunchConfig_t cluster_config(cudaLaunchAttribute *attr, int grid, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kLaunchThreads);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kM; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cfg;
}

// Opt-in attributes; *max_clusters = how many 16-CTA clusters the device keeps resident (0: take the two-kernel path).
cudaError_t rpf_cluster65536_configure(int *max_clusters)
{
    *max_clusters = 0;
    cudaError_t e = cudaFuncSetAttribute(cluster65536_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(cluster65536_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = cluster_config(attr, kM * 64, 0);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, cluster65536_kernel, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return cudaSuccess; }      // no 16-CTA clusters here: not an error
    *max_clusters = n;
    return cudaSuccess;
}

// L.grid = 16 * clusters (<= the resident maximum), L.rows = one 4096-double row per CTA.
cudaError_t rpf_launch_cluster65536(const RpfLaunch &L, const float2 *twbig)
{
    if (L.nspec <= 0) return cudaSuccess;
    int clusters = L.grid / kM;
    if ((int64_t)clusters > L.nspec) clusters = (int)L.nspec;
    cudaLaunchAttribute attr[1];
    cudaLaunchConfig_t cfg = cluster_config(attr, clusters * kM, L.stream);
    return cudaLaunchKernelEx(&cfg, cluster65536_kernel, L.bytes, L.nspec, L.twiddle, twbig, L.rows);
}
