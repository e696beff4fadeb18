// solo65536.cu -- N = M * 4096 (M = 2, 4, 8, 16: 8192 ... 65536 bins) in ONE launch, one CTA per spectrum, thread-private
// hand-off.  Written for 65536 bins (M = 16), which the comments below describe; the outer radix
// is a template parameter (M < 16: an outer stage is 16 / M radix-M butterflies over as many values of a, a spectrum has
// M / 2 iterations, the scratch M x 16 blocks).
//
// n = 4096 n1 + m, k = ko + 16 k'.  The two-kernel four-step path writes the outer radix-16 to a complex64 scratch and
// reads it back (2 + 8 + 8 B of DRAM traffic per sample); the cluster kernel scatters it through distributed shared
// memory and is bound by that.  This kernel rests on one observation: with m = 256 a + t, the thread that runs the
// outer radix-16 of column m produces, for every ko, exactly the value that THREAD t ITSELF needs as input a of pass 1
// of the inner 4096-point transform of outer bin ko.  So if one CTA owns a whole spectrum, the hand-off between the
// outer pass and the 16 inner transforms is private to each thread:
//
//   * the 256 values a thread hands to itself per spectrum go through a per-CTA scratch of 512 KiB in global memory
//     (148 x 512 KiB = 74 MB: meant to stay in the 126 MB L2).  Thread t only ever touches [..][t]: no barrier, no
//     fence, no flag on the compute warps' side beyond three mbarrier arrives per iteration.  Written with st.global (16
//     per outer stage, 256 B per warp); read back by two otherwise idle warps of the copy warp group with 16-byte
//     cp.async (LDGSTS) into a 32 KiB row slot per stream as soon as that slot has been read, a whole iteration before
//     pass 1 needs the row -- only the first two rows of a spectrum also wait for its last outer stage (a first version
//     loaded the rows with ld.global straight into the idle stream's registers: one stage of distance is less than the
//     L2 latency under load, 10 % of all warp samples sat on the first use);
//   * the scratch is reused IN PLACE: spectrum s lives as [ko][a] blocks of 2 KiB, spectrum s + 1 with the fields
//     ko >> 1 and a >> (5 - log2 M) of the block index exchanged -- the outer stages of iteration i of the next spectrum
//     (a >> (5 - log2 M) = i) overwrite exactly the rows ko = 2i, 2i + 1 of the current one, consumed an iteration
//     earlier, in either layout;
//   * a CTA that runs all 16 inner transforms needs 16 x 16 accumulators per thread: they live in TENSOR MEMORY used
//     as a per-thread scratchpad (256 KiB = 256 floats per thread: tcgen05.ld.x16, 32 FFMA, tcgen05.st.x16 per inner
//     transform), flushed to the CTA's float64 rows every 128 spectra like everywhere else;
//   * M <= 4 (8192 / 16384 bins): the hand-off itself fits tensor memory (16 M complex values per thread next to the
//     16 M accumulators): tcgen05.st from the outer stage, one tcgen05.ld.x32 (or two .x16) per row -- no scratch, no
//     helper warps, no barrier for the hand-off, no L2 traffic beyond the raw bytes (602 / 534 Gsamples/s);
//   * the twiddle W_65536^{(256 a + t) ko} of the outer pass is kept in registers and advanced from a to a + 1 by a
//     multiplication with the compile-time constant W_256^{ko}, reloaded exactly at a = 0 of every spectrum.
//
// Iteration = two inner transforms (streams A: ko = 2i, B: ko = 2i + 1 of the same spectrum) + two outer stages of the
// next spectrum, 8 iterations per spectrum, stage order (A,1)(B,1)(A,2)(B,2) O O (A,3)(B,3) as in cluster65536.cu; the
// S1/S2 exchanges and their four split barriers are those of pipe4096.cu.  Raw bytes: 16 bulk copies of 1 KiB per
// iteration (the 512 columns of both outer stages x 16 values of n1) into a two-deep ring with a full/empty mbarrier
// pair per slot.  With a window: byte - 127 (exact), the 16 window values of an outer stage come from global memory (L2)
// and ride in the first radix-4 layer of the outer radix-16 (dft16_first_windowed), as in pipe4096.cu.
//
// Replaces reference src/datastore.cxx:68-87 for N = 8192 ... 65536 (launches of at least 96 spectra).
#include <stdint.h>
#include <type_traits>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kN = 4096;                      // inner transform
constexpr int kThreads = 256;                 // compute threads (8 warps)
constexpr int kLaunchThreads = kThreads + 128;
#ifndef RPF_SOLO_FLUSH
#define RPF_SOLO_FLUSH 128
#endif
constexpr int kFlushSpectra = RPF_SOLO_FLUSH; // float32 partial sums over at most this many spectra
constexpr int kRowBytes = kN * 8;             // one complex64 row of the inner transform
constexpr int kBlk = 256 * 8;                 // scratch block: 256 threads x one complex64 = 2 KiB
constexpr int kRawSlot = 16384;               // raw bytes per iteration: 2 outer stages x 16 values x 256 columns x 2 B
// geometry for outer radix M
template <int M>
struct Geo {
    static_assert(M == 2 || M == 4 || M == 8 || M == 16, "outer radix");
    static constexpr int kBig = kN * M;           // bins
    static constexpr int C = 16 / M;              // values of a (columns per thread) per outer stage
    static constexpr int AL = 32 / M;             // values of a per iteration = 2^(5 - log2 M)
    static constexpr int IPS = M / 2;             // iterations per spectrum (two inner transforms each)
    static constexpr int kChunk = kRawSlot / M;   // raw bytes per n1 and iteration
    // Block index of (ko, a) in the two layouts of the in-place scratch.  Layout 0 is [ko][a]; layout 1 exchanges the
    // fields ko >> 1 and a / AL (what an iteration consumes <-> what its outer stages produce), so that the blocks the
    // outer stages of iteration i write (a / AL = i, every ko) are the blocks of the rows ko = 2i, 2i + 1 in the other
    // layout.  blk1 = g1(ko) + f1(a).  M = 16 (AL = 2) uses the plain transpose [a][ko], which has the same property.
#ifndef RPF_SOLO_L1_TRANSPOSE16
#define RPF_SOLO_L1_TRANSPOSE16 1
#endif
    static constexpr bool kT16 = RPF_SOLO_L1_TRANSPOSE16 && M == 16;
    static __host__ __device__ constexpr int blk0(int ko, int a) { return ko * 16 + a; }
    static __host__ __device__ constexpr int g1(int ko) { return kT16 ? ko : (ko & 1) * 16 + (ko >> 1) * AL; }
    static __host__ __device__ constexpr int f1(int a) { return kT16 ? a * 16 : (a / AL) * 32 + (a % AL); }
    // M <= 4: the hand-off (16 M complex values per thread) fits tensor memory next to the accumulators -- no scratch, no
    // helper warps, no L2 traffic.  Cell (ko, a) = two 32-bit columns at kTmHand + 2 * blk(ko, a); M = 2 keeps layout 0
    // for every spectrum (its single iteration consumes both rows before its outer stages produce anything).
#ifndef RPF_SOLO_TMEM_HANDOFF
#define RPF_SOLO_TMEM_HANDOFF 1
#endif
    static constexpr bool kTm = RPF_SOLO_TMEM_HANDOFF && M <= 4;
    static constexpr int kTmHand = 64;            // first hand-off column (accumulators: columns [0, 16 M))
    static constexpr int kRun = kT16 ? 1 : AL;    // layout 1: a row is runs of kRun blocks ...
    static constexpr int kRunStride1 = kT16 ? 16 : 32;   // ... this many blocks apart (layout 0: contiguous)
};
#ifndef RPF_SOLO_RAWSTG
#define RPF_SOLO_RAWSTG 2
#endif
constexpr int kRawStg = RPF_SOLO_RAWSTG;      // power of two
// shared memory: header | row slots [stream] | S1 [stream] | S2 [stream] | raw ring [slot]
constexpr int kHdr = 256;
constexpr int kOffRow = kHdr;
constexpr int kOffS1 = kOffRow + 2 * kRowBytes;
constexpr int kOffS2 = kOffS1 + 2 * kRowBytes;
constexpr int kHelpers = 64;                  // warps 9, 10 stage the rows; lane 0 of warp 8 issues the raw bulk copies
constexpr int kOffRaw = kOffS2 + 2 * kRowBytes;
constexpr size_t kSmem = (size_t)kOffRaw + kRawStg * kRawSlot;
static_assert(kSmem <= 232448, "shared memory per CTA");
// mbarriers (8 bytes each)
__device__ __forceinline__ constexpr int bar_rd(int s, int e) { return 2 * s + (e - 1); }        // 0..3: 8 warps
__device__ __forceinline__ constexpr int bar_full(int slot) { return 4 + slot; }                 // tx 16 KiB
__device__ __forceinline__ constexpr int bar_empty(int slot) { return 4 + kRawStg + slot; }      // 8 warps
constexpr int kBarWritten = 4 + 2 * kRawStg;          // 8 warps: the whole outer pass of a spectrum is stored
__device__ __forceinline__ constexpr int bar_rowfull(int s) { return kBarWritten + 1 + s; }     // kHelpers cp.async arrivals
__device__ __forceinline__ constexpr int bar_rowfree(int s) { return kBarWritten + 3 + s; }     // 8 warps: row slot s has been read
constexpr int kTmemSlotOff = 8 * (kBarWritten + 5);   // uint32: TMEM base address written by tcgen05.alloc
static_assert(kTmemSlotOff + 4 <= kHdr && (kRawStg & (kRawStg - 1)) == 0, "header");
constexpr int kTmemCols = 512;

// W_256^j (cos, -sin) for the exponents the twiddle recurrence needs: the step from one outer stage to the next is
// W_{M^2}^{ko} = W_256^{ko 256 / M^2}: j = ko (M = 16), 4 ko (8), 16 ko (4), 64 (2)
__host__ __device__ constexpr float w256_re(int j)
{
    return j == 0 ? 1.0f : j == 1 ? 0.99969881869620425f : j == 2 ? 0.99879545620517241f : j == 3 ? 0.99729045667869021f
         : j == 4 ? 0.99518472667219693f : j == 5 ? 0.99247953459870997f : j == 6 ? 0.98917650996478101f
         : j == 7 ? 0.98527764238894122f : j == 8 ? 0.98078528040323043f : j == 9 ? 0.97570213003852857f
         : j == 10 ? 0.97003125319454397f : j == 11 ? 0.96377606579543984f : j == 12 ? 0.95694033573220882f
         : j == 13 ? 0.94952818059303667f : j == 14 ? 0.94154406518302081f : j == 15 ? 0.93299279883473896f
         : j == 16 ? 0.92387953251128674f : j == 20 ? 0.88192126434835505f : j == 24 ? 0.83146961230254524f
         : j == 28 ? 0.77301045336273699f : j == 32 ? 0.70710678118654757f : j == 48 ? 0.38268343236508984f : 0.0f;
}
__host__ __device__ constexpr float w256_im(int j)
{
    return j == 0 ? 0.0f : j == 1 ? -0.024541228522912288f : j == 2 ? -0.049067674327418015f : j == 3 ? -0.073564563599667426f
         : j == 4 ? -0.098017140329560604f : j == 5 ? -0.1224106751992162f : j == 6 ? -0.14673047445536175f
         : j == 7 ? -0.17096188876030122f : j == 8 ? -0.19509032201612825f : j == 9 ? -0.2191012401568698f
         : j == 10 ? -0.24298017990326387f : j == 11 ? -0.26671275747489837f : j == 12 ? -0.29028467725446233f
         : j == 13 ? -0.31368174039889152f : j == 14 ? -0.33688985339222005f : j == 15 ? -0.35989503653498811f
         : j == 16 ? -0.38268343236508978f : j == 20 ? -0.47139673682599764f : j == 24 ? -0.55557023301960218f
         : j == 28 ? -0.63439328416364549f : j == 32 ? -0.70710678118654746f : j == 48 ? -0.92387953251128674f : -1.0f;
}

// RPF_SOLO_HINTS = 1: L2 eviction priorities -- the raw stream is read once (evict_first), the scratch is what should
// stay (evict_last on its stores and loads).  Measured (B200, 4 GiB stream): without 310, with 344 Gsamples/s; DRAM
// traffic 4.8 x -> 2.1 x the algorithmic 2 B/sample, 1.16 x with the persisting set-aside below.
#ifndef RPF_SOLO_HINTS
#define RPF_SOLO_HINTS 1
#endif

// raw[c * M + n1] = the byte pair of column c, digit n1: columns 512 bytes apart, digits CHUNK bytes apart
template <int M, int CHUNK>
__device__ __forceinline__ void lds16x16(uint32_t base, uint32_t (&raw)[16])
{
#pragma unroll
    for (int j = 0; j < 16; j++)
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(raw[j]) : "r"(base + (j / M) * 512 + (j % M) * CHUNK));
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
#if RPF_SOLO_HINTS
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void *src, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void stg_c32_hint(void *p, c32 v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(p), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
#else
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void stg_c32(void *p, c32 v)
{
    asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
#endif

// tensor memory as a per-thread scratchpad: 16 consecutive 32-bit columns of the thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(r[0]), "f"(r[1]), "f"(r[2]), "f"(r[3]), "f"(r[4]), "f"(r[5]), "f"(r[6]), "f"(r[7]),
          "f"(r[8]), "f"(r[9]), "f"(r[10]), "f"(r[11]), "f"(r[12]), "f"(r[13]), "f"(r[14]), "f"(r[15]) : "memory");
}
// 2 N consecutive columns <-> N complex values
__device__ __forceinline__ void tmem_ld_c16(uint32_t taddr, c32 *v)       // 32 columns
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
__device__ __forceinline__ void tmem_ld_c8(uint32_t taddr, c32 *v)        // 16 columns
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(v[0].x), "=f"(v[0].y), "=f"(v[1].x), "=f"(v[1].y), "=f"(v[2].x), "=f"(v[2].y), "=f"(v[3].x), "=f"(v[3].y),
          "=f"(v[4].x), "=f"(v[4].y), "=f"(v[5].x), "=f"(v[5].y), "=f"(v[6].x), "=f"(v[6].y), "=f"(v[7].x), "=f"(v[7].y)
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_c8(uint32_t taddr, const c32 *v)  // 16 columns
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
          "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y) : "memory");
}
__device__ __forceinline__ void tmem_st_c4(uint32_t taddr, const c32 *v)  // 8 columns
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// CTA c takes the spectra c, c + grid, ... of the launch.  scratch: M * 4096 complex64 per CTA; rows: M private rows of
// 4096 doubles per CTA (row M c + ko = outer bin ko, natural order k'), summed by reduce_rows_fourstep().
template <bool WINDOW, int M>
__global__ void __launch_bounds__(kLaunchThreads, 1)
solo65536_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twiddle,
                 const float2 *__restrict__ twbig, const float *__restrict__ window, float2 *__restrict__ scratch,
                 double *__restrict__ rows)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int lo = tid & 15, hi = tid >> 4;       // pass 2: (k1 = lo, n3 = hi); pass 3: (k1 = lo, k2 = hi)
    using G = Geo<M>;
    constexpr int kBig = G::kBig, C = G::C, AL = G::AL, IPS = G::IPS, kChunk = G::kChunk;

    // ---- work range ------------------------------------------------------------------------
    const int mine = (int)((nspec - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x);   // >= 1: the host launches no idle CTA
    if (mine <= 0) return;
    const int nit = IPS * mine;
    const int npairs = IPS + nit;                 // raw pair slots consumed: IPS in the prologue, one per iteration
    auto spectrum_of = [&](int idx) { return (int64_t)blockIdx.x + (int64_t)gridDim.x * min(idx, mine - 1); };

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRawStg; s++) {
            mbar_init(sm + 8 * bar_full(s), 1);
            mbar_init(sm + 8 * bar_empty(s), kThreads / 32);
        }
#pragma unroll
        for (int i = 0; i < 4; i++) mbar_init(sm + 8 * i, kThreads / 32);
        mbar_init(sm + 8 * kBarWritten, kThreads / 32);
        mbar_init(sm + 8 * bar_rowfull(0), kHelpers);
        mbar_init(sm + 8 * bar_rowfull(1), kHelpers);
        mbar_init(sm + 8 * bar_rowfree(0), kThreads / 32);
        mbar_init(sm + 8 * bar_rowfree(1), kThreads / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sm + kTmemSlotOff), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *reinterpret_cast<const volatile uint32_t *>(smem + kTmemSlotOff);

    if (tid >= kThreads) {                             // the copy warp group
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (!G::kTm && tid >= kThreads + 32 && tid < kThreads + 32 + kHelpers) {
            // Row staging: generation g = the rows ko = 2i, 2i + 1 (i = g % IPS) of my spectrum number g / IPS, wanted by
            // iteration g.  A row is fetched as soon as its slot has been read for generation g - 1 (a whole iteration
            // before it is needed); only the first two rows of a spectrum also have to wait for the last outer stages of
            // that spectrum (`written`, once per spectrum), which leaves them two to three stages.
            const int h = tid - (kThreads + 32);
            const uint8_t *cta_scratch = reinterpret_cast<const uint8_t *>(scratch + (size_t)blockIdx.x * kBig);
#if RPF_SOLO_HINTS
            const uint64_t pol_rows = l2_policy_evict_last();
#endif
            for (int g = 0; g <= nit; g++) {       // (generation nit is never transformed: the loop stays branch-free)
                const int gs = g / IPS, gi = g % IPS, par = gs & 1;
#pragma unroll 1
                for (int s = 0; s < 2; s++) {
                    if (g > 0) mbar_wait_parked(sm + 8 * bar_rowfree(s), (uint32_t)((g - 1) & 1), 1000);
                    if (s == 0 && gi == 0) mbar_wait_parked(sm + 8 * kBarWritten, (uint32_t)(gs & 1), 1000);
                    // blocks (ko = 2 gi + s, a), a = 0..15, 2 KiB each: 16 consecutive ones in layout 0, runs of kRun blocks
                    // kRunStride1 apart in layout 1; of every block this thread copies the 16-byte units h and h + 64 (one
                    // pointer increment per run, no index arithmetic: the first version of this loop cost the compute
                    // warps a tenth of their issue slots)
                    constexpr int RUN = G::kRun;
                    const uint8_t *src = cta_scratch + (size_t)(par ? G::g1(2 * gi + s) : G::blk0(2 * gi + s, 0)) * kBlk + h * 16;
                    const size_t run_stride = par ? G::kRunStride1 * kBlk : RUN * kBlk;
                    const uint32_t dst = sm + kOffRow + s * kRowBytes + h * 16;
#pragma unroll
                    for (int a = 0; a < 16; a++) {
#if RPF_SOLO_HINTS
                        cp_async16_hint(dst + a * kBlk, src + (a % RUN) * kBlk, pol_rows);
                        cp_async16_hint(dst + a * kBlk + 1024, src + (a % RUN) * kBlk + 1024, pol_rows);
#else
                        cp_async16(dst + a * kBlk, src + (a % RUN) * kBlk);
                        cp_async16(dst + a * kBlk + 1024, src + (a % RUN) * kBlk + 1024);
#endif
                        if (a % RUN == RUN - 1) src += run_stride;
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(sm + 8 * bar_rowfull(s)) : "memory");
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        if (tid == kThreads) {
#if RPF_SOLO_HINTS
            const uint64_t pol_raw = l2_policy_evict_first();
#endif
            for (int p = 0; p < npairs; p++) {
                const int slot = p & (kRawStg - 1);
                if (p >= kRawStg) mbar_wait(sm + 8 * bar_empty(slot), (uint32_t)((p / kRawStg - 1) & 1));
                // pair p: the two outer stages of iteration p % IPS of my spectrum number p / IPS (one past the end: re-read)
                const uint8_t *src = bytes + (size_t)spectrum_of(p / IPS) * (2 * kBig) + (size_t)(p % IPS) * kChunk;
                const uint32_t bar = sm + 8 * bar_full(slot), dst = sm + kOffRaw + slot * kRawSlot;
                mbar_expect_tx(bar, kRawSlot);
#pragma unroll
                for (int n1 = 0; n1 < M; n1++) {
#if RPF_SOLO_HINTS
                    tma_load_1d_hint(dst + n1 * kChunk, src + (size_t)n1 * 2 * kN, kChunk, bar, pol_raw);
#else
                    tma_load_1d(dst + n1 * kChunk, src + (size_t)n1 * 2 * kN, kChunk, bar);
#endif
                }
            }
        }
        return;
    }
    asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");   // (240 - 168) x 256 = (168 - 24) x 128: exactly what the copy group gave back

    // ---- per-thread constants ----------------------------------------------------------------
    float2 tw1[15], tw2[15], twO[16];                            // twO[c M + ko] (ko = 0 unused)
#pragma unroll
    for (int k = 1; k < 16; k++) {
        tw1[k - 1] = twiddle[(tid * k) & (kN - 1)];             // W_4096^{t k1}
        tw2[k - 1] = twiddle[(hi * k * 16) & (kN - 1)];         // W_256^{n3 k2}
    }
    auto reload_twO = [&]() {                                    // W_N^{(256 c + t) ko}: the columns of outer stage 0 (a = c)
#pragma unroll
        for (int j = 0; j < 16; j++)
            if (j % M != 0) twO[j] = twbig[((256 * (j / M) + tid) * (j % M)) & (kBig - 1)];
    };
    reload_twO();

    // my accumulators: lanes of my warp's quadrant, columns [256 (warp >> 2), +256): column 16 ko + k3
    const uint32_t tm_acc = tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)((tid >> 7) * 256);
    const uint32_t tm_hand = tm_acc + G::kTmHand;                // (M <= 4) my cells of the hand-off
    (void)tm_hand;
    {
        float z[16];
#pragma unroll
        for (int k = 0; k < 16; k++) z[k] = 0.0f;
#pragma unroll 1
        for (int ko = 0; ko < M; ko++) tmem_st16(tm_acc + ko * 16, z);
        tmem_st_wait();
    }

#if RPF_SOLO_HINTS
    const uint64_t pol_keep = l2_policy_evict_last();
#define RPF_SOLO_STG(p, v) stg_c32_hint((p), (v), pol_keep)
#else
#define RPF_SOLO_STG(p, v) stg_c32((p), (v))
#endif
    uint8_t *sbase = reinterpret_cast<uint8_t *>(scratch + (size_t)blockIdx.x * kBig + tid);   // my column of every block
    const uint32_t raw_ld = sm + kOffRaw + 2 * tid;              // + slot * 16 KiB + half * C * 512 + c * 512 + n1 * kChunk
    const uint32_t row_ld = sm + kOffRow + tid * 8;              // + s * 32 KiB + a * 2048: x[256 a + t]
    const uint32_t s1_st = sm + kOffS1 + tid * 128 + ((tid & 7) << 4);
    const uint32_t s1_ld = sm + kOffS1 + hi * 128 + ((((lo >> 1) ^ (hi & 7)) << 4) | ((lo & 1) << 3));   // + n2 * 2048
    const uint32_t s2_st = sm + kOffS2 + hi * 128 + lo * 8;                                              // + k2 * 2048
    const uint32_t s2_ld = sm + kOffS2 + hi * 2048 + lo * 8;                                             // + n3 * 128
    constexpr uint32_t kStreamS = kRowBytes;

    double *row = rows + (size_t)blockIdx.x * kBig + tid;        // + ko * 4096 + 256 k3
    auto flush = [&]() {
        float z[16];
#pragma unroll
        for (int k = 0; k < 16; k++) z[k] = 0.0f;
#pragma unroll 1
        for (int ko = 0; ko < M; ko++) {
            float a16[16];
            tmem_ld16(tm_acc + ko * 16, a16);
            double cur[16];
            double *r = row + ko * kN;
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) cur[k3] = __ldcs(r + 256 * k3);             // streaming: the rows must not displace the scratch in L2
            tmem_ld_wait();
#pragma unroll
            for (int k3 = 0; k3 < 16; k3++) __stcs(r + 256 * k3, cur[k3] + (double)a16[k3]);   // inner bin k' = tid + 256 k3
            tmem_st16(tm_acc + ko * 16, z);
        }
        tmem_st_wait();
    };
    c32 vA[16], vB[16];
    uint32_t raw[16];
    float acc[16];

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
    float win[WINDOW ? 16 : 1];
    auto load_window = [&](int o) {                  // outer stage o: win[c M + n1] = w[4096 n1 + 256 (o C + c) + t]
        if (WINDOW) {
#pragma unroll
            for (int j = 0; j < 16; j++) win[j] = __ldg(window + (j % M) * kN + (o * C + j / M) * 256 + tid);
        }
    };
    // raw bytes of pair p, outer stage `half` of it -> raw[]
    auto fetch_raw = [&](int p, int half) {
        const int slot = p & (kRawStg - 1);
        if (half == 0) mbar_wait(sm + 8 * bar_full(slot), (uint32_t)((p / kRawStg) & 1));
        lds16x16<M, kChunk>(raw_ld + slot * kRawSlot + half * (C * 512), raw);
    };
    // Outer stage o = 2 i + half of a spectrum stored with layout `par`: raw[c M + n1] holds x[4096 n1 + 256 a + t] for my
    // C columns a = o C + c.  Radix-M over n1 on the biased bytes (exact, as in pass 1 of pipe4096.cu), times
    // W_N^{(256 a + t) ko}, 16 stores into my element of the blocks (ko, a); then the twiddles move on by C values of a.
    auto outer = [&](c32 (&v)[16], int i, int half, int par, auto &&fetch) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            v[j] = unpack_biased(raw[j]);
            if (WINDOW) v[j] = __fadd2_rn(v[j], make_float2(-kBias, -kBias));     // byte - 127.0f, exact
            if (WINDOW && M < 16) v[j] = __fmul2_rn(v[j], make_float2(win[j], win[j]));
        }
        // my element of block (ko, a): layout 0 = blk0, layout 1 = blk1 (Geo<M>); the part that depends on (i, half) here,
        // the part that depends on (c, ko) as an immediate below
        uint8_t *p = sbase + (size_t)(par == 0 ? i * AL + half * C : G::f1(i * AL + half * C)) * kBlk;   // (a = i AL + half C + c)
        auto put = [&](int c, int ko, c32 o) {       // (c, ko: compile-time after unrolling)
            if (!WINDOW && ko == 0) o = __fadd2_rn(o, make_float2(-(float)M * kBias, -(float)M * kBias));   // the bias survives in ko = 0 only (exact)
            if (ko != 0) o = cmul(o, twO[c * M + ko]);
            if (par == 0) RPF_SOLO_STG(p + (ko * 16 + c) * kBlk, o);
            else RPF_SOLO_STG(p + (G::g1(ko) + G::f1(c)) * kBlk, o);
            if (ko != 0) twO[c * M + ko] = cmul(twO[c * M + ko], w256_re(ko * (256 / (M * M))), w256_im(ko * (256 / (M * M))));
        };
        if (M == 16) {
            c32 y[16];
            if (WINDOW) dft16_first_windowed(v, win); else dft16_first(v);
            // the 16 stores leave in four groups, each behind the radix-4 that produces its outputs ko = q + 4 k1: sixteen
            // back to back fill the LSU queue, and whatever next overwrites one of their registers waits for it to drain
            radix4_tw<0>(v[0], v[1], v[2], v[3], y[0], y[4], y[8], y[12]);
#pragma unroll
            for (int k1 = 0; k1 < 4; k1++) put(0, 4 * k1, y[4 * k1]);
            radix4_tw<1>(v[4], v[5], v[6], v[7], y[1], y[5], y[9], y[13]);
#pragma unroll
            for (int k1 = 0; k1 < 4; k1++) put(0, 1 + 4 * k1, y[1 + 4 * k1]);
            fetch();
            if (WINDOW) load_window((2 * i + half + 1) & (M - 1));               // for the next outer stage
            radix4_tw<2>(v[8], v[9], v[10], v[11], y[2], y[6], y[10], y[14]);
#pragma unroll
            for (int k1 = 0; k1 < 4; k1++) put(0, 2 + 4 * k1, y[2 + 4 * k1]);
            radix4_tw<3>(v[12], v[13], v[14], v[15], y[3], y[7], y[11], y[15]);
#pragma unroll
            for (int k1 = 0; k1 < 4; k1++) put(0, 3 + 4 * k1, y[3 + 4 * k1]);
        } else if (!G::kTm) {
#pragma unroll
            for (int c = 0; c < C; c++) {
                Dft<M>::run(&v[c * M]);
#pragma unroll
                for (int ko = 0; ko < M; ko++) put(c, ko, v[c * M + ko]);
                if (c == C / 2 - 1) {
                    fetch();
                    if (WINDOW) load_window((2 * i + half + 1) & (M - 1));
                }
            }
        } else {
            // hand-off through tensor memory: per ko the C values of this stage are C consecutive cells
            c32 o[M][C];
#pragma unroll
            for (int c = 0; c < C; c++) {
                Dft<M>::run(&v[c * M]);
#pragma unroll
                for (int ko = 0; ko < M; ko++) {
                    c32 x = v[c * M + ko];
                    if (!WINDOW && ko == 0) x = __fadd2_rn(x, make_float2(-(float)M * kBias, -(float)M * kBias));
                    if (ko != 0) {
                        x = cmul(x, twO[c * M + ko]);
                        twO[c * M + ko] = cmul(twO[c * M + ko], w256_re(ko * (256 / (M * M))), w256_im(ko * (256 / (M * M))));
                    }
                    o[ko][c] = x;
                }
                if (c == C / 2 - 1) {
                    fetch();
                    if (WINDOW) load_window((2 * i + half + 1) & (M - 1));
                }
            }
            const bool l1 = M == 4 && par != 0;
            const uint32_t cell0 = l1 ? (uint32_t)G::f1(i * AL + half * C) : (uint32_t)(i * AL + half * C);
#pragma unroll
            for (int ko = 0; ko < M; ko++) {
                const uint32_t t = tm_hand + 2 * (cell0 + (l1 ? G::g1(ko) : ko * 16));
                if (C == 8) tmem_st_c8(t, o[ko]); else tmem_st_c4(t, o[ko]);
            }
        }
    };
    // my 16 inputs x[256 a + t] of stream s of iteration g, staged by the helper warps -> registers
    auto load_row = [&](c32 (&v)[16], int s, int g) {
        if (G::kTm) {
            const int gi = g % IPS, ko = 2 * gi + s;
            if (M == 2 || ((g / IPS) & 1) == 0) {
                tmem_ld_c16(tm_hand + 2 * (ko * 16), v);                       // layout 0: 16 consecutive cells
            } else {
                tmem_ld_c8(tm_hand + 2 * G::g1(ko), v);                       // layout 1: two runs of 8 cells, 32 apart
                tmem_ld_c8(tm_hand + 2 * (G::g1(ko) + 32), v + 8);
            }
        } else {
            mbar_wait(sm + 8 * bar_rowfull(s), (uint32_t)(g & 1));
            lds64x16<2048>(row_ld + s * kStreamS, v);
        }
    };

    // ---- prologue: the whole outer pass of my first spectrum; rows 0 and 1 on their way ------------------------
    load_window(0);
    fetch_raw(0, 0);
#pragma unroll 1
    for (int pp = 0; pp < IPS; pp++) {
        outer(vB, pp, 0, 0, [&]() { fetch_raw(pp, 1); });
        outer(vB, pp, 1, 0, [&]() {});
        warp_arrive(sm + 8 * bar_empty(pp & (kRawStg - 1)));
        if (pp < IPS - 1) fetch_raw(pp + 1, 0);
    }
    if (G::kTm) tmem_st_wait(); else warp_arrive(sm + 8 * kBarWritten);
    load_row(vA, 0, 0);

    const uint32_t rd = sm + 8 * bar_rd(0, 1);           // bar_rd(s, e) = rd + 8 (2 s + e - 1)
    int since_flush = 0;
    for (int j = 0; j < nit; j++) {
        const uint32_t ph = (uint32_t)(j & 1);
        const int i = j % IPS;                            // streams A, B = outer bins 2i, 2i + 1 of spectrum j / IPS
        const int par = (j / IPS) & 1;
        if (i == 0) { reload_twO(); load_window(0); }     // the outer stages of this spectrum's successor start at a = 0

        // ---- (A,1) | row B ------------------------------------------------------------------------
        if (G::kTm) tmem_ld_wait();                       // row A (tcgen05.ld in (B,3) of the previous iteration)
        radix(vA, [&]() { if (!G::kTm) warp_arrive(sm + 8 * bar_rowfree(0)); load_row(vB, 1, j); });     // (row A has been consumed)
        post1(vA, 0);
        // ---- (B,1) | S1[A] ------------------------------------------------------------------------
        if (G::kTm) tmem_ld_wait();                       // row B
        radix(vB, [&]() { if (!G::kTm) warp_arrive(sm + 8 * bar_rowfree(1)); mbar_wait(rd, ph); lds64x16<2048>(s1_ld, vA); });
        post1(vB, 1);
        // ---- (A,2) | S1[B] ------------------------------------------------------------------------
        radix(vA, [&]() { mbar_wait(rd + 16, ph); lds64x16<2048>(s1_ld + kStreamS, vB); });
        post2(vA, 0, []() {});
        // ---- (B,2) | S2[A]; raw bytes of the first outer stage ---------------------------------------
        radix(vB, [&]() { mbar_wait(rd + 8, ph); lds64x16<128>(s2_ld, vA); });
        post2(vB, 1, [&]() { fetch_raw(IPS + j, 0); });
        // ---- O(2i), O(2i + 1) of the next spectrum (layout par ^ 1) ---------------------------------
        outer(vB, i, 0, par ^ 1, [&]() { fetch_raw(IPS + j, 1); });
        outer(vB, i, 1, par ^ 1, [&]() {});
        warp_arrive(sm + 8 * bar_empty((IPS + j) & (kRawStg - 1)));
        if (!G::kTm && i == IPS - 1) warp_arrive(sm + 8 * kBarWritten);   // the outer pass of the next spectrum is complete
        // ---- (A,3) | S2[B] ------------------------------------------------------------------------
        tmem_st_wait();                                   // (the stores of the previous iteration: long done)
        tmem_ld16(tm_acc + (2 * i) * 16, acc);
        radix(vA, [&]() { mbar_wait(rd + 24, ph); lds64x16<128>(s2_ld + kStreamS, vB); });
        tmem_ld_wait();
        power(vA);
        tmem_st16(tm_acc + (2 * i) * 16, acc);
        // ---- (B,3) | row A of the next iteration -------------------------------------------------------
        tmem_ld16(tm_acc + (2 * i + 1) * 16, acc);
        radix(vB, [&]() { load_row(vA, 0, j + 1); });
        tmem_ld_wait();
        power(vB);
        tmem_st16(tm_acc + (2 * i + 1) * 16, acc);

        if (i == IPS - 1 && ++since_flush == kFlushSpectra) { tmem_st_wait(); flush(); since_flush = 0; }
    }
    tmem_st_wait();
    flush();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
}

}  // namespace

// RPF_SOLO_PERSIST = 1: while a context of this kind exists on the device, the L2 set-aside for persisting (evict_last)
// lines is raised to what the device allows (79 MiB on the B200; the scratch is 74 MB), so that the raw stream and the
// float64 rows cannot displace the scratch: DRAM writes 2.1 GB -> 0.17 GB per 2.1 GB of input.  The previous limit is
// restored by rpf_solo65536_release().
#ifndef RPF_SOLO_PERSIST
#define RPF_SOLO_PERSIST 1
#endif
cudaError_t rpf_solo65536_configure(int M, size_t *saved_limit)
{
    *saved_limit = (size_t)-1;
#if RPF_SOLO_PERSIST
    if (!Geo<2>::kTm || M > 4) {               // (M <= 4 hands over through tensor memory: nothing to keep in L2)
        int dev = 0, maxp = 0;
        size_t cur = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&maxp, cudaDevAttrMaxPersistingL2CacheSize, dev) == cudaSuccess && maxp > 0 &&
            cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize) == cudaSuccess && cur < (size_t)maxp &&
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)maxp) == cudaSuccess)
            *saved_limit = cur;
        cudaGetLastError();                        // best effort: without the set-aside the hints alone still work
    }
#endif
    const void *fns[] = {(const void *)solo65536_kernel<false, 2>, (const void *)solo65536_kernel<true, 2>,
                         (const void *)solo65536_kernel<false, 4>, (const void *)solo65536_kernel<true, 4>,
                         (const void *)solo65536_kernel<false, 8>, (const void *)solo65536_kernel<true, 8>,
                         (const void *)solo65536_kernel<false, 16>, (const void *)solo65536_kernel<true, 16>};
    for (const void *f : fns) {
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

void rpf_solo65536_release(size_t saved_limit)
{
    if (saved_limit != (size_t)-1) {
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, saved_limit);
        cudaGetLastError();
    }
}

size_t rpf_solo65536_scratch_bytes(int M, int grid)
{
    if (Geo<2>::kTm && M <= 4) return 0;       // hand-off through tensor memory
    return (size_t)grid * M * kN * sizeof(float2);
}

template <int M>
static cudaError_t launch_solo(const RpfLaunch &L, const float2 *twbig, float2 *scratch)
{
    const unsigned grid = (unsigned)((int64_t)L.grid < L.nspec ? L.grid : L.nspec);
    if (L.window)
        solo65536_kernel<true, M><<<grid, kLaunchThreads, kSmem, L.stream>>>(L.bytes, L.nspec, L.twiddle, twbig, L.window, scratch, L.rows);
    else
        solo65536_kernel<false, M><<<grid, kLaunchThreads, kSmem, L.stream>>>(L.bytes, L.nspec, L.twiddle, twbig, nullptr, scratch, L.rows);
    return cudaGetLastError();
}

// L.M = outer radix, L.grid = CTAs (<= SMs), L.rows = M rows of 4096 doubles per CTA, scratch = rpf_solo65536_scratch_bytes(M, L.grid).
cudaError_t rpf_launch_solo65536(const RpfLaunch &L, const float2 *twbig, float2 *scratch)
{
    if (L.nspec <= 0) return cudaSuccess;
    switch (L.M) {
        case 2: return launch_solo<2>(L, twbig, scratch);
        case 4: return launch_solo<4>(L, twbig, scratch);
        case 8: return launch_solo<8>(L, twbig, scratch);
        case 16: return launch_solo<16>(L, twbig, scratch);
        default: return cudaErrorInvalidValue;
    }
}
