// outer_tma.cu -- four-step outer pass (N = M * 4096, M = 2, 4, 8, 16) as a TMA-staged streaming kernel.
//
// n = 4096 n1 + m, k = ko + M k'.  The pass reads 2 bytes and writes 8 bytes per sample and does almost no arithmetic,
// so all that matters is how many bytes are in flight.  The round-1 kernel (aux_kernels.cu: one thread per column,
// 2-byte loads with an 8 KiB stride, 8-byte stores) kept the memory system at 4.3 TB/s; here both directions are
// bulk copies, so the bytes in flight do not depend on how many warps are resident:
//
//   tile = 4096 samples of one spectrum: TC = 4096 / M adjacent columns m, all M digits n1
//   load   M bulk copies (8192 / M bytes each: the TC columns of one n1) into a 4-deep shared-memory ring
//   work   thread t owns the C = 16 / M columns t, t + 256, ..: u8 -> float (- 127, * window), radix-M butterfly,
//          twiddles W_N^{m ko} -- the same arithmetic, in the same order, as the round-1 kernel (bit-identical scratch)
//   store  outputs staged in shared memory as [ko][TC] and sent with M bulk stores (32768 / M bytes each,
//          contiguous in scratch[spec][ko][.]); two staging buffers, reuse gated by cp.async.bulk.wait_group.read
//
// A CTA is bound to ONE column range (tile index mt = blockIdx.x % M) for its whole life, so its twiddles (and window
// values) stay in registers; it takes the spectra part, part + P, ...  Two CTAs per SM.
//
// Replaces (with the inner transform that follows) reference src/datastore.cxx:68-82 for N = 8192 ... 65536.
#include <stdint.h>
#include "fft_math.cuh"
#include "rpf_internal.h"
#include "stage_prims.cuh"

namespace {

constexpr int kNI = 4096;
constexpr int kThreadsO = 256;
constexpr int kRawStgO = 4;                   // tiles of raw bytes in flight per CTA
constexpr int kTileRaw = 2 * kNI;             // 8 KiB of raw bytes per tile
constexpr int kTileOut = 8 * kNI;             // 32 KiB of complex64 per tile
constexpr int kHdrO = 128;
constexpr int kOffRawO = kHdrO;
constexpr int kOffOutO = kOffRawO + kRawStgO * kTileRaw;
constexpr size_t kSmemO = (size_t)kOffOutO + 2 * kTileOut;      // 98,432 bytes: two CTAs per SM

__device__ __forceinline__ void tma_store_1d(void *dst, uint32_t src, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

template <int M, bool WINDOW>
__global__ void __launch_bounds__(kThreadsO, 2)
fourstep_outer_tma_kernel(const uint8_t *__restrict__ bytes, int64_t nspec, const float2 *__restrict__ twbig,
                          const float *__restrict__ window, float2 *__restrict__ scratch)
{
    constexpr int NT = M * kNI;
    constexpr int C = 16 / M;                 // columns per thread
    constexpr int TC = kNI / M;               // columns per tile (= 256 C)
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    const int mt = (int)(blockIdx.x % M), part = (int)(blockIdx.x / M), nparts = (int)(gridDim.x / M);
    if (part >= nspec) return;
    const int ntiles = (int)((nspec - part + nparts - 1) / nparts);       // my spectra: part, part + nparts, ...

    float2 tw[C][M];
    float win[WINDOW ? 16 : 1];
#pragma unroll
    for (int c = 0; c < C; c++) {
        const int m = mt * TC + tid + kThreadsO * c;
#pragma unroll
        for (int k = 1; k < M; k++) tw[c][k] = __ldg(&twbig[(m * k) & (NT - 1)]);
        if (WINDOW) {
#pragma unroll
            for (int n1 = 0; n1 < M; n1++) win[c * M + n1] = window[n1 * kNI + m];
        }
    }

    auto issue_load = [&](int i) {            // thread 0: raw bytes of my tile i
        const int slot = i % kRawStgO;
        const uint32_t bar = sm + 8 * slot;
        const uint8_t *src = bytes + (size_t)(part + (int64_t)nparts * i) * (2 * NT) + (size_t)mt * (2 * TC);
        mbar_expect_tx(bar, kTileRaw);
#pragma unroll
        for (int n1 = 0; n1 < M; n1++)
            tma_load_1d(sm + kOffRawO + slot * kTileRaw + n1 * (2 * TC), src + (size_t)n1 * 2 * kNI, 2 * TC, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kRawStgO; s++) mbar_init(sm + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        for (int i = 0; i < kRawStgO && i < ntiles; i++) issue_load(i);
    }
    __syncthreads();

    for (int i = 0; i < ntiles; i++) {
        const int slot = i % kRawStgO;
        mbar_wait(sm + 8 * slot, (uint32_t)((i / kRawStgO) & 1));
        const uint32_t raw = sm + kOffRawO + slot * kTileRaw + 2 * tid;
        uint32_t r[C][M];
#pragma unroll
        for (int c = 0; c < C; c++)
#pragma unroll
            for (int n1 = 0; n1 < M; n1++)
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r[c][n1]) : "r"(raw + n1 * (2 * TC) + 2 * kThreadsO * c));
        const uint32_t out = sm + kOffOutO + (i & 1) * kTileOut + 8 * tid;
#pragma unroll
        for (int c = 0; c < C; c++) {
            c32 v[M];
#pragma unroll
            for (int n1 = 0; n1 < M; n1++) {
                float re = (float)(r[c][n1] & 255u) - 127.0f, im = (float)(r[c][n1] >> 8) - 127.0f;   // src/datastore.cxx:74-75
                if (WINDOW) { re *= win[c * M + n1]; im *= win[c * M + n1]; }                             // :76-77
                v[n1] = make_float2(re, im);
            }
            Dft<M>::run(v);
#pragma unroll
            for (int k = 1; k < M; k++) v[k] = cmul(v[k], tw[c][k]);
#pragma unroll
            for (int k = 0; k < M; k++)
                asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(out + k * (8 * TC) + 8 * kThreadsO * c), "f"(v[k].x), "f"(v[k].y) : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // my staged values -> visible to the bulk stores
        // the other staging buffer is written next: the stores that read it (tile i - 1) must have finished reading
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();                                                   // raw slot read, staging complete
        if (tid == 0) {
            float2 *dst = scratch + (size_t)(part + (int64_t)nparts * i) * NT + (size_t)mt * TC;
#pragma unroll
            for (int k = 0; k < M; k++)
                tma_store_1d(dst + (size_t)k * kNI, sm + kOffOutO + (i & 1) * kTileOut + k * (8 * TC), 8 * TC);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (i + kRawStgO < ntiles) issue_load(i + kRawStgO);
        }
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int M>
cudaError_t launch_m(const uint8_t *bytes, int64_t nspec, const float2 *twbig, const float *window, float2 *scratch,
                     int sms, cudaStream_t stream)
{
    int grid = (2 * sms / M) * M;                      // whole sets of M column ranges
    if (grid < M) grid = M;
    if ((int64_t)grid > nspec * M) grid = (int)(nspec * M);
    if (window)
        fourstep_outer_tma_kernel<M, true><<<grid, kThreadsO, kSmemO, stream>>>(bytes, nspec, twbig, window, scratch);
    else
        fourstep_outer_tma_kernel<M, false><<<grid, kThreadsO, kSmemO, stream>>>(bytes, nspec, twbig, nullptr, scratch);
    return cudaGetLastError();
}

template <int M>
cudaError_t configure_m()
{
    cudaError_t e = cudaFuncSetAttribute(fourstep_outer_tma_kernel<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemO);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(fourstep_outer_tma_kernel<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemO);
}

}  // namespace

cudaError_t rpf_outer_tma_configure()
{
    cudaError_t e;
    if ((e = configure_m<2>()) != cudaSuccess) return e;
    if ((e = configure_m<4>()) != cudaSuccess) return e;
    if ((e = configure_m<8>()) != cudaSuccess) return e;
    return configure_m<16>();
}

cudaError_t rpf_launch_fourstep_outer_tma(int M, const uint8_t *bytes, int64_t nspec, const float2 *twbig,
                                          const float *window, float2 *scratch, int sms, cudaStream_t stream)
{
    if (nspec <= 0) return cudaSuccess;
    switch (M) {
        case 2: return launch_m<2>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 4: return launch_m<4>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 8: return launch_m<8>(bytes, nspec, twbig, window, scratch, sms, stream);
        case 16: return launch_m<16>(bytes, nspec, twbig, window, scratch, sms, stream);
        default: return cudaErrorInvalidValue;
    }
}
