// cluster_probe.cu -- development probe: how many 8- / 16-CTA clusters of the 65536-bin kernel's shape (384 threads,
// ~224 KiB of shared memory) the device keeps resident, the structure of shared::cluster addresses (mapa), and the
// DSMEM store bandwidth of an all-to-all scatter with st.async (8-byte stores, complete_tx on a remote mbarrier).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cluster_probe tools/cluster_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CS>
__global__ void __launch_bounds__(384, 1) scatter_kernel(int iters, uint32_t *out, long long *cycles)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sm = smem_u32(smem);
    const int tid = threadIdx.x;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    // barrier 0: "full" (expects 32 KiB per generation from the whole cluster), data at +1024
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm), "r"(65536u) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm + 8), "r"(65536u) : "memory");
    }
    cluster_sync();
    if (blockIdx.x == 0 && tid == 0)
        for (int r = 0; r < CS; r++) out[r] = mapa(sm, r);
    if (blockIdx.x == 0 && tid == 0) out[16] = sm;
    const long long t0 = clock64();
    if (tid < 256) {
        // per generation every thread sends 8 B to each of 16 (row, cta) pairs: rows = 16 / CS per CTA, 2 "spectra"
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int sp = 0; sp < 2; sp++) {
#pragma unroll
                for (int ko = 0; ko < 16; ko++) {
                    const uint32_t dst_cta = ko % CS, row = (ko / CS) + (16 / CS) * sp;    // rows of 32 KiB? keep it small:
                    const uint32_t local = sm + 1024 + (row & 1) * 32768 + (rank * 256 + tid) * 8 % 32768;
                    const uint32_t ra = mapa(local, dst_cta), rb = mapa(sm + 8 * (it & 1), dst_cta);   // two barriers, alternating: a CTA is at most one generation ahead
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
                                 ::"r"(ra), "r"(tid), "r"(it), "r"(rb) : "memory");
                }
            }
            // every CTA receives 2 spectra x (16 / CS) rows x CS producers x 256 threads x 8 B = 65536 B per generation
            if (tid == 0) {
                uint32_t done;
                do {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(sm + 8 * (it & 1)), "r"((uint32_t)((it >> 1) & 1)) : "memory");
                } while (!done);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm + 8 * (it & 1)), "r"(65536u) : "memory");
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            // (no "slot free" handshake: the probe overwrites the same bytes; only bandwidth is measured)
        }
    }
    const long long t1 = clock64();
    cluster_sync();
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int CS>
static void run(int sms)
{
    const size_t smem = 1024 + 65536 + 160 * 1024;     // ~225 KiB, like the real kernel
    cudaFuncSetAttribute(scatter_kernel<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (CS > 8) cudaFuncSetAttribute(scatter_kernel<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(CS * 32);
    int ncl = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, scatter_kernel<CS>, &cfg);
    printf("cluster size %d: max active clusters %d (%s) -> %d of %d SMs\n", CS, ncl, cudaGetErrorString(e), ncl * CS, sms);
    if (e != cudaSuccess || ncl <= 0) { cudaGetLastError(); return; }
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 64 * 4); cudaMalloc(&cyc, 4096 * 8);
    cfg.gridDim = dim3(CS * ncl);
    const int iters = 2000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        e = cudaLaunchKernelEx(&cfg, scatter_kernel<CS>, iters, out, cyc);
        cudaEventRecord(b);
        cudaError_t e2 = cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        printf("  launch %s / %s: %.3f ms for %d generations of 64 KiB sent per CTA -> %.1f B/clk per SM at 1.965 GHz, %.2f us per generation\n",
               cudaGetErrorString(e), cudaGetErrorString(e2), ms, iters, 65536.0 * iters / (ms * 1e-3 * 1.965e9), ms * 1e3 / iters);
    }
    uint32_t h[17]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("  local %08x, mapa:", h[16]);
    for (int r = 0; r < CS; r++) printf(" %08x", h[r]);
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<8>(p.multiProcessorCount);
    run<16>(p.multiProcessorCount);
    return 0;
}
