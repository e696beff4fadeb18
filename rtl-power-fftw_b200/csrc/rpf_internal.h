// rpf_internal.h -- launch interfaces between rpf_capi.cu and the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// Source of one launch of the fused kernels.
struct RpfLaunch {
    const uint8_t *bytes;     // SRC_U8: 16-byte aligned start of spectrum 0 (2N bytes per spectrum)
    const float2  *cplx;      // SRC_C32: four-step scratch, [spec][k1][n] complex64
    int64_t        nspec;     // spectra (SRC_U8) / spectra in the batch S (SRC_C32)
    int            M;         // four-step outer radix (1 for the direct path)
    const float2  *twiddle;   // W_N^j, j in [0,N): (cos, -sin)
    const float   *window;    // N floats or nullptr
    double        *rows;      // private accumulator rows [grid*G][N*M]
    int            grid;
    cudaStream_t   stream;
};

// Fused unpack -> (window) -> N-point FFT -> |X|^2 accumulate, N in {512,1024,2048,4096}.
cudaError_t rpf_launch_fused3(int N, const RpfLaunch &L);
int rpf_fused3_groups(int N);          // spectra per CTA iteration (G)
bool rpf_fused3_supported(int N);
cudaError_t rpf_fused3_configure();    // opt-in shared memory attributes (once per device)

// N = 512, 1024, 2048: the two-stream pipeline on the generic 3-pass geometry (one CTA per SM): the product path.
cudaError_t rpf_launch_pipen(int N, const RpfLaunch &L);
int rpf_pipen_groups(int N);
bool rpf_pipen_supported(int N);
cudaError_t rpf_pipen_configure();

// N = 4096, two-stream software pipeline (256-thread CTAs, two spectra per thread): the product path.
cudaError_t rpf_launch_pipe4096(const RpfLaunch &L);
cudaError_t rpf_pipe4096_configure();

// N = 2048 / 1024 / 512 on the pipe4096 loop: 2 / 4 / 8 consecutive spectra share a 4096-sample ring slot, pass 1 runs that
// many radix-8 / 4 / 2 butterflies per thread, passes 2 and 3 are the 4096 kernel's.  L.twiddle = the W_4096 table,
// L.rows = (4096 / N) private rows of N doubles per CTA.
bool rpf_pipe4096_small_supported(int N);
cudaError_t rpf_launch_pipe4096_small(int N, const RpfLaunch &L);

// N = 65536 without a window in one launch: 16-CTA clusters, the outer radix-16 handed over through distributed
// shared memory (st.async + remote mbarriers) instead of a complex64 scratch in HBM.  L.grid = 16 * clusters.
cudaError_t rpf_cluster65536_configure(int *max_clusters);
cudaError_t rpf_launch_cluster65536(const RpfLaunch &L, const float2 *twiddle_big);

// N = M * 4096 (M = 2 ... 16) in one launch, one CTA per spectrum: thread-private hand-off of the outer radix-M through an
// L2-resident scratch (rpf_solo65536_scratch_bytes(M, grid) bytes), accumulators in tensor memory.  L.M = M, L.grid = CTAs,
// L.rows = M rows of 4096 doubles per CTA (row M c + ko).
cudaError_t rpf_solo65536_configure(int M, size_t *saved_persisting_l2_limit);   // M >= 8: raises the device's persisting-L2 set-aside
void rpf_solo65536_release(size_t saved_persisting_l2_limit);              // ... and puts it back
size_t rpf_solo65536_scratch_bytes(int M, int grid);
cudaError_t rpf_launch_solo65536(const RpfLaunch &L, const float2 *twiddle_big, float2 *scratch);

// N = 4096 with pass 1 on the tensor cores (tcgen05.mma, fp16 hi+lo DFT-16 matrix; raw bytes, no window): experiment
// selected with RPF_FLAG_TC_PASS1.
cudaError_t rpf_launch_pipe4096_tc(const RpfLaunch &L);
cudaError_t rpf_pipe4096_tc_configure();

// Generic power-of-two N in [2,256]: shared-memory radix-2 Stockham.
cudaError_t rpf_launch_small(int N, const RpfLaunch &L);

// Even N that is not a power of two (2 < N <= 8190): O(N^2) DFT per spectrum.
cudaError_t rpf_launch_direct_dft(int N, const RpfLaunch &L);
cudaError_t rpf_direct_dft_configure();

// Any other even N in O(N log N): Bluestein over M = 2^m >= 2N - 1 (N not a power of two) or a plain FFT_N
// (N = 2^m > 65536), batched global-memory Stockham passes.  Buffers and tables are owned by the caller.
struct RpfBluestein {
    int N, M;                 // M = rpf_bluestein_M(N)
    const float2 *chirp;      // c[n] = exp(-i pi n^2 / N), n < N            (unused when M == N)
    const float2 *bhat;       // FFT_M(conj c, wrapped) / M                  (unused when M == N)
    const float2 *twM;        // W_M^j, j < M
    const float  *window;     // N floats or nullptr
    float2 *buf0, *buf1;      // 2 x batch x M complex64
    double *rows;             // [slices][N] private accumulator rows
    int slices;
    int sms;
    cudaStream_t stream;
};
int rpf_bluestein_M(int N);
cudaError_t rpf_launch_bluestein(const RpfBluestein &B, const uint8_t *bytes, int64_t nspec, int *launches);

// Four-step outer pass for N = M*4096: radix-M over the stride-4096 digit + twiddle.
cudaError_t rpf_launch_fourstep_outer(int M, const uint8_t *bytes, int64_t nspec, const float2 *twiddle_big,
                                      const float *window, float2 *scratch, int sms, cudaStream_t stream);

// The same pass as a TMA-staged streaming kernel (bulk loads -> shared-memory ring -> bulk stores): the outer pass of the
// two-kernel path (launches of fewer than 96 spectra, RPF_FLAG_FOURSTEP).
cudaError_t rpf_launch_fourstep_outer_tma(int M, const uint8_t *bytes, int64_t nspec, const float2 *twiddle_big,
                                          const float *window, float2 *scratch, int sms, cudaStream_t stream);
cudaError_t rpf_outer_tma_configure();

// pwr[i] = sum over rows [0, nrows) (fixed order, float64).
cudaError_t rpf_launch_reduce_rows(const double *rows, int nrows, int64_t rowlen, double *pwr, cudaStream_t stream);

cudaError_t rpf_launch_reduce_rows_fourstep(const double *rows, int nrows, int M, int inner, double *pwr, cudaStream_t stream);

cudaError_t rpf_launch_synth(uint8_t *bytes, size_t len, uint64_t stream_offset, uint64_t seed, int mode,
                             int sms, cudaStream_t stream);

cudaError_t rpf_launch_write_data(const double *pwr, int N, int64_t repeats_done, double tuned_freq,
                                  int samplerate, int linear, const double *baseline, double *freq_out,
                                  double *val_out, float *row_f32_out, cudaStream_t stream);
