// rpf_capi.cu -- the C ABI of include/rpf_b200.h: context lifetime, the
// continuous-stream bookkeeping of Datastore::fftThread (reference
// src/datastore.cxx:48-96: fft_pointer carry-over, repeats limit, truncation)
// and the host->device staging pipeline.  All arithmetic happens in the CUDA
// kernels; there is no CPU fallback -- without a usable CUDA device every
// compute entry point fails with RPF_ERR_CUDA.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../../include/rpf_b200.h"
#include "rpf_internal.h"
#include "synth_gen.h"

namespace {

constexpr int kNumStage = 3;          // staging buffers in flight (copy / compute overlap)
constexpr int kInnerN = 4096;         // inner transform length of the four-step path

enum PlanKind { PLAN_SMALL, PLAN_FUSED3, PLAN_FOURSTEP, PLAN_DIRECT };

thread_local char g_create_err[256] = "";

}  // namespace

struct rpf_ctx {
    int N = 0, device = 0;
    PlanKind plan = PLAN_SMALL;
    int M = 1;                        // four-step outer radix
    int grid = 0, groups = 1, nrows = 0;
    int64_t rowlen = 0;               // doubles per private accumulator row
    size_t spectrum_bytes = 0;        // 2N
    bool has_window = false;
    bool usepipe = false;             // N=4096 (and the four-step inner transform): pipe4096_kernel

    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[kNumStage] = {}, ev_consumed[kNumStage] = {};
    uint8_t *d_stage[kNumStage] = {};
    size_t stage_cap = 0;
    int slot = 0;
    uint8_t *h_carry[kNumStage] = {};   // pinned; carry for push k lives in h_carry[slot]
    size_t carry_bytes = 0;

    float2 *d_twiddle = nullptr;      // W_Ni^j, Ni = inner N
    float2 *d_twiddle_big = nullptr;  // four-step: W_N^j
    float *d_window = nullptr;
    double *d_rows = nullptr, *d_pwr = nullptr, *d_tmp = nullptr, *d_baseline = nullptr;
    float2 *d_scratch = nullptr;
    int64_t scratch_spectra = 0;

    int64_t repeats = INT64_MAX, repeats_done = 0;
    int64_t launches = 0;
    char err[256] = "";
};

namespace {

int fail(rpf_ctx *c, int code, const char *fmt, ...)
{
    char *dst = c ? c->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(c, call)                                                                                   \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return fail((c), RPF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),    \
                        __FILE__, __LINE__);                                                          \
    } while (0)

bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

void fill_twiddle(std::vector<float2> &t, int n)
{
    const double PI = 3.14159265358979323846264338327950288;
    t.resize(n);
    for (int j = 0; j < n; j++) {
        // exact octant reduction is not needed at these sizes; cos/sin of a double
        // argument < 2 pi are accurate to < 1 ulp of float after rounding.
        const double a = -2.0 * PI * (double)j / (double)n;
        t[j] = make_float2((float)cos(a), (float)sin(a));
    }
}

int ensure_stage(rpf_ctx *c, size_t need)
{
    if (need <= c->stage_cap) return RPF_OK;
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    size_t cap = c->stage_cap ? c->stage_cap : (size_t)1 << 20;
    while (cap < need) cap *= 2;
    for (int i = 0; i < kNumStage; i++) {
        if (c->d_stage[i]) CU(c, cudaFree(c->d_stage[i]));
        c->d_stage[i] = nullptr;
        CU(c, cudaMalloc(&c->d_stage[i], cap));
    }
    c->stage_cap = cap;
    return RPF_OK;
}

// Enqueue the FFT work for `nspec` whole spectra starting at the 16-byte
// aligned device pointer `src` on the compute stream.
int launch_spectra(rpf_ctx *c, const uint8_t *src, int64_t nspec)
{
    if (nspec <= 0) return RPF_OK;
    RpfLaunch L;
    L.bytes = src; L.cplx = nullptr; L.nspec = nspec; L.M = 1;
    L.twiddle = c->d_twiddle; L.window = c->d_window; L.rows = c->d_rows;
    L.grid = c->grid; L.stream = c->stream;
    if (c->plan == PLAN_FUSED3) {
        CU(c, c->usepipe ? rpf_launch_pipe4096(L) : rpf_launch_fused3(c->N, L));
        c->launches++;
    } else if (c->plan == PLAN_SMALL) {
        CU(c, rpf_launch_small(c->N, L));
        c->launches++;
    } else if (c->plan == PLAN_DIRECT) {
        CU(c, rpf_launch_direct_dft(c->N, L));
        c->launches++;
    } else {
        for (int64_t done = 0; done < nspec; done += c->scratch_spectra) {
            const int64_t b = (nspec - done) < c->scratch_spectra ? (nspec - done) : c->scratch_spectra;
            CU(c, rpf_launch_fourstep_outer(c->M, src + (size_t)done * c->spectrum_bytes, b, c->d_twiddle_big,
                                            c->d_window, c->d_scratch, c->stream));
            RpfLaunch Li = L;
            Li.bytes = nullptr; Li.cplx = c->d_scratch; Li.nspec = b; Li.M = c->M; Li.window = nullptr;
            CU(c, rpf_launch_pipe4096(Li));
            c->launches += 2;
        }
    }
    return RPF_OK;
}

}  // namespace

extern "C" {

int rpf_abi_version(void) { return RPF_B200_ABI_VERSION; }

const char *rpf_strerror(const rpf_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

int rpf_create(const rpf_config *cfg, rpf_ctx **out)
{
    if (!cfg || !out) return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: NULL argument");
    *out = nullptr;
    const int N = cfg->N;
    if (N < 2 || (N & 1))
        return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: N=%d must be even and >= 2 (src/params.cxx:152-155)", N);
    if ((!is_pow2(N) && N > 8190) || N > 16 * kInnerN)
        return fail(nullptr, RPF_ERR_UNSUPPORTED_N,
                    "rpf_create: N=%d not supported by the CUDA path (powers of two up to 65536, other even N up to 8190)", N);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, RPF_ERR_CUDA, "rpf_create: no usable CUDA device (%s); there is no CPU fallback",
                    cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: device %d out of range [0,%d)", cfg->device, ndev);

    rpf_ctx *c = new (std::nothrow) rpf_ctx;
    if (!c) return fail(nullptr, RPF_ERR_NOMEM, "rpf_create: out of host memory");
    c->N = N; c->device = cfg->device;
    c->spectrum_bytes = (size_t)2 * N;
    c->has_window = cfg->window != nullptr;
    int rc = RPF_OK;
    auto body = [&]() -> int {
        CU(c, cudaSetDevice(c->device));
        int sms = 0;
        CU(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        CU(c, cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        CU(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        for (int i = 0; i < kNumStage; i++) {
            CU(c, cudaEventCreateWithFlags(&c->ev_copied[i], cudaEventDisableTiming));
            CU(c, cudaEventCreateWithFlags(&c->ev_consumed[i], cudaEventDisableTiming));
            CU(c, cudaHostAlloc((void **)&c->h_carry[i], c->spectrum_bytes, cudaHostAllocDefault));
        }
        int inner = N;
        // development A/B switch only: RPF_B200_KERNEL=generic runs N=4096 on the generic 3-pass kernel
        const char *kname = getenv("RPF_B200_KERNEL");
        const bool generic = kname && !strcmp(kname, "generic");
        if (!is_pow2(N)) {
            c->plan = PLAN_DIRECT; c->groups = 1; c->grid = 4 * sms;
            CU(c, rpf_direct_dft_configure());
        } else if (rpf_fused3_supported(N)) {
            c->plan = PLAN_FUSED3;
            c->usepipe = (N == kInnerN) && !generic;
            c->groups = c->usepipe ? 1 : rpf_fused3_groups(N);
            c->grid = c->usepipe ? sms : 2 * sms;
        } else if (N <= 256) {
            c->plan = PLAN_SMALL; c->groups = 1; c->grid = 8 * sms;
        } else {
            // inner transform = pipe4096 in complex-source mode: every CTA is bound to one outer bin
            c->plan = PLAN_FOURSTEP; c->M = N / kInnerN; inner = kInnerN;
            c->usepipe = true; c->groups = 1;
            c->grid = (sms / c->M) * c->M;
            size_t scratch_mb = 256;                                  // batch of complex64 rows (measured: 64 -> 171, 256 -> 205 Gs/s)
            if (const char *e = getenv("RPF_B200_SCRATCH_MB")) scratch_mb = (size_t)atoi(e) > 0 ? (size_t)atoi(e) : 256;
            c->scratch_spectra = (scratch_mb << 20) / ((size_t)N * sizeof(float2));
            if (c->scratch_spectra < 1) c->scratch_spectra = 1;
            CU(c, cudaMalloc(&c->d_scratch, (size_t)c->scratch_spectra * N * sizeof(float2)));
            std::vector<float2> tb;
            fill_twiddle(tb, N);
            CU(c, cudaMalloc(&c->d_twiddle_big, sizeof(float2) * N));
            CU(c, cudaMemcpy(c->d_twiddle_big, tb.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
        }
        if (c->plan != PLAN_SMALL && c->plan != PLAN_DIRECT) {
            CU(c, rpf_fused3_configure());
            CU(c, rpf_pipe4096_configure());
        }
        std::vector<float2> tw;
        fill_twiddle(tw, inner);
        CU(c, cudaMalloc(&c->d_twiddle, sizeof(float2) * inner));
        CU(c, cudaMemcpy(c->d_twiddle, tw.data(), sizeof(float2) * inner, cudaMemcpyHostToDevice));
        if (cfg->window) {
            CU(c, cudaMalloc(&c->d_window, sizeof(float) * N));
            CU(c, cudaMemcpy(c->d_window, cfg->window, sizeof(float) * N, cudaMemcpyHostToDevice));
        }
        c->nrows = c->grid * c->groups;
        c->rowlen = c->plan == PLAN_FOURSTEP ? kInnerN : N;
        CU(c, cudaMalloc(&c->d_rows, sizeof(double) * (size_t)c->nrows * c->rowlen));
        CU(c, cudaMalloc(&c->d_pwr, sizeof(double) * N));
        CU(c, cudaMalloc(&c->d_tmp, sizeof(double) * 2 * N));
        CU(c, cudaMalloc(&c->d_baseline, sizeof(double) * N));
        CU(c, cudaMemset(c->d_rows, 0, sizeof(double) * (size_t)c->nrows * c->rowlen));
        size_t hint = cfg->buf_length > 0 ? (size_t)cfg->buf_length : (size_t)16384 * 100;   // src/params.h:43
        int r = ensure_stage(c, hint + c->spectrum_bytes);
        if (r) return r;
        CU(c, cudaDeviceSynchronize());
        return RPF_OK;
    };
    rc = body();
    if (rc != RPF_OK) {
        snprintf(g_create_err, sizeof(g_create_err), "%s", c->err);
        rpf_destroy(c);
        return rc;
    }
    *out = c;
    return RPF_OK;
}

void rpf_destroy(rpf_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
    for (int i = 0; i < kNumStage; i++) {
        if (c->d_stage[i]) cudaFree(c->d_stage[i]);
        if (c->h_carry[i]) cudaFreeHost(c->h_carry[i]);
        if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]);
        if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]);
    }
    cudaFree(c->d_twiddle); cudaFree(c->d_twiddle_big); cudaFree(c->d_window);
    cudaFree(c->d_rows); cudaFree(c->d_pwr); cudaFree(c->d_tmp); cudaFree(c->d_baseline); cudaFree(c->d_scratch);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    delete c;
}

int rpf_reset(rpf_ctx *c, int64_t repeats)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (repeats < 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_reset: repeats must be >= 0");
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaMemsetAsync(c->d_rows, 0, sizeof(double) * (size_t)c->nrows * c->rowlen, c->stream));
    c->repeats = repeats;
    c->repeats_done = 0;
    c->carry_bytes = 0;
    return RPF_OK;
}

int rpf_push(rpf_ctx *c, const uint8_t *bytes, size_t len)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (len == 0) return RPF_OK;
    if (!bytes) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push: NULL buffer");
    if (len & 1) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push: odd byte count %zu (I/Q pairs, src/datastore.cxx:74)", len);
    if (c->repeats_done >= c->repeats) return RPF_OK;          // src/datastore.cxx:67: discarded
    CU(c, cudaSetDevice(c->device));

    const size_t sb = c->spectrum_bytes;
    const size_t total = c->carry_bytes + len;
    const int s = c->slot, sn = (c->slot + 1) % kNumStage;
    if (total < sb) {                                            // still inside one spectrum
        memcpy(c->h_carry[sn], c->h_carry[s], c->carry_bytes);
        memcpy(c->h_carry[sn] + c->carry_bytes, bytes, len);
        c->carry_bytes = total;
        c->slot = sn;
        return RPF_OK;
    }
    int r = ensure_stage(c, total);
    if (r) return r;
    const int64_t avail = (int64_t)(total / sb);
    const int64_t left = c->repeats - c->repeats_done;
    const int64_t nspec = avail < left ? avail : left;
    const size_t used = (size_t)nspec * sb;                      // bytes the kernels will read
    // copy stream: wait until the kernels that last read this staging buffer are done
    CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[s], 0));
    if (c->carry_bytes)
        CU(c, cudaMemcpyAsync(c->d_stage[s], c->h_carry[s], c->carry_bytes, cudaMemcpyHostToDevice, c->copy_stream));
    const size_t copy_new = used - c->carry_bytes < len ? used - c->carry_bytes : len;
    CU(c, cudaMemcpyAsync(c->d_stage[s] + c->carry_bytes, bytes, copy_new, cudaMemcpyHostToDevice, c->copy_stream));
    CU(c, cudaEventRecord(c->ev_copied[s], c->copy_stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_copied[s], 0));
    r = launch_spectra(c, c->d_stage[s], nspec);
    if (r) return r;
    CU(c, cudaEventRecord(c->ev_consumed[s], c->stream));
    c->repeats_done += nspec;
    if (c->repeats_done >= c->repeats) {
        c->carry_bytes = 0;                                      // rest of the stream is discarded
    } else {
        const size_t tail = total - used;                        // < sb; lies entirely in `bytes`
        memcpy(c->h_carry[sn], bytes + (len - tail), tail);
        c->carry_bytes = tail;
    }
    c->slot = sn;
    CU(c, cudaEventSynchronize(c->ev_copied[s]));                // caller may now reuse `bytes`
    return RPF_OK;
}

int rpf_push_device(rpf_ctx *c, const uint8_t *bytes, size_t len)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (len == 0) return RPF_OK;
    if (!bytes) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push_device: NULL buffer");
    if (len & 1) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push_device: odd byte count %zu", len);
    if (c->repeats_done >= c->repeats) return RPF_OK;
    CU(c, cudaSetDevice(c->device));
    const size_t sb = c->spectrum_bytes;
    const int64_t left = c->repeats - c->repeats_done;

    if (c->carry_bytes == 0 && (((uintptr_t)bytes) & 15) == 0) {
        // zero-copy: spectra are read in place by TMA
        const int64_t avail = (int64_t)(len / sb);
        const int64_t nspec = avail < left ? avail : left;
        int r = launch_spectra(c, bytes, nspec);
        if (r) return r;
        c->repeats_done += nspec;
        if (c->repeats_done >= c->repeats) return RPF_OK;
        const size_t tail = len - (size_t)nspec * sb;
        if (tail) {
            const int sn = c->slot;
            CU(c, cudaMemcpyAsync(c->h_carry[sn], bytes + (len - tail), tail, cudaMemcpyDeviceToHost, c->stream));
            CU(c, cudaStreamSynchronize(c->stream));
            c->carry_bytes = tail;
        }
        return RPF_OK;
    }
    // general case: assemble carry + buffer contiguously in staging (device-to-device)
    const size_t total = c->carry_bytes + len;
    const int s = c->slot, sn = (c->slot + 1) % kNumStage;
    if (total < sb) {
        memcpy(c->h_carry[sn], c->h_carry[s], c->carry_bytes);
        CU(c, cudaMemcpyAsync(c->h_carry[sn] + c->carry_bytes, bytes, len, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        c->carry_bytes = total; c->slot = sn;
        return RPF_OK;
    }
    int r = ensure_stage(c, total);
    if (r) return r;
    const int64_t avail = (int64_t)(total / sb);
    const int64_t nspec = avail < left ? avail : left;
    const size_t used = (size_t)nspec * sb;
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_consumed[s], 0));
    if (c->carry_bytes)
        CU(c, cudaMemcpyAsync(c->d_stage[s], c->h_carry[s], c->carry_bytes, cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaMemcpyAsync(c->d_stage[s] + c->carry_bytes, bytes, used - c->carry_bytes, cudaMemcpyDeviceToDevice, c->stream));
    r = launch_spectra(c, c->d_stage[s], nspec);
    if (r) return r;
    CU(c, cudaEventRecord(c->ev_consumed[s], c->stream));
    c->repeats_done += nspec;
    if (c->repeats_done >= c->repeats) {
        c->carry_bytes = 0;
    } else {
        const size_t tail = total - used;
        if (tail) CU(c, cudaMemcpyAsync(c->h_carry[sn], bytes + (len - tail), tail, cudaMemcpyDeviceToHost, c->stream));
        CU(c, cudaStreamSynchronize(c->stream));
        c->carry_bytes = tail;
    }
    c->slot = sn;
    return RPF_OK;
}

int rpf_sync(rpf_ctx *c)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RPF_OK;
}

int rpf_finish_device(rpf_ctx *c, double *pwr_out_device, int64_t *repeats_done)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    if (pwr_out_device) {
        if (c->plan == PLAN_FOURSTEP)
            CU(c, rpf_launch_reduce_rows_fourstep(c->d_rows, c->nrows, c->M, kInnerN, pwr_out_device, c->stream));
        else
            CU(c, rpf_launch_reduce_rows(c->d_rows, c->nrows, c->N, pwr_out_device, c->stream));
        c->launches++;
    }
    if (repeats_done) *repeats_done = c->repeats_done;
    return RPF_OK;
}

int rpf_finish(rpf_ctx *c, double *pwr_out, int64_t *repeats_done)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    int r = rpf_finish_device(c, pwr_out ? c->d_pwr : nullptr, repeats_done);
    if (r) return r;
    if (pwr_out)
        CU(c, cudaMemcpyAsync(pwr_out, c->d_pwr, sizeof(double) * c->N, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RPF_OK;
}

int64_t rpf_repeats_done(const rpf_ctx *c) { return c ? c->repeats_done : 0; }
int64_t rpf_fft_pointer(const rpf_ctx *c) { return c ? (int64_t)(c->carry_bytes / 2) : 0; }
int64_t rpf_kernel_launches(const rpf_ctx *c) { return c ? c->launches : 0; }

const char *rpf_plan_name(const rpf_ctx *c)
{
    if (!c) return "";
    return c->plan == PLAN_SMALL ? "small" : c->plan == PLAN_FUSED3 ? "fused3" : c->plan == PLAN_DIRECT ? "direct" : "fourstep";
}

int rpf_set_stream(rpf_ctx *c, void *cuda_stream)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    CU(c, cudaSetDevice(c->device));
    CU(c, cudaStreamSynchronize(c->stream));
    CU(c, cudaStreamSynchronize(c->copy_stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return RPF_OK;
}

int rpf_write_data(rpf_ctx *c, double tuned_freq, int32_t samplerate, int32_t linear, const double *baseline,
                   double *freq_out, double *val_out)
{
    if (!c || !val_out) return RPF_ERR_INVALID_ARG;
    if (samplerate <= 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_write_data: samplerate must be > 0");
    CU(c, cudaSetDevice(c->device));
    const int N = c->N;
    if (c->plan == PLAN_FOURSTEP)
        CU(c, rpf_launch_reduce_rows_fourstep(c->d_rows, c->nrows, c->M, kInnerN, c->d_pwr, c->stream));
    else
        CU(c, rpf_launch_reduce_rows(c->d_rows, c->nrows, N, c->d_pwr, c->stream));
    if (baseline)
        CU(c, cudaMemcpyAsync(c->d_baseline, baseline, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
    CU(c, rpf_launch_write_data(c->d_pwr, N, c->repeats_done, tuned_freq, samplerate, linear,
                                baseline ? c->d_baseline : nullptr, c->d_tmp, c->d_tmp + N, c->stream));
    c->launches += 2;
    if (freq_out) CU(c, cudaMemcpyAsync(freq_out, c->d_tmp, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaMemcpyAsync(val_out, c->d_tmp + N, sizeof(double) * N, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    return RPF_OK;
}

void rpf_synth_fill_host(uint8_t *bytes, size_t len, uint64_t off, uint64_t seed, int32_t mode)
{
    size_t i = 0;
    while (i < len && ((off + i) & 7)) { bytes[i] = rpf_synth_byte(seed, mode, off + i); i++; }
    for (; i + 8 <= len; i += 8) {
        const uint64_t v = rpf_synth_word(seed, mode, (off + i) >> 3);
        memcpy(bytes + i, &v, 8);     // little-endian host, as on the device
    }
    for (; i < len; i++) bytes[i] = rpf_synth_byte(seed, mode, off + i);
}

int rpf_synth_fill_device(uint8_t *bytes, size_t len, uint64_t off, uint64_t seed, int32_t mode, int32_t device,
                          void *cuda_stream)
{
    if (!bytes && len) return RPF_ERR_INVALID_ARG;
    if (mode < 0 || mode > 2) return RPF_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return RPF_ERR_CUDA;
    return rpf_launch_synth(bytes, len, off, seed, mode, (cudaStream_t)cuda_stream) == cudaSuccess ? RPF_OK : RPF_ERR_CUDA;
}

int rpf_host_alloc(void **out, size_t bytes)
{
    if (!out) return RPF_ERR_INVALID_ARG;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? RPF_OK : RPF_ERR_CUDA;
}
void rpf_host_free(void *p) { if (p) cudaFreeHost(p); }
int rpf_host_register(void *p, size_t bytes)
{
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? RPF_OK : RPF_ERR_CUDA;
}
int rpf_host_unregister(void *p) { return cudaHostUnregister(p) == cudaSuccess ? RPF_OK : RPF_ERR_CUDA; }

}  // extern "C"
