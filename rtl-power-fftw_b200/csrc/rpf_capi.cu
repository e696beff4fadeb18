// rpf_capi.cu -- the C ABI of include/rpf_b200.h: context lifetime, the
// continuous-stream bookkeeping of Datastore::fftThread (reference
// src/datastore.cxx:48-96: fft_pointer carry-over, repeats limit, truncation),
// the host->device staging pipeline, the dealing of whole spectra over the
// context's GPUs and the final NCCL allreduce of pwr[].  All arithmetic happens
// in the CUDA kernels; there is no CPU fallback -- without a usable CUDA device
// every compute entry point fails with RPF_ERR_CUDA.
#include <dlfcn.h>
#include <math.h>
#include <nccl.h>          // types and enums only: the library is dlopen()ed, see NcclApi
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <algorithm>
#include <new>
#include <vector>

#include "../../include/rpf_b200.h"
#include "rpf_internal.h"
#include "synth_gen.h"

namespace {

constexpr int kNumStage = 3;               // staging buffers in flight per device (copy / compute overlap)
constexpr int kInnerN = 4096;              // inner transform length of the four-step path
constexpr size_t kSmallPush = 256 << 10;   // pushes up to this size go through a pinned bounce buffer (no host wait)
constexpr size_t kMinPartBytes = 1 << 20;  // a push is split over the devices only in parts of at least this size

enum PlanKind { PLAN_SMALL, PLAN_FUSED3, PLAN_FOURSTEP, PLAN_DIRECT, PLAN_BLUESTEIN };
constexpr int kMaxBins = 1 << 22;          // largest N accepted (pwr[] of 32 MiB; the reference has no limit but memory)
constexpr int kDirectMaxN = 128;           // non-power-of-two N up to here: O(N^2) kernel; above: Bluestein
constexpr int kBlueSlices = 64;            // private accumulator rows of the Bluestein path (at most)

thread_local char g_create_err[256] = "";

// stdout is the caller's data channel (the reference prints its spectra there, src/acquisition.cxx:391-399).  NCCL prints
// a version banner to stdout during communicator creation whenever NCCL_DEBUG is set (and ignores NCCL_DEBUG_FILE for
// it on some versions): while a communicator is being created fd 1 points at stderr.
struct StdoutToStderr {
    int saved = -1;
    StdoutToStderr() { fflush(stdout); saved = dup(1); if (saved >= 0) dup2(2, 1); }
    ~StdoutToStderr() { if (saved >= 0) { fflush(stdout); dup2(saved, 1); close(saved); } }
};

// ---- NCCL, resolved at run time ------------------------------------------------------------------
// Single-GPU use must not require libnccl, and a process that already carries an NCCL (PyTorch bundles
// its own) must not get a second copy: the already loaded library is preferred (RTLD_NOLOAD), then the
// system one.  Only the handful of entry points SURVEY.md 8(e) needs.
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};

NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
    api.handle = h;
    bool all = true;
    auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) all = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    api.ok = all;
    return api.ok ? &api : nullptr;
}

// Everything that lives on one GPU.
struct rpf_dev {
    int device = 0, sms = 0;
    int grid = 0, nrows = 0;
    int cluster_grid = 0;             // N = 65536: CTAs of the resident 16-CTA clusters (0: two-kernel four-step path)
    int solo_grid = 0;                // N = 65536: CTAs of the one-CTA-per-spectrum kernel (0: not selected)
    int64_t solo_min = 0;             // ... which takes the launches of at least this many spectra
    float2 *d_solo = nullptr;         // ... and their private scratch, 512 KiB each
    size_t solo_saved_l2 = (size_t)-1;   // the device's persisting-L2 limit before this context raised it
    int rows_dirty = 0;               // private accumulator rows written since the last reset
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_copied[kNumStage] = {}, ev_consumed[kNumStage] = {};
    uint8_t *d_stage[kNumStage] = {};
    uint8_t *h_bounce[kNumStage] = {};   // pinned, kSmallPush + 2N bytes each
    size_t stage_cap = 0;
    int slot = 0;
    float2 *d_twiddle = nullptr;      // W_Ni^j, Ni = inner N
    float2 *d_twiddle_big = nullptr;  // four-step: W_N^j
    float *d_window = nullptr;
    double *d_rows = nullptr, *d_pwr = nullptr, *d_tmp = nullptr, *d_baseline = nullptr;
    float2 *d_scratch = nullptr;
    int64_t scratch_cap = 0;          // spectra the scratch holds now (grows on demand up to rpf_ctx::scratch_spectra)
    float2 *d_chirp = nullptr, *d_bhat = nullptr, *d_blue[2] = {};   // Bluestein tables and ping-pong buffers
    ncclComm_t comm = nullptr;
    int64_t launches = 0;
};

}  // namespace

struct rpf_ctx {
    int N = 0;
    PlanKind plan = PLAN_SMALL;
    int M = 1;                        // four-step outer radix
    int groups = 1;
    int64_t rowlen = 0;               // doubles per private accumulator row
    size_t spectrum_bytes = 0;        // 2N
    bool has_window = false;
    bool usepipe = false;             // N=4096 (and the four-step inner transform): pipe4096_kernel
    bool use_tc = false;              // RPF_FLAG_TC_PASS1 and N=4096 without window: pipe4096_tc_kernel
    bool old_outer = false;           // RPF_FLAG_LOCKSTEP and N > 4096: the round-1 outer-pass kernel (A/B comparisons)
    bool use_batched = false;         // N = 512, 1024, 2048 on the pipe4096 loop (4096 / N spectra per ring slot)
    bool use_pipen = false;           // N = 512, 1024, 2048: pipen_kernel (unless RPF_FLAG_LOCKSTEP)
    bool pipelined = false;           // RPF_FLAG_PIPELINED_PUSH
    int64_t scratch_spectra = 0;
    int blueM = 0;                    // PLAN_BLUESTEIN: transform length M (= N for a plain power-of-two FFT)
    int64_t blue_batch = 0;           // spectra per pass over the ping-pong buffers

    std::vector<rpf_dev> devs;        // devs[0] is the primary: results are read from it
    int cursor = 0;                   // next device of the round-robin
    int comm_rank = 0, comm_nranks = 1;   // multi-process communicator (rpf_comm_init_rank); 1 = none

    uint8_t *h_carry[kNumStage] = {};   // pinned; the partial spectrum kept for the next push
    double *h_pwr = nullptr;          // pinned + device-visible: single-GPU rpf_finish reduces straight into it
    int cslot = 0;
    size_t carry_bytes = 0;
    struct Pending { int dev, slot; };
    std::vector<Pending> pending;     // host->device copies of the last push that were not waited for yet

    int64_t repeats = INT64_MAX, repeats_done = 0;
    bool early_reduced = false;       // the push that completed the integration has already enqueued the reduction into h_pwr
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

#define NC(c, api, call)                                                                              \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                                        \
            return fail((c), RPF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(r_), \
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

// float64 radix-2 FFT (forward) for the Bluestein filter table; n = re.size() is a power of two
void host_fft_pow2(std::vector<double> &re, std::vector<double> &im)
{
    const size_t n = re.size();
    const double PI = 3.14159265358979323846264338327950288;
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) { std::swap(re[i], re[j]); std::swap(im[i], im[j]); }
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1;
        for (size_t k = 0; k < half; k++) {
            const double a = -2.0 * PI * (double)k / (double)len, wr = cos(a), wi = sin(a);
            for (size_t base = 0; base < n; base += len) {
                const size_t p = base + k, q = p + half;
                const double tr = re[q] * wr - im[q] * wi, ti = re[q] * wi + im[q] * wr;
                re[q] = re[p] - tr; im[q] = im[p] - ti;
                re[p] += tr; im[p] += ti;
            }
        }
    }
}

int ensure_stage(rpf_ctx *c, rpf_dev &d, size_t need)
{
    if (need <= d.stage_cap) return RPF_OK;
    CU(c, cudaStreamSynchronize(d.stream));
    CU(c, cudaStreamSynchronize(d.copy_stream));
    size_t cap = d.stage_cap ? d.stage_cap : (size_t)1 << 20;
    while (cap < need) cap *= 2;
    for (int i = 0; i < kNumStage; i++) {
        if (d.d_stage[i]) CU(c, cudaFree(d.d_stage[i]));
        d.d_stage[i] = nullptr;
        CU(c, cudaMalloc(&d.d_stage[i], cap));
    }
    d.stage_cap = cap;
    return RPF_OK;
}

// Private accumulator rows a launch over `nspec` spectra writes (rows [0, result) of d_rows).
int rows_touched(const rpf_ctx *c, const rpf_dev &d, int64_t nspec)
{
    int64_t blocks;
    switch (c->plan) {
        case PLAN_FUSED3:
            blocks = c->usepipe ? (nspec + 1) / 2 : (nspec + c->groups - 1) / c->groups;
            if (c->use_pipen || c->use_batched) blocks = (blocks + 1) / 2;   // pairs of group iterations / ring slots
            break;
        case PLAN_FOURSTEP:
            if (d.solo_grid && nspec >= d.solo_min)
                return (int)(nspec < d.solo_grid ? nspec : d.solo_grid) * c->M;              // M rows per CTA that ran
            return d.grid;
        case PLAN_BLUESTEIN: return d.nrows;
        default: blocks = nspec;
    }
    if (blocks > d.grid) blocks = d.grid;
    return (int)blocks * c->groups;
}

// Enqueue the FFT work for `nspec` whole spectra starting at the 16-byte
// aligned device pointer `src` on the device's compute stream.
int launch_spectra(rpf_ctx *c, rpf_dev &d, const uint8_t *src, int64_t nspec)
{
    if (nspec <= 0) return RPF_OK;
    RpfLaunch L;
    L.bytes = src; L.cplx = nullptr; L.nspec = nspec; L.M = 1;
    L.twiddle = d.d_twiddle; L.window = d.d_window; L.rows = d.d_rows;
    L.grid = d.grid; L.stream = d.stream;
    if (c->plan == PLAN_FUSED3) {
        CU(c, c->use_tc ? rpf_launch_pipe4096_tc(L) : c->usepipe ? rpf_launch_pipe4096(L)
              : c->use_batched ? rpf_launch_pipe4096_small(c->N, L)
              : c->use_pipen ? rpf_launch_pipen(c->N, L) : rpf_launch_fused3(c->N, L));
        d.launches++;
    } else if (c->plan == PLAN_SMALL) {
        CU(c, rpf_launch_small(c->N, L));
        d.launches++;
    } else if (c->plan == PLAN_DIRECT) {
        CU(c, rpf_launch_direct_dft(c->N, L));
        d.launches++;
    } else if (c->plan == PLAN_BLUESTEIN) {
        RpfBluestein B;
        B.N = c->N; B.M = c->blueM; B.chirp = d.d_chirp; B.bhat = d.d_bhat; B.twM = d.d_twiddle; B.window = d.d_window;
        B.buf0 = d.d_blue[0]; B.buf1 = d.d_blue[1]; B.rows = d.d_rows; B.slices = d.grid; B.sms = d.sms;
        B.stream = d.stream;
        for (int64_t done = 0; done < nspec; done += c->blue_batch) {
            const int64_t b = (nspec - done) < c->blue_batch ? (nspec - done) : c->blue_batch;
            int n = 0;
            CU(c, rpf_launch_bluestein(B, src + (size_t)done * c->spectrum_bytes, b, &n));
            d.launches += n;
        }
    } else if (d.solo_grid && nspec >= d.solo_min) {
        RpfLaunch Ls = L;
        Ls.grid = d.solo_grid; Ls.M = c->M;
        CU(c, rpf_launch_solo65536(Ls, d.d_twiddle_big, d.d_solo));
        d.launches++;
    } else if (d.cluster_grid) {
        RpfLaunch Lc = L;
        Lc.grid = d.cluster_grid;
        CU(c, rpf_launch_cluster65536(Lc, d.d_twiddle_big));
        d.launches++;
    } else {
        // the complex64 scratch grows with the largest launch seen, up to the batch size: a short integration does not pay
        // for (or wait on the allocation of) a long one's buffer
        const int64_t want = nspec < c->scratch_spectra ? nspec : c->scratch_spectra;
        if (want > d.scratch_cap) {
            CU(c, cudaStreamSynchronize(d.stream));
            if (d.d_scratch) CU(c, cudaFree(d.d_scratch));
            d.d_scratch = nullptr; d.scratch_cap = 0;
            CU(c, cudaMalloc(&d.d_scratch, (size_t)want * c->N * sizeof(float2)));
            d.scratch_cap = want;
        }
        for (int64_t done = 0; done < nspec; done += d.scratch_cap) {
            const int64_t b = (nspec - done) < d.scratch_cap ? (nspec - done) : d.scratch_cap;
            CU(c, (c->old_outer ? rpf_launch_fourstep_outer : rpf_launch_fourstep_outer_tma)(
                      c->M, src + (size_t)done * c->spectrum_bytes, b, d.d_twiddle_big, d.d_window, d.d_scratch, d.sms, d.stream));
            RpfLaunch Li = L;
            Li.bytes = nullptr; Li.cplx = d.d_scratch; Li.nspec = b; Li.M = c->M; Li.window = nullptr;
            CU(c, rpf_launch_pipe4096(Li));
            d.launches += 2;
        }
    }
    const int r = rows_touched(c, d, nspec);
    if (r > d.rows_dirty) d.rows_dirty = r;
    return RPF_OK;
}

// The copies of an earlier push that the caller was not made to wait for.
int drain_pending(rpf_ctx *c)
{
    for (const rpf_ctx::Pending &p : c->pending) {
        CU(c, cudaSetDevice(c->devs[p.dev].device));
        CU(c, cudaEventSynchronize(c->devs[p.dev].ev_copied[p.slot]));
    }
    c->pending.clear();
    return RPF_OK;
}

int create_device(rpf_ctx *c, rpf_dev &d, const rpf_config *cfg)
{
    const int N = c->N;
    CU(c, cudaSetDevice(d.device));
    CU(c, cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.device));
    CU(c, cudaStreamCreateWithFlags(&d.own_stream, cudaStreamNonBlocking));
    CU(c, cudaStreamCreateWithFlags(&d.copy_stream, cudaStreamNonBlocking));
    d.stream = d.own_stream;
    for (int i = 0; i < kNumStage; i++) {
        CU(c, cudaEventCreateWithFlags(&d.ev_copied[i], cudaEventDisableTiming));
        CU(c, cudaEventCreateWithFlags(&d.ev_consumed[i], cudaEventDisableTiming));
        CU(c, cudaHostAlloc((void **)&d.h_bounce[i], kSmallPush + c->spectrum_bytes, cudaHostAllocDefault));
    }
    const int sms = d.sms;
    int inner = N;
    switch (c->plan) {
        case PLAN_DIRECT:
            d.grid = 4 * sms;
            CU(c, rpf_direct_dft_configure());
            break;
        case PLAN_FUSED3:
            d.grid = (c->usepipe || c->use_pipen || c->use_batched) ? sms : 2 * sms;
            if (c->use_batched) inner = kInnerN;                 // the loop's twiddles come from the W_4096 table
            break;
        case PLAN_SMALL:
            d.grid = 8 * sms;
            break;
        case PLAN_BLUESTEIN: {
            // tables in double on the host: chirp c[n] = exp(-i pi (n^2 mod 2N) / N) and Bhat = FFT_M(b) / M with
            // b[n] = b[M - n] = conj c[n] for n < N, zero elsewhere
            const int M = c->blueM;
            inner = M;
            // private rows: up to 64 slices of the batch, but no more than 64 MiB of accumulators
            int64_t slices = ((int64_t)64 << 20) / ((int64_t)N * (int64_t)sizeof(double));
            d.grid = (int)(slices < 2 ? 2 : slices > kBlueSlices ? kBlueSlices : slices);
            CU(c, cudaMalloc(&d.d_blue[0], (size_t)c->blue_batch * M * sizeof(float2)));
            CU(c, cudaMalloc(&d.d_blue[1], (size_t)c->blue_batch * M * sizeof(float2)));
            if (M != N) {
                std::vector<float2> ch(N), bh(M);
                std::vector<double> br(M, 0.0), bi(M, 0.0);
                const double PI = 3.14159265358979323846264338327950288;
                for (int n = 0; n < N; n++) {
                    const int64_t ph = ((int64_t)n * n) % (2 * (int64_t)N);
                    const double a = PI * (double)ph / (double)N;
                    ch[n] = make_float2((float)cos(a), (float)(-sin(a)));
                    br[n] = cos(a); bi[n] = sin(a);
                    if (n) { br[M - n] = cos(a); bi[M - n] = sin(a); }
                }
                host_fft_pow2(br, bi);                           // float64, in place
                for (int k = 0; k < M; k++) bh[k] = make_float2((float)(br[k] / M), (float)(bi[k] / M));
                CU(c, cudaMalloc(&d.d_chirp, sizeof(float2) * N));
                CU(c, cudaMemcpy(d.d_chirp, ch.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
                CU(c, cudaMalloc(&d.d_bhat, sizeof(float2) * M));
                CU(c, cudaMemcpy(d.d_bhat, bh.data(), sizeof(float2) * M, cudaMemcpyHostToDevice));
            }
            break;
        }
        case PLAN_FOURSTEP: {
            // inner transform = pipe4096 in complex-source mode: every CTA is bound to one outer bin, so the
            // grid is a multiple of M (CTAs are persistent; co-residency is not required for correctness)
            inner = kInnerN;
            d.grid = (sms / c->M) * c->M;
            if (d.grid < c->M) d.grid = c->M;
            std::vector<float2> tb;
            fill_twiddle(tb, N);
            CU(c, cudaMalloc(&d.d_twiddle_big, sizeof(float2) * N));
            CU(c, cudaMemcpy(d.d_twiddle_big, tb.data(), sizeof(float2) * N, cudaMemcpyHostToDevice));
            // Experiment (RPF_FLAG_CLUSTER65536; 65536 bins, no window): one launch, the outer radix-16 scattered through
            // distributed shared memory by 16-CTA clusters (cluster65536.cu).  Correct, but 212 against 299 Gsamples/s for the
            // two-kernel path, so not the default.
            if (c->M == 16 && !cfg->window && (cfg->flags & RPF_FLAG_CLUSTER65536)) {
                int ncl = 0;
                CU(c, rpf_cluster65536_configure(&ncl));
                if (ncl * 16 > d.grid) ncl = d.grid / 16;          // (rows are sized by d.grid)
                d.cluster_grid = ncl * 16;
            }
            // 8192 ... 65536 bins: one launch, one CTA per spectrum (solo65536.cu): 375-383 against 299-309 Gsamples/s for
            // the two-kernel path on long launches.  A launch costs every CTA the time of a whole spectrum (~30 us) however few spectra
            // there are, so short ones (the reference's 1,638,400-byte reads hold 12.5 spectra) stay on the two-kernel
            // path; both add into the same private rows (row index mod 16 = outer bin on either path).
            const uint32_t forced = cfg->flags & (RPF_FLAG_SOLO65536 | RPF_FLAG_FOURSTEP | RPF_FLAG_CLUSTER65536 | RPF_FLAG_LOCKSTEP);
            if (forced ? (forced & RPF_FLAG_SOLO65536) != 0 : cfg->scratch_mib == 0) {
                CU(c, rpf_solo65536_configure(c->M, &d.solo_saved_l2));
                d.solo_grid = sms;
                d.solo_min = (cfg->flags & RPF_FLAG_SOLO65536) ? 1 : 96;
                // with the flag, rpf_config.scratch_mib bounds the scratch (32 M KiB per CTA) and so the CTA count (tuning runs)
                if (forced && cfg->scratch_mib > 0 && cfg->scratch_mib * (32 / c->M) < d.solo_grid) d.solo_grid = cfg->scratch_mib * (32 / c->M);
                if (const size_t sb = rpf_solo65536_scratch_bytes(c->M, d.solo_grid)) {      // (0: hand-off through tensor memory)
                    CU(c, cudaMalloc(&d.d_solo, sb));
                    CU(c, cudaMemset(d.d_solo, 0, sb));
                }
                d.cluster_grid = 0;
            }
            break;
        }
    }
    if (c->plan == PLAN_FUSED3 || c->plan == PLAN_FOURSTEP) {
        CU(c, rpf_fused3_configure());
        CU(c, rpf_pipe4096_configure());
        if (c->plan == PLAN_FOURSTEP) CU(c, rpf_outer_tma_configure());
        if (c->use_pipen) CU(c, rpf_pipen_configure());
        if (c->use_tc) CU(c, rpf_pipe4096_tc_configure());
    }
    std::vector<float2> tw;
    fill_twiddle(tw, inner);
    CU(c, cudaMalloc(&d.d_twiddle, sizeof(float2) * inner));
    CU(c, cudaMemcpy(d.d_twiddle, tw.data(), sizeof(float2) * inner, cudaMemcpyHostToDevice));
    if (cfg->window) {
        CU(c, cudaMalloc(&d.d_window, sizeof(float) * N));
        CU(c, cudaMemcpy(d.d_window, cfg->window, sizeof(float) * N, cudaMemcpyHostToDevice));
    }
    d.nrows = d.grid * c->groups;
    if (d.solo_grid * c->M > d.nrows) d.nrows = d.solo_grid * c->M;   // M private rows per CTA of the solo kernel
    CU(c, cudaMalloc(&d.d_rows, sizeof(double) * (size_t)d.nrows * c->rowlen));
    CU(c, cudaMalloc(&d.d_pwr, sizeof(double) * N));
    CU(c, cudaMalloc(&d.d_tmp, sizeof(double) * 2 * N));
    CU(c, cudaMalloc(&d.d_baseline, sizeof(double) * N));
    CU(c, cudaMemset(d.d_rows, 0, sizeof(double) * (size_t)d.nrows * c->rowlen));
    d.rows_dirty = 0;
    const size_t hint = cfg->buf_length > 0 ? (size_t)cfg->buf_length : (size_t)16384 * 100;   // src/params.h:43
    const int r = ensure_stage(c, d, hint + c->spectrum_bytes);
    if (r) return r;
    CU(c, cudaDeviceSynchronize());
    return RPF_OK;
}

// pwr[] of one device (sum of its private rows, fixed order) into `out` (device memory), on its stream.
int reduce_device(rpf_ctx *c, rpf_dev &d, double *out)
{
    if (d.rows_dirty == 0) {
        CU(c, cudaMemsetAsync(out, 0, sizeof(double) * c->N, d.stream));
        return RPF_OK;
    }
    if (c->plan == PLAN_FOURSTEP)
        CU(c, rpf_launch_reduce_rows_fourstep(d.d_rows, d.rows_dirty, c->M, kInnerN, out, d.stream));
    else
        CU(c, rpf_launch_reduce_rows(d.d_rows, d.rows_dirty, c->N, out, d.stream));
    d.launches++;
    return RPF_OK;
}

// Sum of Datastore::pwr over every GPU of the context and every rank of the communicator
// (src/datastore.cxx:83-85 is a sum, so the order of spectra does not matter): reduce on each device,
// then ONE ncclAllReduce of N doubles per communicator, enqueued on the compute streams.
// Once devs[0].stream has drained, devs[0].d_pwr (or `out0` if given) holds the total.
int reduce_all(rpf_ctx *c, double *out0)
{
    NcclApi *api = nullptr;
    if (c->devs.size() > 1 || c->comm_nranks > 1) {
        api = nccl_api();
        if (!api) return fail(c, RPF_ERR_STATE, "multi-GPU reduction needs libnccl.so.2 (not found)");
    }
    for (size_t i = 0; i < c->devs.size(); i++) {
        rpf_dev &d = c->devs[i];
        CU(c, cudaSetDevice(d.device));
        int r = reduce_device(c, d, (i == 0 && out0) ? out0 : d.d_pwr);
        if (r) return r;
    }
    if (c->devs.size() > 1) {
        NC(c, api, api->GroupStart());
        for (size_t i = 0; i < c->devs.size(); i++) {
            rpf_dev &d = c->devs[i];
            double *buf = (i == 0 && out0) ? out0 : d.d_pwr;
            NC(c, api, api->AllReduce(buf, buf, (size_t)c->N, ncclDouble, ncclSum, d.comm, d.stream));
        }
        NC(c, api, api->GroupEnd());
    } else if (c->comm_nranks > 1) {
        rpf_dev &d = c->devs[0];
        double *buf = out0 ? out0 : d.d_pwr;
        NC(c, api, api->AllReduce(buf, buf, (size_t)c->N, ncclDouble, ncclSum, d.comm, d.stream));
    }
    return RPF_OK;
}

int sync_compute_streams(rpf_ctx *c)
{
    for (rpf_dev &d : c->devs) {
        CU(c, cudaSetDevice(d.device));
        CU(c, cudaStreamSynchronize(d.stream));
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
    if (N > kMaxBins)
        return fail(nullptr, RPF_ERR_UNSUPPORTED_N, "rpf_create: N=%d not supported by the CUDA path (even N up to %d)", N, kMaxBins);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, RPF_ERR_CUDA, "rpf_create: no usable CUDA device (%s); there is no CPU fallback",
                    cudaGetErrorString(e));
    }
    const int nd = cfg->n_devices > 0 ? cfg->n_devices : 1;
    if (nd > RPF_MAX_DEVICES)
        return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: n_devices %d > %d", nd, RPF_MAX_DEVICES);
    for (int i = 0; i < nd; i++) {
        const int dv = cfg->n_devices > 0 ? cfg->devices[i] : cfg->device;
        if (dv < 0 || dv >= ndev)
            return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: device %d out of range [0,%d)", dv, ndev);
        for (int k = 0; k < i; k++)
            if (cfg->devices[k] == dv) return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_create: device %d listed twice", dv);
    }

    rpf_ctx *c = new (std::nothrow) rpf_ctx;
    if (!c) return fail(nullptr, RPF_ERR_NOMEM, "rpf_create: out of host memory");
    c->N = N;
    c->spectrum_bytes = (size_t)2 * N;
    c->has_window = cfg->window != nullptr;
    c->pipelined = (cfg->flags & RPF_FLAG_PIPELINED_PUSH) != 0;
    auto body = [&]() -> int {
        if (!is_pow2(N) && N <= kDirectMaxN) {
            c->plan = PLAN_DIRECT; c->groups = 1;
        } else if (!is_pow2(N) || N > 16 * kInnerN) {
            c->plan = PLAN_BLUESTEIN; c->groups = 1;
            c->blueM = rpf_bluestein_M(N);
            c->blue_batch = ((size_t)128 << 20) / ((size_t)c->blueM * sizeof(float2));   // 2 x 128 MiB ping-pong
            if (c->blue_batch < 1) c->blue_batch = 1;
        } else if (rpf_fused3_supported(N)) {
            c->plan = PLAN_FUSED3;
            c->usepipe = (N == kInnerN);
            c->use_tc = c->usepipe && !cfg->window && (cfg->flags & RPF_FLAG_TC_PASS1);
            // Three kernels cover 512 / 1024 / 2048 bins; measured on the B200 (Gsamples/s; with a Hann window in brackets):
            //                     512          1024         2048
            //   batched on the    693 (695)    716 (671)    661 (641)    4096 loop, 8 / 4 / 2 spectra per ring slot
            //   pipen             774 (768)    621 (601)    574 (556)    two-stream pipeline, exchange barriers per spectrum
            //   fused3            693 (606)    609 (549)    609 (557)    lock-step, two CTAs per SM (round 1)
            // At 512 bins every warp of pipen owns whole spectra (no cross-warp exchange) and the warps drift apart
            // freely; above that the tuned 4096 loop wins.  The flags force a kernel for A/B runs.
            const uint32_t forced = cfg->flags & (RPF_FLAG_LOCKSTEP | RPF_FLAG_PIPELINE | RPF_FLAG_BATCHED4096);
            const bool small = rpf_pipen_supported(N);
            c->use_batched = small && (forced ? forced == RPF_FLAG_BATCHED4096 : N != 512);
            c->use_pipen = small && (forced ? forced == RPF_FLAG_PIPELINE : N == 512);
            c->groups = c->usepipe ? 1 : c->use_batched ? kInnerN / N : rpf_fused3_groups(N);
        } else if (N >= 16 && N <= 256 && !(cfg->flags & RPF_FLAG_LOCKSTEP)) {
            // 256 ... 16 bins: the 4096 loop with 16 ... 256 spectra per ring slot, pass 1 (and part of pass 2) empty:
            // 722 Gsamples/s at 256 bins against 60 for small_kernel (RPF_FLAG_LOCKSTEP selects it for A/B runs)
            c->plan = PLAN_FUSED3; c->use_batched = true; c->groups = kInnerN / N;
        } else if (N <= 256) {
            c->plan = PLAN_SMALL; c->groups = 1;
        } else {
            c->plan = PLAN_FOURSTEP; c->M = N / kInnerN;
            c->old_outer = (cfg->flags & RPF_FLAG_LOCKSTEP) != 0;
            c->usepipe = true; c->groups = 1;
            // batch of complex64 rows (round 1: 64 -> 171, 256 -> 205 Gs/s); rpf_config.scratch_mib overrides (tuning runs)
            const size_t scratch_mb = cfg->scratch_mib > 0 ? (size_t)cfg->scratch_mib : 1024;
            c->scratch_spectra = (scratch_mb << 20) / ((size_t)N * sizeof(float2));
            if (c->scratch_spectra < 1) c->scratch_spectra = 1;
        }
        c->rowlen = c->plan == PLAN_FOURSTEP ? kInnerN : N;
        c->devs.resize(nd);
        for (int i = 0; i < nd; i++) {
            c->devs[i].device = cfg->n_devices > 0 ? cfg->devices[i] : cfg->device;
            int r = create_device(c, c->devs[i], cfg);
            if (r) return r;
        }
        for (int i = 0; i < kNumStage; i++)
            CU(c, cudaHostAlloc((void **)&c->h_carry[i], c->spectrum_bytes, cudaHostAllocDefault));
        CU(c, cudaHostAlloc((void **)&c->h_pwr, sizeof(double) * N, cudaHostAllocMapped));
        if (nd > 1) {
            // single process, ncclCommInitAll over the context's devices
            NcclApi *api = nccl_api();
            if (!api) return fail(c, RPF_ERR_STATE, "rpf_create: n_devices > 1 needs libnccl.so.2 (not found)");
            std::vector<ncclComm_t> comms(nd);
            std::vector<int> ids(nd);
            for (int i = 0; i < nd; i++) ids[i] = c->devs[i].device;
            {
                StdoutToStderr guard;
                NC(c, api, api->CommInitAll(comms.data(), nd, ids.data()));
            }
            for (int i = 0; i < nd; i++) c->devs[i].comm = comms[i];
        }
        return RPF_OK;
    };
    const int rc = body();
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
    NcclApi *api = nccl_api();
    for (rpf_dev &d : c->devs) {
        cudaSetDevice(d.device);
        if (d.stream) cudaStreamSynchronize(d.stream);
        if (d.copy_stream) cudaStreamSynchronize(d.copy_stream);
        if (d.comm && api) api->CommDestroy(d.comm);
        for (int i = 0; i < kNumStage; i++) {
            if (d.d_stage[i]) cudaFree(d.d_stage[i]);
            if (d.h_bounce[i]) cudaFreeHost(d.h_bounce[i]);
            if (d.ev_copied[i]) cudaEventDestroy(d.ev_copied[i]);
            if (d.ev_consumed[i]) cudaEventDestroy(d.ev_consumed[i]);
        }
        cudaFree(d.d_twiddle); cudaFree(d.d_twiddle_big); cudaFree(d.d_window);
        cudaFree(d.d_rows); cudaFree(d.d_pwr); cudaFree(d.d_tmp); cudaFree(d.d_baseline); cudaFree(d.d_scratch); cudaFree(d.d_solo);
        if (d.solo_grid) rpf_solo65536_release(d.solo_saved_l2);
        cudaFree(d.d_chirp); cudaFree(d.d_bhat); cudaFree(d.d_blue[0]); cudaFree(d.d_blue[1]);
        if (d.own_stream) cudaStreamDestroy(d.own_stream);
        if (d.copy_stream) cudaStreamDestroy(d.copy_stream);
    }
    for (int i = 0; i < kNumStage; i++)
        if (c->h_carry[i]) cudaFreeHost(c->h_carry[i]);
    if (c->h_pwr) cudaFreeHost(c->h_pwr);
    cudaGetLastError();
    delete c;
}

int rpf_reset(rpf_ctx *c, int64_t repeats)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (repeats < 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_reset: repeats must be >= 0");
    int r = drain_pending(c);
    if (r) return r;
    for (rpf_dev &d : c->devs) {
        CU(c, cudaSetDevice(d.device));
        if (d.rows_dirty)   // only the rows a launch wrote since the last reset
            CU(c, cudaMemsetAsync(d.d_rows, 0, sizeof(double) * (size_t)d.rows_dirty * c->rowlen, d.stream));
        d.rows_dirty = 0;
    }
    c->repeats = repeats;
    c->repeats_done = 0;
    c->carry_bytes = 0;
    c->early_reduced = false;
    return RPF_OK;
}

int rpf_push(rpf_ctx *c, const uint8_t *bytes, size_t len)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (len == 0) return RPF_OK;
    if (!bytes) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push: NULL buffer");
    if (len & 1) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push: odd byte count %zu (I/Q pairs, src/datastore.cxx:74)", len);
    int r = drain_pending(c);                                    // pipelined mode: the previous push's copies
    if (r) return r;
    if (c->repeats_done >= c->repeats) return RPF_OK;            // src/datastore.cxx:67: discarded

    const size_t sb = c->spectrum_bytes;
    const size_t total = c->carry_bytes + len;
    const int cs = c->cslot, cn = (c->cslot + 1) % kNumStage;
    if (total < sb) {                                            // still inside one spectrum
        memcpy(c->h_carry[cn], c->h_carry[cs], c->carry_bytes);
        memcpy(c->h_carry[cn] + c->carry_bytes, bytes, len);
        c->carry_bytes = total;
        c->cslot = cn;
        return RPF_OK;
    }
    const int64_t avail = (int64_t)(total / sb);
    const int64_t left = c->repeats - c->repeats_done;
    const int64_t nspec = avail < left ? avail : left;
    const size_t used = (size_t)nspec * sb;                      // carry + new bytes the kernels will read

    // Deal the whole spectra of this push over the GPUs: contiguous parts of at least kMinPartBytes, starting at
    // the round-robin cursor (small pushes therefore go to one GPU each, in turn).  The carry was resolved above,
    // once, before dealing: every part starts on a spectrum boundary.
    const int nd = (int)c->devs.size();
    int parts = (int)(used / kMinPartBytes);
    if (parts < 1) parts = 1;
    if (parts > nd) parts = nd;
    if ((int64_t)parts > nspec) parts = (int)nspec;
    size_t consumed = 0;                                         // bytes of (carry + new) dealt so far
    cudaEvent_t in_place_read = nullptr;                         // set when a kernel reads the caller's buffer itself
    for (int p = 0; p < parts; p++) {
        const int64_t s0 = nspec * p / parts, s1 = nspec * (p + 1) / parts;
        const size_t part_bytes = (size_t)(s1 - s0) * sb;
        const int di = (c->cursor + p) % nd;
        rpf_dev &d = c->devs[di];
        CU(c, cudaSetDevice(d.device));
        r = ensure_stage(c, d, part_bytes);
        if (r) return r;
        const int s = d.slot;
        d.slot = (d.slot + 1) % kNumStage;
        const size_t from_carry = consumed < c->carry_bytes ? c->carry_bytes - consumed : 0;   // only part 0
        const uint8_t *src = bytes + (consumed + from_carry - c->carry_bytes);
        const size_t from_new = part_bytes - from_carry;
        if (part_bytes <= kSmallPush + sb && !c->pipelined) {
            // short read (config 1: one 114,688-byte buffer): bounce through pinned memory on the compute stream;
            // the caller's buffer is free as soon as the host memcpy returns, nobody waits for an event
            // (the kernels read the pinned bounce buffer in place -- TMA over PCIe -- so there is no separate copy
            // operation either: for ~100 KB the launch gaps of a memcpy cost more than the transfer)
            CU(c, cudaEventSynchronize(d.ev_consumed[s]));       // bounce slot free (3 pushes ago)
            // A short push that COMPLETES the integration out of a page-locked buffer (the reference-shaped Datastore
            // page-locks its Buffers; config 1 is one such read) skips even the bounce copy: the kernel reads the
            // caller's buffer in place and the call returns when it has -- the caller would wait for this kernel in
            // rpf_finish anyway, and ~10 us of host memcpy leave the critical path.  Other short pushes keep the
            // bounce copy, which lets the caller go on while the GPU works.
            bool in_place = false;
            if (from_carry == 0 && parts == 1 && c->repeats_done + nspec >= c->repeats && (((uintptr_t)src) & 15) == 0) {
                cudaPointerAttributes at;
                if (cudaPointerGetAttributes(&at, src) == cudaSuccess)
                    in_place = at.type == cudaMemoryTypeHost && at.devicePointer == (const void *)src;
                else
                    cudaGetLastError();
            }
            if (in_place) {
                r = launch_spectra(c, d, src, s1 - s0);
                if (r) return r;
                CU(c, cudaEventRecord(d.ev_consumed[s], d.stream));
                in_place_read = d.ev_consumed[s];                // waited for below: `bytes` is free on return
                consumed += part_bytes;
                continue;
            }
            if (from_carry) memcpy(d.h_bounce[s], c->h_carry[cs], from_carry);
            memcpy(d.h_bounce[s] + from_carry, src, from_new);
            r = launch_spectra(c, d, d.h_bounce[s], s1 - s0);
            if (r) return r;
            CU(c, cudaEventRecord(d.ev_consumed[s], d.stream));
        } else {
            // copy stream: wait until the kernels that last read this staging buffer are done
            CU(c, cudaStreamWaitEvent(d.copy_stream, d.ev_consumed[s], 0));
            if (from_carry)
                CU(c, cudaMemcpyAsync(d.d_stage[s], c->h_carry[cs], from_carry, cudaMemcpyHostToDevice, d.copy_stream));
            CU(c, cudaMemcpyAsync(d.d_stage[s] + from_carry, src, from_new, cudaMemcpyHostToDevice, d.copy_stream));
            CU(c, cudaEventRecord(d.ev_copied[s], d.copy_stream));
            CU(c, cudaStreamWaitEvent(d.stream, d.ev_copied[s], 0));
            r = launch_spectra(c, d, d.d_stage[s], s1 - s0);
            if (r) return r;
            CU(c, cudaEventRecord(d.ev_consumed[s], d.stream));
            c->pending.push_back({di, s});
        }
        consumed += part_bytes;
    }
    c->cursor = (c->cursor + parts) % nd;
    c->repeats_done += nspec;
    if (c->repeats_done >= c->repeats) {
        c->carry_bytes = 0;                                      // rest of the stream is discarded
    } else {
        const size_t tail = total - used;                        // < sb; lies entirely in `bytes`
        memcpy(c->h_carry[cn], bytes + (len - tail), tail);
        c->carry_bytes = tail;
    }
    c->cslot = cn;
    // The push that completes the integration (src/datastore.cxx:67: the worker stops here) enqueues the reduction
    // of the private rows right behind its kernels instead of leaving it to rpf_finish: one launch latency less
    // between the last FFT kernel and the result (config 1: a single read per integration).
    if (c->repeats_done >= c->repeats && c->devs.size() == 1 && c->comm_nranks == 1 && !c->early_reduced) {
        CU(c, cudaSetDevice(c->devs[0].device));
        r = reduce_device(c, c->devs[0], c->h_pwr);
        if (r) return r;
        c->early_reduced = true;
    }
    if (in_place_read) CU(c, cudaEventSynchronize(in_place_read));
    if (!c->pipelined) return drain_pending(c);                  // caller may now reuse `bytes`
    return RPF_OK;                                               // pipelined: `bytes` is free after the NEXT call
}

int rpf_push_device(rpf_ctx *c, const uint8_t *bytes, size_t len)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (len == 0) return RPF_OK;
    if (!bytes) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push_device: NULL buffer");
    if (len & 1) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push_device: odd byte count %zu", len);
    int r = drain_pending(c);
    if (r) return r;
    if (c->repeats_done >= c->repeats) return RPF_OK;
    const size_t sb = c->spectrum_bytes;
    const int64_t left = c->repeats - c->repeats_done;

    // which of the context's GPUs owns the buffer
    int di = 0;
    if (c->devs.size() > 1) {
        cudaPointerAttributes at;
        CU(c, cudaPointerGetAttributes(&at, bytes));
        di = -1;
        for (size_t i = 0; i < c->devs.size(); i++)
            if (c->devs[i].device == at.device && at.type == cudaMemoryTypeDevice) di = (int)i;
        if (di < 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_push_device: buffer is not in the memory of one of the context's GPUs");
        if (c->carry_bytes || len % sb)
            return fail(c, RPF_ERR_STATE, "rpf_push_device: with several GPUs device buffers must hold whole spectra "
                                          "(no partial spectrum may be pending or left over)");
    }
    rpf_dev &d = c->devs[di];
    CU(c, cudaSetDevice(d.device));

    if (c->carry_bytes == 0 && (((uintptr_t)bytes) & 15) == 0) {
        // zero-copy: spectra are read in place by TMA
        const int64_t avail = (int64_t)(len / sb);
        const int64_t nspec = avail < left ? avail : left;
        r = launch_spectra(c, d, bytes, nspec);
        if (r) return r;
        c->repeats_done += nspec;
        if (c->repeats_done >= c->repeats) return RPF_OK;
        const size_t tail = len - (size_t)nspec * sb;
        if (tail) {
            CU(c, cudaMemcpyAsync(c->h_carry[c->cslot], bytes + (len - tail), tail, cudaMemcpyDeviceToHost, d.stream));
            CU(c, cudaStreamSynchronize(d.stream));
            c->carry_bytes = tail;
        }
        return RPF_OK;
    }
    // general case: assemble carry + buffer contiguously in staging (device-to-device)
    const size_t total = c->carry_bytes + len;
    const int cs = c->cslot, cn = (c->cslot + 1) % kNumStage;
    if (total < sb) {
        memcpy(c->h_carry[cn], c->h_carry[cs], c->carry_bytes);
        CU(c, cudaMemcpyAsync(c->h_carry[cn] + c->carry_bytes, bytes, len, cudaMemcpyDeviceToHost, d.stream));
        CU(c, cudaStreamSynchronize(d.stream));
        c->carry_bytes = total; c->cslot = cn;
        return RPF_OK;
    }
    r = ensure_stage(c, d, total);
    if (r) return r;
    const int s = d.slot;
    d.slot = (d.slot + 1) % kNumStage;
    const int64_t avail = (int64_t)(total / sb);
    const int64_t nspec = avail < left ? avail : left;
    const size_t used = (size_t)nspec * sb;
    CU(c, cudaStreamWaitEvent(d.stream, d.ev_consumed[s], 0));
    if (c->carry_bytes)
        CU(c, cudaMemcpyAsync(d.d_stage[s], c->h_carry[cs], c->carry_bytes, cudaMemcpyHostToDevice, d.stream));
    CU(c, cudaMemcpyAsync(d.d_stage[s] + c->carry_bytes, bytes, used - c->carry_bytes, cudaMemcpyDeviceToDevice, d.stream));
    r = launch_spectra(c, d, d.d_stage[s], nspec);
    if (r) return r;
    CU(c, cudaEventRecord(d.ev_consumed[s], d.stream));
    c->repeats_done += nspec;
    const size_t tail = c->repeats_done >= c->repeats ? 0 : total - used;
    if (tail) CU(c, cudaMemcpyAsync(c->h_carry[cn], bytes + (len - tail), tail, cudaMemcpyDeviceToHost, d.stream));
    // in every branch: the pinned h_carry[cs] -> staging copy has completed before that slot can be rewritten
    CU(c, cudaStreamSynchronize(d.stream));
    c->carry_bytes = tail;
    c->cslot = cn;
    return RPF_OK;
}

int rpf_sync(rpf_ctx *c)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    int r = drain_pending(c);
    if (r) return r;
    for (rpf_dev &d : c->devs) {
        CU(c, cudaSetDevice(d.device));
        CU(c, cudaStreamSynchronize(d.copy_stream));
        CU(c, cudaStreamSynchronize(d.stream));
    }
    return RPF_OK;
}

int rpf_finish_device(rpf_ctx *c, double *pwr_out_device, int64_t *repeats_done)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    int r = drain_pending(c);
    if (r) return r;
    if (pwr_out_device) {
        r = reduce_all(c, pwr_out_device);
        if (r) return r;
    }
    if (repeats_done) *repeats_done = c->repeats_done;
    return RPF_OK;
}

int rpf_finish(rpf_ctx *c, double *pwr_out, int64_t *repeats_done)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    int r = drain_pending(c);
    if (r) return r;
    rpf_dev &d0 = c->devs[0];
    const bool direct = pwr_out && c->devs.size() == 1 && c->comm_nranks == 1;
    if (direct) {
        // one GPU: the reduction writes pwr[] straight into device-visible pinned host memory (no D2H operation)
        CU(c, cudaSetDevice(d0.device));
        if (!c->early_reduced) {                                 // (else: enqueued by the push that completed the integration)
            r = reduce_device(c, d0, c->h_pwr);
            if (r) return r;
        }
    } else if (pwr_out) {
        r = reduce_all(c, nullptr);
        if (r) return r;
        CU(c, cudaSetDevice(d0.device));
        CU(c, cudaMemcpyAsync(pwr_out, d0.d_pwr, sizeof(double) * c->N, cudaMemcpyDeviceToHost, d0.stream));
    }
    r = sync_compute_streams(c);
    if (r) return r;
    if (direct) memcpy(pwr_out, c->h_pwr, sizeof(double) * c->N);
    if (repeats_done) *repeats_done = c->repeats_done;
    return RPF_OK;
}

int64_t rpf_repeats_done(const rpf_ctx *c) { return c ? c->repeats_done : 0; }
int64_t rpf_fft_pointer(const rpf_ctx *c) { return c ? (int64_t)(c->carry_bytes / 2) : 0; }
int64_t rpf_kernel_launches(const rpf_ctx *c)
{
    int64_t n = 0;
    if (c) for (const rpf_dev &d : c->devs) n += d.launches;
    return n;
}
int32_t rpf_device_count(const rpf_ctx *c) { return c ? (int32_t)c->devs.size() : 0; }

const char *rpf_plan_name(const rpf_ctx *c)
{
    if (!c) return "";
    if (c->use_tc) return "pipe4096_tc";
    if (c->plan == PLAN_BLUESTEIN) return c->blueM == c->N ? "gfft" : "bluestein";
    if (c->plan == PLAN_FUSED3) return c->usepipe ? "pipe4096" : c->use_batched ? "pipe4096_batched" : c->use_pipen ? "pipen" : "fused3";
    if (c->plan == PLAN_FOURSTEP)
        return !c->devs.empty() && c->devs[0].solo_grid ? (c->M == 16 ? "solo65536" : "solo") : !c->devs.empty() && c->devs[0].cluster_grid ? "cluster65536" : c->old_outer ? "fourstep_plain_outer" : "fourstep";
    return c->plan == PLAN_SMALL ? "small" : "direct";
}

int rpf_set_stream(rpf_ctx *c, void *cuda_stream)
{
    if (!c) return RPF_ERR_INVALID_ARG;
    if (c->devs.size() != 1) return fail(c, RPF_ERR_STATE, "rpf_set_stream: only for single-GPU contexts");
    int r = drain_pending(c);
    if (r) return r;
    rpf_dev &d = c->devs[0];
    CU(c, cudaSetDevice(d.device));
    CU(c, cudaStreamSynchronize(d.stream));
    CU(c, cudaStreamSynchronize(d.copy_stream));
    d.stream = cuda_stream ? (cudaStream_t)cuda_stream : d.own_stream;
    return RPF_OK;
}

int rpf_comm_unique_id(void *id_out, size_t id_bytes)
{
    NcclApi *api = nccl_api();
    if (!api) return fail(nullptr, RPF_ERR_STATE, "rpf_comm_unique_id: libnccl.so.2 not found");
    if (!id_out || id_bytes < sizeof(ncclUniqueId))
        return fail(nullptr, RPF_ERR_INVALID_ARG, "rpf_comm_unique_id: need %zu bytes", sizeof(ncclUniqueId));
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return fail(nullptr, RPF_ERR_CUDA, "ncclGetUniqueId failed");
    memcpy(id_out, &id, sizeof(id));
    return RPF_OK;
}

int rpf_comm_init_rank(rpf_ctx *c, const void *id, size_t id_bytes, int32_t rank, int32_t nranks)
{
    if (!c || !id) return RPF_ERR_INVALID_ARG;
    if (c->devs.size() != 1)
        return fail(c, RPF_ERR_STATE, "rpf_comm_init_rank: the context already spans several GPUs (one communicator per context)");
    if (id_bytes < sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks)
        return fail(c, RPF_ERR_INVALID_ARG, "rpf_comm_init_rank: bad arguments");
    NcclApi *api = nccl_api();
    if (!api) return fail(c, RPF_ERR_STATE, "rpf_comm_init_rank: libnccl.so.2 not found");
    rpf_dev &d = c->devs[0];
    CU(c, cudaSetDevice(d.device));
    if (d.comm) { api->CommDestroy(d.comm); d.comm = nullptr; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    {
        StdoutToStderr guard;
        NC(c, api, api->CommInitRank(&d.comm, nranks, uid, rank));
    }
    c->comm_rank = rank; c->comm_nranks = nranks;
    return RPF_OK;
}

int rpf_write_data(rpf_ctx *c, double tuned_freq, int32_t samplerate, int32_t linear, const double *baseline,
                   double *freq_out, double *val_out)
{
    if (!c || !val_out) return RPF_ERR_INVALID_ARG;
    if (samplerate <= 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_write_data: samplerate must be > 0");
    int r = drain_pending(c);
    if (r) return r;
    r = reduce_all(c, nullptr);
    if (r) return r;
    rpf_dev &d = c->devs[0];
    CU(c, cudaSetDevice(d.device));
    const int N = c->N;
    if (baseline)
        CU(c, cudaMemcpyAsync(d.d_baseline, baseline, sizeof(double) * N, cudaMemcpyHostToDevice, d.stream));
    CU(c, rpf_launch_write_data(d.d_pwr, N, c->repeats_done, tuned_freq, samplerate, linear,
                                baseline ? d.d_baseline : nullptr, d.d_tmp, d.d_tmp + N, nullptr, d.stream));
    d.launches++;
    if (freq_out) CU(c, cudaMemcpyAsync(freq_out, d.d_tmp, sizeof(double) * N, cudaMemcpyDeviceToHost, d.stream));
    CU(c, cudaMemcpyAsync(val_out, d.d_tmp + N, sizeof(double) * N, cudaMemcpyDeviceToHost, d.stream));
    return sync_compute_streams(c);
}

int rpf_write_row_f32(rpf_ctx *c, int32_t samplerate, int32_t linear, const double *baseline, float *row_out)
{
    if (!c || !row_out) return RPF_ERR_INVALID_ARG;
    if (samplerate <= 0) return fail(c, RPF_ERR_INVALID_ARG, "rpf_write_row_f32: samplerate must be > 0");
    int r = drain_pending(c);
    if (r) return r;
    r = reduce_all(c, nullptr);
    if (r) return r;
    rpf_dev &d = c->devs[0];
    CU(c, cudaSetDevice(d.device));
    const int N = c->N;
    if (baseline)
        CU(c, cudaMemcpyAsync(d.d_baseline, baseline, sizeof(double) * N, cudaMemcpyHostToDevice, d.stream));
    float *d_row = reinterpret_cast<float *>(d.d_tmp);          // N floats at the start of the 2N-double scratch
    CU(c, rpf_launch_write_data(d.d_pwr, N, c->repeats_done, 0.0, samplerate, linear,
                                baseline ? d.d_baseline : nullptr, nullptr, d.d_tmp + N, d_row, d.stream));
    d.launches++;
    CU(c, cudaMemcpyAsync(row_out, d_row, sizeof(float) * N, cudaMemcpyDeviceToHost, d.stream));
    return sync_compute_streams(c);
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
    int sms = 0;
    if (cudaSetDevice(device) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        return RPF_ERR_CUDA;
    }
    return rpf_launch_synth(bytes, len, off, seed, mode, sms, (cudaStream_t)cuda_stream) == cudaSuccess ? RPF_OK : RPF_ERR_CUDA;
}

// A failing runtime call leaves a non-sticky "last error" behind on the calling thread; the launch wrappers
// return cudaGetLastError() after <<<>>>, so it is cleared here or it would surface at an unrelated launch.
int rpf_host_alloc(void **out, size_t bytes)
{
    if (!out) return RPF_ERR_INVALID_ARG;
    if (cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess) return RPF_OK;
    cudaGetLastError();
    return RPF_ERR_CUDA;
}
void rpf_host_free(void *p) { if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError(); }
int rpf_host_register(void *p, size_t bytes)
{
    if (cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess) return RPF_OK;
    cudaGetLastError();
    return RPF_ERR_CUDA;
}
int rpf_host_unregister(void *p)
{
    if (cudaHostUnregister(p) == cudaSuccess) return RPF_OK;
    cudaGetLastError();
    return RPF_ERR_CUDA;
}

}  // extern "C"
