/*
 * rpf_b200.h -- C ABI of the B200-native FFT worker for rtl_power_fftw.
 *
 * The reference has no plugin/FFI layer: its hot path is the body of
 * Datastore::fftThread() (src/datastore.cxx:48-96) calling libfftw3f.  This
 * header is the boundary a maintainer binds instead of <fftw3.h>: every entry
 * point names the reference code it replaces (paths relative to the reference
 * repository root).  Plain pointers and sizes only; no exceptions cross it;
 * every function returns RPF_OK (0) or a negative rpf_status and leaves a
 * message retrievable with rpf_strerror().
 *
 * Several GPUs (what must be sharded is the ONE Datastore / ONE pwr[] of
 * src/rtl_power_fftw.cxx:112 and the sum of src/datastore.cxx:83-85, read after join at
 * src/acquisition.cxx:347-357).  Two ways, both ending in exactly one NCCL allreduce(sum) of pwr[N]
 * (float64) on the compute streams inside rpf_finish / rpf_finish_device / rpf_write_data:
 *   - one process: rpf_config.devices[] -- rpf_push deals whole spectra over the GPUs (the partial
 *     spectrum carried between buffers is resolved once, before dealing), ncclCommInitAll;
 *   - one process per GPU: rpf_comm_unique_id on rank 0, ship the id by any means, then
 *     rpf_comm_init_rank on every rank's single-GPU context.
 * repeats_done is the local count of spectra (summed over the context's own GPUs).  Precision: bins are
 * summed in fp32 over at most 128 spectra per CTA, then in float64 (private rows, fixed-order reduction);
 * the reference adds every spectrum in double (src/datastore.cxx:84) -- the difference is < 1e-6 relative.
 * libnccl.so.2 is looked up at run time, and only when a second GPU or rank is involved.
 *
 * Threading: a context is thread-compatible, not thread-safe -- one caller at
 * a time, possibly a different thread per acquisition (the reference starts a
 * fresh worker thread per hop, src/acquisition.cxx:256).  Every entry point
 * selects the context's CUDA device itself.
 *
 * Index convention of pwr[]: identical to the reference's Datastore::pwr --
 * pwr[i] is the summed |X|^2 of the bin at frequency (i - N/2) * fs / N
 * (DC at i = N/2), consumed unchanged by Acquisition::write_data()
 * (src/acquisition.cxx:377,391).
 */
#ifndef RPF_B200_H
#define RPF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RPF_B200_ABI_VERSION 2
#define RPF_MAX_DEVICES 16

typedef enum rpf_status {
    RPF_OK = 0,
    RPF_ERR_INVALID_ARG = -1,   /* NULL pointer, odd byte count, N odd or < 2 ...            */
    RPF_ERR_UNSUPPORTED_N = -2, /* N is valid for the reference but not implemented on GPU: every even N
                                   from 2 to 4194304 is supported (fused kernels for powers of two up to
                                   65536, Bluestein / plain O(N log N) FFT passes for the rest)          */
    RPF_ERR_CUDA = -3,          /* a CUDA runtime call or kernel failed (message has detail) */
    RPF_ERR_NOMEM = -4,
    RPF_ERR_STATE = -5          /* call sequence error                                       */
} rpf_status;

typedef struct rpf_ctx rpf_ctx;

/* Mirrors the Params fields the hot path reads (src/params.h:33,42-43,52) plus
 * placement.  Zero-initialise, then set what you need. */
typedef struct rpf_config {
    int32_t      N;             /* Params::N: number of bins; even (src/params.cxx:152-155)  */
    int32_t      device;        /* CUDA device ordinal                                       */
    const float *window;        /* NULL, or N window values (copied): AuxData::window_values,
                                   applied as in src/datastore.cxx:76-77                     */
    int64_t      buf_length;    /* Params::buf_length: typical bytes per push (sizing hint)  */
    uint32_t     flags;         /* RPF_FLAG_*                                                */
    int32_t      n_devices;     /* 0: one GPU, `device`.  > 0: the context spans devices[0..n_devices):
                                   one Datastore, one pwr[], its spectra dealt over these GPUs and
                                   summed by one NCCL allreduce in rpf_finish (see below)       */
    int32_t      devices[RPF_MAX_DEVICES];
    int32_t      scratch_mib;   /* 0: default.  N = 8192 ... 65536: size of the complex64 scratch between the two
                                   kernels of the four-step path, i.e. spectra per batch (tuning runs)          */
    int32_t      reserved[3];
} rpf_config;

#define RPF_FLAG_NONE 0u
/* Experimental (N = 4096, no window): first radix-16 pass on the tensor cores (tcgen05.mma on the exact
 * 8-bit samples, DFT-16 matrix split hi + lo in fp16, fp32 accumulation in TMEM).  Same results within the
 * 1e-5 tolerance; not the default.  Ignored for other N / with a window. */
#define RPF_FLAG_TC_PASS1 1u
/* rpf_push returns as soon as the host->device copies are ENQUEUED; the caller keeps `bytes` untouched
 * until the next rpf_push / rpf_sync / rpf_finish on this context returns (so two buffers are in flight
 * and the PCIe link never idles between buffers).  Without the flag `bytes` is free on return. */
#define RPF_FLAG_PIPELINED_PUSH 2u
/* Kernel choice at N = 512, 1024, 2048 (A/B comparisons; at most one of the three; default: BATCHED4096 at 1024 and
 * 2048, PIPELINE at 512 -- see the table in rpf_create):
 *   RPF_FLAG_LOCKSTEP     the round-1 kernel: two CTAs per SM, __syncthreads between the passes.  Also, for
 *                         N = 8192 ... 65536: the round-1 outer pass (plain loads and stores) instead of the TMA-staged one.
 *   RPF_FLAG_PIPELINE     the two-stream mbarrier pipeline on the generic three-pass geometry (pipen.cu).
 *   RPF_FLAG_BATCHED4096  the 4096-bin kernel's loop with 8 / 4 / 2 consecutive spectra per 4096-sample ring slot (pass 1
 *                         becomes that many radix-2 / 4 / 8 butterflies per thread; passes 2, 3 and both exchanges shared). */
#define RPF_FLAG_LOCKSTEP 4u
#define RPF_FLAG_PIPELINE 8u
#define RPF_FLAG_BATCHED4096 32u
/* Experiment (N = 65536 without a window): the whole transform in one launch of 16-CTA clusters, the outer radix-16
 * handed to the inner 4096-point pipelines through distributed shared memory (st.async + remote mbarriers) instead of
 * a complex64 scratch in HBM.  Same results; measured slower than the default two-kernel path. */
#define RPF_FLAG_CLUSTER65536 16u
/* N = 8192 ... 65536 (N = M * 4096), kernel choice (A/B comparisons; default: SOLO65536 for launches of at least 96
 * spectra, the two-kernel four-step path for shorter ones -- both add into the same private accumulator rows):
 *   RPF_FLAG_SOLO65536  every launch in one kernel, one CTA per spectrum: the outer radix-M is handed to the M inner
 *                       4096-point transforms thread-privately through a 32 M KiB per-CTA scratch that stays in L2, the
 *                       N accumulators of a CTA live in tensor memory (csrc/solo65536.cu).  While
 *                       such a context exists the device's persisting-L2 set-aside is raised to its maximum.
 *   RPF_FLAG_FOURSTEP   every launch on the two-kernel path (outer radix-16 into a complex64 scratch in HBM, 4096-point
 *                       transforms of its rows). */
#define RPF_FLAG_SOLO65536 64u
#define RPF_FLAG_FOURSTEP 128u

/* Plays the role of fftwf_alloc_complex x2 + fftwf_plan_dft_1d(N, in, out,
 * FFTW_FORWARD, FFTW_MEASURE) in Datastore::Datastore (src/datastore.cxx:30-33):
 * builds twiddle tables, allocates device staging and accumulators.
 * The context starts reset with repeats = INT64_MAX. */
int rpf_create(const rpf_config *cfg, rpf_ctx **out);

/* fftwf_destroy_plan + fftwf_free x2 in Datastore::~Datastore
 * (src/datastore.cxx:43-45).  NULL is allowed. */
void rpf_destroy(rpf_ctx *ctx);

/* Start of one Acquisition::run(): std::fill(pwr, 0), repeats_done = 0
 * (src/acquisition.cxx:252-254) and the fresh worker's fft_pointer = 0
 * (src/datastore.cxx:52).  `repeats` is Params::repeats (src/params.h:55):
 * the worker stops after that many spectra (src/datastore.cxx:67). */
int rpf_reset(rpf_ctx *ctx, int64_t repeats);

/* The worker body for one Buffer (src/datastore.cxx:66-89): consume `len`
 * bytes (interleaved I,Q uint8, bias 127) of the continuous sample stream.
 * Spectrum boundaries are independent of buffer boundaries (fft_pointer
 * persists, src/datastore.cxx:52,68,81,87); bytes arriving after `repeats`
 * spectra are discarded (src/datastore.cxx:67); a trailing partial spectrum is
 * kept for the next push.  `len` must be even (the reference reads
 * buffer[p+1] unconditionally, src/datastore.cxx:74).
 * On return `bytes` may be reused (the host->device copy has completed); the
 * FFT work itself continues asynchronously on the context's stream. */
int rpf_push(rpf_ctx *ctx, const uint8_t *bytes, size_t len);

/* Same, for a buffer already resident in this context's device memory (the
 * HBM-resident benchmark path).  Enqueued on the context's stream; `bytes`
 * must stay valid until rpf_sync()/rpf_finish().  Zero-copy when no partial
 * spectrum is pending and `bytes` is 16-byte aligned. */
int rpf_push_device(rpf_ctx *ctx, const uint8_t *bytes, size_t len);

/* Wait for everything enqueued so far. */
int rpf_sync(rpf_ctx *ctx);

/* What the producer thread reads after t.join() (src/acquisition.cxx:347,
 * 352-357,377-398): pwr_out[N] = Datastore::pwr (summed, not averaged,
 * float64), *repeats_done = Datastore::repeats_done.  Either may be NULL.
 * Does not reset; may be called repeatedly. */
int rpf_finish(rpf_ctx *ctx, double *pwr_out, int64_t *repeats_done);

/* As rpf_finish but pwr_out_device is device memory (N doubles) and the call
 * only enqueues work on the context's stream -- used before a cross-GPU
 * allreduce of pwr[] (the caller's collective runs on the same stream). */
int rpf_finish_device(rpf_ctx *ctx, double *pwr_out_device, int64_t *repeats_done);

/* Datastore::repeats_done (src/datastore.h:38) without synchronising. */
int64_t rpf_repeats_done(const rpf_ctx *ctx);

/* Number of complex samples of a partial spectrum currently held
 * (the reference's local fft_pointer, src/datastore.cxx:52). */
int64_t rpf_fft_pointer(const rpf_ctx *ctx);

/* (single-GPU contexts only) Run the context on a caller-owned CUDA stream (a cudaStream_t passed as
 * void*).  NULL restores the context's own non-blocking stream -- to run on the
 * legacy default stream pass cudaStreamLegacy, not 0.  Synchronises first. */
int rpf_set_stream(rpf_ctx *ctx, void *cuda_stream);

/* Message for the last error on this context (or for create failures when
 * ctx is NULL).  Never NULL. */
const char *rpf_strerror(const rpf_ctx *ctx);

/* ---- one process per GPU: join single-GPU contexts into one integration ----
 * rpf_comm_unique_id fills `id_out` (id_bytes >= 128) with an ncclUniqueId; every rank passes the same id
 * to rpf_comm_init_rank (collective: all `nranks` ranks must call it).  Afterwards rpf_finish*,
 * rpf_write_data and rpf_write_row_f32 return the sum over all ranks and must be called by every rank. */
int rpf_comm_unique_id(void *id_out, size_t id_bytes);
int rpf_comm_init_rank(rpf_ctx *ctx, const void *id, size_t id_bytes, int32_t rank, int32_t nranks);

/* Number of GPUs the context spans. */
int32_t rpf_device_count(const rpf_ctx *ctx);

/* Number of this library's kernel launches issued by the context so far. */
int64_t rpf_kernel_launches(const rpf_ctx *ctx);

/* Name of the kernel family selected for this context ("small", "pipen", "fused3", "pipe4096", "pipe4096_tc",
 * "fourstep", "fourstep_plain_outer", "cluster65536", "direct", "bluestein", "gfft"). */
const char *rpf_plan_name(const rpf_ctx *ctx);

/* ---- Acquisition::write_data() arithmetic on the device ------------------
 * (src/acquisition.cxx:377,390-399): DC-bin interpolation, /repeats_done/N/fs,
 * 10*log10 unless linear, baseline subtraction.  Leaves the context's pwr
 * untouched (works on a copy, as write_data mutates pwr only for output).
 * freq_out / val_out: host arrays of N doubles (freq_out may be NULL);
 * baseline: NULL or N host doubles. */
int rpf_write_data(rpf_ctx *ctx, double tuned_freq, int32_t actual_samplerate, int32_t linear,
                   const double *baseline, double *freq_out, double *val_out);

/* The matrix-mode row of Acquisition::write_data() (src/acquisition.cxx:385-388,400-405): the same values,
 * narrowed to float, N of them, ready to be appended to the .bin file.  row_out: host, N floats. */
int rpf_write_row_f32(rpf_ctx *ctx, int32_t actual_samplerate, int32_t linear, const double *baseline,
                      float *row_out);

/* ---- synthetic IQ source (stands in for Rtlsdr::read, src/device.cxx:92-97) --
 * Counter-based generator, bit-identical on host and device (integer
 * arithmetic only): byte i of the stream depends only on (seed, mode, i).
 *   mode 0: uniform on [0,255]
 *   mode 1: clipped Gaussian, mean 127.4, sigma 30  (a dongle with AGC)
 *   mode 2: tone at +fs/8 (bin N/2 + N/8 for any N divisible by 8), amplitude
 *           100, plus Gaussian noise sigma 5        (exercises bin indexing)
 * `stream_offset` is the absolute byte offset of bytes[0] in the stream. */
void rpf_synth_fill_host(uint8_t *bytes, size_t len, uint64_t stream_offset, uint64_t seed,
                         int32_t mode);
int rpf_synth_fill_device(uint8_t *bytes_device, size_t len, uint64_t stream_offset, uint64_t seed,
                          int32_t mode, int32_t device, void *cuda_stream);

/* Pinned host memory helpers so the caller's Buffers can be DMA targets
 * (Datastore's Buffer pool, src/datastore.cxx:27-28). */
int rpf_host_alloc(void **out, size_t bytes);
void rpf_host_free(void *p);
int rpf_host_register(void *p, size_t bytes);
int rpf_host_unregister(void *p);

int rpf_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RPF_B200_H */
