/*
 * rpf_oracle.c -- CPU restatement of rtl_power_fftw's FFT worker path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under rtl-power-fftw_b200/ may include,
 * link or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the
 * reported CPU baseline -- never as the product path.
 *
 * PARITY UNPINNED at the FFT boundary: the reference ships no tests, golden
 * vectors or fixtures, and its FFT is libfftw3f
 * (version unpinned, CMakeLists.txt:7; not installed here).  What pins this
 * restatement instead:
 *   (1) oracle/_ref: the reference's *own* src/datastore.cxx compiled from
 *       /root/reference against an fftw3.h shim -- identical loop, carry-over,
 *       truncation, index order and double accumulation; only the DFT routine
 *       is a stand-in (tests/test_oracle.py, tests/test_golden.py),
 *   (2) the DFT definition (forward, exp(-2 pi i nk/N), unnormalised,
 *       FFTW_FORWARD at src/datastore.cxx:33) cross-checked against numpy's
 *       float64 pocketfft and an O(N^2) long-double DFT,
 *   (3) analytic known-answer tests derived from the reference source
 *       (SURVEY.md section 4).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

/* ------------------------------------------------------------------------ */
/* DFT back ends.  X[k] = sum_n x[n] exp(-2 pi i n k / N), unnormalised.     */
/* (FFTW_FORWARD, src/datastore.cxx:32-33; executed at src/datastore.cxx:82) */
/* ------------------------------------------------------------------------ */

static int is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

/* float64 iterative radix-2 DIT, exact-per-index twiddles (no recurrence). */
static void fft_pow2_f64(int N, double *re, double *im)
{
    int lg = 0;
    while ((1 << lg) < N) lg++;
    for (int i = 0; i < N; i++) {           /* bit reversal */
        int j = 0;
        for (int b = 0; b < lg; b++) if (i & (1 << b)) j |= 1 << (lg - 1 - b);
        if (j > i) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    double *wr = (double *)malloc(sizeof(double) * (N / 2 + 1));
    double *wi = (double *)malloc(sizeof(double) * (N / 2 + 1));
    for (int k = 0; k < N / 2; k++) {
        wr[k] = cos(-2.0 * M_PI * (double)k / (double)N);
        wi[k] = sin(-2.0 * M_PI * (double)k / (double)N);
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1, step = N / len;
        for (int base = 0; base < N; base += len)
            for (int j = 0; j < half; j++) {
                double ur = re[base + j], ui = im[base + j];
                double xr = re[base + j + half], xi = im[base + j + half];
                double tr = xr * wr[j * step] - xi * wi[j * step];
                double ti = xr * wi[j * step] + xi * wr[j * step];
                re[base + j] = ur + tr;        im[base + j] = ui + ti;
                re[base + j + half] = ur - tr; im[base + j + half] = ui - ti;
            }
    }
    free(wr); free(wi);
}

/* float32 flavour of the same algorithm: "roughly what a single-precision
 * library gives" -- used to show the oracle's own f32-vs-f64 spread. */
static void fft_pow2_f32(int N, float *re, float *im)
{
    int lg = 0;
    while ((1 << lg) < N) lg++;
    for (int i = 0; i < N; i++) {
        int j = 0;
        for (int b = 0; b < lg; b++) if (i & (1 << b)) j |= 1 << (lg - 1 - b);
        if (j > i) {
            float t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    float *wr = (float *)malloc(sizeof(float) * (N / 2 + 1));
    float *wi = (float *)malloc(sizeof(float) * (N / 2 + 1));
    for (int k = 0; k < N / 2; k++) {
        wr[k] = (float)cos(-2.0 * M_PI * (double)k / (double)N);
        wi[k] = (float)sin(-2.0 * M_PI * (double)k / (double)N);
    }
    for (int len = 2; len <= N; len <<= 1) {
        int half = len >> 1, step = N / len;
        for (int base = 0; base < N; base += len)
            for (int j = 0; j < half; j++) {
                float ur = re[base + j], ui = im[base + j];
                float xr = re[base + j + half], xi = im[base + j + half];
                float tr = xr * wr[j * step] - xi * wi[j * step];
                float ti = xr * wi[j * step] + xi * wr[j * step];
                re[base + j] = ur + tr;        im[base + j] = ui + ti;
                re[base + j + half] = ur - tr; im[base + j + half] = ui - ti;
            }
    }
    free(wr); free(wi);
}

/* Any N (the reference only forces N even, src/params.cxx:152-155; the man
 * page uses -b 500): direct O(N^2) DFT with the phase index reduced mod N in
 * integers, so the argument of cos/sin never loses accuracy. */
static void dft_any_f64(int N, const double *xr, const double *xi, double *Xr, double *Xi)
{
    double *cr = (double *)malloc(sizeof(double) * N);
    double *ci = (double *)malloc(sizeof(double) * N);
    for (int m = 0; m < N; m++) {
        cr[m] = cos(-2.0 * M_PI * (double)m / (double)N);
        ci[m] = sin(-2.0 * M_PI * (double)m / (double)N);
    }
    for (int k = 0; k < N; k++) {
        long double sr = 0.0L, si = 0.0L;
        int64_t idx = 0;
        for (int n = 0; n < N; n++) {
            sr += (long double)xr[n] * cr[idx] - (long double)xi[n] * ci[idx];
            si += (long double)xr[n] * ci[idx] + (long double)xi[n] * cr[idx];
            idx += k; if (idx >= N) idx -= N;
        }
        Xr[k] = (double)sr; Xi[k] = (double)si;
    }
    free(cr); free(ci);
}

/* Exported for direct DFT tests: in-place forward DFT of interleaved
 * complex128 (kind 0/2) ; kind: 0 = f64 fast (pow2) or O(N^2) (other N),
 * 2 = force O(N^2). */
int rpf_oracle_dft(int N, double *re, double *im, int kind)
{
    if (N <= 0) return -1;
    if (kind == 0 && is_pow2(N)) { fft_pow2_f64(N, re, im); return 0; }
    double *Xr = (double *)malloc(sizeof(double) * N);
    double *Xi = (double *)malloc(sizeof(double) * N);
    dft_any_f64(N, re, im, Xr, Xi);
    memcpy(re, Xr, sizeof(double) * N); memcpy(im, Xi, sizeof(double) * N);
    free(Xr); free(Xi);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* The worker: Datastore::fftThread(), src/datastore.cxx:48-96.              */
/* ------------------------------------------------------------------------ */

typedef struct rpf_oracle_state {
    int      N;               /* params.N            (src/params.h:33)        */
    int64_t  repeats;         /* params.repeats      (src/params.h:55)        */
    int      window;          /* params.window       (src/params.h:52)        */
    float   *window_values;   /* Datastore::window_values (src/datastore.h:49)*/
    int      fft_kind;        /* 0: float64 DFT (ground truth), 1: float32    */
    int      fft_pointer;     /* src/datastore.cxx:52 -- persists over buffers*/
    int64_t  repeats_done;    /* Datastore::repeats_done (src/datastore.h:38) */
    float   *inbuf_re, *inbuf_im;   /* complex<float> inbuf (datastore.h:51)  */
    double  *pwr;             /* std::vector<double> pwr (src/datastore.h:53) */
} rpf_oracle_state;

rpf_oracle_state *rpf_oracle_create(int N, int64_t repeats, const float *window_values, int fft_kind)
{
    if (N <= 0 || (N & 1)) return NULL;             /* src/params.cxx:152-155 */
    rpf_oracle_state *s = (rpf_oracle_state *)calloc(1, sizeof(*s));
    s->N = N; s->repeats = repeats; s->fft_kind = fft_kind;
    s->window = window_values != NULL;
    if (s->window) {
        s->window_values = (float *)malloc(sizeof(float) * N);
        memcpy(s->window_values, window_values, sizeof(float) * N);
    }
    s->inbuf_re = (float *)calloc(N, sizeof(float));
    s->inbuf_im = (float *)calloc(N, sizeof(float));
    s->pwr = (double *)calloc(N, sizeof(double));
    return s;
}

void rpf_oracle_destroy(rpf_oracle_state *s)
{
    if (!s) return;
    free(s->window_values); free(s->inbuf_re); free(s->inbuf_im); free(s->pwr); free(s);
}

/* Acquisition::run() reset, src/acquisition.cxx:252-254, plus the fresh
 * worker thread's local fft_pointer = 0 (src/datastore.cxx:52). */
void rpf_oracle_reset(rpf_oracle_state *s, int64_t repeats)
{
    memset(s->pwr, 0, sizeof(double) * s->N);
    s->repeats_done = 0;
    s->fft_pointer = 0;
    s->repeats = repeats;
}

/* One spectrum: fftwf_execute + hot loop B, src/datastore.cxx:82-87. */
static void oracle_execute_and_accumulate(rpf_oracle_state *s)
{
    const int N = s->N;
    if (s->fft_kind == 1 && is_pow2(N)) {
        float *re = (float *)malloc(sizeof(float) * N), *im = (float *)malloc(sizeof(float) * N);
        memcpy(re, s->inbuf_re, sizeof(float) * N); memcpy(im, s->inbuf_im, sizeof(float) * N);
        fft_pow2_f32(N, re, im);
        for (int i = 0; i < N; i++)              /* pow(float,2) -> double, :84 */
            s->pwr[i] += (double)re[i] * (double)re[i] + (double)im[i] * (double)im[i];
        free(re); free(im);
    } else {
        double *re = (double *)malloc(sizeof(double) * N), *im = (double *)malloc(sizeof(double) * N);
        for (int i = 0; i < N; i++) { re[i] = s->inbuf_re[i]; im[i] = s->inbuf_im[i]; }
        rpf_oracle_dft(N, re, im, 0);
        for (int i = 0; i < N; i++)
            s->pwr[i] += re[i] * re[i] + im[i] * im[i];
        free(re); free(im);
    }
    s->repeats_done++;                           /* :86 */
    s->fft_pointer = 0;                          /* :87 */
}

/* One Buffer through the worker body, src/datastore.cxx:66-89.  The buffer
 * is consumed (and then returned to the empty queue, :91-94) whether or not
 * its bytes were used. */
int rpf_oracle_push(rpf_oracle_state *s, const uint8_t *buffer, size_t size)
{
    if (size & 1) return -1;   /* the reference reads buffer[p+1] unconditionally (:74) */
    size_t buffer_pointer = 0;                                        /* :66 */
    while (buffer_pointer < size && s->repeats_done < s->repeats) {   /* :67 */
        while (s->fft_pointer < s->N && buffer_pointer < size) {      /* :68 */
            /* (-1)^n phase flip, :73; bias 127.0 on both rails, :74-75 */
            const float multiplier = (s->fft_pointer % 2 == 0 ? 1.0f : -1.0f);
            float vr = (float)buffer[buffer_pointer], vi = (float)buffer[buffer_pointer + 1];
            vr = (vr - 127.0f) * multiplier;
            vi = (vi - 127.0f) * multiplier;
            if (s->window) {                                          /* :76-77 */
                vr *= s->window_values[s->fft_pointer];
                vi *= s->window_values[s->fft_pointer];
            }
            s->inbuf_re[s->fft_pointer] = vr; s->inbuf_im[s->fft_pointer] = vi;
            buffer_pointer += 2;                                      /* :78 */
            s->fft_pointer++;                                         /* :79 */
        }
        if (s->fft_pointer == s->N)                                   /* :81 */
            oracle_execute_and_accumulate(s);
    }
    return 0;
}

const double *rpf_oracle_pwr(const rpf_oracle_state *s) { return s->pwr; }
int64_t rpf_oracle_repeats_done(const rpf_oracle_state *s) { return s->repeats_done; }
int rpf_oracle_fft_pointer(const rpf_oracle_state *s) { return s->fft_pointer; }

/* ------------------------------------------------------------------------ */
/* Producer-side arithmetic the path depends on.                            */
/* ------------------------------------------------------------------------ */

#define RPF_ORACLE_BASE_BUF 16384            /* src/params.h:26 */
#define RPF_ORACLE_DEFAULT_BUF_MULT 100      /* src/params.h:27 */

/* Plan::Plan buffer auto-sizing, src/acquisition.cxx:166-176. */
int64_t rpf_oracle_plan_buf_length(int N, int64_t repeats, int64_t buf_length, int buf_length_is_set)
{
    if (!buf_length_is_set) {
        int64_t m = (int64_t)ceil((2.0 * N * repeats) / RPF_ORACLE_BASE_BUF);
        if (m <= RPF_ORACLE_DEFAULT_BUF_MULT)
            buf_length = RPF_ORACLE_BASE_BUF * ((m == 0) ? 1 : m);
    }
    return buf_length;
}

/* Sizes of the successive device reads of one Acquisition::run(),
 * src/acquisition.cxx:273-302 (no dropped reads, no strict-time break).
 * Returns the number of reads; writes at most max_out sizes. */
int64_t rpf_oracle_read_sizes(int N, int64_t repeats, int64_t buf_length, int64_t *sizes, int64_t max_out)
{
    int64_t dataTotal = 2 * (int64_t)N * repeats;                     /* :273 */
    int64_t dataRead = 0, n = 0;
    while (dataRead < dataTotal) {                                    /* :276 */
        int64_t dataNeeded = dataTotal - dataRead;                    /* :288 */
        if (dataNeeded >= buf_length) dataNeeded = buf_length;        /* :289-291 */
        else {
            dataNeeded = RPF_ORACLE_BASE_BUF * (int64_t)ceil((double)dataNeeded / RPF_ORACLE_BASE_BUF); /* :295 */
            if (dataNeeded > buf_length) dataNeeded = buf_length;     /* :296-299 */
        }
        if (n < max_out) sizes[n] = dataNeeded;
        n++;
        dataRead += dataNeeded;                                       /* :320 */
    }
    return n;
}

/* Acquisition::write_data numeric part, src/acquisition.cxx:377,390-399:
 * DC-bin interpolation (mutates pwr), frequency axis, normalisation, dB,
 * baseline subtraction.  baseline may be NULL. */
void rpf_oracle_write_data(int N, double *pwr, int64_t repeats_done, double tuned_freq,
                           int actual_samplerate, int linear, const double *baseline,
                           double *freq_out, double *val_out)
{
    pwr[N / 2] = (pwr[N / 2 - 1] + pwr[N / 2 + 1]) / 2;              /* :377 */
    for (int i = 0; i < N; i++) {
        freq_out[i] = tuned_freq + (i - N / 2.0) * actual_samplerate / N;      /* :391 */
        double v = pwr[i] / repeats_done / N / actual_samplerate;              /* :393,397 */
        if (!linear) v = 10 * log10(v);                                        /* :397 */
        val_out[i] = v - (baseline ? baseline[i] : 0);
    }
}

/* Plan::Plan, repeats from the integration time at the true sample rate,
 * src/acquisition.cxx:162-163 (params.repeats is int64_t, the product is
 * evaluated in double). */
int64_t rpf_oracle_plan_repeats(int N, int actual_samplerate, double integration_time)
{
    return (int64_t)ceil(actual_samplerate * integration_time / N);
}

/* Plan::Plan, list of centre frequencies, src/acquisition.cxx:180-198.
 * Returns the number of hops; writes at most max_out frequencies.
 * Types as in the reference: startfreq/stopfreq int64_t (src/params.h:36-37),
 * min_overlap double (percent), hops and overhang int, list of int64_t
 * (src/acquisition.h:53) -- so startfreq + samplerate/2.0 is truncated. */
int rpf_oracle_plan_freqs(int64_t startfreq, int64_t stopfreq, int actual_samplerate, double min_overlap,
                          int64_t *freqs, int max_out)
{
    double min_overhang = actual_samplerate * min_overlap / 100;                                    /* :181 */
    int hops = (int)ceil(((double)(stopfreq - startfreq) - min_overhang) /
                         ((double)actual_samplerate - min_overhang));                               /* :182 */
    int n = 0;
    if (hops > 1) {                                                                                 /* :183 */
        int overhang = (int)(((int64_t)hops * actual_samplerate - (stopfreq - startfreq)) / (hops - 1));   /* :184 */
        int64_t f = (int64_t)(startfreq + actual_samplerate / 2.0);                                 /* :185 */
        if (n < max_out) freqs[n] = f;
        n++;
        for (int hop = 1; hop < hops; hop++) {                                                      /* :187-189 */
            f = f + actual_samplerate - overhang;
            if (n < max_out) freqs[n] = f;
            n++;
        }
    } else {
        if (n < max_out) freqs[n] = (startfreq + stopfreq) / 2;                                     /* :192 */
        n++;
    }
    return n;
}
