/*
 * shim_fft.cpp -- stand-in for libfftw3f behind oracle/ref_shim/fftw3.h.
 * TEST INFRASTRUCTURE ONLY.  Single-precision, forward, unnormalised,
 * out-of-place complex DFT, same contract as fftwf_plan_dft_1d(...,
 * FFTW_FORWARD, ...) + fftwf_execute (src/datastore.cxx:32-33,82).
 *
 * Power-of-two N: Stockham autosort, radix-4 (+ one radix-2 stage when log2 N
 * is odd), split re/im work arrays so gcc vectorises the inner loops, twiddles
 * tabulated from double precision.  Other N: O(N^2) DFT in double.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "fftw3.h"

struct rpf_shim_plan_s {
    int n;
    float *in, *out;             // interleaved complex<float>
    bool pow2;
    std::vector<float> ar, ai, br, bi;          // ping-pong SoA work arrays
    std::vector<std::vector<float>> twr, twi;   // per-stage twiddles
    std::vector<int> radix;                     // per-stage radix (4 or 2)
    std::vector<double> cr, ci;                 // O(N^2) path table
};

extern "C" fftwf_complex *fftwf_alloc_complex(unsigned long n)
{
    void *p = nullptr;
    if (posix_memalign(&p, 64, sizeof(fftwf_complex) * (n ? n : 1)) != 0) return nullptr;
    return (fftwf_complex *)p;
}
extern "C" void fftwf_free(void *p) { free(p); }

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
    if (sign != FFTW_FORWARD || n <= 0) return nullptr;
    rpf_shim_plan_s *p = new rpf_shim_plan_s;
    p->n = n; p->in = (float *)in; p->out = (float *)out;
    p->pow2 = (n & (n - 1)) == 0;
    const double PI = 3.14159265358979323846264338327950288;
    if (p->pow2) {
        p->ar.resize(n); p->ai.resize(n); p->br.resize(n); p->bi.resize(n);
        int lg = 0; while ((1 << lg) < n) lg++;
        int l = 1;
        int rem = lg;
        while (rem > 0) {
            int r = (rem >= 2) ? 4 : 2;
            if (rem & 1) { r = 2; }          // do the single radix-2 stage first
            int m = n / (l * r);
            std::vector<float> wr((size_t)l * (r - 1)), wi((size_t)l * (r - 1));
            for (int j = 0; j < l; j++)
                for (int q = 1; q < r; q++) {
                    double a = -2.0 * PI * (double)(j * q) / (double)(l * r);
                    wr[(size_t)(q - 1) * l + j] = (float)std::cos(a);
                    wi[(size_t)(q - 1) * l + j] = (float)std::sin(a);
                }
            p->twr.push_back(wr); p->twi.push_back(wi); p->radix.push_back(r);
            (void)m;
            l *= r; rem -= (r == 4) ? 2 : 1;
        }
    } else {
        p->cr.resize(n); p->ci.resize(n);
        for (int m = 0; m < n; m++) {
            p->cr[m] = std::cos(-2.0 * PI * (double)m / (double)n);
            p->ci[m] = std::sin(-2.0 * PI * (double)m / (double)n);
        }
    }
    return p;
}

extern "C" void fftwf_destroy_plan(fftwf_plan plan) { delete plan; }

// One Stockham DIT stage.  Before the stage the array holds, for every
// k in [0, n/l), the l-point DFT of the stride-(n/l) subsequence starting at k,
// stored as X_l[j][k] at index j*(n/l) + k.  With m = n/(l*r) that is the view
// x[j][q][k] = x[(j*r + q)*m + k]; the stage forms
//   y[q'][j][k] = sum_q  w_{l r}^{j q} x[j][q][k]  W_r^{q q'}
// at index (q'*l + j)*m + k.  The contiguous (vectorised) index is k.
static inline void stage4(int l, int m, const float *xr, const float *xi, float *yr, float *yi,
                          const float *wr, const float *wi)
{
    for (int j = 0; j < l; j++) {
        const float w1r = wr[j], w1i = wi[j], w2r = wr[l + j], w2i = wi[l + j], w3r = wr[2 * l + j], w3i = wi[2 * l + j];
        const float *x0r = xr + (size_t)(j * 4 + 0) * m, *x0i = xi + (size_t)(j * 4 + 0) * m;
        const float *x1r = xr + (size_t)(j * 4 + 1) * m, *x1i = xi + (size_t)(j * 4 + 1) * m;
        const float *x2r = xr + (size_t)(j * 4 + 2) * m, *x2i = xi + (size_t)(j * 4 + 2) * m;
        const float *x3r = xr + (size_t)(j * 4 + 3) * m, *x3i = xi + (size_t)(j * 4 + 3) * m;
        float *y0r = yr + (size_t)(0 * l + j) * m, *y0i = yi + (size_t)(0 * l + j) * m;
        float *y1r = yr + (size_t)(1 * l + j) * m, *y1i = yi + (size_t)(1 * l + j) * m;
        float *y2r = yr + (size_t)(2 * l + j) * m, *y2i = yi + (size_t)(2 * l + j) * m;
        float *y3r = yr + (size_t)(3 * l + j) * m, *y3i = yi + (size_t)(3 * l + j) * m;
        for (int k = 0; k < m; k++) {
            float ar = x0r[k], ai = x0i[k];
            float br = x1r[k] * w1r - x1i[k] * w1i, bi = x1r[k] * w1i + x1i[k] * w1r;
            float cr = x2r[k] * w2r - x2i[k] * w2i, ci = x2r[k] * w2i + x2i[k] * w2r;
            float dr = x3r[k] * w3r - x3i[k] * w3i, di = x3r[k] * w3i + x3i[k] * w3r;
            float s0r = ar + cr, s0i = ai + ci, s1r = ar - cr, s1i = ai - ci;
            float s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
            y0r[k] = s0r + s2r; y0i[k] = s0i + s2i;
            y1r[k] = s1r + s3i; y1i[k] = s1i - s3r;      // s1 - i*s3
            y2r[k] = s0r - s2r; y2i[k] = s0i - s2i;
            y3r[k] = s1r - s3i; y3i[k] = s1i + s3r;      // s1 + i*s3
        }
    }
}
static inline void stage2(int l, int m, const float *xr, const float *xi, float *yr, float *yi,
                          const float *wr, const float *wi)
{
    for (int j = 0; j < l; j++) {
        const float w1r = wr[j], w1i = wi[j];
        const float *x0r = xr + (size_t)(j * 2) * m, *x0i = xi + (size_t)(j * 2) * m;
        const float *x1r = xr + (size_t)(j * 2 + 1) * m, *x1i = xi + (size_t)(j * 2 + 1) * m;
        float *y0r = yr + (size_t)j * m, *y0i = yi + (size_t)j * m;
        float *y1r = yr + (size_t)(l + j) * m, *y1i = yi + (size_t)(l + j) * m;
        for (int k = 0; k < m; k++) {
            float br = x1r[k] * w1r - x1i[k] * w1i, bi = x1r[k] * w1i + x1i[k] * w1r;
            y0r[k] = x0r[k] + br; y0i[k] = x0i[k] + bi;
            y1r[k] = x0r[k] - br; y1i[k] = x0i[k] - bi;
        }
    }
}

extern "C" void fftwf_execute(const fftwf_plan p)
{
    const int n = p->n;
    if (!p->pow2) {
        for (int k = 0; k < n; k++) {
            double sr = 0, si = 0; long idx = 0;
            for (int t = 0; t < n; t++) {
                double xr = p->in[2 * t], xi = p->in[2 * t + 1];
                sr += xr * p->cr[idx] - xi * p->ci[idx];
                si += xr * p->ci[idx] + xi * p->cr[idx];
                idx += k; if (idx >= n) idx -= n;
            }
            p->out[2 * k] = (float)sr; p->out[2 * k + 1] = (float)si;
        }
        return;
    }
    float *xr = p->ar.data(), *xi = p->ai.data(), *yr = p->br.data(), *yi = p->bi.data();
    for (int t = 0; t < n; t++) { xr[t] = p->in[2 * t]; xi[t] = p->in[2 * t + 1]; }
    int l = 1;
    for (size_t s = 0; s < p->radix.size(); s++) {
        int r = p->radix[s], m = n / (l * r);
        if (r == 4) stage4(l, m, xr, xi, yr, yi, p->twr[s].data(), p->twi[s].data());
        else        stage2(l, m, xr, xi, yr, yi, p->twr[s].data(), p->twi[s].data());
        float *t;
        t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t;
        l *= r;
    }
    for (int t = 0; t < n; t++) { p->out[2 * t] = xr[t]; p->out[2 * t + 1] = xi[t]; }
}
