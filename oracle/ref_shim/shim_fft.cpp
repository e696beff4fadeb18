/*
 * shim_fft.cpp -- stand-in for libfftw3f behind oracle/ref_shim/fftw3.h.
 * TEST INFRASTRUCTURE ONLY.  Single-precision, forward, unnormalised,
 * out-of-place complex DFT, same contract as fftwf_plan_dft_1d(...,
 * FFTW_FORWARD, ...) + fftwf_execute (src/datastore.cxx:32-33,82).
 *
 * Power-of-two N >= 64: four-step N = N1 x N2: batched Stockham autosort (radix-4,
 * + one radix-2 stage when the log is odd) down the columns of the N1 x N2 view,
 * twiddle + transpose, batched Stockham down the columns of the N2 x N1 view.
 * Every butterfly loop runs over a contiguous run of >= min(N1,N2) floats, so gcc
 * vectorises all of them (split re/im arrays, twiddles tabulated from double).
 * Smaller power-of-two N: plain Stockham.  Other N: O(N^2) DFT in double.
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
    // four-step
    int n1 = 0, n2 = 0;
    std::vector<std::vector<float>> twr1, twi1, twr2, twi2;   // per-stage twiddles of the two column passes
    std::vector<int> radix1, radix2;
    std::vector<float> fwr, fwi;                // W_N^{n2 k1}, [k1][n2]
};

extern "C" fftwf_complex *fftwf_alloc_complex(unsigned long n)
{
    void *p = nullptr;
    if (posix_memalign(&p, 64, sizeof(fftwf_complex) * (n ? n : 1)) != 0) return nullptr;
    return (fftwf_complex *)p;
}
extern "C" void fftwf_free(void *p) { free(p); }

static void build_stages(int n, std::vector<int> &radix, std::vector<std::vector<float>> &twr,
                         std::vector<std::vector<float>> &twi)
{
    const double PI = 3.14159265358979323846264338327950288;
    int lg = 0; while ((1 << lg) < n) lg++;
    int l = 1, rem = lg;
    while (rem > 0) {
        int r = (rem & 1) ? 2 : 4;               // the single radix-2 stage (odd log) goes first
        std::vector<float> wr((size_t)l * (r - 1)), wi((size_t)l * (r - 1));
        for (int j = 0; j < l; j++)
            for (int q = 1; q < r; q++) {
                double a = -2.0 * PI * (double)(j * q) / (double)(l * r);
                wr[(size_t)(q - 1) * l + j] = (float)std::cos(a);
                wi[(size_t)(q - 1) * l + j] = (float)std::sin(a);
            }
        twr.push_back(wr); twi.push_back(wi); radix.push_back(r);
        l *= r; rem -= (r == 4) ? 2 : 1;
    }
}

extern "C" fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex *in, fftwf_complex *out, int sign, unsigned)
{
    if (sign != FFTW_FORWARD || n <= 0) return nullptr;
    rpf_shim_plan_s *p = new rpf_shim_plan_s;
    p->n = n; p->in = (float *)in; p->out = (float *)out;
    p->pow2 = (n & (n - 1)) == 0;
    const double PI = 3.14159265358979323846264338327950288;
    if (p->pow2) {
        p->ar.resize(n); p->ai.resize(n); p->br.resize(n); p->bi.resize(n);
        if (n >= 64) {
            int lg = 0; while ((1 << lg) < n) lg++;
            p->n1 = 1 << ((lg + 1) / 2); p->n2 = n / p->n1;
            build_stages(p->n1, p->radix1, p->twr1, p->twi1);
            build_stages(p->n2, p->radix2, p->twr2, p->twi2);
            p->fwr.resize(n); p->fwi.resize(n);
            for (int k1 = 0; k1 < p->n1; k1++)
                for (int c = 0; c < p->n2; c++) {
                    double a = -2.0 * PI * (double)((long)k1 * c % n) / (double)n;
                    p->fwr[(size_t)k1 * p->n2 + c] = (float)std::cos(a);
                    p->fwi[(size_t)k1 * p->n2 + c] = (float)std::sin(a);
                }
        } else {
            build_stages(n, p->radix, p->twr, p->twi);
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
// unit-stride radix-4 butterfly over a run of m elements (kept in its own function with restrict
// pointers so that gcc vectorises it)
static inline void bfly4_run(int m, const float *__restrict__ x0r, const float *__restrict__ x0i,
                             const float *__restrict__ x1r, const float *__restrict__ x1i,
                             const float *__restrict__ x2r, const float *__restrict__ x2i,
                             const float *__restrict__ x3r, const float *__restrict__ x3i,
                             float *__restrict__ y0r, float *__restrict__ y0i, float *__restrict__ y1r,
                             float *__restrict__ y1i, float *__restrict__ y2r, float *__restrict__ y2i,
                             float *__restrict__ y3r, float *__restrict__ y3i, float w1r, float w1i, float w2r,
                             float w2i, float w3r, float w3i)
{
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
static inline void stage4(int l, int m, const float *xr, const float *xi, float *yr, float *yi,
                          const float *wr, const float *wi)
{
    for (int j = 0; j < l; j++) {
        const size_t xb = (size_t)j * 4 * m, yb = (size_t)j * m, ys = (size_t)l * m;
        bfly4_run(m, xr + xb, xi + xb, xr + xb + m, xi + xb + m, xr + xb + 2 * (size_t)m, xi + xb + 2 * (size_t)m,
                  xr + xb + 3 * (size_t)m, xi + xb + 3 * (size_t)m, yr + yb, yi + yb, yr + yb + ys, yi + yb + ys,
                  yr + yb + 2 * ys, yi + yb + 2 * ys, yr + yb + 3 * ys, yi + yb + 3 * ys, wr[j], wi[j], wr[l + j],
                  wi[l + j], wr[2 * l + j], wi[2 * l + j]);
    }
}
static inline void stage2(int l, int m, const float *__restrict__ xr, const float *__restrict__ xi, float *__restrict__ yr, float *__restrict__ yi,
                          const float *__restrict__ wr, const float *__restrict__ wi)
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
    // Batched Stockham: a length-L transform down the columns of an L x C row-major array is the
    // plain Stockham recursion with every run length m multiplied by C.
    auto columns = [&](int L, int C, const std::vector<int> &radix, const std::vector<std::vector<float>> &twr,
                       const std::vector<std::vector<float>> &twi) {
        int l = 1;
        for (size_t s = 0; s < radix.size(); s++) {
            int r = radix[s], m = L / (l * r);
            if (r == 4) stage4(l, m * C, xr, xi, yr, yi, twr[s].data(), twi[s].data());
            else        stage2(l, m * C, xr, xi, yr, yi, twr[s].data(), twi[s].data());
            float *t;
            t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t;
            l *= r;
        }
    };
    if (p->n1 == 0) {
        columns(n, 1, p->radix, p->twr, p->twi);
    } else {
        const int N1 = p->n1, N2 = p->n2;
        columns(N1, N2, p->radix1, p->twr1, p->twi1);            // Y[k1][n2], n = n1*N2 + n2
        for (int k1 = 0; k1 < N1; k1++)                           // T[n2][k1] = Y[k1][n2] * W_N^{n2 k1}
            for (int c = 0; c < N2; c++) {
                const float a = xr[(size_t)k1 * N2 + c], b = xi[(size_t)k1 * N2 + c];
                const float wr = p->fwr[(size_t)k1 * N2 + c], wi = p->fwi[(size_t)k1 * N2 + c];
                yr[(size_t)c * N1 + k1] = a * wr - b * wi;
                yi[(size_t)c * N1 + k1] = a * wi + b * wr;
            }
        { float *t; t = xr; xr = yr; yr = t; t = xi; xi = yi; yi = t; }
        columns(N2, N1, p->radix2, p->twr2, p->twi2);            // X[k2][k1] = out[k1 + N1*k2]
    }
    for (int t = 0; t < n; t++) { p->out[2 * t] = xr[t]; p->out[2 * t + 1] = xi[t]; }
}
