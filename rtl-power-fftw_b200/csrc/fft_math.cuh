// fft_math.cuh -- in-register radix-2/4/8/16 DFT butterflies on packed FP32.
//
// One complex value = one 64-bit register pair (float2: x = re, y = im).
// sm_100a executes FADD2/FMUL2/FFMA2 on such pairs in a single issue slot, and
// (verified in SASS) takes a free half-swap (.LO_HI), a free per-half negate
// (.NP) and a free scalar broadcast (Rn.F32 / immediate) on its operands.  So
//   complex add/sub            = 1 FADD2
//   multiply by -i / +i        = folded into the consumer's operand modifiers
//   complex * (wr + i wi)      = FMUL2 + FFMA2 with *scalar* twiddle registers
// and a radix-4 butterfly is 8 issue slots, a radix-16 is 80.
//
// Forward transform, exponent sign -1, unnormalised: the contract of
// fftwf_plan_dft_1d(..., FFTW_FORWARD, ...) at reference src/datastore.cxx:32-33.
#pragma once
#include <cuda_runtime.h>

typedef float2 c32;

__device__ __forceinline__ c32 cadd(c32 a, c32 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ c32 csub(c32 a, c32 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
// a + (-i) b  and  a + (+i) b
__device__ __forceinline__ c32 cadd_mi(c32 a, c32 b) { return __fadd2_rn(a, make_float2(b.y, -b.x)); }
__device__ __forceinline__ c32 cadd_pi(c32 a, c32 b) { return __fadd2_rn(a, make_float2(-b.y, b.x)); }
// (ar + i ai)(wr + i wi) = (ar wr - ai wi) + i (ai wr + ar wi)
__device__ __forceinline__ c32 cmul(c32 a, float wr, float wi)
{
    const c32 t = __fmul2_rn(a, make_float2(wr, wr));
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-wi, wi), t);
}
__device__ __forceinline__ c32 cmul(c32 a, c32 w) { return cmul(a, w.x, w.y); }

// cos(2 pi j / 16), j = 0..15, as compile-time constants.
__host__ __device__ constexpr float rpf_cos16(int j)
{
    return j == 0 ? 1.0f
         : j == 1 ? 0.923879532511286756f
         : j == 2 ? 0.707106781186547524f
         : j == 3 ? 0.382683432365089772f
         : j == 4 ? 0.0f
         : j == 5 ? -0.382683432365089772f
         : j == 6 ? -0.707106781186547524f
         : j == 7 ? -0.923879532511286756f
         : j == 8 ? -1.0f
         : j == 9 ? -0.923879532511286756f
         : j == 10 ? -0.707106781186547524f
         : j == 11 ? -0.382683432365089772f
         : j == 12 ? 0.0f
         : j == 13 ? 0.382683432365089772f
         : j == 14 ? 0.707106781186547524f
         : 0.923879532511286756f;
}
// multiply by W_16^J = exp(-2 pi i J / 16), J a compile-time constant
template <int J>
__device__ __forceinline__ c32 mul_w16(c32 a)
{
    constexpr int j = ((J % 16) + 16) % 16;
    if (j == 0) return a;
    if (j == 4) return make_float2(a.y, -a.x);     // -i
    if (j == 8) return make_float2(-a.x, -a.y);    // -1
    if (j == 12) return make_float2(-a.y, a.x);    // +i
    // wr = cos(2 pi j/16), wi = -sin(2 pi j/16) = -cos(2 pi (j-4)/16)
    return cmul(a, rpf_cos16(j), -rpf_cos16((j + 12) % 16));
}

template <int R> struct Dft;

template <> struct Dft<1> {
    static __device__ __forceinline__ void run(c32 *) {}
};
template <> struct Dft<2> {
    static __device__ __forceinline__ void run(c32 *x)
    {
        const c32 a = x[0], b = x[1];
        x[0] = cadd(a, b);
        x[1] = csub(a, b);
    }
};
template <> struct Dft<4> {
    static __device__ __forceinline__ void run(c32 *x)
    {
        const c32 s0 = cadd(x[0], x[2]), s1 = csub(x[0], x[2]);
        const c32 s2 = cadd(x[1], x[3]), s3 = csub(x[1], x[3]);
        x[0] = cadd(s0, s2);
        x[1] = cadd_mi(s1, s3);
        x[2] = csub(s0, s2);
        x[3] = cadd_pi(s1, s3);
    }
};

// Cooley-Tukey inside a thread: R = R1*R2, n = R2*n1 + n0, k = k0 + R1*k1.
//   A[n0][k0] = DFT_R1 over n1;  A *= W_R^{n0 k0};  X[k0 + R1 k1] = DFT_R2 over n0.
template <int R1, int R2>
__device__ __forceinline__ void dft_composite(c32 *x)
{
    constexpr int R = R1 * R2;
    constexpr int S = 16 / R;   // W_R^j = W_16^{S j}
    c32 a[R2][R1];
#pragma unroll
    for (int n0 = 0; n0 < R2; n0++) {
        c32 col[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; n1++) col[n1] = x[R2 * n1 + n0];
        Dft<R1>::run(col);
#pragma unroll
        for (int k0 = 0; k0 < R1; k0++) a[n0][k0] = col[k0];
    }
    // twiddles: n0*k0 products are compile-time after full unrolling; the
    // switch collapses to one case per (n0,k0).
#pragma unroll
    for (int n0 = 1; n0 < R2; n0++) {
#pragma unroll
        for (int k0 = 1; k0 < R1; k0++) {
            const int j = (S * n0 * k0) % 16;
            c32 v = a[n0][k0];
            switch (j) {
                case 0: break;
                case 1: v = mul_w16<1>(v); break;
                case 2: v = mul_w16<2>(v); break;
                case 3: v = mul_w16<3>(v); break;
                case 4: v = mul_w16<4>(v); break;
                case 5: v = mul_w16<5>(v); break;
                case 6: v = mul_w16<6>(v); break;
                case 7: v = mul_w16<7>(v); break;
                case 8: v = mul_w16<8>(v); break;
                case 9: v = mul_w16<9>(v); break;
                case 10: v = mul_w16<10>(v); break;
                case 11: v = mul_w16<11>(v); break;
                case 12: v = mul_w16<12>(v); break;
                case 13: v = mul_w16<13>(v); break;
                case 14: v = mul_w16<14>(v); break;
                default: v = mul_w16<15>(v); break;
            }
            a[n0][k0] = v;
        }
    }
#pragma unroll
    for (int k0 = 0; k0 < R1; k0++) {
        c32 row[R2];
#pragma unroll
        for (int n0 = 0; n0 < R2; n0++) row[n0] = a[n0][k0];
        Dft<R2>::run(row);
#pragma unroll
        for (int k1 = 0; k1 < R2; k1++) x[k0 + R1 * k1] = row[k1];
    }
}
template <> struct Dft<8> {
    static __device__ __forceinline__ void run(c32 *x) { dft_composite<2, 4>(x); }
};
__device__ __forceinline__ void dft16_first(c32 *x);
__device__ __forceinline__ void dft16_second(c32 *x);
template <> struct Dft<16> {   // two radix-4 layers, internal twiddles in FMA form (75 packed ops)
    static __device__ __forceinline__ void run(c32 *x) { dft16_first(x); dft16_second(x); }
};

// Radix-16 in two halves, so that a caller can issue independent memory operations in between
// (software pipelining).  dft16_first: 4 radix-4 butterflies over n1 (x[4 n1 + n0]), result
// a[n0][k0] left in x[4 k0 + n0] -- WITHOUT the internal W_16^{n0 k0} twiddles.  dft16_second applies
// those twiddles fused into its radix-4 butterflies (FMA form: s0 = a0 + w2 a2 is two FFMA2,
// s1 = 2 a0 - s0 one more; 43 packed ops instead of 48) and leaves X[k0 + 4 k1] in natural order.
// dft16_first followed by dft16_second == Dft<16>::run up to rounding.
__device__ __forceinline__ void dft16_first(c32 *x)
{
    c32 a[16];
#pragma unroll
    for (int n0 = 0; n0 < 4; n0++) {
        c32 col[4] = {x[n0], x[4 + n0], x[8 + n0], x[12 + n0]};
        Dft<4>::run(col);
#pragma unroll
        for (int k0 = 0; k0 < 4; k0++) a[4 * k0 + n0] = col[k0];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = a[i];
}
// columns [C0, C1) of dft16_first, in place (column n0 touches x[n0], x[4+n0], x[8+n0], x[12+n0] only)
template <int C0, int C1>
__device__ __forceinline__ void dft16_first_cols(c32 *x)
{
#pragma unroll
    for (int n0 = C0; n0 < C1; n0++) {
        c32 col[4] = {x[n0], x[4 + n0], x[8 + n0], x[12 + n0]};
        Dft<4>::run(col);
#pragma unroll
        for (int k0 = 0; k0 < 4; k0++) x[4 * k0 + n0] = col[k0];
    }
}
// dft16_first on x[i] * w[i] (w real): the window multiplication rides in the butterflies' first adds,
//   s0 = x0 w0 + x2 w2 = FFMA2(x2, w2, FMUL2(x0, w0)),  s1 = FFMA2(x2, -w2, ...)  -- 14 packed ops per radix-4
// instead of 4 + 8, with one rounding less per sum than multiply-then-add.
template <int NW>
__device__ __forceinline__ void dft16_first_windowed(c32 *x, const float (&w)[NW])
{
    static_assert(NW == 16 || NW == 1, "window registers");
#pragma unroll
    for (int n0 = 0; n0 < 4; n0++) {
        const float w0 = w[(0 + n0) % NW], w1 = w[(4 + n0) % NW], w2 = w[(8 + n0) % NW], w3 = w[(12 + n0) % NW];
        const c32 t0 = __fmul2_rn(x[n0], make_float2(w0, w0)), t1 = __fmul2_rn(x[4 + n0], make_float2(w1, w1));
        const c32 s0 = __ffma2_rn(x[8 + n0], make_float2(w2, w2), t0), s1 = __ffma2_rn(x[8 + n0], make_float2(-w2, -w2), t0);
        const c32 s2 = __ffma2_rn(x[12 + n0], make_float2(w3, w3), t1), s3 = __ffma2_rn(x[12 + n0], make_float2(-w3, -w3), t1);
        x[n0] = cadd(s0, s2);
        x[4 + n0] = cadd_mi(s1, s3);
        x[8 + n0] = csub(s0, s2);
        x[12 + n0] = cadd_pi(s1, s3);
    }
}
// a + w b with w = W_16^J (compile-time): two FFMA2
template <int J>
__device__ __forceinline__ c32 fma_w16(c32 a, c32 b)
{
    constexpr int j = ((J % 16) + 16) % 16;
    constexpr float wr = rpf_cos16(j), wi = -rpf_cos16((j + 12) % 16);
    const c32 t = __ffma2_rn(b, make_float2(wr, wr), a);
    return __ffma2_rn(make_float2(b.y, b.x), make_float2(-wi, wi), t);
}
// 2 a - s
__device__ __forceinline__ c32 twice_minus(c32 a, c32 s) { return __ffma2_rn(a, make_float2(2.0f, 2.0f), make_float2(-s.x, -s.y)); }

// radix-4 over (a0, w1 a1, w2 a2, w3 a3) with w_q = W_16^{q K0}
template <int K0>
__device__ __forceinline__ void radix4_tw(c32 a0, c32 a1, c32 a2, c32 a3, c32 &y0, c32 &y1, c32 &y2, c32 &y3)
{
    c32 s0, s1, s2, s3;
    if (K0 == 0) {
        s0 = cadd(a0, a2); s1 = csub(a0, a2); s2 = cadd(a1, a3); s3 = csub(a1, a3);
    } else {
        if ((2 * K0) % 16 == 4) {            // w2 = -i
            s0 = cadd_mi(a0, a2); s1 = cadd_pi(a0, a2);
        } else {
            s0 = fma_w16<2 * K0>(a0, a2); s1 = twice_minus(a0, s0);
        }
        const c32 p1 = mul_w16<K0>(a1);
        s2 = fma_w16<3 * K0>(p1, a3); s3 = twice_minus(p1, s2);
    }
    y0 = cadd(s0, s2);
    y1 = cadd_mi(s1, s3);
    y2 = csub(s0, s2);
    y3 = cadd_pi(s1, s3);
}
// dft16_second in two halves (a caller may put memory operations in between): y is scratch, the
// result is in x after dft16_second_b.
__device__ __forceinline__ void dft16_second_a(const c32 *x, c32 *y)
{
    radix4_tw<0>(x[0], x[1], x[2], x[3], y[0], y[4], y[8], y[12]);
    radix4_tw<1>(x[4], x[5], x[6], x[7], y[1], y[5], y[9], y[13]);
}
__device__ __forceinline__ void dft16_second_b(c32 *x, c32 *y)
{
    radix4_tw<2>(x[8], x[9], x[10], x[11], y[2], y[6], y[10], y[14]);
    radix4_tw<3>(x[12], x[13], x[14], x[15], y[3], y[7], y[11], y[15]);
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = y[i];
}
__device__ __forceinline__ void dft16_second(c32 *x)
{
    c32 y[16];
    dft16_second_a(x, y);
    dft16_second_b(x, y);
}
