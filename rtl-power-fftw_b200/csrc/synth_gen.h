// synth_gen.h -- counter-based synthetic IQ byte source, shared by host and
// device code.  Stands in for Rtlsdr::read() (reference src/device.cxx:92-97),
// which needs a USB dongle.  Integer arithmetic only, so host and device
// produce bit-identical streams.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RPF_HD __host__ __device__ __forceinline__
#else
#define RPF_HD inline
#endif

RPF_HD uint64_t rpf_mix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
RPF_HD uint64_t rpf_hash64(uint64_t seed, uint64_t counter) { return rpf_mix64(seed ^ rpf_mix64(counter)); }

// Sum of the 8 bytes of h: mean 1020, sigma 209.0 (Irwin-Hall, close to normal).
RPF_HD int rpf_bytesum(uint64_t h)
{
    h = (h & 0x00FF00FF00FF00FFull) + ((h >> 8) & 0x00FF00FF00FF00FFull);
    h = (h & 0x0000FFFF0000FFFFull) + ((h >> 16) & 0x0000FFFF0000FFFFull);
    return (int)((h & 0xFFFFFFFFull) + (h >> 32));
}
RPF_HD int rpf_clamp_u8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

// 8 consecutive stream bytes starting at the 8-aligned absolute offset 8*w.
RPF_HD uint64_t rpf_synth_word(uint64_t seed, int mode, uint64_t w)
{
    if (mode == 0) return rpf_hash64(seed, w);
    uint64_t out = 0;
    for (int b = 0; b < 8; b++) {
        const uint64_t i = 8 * w + (uint64_t)b;
        const int g = rpf_bytesum(rpf_hash64(seed, i)) - 1020;
        int v;
        if (mode == 1) {
            v = 127 + ((g * 147 + 410 + 512) >> 10);            // 127.4 + g*30/209
        } else {
            const int ph = (int)((i >> 1) & 7);                    // sample index mod 8: fs/8 tone
            const int c8[8] = {100, 71, 0, -71, -100, -71, 0, 71};
            const int tone = (i & 1) ? c8[(ph + 6) & 7] : c8[ph];  // Q = sin, I = cos
            v = 127 + tone + ((g * 49 + 1024) >> 11);            // noise sigma ~5
        }
        out |= (uint64_t)rpf_clamp_u8(v) << (8 * b);
    }
    return out;
}
RPF_HD uint8_t rpf_synth_byte(uint64_t seed, int mode, uint64_t i)
{
    return (uint8_t)(rpf_synth_word(seed, mode, i >> 3) >> (8 * (i & 7)));
}
