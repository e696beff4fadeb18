// stage_prims.cuh -- mbarrier / 1-D TMA bulk-copy primitives and the exact u8 -> f32 unpack,
// shared by the fused kernels.
#pragma once
#include <stdint.h>
#include "fft_math.cuh"

// ---- mbarrier / TMA (1-D bulk) primitives ---------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}

// Same wait for a thread with nothing else to do (the bulk-copy lane): the try_wait carries a suspend-time hint, so
// the warp is parked by the hardware instead of re-issuing the probe back to back.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t hint_ns)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
    } while (!done);
}

// u8 pair (I | Q<<8) -> complex float carrying a +32768 offset on both rails:
// PRMT builds 0x4700XX00 = 32768.0f + XX exactly (byte into mantissa bits 15:8).
// No window: the offset (32768+127 per rail) is left in; every butterfly sum of
// <= 16 such values is an exact integer < 2^24, the offset cancels exactly in
// every output of the first radix pass except k1 = 0, where R1*32895 is
// subtracted exactly (fix_bias()).  Result: bit-identical to converting and
// subtracting 127.0f per sample (src/datastore.cxx:74-75), at 2 instead of 3
// issue slots per sample.
__device__ __forceinline__ c32 unpack_biased(uint32_t raw)
{
    c32 v;
    v.x = __uint_as_float(__byte_perm(raw, 0x47000000u, 0x7604));
    v.y = __uint_as_float(__byte_perm(raw, 0x47000000u, 0x7614));
    return v;
}
constexpr float kBias = 32768.0f + 127.0f;
