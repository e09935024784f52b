// tc_common.cuh — what the two tcgen05 kernels (mfcc_tc.cu: DCT, mel_tc.cu: mel projection) share: UMMA shared-memory
// descriptors, the TF32 MMA issue, the 3xTF32 operand split, mbarrier waits.
#pragma once
#include "common.cuh"

namespace acids {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), K-major, SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address, bits [0, 14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;       // leading byte offset, bits [16, 30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;       // stride byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
    return d;                                               // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, dense, M x N
__host__ __device__ constexpr uint32_t make_idesc_mn(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);     // sign + exponent + 10 mantissa bits
    lo = x - hi;                                                // exact; the tensor core reads its TF32 prefix
}

__device__ __forceinline__ void mbar_wait_parity(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}

}  // namespace tc
}  // namespace acids
