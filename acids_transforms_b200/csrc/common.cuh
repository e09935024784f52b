// common.cuh — error plumbing, launch helpers and the shared epilogue of the spectral kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/acids_b200.h"
#include "fft_core.cuh"

namespace acids {

void set_error(const char* fmt, ...);           // capi.cu (thread-local message)
int num_sms();                                  // cached cudaDevAttrMultiProcessorCount

#define ACIDS_REQUIRE(cond, code, ...)   \
    do {                                 \
        if (!(cond)) {                   \
            acids::set_error(__VA_ARGS__); \
            return (code);               \
        }                                \
    } while (0)

#define ACIDS_CHECK_LAUNCH(what)                                                  \
    do {                                                                          \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            acids::set_error("%s: CUDA error: %s", what, cudaGetErrorString(e__)); \
            return ACIDS_ECUDA;                                                   \
        }                                                                         \
    } while (0)

// Barrier over the T threads that share one frame.  T <= 32: the groups of a warp run in
// lockstep through the same sequence, a warp barrier suffices.  Larger groups are warp aligned
// and use their own named barrier (id 0 stays reserved for __syncthreads()).
template <int T, int THREADS>
__device__ __forceinline__ void group_sync(int group) {
    if (T <= 32) {
        __syncwarp();
    } else if (T == THREADS) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(T) : "memory");
    }
}

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream2(float2* p, float x, float y) {
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float x) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
}

// ---- epilogue shared by the fused STFT kernel and the stand-alone Magnitude kernel -------------
struct EpiParams {
    const int32_t* meta;     // banded matrix (or nullptr)
    const float* coef;
    int n_cols;              // columns before drop_first
    int contrast;
    float eps;
    float offset;            // already loaded from the device scalars
    float inv_scale;
    int drop_first;
};

__device__ __forceinline__ float apply_contrast(float a, int mode, float eps) {
    // spectral_repr.py:191-201.  log(1 + m) is evaluated literally (not log1p), like the reference.
    if (mode == ACIDS_CONTRAST_LOG1P) return logf(1.0f + a);
    if (mode == ACIDS_CONTRAST_LOG) return logf(fmaxf(a, eps));
    if (mode == ACIDS_CONTRAST_LOG10) return log10f(fmaxf(a, eps));
    return a;
}

__device__ __forceinline__ float invert_contrast(float y, int mode, float eps) {
    // spectral_repr.py:203-213
    if (mode == ACIDS_CONTRAST_LOG1P) return expf(y) - 1.0f;
    if (mode == ACIDS_CONTRAST_LOG) return expf(y) - eps;
    if (mode == ACIDS_CONTRAST_LOG10) return powf(10.0f, y);
    return y;
}

// One row: `val` holds n_in non-negative values in shared memory (|X| or |X|^p); the T threads of
// the group produce columns tid, tid+T, ... : banded projection -> contrast -> normalise -> store.
template <int T>
__device__ __forceinline__ void epilogue_row(const float* __restrict__ val, int tid, const EpiParams& ep,
                                             float* __restrict__ out_row, int64_t col_stride, bool valid) {
    for (int m = tid; m < ep.n_cols; m += T) {
        float a;
        if (ep.meta != nullptr) {
            const int2 me = __ldg(reinterpret_cast<const int2*>(ep.meta) + m);
            const int start = me.x & 0xffff, cnt = me.x >> 16;
            const float* c = ep.coef + me.y;
            a = 0.f;
            for (int u = 0; u < cnt; ++u) a = fmaf(val[start + u], __ldg(c + u), a);
        } else {
            a = val[m];
        }
        a = apply_contrast(a, ep.contrast, ep.eps);
        a = (a - ep.offset) * ep.inv_scale;
        if (valid && m >= ep.drop_first) stg_stream1(out_row + (int64_t)(m - ep.drop_first) * col_stride, a);
    }
}

__device__ __forceinline__ void load_norm(const float* offset, const float* scale, float& off, float& inv) {
    off = offset ? __ldg(offset) : 0.f;
    inv = scale ? 1.0f / __ldg(scale) : 1.0f;
}

}  // namespace acids
