// pointwise.cu — mu-law / one-hot (A18, A19), Normalize statistics (A7), Mono / MidSide (A20) and the
// opt-in MFCC dB + DCT tail (A9).  Pure HBM-bound elementwise / reduction kernels.
#include <float.h>
#include "common.cuh"

namespace acids {

// ---------------------------------------------------------------------------------------------
// mu-law.  Bit-exactness with the eager torch chain needs every float32 rounding of that chain:
// explicit _rn intrinsics keep nvcc from contracting mul+add into FMA.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t mulaw_quantise(float x, float mu, float l1p, float inv_l1p, int recip) {
    // functional.py:698-699: sign(x) * log1p(mu * |x|) / log1p(mu); ((. + 1) / 2 * mu + 0.5).to(int64)
    const float sgn = (x > 0.f) ? 1.f : ((x < 0.f) ? -1.f : 0.f);
    const float num = __fmul_rn(sgn, log1pf(__fmul_rn(mu, fabsf(x))));
    float q = recip ? __fmul_rn(num, inv_l1p) : __fdiv_rn(num, l1p);
    q = __fadd_rn(q, 1.0f);
    q = __fmul_rn(q, 0.5f);       // "/ 2" is exact either way
    q = __fmul_rn(q, mu);
    q = __fadd_rn(q, 0.5f);
    return (int64_t)q;            // truncation toward zero, like .to(torch.int64)
}

// four samples per thread: one 16-byte load, two 16-byte streaming stores (int64 output: 8 B per sample)
__global__ void __launch_bounds__(256) mulaw_encode_kernel(const float* __restrict__ x, int64_t n, float mu, float l1p, int recip,
                                                           int64_t* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, l1p);
    const bool vec = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
        const long long q0 = mulaw_quantise(v.x, mu, l1p, inv, recip), q1 = mulaw_quantise(v.y, mu, l1p, inv, recip);
        const long long q2 = mulaw_quantise(v.z, mu, l1p, inv, recip), q3 = mulaw_quantise(v.w, mu, l1p, inv, recip);
        long long* o = reinterpret_cast<long long*>(out) + 4 * i;
        asm volatile("st.global.L1::no_allocate.v2.s64 [%0], {%1, %2};" ::"l"(o), "l"(q0), "l"(q1) : "memory");
        asm volatile("st.global.L1::no_allocate.v2.s64 [%0], {%1, %2};" ::"l"(o + 2), "l"(q2), "l"(q3) : "memory");
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = mulaw_quantise(__ldg(x + i), mu, l1p, inv, recip);
}

// fused encode + one-hot.  CATEGORICAL: out[i, c]; one warp writes whole 8*C-byte rows with 16-byte
// stores.  CHANNEL: out[o, c, j] with i = o*inner + j; threads run along j for coalescing.
__global__ void mulaw_onehot_categorical_kernel(const float* __restrict__ x, int64_t n, int C, float mu, float l1p,
                                                int recip, int64_t* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, l1p);
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t q = mulaw_quantise(__ldg(x + i), mu, l1p, inv, recip);
        int64_t* row = out + i * C;
        if ((C & 1) == 0) {
            longlong2* row2 = reinterpret_cast<longlong2*>(row);
            for (int c = lane; c < C / 2; c += 32) row2[c] = make_longlong2(q == 2 * c, q == 2 * c + 1);
        } else {
            for (int c = lane; c < C; c += 32) row[c] = (q == c);
        }
    }
}

__global__ void mulaw_onehot_channel_kernel(const float* __restrict__ x, int64_t outer, int64_t inner, int C, float mu,
                                            float l1p, int recip, int64_t* __restrict__ out) {
    const float inv = __fdiv_rn(1.0f, l1p);
    const int64_t n = outer * inner;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = i / inner, j = i - o * inner;
        const int64_t q = mulaw_quantise(__ldg(x + i), mu, l1p, inv, recip);
        int64_t* base = out + o * C * inner + j;
        for (int c = 0; c < C; ++c) base[(int64_t)c * inner] = (q == c);
    }
}

__global__ void mulaw_decode_kernel(const int64_t* __restrict__ q, int64_t n, float mu, float l1p, int recip,
                                    float* __restrict__ out) {
    // functional.py:727-728: x = q / mu * 2 - 1;  sign(x) * (exp(|x| * log1p(mu)) - 1) / mu
    // (both divisions are by a host scalar: the CUDA eager chain multiplies by 1/mu instead)
    const float inv_mu = __fdiv_rn(1.0f, mu);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = recip ? __fmul_rn((float)q[i], inv_mu) : __fdiv_rn((float)q[i], mu);
        v = __fadd_rn(__fmul_rn(v, 2.0f), -1.0f);
        const float sgn = (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f);
        const float e = __fadd_rn(expf(__fmul_rn(fabsf(v), l1p)), -1.0f);
        const float num = __fmul_rn(sgn, e);
        out[i] = recip ? __fmul_rn(num, inv_mu) : __fdiv_rn(num, mu);
    }
}

__global__ void one_hot_kernel(const int64_t* __restrict__ q, int64_t n, int C, int64_t* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t v = q[i];
        int64_t* row = out + i * C;
        for (int c = lane; c < C; c += 32) row[c] = (v == c);
    }
}

// ---------------------------------------------------------------------------------------------
// statistics: min, max, mean, unbiased std in one pass (double accumulators), two tiny stages
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stats_partial_kernel(const float* __restrict__ x, int64_t n, int kind, int contrast,
                                                            float eps, StatAcc* __restrict__ part) {
    StatAcc a{DBL_MAX, -DBL_MAX, 0.0, 0.0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v;
        if (kind == ACIDS_STATS_CABS_CONTRAST) {
            const float2 c = ldg_stream2(reinterpret_cast<const float2*>(x) + i);
            v = apply_contrast(sqrtf(c.x * c.x + c.y * c.y), contrast, eps);
        } else if (kind == ACIDS_STATS_ABS_CONTRAST) {
            v = apply_contrast(fabsf(__ldg(x + i)), contrast, eps);
        } else {
            v = __ldg(x + i);
        }
        const double d = (double)v;
        a.mn = fmin(a.mn, d);
        a.mx = fmax(a.mx, d);
        a.s += d;
        a.s2 += d * d;
    }
    a = stat_block_reduce(a);
    if (threadIdx.x == 0) part[blockIdx.x] = a;
}

__global__ void __launch_bounds__(256) stats_final_kernel(const StatAcc* __restrict__ part, int nparts, int64_t n,
                                                          double* __restrict__ out4) {
    StatAcc a{DBL_MAX, -DBL_MAX, 0.0, 0.0};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) stat_merge(a, part[i]);
    a = stat_block_reduce(a);
    if (threadIdx.x == 0) {
        const double mean = a.s / (double)n;
        double var = (a.s2 - a.s * a.s / (double)n) / (double)(n - 1);   // unbiased, like torch.std
        if (!(var > 0.0)) var = (n > 1) ? 0.0 : nan("");
        out4[0] = a.mn;
        out4[1] = a.mx;
        out4[2] = mean;
        out4[3] = sqrt(var);
    }
}

// ---------------------------------------------------------------------------------------------
// Mono mix and MidSide
// ---------------------------------------------------------------------------------------------
// grid = (chunks of a clip, clips), like midside_kernel: no division per sample, 16-byte accesses when the rows allow it
__global__ void __launch_bounds__(256) mono_mix_kernel(const float* __restrict__ x, int64_t B, int64_t L, float* __restrict__ out) {
    const bool vec = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const float* __restrict__ x0 = x + 2 * b * L;
        const float* __restrict__ x1 = x0 + L;
        float* __restrict__ y = out + b * L;
        // raw.py:39: x.sum(-2) / 2  (x / 2 == x * 0.5 exactly)
        if (vec) {
            for (int64_t j = j0; j < (L >> 2); j += step) {
                const float4 l = __ldg(reinterpret_cast<const float4*>(x0) + j), r = __ldg(reinterpret_cast<const float4*>(x1) + j);
                reinterpret_cast<float4*>(y)[j] = make_float4(__fmul_rn(__fadd_rn(l.x, r.x), 0.5f), __fmul_rn(__fadd_rn(l.y, r.y), 0.5f),
                                                              __fmul_rn(__fadd_rn(l.z, r.z), 0.5f), __fmul_rn(__fadd_rn(l.w, r.w), 0.5f));
            }
        } else {
            for (int64_t j = j0; j < L; j += step) y[j] = __fmul_rn(__fadd_rn(__ldg(x0 + j), __ldg(x1 + j)), 0.5f);
        }
    }
}

__device__ __forceinline__ void midside_pair(float a, float c, int pad_mid, int inverse, float& o0, float& o1) {
    const float rt2 = 1.41421356237309504880f;   // float32(math.sqrt(2))
    if (!inverse) {        // raw.py:155-160; x / 2 == x * 0.5 exactly, the division by sqrt(2) stays an IEEE division
        o0 = __fmul_rn(__fadd_rn(a, c), 0.5f);
        o1 = __fmul_rn(__fadd_rn(a, -c), 0.5f);
        if (pad_mid) o0 = __fdiv_rn(o0, rt2);
    } else {               // raw.py:172-178
        const float mid = pad_mid ? __fmul_rn(a, rt2) : a;
        o0 = __fadd_rn(mid, c);
        o1 = __fadd_rn(mid, -c);
    }
}

// grid = (chunks of a clip, clips): no division per element; 16-byte accesses when the clip rows allow it
__global__ void __launch_bounds__(256) midside_kernel(const float* __restrict__ x, int64_t B, int64_t L, int pad_mid, int inverse,
                                                      float* __restrict__ out) {
    const bool vec = (L & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t j0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, step = (int64_t)gridDim.x * blockDim.x;
    for (int64_t b = blockIdx.y; b < B; b += gridDim.y) {
        const float* __restrict__ x0 = x + 2 * b * L;
        const float* __restrict__ x1 = x0 + L;
        float* __restrict__ y0 = out + 2 * b * L;
        float* __restrict__ y1 = y0 + L;
        if (vec) {
            for (int64_t j = j0; j < (L >> 2); j += step) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(x0) + j), c = __ldg(reinterpret_cast<const float4*>(x1) + j);
                float4 p, q;
                midside_pair(a.x, c.x, pad_mid, inverse, p.x, q.x);
                midside_pair(a.y, c.y, pad_mid, inverse, p.y, q.y);
                midside_pair(a.z, c.z, pad_mid, inverse, p.z, q.z);
                midside_pair(a.w, c.w, pad_mid, inverse, p.w, q.w);
                reinterpret_cast<float4*>(y0)[j] = p;
                reinterpret_cast<float4*>(y1)[j] = q;
            }
        } else {
            for (int64_t j = j0; j < L; j += step) {
                float p, q;
                midside_pair(__ldg(x0 + j), __ldg(x1 + j), pad_mid, inverse, p, q);
                y0[j] = p;
                y1[j] = q;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// MFCC tail: per-group max of the mel power, dB with top_db floor, ortho DCT-II.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_max_kernel(const float* __restrict__ mel, int64_t per_group,
                                                        float* __restrict__ gmax) {
    // grid = (chunks, groups): every CTA reduces a strided share of its group and merges with one atomic.  Only
    // max(x, 0) matters (the dB conversion clamps at 1e-10 anyway), and for non-negative floats the integer order of
    // the bit patterns is the float order: an integer atomicMax on a zero-initialised word is exact.
    // (log10 is monotone, so the dB max follows from the max of the mel power.)
    const float* src = mel + (int64_t)blockIdx.y * per_group;
    float m = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    const int64_t n4 = ((reinterpret_cast<uintptr_t>(src) & 15) == 0) ? (per_group & ~(int64_t)3) : 0;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + i));
        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (int64_t i = n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_group; i += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, __ldg(src + i));
    __shared__ float sh[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) m = fmaxf(m, sh[i]);
        atomicMax(reinterpret_cast<int*>(gmax + blockIdx.y), __float_as_int(m));
    }
}

// block: 32 frames x 8 coefficient lanes; dB tile [n_mels][32] and DCT matrix in shared memory
__global__ void __launch_bounds__(256) mfcc_dct_kernel(const float* __restrict__ mel, int64_t B, int n_mels,
                                                       int64_t n_frames, const float* __restrict__ dct, int n_mfcc,
                                                       float top_db, const float* __restrict__ gmax,
                                                       int64_t clips_per_group, float* __restrict__ out) {
    extern __shared__ float smem[];
    float* db = smem;                         // [n_mels][33]
    float* dm = smem + (size_t)n_mels * 33;   // [n_mels][n_mfcc]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_mels * n_mfcc; i += 256) dm[i] = __ldg(dct + i);
    const int64_t tiles_per_clip = (n_frames + 31) / 32;
    for (int64_t tile = blockIdx.x; tile < B * tiles_per_clip; tile += gridDim.x) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * 32;
        float floor_db = -FLT_MAX;
        if (top_db >= 0.f) {
            const float gm = __ldg(gmax + b / clips_per_group);
            floor_db = 10.0f * log10f(fmaxf(gm, 1e-10f)) - top_db;
        }
        __syncthreads();
        const float* src = mel + b * (int64_t)n_mels * n_frames;
        for (int m = ty; m < n_mels; m += 8) {
            const int64_t t = t0 + tx;
            float v = 0.f;
            if (t < n_frames) {
                v = 10.0f * log10f(fmaxf(__ldg(src + (int64_t)m * n_frames + t), 1e-10f));   // functional.py:390-391 (db_multiplier = 0)
                v = fmaxf(v, floor_db);
            }
            db[m * 33 + tx] = v;
        }
        __syncthreads();
        const int64_t t = t0 + tx;
        for (int k = ty; k < n_mfcc; k += 8) {
            float acc = 0.f;
            for (int m = 0; m < n_mels; ++m) acc = fmaf(db[m * 33 + tx], dm[m * n_mfcc + k], acc);
            if (t < n_frames) out[(b * n_mfcc + k) * n_frames + t] = acc;
        }
    }
}

// Register-tiled variant for n_mfcc <= 64: a CTA takes 64 frames; thread (lane, warp) owns frames lane and lane + 32
// and the CPT consecutive coefficients warp * CPT ... : per mel band 2 conflict-free loads of the dB tile, CPT broadcast
// loads of the DCT row and 2 CPT FMAs.  dB tile [n_mels][65], DCT matrix [n_mels][8 CPT] (zero padded) in shared memory.
template <int CPT>
__global__ void __launch_bounds__(256) mfcc_dct_tiled_kernel(const float* __restrict__ mel, int64_t B, int n_mels,
                                                             int64_t n_frames, const float* __restrict__ dct, int n_mfcc,
                                                             float top_db, const float* __restrict__ gmax,
                                                             int64_t clips_per_group, float* __restrict__ out) {
    extern __shared__ float smem[];
    constexpr int KP = 8 * CPT;
    float* db = smem;                         // [n_mels][65]
    float* dm = smem + (size_t)n_mels * 65;   // [n_mels][KP]
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < n_mels * KP; i += 256) {
        const int m = i / KP, k = i - m * KP;
        dm[i] = k < n_mfcc ? __ldg(dct + (size_t)m * n_mfcc + k) : 0.f;
    }
    const int64_t tiles_per_clip = (n_frames + 63) / 64;
    for (int64_t tile = blockIdx.x; tile < B * tiles_per_clip; tile += gridDim.x) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * 64;
        float floor_db = -FLT_MAX;
        if (top_db >= 0.f) {
            const float gm = __ldg(gmax + b / clips_per_group);
            floor_db = 10.0f * log10f(fmaxf(gm, 1e-10f)) - top_db;
        }
        __syncthreads();
        const float* src = mel + b * (int64_t)n_mels * n_frames;
        for (int i = threadIdx.x; i < n_mels * 64; i += 256) {
            const int m = i >> 6, c = i & 63;
            const int64_t t = t0 + c;
            float v = 0.f;
            if (t < n_frames) {
                v = 10.0f * log10f(fmaxf(__ldg(src + (int64_t)m * n_frames + t), 1e-10f));   // functional.py:390-391 (db_multiplier = 0)
                v = fmaxf(v, floor_db);
            }
            db[m * 65 + c] = v;
        }
        __syncthreads();
        float a0[CPT], a1[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) a0[j] = a1[j] = 0.f;
        const float* dr = dm + ty * CPT;
#pragma unroll 4
        for (int m = 0; m < n_mels; ++m) {
            const float x0 = db[m * 65 + tx], x1 = db[m * 65 + tx + 32];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                const float w = dr[m * KP + j];
                a0[j] = fmaf(x0, w, a0[j]);
                a1[j] = fmaf(x1, w, a1[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            const int k = ty * CPT + j;
            if (k < n_mfcc) {
                float* o = out + (b * n_mfcc + k) * n_frames + t0;
                if (t0 + tx < n_frames) o[tx] = a0[j];
                if (t0 + tx + 32 < n_frames) o[tx + 32] = a1[j];
            }
        }
    }
}

static inline unsigned grid_for(int64_t n, int threads, int per_sm) {
    int64_t g = (n + threads - 1) / threads;
    const int64_t cap = (int64_t)num_sms() * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (unsigned)g;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int acids_mulaw_encode(const float* x, int64_t outer, int64_t inner, int channels, float log1p_mu,
                                  int reciprocal_divide, int one_hot, int64_t* out, void* stream) {
    ACIDS_REQUIRE(x && out, ACIDS_EINVAL, "mulaw_encode: NULL pointer");
    ACIDS_REQUIRE(outer >= 0 && inner >= 0 && channels >= 2, ACIDS_EINVAL, "mulaw_encode: bad sizes");
    ACIDS_REQUIRE(one_hot >= 0 && one_hot <= 2, ACIDS_EINVAL, "mulaw_encode: unknown one-hot layout %d", one_hot);
    const int64_t n = outer * inner;
    if (n == 0) return ACIDS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float mu = (float)(channels - 1.0);
    if (one_hot == ACIDS_ONEHOT_NONE)
        mulaw_encode_kernel<<<grid_for((n + 3) / 4, 256, 8), 256, 0, st>>>(x, n, mu, log1p_mu, reciprocal_divide, out);
    else if (one_hot == ACIDS_ONEHOT_CATEGORICAL)
        mulaw_onehot_categorical_kernel<<<grid_for(n * 32, 256, 16), 256, 0, st>>>(x, n, channels, mu, log1p_mu, reciprocal_divide, out);
    else
        mulaw_onehot_channel_kernel<<<grid_for(n, 256, 16), 256, 0, st>>>(x, outer, inner, channels, mu, log1p_mu, reciprocal_divide, out);
    ACIDS_CHECK_LAUNCH("mulaw_encode");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_mulaw_decode(const int64_t* q, int64_t n, int channels, float log1p_mu, int reciprocal_divide,
                                            float* out, void* stream) {
    ACIDS_REQUIRE(q && out, ACIDS_EINVAL, "mulaw_decode: NULL pointer");
    ACIDS_REQUIRE(n >= 0 && channels >= 2, ACIDS_EINVAL, "mulaw_decode: bad sizes");
    if (n == 0) return ACIDS_OK;
    mulaw_decode_kernel<<<grid_for(n, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n, (float)(channels - 1.0), log1p_mu, reciprocal_divide, out);
    ACIDS_CHECK_LAUNCH("mulaw_decode");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_one_hot(const int64_t* q, int64_t n, int n_classes, int64_t* out, void* stream) {
    ACIDS_REQUIRE(q && out, ACIDS_EINVAL, "one_hot: NULL pointer");
    ACIDS_REQUIRE(n >= 0 && n_classes >= 1, ACIDS_EINVAL, "one_hot: bad sizes");
    if (n == 0) return ACIDS_OK;
    one_hot_kernel<<<grid_for(n * 32, 256, 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, n, n_classes, out);
    ACIDS_CHECK_LAUNCH("one_hot");
    return ACIDS_OK;
}

namespace acids {
int launch_stats_final(const StatAcc* part, int nparts, int64_t n, double* out4, cudaStream_t st) {
    stats_final_kernel<<<1, 256, 0, st>>>(part, nparts, n, out4);
    ACIDS_CHECK_LAUNCH("stats");
    return ACIDS_OK;
}
}  // namespace acids

extern "C" ACIDS_API int64_t acids_stats_scratch_bytes(void) { return (int64_t)kStatsBlocks * (int64_t)sizeof(StatAcc); }

extern "C" ACIDS_API int acids_stats(const float* x, int64_t n, int kind, int contrast, float eps, void* scratch, double* out4,
                           void* stream) {
    ACIDS_REQUIRE(x && scratch && out4, ACIDS_EINVAL, "stats: NULL pointer");
    ACIDS_REQUIRE(n >= 1, ACIDS_EINVAL, "stats: empty input");
    ACIDS_REQUIRE(kind == ACIDS_STATS_REAL || kind == ACIDS_STATS_CABS_CONTRAST || kind == ACIDS_STATS_ABS_CONTRAST, ACIDS_EINVAL, "stats: unknown kind %d", kind);
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int64_t blocks = (n + 255) / 256;
    if (blocks > kStatsBlocks) blocks = kStatsBlocks;
    stats_partial_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, kind, contrast, eps, static_cast<StatAcc*>(scratch));
    stats_final_kernel<<<1, 256, 0, st>>>(static_cast<const StatAcc*>(scratch), (int)blocks, n, out4);
    ACIDS_CHECK_LAUNCH("stats");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_mono_mix(const float* x, int64_t B, int64_t L, float* out, void* stream) {
    ACIDS_REQUIRE(x && out, ACIDS_EINVAL, "mono_mix: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && L >= 0, ACIDS_EINVAL, "mono_mix: bad sizes");
    if (B * L == 0) return ACIDS_OK;
    const int64_t per_clip = ((L & 3) == 0 ? L / 4 : L);
    int64_t gx = (per_clip + 255) / 256;
    if (gx > 4096) gx = 4096;
    const int64_t gy = B < 65535 ? B : 65535;
    mono_mix_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, L, out);
    ACIDS_CHECK_LAUNCH("mono_mix");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_midside(const float* x, int64_t B, int64_t L, int pad_mid, int inverse, float* out, void* stream) {
    ACIDS_REQUIRE(x && out, ACIDS_EINVAL, "midside: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && L >= 0, ACIDS_EINVAL, "midside: bad sizes");
    if (B * L == 0) return ACIDS_OK;
    const int64_t per_clip = ((L & 3) == 0 ? L / 4 : L);
    int64_t gx = (per_clip + 255) / 256;
    if (gx > 4096) gx = 4096;
    if (gx < 1) gx = 1;
    const int64_t gy = B < 65535 ? B : 65535;
    midside_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, B, L, pad_mid, inverse, out);
    ACIDS_CHECK_LAUNCH("midside");
    return ACIDS_OK;
}

static int launch_group_max(const float* mel, int64_t B, int n_mels, int64_t n_frames, float top_db, int64_t clips_per_group,
                            float* group_max, cudaStream_t st) {
    if (top_db < 0.f) return ACIDS_OK;
    const int64_t groups = B / clips_per_group, per_group = clips_per_group * n_mels * n_frames;
    ACIDS_REQUIRE(groups < 65536, ACIDS_EINVAL, "mfcc_dct: more than 65535 top_db groups");
    if (cudaMemsetAsync(group_max, 0, (size_t)groups * sizeof(float), st) != cudaSuccess) {
        set_error("mfcc_dct: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
        return ACIDS_ECUDA;
    }
    // enough CTAs to stream the whole tensor at HBM speed, split over the groups
    int64_t chunks = (per_group + 256 * 16 - 1) / (256 * 16);
    const int64_t cap = ((int64_t)num_sms() * 8 + groups - 1) / groups;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    group_max_kernel<<<dim3((unsigned)chunks, (unsigned)groups), 256, 0, st>>>(mel, per_group, group_max);
    ACIDS_CHECK_LAUNCH("mfcc group max");
    return ACIDS_OK;
}

int acids_mfcc_dct_tc_launch(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct, int n_mfcc, float top_db,
                             int64_t clips_per_group, const float* group_max, float* out, cudaStream_t st);   // mfcc_tc.cu

extern "C" ACIDS_API int acids_mfcc_dct(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct, int n_mfcc,
                              float top_db, int64_t clips_per_group, float* group_max, float* out, void* stream) {
    ACIDS_REQUIRE(mel && dct && out, ACIDS_EINVAL, "mfcc_dct: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_mels > 0 && n_frames > 0 && n_mfcc > 0, ACIDS_EINVAL, "mfcc_dct: bad sizes");
    ACIDS_REQUIRE(top_db < 0.f || (group_max && clips_per_group >= 1 && B % clips_per_group == 0), ACIDS_EINVAL,
                  "mfcc_dct: top_db needs group_max scratch and a group size dividing B");
    if (B == 0) return ACIDS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    {
        const int rc = launch_group_max(mel, B, n_mels, n_frames, top_db, clips_per_group, group_max, st);
        if (rc) return rc;
    }
    const int64_t gsize = clips_per_group > 0 ? clips_per_group : 1;
    if (n_mfcc <= 64) {
        const int cpt = (n_mfcc + 7) / 8;
        const size_t tsmem = ((size_t)n_mels * 65 + (size_t)n_mels * 8 * cpt) * sizeof(float);
        if (tsmem <= 200 * 1024) {
            void (*kern)(const float*, int64_t, int, int64_t, const float*, int, float, const float*, int64_t, float*) = nullptr;
            switch (cpt) {
                case 1: kern = mfcc_dct_tiled_kernel<1>; break;
                case 2: kern = mfcc_dct_tiled_kernel<2>; break;
                case 3: kern = mfcc_dct_tiled_kernel<3>; break;
                case 4: kern = mfcc_dct_tiled_kernel<4>; break;
                case 5: kern = mfcc_dct_tiled_kernel<5>; break;
                case 6: kern = mfcc_dct_tiled_kernel<6>; break;
                case 7: kern = mfcc_dct_tiled_kernel<7>; break;
                default: kern = mfcc_dct_tiled_kernel<8>; break;
            }
            ACIDS_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(tsmem > 48 * 1024 ? tsmem : 48 * 1024)) == cudaSuccess,
                          ACIDS_ECUDA, "mfcc_dct: cannot reserve %zu B of shared memory", tsmem);
            int64_t tgrid = B * ((n_frames + 63) / 64);
            const int64_t tcap = (int64_t)num_sms() * (tsmem <= 56 * 1024 ? 4 : (tsmem <= 112 * 1024 ? 2 : 1));
            if (tgrid > tcap) tgrid = tcap;
            kern<<<(unsigned)tgrid, 256, tsmem, st>>>(mel, B, n_mels, n_frames, dct, n_mfcc, top_db, group_max, gsize, out);
            ACIDS_CHECK_LAUNCH("mfcc_dct");
            return ACIDS_OK;
        }
    }
    const size_t smem = ((size_t)n_mels * 33 + (size_t)n_mels * n_mfcc) * sizeof(float);
    ACIDS_REQUIRE(smem <= 227 * 1024, ACIDS_ENOTSUP, "mfcc_dct: n_mels * (33 + n_mfcc) floats exceed shared memory");
    if (smem > 48 * 1024)
        ACIDS_REQUIRE(cudaFuncSetAttribute(mfcc_dct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                      ACIDS_ECUDA, "mfcc_dct: cannot reserve %zu B of shared memory", smem);
    const int64_t tiles = B * ((n_frames + 31) / 32);
    int64_t grid = tiles;
    const int64_t cap = (int64_t)num_sms() * 4;
    if (grid > cap) grid = cap;
    mfcc_dct_kernel<<<(unsigned)grid, 256, smem, st>>>(mel, B, n_mels, n_frames, dct, n_mfcc, top_db, group_max,
                                                        clips_per_group > 0 ? clips_per_group : 1, out);
    ACIDS_CHECK_LAUNCH("mfcc_dct");
    return ACIDS_OK;
}

// The same tail with the DCT on the tensor cores (tcgen05, 3xTF32; mfcc_tc.cu).  Shapes: n_mels a multiple of 8 up to
// 128, n_mfcc <= 48; other shapes return ACIDS_ENOTSUP (use acids_mfcc_dct).
extern "C" ACIDS_API int acids_mfcc_dct_tc(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct, int n_mfcc,
                                 float top_db, int64_t clips_per_group, float* group_max, float* out, void* stream) {
    ACIDS_REQUIRE(mel && dct && out, ACIDS_EINVAL, "mfcc_dct_tc: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_mels > 0 && n_frames > 0 && n_mfcc > 0, ACIDS_EINVAL, "mfcc_dct_tc: bad sizes");
    ACIDS_REQUIRE(n_mels % 8 == 0 && n_mels <= 128 && n_mfcc <= 48, ACIDS_ENOTSUP,
                  "mfcc_dct_tc: needs n_mels %% 8 == 0, n_mels <= 128, n_mfcc <= 48 (got %d, %d)", n_mels, n_mfcc);
    ACIDS_REQUIRE(top_db < 0.f || (group_max && clips_per_group >= 1 && B % clips_per_group == 0), ACIDS_EINVAL,
                  "mfcc_dct_tc: top_db needs group_max scratch and a group size dividing B");
    if (B == 0) return ACIDS_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int rc = launch_group_max(mel, B, n_mels, n_frames, top_db, clips_per_group, group_max, st);
    if (rc) return rc;
    return acids_mfcc_dct_tc_launch(mel, B, n_mels, n_frames, dct, n_mfcc, top_db, clips_per_group, group_max, out, st);
}
