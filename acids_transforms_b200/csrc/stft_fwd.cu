// stft_fwd.cu — fused framing + reflect padding + window + real FFT (+ magnitude / mel / contrast /
// normalise epilogue).  Rows A3, A6, A9, A17 of SURVEY.md §8(a).
//
// Work decomposition: a "unit" is G consecutive frames of one clip (G = THREADS / T frame groups
// per CTA, T threads per frame from plans.cuh).  CTA c of a persistent grid owns the contiguous
// unit range [c*U/grid, (c+1)*U/grid): consecutive units share (n_fft - hop)/n_fft of their input,
// which is then served by L1 instead of L2/HBM.  Each sample is fetched from HBM once.
//
// HBM traffic per frame (ideal = achieved): hop*4 B in, (n_fft/2+1)*8 B out (complex mode) or
// n_cols*4 B out (fused mode).  Everything else lives in registers and 4.3 KB of shared memory
// per frame group (+ the banded mel matrix, staged once per CTA).
#include "common.cuh"
#include "plans.cuh"

namespace acids {

enum { MODE_COMPLEX = 0, MODE_REAL = 1 };

struct FwdParams {
    const float* x;
    int64_t B, L, ldx;
    int hop, pad;
    int64_t n_frames;
    const float* window;
    float* out;
    int64_t out_clip_stride, out_row_stride, out_col_stride;   // MODE_REAL, in floats
    EpiParams ep;
    const float* offset_ptr;
    const float* scale_ptr;
    float power;     // MODE_REAL: value = |X|^power (1 -> magnitude, 2 -> power spectrum)
    int vec_ok;      // rows and frame starts are 8-byte aligned: float2 loads allowed
};

// launch shape per plan: small frame groups run 128-thread CTAs at 4 CTAs / SM (<= 128 registers)
#ifndef ACIDS_FWD_MINB_SMALL
#define ACIDS_FWD_MINB_SMALL 4
#endif
template <class P>
struct FwdCfg {
    static constexpr int THREADS = P::T <= 32 ? 128 : (P::T > 256 ? P::T : 256);
    static constexpr int MINB = P::T <= 32 ? ACIDS_FWD_MINB_SMALL : (P::T <= 256 ? 2 : 1);
    static constexpr int G = THREADS / P::T;
};

// |X|^power from the squared magnitude; pmode: 1 -> magnitude, 2 -> power spectrum, 0 -> general exponent
__device__ __forceinline__ float pow_value(float re, float im, int pmode, float power) {
    const float p2 = re * re + im * im;
    if (pmode == 1) return fast_sqrt(p2);
    if (pmode == 2) return p2;
    return powf(fast_sqrt(p2), power);
}

template <class P, int MODE>
__global__ void __launch_bounds__(FwdCfg<P>::THREADS, FwdCfg<P>::MINB) stft_fwd_kernel(const FwdParams p) {
    constexpr int THREADS = FwdCfg<P>::THREADS;
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V, G = THREADS / T;
    constexpr int R0 = P::radix(0), B0 = P::bpt(0), NB0 = P::nb(0);
    using FFT = FrameFFT<P, false>;
    using PR = typename FFT::PR;
    constexpr int RP = PR::R, NBP = PR::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    cf* s = reinterpret_cast<cf*>(smem_raw) + (size_t)g * P::SMEM_CF;
    auto gsync = [&]() { group_sync<T, THREADS>(g); };

    // ---- frame-invariant registers: twiddles and this thread's window taps (pre-scaled by 1/2,
    //      the factor of the even/odd split) ----
    FFT fft;
    fft.init(tid);
    // analysis window as (w[2n], w[2n+1]) / 2 pairs in shared memory: the pass-0 operands of a thread are
    // consecutive float2 across the group, so the reads are conflict free and cost no registers
    float2* swin = reinterpret_cast<float2*>(smem_raw + (size_t)G * P::SMEM_CF * sizeof(cf));
    for (int n = threadIdx.x; n < M; n += THREADS)
        swin[n] = make_float2(0.5f * __ldg(p.window + 2 * n), 0.5f * __ldg(p.window + 2 * n + 1));
    EpiParams ep = p.ep;
    const int32_t* bmeta = nullptr;
    const float* bcoef = nullptr;
    if (MODE == MODE_REAL) {
        load_norm(p.offset_ptr, p.scale_ptr, ep);
        stage_band(ep, smem_raw + (size_t)G * P::SMEM_CF * sizeof(cf) + (size_t)M * sizeof(float2), bmeta, bcoef);
    }
    __syncthreads();

    const int pmode = p.power == 1.0f ? 1 : (p.power == 2.0f ? 2 : 0);
    const int64_t upc = (p.n_frames + G - 1) / G;        // units per clip
    const int64_t total = p.B * upc;
    const int64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
    // |X| rows of the current unit (MODE_REAL): their own region, so that the next unit's FFT never waits for
    // the slowest epilogue thread of this one
    constexpr int VSTR = (P::F + 3) & ~3;
    float* vrows = reinterpret_cast<float*>(smem_raw + (size_t)G * P::SMEM_CF * sizeof(cf) + (size_t)M * sizeof(float2) +
                                            (size_t)p.ep.band_bytes_meta + p.ep.band_bytes_coef);

    // raw (un-windowed) samples of frame (b, t) in pass-0 operand order; edge frames reflect (torch.stft
    // center=True, pad_mode="reflect")
    auto fetch = [&](cf* v, int64_t b, int64_t t) {
        const bool valid = t < p.n_frames;
        const int64_t s0 = t * p.hop - p.pad;
        const float* __restrict__ xb = p.x + b * p.ldx;
        if (valid && p.vec_ok && s0 >= 0 && s0 + N <= p.L) {
#pragma unroll
            for (int b0 = 0; b0 < B0; ++b0) {
                const float2* __restrict__ src = reinterpret_cast<const float2*>(xb + s0) + (tid + T * b0);
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    const float2 a = __ldg(src + r * NB0);
                    v[b0 * R0 + r] = mk(a.x, a.y);
                }
            }
        } else if (valid) {
#pragma unroll
            for (int b0 = 0; b0 < B0; ++b0)
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    const int n = fft.template in_index<0>(b0, r);
                    float e[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int64_t i = s0 + 2 * n + h;
                        if (i < 0) i = -i;
                        if (i >= p.L) i = 2 * (p.L - 1) - i;
                        i = i < 0 ? 0 : (i >= p.L ? p.L - 1 : i);
                        e[h] = __ldg(xb + i);
                    }
                    v[b0 * R0 + r] = mk(e[0], e[1]);
                }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = mk(0.f, 0.f);
        }
    };

    int64_t b = u0 / upc;                   // clip and unit-in-clip advance incrementally: no division per frame
    int64_t uc = u0 - b * upc;
    cf v[V];
    if (u0 < u1) fetch(v, b, uc * G + g);
    for (int64_t u = u0; u < u1; ++u) {
        const int64_t t = uc * G + g;
        const bool valid = t < p.n_frames;
        const int64_t cur_b = b, cur_uc = uc;
        if (++uc == upc) {
            uc = 0;
            ++b;
        }

        // ---- window (pass-0 operand order) ----
#pragma unroll
        for (int b0 = 0; b0 < B0; ++b0) {
            const float2* __restrict__ wv = swin + (tid + T * b0);
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const float2 w = wv[r * NB0];
                v[b0 * R0 + r] = mk(v[b0 * R0 + r].x * w.x, v[b0 * R0 + r].y * w.y);
            }
        }

        // ---- passes ----
        fft.template butterflies<0>(v);
        gsync();   // the previous frame's readers of s are done
        fft.template store<0>(v, s);
        gsync();
        fft.template load<1>(v, s);
        fft.template butterflies<1>(v);
        if constexpr (P::NP > 2) {
            gsync();
            fft.template store<1>(v, s);
            gsync();
            fft.template load<2>(v, s);
            fft.template butterflies<2>(v);
        }
        if constexpr (P::NP > 3) {
            gsync();
            fft.template store<2>(v, s);
            gsync();
            fft.template load<3>(v, s);
            fft.template butterflies<3>(v);
        }

        // ---- untangle in registers ----
        cf o1[V / 2], o2[V / 2], ex;
        fft.untangle_fwd(v, o1, o2, ex);

        if (MODE == MODE_COMPLEX) {
            if (valid) {
                float2* __restrict__ row = reinterpret_cast<float2*>(p.out) + (cur_b * p.n_frames + t) * (int64_t)P::F;
#pragma unroll
                for (int c = 0; c < PR::PC; ++c) {
                    // bins k = base + q*NB and M - k: two per-thread bases, compile-time offsets
                    float2* lo = row + PR::klo(tid, c);
                    float2* hi = row + PR::khi(tid, c);
                    float2* mlo = row + (M - PR::klo(tid, c));
                    float2* mhi = row + (M - PR::khi(tid, c));
#pragma unroll
                    for (int q = 0; q < RP; ++q) {
                        stg_stream2((q < RP / 2 ? lo : hi) + q * NBP, o1[c * RP + q].x, o1[c * RP + q].y);
                        stg_stream2((q < RP / 2 ? mlo : mhi) - q * NBP, o2[c * RP + q].x, o2[c * RP + q].y);
                    }
                }
                if (tid == 0) stg_stream2(row + M / 2, ex.x, ex.y);
            }
            if (u + 1 < u1) fetch(v, b, uc * G + g);
        } else {
            float* __restrict__ val = vrows + g * VSTR;
            __syncthreads();   // every thread is past the previous unit's epilogue: the rows may be overwritten
#pragma unroll
            for (int c = 0; c < PR::PC; ++c) {
                float* lo = val + PR::klo(tid, c);
                float* hi = val + PR::khi(tid, c);
                float* mlo = val + (M - PR::klo(tid, c));
                float* mhi = val + (M - PR::khi(tid, c));
#pragma unroll
                for (int q = 0; q < RP; ++q) {
                    (q < RP / 2 ? lo : hi)[q * NBP] = pow_value(o1[c * RP + q].x, o1[c * RP + q].y, pmode, p.power);
                    (q < RP / 2 ? mlo : mhi)[-q * NBP] = pow_value(o2[c * RP + q].x, o2[c * RP + q].y, pmode, p.power);
                }
            }
            if (tid == 0) val[M / 2] = pow_value(ex.x, ex.y, pmode, p.power);
            // v, o1, o2 are dead: start fetching the next frame's samples, they land during the epilogue
            if (u + 1 < u1) fetch(v, b, uc * G + g);
            // the G rows of this unit are projected together by the whole CTA: per-column band metadata and
            // coefficients are fetched once per row chunk
            __syncthreads();
            const int64_t t0 = cur_uc * G;
            const int n_valid = (int)min((int64_t)G, p.n_frames - t0);
            float* out_row0 = p.out + cur_b * p.out_clip_stride + t0 * p.out_row_stride;
            epilogue_rows<THREADS, G>(vrows, VSTR, threadIdx.x, ep, bmeta, bcoef, out_row0, p.out_col_stride, p.out_row_stride, n_valid);
        }
    }
}

static const size_t kBandSmemBudget = 24 * 1024;

template <class P, int MODE>
static int launch_fwd(FwdParams p, cudaStream_t st) {
    constexpr int THREADS = FwdCfg<P>::THREADS;
    constexpr int G = FwdCfg<P>::G;
    const size_t band_bytes = (MODE == MODE_REAL) ? (size_t)p.ep.band_bytes_meta + p.ep.band_bytes_coef : 0;
    const size_t rows_bytes = (MODE == MODE_REAL) ? (size_t)G * ((P::F + 3) & ~3) * sizeof(float) : 0;
    const size_t smem = (size_t)G * P::SMEM_CF * sizeof(cf) + (size_t)P::M * sizeof(float2) + band_bytes + rows_bytes;
    auto kern = stft_fwd_kernel<P, MODE>;
    static size_t reserved = 0;
    static int ctas_per_sm = 0;
    if (smem > reserved || ctas_per_sm == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("stft_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
            return ACIDS_ECUDA;
        }
        reserved = smem;
        int nb = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, smem);
        ctas_per_sm = nb > 0 ? nb : 1;
    }
    const int64_t total = p.B * ((p.n_frames + G - 1) / G);
    if (total == 0) return ACIDS_OK;
    int64_t grid = (int64_t)num_sms() * ctas_per_sm;
    if (grid > total) grid = total;
    kern<<<(unsigned)grid, THREADS, smem, st>>>(p);
    ACIDS_CHECK_LAUNCH("stft_fwd");
    return ACIDS_OK;
}

template <int MODE>
static int dispatch_fwd(int n_fft, const FwdParams& p, cudaStream_t st) {
    switch (n_fft) {
        case 32: return launch_fwd<Fwd32, MODE>(p, st);
        case 64: return launch_fwd<Fwd64, MODE>(p, st);
        case 128: return launch_fwd<Fwd128, MODE>(p, st);
        case 256: return launch_fwd<Fwd256, MODE>(p, st);
        case 512: return launch_fwd<Fwd512, MODE>(p, st);
        case 1024: return launch_fwd<Fwd1024, MODE>(p, st);
        case 2048: return launch_fwd<Fwd2048, MODE>(p, st);
        case 4096: return launch_fwd<Fwd4096, MODE>(p, st);
        case 8192: return launch_fwd<Fwd8192, MODE>(p, st);
        case 16384: return launch_fwd<Fwd16384, MODE>(p, st);
        default:
            set_error("n_fft=%d is not supported (power of two in [32, 16384])", n_fft);
            return ACIDS_ENOTSUP;
    }
}

static int fill_common(FwdParams& p, const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                       int n_fft, int hop, int center, int64_t n_frames) {
    ACIDS_REQUIRE(x && window, ACIDS_EINVAL, "stft: NULL input or window");
    ACIDS_REQUIRE(B >= 0 && L > 0 && ldx >= L && hop > 0 && n_frames >= 0, ACIDS_EINVAL,
                  "stft: bad sizes B=%lld L=%lld ldx=%lld hop=%d frames=%lld", (long long)B, (long long)L,
                  (long long)ldx, hop, (long long)n_frames);
    const int pad = center ? n_fft / 2 : 0;
    if (center) {
        ACIDS_REQUIRE(L > pad, ACIDS_EINVAL, "stft: reflect padding %d needs an input longer than that (L=%lld)", pad, (long long)L);
        ACIDS_REQUIRE(n_frames <= 1 + L / hop, ACIDS_EINVAL, "stft: n_frames=%lld exceeds 1 + L / hop", (long long)n_frames);
    } else {
        ACIDS_REQUIRE(n_frames == 0 || (n_frames - 1) * hop + n_fft <= L, ACIDS_EINVAL, "stft: frames run past the input");
    }
    p.x = x; p.B = B; p.L = L; p.ldx = ldx; p.hop = hop; p.pad = pad; p.n_frames = n_frames; p.window = window;
    p.vec_ok = ((ldx & 1) == 0) && ((hop & 1) == 0) && ((pad & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    return ACIDS_OK;
}

int fill_epilogue(EpiParams& ep, acids_band band, int n_bins, int contrast, float eps, int drop_first, size_t smem_budget) {
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    ACIDS_REQUIRE(!band.meta || (band.coef && band.n_out > 0 && band.coef_len >= 0 && (band.coef_len & 31) == 0), ACIDS_EINVAL,
                  "malformed banded matrix (n_out=%d coef_len=%d)", band.n_out, band.coef_len);
    ep.meta = band.meta; ep.coef = band.coef;
    ep.n_cols = band.meta ? band.n_out : n_bins;
    ep.contrast = contrast; ep.eps = eps; ep.inv_scale = 1.f; ep.neg_off_scaled = 0.f; ep.drop_first = drop_first;
    const int64_t meta_ints = band.meta ? (int64_t)band.n_out + 2 * (((int64_t)band.n_out + 31) / 32) : 0;
    band_smem_plan(band, band.coef_len, meta_ints, smem_budget, ep);
    return ACIDS_OK;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int acids_stft_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                              int hop, int center, int64_t n_frames, float* out, void* stream) {
    FwdParams p{};
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, center, n_frames);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "stft_fwd: NULL output");
    p.out = out;
    return dispatch_fwd<MODE_COMPLEX>(n_fft, p, static_cast<cudaStream_t>(stream));
}

extern "C" ACIDS_API int acids_stft_mag_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                                  int hop, int center, int64_t n_frames, acids_band band, int contrast, float eps,
                                  const float* offset, const float* scale, int drop_first, float* out,
                                  int64_t out_clip_stride, int64_t out_row_stride, void* stream) {
    FwdParams p{};
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, center, n_frames);
    if (rc) return rc;
    rc = fill_epilogue(p.ep, band, n_fft / 2 + 1, contrast, eps, drop_first, kBandSmemBudget);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "stft_mag_fwd: NULL output");
    p.offset_ptr = offset; p.scale_ptr = scale;
    p.out = out; p.out_clip_stride = out_clip_stride; p.out_row_stride = out_row_stride; p.out_col_stride = 1;
    p.power = 1.0f;
    return dispatch_fwd<MODE_REAL>(n_fft, p, static_cast<cudaStream_t>(stream));
}

extern "C" ACIDS_API int acids_melspec_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                                 int hop, int64_t n_frames, acids_band mel, float power, const float* offset,
                                 const float* scale, float* out, void* stream) {
    FwdParams p{};
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, 1, n_frames);
    if (rc) return rc;
    ACIDS_REQUIRE(mel.meta && mel.coef && mel.n_out > 0, ACIDS_EINVAL, "melspec_fwd: mel bank required");
    ACIDS_REQUIRE(power > 0.f, ACIDS_EINVAL, "melspec_fwd: power must be > 0");
    rc = fill_epilogue(p.ep, mel, n_fft / 2 + 1, ACIDS_CONTRAST_NONE, 0.f, 0, kBandSmemBudget);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "melspec_fwd: NULL output");
    p.offset_ptr = offset; p.scale_ptr = scale;
    // frequency-major output [B, n_mels, n_frames] like torchaudio (mel.py:70)
    p.out = out; p.out_clip_stride = (int64_t)mel.n_out * n_frames; p.out_row_stride = 1; p.out_col_stride = n_frames;
    p.power = power;
    return dispatch_fwd<MODE_REAL>(n_fft, p, static_cast<cudaStream_t>(stream));
}
