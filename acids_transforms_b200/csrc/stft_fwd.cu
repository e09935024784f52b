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
// per frame group.
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

template <class P, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) stft_fwd_kernel(const FwdParams p) {
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V, G = THREADS / T;
    constexpr int R0 = P::radix(0), B0 = P::bpt(0);
    using FFT = FrameFFT<P, false>;
    using PR = typename FFT::PR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    cf* s = reinterpret_cast<cf*>(smem_raw) + (size_t)g * P::SMEM_CF;
    auto gsync = [&]() { group_sync<T, THREADS>(g); };

    // ---- frame-invariant registers: twiddles and this thread's window taps (pre-scaled by 1/2,
    //      the factor of the even/odd split) ----
    FFT fft;
    fft.init(tid);
    float2 win[V];
#pragma unroll
    for (int b = 0; b < B0; ++b)
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const int n = fft.template in_index<0>(b, r);
            win[b * R0 + r] = make_float2(0.5f * __ldg(p.window + 2 * n), 0.5f * __ldg(p.window + 2 * n + 1));
        }
    EpiParams ep = p.ep;
    if (MODE == MODE_REAL) load_norm(p.offset_ptr, p.scale_ptr, ep.offset, ep.inv_scale);

    const int64_t upc = (p.n_frames + G - 1) / G;        // units per clip
    const int64_t total = p.B * upc;
    const int64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;

    for (int64_t u = u0; u < u1; ++u) {
        const int64_t b = u / upc;
        const int64_t t = (u - b * upc) * G + g;
        const bool valid = t < p.n_frames;
        const int64_t s0 = t * p.hop - p.pad;
        const float* __restrict__ xb = p.x + b * p.ldx;

        // ---- load + window (pass-0 operand order) ----
        cf v[V];
        if (valid && p.vec_ok && s0 >= 0 && s0 + N <= p.L) {
            const float2* __restrict__ src = reinterpret_cast<const float2*>(xb + s0);
#pragma unroll
            for (int b0 = 0; b0 < B0; ++b0)
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    const float2 a = __ldg(src + fft.template in_index<0>(b0, r));
                    v[b0 * R0 + r] = mk(a.x * win[b0 * R0 + r].x, a.y * win[b0 * R0 + r].y);
                }
        } else if (valid) {
            // edge frame: reflect padding (torch.stft center=True, pad_mode="reflect")
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
                    v[b0 * R0 + r] = mk(e[0] * win[b0 * R0 + r].x, e[1] * win[b0 * R0 + r].y);
                }
        } else {
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = mk(0.f, 0.f);
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
                cf* __restrict__ row = reinterpret_cast<cf*>(p.out) + (b * p.n_frames + t) * (int64_t)P::F;
#pragma unroll
                for (int c = 0; c < PR::PC; ++c)
#pragma unroll
                    for (int q = 0; q < PR::R; ++q) {
                        const int k = PR::k1(tid, c, q);
                        stg_stream2(reinterpret_cast<float2*>(row + k), o1[c * PR::R + q].x, o1[c * PR::R + q].y);
                        stg_stream2(reinterpret_cast<float2*>(row + (M - k)), o2[c * PR::R + q].x, o2[c * PR::R + q].y);
                    }
                if (tid == 0) stg_stream2(reinterpret_cast<float2*>(row + M / 2), ex.x, ex.y);
            }
        } else {
            float* __restrict__ val = reinterpret_cast<float*>(s);
            gsync();   // every thread has finished reading s for the last pass
#pragma unroll
            for (int c = 0; c < PR::PC; ++c)
#pragma unroll
                for (int q = 0; q < PR::R; ++q) {
                    const int k = PR::k1(tid, c, q);
                    const cf a = o1[c * PR::R + q], d = o2[c * PR::R + q];
                    float pa = a.x * a.x + a.y * a.y, pd = d.x * d.x + d.y * d.y;
                    if (p.power == 1.0f) { pa = sqrtf(pa); pd = sqrtf(pd); }
                    else if (p.power != 2.0f) { pa = powf(sqrtf(pa), p.power); pd = powf(sqrtf(pd), p.power); }
                    val[k] = pa;
                    val[M - k] = pd;
                }
            if (tid == 0) {
                float pe = ex.x * ex.x + ex.y * ex.y;
                if (p.power == 1.0f) pe = sqrtf(pe);
                else if (p.power != 2.0f) pe = powf(sqrtf(pe), p.power);
                val[M / 2] = pe;
            }
            gsync();
            float* out_row = p.out + b * p.out_clip_stride + t * p.out_row_stride;
            epilogue_row<T>(val, tid, ep, out_row, p.out_col_stride, valid);
        }
    }
}

template <class P, int MODE>
static int launch_fwd(const FwdParams& p, cudaStream_t st) {
    constexpr int THREADS = P::T > 256 ? P::T : 256;
    constexpr int G = THREADS / P::T;
    constexpr size_t smem = (size_t)G * P::SMEM_CF * sizeof(cf);
    auto kern = stft_fwd_kernel<P, MODE, THREADS>;
    static int ctas_per_sm = 0;
    if (ctas_per_sm == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("stft_fwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
            return ACIDS_ECUDA;
        }
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

static int fill_epilogue(FwdParams& p, acids_band band, int n_bins, int contrast, float eps, const float* offset,
                         const float* scale, int drop_first) {
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    p.ep.meta = band.meta; p.ep.coef = band.coef;
    p.ep.n_cols = band.meta ? band.n_out : n_bins;
    ACIDS_REQUIRE(!band.meta || (band.coef && band.n_out > 0), ACIDS_EINVAL, "banded matrix without coefficients");
    p.ep.contrast = contrast; p.ep.eps = eps; p.ep.offset = 0.f; p.ep.inv_scale = 1.f; p.ep.drop_first = drop_first;
    p.offset_ptr = offset; p.scale_ptr = scale;
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
    rc = fill_epilogue(p, band, n_fft / 2 + 1, contrast, eps, offset, scale, drop_first);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "stft_mag_fwd: NULL output");
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
    rc = fill_epilogue(p, mel, n_fft / 2 + 1, ACIDS_CONTRAST_NONE, 0.f, offset, scale, 0);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "melspec_fwd: NULL output");
    // frequency-major output [B, n_mels, n_frames] like torchaudio (mel.py:70)
    p.out = out; p.out_clip_stride = (int64_t)mel.n_out * n_frames; p.out_row_stride = 1; p.out_col_stride = n_frames;
    p.power = power;
    return dispatch_fwd<MODE_REAL>(n_fft, p, static_cast<cudaStream_t>(stream));
}
