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
#include "stft_fwd_kernel.cuh"

namespace acids {

static int dispatch_fwd(int variant, int n_fft, const FwdParams& p, cudaStream_t st) {
    switch (n_fft) {
        case 32: return launch_fwd_plan_32(variant, p, st);
        case 64: return launch_fwd_plan_64(variant, p, st);
        case 128: return launch_fwd_plan_128(variant, p, st);
        case 256: return launch_fwd_plan_256(variant, p, st);
        case 512: return launch_fwd_plan_512(variant, p, st);
        case 1024: return launch_fwd_plan_1024(variant, p, st);
        case 2048: return launch_fwd_plan_2048(variant, p, st);
        case 4096: return launch_fwd_plan_4096(variant, p, st);
        case 8192: return launch_fwd_plan_8192(variant, p, st);
        case 16384: return launch_fwd_plan_16384(variant, p, st);
        default:
            set_error("n_fft=%d is not supported (power of two in [32, 16384])", n_fft);
            return ACIDS_ENOTSUP;
    }
}

static int fill_common(FwdParams& p, const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                       int n_fft, int hop, int center, int64_t n_frames) {
    ACIDS_REQUIRE(x && window, ACIDS_EINVAL, "stft: NULL input or window");
    ACIDS_REQUIRE(L < ((int64_t)1 << 31) - 2 * n_fft, ACIDS_ENOTSUP, "stft: clips of 2^31 samples or more are not supported (L=%lld)", (long long)L);
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
    // bit 0: every frame start is 8-byte aligned, bit 1: 16-byte aligned (which one a plan needs: FwdCfg::VW)
    const bool a2 = ((ldx & 1) == 0) && ((hop & 1) == 0) && ((pad & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    const bool a4 = ((ldx & 3) == 0) && ((hop & 3) == 0) && ((pad & 3) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    p.vec_ok = (a2 ? 1 : 0) | (a4 ? 2 : 0);
    return ACIDS_OK;
}

int fill_epilogue(EpiParams& ep, acids_band band, int n_bins, int contrast, float eps, int drop_first, size_t smem_budget) {
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    ACIDS_REQUIRE(!band.meta || (band.coef && band.n_out > 0 && band.coef_len >= 0 && (band.coef_len & 31) == 0), ACIDS_EINVAL,
                  "malformed banded matrix (n_out=%d coef_len=%d)", band.n_out, band.coef_len);
    // the epilogue indexes the |X| row at start[m] + u with no bound of its own: the matrix must have been built for
    // exactly this row length (the reference's matmul raises for STFT(512) + Magnitude(n_fft=1024), spectral_repr.py:219)
    ACIDS_REQUIRE(!band.meta || band.n_in == n_bins, ACIDS_EINVAL,
                  "mat1 and mat2 shapes cannot be multiplied (rows of %d values and a %dx%d matrix)", n_bins, band.n_in, band.n_out);
    ep.meta = band.meta; ep.coef = band.coef;
    ep.n_cols = band.meta ? band.n_out : n_bins;
    ep.contrast = contrast; ep.eps = eps; ep.drop_first = drop_first;
    band_smem_plan(band, band.coef_len, smem_budget, ep);
    return ACIDS_OK;
}

// which kernel variant serves a fused launch: where the band lives, and (mel-spectrogram) the exponent
static int real_variant(FwdParams& p, bool melspec, bool polar = false) {
    const bool band = p.ep.meta != nullptr;
    const bool in_smem = band && p.ep.band_bytes_meta > 0;
    p.band_smem_bytes = in_smem ? (int)round16((size_t)p.ep.band_bytes_meta + p.ep.band_bytes_coef) : 0;
    if (polar) return !band ? VAR_POLAR_NOBAND : (in_smem ? VAR_POLAR_SMEM : VAR_POLAR_GLOBAL);
    if (!melspec) return !band ? VAR_MAG_NOBAND : (in_smem ? VAR_MAG_SMEM : VAR_MAG_GLOBAL);
    if (p.power == 2.0f) return in_smem ? VAR_MEL_POWER_SMEM : VAR_MEL_POWER_GLOBAL;
    return in_smem ? VAR_MEL_ANY_SMEM : VAR_MEL_ANY_GLOBAL;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int acids_stft_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                              int hop, int center, int64_t n_frames, float* out, void* stream) {
    FwdParams p{};
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, center, n_frames);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "stft_fwd: NULL output");
    ACIDS_REQUIRE(n_frames * (n_fft / 2 + 1) < ((int64_t)1 << 30), ACIDS_ENOTSUP, "stft_fwd: more than 2^30 bins per clip");
    p.out = out;
    return dispatch_fwd(VAR_COMPLEX, n_fft, p, static_cast<cudaStream_t>(stream));
}

extern "C" ACIDS_API int acids_midside_stft_fwd(const float* x, int64_t B, int64_t L, const float* window, int n_fft, int hop,
                                      int64_t n_frames, int midside, float* out, void* stream) {
    FwdParams p{};
    ACIDS_REQUIRE(midside == 1 || midside == 2, ACIDS_EINVAL, "midside_stft_fwd: midside must be 1 (pad_mid=False) or 2 (pad_mid=True)");
    ACIDS_REQUIRE(B % 2 == 0, ACIDS_EINVAL, "midside_stft_fwd: the MidSide prologue takes contiguous stereo pairs (B=%lld must be even)", (long long)B);
    int rc = fill_common(p, x, B, L, L, window, n_fft, hop, 1, n_frames);
    if (rc) return rc;
    ACIDS_REQUIRE(out, ACIDS_EINVAL, "midside_stft_fwd: NULL output");
    ACIDS_REQUIRE(n_frames * (n_fft / 2 + 1) < ((int64_t)1 << 30), ACIDS_ENOTSUP, "midside_stft_fwd: more than 2^30 bins per clip");
    p.out = out;
    p.midside = midside;
    return dispatch_fwd(VAR_COMPLEX, n_fft, p, static_cast<cudaStream_t>(stream));
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
    ACIDS_REQUIRE(out_row_stride >= 0 && (n_frames + 4) * out_row_stride < ((int64_t)1 << 31), ACIDS_ENOTSUP,
                  "stft_mag_fwd: more than 2^31 output elements per clip");
    p.offset_ptr = offset; p.scale_ptr = scale;
    p.out = out; p.out_clip_stride = out_clip_stride; p.out_row_stride = out_row_stride; p.out_col_stride = 1;
    p.power = 1.0f;
    return dispatch_fwd(real_variant(p, false), n_fft, p, static_cast<cudaStream_t>(stream));
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
    ACIDS_REQUIRE((int64_t)mel.n_out * n_frames < ((int64_t)1 << 31), ACIDS_ENOTSUP, "melspec_fwd: more than 2^31 output elements per clip");
    p.offset_ptr = offset; p.scale_ptr = scale;
    // the tap-split epilogue (common.cuh: epilogue_split) re-packs the coefficients with every group's tap count rounded up to a
    // multiple of 4: reserve 3 x 32 floats per group on top of the coefficients when the bank lives in shared memory
    if (ACIDS_EPI_SPLIT && p.ep.band_bytes_meta > 0) {
        const int n_groups = (mel.n_out + 31) / 32;
        const size_t grown = (size_t)p.ep.band_bytes_coef + (size_t)n_groups * 96 * sizeof(float);
        if ((size_t)p.ep.band_bytes_meta + grown <= kBandSmemBudget + 4096 && grown / 4 < 65536) {
            p.ep.band_bytes_coef = (int)grown;
            p.ep.split_reserve = 1;
        }
    }
    // frequency-major output [B, n_mels, n_frames] like torchaudio (mel.py:70)
    p.out = out; p.out_clip_stride = (int64_t)mel.n_out * n_frames; p.out_row_stride = 1; p.out_col_stride = n_frames;
    p.power = power;
    return dispatch_fwd(real_variant(p, true), n_fft, p, static_cast<cudaStream_t>(stream));
}

extern "C" ACIDS_API int acids_stft_polar_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                                    int hop, int64_t n_frames, int midside, acids_band band, int contrast, float eps,
                                    const float* mag_offset, const float* mag_scale, int phase_mode, int if_method,
                                    int weighted, const float* ph_offset, const float* ph_scale, int drop_first,
                                    float* mag_out, int64_t mag_clip_stride, int64_t mag_row_stride, float* ph_out,
                                    int64_t ph_clip_stride, int64_t ph_row_stride, void* stream) {
    FwdParams p{};
    ACIDS_REQUIRE(midside >= 0 && midside <= 2, ACIDS_EINVAL, "stft_polar_fwd: midside must be 0, 1 or 2");
    ACIDS_REQUIRE(!midside || (B % 2 == 0 && ldx == L), ACIDS_EINVAL,
                  "stft_polar_fwd: the MidSide prologue takes contiguous stereo pairs (B=%lld even, ldx == L)", (long long)B);
    // the frame-difference scheme that needs no scan over the frames: raw phase, or forward-difference IF; the unwrapped
    // phase and the backward / central differences stay on acids_phase_fwd (the caller falls back)
    ACIDS_REQUIRE(phase_mode == ACIDS_PHASE_RAW || (phase_mode == ACIDS_PHASE_IF && if_method == ACIDS_IF_FORWARD), ACIDS_ENOTSUP,
                  "stft_polar_fwd: only the raw phase and the forward-difference IF are fused (mode %d, method %d)", phase_mode, if_method);
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, 1, n_frames);
    if (rc) return rc;
    rc = fill_epilogue(p.ep, band, n_fft / 2 + 1, contrast, eps, drop_first, kBandSmemBudget);
    if (rc) return rc;
    ACIDS_REQUIRE(p.ep.n_cols == n_fft / 2 + 1, ACIDS_EINVAL,
                  "stack expects each tensor to be equal size, but got [%d] and [%d] bins", p.ep.n_cols - drop_first, n_fft / 2 + 1 - drop_first);
    ACIDS_REQUIRE(mag_out && ph_out, ACIDS_EINVAL, "stft_polar_fwd: NULL output");
    ACIDS_REQUIRE(mag_row_stride >= 0 && ph_row_stride >= 0 && (n_frames + 4) * mag_row_stride < ((int64_t)1 << 31) &&
                  (n_frames + 4) * ph_row_stride < ((int64_t)1 << 31), ACIDS_ENOTSUP, "stft_polar_fwd: more than 2^31 output elements per clip");
    p.offset_ptr = mag_offset; p.scale_ptr = mag_scale;
    p.out = mag_out; p.out_clip_stride = mag_clip_stride; p.out_row_stride = mag_row_stride; p.out_col_stride = 1;
    p.power = 1.0f;
    p.midside = midside;
    p.ph_out = ph_out; p.ph_clip_stride = ph_clip_stride; p.ph_row_stride = ph_row_stride;
    p.ph_offset_ptr = ph_offset; p.ph_scale_ptr = ph_scale;
    p.ph_mode = phase_mode; p.ph_weighted = weighted;
    return dispatch_fwd(real_variant(p, false, true), n_fft, p, static_cast<cudaStream_t>(stream));
}

extern "C" ACIDS_API int acids_stft_stats(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                                int hop, int64_t n_frames, int contrast, float eps, void* scratch, double* out4, void* stream) {
    FwdParams p{};
    int rc = fill_common(p, x, B, L, ldx, window, n_fft, hop, 1, n_frames);
    if (rc) return rc;
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(scratch && out4, ACIDS_EINVAL, "stft_stats: NULL pointer");
    ACIDS_REQUIRE(B * n_frames >= 1, ACIDS_EINVAL, "stft_stats: empty input");
    p.ep.contrast = contrast; p.ep.eps = eps;
    p.stat_part = static_cast<StatAcc*>(scratch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int parts = dispatch_fwd(VAR_STATS, n_fft, p, st);
    if (parts < 0) return parts;
    ACIDS_REQUIRE(parts >= 1 && parts <= kStatsBlocks, ACIDS_ECUDA, "stft_stats: %d partials do not fit the scratch buffer", parts);
    return launch_stats_final(p.stat_part, parts, B * n_frames * (int64_t)(n_fft / 2 + 1), out4, st);
}
