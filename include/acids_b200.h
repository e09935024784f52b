/* acids_b200.h — C ABI of libacids_b200.so, the B200 (sm_100a) implementation of the
 * acids_transforms spectral hot path.
 *
 * This is the drop-in boundary.  The reference is a pure-Python package whose
 * arithmetic lives in torch/torchaudio calls; each entry point below replaces the call site(s)
 * cited next to it (paths relative to the reference repository root).  The Python host in
 * acids_transforms_b200/ binds these symbols with ctypes and exposes them through the
 * reference's own nn.Module API; INTEGRATION.md shows the stub a maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch types, no C++ exceptions cross this boundary;
 *   - every pointer is a DEVICE pointer to contiguous memory unless a stride is passed;
 *   - `stream` is a cudaStream_t passed as void*; calls enqueue work on it and return without
 *     synchronising; the library never allocates or frees device memory;
 *   - complex tensors are interleaved (re, im) float32 pairs, i.e. torch.complex64;
 *   - return value: 0 on success, a negative ACIDS_E* code otherwise; acids_last_error()
 *     gives the message for the calling thread;
 *   - `offset` / `scale` are DEVICE pointers to one float each (the Normalize buffers,
 *     norm.py:22-23) or NULL for "no normalisation" — no host sync is needed to use them.
 */
#ifndef ACIDS_B200_H
#define ACIDS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACIDS_ABI_VERSION 5

#if defined(__GNUC__)
#define ACIDS_API __attribute__((visibility("default")))
#else
#define ACIDS_API
#endif

#define ACIDS_OK 0
#define ACIDS_EINVAL (-1)      /* bad argument (shape, NULL pointer, alignment)            */
#define ACIDS_ENOTSUP (-2)     /* n_fft not a power of two in [32, 16384], or overlap too large */
#define ACIDS_ECUDA (-3)       /* CUDA runtime error at launch                              */
#define ACIDS_EWORKSPACE (-4)  /* caller-provided workspace too small                       */

/* contrast ids — Magnitude.contrast, spectral_repr.py:191-201 */
#define ACIDS_CONTRAST_NONE 0
#define ACIDS_CONTRAST_LOG1P 1   /* log(1 + m) */
#define ACIDS_CONTRAST_LOG 2     /* log(clamp(m, eps)) */
#define ACIDS_CONTRAST_LOG10 3   /* log10(clamp(m, eps)) */

/* IF methods — IF.get_if, spectral_repr.py:319-335 */
#define ACIDS_IF_FORWARD 0
#define ACIDS_IF_BACKWARD 1
#define ACIDS_IF_CENTRAL 2

/* phase-kernel modes */
#define ACIDS_PHASE_RAW 0        /* angle(X)                      Phase(unwrap=False) */
#define ACIDS_PHASE_UNWRAP 1     /* unwrap(angle(X)) over frames  Phase(unwrap=True)  */
#define ACIDS_PHASE_IF 2         /* instantaneous frequency       IF                  */

/* statistics input kinds for acids_stats */
#define ACIDS_STATS_REAL 0           /* float32 values as they are                       */
#define ACIDS_STATS_CABS_CONTRAST 1  /* complex64 in, statistics of contrast(|x|)        */
#define ACIDS_STATS_ABS_CONTRAST 2   /* float32 in, statistics of contrast(|x|) (Magnitude.scale_data on a real spectrogram) */

/* one-hot layouts — MuLaw.one_hot, raw.py:285-292 */
#define ACIDS_ONEHOT_NONE 0
#define ACIDS_ONEHOT_CATEGORICAL 1   /* [..., L, C] */
#define ACIDS_ONEHOT_CHANNEL 2       /* [..., C, L] */

ACIDS_API int acids_abi_version(void);
ACIDS_API const char* acids_last_error(void);

/* A banded (sparse-by-columns) matrix in "group-ELL" layout.  This is how the 99.6 %-sparse mel banks of
 * spectral_repr.py:173-189 are applied.  Output columns are taken in groups of 32; group g applies cnt[g]
 * taps to each of its columns, column m reading input rows start[m] .. start[m] + cnt[g] - 1 (narrower
 * bands are zero padded; start[m] + cnt[g] never exceeds the number of input rows).
 *   meta = (cnt[g], base[g]) for g in [0, ceil(n_out / 32))  followed by  start[0 .. n_out)
 *   tap u of column m = coef[(base[g] + u) * 32 + (m & 31)],  coef_len = 32 * sum_g cnt[g]             */
typedef struct acids_band {
    const int32_t* meta;   /* device, n_out + 2 * ceil(n_out / 32) ints, or NULL for "no projection" */
    const float* coef;     /* device, coef_len floats */
    int32_t n_out;         /* number of output columns */
    int32_t coef_len;
    int32_t n_in;          /* number of input rows the matrix was built for; every entry point checks it against the
                            * row length it is applied to (the reference's matmul raises on a mismatch, e.g. an
                            * STFT(512) feeding Magnitude(n_fft=1024)) */
} acids_band;

/* ---- (1) framing + window + real FFT ------------------------------------------------------
 * Replaces torch.stft(...).transpose(-2,-1) at stft.py:101-102 and dgt.py:67-68 (centre=1:
 * reflect padding by n_fft/2, n_frames = 1 + L / hop), and torch.fft.rfft(x * window) on
 * pre-framed input at stft.py:251 / dgt.py:287 (centre=0, hop = n_fft, L = n_frames * n_fft).
 *   x [B, L] (row stride ldx), window [n_fft]  ->  out complex64 [B, n_frames, n_fft/2 + 1]   */
ACIDS_API int acids_stft_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                   int n_fft, int hop, int center, int64_t n_frames, float* out, void* stream);

/* The same with MidSide.forward (raw.py:145-161) folded into the sample loads: x is stereo [B/2, 2, L] (contiguous) and clip
 * 2p + c of the output is the spectrum of channel c of (mid, side) = ((l + r) / 2 [/ sqrt 2 when midside == 2], (l - r) / 2).
 * Replaces raw.py:145-161 -> stft.py:101-102 without writing the mid/side waveform (one kernel and 8 B/sample less).     */
ACIDS_API int acids_midside_stft_fwd(const float* x, int64_t B, int64_t L, const float* window, int n_fft, int hop,
                           int64_t n_frames, int midside, float* out, void* stream);

/* ---- (2) fused STFT + |.| + banded mel + contrast + normalise ------------------------------
 * Replaces the chain stft.py:101-102 -> spectral_repr.py:217-225 -> norm.py:41 without
 * materialising the complex spectrum.  out float32: row (b, t) starts at
 * out + b * out_clip_stride + t * out_row_stride and holds (n_cols - drop_first) values,
 * where n_cols = band.n_out (or n_fft/2+1 without projection) and drop_first = 1 reproduces
 * `keep_nyquist=False` (spectral_repr.py:224-225, which drops bin 0).                          */
ACIDS_API int acids_stft_mag_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                       int n_fft, int hop, int center, int64_t n_frames, acids_band band,
                       int contrast, float eps, const float* offset, const float* scale,
                       int drop_first, float* out, int64_t out_clip_stride, int64_t out_row_stride,
                       void* stream);

/* ---- (2a) fused [MidSide ->] STFT -> Polar / PolarIF ------------------------------------------
 * Replaces the chain raw.py:145-162 (optional) -> stft.py:101-102 -> spectral_repr.py:431-440 (magnitude(x),
 * phase(x), stack) in ONE kernel: the spectrum never reaches HBM.  The magnitude rows are those of (2); the
 * phase rows (n_fft/2+1 - drop_first values at ph_out + b * ph_clip_stride + t * ph_row_stride) hold
 *   phase_mode = ACIDS_PHASE_RAW: angle(X)                                   (spectral_repr.py:270-278)
 *   phase_mode = ACIDS_PHASE_IF, if_method = ACIDS_IF_FORWARD: the forward-difference instantaneous frequency
 *     (spectral_repr.py:319-323, :352-356) evaluated as the wrapped difference of two consecutive raw phases —
 *     the difference of the unwrapped phases of utils/misc.py:12-26 without their running sum;
 * any other phase mode returns ACIDS_ENOTSUP (use acids_stft_fwd + acids_polar_fwd).  Both outputs usually point
 * into the two slots of one stacked [B, T, 2, F'] tensor.
 * midside: 0 = x is [B, L]; 1 / 2 = x is stereo [B/2, 2, L] and clip 2p+c is channel c of MidSide.forward
 * (1: pad_mid=False, 2: pad_mid=True, mid additionally / sqrt 2), evaluated in the sample loads.               */
ACIDS_API int acids_stft_polar_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                       int n_fft, int hop, int64_t n_frames, int midside, acids_band band, int contrast,
                       float eps, const float* mag_offset, const float* mag_scale, int phase_mode,
                       int if_method, int weighted, const float* ph_offset, const float* ph_scale,
                       int drop_first, float* mag_out, int64_t mag_clip_stride, int64_t mag_row_stride,
                       float* ph_out, int64_t ph_clip_stride, int64_t ph_row_stride, void* stream);

/* Same epilogue on an existing spectrum: Magnitude.forward, spectral_repr.py:215-226.
 *   X complex64 [rows, n_bins] -> out float32, row r at out + r * out_row_stride.              */
ACIDS_API int acids_mag_epilogue(const float* X, int64_t rows, int n_bins, acids_band band, int contrast,
                       float eps, const float* offset, const float* scale, int drop_first,
                       float* out, int64_t out_row_stride, void* stream);

/* Magnitude.invert, spectral_repr.py:228-240: m = contrast^-1(y * scale + offset) [zero-padded by
 * one bin when pad_last] @ inverse bank.  y [rows, n_in] (row stride y_row_stride) -> out [rows, n_out]. */
ACIDS_API int acids_mag_invert(const float* y, int64_t rows, int n_in, int64_t y_row_stride, int pad_last,
                     acids_band inverse_band, int contrast, float eps, const float* offset,
                     const float* scale, float* out, void* stream);

/* ---- (2b) mel spectrogram / MFCC: mel.py:68-73 (torchaudio MelSpectrogram) ------------------
 * periodic-Hann STFT -> |X|^power -> banded mel [n_fft/2+1 -> n_mels] -> out [B, n_mels, n_frames]
 * (frequency-major like torchaudio), optionally normalised.                                    */
ACIDS_API int acids_melspec_fwd(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window,
                      int n_fft, int hop, int64_t n_frames, acids_band mel, float power,
                      const float* offset, const float* scale, float* out, void* stream);

/* Opt-in DCT tail following torchaudio.transforms.MFCC (functional.py:390-404, :636-667):
 * dB = 10 log10(clamp(mel, 1e-10)), floored at (max over a GROUP of clips) - top_db, then
 * out = dct^T dB.  torchaudio takes that max per leading batch item over the packed
 * (channel, mel, time) dims — and over the WHOLE tensor for inputs of 3 or fewer dims — so the
 * caller says how many consecutive clips share one max (clips_per_group, must divide B).
 *   mel [B, n_mels, n_frames], dct [n_mels, n_mfcc] -> out [B, n_mfcc, n_frames].
 * group_max: device scratch of B / clips_per_group floats.  top_db < 0 disables the floor.      */
ACIDS_API int acids_mfcc_dct(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct,
                   int n_mfcc, float top_db, int64_t clips_per_group, float* group_max, float* out,
                   void* stream);

/* The same tail with the DCT as a tensor-core GEMM (tcgen05.mma kind::tf32 with 3xTF32 operand splitting, accumulator
 * in TMEM; csrc/mfcc_tc.cu).  Same arguments; n_mels % 8 == 0, n_mels <= 128, n_mfcc <= 48, else ACIDS_ENOTSUP.   */
ACIDS_API int acids_mfcc_dct_tc(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct,
                   int n_mfcc, float top_db, int64_t clips_per_group, float* group_max, float* out,
                   void* stream);

/* ---- (3) phase / unwrap / instantaneous frequency ------------------------------------------
 * Phase.forward (spectral_repr.py:270-278), unwrap (utils/misc.py:12-26), IF.forward
 * (spectral_repr.py:319-357).  X complex64 [B, n_frames, n_bins]; out float32 rows like (2).
 * `weighted` applies the parabolic frame weighting of spectral_repr.py:337-345.               */
ACIDS_API int acids_phase_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, int mode,
                    int if_method, int weighted, const float* offset, const float* scale,
                    int drop_first, float* out, int64_t out_clip_stride, int64_t out_row_stride,
                    void* stream);

/* Polar / PolarIF on an existing spectrum WITH a mel bank, one read of X (spectral_repr.py:431-440: magnitude(x), phase(x),
 * stack): row-tile kernel, the magnitude rows go through the banded projection while the raw phase of the same rows is
 * kept in a second shared-memory tile; phase_mode ACIDS_PHASE_RAW, or ACIDS_PHASE_IF with ACIDS_IF_FORWARD (the wrapped
 * difference of consecutive raw phases); other modes and rows longer than 4352 bins return ACIDS_ENOTSUP (use
 * acids_mag_epilogue + acids_phase_fwd).  Row (b, t) of both outputs is at base + (b * n_frames + t) * row_stride.     */
ACIDS_API int acids_polar_rows_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, acids_band band, int contrast,
                         float eps, const float* mag_offset, const float* mag_scale, int phase_mode, int if_method,
                         int weighted, const float* ph_offset, const float* ph_scale, int drop_first, float* mag_out,
                         int64_t mag_row_stride, float* ph_out, int64_t ph_row_stride, void* stream);

/* SpectralRepresentation.forward for Polar / PolarIF without a mel bank (spectral_repr.py:434-440: magnitude(x),
 * phase(x), stack): acids_mag_epilogue (band-less) and acids_phase_fwd from ONE read of the spectrum.  mag_out /
 * ph_out rows like (2); both may point into the same stacked tensor.                                       */
ACIDS_API int acids_polar_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, int contrast, float eps,
                    const float* mag_offset, const float* mag_scale, int mode, int if_method, int weighted,
                    const float* ph_offset, const float* ph_scale, int drop_first,
                    float* mag_out, int64_t mag_clip_stride, int64_t mag_row_stride,
                    float* ph_out, int64_t ph_clip_stride, int64_t ph_row_stride, void* stream);

/* IF.invert (spectral_repr.py:359-375; fint_* utils/misc.py:82-104) and Phase.invert
 * (mode RAW/UNWRAP: de-normalise only).  y rows like (2) -> phase float32 [B, n_frames, n_bins]
 * (n_bins = n_in + pad_last).                                                                  */
ACIDS_API int acids_phase_inv(const float* y, int64_t B, int64_t n_frames, int n_in, int64_t y_clip_stride,
                    int64_t y_row_stride, int pad_last, int mode, int if_method,
                    const float* offset, const float* scale, float* out, void* stream);

/* acids_phase_inv and the recombination of spectral_repr.py:452 in one pass (SpectralRepresentation.invert,
 * spectral_repr.py:447-452): out complex64 [B, n_frames, n_bins] = mag * exp(i phase), mag float32 of the
 * same shape, contiguous.  IF method `central` is refused (ACIDS_EINVAL): call the two entry points.    */
ACIDS_API int acids_phase_inv_polar(const float* y, int64_t B, int64_t n_frames, int n_in, int64_t y_clip_stride,
                    int64_t y_row_stride, int pad_last, int mode, int if_method,
                    const float* offset, const float* scale, const float* mag, float* out, void* stream);

/* SpectralRepresentation.invert tail, spectral_repr.py:452: out = mag * exp(i phase).          */
ACIDS_API int acids_polar_to_complex(const float* mag, const float* phase, int64_t n, float* out, void* stream);

/* One fast-Griffin-Lim update (STFT.griffin_lim, stft.py:174-178 -> torchaudio functional.py:336-350):
 *   a = rebuilt - momentum / (1 + momentum) * tprev;   out = mag * a / (|a| + 1e-16)
 * rebuilt, tprev, out: complex64 [n] (16-byte aligned), mag float32 [n], n even.  The caller keeps `rebuilt` as the
 * next call's `tprev`.                                                                         */
ACIDS_API int acids_griffinlim_update(const float* rebuilt, const float* tprev, const float* mag, float momentum,
                            int64_t n, float* out, void* stream);

/* ---- (4) inverse rFFT + synthesis window + overlap-add + envelope normalisation -------------
 * Replaces torch.istft at stft.py:126-127 / dgt.py:92 (centre=1: output trimmed by n_fft/2 on both
 * sides, length hop*(n_frames-1)), atomic-free: each CTA owns a span of output samples.
 *   X complex64 [B, n_frames, n_fft/2+1], window [n_fft] -> out [B, hop*(n_frames-1)].
 * workspace: device scratch of acids_istft_workspace_bytes(...) bytes (0 for the fused path).  */
ACIDS_API int64_t acids_istft_workspace_bytes(int64_t B, int64_t n_frames, int n_fft, int hop);
ACIDS_API int acids_istft_ola(const float* X, int64_t B, int64_t n_frames, int n_fft, int hop,
                    const float* window, float* out, void* workspace, int64_t workspace_bytes,
                    void* stream);

/* Per-frame inverse: torch.fft.irfft(X) * inv_window at stft.py:266 / dgt.py:302.
 *   X complex64 [rows, n_fft/2+1] -> out [rows, n_fft].                                        */
ACIDS_API int acids_irfft_frames(const float* X, int64_t rows, int n_fft, const float* window, float* out,
                       void* stream);

/* Streaming overlap-add with carry: OverlapAdd.invert, oadd.py:91-104.
 *   frames [B, n, n_fft], carry_in [B, keep] (keep = (n_fft/hop - 1) * hop; NULL = zeros)
 *   -> out [B, n*hop + n_fft - hop - keep... see oadd.py:97], carry_out [B, keep]; out is divided
 *   by `gain`.  out_len = (n-1)*hop + n_fft - keep.                                             */
ACIDS_API int acids_ola_stream(const float* frames, int64_t B, int64_t n, int n_fft, int hop, int64_t keep,
                     const float* carry_in, float gain, float* out, float* carry_out, void* stream);

/* ---- (4b) block-by-block streaming step as ONE kernel -------------------------------
 * The reference's real-time use (README.md:4): per block of n * hop new samples per stream,
 *   analysis :  OverlapAdd.forward (oadd.py:70-74: saved tail ++ block, framed with hop) ->
 *               RealtimeSTFT / RealtimeDGT.forward (stft.py:248-253, dgt.py:284-289: rfft(frame * window))
 *   synthesis:  RealtimeSTFT / RealtimeDGT.invert (stft.py:259-266, dgt.py:296-302: irfft(X) * inv_window) ->
 *               OverlapAdd.invert (oadd.py:91-104: carried tail + rectangular overlap-add, / gain, next carry)
 * One CTA per stream; `tail` (OverlapAdd.input_buffer) and `carry` (OverlapAdd.output_buffer), both [B, n_fft - hop],
 * live at fixed addresses and are advanced IN PLACE, so a step is one launch and capturable in a CUDA graph as is.
 * hop must divide n_fft, and a block must be at least as long as the carried tail (n * hop >= n_fft - hop).
 * Synthesis is bit-identical to acids_irfft_frames + acids_ola_stream; analysis agrees with acids_stft_fwd(center=0) to
 * rounding (<= 2e-6 of the peak; the two kernels' packed multiply-adds are contracted differently by ptxas).
 *   x [B, n*hop], X complex64 [B, n, n_fft/2+1], out [B, n*hop].
 * acids_stream_roundtrip is analysis followed by synthesis with the spectrum kept in registers (X_out may be NULL).     */
ACIDS_API int acids_stream_analysis(const float* x, int64_t B, int64_t n, int n_fft, int hop, const float* window,
                          float* tail, float* X_out, void* stream);
ACIDS_API int acids_stream_synthesis(const float* X, int64_t B, int64_t n, int n_fft, int hop, const float* inv_window,
                           float gain, float* carry, float* out, void* stream);
ACIDS_API int acids_stream_roundtrip(const float* x, int64_t B, int64_t n, int n_fft, int hop, const float* window,
                           const float* inv_window, float gain, float* tail, float* carry, float* X_out, float* out,
                           void* stream);

/* Dense mel projection on the tensor cores (mel.py:38-44, torchaudio MelScale: spec @ fb): spec float32 [B, n_frames, n_bins]
 * (|X| or |X|^2 rows, unit stride) x bank float32 [n_bins, n_mels] (n_mels <= 128) -> out float32 [B, n_mels, n_frames].
 * tcgen05.mma kind::tf32 with 3xTF32 operand splitting (fp32 fidelity), accumulator in TMEM, the packed bank streamed by bulk
 * asynchronous copies.  The fused kernels (acids_melspec_fwd, acids_stft_mag_fwd) use the banded FP32 form instead, which
 * never materialises the spectrum; this is the dense formulation for callers that hold one.  workspace (16-byte aligned):
 * acids_mel_tc_workspace_bytes(n_bins).                                                                                     */
ACIDS_API int64_t acids_mel_tc_workspace_bytes(int n_bins);
ACIDS_API int acids_mel_tc(const float* spec, int64_t B, int64_t n_frames, int n_bins, const float* bank, int n_mels, void* workspace,
                 int64_t workspace_bytes, float* out, void* stream);

/* ---- (4c) phase-gradient heap integration (PGHI): DGT.pghi, dgt.py:156-236 — the reference's default inversion of a magnitude ----
 * mag float32 [B, n_frames, n_bins] (clamped at eps inside) -> phase float32 of the same shape.  One CTA per clip: log-magnitude,
 * threshold tol * max and every region's arg-max seed in parallel, the priority-queue flood fill by one thread in the
 * reference's visiting order ((-|X|, t, k) keys) and float32 operation order.  workspace: acids_pghi_workspace_bytes().     */
ACIDS_API int64_t acids_pghi_workspace_bytes(int64_t B, int64_t n_frames, int n_bins);
ACIDS_API int acids_pghi(const float* mag, int64_t B, int64_t n_frames, int n_bins, float gamma, int n_fft, int hop, double tol,
               float eps, void* workspace, int64_t workspace_bytes, float* phase, void* stream);

/* Frame-by-frame variant (RealtimeDGT.pghi, dgt.py:338-452): phase of n_frames NEW frames per stream, continued from the two
 * remembered frames hist_mag [B, 2, n_bins] and the phase hist_phase [B, n_bins] of the newer one.  Every new frame is filled
 * from the previous frame's audible bins (steps along time) and its own loudest bin (steps along bins); the heap and the
 * frame's rows live in shared memory, one CTA per stream.  noise [B, n_frames, n_bins] (may be NULL: zeros) is what bins
 * below max(tol * max, eps) receive (the reference draws randn).  The stencil row before the first remembered frame
 * replicates it (the reference reads uninitialised memory there).  workspace: acids_rt_pghi_workspace_bytes().          */
ACIDS_API int64_t acids_rt_pghi_workspace_bytes(int64_t B, int64_t n_frames, int n_bins);
ACIDS_API int acids_rt_pghi(const float* mag, const float* hist_mag, const float* hist_phase, const float* noise, int64_t B,
                  int64_t n_frames, int n_bins, float gamma, int n_fft, int hop, float tol, float eps, void* workspace,
                  int64_t workspace_bytes, float* phase, void* stream);

/* ---- (5) mu-law and one-hot ------------------------------------------------------------------
 * torchaudio mu_law_encoding / decoding (functional.py:690-700, :723-729) as used by raw.py:282-316.
 * log1p_mu = float32 log1p(channels-1) evaluated by the caller the way the reference does (host).
 * reciprocal_divide: 0 = IEEE division by log1p_mu (CPU eager chain), 1 = multiply by its
 * float32 reciprocal (what the CUDA eager chain computes).  out int64.
 * one_hot: ACIDS_ONEHOT_*; for CATEGORICAL out is [n, C], for CHANNEL [outer, C, inner]
 * with n = outer * inner.                                                                       */
ACIDS_API int acids_mulaw_encode(const float* x, int64_t outer, int64_t inner, int channels, float log1p_mu,
                       int reciprocal_divide, int one_hot, int64_t* out, void* stream);
ACIDS_API int acids_mulaw_decode(const int64_t* q, int64_t n, int channels, float log1p_mu,
                                 int reciprocal_divide, float* out, void* stream);
/* F.one_hot for OneHot.forward, misc.py:176-179: q [n] int64 -> out [n, n_classes] int64.        */
ACIDS_API int acids_one_hot(const int64_t* q, int64_t n, int n_classes, int64_t* out, void* stream);

/* ---- Normalize.scale_data statistics: norm.py:26-38, spectral_repr.py:242-245 -----------------
 * out4 (device, 4 doubles): min, max, mean, unbiased std of the (transformed) values.
 * kind = ACIDS_STATS_REAL: x float32 [n]; ACIDS_STATS_CABS_CONTRAST: x complex64 [n], statistics of
 * contrast(|x|); ACIDS_STATS_ABS_CONTRAST: x float32 [n], statistics of contrast(|x|).
 * scratch: device buffer of acids_stats_scratch_bytes() bytes.                                    */
ACIDS_API int64_t acids_stats_scratch_bytes(void);
ACIDS_API int acids_stats(const float* x, int64_t n, int kind, int contrast, float eps, void* scratch,
                double* out4, void* stream);

/* The same statistics for the spectrum of a waveform WITHOUT materialising it: Magnitude.scale_data(STFT(x))
 * (base.py:144-148 -> stft.py:101-102 -> spectral_repr.py:242-245 -> norm.py:26-38) in one pass of the fused
 * forward kernel — min / max / sum / sum of squares of contrast(|X|) over every bin are kept per CTA and merged by a
 * one-block kernel.  x [B, L] (row stride ldx), centre-padded framing like acids_stft_fwd; scratch / out4 as above.   */
ACIDS_API int acids_stft_stats(const float* x, int64_t B, int64_t L, int64_t ldx, const float* window, int n_fft,
                     int hop, int64_t n_frames, int contrast, float eps, void* scratch, double* out4, void* stream);

/* ---- raw-domain prologues: raw.py:34-49 (Mono mix), raw.py:145-180 (MidSide) -------------------
 * x [B, 2, L] -> mono [B, L] = (l + r) / 2;  mid/side [B, 2, L] (pad_mid: mid / sqrt(2)).        */
ACIDS_API int acids_mono_mix(const float* x, int64_t B, int64_t L, float* out, void* stream);
ACIDS_API int acids_midside(const float* x, int64_t B, int64_t L, int pad_mid, int inverse, float* out,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ACIDS_B200_H */
