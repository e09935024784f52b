// spectral_repr.cu — representation kernels on an existing complex spectrum.
// Rows A6, A8 (Magnitude forward / invert), A10-A13 (Phase, unwrap, IF forward / invert), A14 (polar
// recombination) of SURVEY.md §8(a).  All HBM bound: one read of the input, one write of the output.
#include "common.cuh"
#include <type_traits>

namespace acids {

// ---------------------------------------------------------------------------------------------
// Magnitude.forward on a spectrum: one warp per row; |X| staged in shared memory so that the banded
// mel projection can read neighbouring bins.
// ---------------------------------------------------------------------------------------------
struct MagParams {
    const float2* X;
    int64_t rows;
    int n_bins;
    EpiParams ep;
    const float* offset_ptr;
    const float* scale_ptr;
    float* out;
    int64_t out_row_stride;
};

template <int BAND, int ROWS, bool PF, int NT = 256>
__global__ void __launch_bounds__(NT) mag_epilogue_kernel(const MagParams p) {
    // ROWS rows per CTA iteration: 8 / ROWS warps stage |X| of one row, then the CTA projects the rows together (tiles
    // of up to 4).  PF: a lane's share of the row fits 17 registers pairs, so the NEXT iteration's spectrum is loaded
    // before this iteration's projection and lands while it runs (ncu on the unpipelined kernel: 57 % of the stall
    // samples on the first use of the staging loads, the CTAs of an SM loading and projecting in lock-step).
    // NT = 512 for rows of 1089 .. 4352 bins: 4 (2) rows per iteration keep the 4-row (2-row) projection tiles while a
    // lane's share still fits the prefetch registers (with 256 threads and 2-row tiles the kernel is MIO bound instead).
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPR = (NT / 32) / ROWS;              // warps per row
    constexpr int NF = ROWS < 4 ? ROWS : 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wrow = warp / WPR, wsub = warp % WPR;
    const int stride = (p.n_bins + 3) & ~3;
    float* val = reinterpret_cast<float*>(smem_raw);
    int32_t* smeta = reinterpret_cast<int32_t*>(smem_raw + (size_t)ROWS * stride * sizeof(float));
    float* scoef = reinterpret_cast<float*>(smem_raw + (size_t)ROWS * stride * sizeof(float) + p.ep.band_bytes_meta);
    if (BAND == BAND_SMEM) stage_band(p.ep, smeta, scoef);
    const EpiArgs ea = make_epi_args(p.ep, BAND == BAND_SMEM ? smeta : p.ep.meta, BAND == BAND_SMEM ? scoef : p.ep.coef,
                                     p.offset_ptr, p.scale_ptr);
    __syncthreads();
    const int rs = (int)p.out_row_stride;
    auto project = [&](int64_t r0) {
        const int n_valid = (int)min((int64_t)ROWS, p.rows - r0);
#pragma unroll 1
        for (int g0 = 0; g0 < ROWS && g0 < n_valid; g0 += NF)
            epilogue_dispatch<NT, NF, -1, BAND, false>(p.ep.contrast, val + g0 * stride, stride, threadIdx.x, ea,
                                                       p.out + (r0 + g0) * p.out_row_stride, rs, 1, n_valid - g0);
    };
    const int64_t step = (int64_t)gridDim.x * ROWS;
    if constexpr (PF) {
        constexpr int U = 17;                          // 32 * WPR * U >= n_bins (checked by the launcher)
        const int kl = lane + 32 * wsub;               // this lane's bins: kl + 32 * WPR * j
        float2 a[U];
        auto issue = [&](int64_t r0) {
            const int64_t r = r0 + wrow;
            const float2* __restrict__ row = p.X + r * p.n_bins + kl;
#pragma unroll
            for (int j = 0; j < U; ++j)
                a[j] = (r < p.rows && kl + 32 * WPR * j < p.n_bins) ? ldg_stream2(row + 32 * WPR * j) : make_float2(0.f, 0.f);
        };
        int64_t r0 = (int64_t)blockIdx.x * ROWS;
        if (r0 < p.rows) issue(r0);
        for (; r0 < p.rows; r0 += step) {
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (kl + 32 * WPR * j < p.n_bins) val[wrow * stride + kl + 32 * WPR * j] = fast_sqrt(a[j].x * a[j].x + a[j].y * a[j].y);
            __syncthreads();
            if (r0 + step < p.rows) issue(r0 + step);
            project(r0);
            __syncthreads();
        }
    } else {
        for (int64_t r0 = (int64_t)blockIdx.x * ROWS; r0 < p.rows; r0 += step) {
            const int64_t r = r0 + wrow;
            if (r < p.rows) {
                const float2* __restrict__ row = p.X + r * p.n_bins;
                // U independent loads in flight per lane before the first use (the row is streamed once)
                constexpr int U = ROWS == 4 ? 16 : 8;
                for (int k0 = lane + 32 * U * wsub; k0 < p.n_bins; k0 += 32 * U * WPR) {
                    float2 a[U];
#pragma unroll
                    for (int j = 0; j < U; ++j) a[j] = (k0 + 32 * j < p.n_bins) ? ldg_stream2(row + k0 + 32 * j) : make_float2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        if (k0 + 32 * j < p.n_bins) val[wrow * stride + k0 + 32 * j] = fast_sqrt(a[j].x * a[j].x + a[j].y * a[j].y);
                }
            }
            __syncthreads();
            project(r0);
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Polar / PolarIF on a spectrum WITH a mel bank, one read of X (spectral_repr.py:431-440): the magnitude half needs whole
// rows (banded projection), the IF half the previous frame's phase.  Row-tile kernel derived from mag_epilogue_kernel<PF>:
// while a lane holds X for its share of a row it parks |X| in the magnitude tile AND the raw phase in a second tile; the
// CTA then projects the magnitude rows and emits the phase rows — raw, or the forward-difference IF as the wrapped
// difference of two consecutive raw phases (the arithmetic of the fused STFT -> Polar epilogue).  A CTA owns a CONTIGUOUS
// run of row tiles, so the phase row preceding a tile is the last row of the previous tile, kept in a ring of ROWS + 1
// phase rows; only the first tile of a run recomputes it from the spectrum (one extra row per CTA).
// 8 B read + 8 B written per bin instead of 16 + 8 for the two kernels it replaces.
// ---------------------------------------------------------------------------------------------
struct PolarRowsParams {
    const float2* X;
    int64_t rows;            // B * n_frames
    int n_frames, n_bins;
    EpiParams ep;
    const float* offset_ptr;
    const float* scale_ptr;
    float* out;              // magnitude rows: row r at out + r * out_row_stride
    int64_t out_row_stride;
    float* ph_out;           // phase rows: row r at ph_out + r * ph_row_stride
    int64_t ph_row_stride;
    const float* ph_offset_ptr;
    const float* ph_scale_ptr;
    int ph_mode, ph_weighted;
};

template <int BAND, int ROWS, int NT>
__global__ void __launch_bounds__(NT) polar_rows_kernel(const PolarRowsParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPR = (NT / 32) / ROWS;              // warps per row
    constexpr int NF = ROWS < 4 ? ROWS : 4;
    constexpr int U = 17;                              // 32 * WPR * U >= n_bins (checked by the launcher)
    constexpr int RING = ROWS + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wrow = warp / WPR, wsub = warp % WPR;
    const int stride = (p.n_bins + 3) & ~3;
    float* val = reinterpret_cast<float*>(smem_raw);
    float* phs = val + (size_t)ROWS * stride;          // ring of ROWS + 1 raw-phase rows
    unsigned char* bandmem = reinterpret_cast<unsigned char*>(phs + (size_t)RING * stride);
    int32_t* smeta = reinterpret_cast<int32_t*>(bandmem);
    float* scoef = reinterpret_cast<float*>(bandmem + p.ep.band_bytes_meta);
    if (BAND == BAND_SMEM) stage_band(p.ep, smeta, scoef);
    const EpiArgs ea = make_epi_args(p.ep, BAND == BAND_SMEM ? smeta : p.ep.meta, BAND == BAND_SMEM ? scoef : p.ep.coef,
                                     p.offset_ptr, p.scale_ptr);
    const float ph_off = p.ph_offset_ptr ? __ldg(p.ph_offset_ptr) : 0.f;
    const float ph_inv = p.ph_scale_ptr ? 1.0f / __ldg(p.ph_scale_ptr) : 1.0f;
    const bool ph_if = p.ph_mode == ACIDS_PHASE_IF;
    const int T = p.n_frames;
    const int rs = (int)p.out_row_stride;
    const int df = p.ep.drop_first, n_keep = p.n_bins - df;
    // contiguous run of row tiles
    const int64_t tiles = (p.rows + ROWS - 1) / ROWS;
    const int64_t i0 = tiles * blockIdx.x / gridDim.x, i1 = tiles * (blockIdx.x + 1) / gridDim.x;
    if (i0 >= i1) return;
    int64_t r0 = i0 * ROWS;
    int t_first = (int)(r0 % T);                        // frame index of the tile's first row (advanced incrementally)
    int base = 0;                                       // ring slot of the row preceding the tile
    if (ph_if && t_first != 0) {
        // the phase of the row before the run (same clip): recomputed once per CTA
        const float2* __restrict__ row = p.X + (r0 - 1) * p.n_bins;
        for (int k = threadIdx.x; k < p.n_bins; k += NT) {
            const float2 a = __ldg(row + k);
            phs[k] = fast_atan2f(a.y, a.x);
        }
    }
    __syncthreads();
    const int kl = lane + 32 * wsub;                    // this lane's bins: kl + 32 * WPR * j
    float2 a[U];
    auto issue = [&](int64_t rr0) {
        const int64_t r = rr0 + wrow;
        const float2* __restrict__ row = p.X + r * p.n_bins + kl;
#pragma unroll
        for (int j = 0; j < U; ++j)
            a[j] = (r < p.rows && kl + 32 * WPR * j < p.n_bins) ? ldg_stream2(row + 32 * WPR * j) : make_float2(0.f, 0.f);
    };
    issue(r0);
    for (int64_t i = i0; i < i1; ++i, r0 += ROWS) {
        {
            int slot = base + 1 + wrow;
            if (slot >= RING) slot -= RING;
            float* __restrict__ vrow = val + wrow * stride;
            float* __restrict__ prow = phs + slot * stride;
#pragma unroll
            for (int j = 0; j < U; ++j)
                if (kl + 32 * WPR * j < p.n_bins) {
                    vrow[kl + 32 * WPR * j] = fast_sqrt(a[j].x * a[j].x + a[j].y * a[j].y);
                    prow[kl + 32 * WPR * j] = fast_atan2f(a[j].y, a[j].x);
                }
        }
        __syncthreads();
        if (i + 1 < i1) issue(r0 + ROWS);
        const int n_valid = (int)min((int64_t)ROWS, p.rows - r0);
        // magnitude rows: banded projection -> contrast -> normalise
#pragma unroll 1
        for (int g0 = 0; g0 < ROWS && g0 < n_valid; g0 += NF)
            epilogue_dispatch<NT, NF, -1, BAND, false>(p.ep.contrast, val + g0 * stride, stride, threadIdx.x, ea,
                                                       p.out + (r0 + g0) * p.out_row_stride, rs, 1, n_valid - g0);
        // phase rows (spectral_repr.py:270-278 raw; :319-323 + :352-356 forward-difference IF, weighting, normalisation)
        int t = t_first;
#pragma unroll 1
        for (int g0 = 0; g0 < n_valid; ++g0) {
            int sc = base + 1 + g0, sp = base + g0;
            if (sc >= RING) sc -= RING;
            if (sp >= RING) sp -= RING;
            const float* __restrict__ cur = phs + sc * stride + df;
            const float* __restrict__ prv = phs + sp * stride + df;
            float* __restrict__ o = p.ph_out + (r0 + g0) * p.ph_row_stride;
            const bool diff = ph_if && t > 0;
            const float s_pi = (ph_if && t < T - 1) ? ACIDS_INV_PI_F : 1.f;
            const float wgt = p.ph_weighted ? if_weight(t, T) : 1.f;
            for (int k = threadIdx.x; k < n_keep; k += NT) {
                float v = cur[k];
                if (diff) {
                    const float d = v - prv[k];
                    v = (d + unwrap_correction(d)) * 0.5f;
                }
                v = v * s_pi;
                if (p.ph_weighted) v *= wgt;
                stg_stream1(o + k, (v - ph_off) * ph_inv);
            }
            if (++t == T) t = 0;
        }
        t_first += ROWS;
        while (t_first >= T) t_first -= T;
        base += ROWS;
        if (base >= RING) base -= RING;
        __syncthreads();
    }
}

// Magnitude.invert: m = contrast^-1(y * scale + offset) [zero padded] @ inverse band
struct MagInvParams {
    const float* y;
    int64_t rows;
    int n_in;
    int64_t y_row_stride;
    int pad_last;
    EpiParams ep;            // the inverse band (ep.n_cols = n_out), contrast unused
    int n_out;
    int contrast;
    float eps;
    const float* offset_ptr;
    const float* scale_ptr;
    float* out;
};

// 8 rows per CTA iteration: each warp de-normalises and inverts the contrast of one row into shared memory, then the
// CTA applies the banded inverse matrix to the rows together with the forward kernels' row-tile epilogue (tiles of 4,
// identity contrast and normalisation): per-column metadata and coefficients are fetched once per 4 rows.
template <int BAND, int ROWS>
__global__ void __launch_bounds__(256) mag_invert_kernel(const MagInvParams p) {
    // ROWS rows per CTA iteration, 8 / ROWS warps stage one row.  ncu (cfg 4, 8 rows, 8 loads per lane): 2 CTAs / SM,
    // long-scoreboard bound at 41 % of HBM peak -> ROWS = 4 for long rows (3 CTAs / SM) and 16 loads in flight per lane.
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int WPR = 8 / ROWS, U = 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wrow = warp / WPR, wsub = warp % WPR;
    const int n_val = p.n_in + p.pad_last;
    const int stride = (n_val + 3) & ~3;
    float* val = reinterpret_cast<float*>(smem_raw);
    int32_t* smeta = reinterpret_cast<int32_t*>(smem_raw + (size_t)ROWS * stride * sizeof(float));
    float* scoef = reinterpret_cast<float*>(smem_raw + (size_t)ROWS * stride * sizeof(float) + p.ep.band_bytes_meta);
    if (BAND == BAND_SMEM) stage_band(p.ep, smeta, scoef);
    EpiArgs ea = make_epi_args(p.ep, BAND == BAND_SMEM ? smeta : p.ep.meta, BAND == BAND_SMEM ? scoef : p.ep.coef, nullptr, nullptr);
    const float off = p.offset_ptr ? __ldg(p.offset_ptr) : 0.f;
    const float sc = p.scale_ptr ? __ldg(p.scale_ptr) : 1.f;
    __syncthreads();
    for (int64_t r0 = (int64_t)blockIdx.x * ROWS; r0 < p.rows; r0 += (int64_t)gridDim.x * ROWS) {
        const int64_t r = r0 + wrow;
        if (r < p.rows) {
            const float* __restrict__ row = p.y + r * p.y_row_stride;
            // U independent loads in flight per lane before the first use
            for (int k0 = lane + 32 * U * wsub; k0 < n_val; k0 += 32 * U * WPR) {
                float a[U];
#pragma unroll
                for (int j = 0; j < U; ++j) a[j] = (k0 + 32 * j < p.n_in) ? __ldg(row + k0 + 32 * j) : 0.f;
#pragma unroll
                for (int j = 0; j < U; ++j) {
                    const int k = k0 + 32 * j;
                    // the zero pad is appended BEFORE the contrast inversion (spectral_repr.py:230-234)
                    if (k < n_val) val[wrow * stride + k] = invert_contrast(k < p.n_in ? a[j] * sc + off : 0.f, p.contrast, p.eps);
                }
            }
        }
        __syncthreads();
        const int n_valid = (int)min((int64_t)ROWS, p.rows - r0);
#pragma unroll 1
        for (int g0 = 0; g0 < ROWS && g0 < n_valid; g0 += 4)
            epilogue_dispatch<256, 4, ACIDS_CONTRAST_NONE, BAND, false>(0, val + g0 * stride, stride, threadIdx.x, ea,
                                                                        p.out + (r0 + g0) * (int64_t)p.n_out, p.n_out, 1, n_valid - g0);
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// Phase / unwrap / IF: one thread per (clip, bin) column walks the frames; the unwrap correction is a
// sequential float32 running sum exactly like torch.cumsum along the frame axis (utils/misc.py:25).
// ---------------------------------------------------------------------------------------------
struct PhaseParams {
    const float2* X;
    int64_t B;
    int n_frames, n_bins;
    int mode, method, weighted;
    const float* offset_ptr;
    const float* scale_ptr;
    int drop_first;
    float* out;
    int64_t out_clip_stride, out_row_stride;
    // MAGC >= 0 only: the band-less Magnitude of the same spectrum, written from the same pass
    float* mag_out;
    int64_t mag_clip_stride, mag_row_stride;
    int contrast;
    float eps;
    const float* mag_offset_ptr;
    const float* mag_scale_ptr;
};

// exp(i x) for the polar recombination (spectral_repr.py:452).  ncu showed phase_inv_kernel<POLAR> issue bound (82 % of
// issue slots) on sincosf's ~40 instructions.  Here: x is reduced to [-pi, pi] by the nearest multiple of 2 pi (two-constant
// Cody-Waite, the products exact inside the FMAs), then the SFU sine / cosine (|abs error| <= 2^-21.4 on that interval):
// |error| < 1e-6 absolute for |x| up to ~1e5 rad, two orders below the 1e-4 parity budget.  exp(i 0) = 1 exactly.
__device__ __forceinline__ void polar_sincos(float x, float* s, float* c) {
    const float k = rintf(x * ACIDS_INV_2PI_F);
    float r = fmaf(k, -6.2831854820251465f, x);      // float(2 pi)
    r = fmaf(k, 1.7484555e-7f, r);                   // float(2 pi) - 2 pi
    *s = __sinf(r);
    *c = __cosf(r);
}

// One thread per (clip, bin) column walks the frames.  The arctangents and unwrap corrections of KB consecutive
// frames only depend on the raw spectrum: they are loaded and evaluated as a block (memory- and instruction-level
// parallelism); only the running sum and the differences run serially, in torch.cumsum's order.
// MODE / METHOD / WEIGHTED are compile-time: no branching per element.
// MAGC >= 0 (a contrast id): Polar / PolarIF without a mel bank (SpectralRepresentation.forward, spectral_repr.py:434-440)
// — the thread already holds X, so |X| -> contrast -> normalise goes to the magnitude slot from the same load instead
// of a second pass over the spectrum (same arithmetic as mag_epilogue_kernel<BAND_NONE>).
template <int MODE, int METHOD, bool WEIGHTED, int MAGC>
__global__ void __launch_bounds__(128) phase_fwd_kernel(const PhaseParams p) {
    constexpr int KB = 8;
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.B * p.n_bins) return;
    const int64_t b = col / p.n_bins;
    const int f = (int)(col - b * p.n_bins);
    if (f < p.drop_first) return;                       // the dropped bin produces no output at all
    const int T = p.n_frames;
    const int64_t xs = p.n_bins, os = p.out_row_stride;
    const float2* __restrict__ xp = p.X + b * (int64_t)T * xs + f;          // next spectrum row to load
    float* __restrict__ out = p.out + b * p.out_clip_stride + (f - p.drop_first);
    const float off = p.offset_ptr ? __ldg(p.offset_ptr) : 0.f;
    const float inv = p.scale_ptr ? 1.0f / __ldg(p.scale_ptr) : 1.0f;
    auto emit = [&](int t, float v) {
        if (WEIGHTED) v *= if_weight(t, T);
        stg_stream1(out + t * os, (v - off) * inv);
    };
    float* __restrict__ mout = MAGC >= 0 ? p.mag_out + b * p.mag_clip_stride + (f - p.drop_first) : nullptr;
    float mgain = 1.f, mbias = 0.f;
    if (MAGC >= 0) {     // (x - offset) / scale as one FFMA with the lg2 constant folded in, like make_epi_args
        const float moff = p.mag_offset_ptr ? __ldg(p.mag_offset_ptr) : 0.f;
        const float minv = p.mag_scale_ptr ? 1.0f / __ldg(p.mag_scale_ptr) : 1.0f;
        mgain = minv * contrast_gain(MAGC);
        mbias = -moff * minv;
    }
    float prev_raw = 0.f, cum = 0.f;
    float u1 = 0.f, u2 = 0.f;   // unwrapped phase at t-1, t-2
    // EDGE = false: every frame of the block is interior (1 <= t, t + 1 <= T - 2): no bounds or first / last frame cases
    auto block = [&](int t0, auto edge_) {
        constexpr bool EDGE = decltype(edge_)::value;
        float2 a[KB];
        float rawv[KB], corr[KB];
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            a[k] = (!EDGE || t0 + k < T) ? ldg_stream2(xp) : make_float2(1.f, 0.f);
            xp += xs;
        }
        if (MAGC >= 0) {
#pragma unroll
            for (int k = 0; k < KB; ++k)
                if (!EDGE || t0 + k < T)
                    stg_stream1(mout + (int64_t)(t0 + k) * p.mag_row_stride,
                                fmaf(contrast_core<MAGC>(fast_sqrt(a[k].x * a[k].x + a[k].y * a[k].y), p.eps), mgain, mbias));
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) rawv[k] = fast_atan2f(a[k].y, a[k].x);
        if (MODE != ACIDS_PHASE_RAW) {
#pragma unroll
            for (int k = 0; k < KB; ++k) corr[k] = unwrap_correction(rawv[k] - (k == 0 ? prev_raw : rawv[k - 1]));
            prev_raw = rawv[KB - 1];
        }
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const int t = t0 + k;
            if (EDGE && t >= T) break;
            const float raw = rawv[k];
            if (MODE == ACIDS_PHASE_RAW) {
                emit(t, raw);
                continue;
            }
            if (!EDGE || t > 0) cum += corr[k];
            const float un = (!EDGE || t > 0) ? raw + cum : raw;
            if (MODE == ACIDS_PHASE_UNWRAP) {
                emit(t, un);
            } else if (METHOD == ACIDS_IF_FORWARD) {
                // row 0 = phi_0, row t = (phi_t - phi_{t-1}) / 2; rows [:-1] then divided by pi
                float v = (EDGE && t == 0) ? un : (un - u1) * 0.5f;
                if (!EDGE || t < T - 1) v = v * ACIDS_INV_PI_F;        // 1 ulp from the reference's division, 1e-7 of the budget
                emit(t, v);
            } else if (METHOD == ACIDS_IF_BACKWARD) {
                // row t = (phi_t - phi_{t+1}) / 2 for t < T-1, row T-1 = phi_{T-1}; rows [1:] divided by -pi
                if (!EDGE || t > 0) {
                    float v = (u1 - un) * 0.5f;
                    if (!EDGE || t - 1 >= 1) v = v * -ACIDS_INV_PI_F;
                    emit(t - 1, v);
                }
                if (EDGE && t == T - 1) emit(t, t >= 1 ? un * -ACIDS_INV_PI_F : un);
            } else {
                // central: row 0 = phi_0, row t = (phi_{t+1} - phi_{t-1}) / 4 / (2 pi), row T-1 = phi_{T-1}
                if (EDGE && t == 0) emit(0, un);
                if (!EDGE || t >= 2) emit(t - 1, (un - u2) * 0.25f * ACIDS_INV_2PI_F);
                if (EDGE && t == T - 1 && t > 0) emit(t, un);
            }
            u2 = u1;
            u1 = un;
        }
    };
    // frames 0, 1 and the last two need the special cases of the difference schemes: first and last block are general
    int t0 = 0;
    block(t0, std::true_type{});
    for (t0 = KB; t0 + KB <= T - 2; t0 += KB) block(t0, std::false_type{});
    for (; t0 < T; t0 += KB) block(t0, std::true_type{});
}

typedef void (*PhaseFwdKernel)(const PhaseParams);
template <int MAGC>
static PhaseFwdKernel pick_phase_kernel_c(int mode, int method, int weighted) {
    if (mode == ACIDS_PHASE_RAW) return phase_fwd_kernel<ACIDS_PHASE_RAW, 0, false, MAGC>;
    if (mode == ACIDS_PHASE_UNWRAP) return phase_fwd_kernel<ACIDS_PHASE_UNWRAP, 0, false, MAGC>;
#define ACIDS_IFK(M) (weighted ? phase_fwd_kernel<ACIDS_PHASE_IF, M, true, MAGC> : phase_fwd_kernel<ACIDS_PHASE_IF, M, false, MAGC>)
    if (method == ACIDS_IF_FORWARD) return ACIDS_IFK(ACIDS_IF_FORWARD);
    if (method == ACIDS_IF_BACKWARD) return ACIDS_IFK(ACIDS_IF_BACKWARD);
    return ACIDS_IFK(ACIDS_IF_CENTRAL);
#undef ACIDS_IFK
}
// contrast < 0: phase only
static PhaseFwdKernel pick_phase_kernel(int mode, int method, int weighted, int contrast = -1) {
    switch (contrast) {
        case ACIDS_CONTRAST_NONE: return pick_phase_kernel_c<ACIDS_CONTRAST_NONE>(mode, method, weighted);
        case ACIDS_CONTRAST_LOG1P: return pick_phase_kernel_c<ACIDS_CONTRAST_LOG1P>(mode, method, weighted);
        case ACIDS_CONTRAST_LOG: return pick_phase_kernel_c<ACIDS_CONTRAST_LOG>(mode, method, weighted);
        case ACIDS_CONTRAST_LOG10: return pick_phase_kernel_c<ACIDS_CONTRAST_LOG10>(mode, method, weighted);
        default: return pick_phase_kernel_c<-1>(mode, method, weighted);
    }
}

struct PhaseInvParams {
    const float* y;
    int64_t B;
    int n_frames, n_in;
    int64_t y_clip_stride, y_row_stride;
    int pad_last, mode, method;
    const float* offset_ptr;
    const float* scale_ptr;
    float* out;   // [B, n_frames, n_in + pad_last]; complex64 (interleaved) of the same shape when POLAR
    const float* mag;   // POLAR only: [B, n_frames, n_in + pad_last]
};

// POLAR: the recombination of spectral_repr.py:452 happens in the same pass — the column's phase never goes to
// memory, the thread reads the matching magnitude and stores mag * exp(i phase).  (Not for the central method,
// whose two recurrences use the output column as scratch.)
template <bool POLAR>
__global__ void __launch_bounds__(128) phase_inv_kernel(const PhaseInvParams p) {
    const int nb = p.n_in + p.pad_last;
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= p.B * nb) return;
    const int64_t b = col / nb;
    const int f = (int)(col - b * nb);
    const int64_t col0 = b * (int64_t)p.n_frames * nb + f;
    float* __restrict__ out = p.out + (POLAR ? 2 * col0 : col0);
    const float* __restrict__ mag = POLAR ? p.mag + col0 : nullptr;
    const int T = p.n_frames;
    auto emit = [&](int t, float ph, float m) {
        if constexpr (POLAR) {
            float s, c;
            polar_sincos(ph, &s, &c);
            stg_stream2(reinterpret_cast<float2*>(out) + (int64_t)t * nb, m * c, m * s);
        } else {
            out[(int64_t)t * nb] = ph;
        }
    };
    if (f >= p.n_in) {   // the zero bin appended when keep_nyquist=False (spectral_repr.py:371-374)
        for (int t = 0; t < T; ++t) emit(t, 0.f, POLAR ? __ldg(mag + (int64_t)t * nb) : 0.f);   // phase 0: out = mag + 0i
        return;
    }
    const float* __restrict__ y = p.y + b * p.y_clip_stride + f;
    const float off = p.offset_ptr ? __ldg(p.offset_ptr) : 0.f;
    const float sc = p.scale_ptr ? __ldg(p.scale_ptr) : 1.f;
    auto den = [&](int t) { return __ldg(y + (int64_t)t * p.y_row_stride) * sc + off; };
    // Every mode but IF-central is one running sum (or none) along the column.  The loads of KB frames are issued
    // before the first dependent add / sincos so that each thread keeps 2 KB requests in flight: with one load per
    // iteration the column walk is latency-bound (measured: the POLAR form slower than the two kernels it replaces).
    //   IF forward : rows[:-1] *= pi; rows[1:] *= 2; cumsum over frames        (spectral_repr.py:365-367, utils/misc.py:82-86)
    //   IF backward: rows[1:] *= -pi; flip; rows[1:] *= 2; cumsum; flip  == suffix sum with the LAST row undoubled
    const bool integrate = p.mode == ACIDS_PHASE_IF;
    if (!integrate || p.method != ACIDS_IF_CENTRAL) {
        constexpr int KB = 8;
        const bool reverse = integrate && p.method == ACIDS_IF_BACKWARD;
        float acc = 0.f;
        for (int i0 = 0; i0 < T; i0 += KB) {
            float v[KB], m[KB];
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int i = i0 + k;
                const int t = reverse ? T - 1 - i : i;
                const bool ok = i < T;
                v[k] = ok ? __ldg(y + (int64_t)t * p.y_row_stride) : 0.f;
                m[k] = (POLAR && ok) ? __ldg(mag + (int64_t)t * nb) : 0.f;
            }
#pragma unroll
            for (int k = 0; k < KB; ++k) {
                const int i = i0 + k;
                if (i >= T) break;
                const int t = reverse ? T - 1 - i : i;
                float x = v[k] * sc + off;
                if (integrate) {
                    if (reverse) {
                        if (t >= 1) x *= -ACIDS_PI_F;
                        if (t < T - 1) x *= 2.f;
                    } else {
                        if (t < T - 1) x *= ACIDS_PI_F;
                        if (t >= 1) x *= 2.f;
                    }
                    acc += x;
                } else {
                    acc = x;
                }
                emit(t, acc, m[k]);
            }
        }
    } else if constexpr (!POLAR) {
        // central: rows[1:-1] *= 2 pi, then fint_central's two recurrences, python negative-index
        // wrap-around included (utils/misc.py:96-104).  The column is its own scratch.
        auto X = [&](int t) {
            float v = den(t);
            if (t >= 1 && t < T - 1) v *= ACIDS_2PI_F;
            return v;
        };
        auto O = [&](int t) -> float& { return out[(int64_t)((t % T + T) % T) * nb]; };
        for (int t = 0; t < T; ++t) O(t) = 0.f;
        O(0) = X(0);
        O(T - 1) = X(T - 1);
        for (int i = 2; i < T; i += 2) O(i) = O(i - 2) + 4.f * X(i - 1);
        for (int i = T - 1; i > 0; i -= 2) O(i - 2) = O(i) - 4.f * X(i - 1);
    }
}

// four bins per thread: 16-byte loads of mag and phase, two 16-byte streaming stores.  Library sincosf here: this
// kernel is memory bound either way and measured 9 % SLOWER with polar_sincos (0.54 -> 0.59 ms at the cfg-4 shape)
__global__ void __launch_bounds__(256) polar_to_complex_kernel(const float* __restrict__ mag, const float* __restrict__ phase, int64_t n,
                                                               float2* __restrict__ out) {
    const bool vec = ((reinterpret_cast<uintptr_t>(mag) | reinterpret_cast<uintptr_t>(phase) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(mag) + i);
        const float4 p = __ldg(reinterpret_cast<const float4*>(phase) + i);
        float s0, c0, s1, c1, s2, c2, s3, c3;
        sincosf(p.x, &s0, &c0);
        sincosf(p.y, &s1, &c1);
        sincosf(p.z, &s2, &c2);
        sincosf(p.w, &s3, &c3);
        float4* o = reinterpret_cast<float4*>(out + 4 * i);
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(m.x * c0), "f"(m.x * s0), "f"(m.y * c1), "f"(m.y * s1) : "memory");
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 1), "f"(m.z * c2), "f"(m.z * s2), "f"(m.w * c3), "f"(m.w * s3) : "memory");
    }
    for (int64_t i = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float s, c;
        sincosf(__ldg(phase + i), &s, &c);
        const float m = __ldg(mag + i);
        stg_stream2(out + i, m * c, m * s);
    }
}

// One fast-Griffin-Lim update (torchaudio functional.py:336-350, called by stft.py:174-178):
//   a = rebuilt - mom * tprev;  X = mag * a / (|a| + 1e-16)
// One pass instead of six eager complex kernels: 20 B read + 8 B written per bin.  The caller keeps `rebuilt` as the
// next iteration's `tprev` (a pointer swap, nothing is copied).
__global__ void __launch_bounds__(256) gl_update_kernel(const float4* __restrict__ rebuilt, const float4* __restrict__ tprev,
                                                        const float2* __restrict__ mag, float mom, int64_t n_pairs,
                                                        float4* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 r = __ldg(rebuilt + i), t = __ldg(tprev + i);
        const float2 m = __ldg(mag + i);
        const float ax = r.x - mom * t.x, ay = r.y - mom * t.y, bx = r.z - mom * t.z, by = r.w - mom * t.w;
        const float ga = m.x / (sqrtf(ax * ax + ay * ay) + 1e-16f), gb = m.y / (sqrtf(bx * bx + by * by) + 1e-16f);
        float4 o = make_float4(ax * ga, ay * ga, bx * gb, by * gb);
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(out + i), "f"(o.x), "f"(o.y), "f"(o.z), "f"(o.w) : "memory");
    }
}

static int check_band(const acids_band& band) {
    ACIDS_REQUIRE(!band.meta || (band.coef && band.n_out > 0 && (band.coef_len & 31) == 0), ACIDS_EINVAL,
                  "malformed banded matrix (n_out=%d coef_len=%d)", band.n_out, band.coef_len);
    return ACIDS_OK;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int acids_mag_epilogue(const float* X, int64_t rows, int n_bins, acids_band band, int contrast, float eps,
                                  const float* offset, const float* scale, int drop_first, float* out,
                                  int64_t out_row_stride, void* stream) {
    ACIDS_REQUIRE(X && out, ACIDS_EINVAL, "mag_epilogue: NULL pointer");
    ACIDS_REQUIRE(rows >= 0 && n_bins > 0 && n_bins < 65536, ACIDS_EINVAL, "mag_epilogue: bad sizes");
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    int rc = check_band(band);
    if (rc) return rc;
    if (rows == 0) return ACIDS_OK;
    MagParams p{};
    p.X = reinterpret_cast<const float2*>(X); p.rows = rows; p.n_bins = n_bins;
    // rows per CTA iteration: the largest for which a lane's share of a row (n_bins / (32 * warps per row)) fits the 17
    // prefetch registers; beyond 4352 bins the unpipelined 4-row kernel
    const bool pf = n_bins <= 4352;
    const int threads = (pf && n_bins > 1088) ? 512 : 256;
    const int rows_per_iter = n_bins <= 544 ? 8 : (n_bins <= 2176 ? 4 : (pf ? 2 : 4));
    const size_t rows_bytes = (size_t)rows_per_iter * ((n_bins + 3) & ~3) * sizeof(float);
    rc = fill_epilogue(p.ep, band, n_bins, contrast, eps, drop_first, rows_bytes <= 34 * 1024 ? 40 * 1024 : 24 * 1024);
    if (rc) return rc;
    p.offset_ptr = offset; p.scale_ptr = scale; p.out = out; p.out_row_stride = out_row_stride;
    const size_t smem = rows_bytes + p.ep.band_bytes_meta + p.ep.band_bytes_coef;
    const int bsel = !band.meta ? BAND_NONE : (p.ep.band_bytes_meta > 0 ? BAND_SMEM : BAND_GLOBAL);
    void (*kern)(const MagParams) = nullptr;
#define ACIDS_MAGK(R, P, NT)                                                                                            \
    kern = bsel == BAND_NONE ? mag_epilogue_kernel<BAND_NONE, R, P, NT>                                                 \
                             : (bsel == BAND_SMEM ? mag_epilogue_kernel<BAND_SMEM, R, P, NT> : mag_epilogue_kernel<BAND_GLOBAL, R, P, NT>)
    if (!pf) { ACIDS_MAGK(4, false, 256); }
    else if (rows_per_iter == 8) { ACIDS_MAGK(8, true, 256); }
    else if (threads == 256) { ACIDS_MAGK(4, true, 256); }
    else if (rows_per_iter == 4) { ACIDS_MAGK(4, true, 512); }
    else { ACIDS_MAGK(2, true, 512); }
#undef ACIDS_MAGK
    ACIDS_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)) == cudaSuccess,
                  ACIDS_ECUDA, "mag_epilogue: cannot reserve %zu B of shared memory", smem);
    int64_t grid = (rows + rows_per_iter - 1) / rows_per_iter;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("mag_epilogue");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_mag_invert(const float* y, int64_t rows, int n_in, int64_t y_row_stride, int pad_last,
                                acids_band inverse_band, int contrast, float eps, const float* offset,
                                const float* scale, float* out, void* stream) {
    ACIDS_REQUIRE(y && out, ACIDS_EINVAL, "mag_invert: NULL pointer");
    ACIDS_REQUIRE(rows >= 0 && n_in > 0 && n_in < 65535 && (pad_last == 0 || pad_last == 1), ACIDS_EINVAL, "mag_invert: bad sizes");
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    int rc = check_band(inverse_band);
    if (rc) return rc;
    if (rows == 0) return ACIDS_OK;
    MagInvParams p{};
    p.y = y; p.rows = rows; p.n_in = n_in; p.y_row_stride = y_row_stride; p.pad_last = pad_last;
    rc = fill_epilogue(p.ep, inverse_band, n_in + pad_last, ACIDS_CONTRAST_NONE, eps, 0, 48 * 1024);
    if (rc) return rc;
    p.n_out = p.ep.n_cols;
    p.contrast = contrast; p.eps = eps; p.offset_ptr = offset; p.scale_ptr = scale; p.out = out;
    const int rows_per_iter = n_in > 1024 ? 4 : 8;
    const size_t rows_bytes = (size_t)rows_per_iter * ((n_in + pad_last + 3) & ~3) * sizeof(float);
    const size_t smem = rows_bytes + p.ep.band_bytes_meta + p.ep.band_bytes_coef;
    ACIDS_REQUIRE(smem <= 227 * 1024, ACIDS_ENOTSUP, "mag_invert: %d bins exceed shared memory", n_in);
    void (*kern)(const MagInvParams);
    if (rows_per_iter == 4)
        kern = !inverse_band.meta ? mag_invert_kernel<BAND_NONE, 4>
                                  : (p.ep.band_bytes_meta > 0 ? mag_invert_kernel<BAND_SMEM, 4> : mag_invert_kernel<BAND_GLOBAL, 4>);
    else
        kern = !inverse_band.meta ? mag_invert_kernel<BAND_NONE, 8>
                                  : (p.ep.band_bytes_meta > 0 ? mag_invert_kernel<BAND_SMEM, 8> : mag_invert_kernel<BAND_GLOBAL, 8>);
    ACIDS_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)) == cudaSuccess,
                  ACIDS_ECUDA, "mag_invert: cannot reserve %zu B of shared memory", smem);
    int64_t grid = (rows + rows_per_iter - 1) / rows_per_iter;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    kern<<<(unsigned)grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("mag_invert");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_phase_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, int mode, int if_method,
                               int weighted, const float* offset, const float* scale, int drop_first, float* out,
                               int64_t out_clip_stride, int64_t out_row_stride, void* stream) {
    ACIDS_REQUIRE(X && out, ACIDS_EINVAL, "phase_fwd: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 31) && n_bins > 0, ACIDS_EINVAL, "phase_fwd: bad sizes");
    ACIDS_REQUIRE(mode >= 0 && mode <= 2 && if_method >= 0 && if_method <= 2, ACIDS_EINVAL, "phase_fwd: bad mode/method");
    ACIDS_REQUIRE(!(mode == ACIDS_PHASE_IF && if_method == ACIDS_IF_CENTRAL && n_frames < 2), ACIDS_EINVAL,
                  "phase_fwd: central differences need at least 2 frames");
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    if (B == 0) return ACIDS_OK;
    PhaseParams p{};
    p.X = reinterpret_cast<const float2*>(X); p.B = B; p.n_frames = (int)n_frames; p.n_bins = n_bins;
    p.mode = mode; p.method = if_method; p.weighted = (mode == ACIDS_PHASE_IF) ? weighted : 0;
    p.offset_ptr = offset; p.scale_ptr = scale; p.drop_first = drop_first; p.out = out;
    p.out_clip_stride = out_clip_stride; p.out_row_stride = out_row_stride;
    const int64_t cols = B * n_bins;
    pick_phase_kernel(mode, if_method, p.weighted)<<<(unsigned)((cols + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("phase_fwd");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_polar_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, int contrast, float eps,
                               const float* mag_offset, const float* mag_scale, int mode, int if_method, int weighted,
                               const float* ph_offset, const float* ph_scale, int drop_first, float* mag_out,
                               int64_t mag_clip_stride, int64_t mag_row_stride, float* ph_out, int64_t ph_clip_stride,
                               int64_t ph_row_stride, void* stream) {
    ACIDS_REQUIRE(X && mag_out && ph_out, ACIDS_EINVAL, "polar_fwd: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 31) && n_bins > 0, ACIDS_EINVAL, "polar_fwd: bad sizes");
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(mode >= 0 && mode <= 2 && if_method >= 0 && if_method <= 2, ACIDS_EINVAL, "polar_fwd: bad mode/method");
    ACIDS_REQUIRE(!(mode == ACIDS_PHASE_IF && if_method == ACIDS_IF_CENTRAL && n_frames < 2), ACIDS_EINVAL,
                  "polar_fwd: central differences need at least 2 frames");
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    if (B == 0) return ACIDS_OK;
    PhaseParams p{};
    p.X = reinterpret_cast<const float2*>(X); p.B = B; p.n_frames = (int)n_frames; p.n_bins = n_bins;
    p.mode = mode; p.method = if_method; p.weighted = (mode == ACIDS_PHASE_IF) ? weighted : 0;
    p.offset_ptr = ph_offset; p.scale_ptr = ph_scale; p.drop_first = drop_first; p.out = ph_out;
    p.out_clip_stride = ph_clip_stride; p.out_row_stride = ph_row_stride;
    p.mag_out = mag_out; p.mag_clip_stride = mag_clip_stride; p.mag_row_stride = mag_row_stride;
    p.contrast = contrast; p.eps = eps; p.mag_offset_ptr = mag_offset; p.mag_scale_ptr = mag_scale;
    const int64_t cols = B * n_bins;
    pick_phase_kernel(mode, if_method, p.weighted, contrast)<<<(unsigned)((cols + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("polar_fwd");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_polar_rows_fwd(const float* X, int64_t B, int64_t n_frames, int n_bins, acids_band band, int contrast,
                                    float eps, const float* mag_offset, const float* mag_scale, int phase_mode, int if_method,
                                    int weighted, const float* ph_offset, const float* ph_scale, int drop_first, float* mag_out,
                                    int64_t mag_row_stride, float* ph_out, int64_t ph_row_stride, void* stream) {
    ACIDS_REQUIRE(X && mag_out && ph_out, ACIDS_EINVAL, "polar_rows_fwd: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 31) && n_bins > 0 && n_bins < 65536, ACIDS_EINVAL, "polar_rows_fwd: bad sizes");
    ACIDS_REQUIRE(contrast >= 0 && contrast <= 3, ACIDS_EINVAL, "unknown contrast id %d", contrast);
    ACIDS_REQUIRE(drop_first == 0 || drop_first == 1, ACIDS_EINVAL, "drop_first must be 0 or 1");
    // the phase modes that need no scan over the frames; everything else stays on acids_mag_epilogue + acids_phase_fwd
    ACIDS_REQUIRE(phase_mode == ACIDS_PHASE_RAW || (phase_mode == ACIDS_PHASE_IF && if_method == ACIDS_IF_FORWARD), ACIDS_ENOTSUP,
                  "polar_rows_fwd: only the raw phase and the forward-difference IF are fused (mode %d, method %d)", phase_mode, if_method);
    ACIDS_REQUIRE(n_bins <= 4352, ACIDS_ENOTSUP, "polar_rows_fwd: rows of %d bins do not fit the row-tile kernel (max 4352)", n_bins);
    int rc = check_band(band);
    if (rc) return rc;
    if (B == 0) return ACIDS_OK;
    PolarRowsParams p{};
    p.X = reinterpret_cast<const float2*>(X); p.rows = B * n_frames; p.n_frames = (int)n_frames; p.n_bins = n_bins;
    // (rows per iteration, threads): a lane's share of a row must fit the 17 prefetch registers.  Rows of 1089 .. 2176 bins run
    // 2-row tiles in 256-thread CTAs (3 CTAs / SM) rather than mag_epilogue's 4 rows x 512 threads: the arctangents make a
    // 512-thread CTA need 128 registers, i.e. ONE CTA per SM whose phases (stage, project, emit) cannot overlap (1.22 vs ... ms)
#ifndef ACIDS_POLAR_ROWS_MID
#define ACIDS_POLAR_ROWS_MID 2
#endif
    const int threads = (n_bins > 2176 || (n_bins > 1088 && ACIDS_POLAR_ROWS_MID == 4)) ? 512 : 256;
    const int rows_per_iter = n_bins <= 544 ? 8 : (n_bins <= 1088 ? 4 : (n_bins <= 2176 ? ACIDS_POLAR_ROWS_MID : 2));
    const size_t stride = (size_t)((n_bins + 3) & ~3);
    const size_t rows_bytes = (size_t)(2 * rows_per_iter + 1) * stride * sizeof(float);
    rc = fill_epilogue(p.ep, band, n_bins, contrast, eps, drop_first, rows_bytes <= 80 * 1024 ? 40 * 1024 : 24 * 1024);
    if (rc) return rc;
    ACIDS_REQUIRE(p.ep.n_cols == n_bins, ACIDS_EINVAL, "stack expects each tensor to be equal size, but got [%d] and [%d] bins",
                  p.ep.n_cols - drop_first, n_bins - drop_first);
    p.offset_ptr = mag_offset; p.scale_ptr = mag_scale; p.out = mag_out; p.out_row_stride = mag_row_stride;
    p.ph_out = ph_out; p.ph_row_stride = ph_row_stride; p.ph_offset_ptr = ph_offset; p.ph_scale_ptr = ph_scale;
    p.ph_mode = phase_mode; p.ph_weighted = phase_mode == ACIDS_PHASE_IF ? weighted : 0;
    const size_t smem = rows_bytes + p.ep.band_bytes_meta + p.ep.band_bytes_coef;
    const int bsel = !band.meta ? BAND_NONE : (p.ep.band_bytes_meta > 0 ? BAND_SMEM : BAND_GLOBAL);
    void (*kern)(const PolarRowsParams) = nullptr;
#define ACIDS_PRK(R, NT)                                                                                      \
    kern = bsel == BAND_NONE ? polar_rows_kernel<BAND_NONE, R, NT>                                            \
                             : (bsel == BAND_SMEM ? polar_rows_kernel<BAND_SMEM, R, NT> : polar_rows_kernel<BAND_GLOBAL, R, NT>)
    if (rows_per_iter == 8) { ACIDS_PRK(8, 256); }
    else if (threads == 256 && rows_per_iter == 4) { ACIDS_PRK(4, 256); }
    else if (threads == 256) { ACIDS_PRK(2, 256); }
    else if (rows_per_iter == 4) { ACIDS_PRK(4, 512); }
    else { ACIDS_PRK(2, 512); }
#undef ACIDS_PRK
    ACIDS_REQUIRE(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)) == cudaSuccess,
                  ACIDS_ECUDA, "polar_rows_fwd: cannot reserve %zu B of shared memory", smem);
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, threads, smem);
    if (nb < 1) nb = 1;
    const int64_t tiles = (p.rows + rows_per_iter - 1) / rows_per_iter;
    int64_t grid = (int64_t)num_sms() * nb;
    if (grid > tiles) grid = tiles;
    kern<<<(unsigned)grid, threads, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("polar_rows_fwd");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_phase_inv(const float* y, int64_t B, int64_t n_frames, int n_in, int64_t y_clip_stride,
                               int64_t y_row_stride, int pad_last, int mode, int if_method, const float* offset,
                               const float* scale, float* out, void* stream) {
    ACIDS_REQUIRE(y && out, ACIDS_EINVAL, "phase_inv: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 31) && n_in > 0, ACIDS_EINVAL, "phase_inv: bad sizes");
    ACIDS_REQUIRE(mode >= 0 && mode <= 2 && if_method >= 0 && if_method <= 2, ACIDS_EINVAL, "phase_inv: bad mode/method");
    ACIDS_REQUIRE(pad_last == 0 || pad_last == 1, ACIDS_EINVAL, "pad_last must be 0 or 1");
    if (B == 0) return ACIDS_OK;
    PhaseInvParams p{};
    p.y = y; p.B = B; p.n_frames = (int)n_frames; p.n_in = n_in; p.y_clip_stride = y_clip_stride;
    p.y_row_stride = y_row_stride; p.pad_last = pad_last; p.mode = mode; p.method = if_method;
    p.offset_ptr = offset; p.scale_ptr = scale; p.out = out;
    const int64_t cols = B * (n_in + pad_last);
    phase_inv_kernel<false><<<(unsigned)((cols + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("phase_inv");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_phase_inv_polar(const float* y, int64_t B, int64_t n_frames, int n_in, int64_t y_clip_stride,
                                     int64_t y_row_stride, int pad_last, int mode, int if_method, const float* offset,
                                     const float* scale, const float* mag, float* out, void* stream) {
    ACIDS_REQUIRE(y && mag && out, ACIDS_EINVAL, "phase_inv_polar: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 31) && n_in > 0, ACIDS_EINVAL, "phase_inv_polar: bad sizes");
    ACIDS_REQUIRE(mode >= 0 && mode <= 2 && if_method >= 0 && if_method <= 2, ACIDS_EINVAL, "phase_inv_polar: bad mode/method");
    ACIDS_REQUIRE(!(mode == ACIDS_PHASE_IF && if_method == ACIDS_IF_CENTRAL), ACIDS_EINVAL,
                  "phase_inv_polar: the central method needs the phase column as scratch; use phase_inv + polar_to_complex");
    ACIDS_REQUIRE(pad_last == 0 || pad_last == 1, ACIDS_EINVAL, "pad_last must be 0 or 1");
    ACIDS_REQUIRE((reinterpret_cast<uintptr_t>(out) & 7) == 0, ACIDS_EINVAL, "phase_inv_polar: out must be 8-byte aligned");
    if (B == 0) return ACIDS_OK;
    PhaseInvParams p{};
    p.y = y; p.B = B; p.n_frames = (int)n_frames; p.n_in = n_in; p.y_clip_stride = y_clip_stride;
    p.y_row_stride = y_row_stride; p.pad_last = pad_last; p.mode = mode; p.method = if_method;
    p.offset_ptr = offset; p.scale_ptr = scale; p.out = out; p.mag = mag;
    const int64_t cols = B * (n_in + pad_last);
    phase_inv_kernel<true><<<(unsigned)((cols + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("phase_inv_polar");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_polar_to_complex(const float* mag, const float* phase, int64_t n, float* out, void* stream) {
    ACIDS_REQUIRE(mag && phase && out, ACIDS_EINVAL, "polar_to_complex: NULL pointer");
    ACIDS_REQUIRE(n >= 0, ACIDS_EINVAL, "polar_to_complex: negative size");
    if (n == 0) return ACIDS_OK;
    int64_t grid = ((n + 3) / 4 + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    polar_to_complex_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mag, phase, n,
                                                                                           reinterpret_cast<float2*>(out));
    ACIDS_CHECK_LAUNCH("polar_to_complex");
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_griffinlim_update(const float* rebuilt, const float* tprev, const float* mag, float momentum,
                                       int64_t n, float* out, void* stream) {
    ACIDS_REQUIRE(rebuilt && tprev && mag && out, ACIDS_EINVAL, "griffinlim_update: NULL pointer");
    ACIDS_REQUIRE(n >= 0 && (n & 1) == 0, ACIDS_EINVAL, "griffinlim_update: the number of bins must be even (pad the caller's view)");
    ACIDS_REQUIRE(((reinterpret_cast<uintptr_t>(rebuilt) | reinterpret_cast<uintptr_t>(tprev) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(mag) & 7) == 0,
                  ACIDS_EINVAL, "griffinlim_update: buffers must be 16-byte aligned");
    if (n == 0) return ACIDS_OK;
    const int64_t pairs = n / 2;
    int64_t grid = (pairs + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    gl_update_kernel<<<(unsigned)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const float4*>(rebuilt), reinterpret_cast<const float4*>(tprev), reinterpret_cast<const float2*>(mag),
        momentum / (1.0f + momentum), pairs, reinterpret_cast<float4*>(out));
    ACIDS_CHECK_LAUNCH("griffinlim_update");
    return ACIDS_OK;
}
