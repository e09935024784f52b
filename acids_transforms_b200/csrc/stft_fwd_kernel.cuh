// stft_fwd_kernel.cuh — the fused forward kernel template, its launch logic and the per-plan entry points.
// stft_fwd_plan.cu instantiates it once per FFT plan (one translation unit each, built in parallel);
// stft_fwd.cu holds the C ABI and dispatches on n_fft.
#pragma once
#include "common.cuh"
#include "plans.cuh"

namespace acids {

// MODE_POLAR = MODE_REAL plus a second output: the raw phase or the forward-difference instantaneous frequency of every
// bin (Polar / PolarIF right after the STFT, spectral_repr.py:431-440) — the spectrum never reaches HBM
// MODE_STATS: nothing is stored; min / max / sum / sum of squares of contrast(|X|) over every bin go to one StatAcc per CTA
// (Magnitude.scale_data on the spectrum, spectral_repr.py:242-245 + norm.py:26-38, without materialising it)
enum { MODE_COMPLEX = 0, MODE_REAL = 1, MODE_POLAR = 2, MODE_STATS = 3 };
enum { VAR_COMPLEX = 0, VAR_MAG_NOBAND, VAR_MAG_SMEM, VAR_MAG_GLOBAL, VAR_MEL_POWER_SMEM, VAR_MEL_POWER_GLOBAL, VAR_MEL_ANY_SMEM,
       VAR_MEL_ANY_GLOBAL, VAR_POLAR_NOBAND, VAR_POLAR_SMEM, VAR_POLAR_GLOBAL, VAR_STATS };

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
    int vec_ok;      // clip rows and frame starts are aligned for the first pass's vector loads (bit 0: 8 B, bit 1: 16 B)
    int band_smem_bytes;   // shared memory reserved for the banded matrix (16-byte multiple; 0: read it from global)
    // raw-domain prologue folded into the loads: x is [B / 2, 2, L] stereo and clip 2 b + c is channel c of
    // MidSide.forward (raw.py:145-161): mid = (l + r) / 2 [/ sqrt 2 when midside == 2], side = (l - r) / 2
    int midside;
    // MODE_POLAR: the phase output (rows like `out`), its normalisation, ACIDS_PHASE_RAW or ACIDS_PHASE_IF (forward differences)
    float* ph_out;
    int64_t ph_clip_stride, ph_row_stride;
    const float* ph_offset_ptr;
    const float* ph_scale_ptr;
    int ph_mode, ph_weighted;
    StatAcc* stat_part;      // MODE_STATS: one partial per CTA (gridDim.x of them)
};

// launch shape per plan: small frame groups run 128-thread CTAs at 4 CTAs / SM (<= 128 registers)
#ifndef ACIDS_FWD_MINB_SMALL
#define ACIDS_FWD_MINB_SMALL 4
#endif
#ifndef ACIDS_FWD_RIE
#define ACIDS_FWD_RIE 1
#endif
#ifndef ACIDS_FWD_FPW_1024
#define ACIDS_FWD_FPW_1024 1           // 2: two frames in flight per warp at 168 registers, measured SLOWER (1.35 vs 1.16 ms fused)
#endif
#ifndef ACIDS_FWD_THREADS_1024
#define ACIDS_FWD_THREADS_1024 256     // T = 32 plan: 8 frames per unit, the epilogue amortises a column's metadata / coefficients over 8 rows
#endif
#ifndef ACIDS_FWD_THREADS_1024R
#define ACIDS_FWD_THREADS_1024R 128    // T = 16 plan: 8 frames per unit in 4 warps; 4 CTAs / SM (256 threads x 2: 1.21 vs 1.06 ms, barrier bound)
#endif
template <class P, int MODE>
struct FwdCfg {
#ifndef ACIDS_FWD_MID_THREADS
#define ACIDS_FWD_MID_THREADS 128      // CTA size of the T = 64 / 128 plans (n_fft 2048 / 4096): 3 CTAs x 168 registers, no spills
#endif
#ifndef ACIDS_FWD_MID_MINB
#define ACIDS_FWD_MID_MINB 3
#endif
    // FPW: frames a thread group keeps in flight.  With two, every phase of the transform (exchange loads, butterflies,
    // exchange stores) has two independent instruction streams per warp, the window taps are read from shared memory once
    // for both, and the epilogue amortises a column's metadata / coefficients / dispatch over twice the rows.  Costs
    // registers: 168 per thread, 3 CTAs / SM (n_fft = 1024: 1.17 -> see DESIGN.md section 5).
    static constexpr int FPW = P::N == 1024 ? ACIDS_FWD_FPW_1024 : 1;
    static constexpr int THREADS = (P::N == 1024 && MODE == 1 /* MODE_REAL */) ? (P::T == 16 ? ACIDS_FWD_THREADS_1024R : ACIDS_FWD_THREADS_1024)
                                                : (P::T <= 32 ? 128 : (P::T > 256 ? P::T : (P::T <= 128 ? ACIDS_FWD_MID_THREADS : 256)));
    // MODE_POLAR (2), small plans: 168 registers (3 CTAs / SM) — the arctangents next to the |X| rows spill 100+ bytes at 128
    static constexpr int MINB = (FPW > 1 || (MODE == 2 && P::T <= 32)) ? 3 : (P::T <= 32 ? ACIDS_FWD_MINB_SMALL * 128 / THREADS : (P::T <= 128 ? ACIDS_FWD_MID_MINB : (P::T <= 256 ? 2 : 1)));
    // complex output: no epilogue to hide the next frame's loads behind, so they are issued a whole FFT early into a
    // second register set; that needs ~160 registers -> one CTA less per SM for the small plans
    static constexpr int MINB_COMPLEX = P::T <= 32 ? 3 : MINB;
#ifndef ACIDS_FWD_LAG
#define ACIDS_FWD_LAG 0      // measured slower (1.32 vs 1.15 ms on the two-exchange plan): see LAG in the kernel
#endif
    static constexpr bool LAG = ACIDS_FWD_LAG && MODE == 1 /* MODE_REAL */ && FPW == 1 && !(ACIDS_FWD_RIE && P::N == 1024 && P::T == 16);
    static constexpr int GT = THREADS / P::T;                   // thread groups per CTA
    static constexpr int G = GT * FPW;                          // frames per unit
    static constexpr int NF = (P::N == 1024 && G % 8 == 0) ? 8 : (G < 4 ? G : 4);   // rows per epilogue tile
    // FPW > 1: the |X| row of a frame is parked in the frame's own exchange buffer (free once the last pass has read
    // it) instead of a separate double-buffered tile: 51 KB instead of 84 KB per CTA, three CTAs fit an SM
    // the 16-thread n_fft = 1024 plan: two frames per warp double the exchange and row buffers per warp; with separate
    // double-buffered rows only 2 CTAs fit an SM (1.22 ms), with the rows parked in the exchange buffers 4 do (1.00 ms)
    static constexpr bool ROWS_IN_EXCH = FPW > 1 || (ACIDS_FWD_RIE && P::N == 1024 && P::T == 16 && MODE == 1);
    static constexpr int VSTR = ROWS_IN_EXCH ? 2 * P::SMEM_CF : ((P::F + 3) & ~3);   // |X| row stride in shared memory (floats)
    static constexpr int VW = (P::bpt(0) % 2 == 0) ? 4 : 2;     // floats per vector load of the first pass
    static_assert(G % NF == 0, "rows per CTA must be a multiple of the row tile");
    static_assert(!ROWS_IN_EXCH || 2 * P::SMEM_CF >= P::F, "a row must fit the exchange buffer");
    static constexpr size_t exch_bytes() { return (size_t)G * P::SMEM_CF * sizeof(cf); }
    static constexpr size_t win_bytes() { return (size_t)P::M * sizeof(float2); }
};

static inline size_t round16(size_t n) { return (n + 15) & ~(size_t)15; }

// |X|^power from the squared magnitude; PMODE: 1 -> magnitude, 2 -> power spectrum, 0 -> general exponent
template <int PMODE>
__device__ __forceinline__ float pow_value(cf a, float power) {
    const float p2 = fmaf(a.x, a.x, a.y * a.y);
    if (PMODE == 1) return fast_sqrt(p2);
    if (PMODE == 2) return p2;
    return powf(fast_sqrt(p2), power);
}

// CSEL: contrast known at compile time (ACIDS_CONTRAST_*) or -1 (dispatched once per row tile)
// MS: the MidSide prologue is folded into the sample loads (a compile-time switch: as a run-time one it cost every variant
// registers — 16..120 bytes of spills at 128 registers, the headline kernel 1.11 -> 1.44 ms)
template <class P, int MODE, int PMODE, int CSEL, int BAND, bool TRANSPOSED, bool MS = false>
__global__ void __launch_bounds__(FwdCfg<P, MODE>::THREADS, MODE == MODE_COMPLEX ? FwdCfg<P, MODE>::MINB_COMPLEX : FwdCfg<P, MODE>::MINB)
    stft_fwd_kernel(const FwdParams p) {
    using C = FwdCfg<P, MODE>;
    constexpr int THREADS = C::THREADS;
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V, G = C::G, NF = C::NF, VSTR = C::VSTR, VW = C::VW, FPW = C::FPW;
    constexpr int R0 = P::radix(0), B0 = P::bpt(0), NB0 = P::nb(0);
    constexpr bool REALISH = MODE == MODE_REAL || MODE == MODE_POLAR, POLAR = MODE == MODE_POLAR, STATS = MODE == MODE_STATS;
    constexpr bool RIE = REALISH && C::ROWS_IN_EXCH;
    static_assert(!POLAR || !RIE, "the polar epilogue keeps its rows in their own buffers");
    // LAG: the epilogue of unit u runs AFTER the transform of unit u + 1, and the CTA barrier between "rows written" and
    // "rows read" becomes a split mbarrier arrive / wait one whole FFT apart: nobody actually waits unless a warp falls a
    // full unit behind (ncu before: 1.37 stall cycles per issue on the barrier, the warps of a CTA moved in lock-step).
    constexpr bool LAG = C::LAG;
    __shared__ uint64_t lag_full[2], lag_empty[2];
    using FFT = FrameFFT<P, false>;
    using PR = typename FFT::PR;
    constexpr int RP = PR::R, NBP = PR::NB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    // exchange buffer of frame f of this group: frames g * FPW + f of a unit
    cf* const s0 = reinterpret_cast<cf*>(smem_raw) + (size_t)g * FPW * P::SMEM_CF;
    auto gsync = [&]() { group_sync<T, THREADS>(g); };

    // ---- frame-invariant state: twiddles in registers; the analysis window as (w[2n], w[2n+1]) / 2 pairs in shared
    //      memory (1/2 = the factor of the even/odd split); the banded matrix; the normalisation constants ----
    FFT fft;
    fft.init(tid);
    float2* const swin = reinterpret_cast<float2*>(smem_raw + C::exch_bytes());
    for (int n = threadIdx.x; n < M; n += THREADS)
        swin[n] = make_float2(0.5f * __ldg(p.window + 2 * n), 0.5f * __ldg(p.window + 2 * n + 1));
    EpiArgs ea{};
    float* vrows = nullptr;
    float* prows = nullptr;      // POLAR: raw phase rows of the unit
    float* pcarry = nullptr;     // POLAR: last phase row of the previous units (rotating)
    if (REALISH) {
        unsigned char* bandmem = smem_raw + C::exch_bytes() + C::win_bytes();
        int32_t* smeta = reinterpret_cast<int32_t*>(bandmem);
        float* scoef = reinterpret_cast<float*>(bandmem + p.ep.band_bytes_meta);
        bool split = false;
        if (BAND == BAND_SMEM) {
            split = TRANSPOSED && band_wants_split(p.ep);
            if (split) stage_band_split(p.ep, smeta, scoef);
            else stage_band(p.ep, smeta, scoef);
        }
        ea = make_epi_args(p.ep, BAND == BAND_SMEM ? smeta : p.ep.meta, BAND == BAND_SMEM ? scoef : p.ep.coef, p.offset_ptr,
                           p.scale_ptr);
        ea.split = split ? 1 : 0;
        // |X| rows of a unit: in the frames' exchange buffers (RIE), or double buffered so that the next unit's FFT never
        // waits for the slowest epilogue thread
        vrows = RIE ? reinterpret_cast<float*>(smem_raw) : reinterpret_cast<float*>(bandmem + (BAND == BAND_SMEM ? p.band_smem_bytes : 0));
        // POLAR: raw phase rows.  G == 1: three rotating rows (the previous frame's row is the previous unit's);
        // G > 1: double-buffered unit tiles like vrows plus three rotating carry rows holding the last row of a unit.
        // Three, not two: a thread that has left the epilogue of unit u writes unit u + 1's rows while a slower one may
        // still read unit u - 1's last row; by unit u + 2 everybody has passed the barrier of unit u + 1.
        if (POLAR) {
            prows = vrows + 2 * G * VSTR;
            pcarry = G == 1 ? prows : prows + 2 * G * VSTR;
        }
    }
    if (LAG && threadIdx.x == 0) {
        mbar_init(&lag_full[0], THREADS / 32);
        mbar_init(&lag_full[1], THREADS / 32);
        mbar_init(&lag_empty[0], THREADS / 32);
        mbar_init(&lag_empty[1], THREADS / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float ph_off = 0.f, ph_inv = 1.f;
    if (POLAR) {
        ph_off = p.ph_offset_ptr ? __ldg(p.ph_offset_ptr) : 0.f;
        ph_inv = p.ph_scale_ptr ? 1.0f / __ldg(p.ph_scale_ptr) : 1.0f;
    }
    __syncthreads();

    const int n_frames = (int)p.n_frames;
    const int upc = (n_frames + G - 1) / G;              // units per clip
    const int64_t total = p.B * upc;
    const int64_t u0 = total * blockIdx.x / gridDim.x, u1 = total * (blockIdx.x + 1) / gridDim.x;
    const int L = (int)p.L;                              // < 2^31 - 2 n_fft (checked by the host)
    // POLAR with frame differences: a CTA whose run starts inside a clip first runs the unit before it without
    // storing anything (a halo), so that the phase row of the frame preceding its first one is there (0.3 % extra work)
    const bool ph_if = POLAR && p.ph_mode == ACIDS_PHASE_IF;
    const int64_t ustart = (ph_if && u0 < u1 && (u0 % upc) != 0) ? u0 - 1 : u0;

    // Raw (un-windowed) samples of frame (b, t) in first-pass operand order.  Interior, aligned frames: vector
    // loads straight into registers.  Edge frames (torch.stft center=True, pad_mode="reflect") and unaligned
    // inputs: the group stages the frame in its exchange buffer with scalar loads, then reads its operands.
    // fetch_fast issues the vector loads when the frame allows it and says whether the frame still needs staging.
    // Frames past the end of a clip (the last unit of a clip is padded to G frames) are never stored: instead of
    // zero-filling their registers they re-read an interior frame of the same clip when there is one.
    const int t_safe = (p.pad + p.hop - 1) / p.hop;
    const bool has_safe = t_safe < n_frames && t_safe * p.hop - p.pad + N <= L;
    // MS_DEFER (MidSide + the early prefetch of the complex kernels): combining the two channels where they are loaded would
    // consume the loads at once and expose a full memory latency per frame (ncu, cfg 4: 26 % of all stall samples on that one
    // FFMA).  Instead the early fetch loads the first channel raw and only PREFETCHES the second into L1; ms_finish() loads it
    // (an L1 hit by then) and forms mid / side right before the frame is used.
    constexpr bool MS_DEFER = MS && MODE == MODE_COMPLEX && T <= 128 && FPW == 1;
    bool fin_on = false;                 // the prefetched frame still needs ms_finish
    const float* __restrict__ fin_x = nullptr;
    int fin_t = 0;
    float fin_s = 1.f, fin_g = 1.f;
    auto fetch_fast = [&](cf* v, const float* __restrict__ xb, int t, float ms_sign, float ms_gain) -> bool {
        const bool valid = t < n_frames;
        if (!valid && has_safe) t = t_safe;
#ifdef ACIDS_DEBUG_NOLOAD       // tuning experiment only: every frame reads the same (cached) samples
        const int s0i = N;
        xb = p.x;
#else
        const int s0i = t * p.hop - p.pad;
#endif
        const bool fast = (p.vec_ok & (VW == 4 ? 2 : 1)) && s0i >= 0 && s0i + N <= L;
        bool stage = valid && !fast;
        if (T < 32) stage = __any_sync(0xffffffffu, stage);     // frame groups sharing a warp take the same path
        if (!stage) {
            if (valid || (has_safe && fast)) {
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    // this thread's B0 butterflies are consecutive: B0 adjacent complex operands per radix slot
                    const float* __restrict__ src = xb + s0i + 2 * (tid * B0 + r * NB0);
                    if (VW == 4) {
#pragma unroll
                        for (int b0 = 0; b0 < B0; b0 += 2) {
                            float4 a = __ldg(reinterpret_cast<const float4*>(src) + (b0 >> 1));
                            if (MS_DEFER) {
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float4*>(src + p.ldx) + (b0 >> 1)));
                            } else if (MS) {
                                const float4 o = __ldg(reinterpret_cast<const float4*>(src + p.ldx) + (b0 >> 1));
                                a = make_float4(fmaf(ms_sign, o.x, a.x) * ms_gain, fmaf(ms_sign, o.y, a.y) * ms_gain,
                                                fmaf(ms_sign, o.z, a.z) * ms_gain, fmaf(ms_sign, o.w, a.w) * ms_gain);
                            }
                            v[b0 * R0 + r] = mk(a.x, a.y);
                            v[(b0 + 1) * R0 + r] = mk(a.z, a.w);
                        }
                    } else {
#pragma unroll
                        for (int b0 = 0; b0 < B0; ++b0) {
                            float2 a = __ldg(reinterpret_cast<const float2*>(src) + b0);
                            if (MS_DEFER) {
                                asm volatile("prefetch.global.L1 [%0];" ::"l"(reinterpret_cast<const float2*>(src + p.ldx) + b0));
                            } else if (MS) {
                                const float2 o = __ldg(reinterpret_cast<const float2*>(src + p.ldx) + b0);
                                a = make_float2(fmaf(ms_sign, o.x, a.x) * ms_gain, fmaf(ms_sign, o.y, a.y) * ms_gain);
                            }
                            v[b0 * R0 + r] = mk(a.x, a.y);
                        }
                    }
                }
                if (MS_DEFER) {
                    fin_on = true;
                    fin_x = xb;
                    fin_t = t;
                    fin_s = ms_sign;
                    fin_g = ms_gain;
                }
            } else {
#pragma unroll
                for (int i = 0; i < V; ++i) v[i] = mk(0.f, 0.f);
            }
        }
        return stage;
    };
    // second half of a deferred MidSide fetch: v holds the first channel's raw samples of frame fin_t
    auto ms_finish = [&](cf* v) {
        if (!MS_DEFER || !fin_on) return;
        fin_on = false;
        const int s0i = fin_t * p.hop - p.pad;
#pragma unroll
        for (int r = 0; r < R0; ++r) {
            const float* __restrict__ src = fin_x + p.ldx + s0i + 2 * (tid * B0 + r * NB0);
            if (VW == 4) {
#pragma unroll
                for (int b0 = 0; b0 < B0; b0 += 2) {
                    const float4 o = __ldg(reinterpret_cast<const float4*>(src) + (b0 >> 1));
                    v[b0 * R0 + r] = mk(fmaf(fin_s, o.x, v[b0 * R0 + r].x) * fin_g, fmaf(fin_s, o.y, v[b0 * R0 + r].y) * fin_g);
                    v[(b0 + 1) * R0 + r] = mk(fmaf(fin_s, o.z, v[(b0 + 1) * R0 + r].x) * fin_g, fmaf(fin_s, o.w, v[(b0 + 1) * R0 + r].y) * fin_g);
                }
            } else {
#pragma unroll
                for (int b0 = 0; b0 < B0; ++b0) {
                    const float2 o = __ldg(reinterpret_cast<const float2*>(src) + b0);
                    v[b0 * R0 + r] = mk(fmaf(fin_s, o.x, v[b0 * R0 + r].x) * fin_g, fmaf(fin_s, o.y, v[b0 * R0 + r].y) * fin_g);
                }
            }
        }
    };
    auto fetch_staged = [&](cf* v, cf* s, const float* __restrict__ xb, int t, float ms_sign, float ms_gain) {
        const bool valid = t < n_frames;
        const int s0i = t * p.hop - p.pad;
        float* sf = reinterpret_cast<float*>(s);
        gsync();   // the group is done reading the exchange buffer
        for (int i = tid; i < N; i += T) {
            int k = s0i + i;
            if (k < 0) k = -k;
            if (k >= L) k = 2 * (L - 1) - k;
            k = k < 0 ? 0 : (k >= L ? L - 1 : k);
            float smp = valid ? __ldg(xb + k) : 0.f;
            if (MS && valid) smp = fmaf(ms_sign, __ldg(xb + p.ldx + k), smp) * ms_gain;
            sf[i] = smp;
        }
        gsync();
#pragma unroll
        for (int b0 = 0; b0 < B0; ++b0)
#pragma unroll
            for (int r = 0; r < R0; ++r) v[b0 * R0 + r] = *reinterpret_cast<const cf*>(sf + 2 * fft.template in_index<0>(b0, r));
    };
    // MidSide folded into the loads: (sign, gain) of the clip being fetched — channel 0: (l + r) gain, channel 1: (l - r) / 2
    float ms_s = 1.f, ms_g = 1.f;
    auto fetch = [&](cf* v, cf* s, const float* __restrict__ xb, int t) {
        if (fetch_fast(v, xb, t, ms_s, ms_g)) fetch_staged(v, s, xb, t, ms_s, ms_g);
    };

    // clip and unit-in-clip advance incrementally: no division, no 64-bit multiply per frame
    int64_t b = ustart / upc;
    int uc = (int)(ustart - b * upc);
    // clip of the NEXT unit (the one being fetched); with MidSide: the LEFT row of the clip's stereo pair
    const float* __restrict__ xclip = MS ? p.x + (b >> 1) * 2 * p.ldx : p.x + b * p.ldx;
    int ms_ch = (int)(b & 1);
    const float ms_mid_gain = p.midside == 2 ? 0.35355339059327376220f : 0.5f;      // 1/2 or 1/(2 sqrt 2) (raw.py:151-155)
    auto ms_update = [&]() {
        if (MS) {
            ms_s = ms_ch ? -1.f : 1.f;
            ms_g = ms_ch ? 0.5f : ms_mid_gain;
        }
    };
    ms_update();
    float* __restrict__ oclip = MODE == MODE_COMPLEX ? p.out + b * ((int64_t)n_frames * P::F * 2) : p.out + b * p.out_clip_stride;
    const int64_t oclip_step = MODE == MODE_COMPLEX ? (int64_t)n_frames * P::F * 2 : p.out_clip_stride;
    int buf = 0;
    int pidx = 0;                    // POLAR: rotating index (0..2) of the carry row this unit writes
    cf v[FPW][V];
    bool pend[FPW];                  // RIE: frames of the next unit that must be staged once the rows have been consumed
    const float* __restrict__ pend_clip = xclip;
    float pend_s = 1.f, pend_g = 1.f;
#pragma unroll
    for (int f = 0; f < FPW; ++f) pend[f] = false;
    float* __restrict__ phclip = POLAR ? p.ph_out + b * p.ph_clip_stride : nullptr;
    StatAcc acc{DBL_MAX, -DBL_MAX, 0.0, 0.0};
    if (u0 < u1) {
#pragma unroll
        for (int f = 0; f < FPW; ++f) fetch(v[f], s0 + f * P::SMEM_CF, xclip, uc * G + g * FPW + f);
        ms_finish(v[0]);
    }
    // the row-tile epilogue of one unit: rows `vb` (shared memory) -> banded projection -> contrast -> normalise -> `outc`
    // halo (POLAR): the unit before the CTA's run — rows only, nothing is stored
    auto run_epilogue = [&](const float* __restrict__ vb, float* __restrict__ outc, int unit_in_clip, bool halo) {
        const int t0 = unit_in_clip * G;
        const int n_valid = halo ? 0 : min(G, n_frames - t0);
        const int rs = (int)p.out_row_stride, cs = (int)p.out_col_stride;
        float* out0 = outc + (TRANSPOSED ? t0 : t0 * rs);
        constexpr bool SPREAD = !TRANSPOSED;
#pragma unroll 1
        for (int g0 = 0; g0 < G; g0 += NF) {
            if (g0 >= n_valid) break;
            if (!TRANSPOSED && N == 1024 && rs == P::F)
                epilogue_dispatch<THREADS, NF, CSEL, BAND, TRANSPOSED, P::F, SPREAD>(p.ep.contrast, vb + g0 * VSTR, VSTR, threadIdx.x, ea,
                                                                                      out0 + g0 * rs, rs, cs, n_valid - g0);
            else
                epilogue_dispatch<THREADS, NF, CSEL, BAND, TRANSPOSED, 0, SPREAD>(p.ep.contrast, vb + g0 * VSTR, VSTR, threadIdx.x, ea,
                                                                                   out0 + (TRANSPOSED ? g0 : g0 * rs), rs, cs, n_valid - g0);
        }
        return n_valid;
    };
    int it = 0;                       // LAG: units this CTA has transformed so far
    float* __restrict__ lag_out = nullptr;
    int lag_uc = 0;
    for (int64_t u = ustart; u < u1; ++u) {
        const int tb = uc * G + g * FPW;             // first frame of this group in the unit
        const int cur_uc = uc;
        float* __restrict__ const cur_out = oclip;
        float* __restrict__ const cur_ph = phclip;
        if (++uc == upc) {
            uc = 0;
            if (MS) {
                ms_ch ^= 1;
                if (!ms_ch) xclip += 2 * p.ldx;
                ms_update();
            } else {
                xclip += p.ldx;
            }
            oclip += oclip_step;
            if (POLAR) phclip += p.ph_clip_stride;
        }

        if (RIE) {
            // the previous unit's rows (in the exchange buffers) have been consumed by every thread of the CTA before
            // anybody overwrites them: frames that need staging first, all others as late as the first exchange store
            bool any_pend = false;
#pragma unroll
            for (int f = 0; f < FPW; ++f) any_pend |= pend[f];
            if (any_pend) {
                __syncthreads();
#pragma unroll
                for (int f = 0; f < FPW; ++f)
                    if (pend[f]) fetch_staged(v[f], s0 + f * P::SMEM_CF, pend_clip, tb + f, pend_s, pend_g);
            }
            // ---- window (first-pass operand order, same adjacency as the loads) ----
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const float2* __restrict__ wv = swin + (tid * B0 + r * NB0);
                if (VW == 4) {
#pragma unroll
                    for (int b0 = 0; b0 < B0; b0 += 2) {
                        const float4 w = *reinterpret_cast<const float4*>(wv + b0);
#pragma unroll
                        for (int f = 0; f < FPW; ++f) {
                            v[f][b0 * R0 + r] = cmul2(v[f][b0 * R0 + r], mk(w.x, w.y));
                            v[f][(b0 + 1) * R0 + r] = cmul2(v[f][(b0 + 1) * R0 + r], mk(w.z, w.w));
                        }
                    }
                } else {
#pragma unroll
                    for (int b0 = 0; b0 < B0; ++b0) {
                        const float2 w = wv[b0];
#pragma unroll
                        for (int f = 0; f < FPW; ++f) v[f][b0 * R0 + r] = cmul2(v[f][b0 * R0 + r], mk(w.x, w.y));
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template butterflies<0>(v[f]);
            if (!any_pend) __syncthreads();
        } else {
            // ---- window (first-pass operand order, same adjacency as the loads) ----
#pragma unroll
            for (int r = 0; r < R0; ++r) {
                const float2* __restrict__ wv = swin + (tid * B0 + r * NB0);
                if (VW == 4) {
#pragma unroll
                    for (int b0 = 0; b0 < B0; b0 += 2) {
#ifdef ACIDS_DEBUG_NOWINDOW     // tuning experiment only: what would the kernel cost if the window taps were free?
                        const float4 w = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
#else
                        const float4 w = *reinterpret_cast<const float4*>(wv + b0);
#endif
#pragma unroll
                        for (int f = 0; f < FPW; ++f) {
                            v[f][b0 * R0 + r] = cmul2(v[f][b0 * R0 + r], mk(w.x, w.y));
                            v[f][(b0 + 1) * R0 + r] = cmul2(v[f][(b0 + 1) * R0 + r], mk(w.z, w.w));
                        }
                    }
                } else {
#pragma unroll
                    for (int b0 = 0; b0 < B0; ++b0) {
                        const float2 w = wv[b0];
#pragma unroll
                        for (int f = 0; f < FPW; ++f) v[f][b0 * R0 + r] = cmul2(v[f][b0 * R0 + r], mk(w.x, w.y));
                    }
                }
            }
        }

        // complex output, plans that run 3 CTAs / SM (168 registers, n_fft <= 4096): the next frame's samples are fetched a
        // whole FFT early into a second register set (n_fft = 4096: 1.09 -> 0.96 ms); larger plans (128 registers) fetch
        // after the stores instead; with two frames in flight the second register set IS the second frame
        constexpr bool EARLY = MODE == MODE_COMPLEX && T <= 128 && FPW == 1;
        cf nv[EARLY ? V : 1];
        if (EARLY) {
            if (u + 1 < u1) fetch(nv, s0, xclip, uc * G + g);       // lands during this frame's FFT
        }

        // ---- passes ----
        if (!RIE) {
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template butterflies<0>(v[f]);
            gsync();   // the previous frame's readers of s are done
        }
#pragma unroll
        for (int f = 0; f < FPW; ++f) fft.template store<0>(v[f], s0 + f * P::SMEM_CF);
        gsync();
#pragma unroll
        for (int f = 0; f < FPW; ++f) fft.template load<1>(v[f], s0 + f * P::SMEM_CF);
#pragma unroll
        for (int f = 0; f < FPW; ++f) fft.template butterflies<1>(v[f]);
        if constexpr (P::NP > 2) {
            gsync();
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template store<1>(v[f], s0 + f * P::SMEM_CF);
            gsync();
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template load<2>(v[f], s0 + f * P::SMEM_CF);
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template butterflies<2>(v[f]);
        }
        if constexpr (P::NP > 3) {
            gsync();
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template store<2>(v[f], s0 + f * P::SMEM_CF);
            gsync();
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template load<3>(v[f], s0 + f * P::SMEM_CF);
#pragma unroll
            for (int f = 0; f < FPW; ++f) fft.template butterflies<3>(v[f]);
        }
        if (RIE) gsync();    // every lane has read the last pass's operands: the buffers may take the |X| rows

        if (STATS) {
#pragma unroll
            for (int f = 0; f < FPW; ++f) {
                cf o1[V / 2], o2[V / 2], ex;
                fft.untangle_fwd(v[f], o1, o2, ex);
                // this thread's V bins of the frame (+ bin M/2 on thread 0): float partials per frame, double across frames
                float fmn = 3.4e38f, fmx = -3.4e38f, fs = 0.f, fs2 = 0.f;
#pragma unroll
                for (int i = 0; i < V / 2; ++i) {
                    const float a1 = apply_contrast(fast_sqrt(fmaf(o1[i].x, o1[i].x, o1[i].y * o1[i].y)), p.ep.contrast, p.ep.eps);
                    const float a2 = apply_contrast(fast_sqrt(fmaf(o2[i].x, o2[i].x, o2[i].y * o2[i].y)), p.ep.contrast, p.ep.eps);
                    fmn = fminf(fmn, fminf(a1, a2));
                    fmx = fmaxf(fmx, fmaxf(a1, a2));
                    fs += a1 + a2;
                    fs2 = fmaf(a1, a1, fmaf(a2, a2, fs2));
                }
                if (tid == 0) {
                    const float ae = apply_contrast(fast_sqrt(fmaf(ex.x, ex.x, ex.y * ex.y)), p.ep.contrast, p.ep.eps);
                    fmn = fminf(fmn, ae);
                    fmx = fmaxf(fmx, ae);
                    fs += ae;
                    fs2 = fmaf(ae, ae, fs2);
                }
                if (tb + f < n_frames) {
                    acc.mn = fmin(acc.mn, (double)fmn);
                    acc.mx = fmax(acc.mx, (double)fmx);
                    acc.s += (double)fs;
                    acc.s2 += (double)fs2;
                }
            }
            if (u + 1 < u1) {
#pragma unroll
                for (int f = 0; f < FPW; ++f) fetch(v[f], s0 + f * P::SMEM_CF, xclip, uc * G + g * FPW + f);
            }
        } else if (!REALISH) {
#pragma unroll
            for (int f = 0; f < FPW; ++f) {
                // ---- untangle in registers ----
                cf o1[V / 2], o2[V / 2], ex;
                fft.untangle_fwd(v[f], o1, o2, ex);
                const int t = tb + f;
                if (t < n_frames) {
                    float2* __restrict__ row = reinterpret_cast<float2*>(cur_out) + t * P::F;
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
            }
            if (EARLY) {
#pragma unroll
                for (int i = 0; i < V; ++i) v[0][i] = nv[EARLY ? i : 0];
                ms_finish(v[0]);
            } else {
                if (u + 1 < u1) {
#pragma unroll
                    for (int f = 0; f < FPW; ++f) fetch(v[f], s0 + f * P::SMEM_CF, xclip, uc * G + g * FPW + f);
                }
            }
        } else {
            float* __restrict__ vbuf = RIE ? vrows : vrows + buf * (G * VSTR);
            const int t0 = cur_uc * G;
            int n_valid = 0;
            // LAG: this row buffer was last read by the epilogue of the unit two back; every warp ran that epilogue before
            // its previous transform, so the wait is a formality unless a warp is a whole unit behind
            if (LAG && it >= 2) mbar_wait(&lag_empty[buf], ((it >> 1) & 1) ^ 1);
#pragma unroll
            for (int f = 0; f < FPW; ++f) {
                cf o1[V / 2], o2[V / 2], ex;
                fft.untangle_fwd(v[f], o1, o2, ex);
                float* __restrict__ val = vbuf + (g * FPW + f) * VSTR;
#pragma unroll
                for (int c = 0; c < PR::PC; ++c) {
                    float* lo = val + PR::klo(tid, c);
                    float* hi = val + PR::khi(tid, c);
                    float* mlo = val + (M - PR::klo(tid, c));
                    float* mhi = val + (M - PR::khi(tid, c));
#pragma unroll
                    for (int q = 0; q < RP; ++q) {
                        (q < RP / 2 ? lo : hi)[q * NBP] = pow_value<PMODE>(o1[c * RP + q], p.power);
                        (q < RP / 2 ? mlo : mhi)[-q * NBP] = pow_value<PMODE>(o2[c * RP + q], p.power);
                    }
                }
                if (tid == 0) val[M / 2] = pow_value<PMODE>(ex, p.power);
                if (POLAR) {
                    // raw phase of every bin, same row layout; the unit's last row also goes to the rotating carry row
                    float* __restrict__ pval = G == 1 ? prows + pidx * VSTR : prows + buf * (G * VSTR) + (g * FPW + f) * VSTR;
                    const bool last = G > 1 && (g * FPW + f) == G - 1;
                    const int cd = last ? (int)((pcarry + pidx * VSTR) - pval) : 0;      // carry row relative to the tile row
#pragma unroll
                    for (int c = 0; c < PR::PC; ++c) {
                        float* lo = pval + PR::klo(tid, c);
                        float* hi = pval + PR::khi(tid, c);
                        float* mlo = pval + (M - PR::klo(tid, c));
                        float* mhi = pval + (M - PR::khi(tid, c));
#pragma unroll
                        for (int q = 0; q < RP; ++q) {
                            const float a1 = fast_atan2f(o1[c * RP + q].y, o1[c * RP + q].x);
                            const float a2 = fast_atan2f(o2[c * RP + q].y, o2[c * RP + q].x);
                            (q < RP / 2 ? lo : hi)[q * NBP] = a1;
                            (q < RP / 2 ? mlo : mhi)[-q * NBP] = a2;
                            if (last) {
                                (q < RP / 2 ? lo : hi)[q * NBP + cd] = a1;
                                (q < RP / 2 ? mlo : mhi)[-q * NBP + cd] = a2;
                            }
                        }
                    }
                    if (tid == 0) {
                        const float ae = fast_atan2f(ex.y, ex.x);
                        pval[M / 2] = ae;
                        if (last) pval[M / 2 + cd] = ae;
                    }
                }
            }
            // v, o1, o2 are dead: start fetching the next frame's samples, they land during the epilogue
            if (u + 1 < u1) {
                pend_clip = xclip;
                pend_s = ms_s;
                pend_g = ms_g;
#pragma unroll
                for (int f = 0; f < FPW; ++f) {
                    if (RIE) pend[f] = fetch_fast(v[f], xclip, uc * G + g * FPW + f, ms_s, ms_g);
                    else fetch(v[f], s0 + f * P::SMEM_CF, xclip, uc * G + g * FPW + f);
                }
            }
            // One barrier per unit: the rows are complete.  (Double-buffered rows: writers of this buffer two units from now
            // have passed the next barrier, i.e. every thread has left this epilogue.  RIE: see the barrier at the loop top.)
            if (LAG) {
                __syncwarp();
                if ((threadIdx.x & 31) == 0) mbar_arrive(&lag_full[buf]);
                if (it > 0) {
                    mbar_wait(&lag_full[buf ^ 1], ((it - 1) >> 1) & 1);
                    run_epilogue(vrows + (buf ^ 1) * (G * VSTR), lag_out, lag_uc, false);
                    __syncwarp();
                    if ((threadIdx.x & 31) == 0) mbar_arrive(&lag_empty[buf ^ 1]);
                }
                lag_out = cur_out;
                lag_uc = cur_uc;
                ++it;
            } else {
                __syncthreads();
                n_valid = run_epilogue(vbuf, cur_out, cur_uc, POLAR && u < u0);
            }
            if (POLAR) {
                // ---- phase epilogue: raw phase, or the forward-difference IF of spectral_repr.py:319-323 as the wrapped
                // difference of two consecutive raw phases (= the difference of the unwrapped phases of utils/misc.py:12-26
                // without the running sum), the +-pi scalings of spectral_repr.py:352-356, weighting, normalisation ----
                const float* __restrict__ pb = G == 1 ? prows + pidx * VSTR : prows + buf * (G * VSTR);
                const float* __restrict__ pprev = pcarry + (pidx == 0 ? 2 : pidx - 1) * VSTR;
                const int df = p.ep.drop_first, n_keep = P::F - df;
                const int prs = (int)p.ph_row_stride;
#pragma unroll 1
                for (int g0 = 0; g0 < n_valid; ++g0) {
                    const int t = t0 + g0;
                    const float* __restrict__ cur = pb + g0 * VSTR + df;
                    const float* __restrict__ prv = g0 > 0 ? cur - VSTR : pprev + df;
                    float* __restrict__ o = cur_ph + (int64_t)t * prs;
                    const bool diff = ph_if && t > 0;
                    const float s_pi = (ph_if && t < n_frames - 1) ? ACIDS_INV_PI_F : 1.f;
                    const float wgt = p.ph_weighted ? if_weight(t, n_frames) : 1.f;
                    for (int k = threadIdx.x; k < n_keep; k += THREADS) {
                        float v = cur[k];
                        if (diff) {
                            const float d = v - prv[k];
                            v = (d + unwrap_correction(d)) * 0.5f;
                        }
                        v = v * s_pi;
                        if (p.ph_weighted) v *= wgt;
                        stg_stream1(o + k, (v - ph_off) * ph_inv);
                    }
                }
                pidx = pidx == 2 ? 0 : pidx + 1;
            }
#ifdef ACIDS_FWD_SINGLE_ROWBUF     // tuning experiment: one |X| row buffer (8 KB less shared memory per CTA), two barriers per unit
            __syncthreads();
#else
            if (!RIE) buf ^= 1;
#endif
        }
    }
    if (LAG && it > 0) {       // the last unit's epilogue
        const int lb = (it - 1) & 1;
        mbar_wait(&lag_full[lb], ((it - 1) >> 1) & 1);
        run_epilogue(vrows + lb * (G * VSTR), lag_out, lag_uc, false);
    }
    if (STATS) {
        __syncthreads();
        acc = stat_block_reduce(acc);
        if (threadIdx.x == 0) p.stat_part[blockIdx.x] = acc;
    }
}

#ifndef ACIDS_FWD_BAND_BUDGET
#define ACIDS_FWD_BAND_BUDGET (24 * 1024)
#endif
static const size_t kBandSmemBudget = ACIDS_FWD_BAND_BUDGET;

template <class P, int MODE, int PMODE, int CSEL, int BAND, bool TRANSPOSED, bool MS = false>
static int launch_fwd(FwdParams p, cudaStream_t st) {
    using C = FwdCfg<P, MODE>;
    constexpr int THREADS = C::THREADS;
    constexpr int G = C::G;
    size_t smem = C::exch_bytes() + C::win_bytes();
#ifdef ACIDS_FWD_SINGLE_ROWBUF
    constexpr int kRowBufs = 1;
#else
    constexpr int kRowBufs = 2;
#endif
    if (MODE == MODE_REAL || MODE == MODE_POLAR) smem += (BAND == BAND_SMEM ? (size_t)p.band_smem_bytes : 0) + (C::ROWS_IN_EXCH ? 0 : (size_t)kRowBufs * G * C::VSTR * sizeof(float));
    if (MODE == MODE_POLAR) smem += (size_t)(G == 1 ? 3 : 2 * G + 3) * C::VSTR * sizeof(float);
    auto kern = stft_fwd_kernel<P, MODE, PMODE, CSEL, BAND, TRANSPOSED, MS>;
    static PerDevice cache[kMaxDevices];
    PerDevice& pd = per_device(cache);
    size_t& reserved = pd.reserved;
    int& ctas_per_sm = pd.ctas_per_sm;
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
    if (MODE == MODE_STATS) return (int)grid;      // number of partials written (>= 1)
    return ACIDS_OK;
}

// plan of the real-output (magnitude / mel) kernels of an n_fft: n_fft = 1024 runs them on the one-exchange plan
template <class P> struct RealPlan { using type = P; };
template <> struct RealPlan<Fwd1024> { using type = Fwd1024R; };

// ACIDS_FWD_ONLY=<variant>: tuning builds instantiate one variant only (the full set takes minutes per plan)
#ifndef ACIDS_FWD_ONLY
#define ACIDS_FWD_ONLY -1
#endif
template <class P>
static int launch_any(int variant, const FwdParams& p, cudaStream_t st) {
    if (ACIDS_FWD_ONLY >= 0 && variant != ACIDS_FWD_ONLY) {
        set_error("stft_fwd: this tuning build only holds kernel variant %d (asked for %d)", (int)ACIDS_FWD_ONLY, variant);
        return ACIDS_ENOTSUP;
    }
    using PR_ = typename RealPlan<P>::type;
    if constexpr (ACIDS_FWD_ONLY >= 0) {
        if constexpr (ACIDS_FWD_ONLY == VAR_MAG_SMEM) return launch_fwd<PR_, MODE_REAL, 1, -1, BAND_SMEM, false>(p, st);
        else if constexpr (ACIDS_FWD_ONLY == VAR_COMPLEX) return p.midside ? launch_fwd<P, MODE_COMPLEX, 1, ACIDS_CONTRAST_NONE, BAND_NONE, false, true>(p, st)
                                                                           : launch_fwd<P, MODE_COMPLEX, 1, ACIDS_CONTRAST_NONE, BAND_NONE, false>(p, st);
        else if constexpr (ACIDS_FWD_ONLY == VAR_MAG_NOBAND) return launch_fwd<PR_, MODE_REAL, 1, -1, BAND_NONE, false>(p, st);
        else if constexpr (ACIDS_FWD_ONLY == VAR_MEL_POWER_SMEM) return launch_fwd<PR_, MODE_REAL, 2, ACIDS_CONTRAST_NONE, BAND_SMEM, true>(p, st);
        else return ACIDS_ENOTSUP;
    } else
    switch (variant) {
        case VAR_COMPLEX: return p.midside ? launch_fwd<P, MODE_COMPLEX, 1, ACIDS_CONTRAST_NONE, BAND_NONE, false, true>(p, st)
                                           : launch_fwd<P, MODE_COMPLEX, 1, ACIDS_CONTRAST_NONE, BAND_NONE, false>(p, st);
        case VAR_MAG_NOBAND: return launch_fwd<PR_, MODE_REAL, 1, -1, BAND_NONE, false>(p, st);
        case VAR_MAG_SMEM: return launch_fwd<PR_, MODE_REAL, 1, -1, BAND_SMEM, false>(p, st);
        case VAR_MAG_GLOBAL: return launch_fwd<PR_, MODE_REAL, 1, -1, BAND_GLOBAL, false>(p, st);
        case VAR_MEL_POWER_SMEM: return launch_fwd<PR_, MODE_REAL, 2, ACIDS_CONTRAST_NONE, BAND_SMEM, true>(p, st);
        case VAR_MEL_POWER_GLOBAL: return launch_fwd<PR_, MODE_REAL, 2, ACIDS_CONTRAST_NONE, BAND_GLOBAL, true>(p, st);
        case VAR_MEL_ANY_SMEM: return launch_fwd<PR_, MODE_REAL, 0, ACIDS_CONTRAST_NONE, BAND_SMEM, true>(p, st);
        case VAR_MEL_ANY_GLOBAL: return launch_fwd<PR_, MODE_REAL, 0, ACIDS_CONTRAST_NONE, BAND_GLOBAL, true>(p, st);
        case VAR_STATS: return launch_fwd<P, MODE_STATS, 1, -1, BAND_NONE, false>(p, st);
        case VAR_POLAR_NOBAND: return p.midside ? launch_fwd<P, MODE_POLAR, 1, -1, BAND_NONE, false, true>(p, st) : launch_fwd<P, MODE_POLAR, 1, -1, BAND_NONE, false>(p, st);
        case VAR_POLAR_SMEM: return p.midside ? launch_fwd<P, MODE_POLAR, 1, -1, BAND_SMEM, false, true>(p, st) : launch_fwd<P, MODE_POLAR, 1, -1, BAND_SMEM, false>(p, st);
        case VAR_POLAR_GLOBAL: return p.midside ? launch_fwd<P, MODE_POLAR, 1, -1, BAND_GLOBAL, false, true>(p, st) : launch_fwd<P, MODE_POLAR, 1, -1, BAND_GLOBAL, false>(p, st);
    }
    set_error("stft_fwd: unknown kernel variant %d", variant);
    return ACIDS_EINVAL;
}


// one definition per plan, each in its own translation unit (stft_fwd_plan.cu with -DACIDS_FWD_PLAN_N=<n_fft>)
#define ACIDS_DECLARE_FWD_PLAN(n) int launch_fwd_plan_##n(int variant, const FwdParams& p, cudaStream_t st)
ACIDS_DECLARE_FWD_PLAN(32); ACIDS_DECLARE_FWD_PLAN(64); ACIDS_DECLARE_FWD_PLAN(128); ACIDS_DECLARE_FWD_PLAN(256);
ACIDS_DECLARE_FWD_PLAN(512); ACIDS_DECLARE_FWD_PLAN(1024); ACIDS_DECLARE_FWD_PLAN(2048); ACIDS_DECLARE_FWD_PLAN(4096);
ACIDS_DECLARE_FWD_PLAN(8192); ACIDS_DECLARE_FWD_PLAN(16384);

}  // namespace acids
