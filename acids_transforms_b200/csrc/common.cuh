// common.cuh — error plumbing, launch helpers and the shared epilogue of the spectral kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <float.h>
#include "../../include/acids_b200.h"
#include "fft_core.cuh"

namespace acids {

struct EpiParams;
void set_error(const char* fmt, ...);           // capi.cu (thread-local message)
int fill_epilogue(EpiParams& ep, acids_band band, int n_bins, int contrast, float eps, int drop_first, size_t smem_budget);  // stft_fwd.cu
int num_sms();                                  // cached cudaDevAttrMultiProcessorCount

// Function attributes (opt-in shared memory) and occupancy are per DEVICE: launch sites cache them per device ordinal
// (one process may drive several GPUs).  Races between host threads are benign: the cached values are idempotent.
constexpr int kMaxDevices = 64;
struct PerDevice {
    size_t reserved = 0;     // dynamic shared memory the kernel has been opted in to
    int ctas_per_sm = 0;     // occupancy at occ_smem bytes
    size_t occ_smem = 0;
};
static inline PerDevice& per_device(PerDevice (&tab)[kMaxDevices]) {
    int d = 0;
    cudaGetDevice(&d);
    return tab[(d >= 0 && d < kMaxDevices) ? d : 0];
}

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

// ---- mbarrier (shared-memory phase barrier): split arrive / wait, so that a warp can signal "my rows are written" and only
// wait much later, when it needs everybody's rows (the lagged epilogue of the fused forward kernel) ----
__device__ __forceinline__ uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {      // release at CTA scope
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_addr_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {      // acquire at CTA scope
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_addr_u32(bar)), "r"(parity) : "memory");
}

// streaming (read-once / write-once) global accesses: keep them out of L1
__device__ __forceinline__ float2 ldg_stream2(const float2* p) {
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream2(float2* p, float x, float y) {
#ifdef ACIDS_DEBUG_NOSTORE      // tuning experiment only: keep the value alive, skip the store unless it is NaN-tagged
    if (x != 123456.789f) return;
#endif
    asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ void stg_stream1(float* p, float x) {
#ifdef ACIDS_DEBUG_NOSTORE
    if (x != 123456.789f) return;
#endif
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(x) : "memory");
}

// ---- phase helpers shared by the phase kernels (spectral_repr.cu) and the fused STFT -> Polar epilogue ----
#define ACIDS_PI_F 3.14159265358979323846f
#define ACIDS_2PI_F 6.28318530717958647692f
#define ACIDS_INV_PI_F 0.31830988618379067154f
#define ACIDS_INV_2PI_F 0.15915494309189533577f

// atan2 for the phase kernels: |error| < 1e-7 rad (3e-8 of the +-pi range; the parity budget is 1e-4).
// One approximate division to map the argument into [0, 1], the degree-16 even minimax polynomial of atan(a)/a
// (Abramowitz & Stegun 4.4.49, |eps| <= 2e-8), octant fix-ups, IEEE sign of zero: atan2(+0, x < 0) = +pi and
// atan2(-0, x < 0) = -pi (the DC and Nyquist bins carry an exact +0 imaginary part).  ~25 instructions instead of the
// ~50 of atan2f: the phase kernel is issue bound (82 % issue-slot utilisation measured), not memory bound.
__device__ __forceinline__ float fast_atan2f(float y, float x) {
    const float ax = fabsf(x), ay = fabsf(y);
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float a = mx > 0.f ? __fdividef(mn, mx) : 0.f;
    const float z = a * a;
    float p = 0.0028662257f;
    p = fmaf(p, z, -0.0161657367f);
    p = fmaf(p, z, 0.0429096138f);
    p = fmaf(p, z, -0.0752896400f);
    p = fmaf(p, z, 0.1065626393f);
    p = fmaf(p, z, -0.1420889944f);
    p = fmaf(p, z, 0.1999355085f);
    p = fmaf(p, z, -0.3333314528f);
    float r = fmaf(p * z, a, a);
    r = ay > ax ? 1.57079632679489661923f - r : r;
    r = x < 0.f ? 3.14159265358979323846f - r : r;
    return copysignf(r, y);
}

__device__ __forceinline__ float unwrap_correction(float d) {
    // utils/misc.py:19-24: ddmod = (d + pi) % 2pi - pi (python remainder); +pi when it lands on -pi going up
    // d is a difference of two principal values, so x = d + pi lies in [-pi, 3 pi]: the float remainder (exact, like
    // fmodf) reduces to one conditional subtraction (exact by Sterbenz' lemma), then python's sign fix-up
    const float x = d + ACIDS_PI_F;
    float r = x >= ACIDS_2PI_F ? x - ACIDS_2PI_F : x;
    if (r < 0.f) r += ACIDS_2PI_F;
    float dd = r - ACIDS_PI_F;
    if (dd == -ACIDS_PI_F && d > 0.f) dd = ACIDS_PI_F;
    return fabsf(d) < ACIDS_PI_F ? 0.f : dd - d;
}

__device__ __forceinline__ float if_weight(int t, int T) {
    // spectral_repr.py:341-342
    const float N = (float)T, n = (float)t;
    const float a = (n - (N / 2.f - 1.f)) / (N / 2.f);
    return (1.5f * N) / (N * N - 1.f) * (1.f - a * a);
}

// ---- one-pass statistics (min, max, sum, sum of squares in double) shared by acids_stats and the STFT statistics mode ----
struct StatAcc {
    double mn, mx, s, s2;
};
constexpr int kStatsBlocks = 1024;       // partials the scratch buffer of acids_stats_scratch_bytes() holds
int launch_stats_final(const StatAcc* part, int nparts, int64_t n, double* out4, cudaStream_t st);     // pointwise.cu

__device__ __forceinline__ void stat_merge(StatAcc& a, const StatAcc& b) {
    a.mn = fmin(a.mn, b.mn);
    a.mx = fmax(a.mx, b.mx);
    a.s += b.s;
    a.s2 += b.s2;
}

// all threads of the CTA call it; the result is valid on thread 0
__device__ __forceinline__ StatAcc stat_block_reduce(StatAcc a) {
    __shared__ StatAcc sh[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        StatAcc b;
        b.mn = __shfl_xor_sync(0xffffffffu, a.mn, o);
        b.mx = __shfl_xor_sync(0xffffffffu, a.mx, o);
        b.s = __shfl_xor_sync(0xffffffffu, a.s, o);
        b.s2 = __shfl_xor_sync(0xffffffffu, a.s2, o);
        stat_merge(a, b);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) sh[warp] = a;
    __syncthreads();
    if (warp == 0) {
        StatAcc b = lane < nw ? sh[lane] : StatAcc{DBL_MAX, -DBL_MAX, 0.0, 0.0};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            StatAcc c;
            c.mn = __shfl_xor_sync(0xffffffffu, b.mn, o);
            c.mx = __shfl_xor_sync(0xffffffffu, b.mx, o);
            c.s = __shfl_xor_sync(0xffffffffu, b.s, o);
            c.s2 = __shfl_xor_sync(0xffffffffu, b.s2, o);
            stat_merge(b, c);
        }
        a = b;
    }
    return a;
}

// ---- epilogue shared by the fused STFT kernel and the stand-alone Magnitude kernels -------------
//
// Banded matrix, "group-ELL" layout (built by the host, see ops.BandedMatrix): output columns are taken
// in groups of 32 (one warp); group g applies cnt[g] taps to every column, column m reading the input rows
// start[m] .. start[m] + cnt[g] - 1 (columns with a narrower band are zero padded, start[] is shifted so the
// window never leaves the input).  meta = [(cnt[g], base[g]) pairs | start[0..n_out)], coefficient u of
// column m sits at coef[(base[g] + u) * 32 + (m & 31)]: lanes read consecutive floats, the trip count is
// warp uniform, nothing diverges.
struct EpiParams {
    const int32_t* meta;     // banded matrix (or nullptr)
    const float* coef;
    int n_cols;              // output columns before drop_first
    int contrast;
    float eps;
    int drop_first;
    int band_bytes_meta;     // bytes of meta / coef to stage in shared memory (0: read from global memory)
    int band_bytes_coef;
    int coef_floats;         // coefficients actually present in `coef` (band_bytes_coef may reserve more, see split_reserve)
    int split_reserve;       // the shared-memory coefficient area was enlarged for the tap-split layout (transposed launches)
};

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float apply_contrast(float a, int mode, float eps) {
    // spectral_repr.py:191-201.  log(1 + m) is evaluated literally (not log1p), like the reference;
    // lg2.approx has 2^-22 relative error, two orders below the 1e-4 parity budget.
    if (mode == ACIDS_CONTRAST_LOG1P) return __logf(1.0f + a);
    if (mode == ACIDS_CONTRAST_LOG) return __logf(fmaxf(a, eps));
    if (mode == ACIDS_CONTRAST_LOG10) return __log10f(fmaxf(a, eps));
    return a;
}

__device__ __forceinline__ float invert_contrast(float y, int mode, float eps) {
    // spectral_repr.py:203-213
    if (mode == ACIDS_CONTRAST_LOG1P) return expf(y) - 1.0f;
    if (mode == ACIDS_CONTRAST_LOG) return expf(y) - eps;
    if (mode == ACIDS_CONTRAST_LOG10) return powf(10.0f, y);
    return y;
}

__device__ __forceinline__ float fast_lg2(float x) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // x >= eps > FLT_MIN on every path that calls this
    return r;
}

// contrast up to the constant factor that epilogue folds into the normalisation FFMA (see contrast_gain)
template <int CONTRAST>
__device__ __forceinline__ float contrast_core(float a, float eps) {
    // spectral_repr.py:191-201.  log(1 + m) is evaluated literally (not log1p), like the reference;
    // lg2.approx has 2^-22 relative error, two orders below the 1e-4 parity budget.
    if (CONTRAST == ACIDS_CONTRAST_LOG1P) return fast_lg2(1.0f + a);
    if (CONTRAST == ACIDS_CONTRAST_LOG || CONTRAST == ACIDS_CONTRAST_LOG10) return fast_lg2(fmaxf(a, eps));
    return a;
}
__device__ __forceinline__ float contrast_gain(int contrast) {
    if (contrast == ACIDS_CONTRAST_LOG1P || contrast == ACIDS_CONTRAST_LOG) return 0.69314718055994530942f;   // ln 2
    if (contrast == ACIDS_CONTRAST_LOG10) return 0.30102999566398119521f;                                      // log10 2
    return 1.0f;
}

#ifndef ACIDS_EPI_SNAKE
#define ACIDS_EPI_SNAKE 1
#endif
// ---- the row-tile epilogue --------------------------------------------------------------------------------------
// `val` (shared memory) holds the non-negative inputs (|X| or |X|^p) of NF consecutive rows, val_stride floats
// apart.  NT threads produce output columns tid, tid + NT, ... of every row:
//     banded projection -> contrast -> normalise -> streaming store.
// Everything that is uniform over the launch is a template parameter (contrast, where the band lives, output
// orientation, whether all NF rows exist), so the column loop is branch free apart from the warp-uniform
// tap-count dispatch; per-column work (band metadata, coefficients, dispatch) is paid once for the NF rows.
enum { BAND_NONE = 0, BAND_SMEM = 1, BAND_GLOBAL = 2 };

// BAND_SMEM: the kernel re-packs the group-ELL metadata into ONE 8-byte record per column when it stages the band:
//   x = first input row of the column's window, y = (taps << 16) | index of its first coefficient
struct EpiArgs {
    const int32_t* meta;     // BAND_GLOBAL: group-ELL metadata in global memory; BAND_SMEM: int2 records in shared memory
    const float* coef;       // coefficients, same address space
    int n_out;               // output columns (= n_cols - drop_first)
    int n_groups;            // ceil(n_cols / 32)
    int drop_first;
    float gain;              // contrast_gain / scale
    float bias;              // -offset / scale
    float eps;
    int split;               // BAND_SMEM, transposed output: tap-split layout (epilogue_split), see stage_band_split
};

// K taps of one column for NF rows: the K coefficients are loaded once and reused for every row
template <int K, int NF, int BAND>
__device__ __forceinline__ void taps(const float* __restrict__ v, int val_stride, const float* __restrict__ c, float (&a)[NF]) {
    float w[K > 0 ? K : 1];
#pragma unroll
    for (int u = 0; u < K; ++u) w[u] = BAND == BAND_GLOBAL ? __ldg(c + (u << 5)) : c[u << 5];
#pragma unroll
    for (int f = 0; f < NF; ++f) {
        float acc = K > 0 ? v[f * val_stride] * w[0] : 0.f;     // == fmaf(., ., 0): no accumulator to clear
#pragma unroll
        for (int u = 1; u < K; ++u) acc = fmaf(v[f * val_stride + u], w[u], acc);
        a[f] = acc;
    }
}

// Banded projection of column m for NF rows (rows are val_stride floats apart).  The tap count is uniform over
// a group of 32 columns, so the switch does not diverge inside a warp and every case is straight-line code.
template <int NF, int BAND>
__device__ __forceinline__ void band_column(const float* __restrict__ val, int val_stride, const int32_t* __restrict__ meta,
                                            const float* __restrict__ coef, int n_groups, int m, float (&a)[NF]) {
    int cnt;
    const float* __restrict__ v;
    const float* __restrict__ c;
    if (BAND == BAND_SMEM) {
        const int2 rec = reinterpret_cast<const int2*>(meta)[m];
        cnt = (int)((unsigned)rec.y >> 16);
        v = val + rec.x;
        c = coef + (rec.y & 0xffff);
    } else {
        cnt = __ldg(meta + 2 * (m >> 5));
        const int base = __ldg(meta + 2 * (m >> 5) + 1);
        v = val + __ldg(meta + 2 * n_groups + m);
        c = coef + (base << 5) + (m & 31);
    }
    // warp-uniform dispatch on the tap count as a shallow compare tree (cheaper than a jump table: no constant-bank
    // load feeding an indirect branch); the square mel banks use 0..7 taps, most columns 1..4
#define ACIDS_TAPS(K) taps<K, NF, BAND>(v, val_stride, c, a)
    if (cnt <= 4) {
        if (cnt <= 2) {
            if (cnt == 2) ACIDS_TAPS(2);
            else if (cnt == 1) ACIDS_TAPS(1);
            else ACIDS_TAPS(0);
        } else {
            if (cnt == 3) ACIDS_TAPS(3);
            else ACIDS_TAPS(4);
        }
    } else if (cnt <= 8) {
        if (cnt <= 6) {
            if (cnt == 5) ACIDS_TAPS(5);
            else ACIDS_TAPS(6);
        } else {
            if (cnt == 7) ACIDS_TAPS(7);
            else ACIDS_TAPS(8);
        }
    } else {
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            float acc = 0.f;
            for (int u = 0; u < cnt; ++u) acc = fmaf(v[f * val_stride + u], BAND == BAND_GLOBAL ? __ldg(c + (u << 5)) : c[u << 5], acc);
            a[f] = acc;
        }
    }
#undef ACIDS_TAPS
}

// generic (runtime tap count) projection of ONE (row, column) pair — the spread tail of epilogue_tile
template <int BAND>
__device__ __forceinline__ float band_single(const float* __restrict__ vrow, const int32_t* __restrict__ meta,
                                             const float* __restrict__ coef, int n_groups, int m) {
    int cnt;
    const float* __restrict__ v;
    const float* __restrict__ c;
    if (BAND == BAND_SMEM) {
        const int2 rec = reinterpret_cast<const int2*>(meta)[m];
        cnt = (int)((unsigned)rec.y >> 16);
        v = vrow + rec.x;
        c = coef + (rec.y & 0xffff);
    } else {
        cnt = __ldg(meta + 2 * (m >> 5));
        const int base = __ldg(meta + 2 * (m >> 5) + 1);
        v = vrow + __ldg(meta + 2 * n_groups + m);
        c = coef + (base << 5) + (m & 31);
    }
    float acc = 0.f;
    for (int u = 0; u < cnt; ++u) acc = fmaf(v[u], BAND == BAND_GLOBAL ? __ldg(c + (u << 5)) : c[u << 5], acc);
    return acc;
}

// ---- tap-split epilogue for wide bands (the 1025 -> 128 MFCC mel bank: 4 / 10 / 23 / 55 taps over its four 32-column groups) ----
// With one column per lane a warp's work is its group's tap count, and the warp holding the widest group sets the pace of the
// CTA (ncu, cfg 3: 1.9 barrier-stall cycles per issue).  Here a warp takes 8 columns at a time and FOUR lanes share a column,
// each a quarter of its taps; two shuffle-adds combine the partial sums (fixed order: deterministic).  Every warp then
// walks ceil(cnt / 4) taps of every group.  The coefficients are re-packed once per CTA so that the 32 lanes of a step read
// 32 consecutive floats; rec.y of a column holds (taps << 16 | float offset of its group's split block).
// Measured on B200 (cfg 3, 512 x 10 s): 1.773 ms against 1.644 ms for one column per lane — with three CTAs per SM the
// other CTAs fill the SM while a CTA's wide-group warp finishes, so the imbalance costs less than the shuffles and index
// arithmetic of the split.  Correct (the whole GPU suite passes with it on) but opt-in: -DACIDS_EPI_SPLIT=1.
#ifndef ACIDS_EPI_SPLIT
#define ACIDS_EPI_SPLIT 0
#endif
#ifndef ACIDS_SPLIT_MIN_TAPS
#define ACIDS_SPLIT_MIN_TAPS 12
#endif
template <int NT, int NF, int CONTRAST>
__device__ __forceinline__ void epilogue_split(const float* __restrict__ val, int val_stride, int tid, const EpiArgs& ep,
                                               float* __restrict__ out0, int col_step, int n_valid) {
    constexpr int NW = NT / 32;
    const int lane = tid & 31, warp = tid >> 5, j = lane & 7, q = lane >> 3;
    const int n_chunks = (ep.n_out + 7) >> 3;
    for (int c = warp; c < n_chunks; c += NW) {
        const int col = 8 * c + j;
        const int colc = min(col, ep.n_out - 1) + ep.drop_first;
        const int2 rec = reinterpret_cast<const int2*>(ep.meta)[colc];
        const int cnt = (int)((unsigned)rec.y >> 16);
        const int kq = (cnt + 3) >> 2;
        const float* __restrict__ cs = ep.coef + (rec.y & 0xffff) + (c & 3) * kq * 32 + lane;       // chunk c & 3 of its group (drop_first == 0)
        const float* __restrict__ v = val + rec.x;
        const int u0 = q * kq, ulast = cnt - 1;
        float acc[NF];
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[f] = 0.f;
        for (int up = 0; up < kq; ++up) {
            const float w = cs[up * 32];                       // zero beyond the column's taps
            const int u = min(u0 + up, ulast);                  // keep the (unused) read inside the window
#pragma unroll
            for (int f = 0; f < NF; ++f) acc[f] = fmaf(v[f * val_stride + u], w, acc[f]);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], 8);
            acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], 16);
        }
        // the four lanes of a column share out its rows
        if (col < ep.n_out) {
            float* __restrict__ o = out0 + (int64_t)col * col_step;
#pragma unroll
            for (int f = 0; f < NF; ++f)
                if ((f & 3) == q && f < n_valid) stg_stream1(o + f, fmaf(contrast_core<CONTRAST>(acc[f], ep.eps), ep.gain, ep.bias));
        }
    }
}

// FULL: all NF rows exist (no row predicates); otherwise rows f >= n_valid are skipped.
// RS > 0: the output row step is known at compile time (the common contiguous [.., T, F] output): the NF stores of a
// column are one pointer plus immediates instead of NF 64-bit additions.
// SPREAD: when the columns left over after the last full sweep of NT, times the NF rows, fit one sweep of the CTA
// (n_out = 513 = 4 * 128 + 1 is the common case), those (row, column) pairs are dealt one per thread to the LAST warps
// instead of costing warp 0 a whole extra iteration with one active lane — the slowest warp sets the CTA's pace.
template <int NT, int NF, int CONTRAST, int BAND, bool TRANSPOSED, bool FULL, int RS = 0, bool SPREAD = false>
__device__ __forceinline__ void epilogue_tile(const float* __restrict__ val, int val_stride, int tid, const EpiArgs& ep,
                                              float* __restrict__ out0, int row_step, int col_step, int n_valid) {
    // out0: element (row 0, first output column) of the tile; rows are row_step floats apart, columns col_step
    // (TRANSPOSED: rows are adjacent, row_step is ignored; otherwise columns are adjacent, col_step is ignored)
    // one pointer per column (row 0) plus loop-invariant row offsets
    if constexpr (TRANSPOSED && BAND == BAND_SMEM) {
        if (ep.split) {
            epilogue_split<NT, NF, CONTRAST>(val, val_stride, tid, ep, out0, col_step, FULL ? NF : n_valid);
            return;
        }
    }
    int64_t roff[NF];
#pragma unroll
    for (int f = 0; f < NF; ++f) roff[f] = TRANSPOSED ? (int64_t)f : (RS > 0 ? (int64_t)f * RS : (int64_t)f * row_step);
    int n_main = ep.n_out;
    if (SPREAD) {
        const int rem = ep.n_out % NT;
        if (rem * NF <= NT) n_main = ep.n_out - rem;
    }
    // Odd sweeps hand the 32-column groups to the warps in REVERSE order: a bank whose tap count grows with the column index
    // (mel: 1 ... 7 taps over the 16 groups of the 513-column bank) then gives every warp a similar total instead of loading
    // the last warp with the widest group of every sweep (fused cfg-2 kernel 0.999 -> 0.982 ms).  ACIDS_EPI_SNAKE=0: plain order.
    int sweep = 0;
    for (int mo0 = tid; mo0 < n_main; mo0 += NT, ++sweep) {
        int mo = mo0;
        if (ACIDS_EPI_SNAKE && !TRANSPOSED && (sweep & 1) && (mo0 - tid) + NT <= n_main)       // full sweeps only
            mo = (mo0 - tid) + 32 * (NT / 32 - 1 - (tid >> 5)) + (tid & 31);
        float* __restrict__ o = out0 + (TRANSPOSED ? (int64_t)mo * col_step : (int64_t)mo);
        const int m = mo + ep.drop_first;
        float a[NF];
        if (BAND != BAND_NONE) {
            band_column<NF, BAND>(val, val_stride, ep.meta, ep.coef, ep.n_groups, m, a);
        } else {
#pragma unroll
            for (int f = 0; f < NF; ++f) a[f] = val[f * val_stride + m];
        }
#pragma unroll
        for (int f = 0; f < NF; ++f)
            if (FULL || f < n_valid) stg_stream1(o + roff[f], fmaf(contrast_core<CONTRAST>(a[f], ep.eps), ep.gain, ep.bias));
    }
    if (SPREAD && n_main < ep.n_out) {
        const int i = NT - 1 - tid;                      // dealt from the last thread backwards
        const int f = i % NF, c = i / NF;                // row of the tile, left-over column
        if (c < ep.n_out - n_main && (FULL || f < n_valid)) {
            const int mo = n_main + c, m = mo + ep.drop_first;
            const float a = BAND != BAND_NONE ? band_single<BAND>(val + f * val_stride, ep.meta, ep.coef, ep.n_groups, m)
                                              : val[f * val_stride + m];
            float* __restrict__ oo = out0 + (TRANSPOSED ? (int64_t)mo * col_step + f
                                                        : (int64_t)mo + (RS > 0 ? (int64_t)f * RS : (int64_t)f * row_step));
            stg_stream1(oo, fmaf(contrast_core<CONTRAST>(a, ep.eps), ep.gain, ep.bias));
        }
    }
}

// runtime contrast / row count -> compile-time (one dispatch per tile; call sites that know the contrast at compile
// time pass it as CSEL >= 0 and pay no dispatch)
template <int NT, int NF, int CSEL, int BAND, bool TRANSPOSED, int RS = 0, bool SPREAD = false>
__device__ __forceinline__ void epilogue_dispatch(int contrast, const float* __restrict__ val, int val_stride, int tid,
                                                  const EpiArgs& ep, float* __restrict__ out0, int row_step, int col_step, int n_valid) {
#define ACIDS_TILE(C)                                                                                                   \
    do {                                                                                                                \
        if (n_valid >= NF) epilogue_tile<NT, NF, C, BAND, TRANSPOSED, true, RS, SPREAD>(val, val_stride, tid, ep, out0, row_step, col_step, NF); \
        else epilogue_tile<NT, NF, C, BAND, TRANSPOSED, false, RS, SPREAD>(val, val_stride, tid, ep, out0, row_step, col_step, n_valid);       \
    } while (0)
    if (CSEL >= 0) {
        ACIDS_TILE((CSEL >= 0 ? CSEL : 0));
    } else {
        switch (contrast) {
            case ACIDS_CONTRAST_LOG1P: ACIDS_TILE(ACIDS_CONTRAST_LOG1P); break;
            case ACIDS_CONTRAST_LOG: ACIDS_TILE(ACIDS_CONTRAST_LOG); break;
            case ACIDS_CONTRAST_LOG10: ACIDS_TILE(ACIDS_CONTRAST_LOG10); break;
            default: ACIDS_TILE(ACIDS_CONTRAST_NONE); break;
        }
    }
#undef ACIDS_TILE
}

// stage the banded matrix in shared memory (all threads of the CTA; caller synchronises): one int2 record per
// column (see EpiArgs) + the coefficients
__device__ __forceinline__ void stage_band(const EpiParams& ep, int32_t* srec, float* scoef) {
    const int n_groups = (ep.n_cols + 31) >> 5;
    for (int m = threadIdx.x; m < ep.n_cols; m += blockDim.x) {
        const int cnt = __ldg(ep.meta + 2 * (m >> 5)), base = __ldg(ep.meta + 2 * (m >> 5) + 1);
        reinterpret_cast<int2*>(srec)[m] = make_int2(__ldg(ep.meta + 2 * n_groups + m), (cnt << 16) | ((base << 5) + (m & 31)));
    }
    for (int i = threadIdx.x; i < ep.coef_floats; i += blockDim.x) scoef[i] = __ldg(ep.coef + i);
}

// does the bank want the tap-split epilogue?  (uniform over the CTA; only where the launcher reserved the room)
__device__ __forceinline__ bool band_wants_split(const EpiParams& ep) {
    if (!ACIDS_EPI_SPLIT || !ep.split_reserve || ep.drop_first != 0) return false;
    const int n_groups = (ep.n_cols + 31) >> 5;
    int mx = 0;
    for (int g = 0; g < n_groups; ++g) mx = max(mx, __ldg(ep.meta + 2 * g));
    return mx >= ACIDS_SPLIT_MIN_TAPS;
}

// the same, tap-split layout (epilogue_split): records carry (taps << 16 | offset of the group's split block); the block of
// group g holds, for each of its four 8-column chunks k and each step u', the 32 lanes' coefficients
//   lane (q = lane >> 3, j = lane & 7) -> coefficient q * kq + u' of column 32 g + 8 k + j   (zero beyond the column's taps)
__device__ __forceinline__ void stage_band_split(const EpiParams& ep, int32_t* srec, float* scoef) {
    const int n_groups = (ep.n_cols + 31) >> 5;
    int gbase = 0;
    for (int g = 0; g < n_groups; ++g) {
        const int cnt = __ldg(ep.meta + 2 * g), base = __ldg(ep.meta + 2 * g + 1);
        const int kq = (cnt + 3) >> 2, block = 4 * kq * 32;
        for (int m = 32 * g + threadIdx.x; m < min(32 * g + 32, ep.n_cols); m += blockDim.x)
            reinterpret_cast<int2*>(srec)[m] = make_int2(__ldg(ep.meta + 2 * n_groups + m), (cnt << 16) | gbase);
        for (int i = threadIdx.x; i < block; i += blockDim.x) {
            const int k = i / (kq * 32), r = i - k * kq * 32, up = r >> 5, ln = r & 31;
            const int u = (ln >> 3) * kq + up;
            scoef[gbase + i] = u < cnt ? __ldg(ep.coef + (((base + u) << 5) + 8 * k + (ln & 7))) : 0.f;
        }
        gbase += block;
    }
}

__device__ __forceinline__ EpiArgs make_epi_args(const EpiParams& ep, const int32_t* meta, const float* coef, const float* offset,
                                                 const float* scale) {
    // (x - offset) / scale as one FFMA: x * (1/scale) + (-offset/scale)   (norm.py:40-41); the constant factor of
    // the lg2-based contrast is folded into the same multiply
    EpiArgs a;
    a.meta = meta;
    a.coef = coef;
    a.n_out = ep.n_cols - ep.drop_first;
    a.n_groups = (ep.n_cols + 31) >> 5;
    a.drop_first = ep.drop_first;
    const float off = offset ? __ldg(offset) : 0.f;
    const float inv = scale ? 1.0f / __ldg(scale) : 1.0f;
    a.gain = inv * contrast_gain(ep.contrast);
    a.bias = -off * inv;
    a.eps = ep.eps;
    a.split = 0;
    return a;
}

// bytes of shared memory the banded matrix needs, or 0 when it should stay in global memory
static inline void band_smem_plan(const acids_band& band, int64_t coef_floats, size_t budget, EpiParams& ep) {
    ep.band_bytes_meta = ep.band_bytes_coef = 0;
    if (!band.meta) return;
    // in shared memory: one 8-byte record per column + the coefficients (indexable with 16 bits, see EpiArgs)
    const size_t need = (size_t)band.n_out * 8 + (size_t)coef_floats * 4;
    ep.coef_floats = (int)coef_floats;
    ep.split_reserve = 0;
    if (need <= budget && coef_floats < 65536 && band.n_out < 65536) {
        ep.band_bytes_meta = band.n_out * 8;
        ep.band_bytes_coef = (int)(coef_floats * 4);
    }
}

}  // namespace acids
