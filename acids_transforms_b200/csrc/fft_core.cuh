// fft_core.cuh — register/shared-memory real FFT core for the acids_transforms hot path (sm_100a).
//
// One frame of N real samples is transformed as an M = N/2 point complex FFT of
// z[n] = x[2n] + i x[2n+1] by T cooperating threads that each keep V = M/T complex values in
// registers.  Passes are Stockham (self-sorting) radix-R steps: a pass reads its R inputs at the
// fixed stride M/R and writes its outputs at stride Ns (the product of the earlier radices), so
// after the last pass the spectrum is in natural order.  Between passes the values cross threads
// through a swizzled shared-memory buffer private to the frame's thread group.
//
// The real<->half-complex "untangle" needs Z[k] together with Z[M-k].  In the pass that touches
// the spectrum side (last pass forward, first pass inverse) butterfly j holds Z[j + q*M/R] and
// butterfly M/R - j holds exactly the mirrored bins, so each thread takes butterflies in such
// PAIRS and the untangle happens in registers: no extra exchange.  Butterflies 0 and M/(2R) are
// their own mirrors; thread 0 takes both and re-routes its operands with selects.
//
// Everything that only depends on the thread's role (pass twiddles, untangle twiddles, analysis
// window taps) is computed once per kernel and kept in registers across the frame loop.
//
// The per-thread phases are __host__ __device__ so tests/emu/emu_fft.cpp can run the very same
// index arithmetic on the CPU (there is no GPU in the build container).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define ACIDS_HD __host__ __device__ __forceinline__
#define ACIDS_ALIGN8 __align__(8)
#else
#define ACIDS_HD inline
#define ACIDS_ALIGN8 alignas(8)
#endif

namespace acids {

struct ACIDS_ALIGN8 cf {
    float x, y;
};

#if defined(__CUDACC__)
struct __align__(16) cf2 { cf a, b; };     // two adjacent complex slots: one 128-bit access
#else
struct alignas(16) cf2 { cf a, b; };
#endif

ACIDS_HD cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }
#if defined(__CUDA_ARCH__) && !defined(ACIDS_NO_PACKED)     // ACIDS_NO_PACKED: tuning experiment (scalar FADD / FFMA instead)
// Blackwell packed FP32 (CUDA 12.9 float2 builtins -> FADD2 / FMUL2 / FFMA2): a complex value is one aligned
// register pair, so a complex add / sub is ONE instruction and a complex multiply TWO
// (FMUL2 a, w.x ; FFMA2 swap(a) * (-w.y, +w.y) + .) — ptxas folds the swap, the per-half sign and the scalar
// broadcast into operand modifiers (.LO_HI, .NP, .F32).  The FP32 pipe still retires 128 results / clk / SM;
// what is saved are issue slots (tools/micro/f32x2.cu).
__device__ __forceinline__ cf operator+(cf a, cf b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
    return mk(r.x, r.y);
}
__device__ __forceinline__ cf operator-(cf a, cf b) {
    const float2 r = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    return mk(r.x, r.y);
}
__device__ __forceinline__ cf cmul(cf a, cf b) {
    const float2 t = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.x));
    const float2 r = __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), t);
    return mk(r.x, r.y);
}
__device__ __forceinline__ cf cmul2(cf a, cf w) {      // element-wise (a.x w.x, a.y w.y): one FMUL2
    const float2 r = __fmul2_rn(make_float2(a.x, a.y), make_float2(w.x, w.y));
    return mk(r.x, r.y);
}
#else
ACIDS_HD cf cmul2(cf a, cf w) { return mk(a.x * w.x, a.y * w.y); }
ACIDS_HD cf operator+(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
ACIDS_HD cf operator-(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
ACIDS_HD cf cmul(cf a, cf b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
#endif
ACIDS_HD cf cconj(cf a) { return mk(a.x, -a.y); }
ACIDS_HD cf mul_pi(cf a) { return mk(-a.y, a.x); }   // a * (+i)
ACIDS_HD cf mul_mi(cf a) { return mk(a.y, -a.x); }   // a * (-i)
ACIDS_HD cf csel(bool p, cf a, cf b) { return mk(p ? a.x : b.x, p ? a.y : b.y); }

// a * e^{-+ 2 pi i Q / R}  (forward: minus).  Q, R compile-time, R | 32, 0 <= Q < R/2.
template <int R, int Q, bool INV>
ACIDS_HD cf mul_root(cf a) {
    constexpr int t = Q * (32 / R);
    if (t == 0) return a;
    if (t == 8) return INV ? mul_pi(a) : mul_mi(a);
    // literal constants so they fold into FFMA immediates
    constexpr float c = (t == 1) ? 0.980785280f : (t == 2) ? 0.923879533f : (t == 3) ? 0.831469612f :
                        (t == 4) ? 0.707106781f : (t == 5) ? 0.555570233f : (t == 6) ? 0.382683432f :
                        (t == 7) ? 0.195090322f : (t == 9) ? -0.195090322f : (t == 10) ? -0.382683432f :
                        (t == 11) ? -0.555570233f : (t == 12) ? -0.707106781f : (t == 13) ? -0.831469612f :
                        (t == 14) ? -0.923879533f : -0.980785280f;
    constexpr float s = (t == 1) ? 0.195090322f : (t == 2) ? 0.382683432f : (t == 3) ? 0.555570233f :
                        (t == 4) ? 0.707106781f : (t == 5) ? 0.831469612f : (t == 6) ? 0.923879533f :
                        (t == 7) ? 0.980785280f : (t == 9) ? 0.980785280f : (t == 10) ? 0.923879533f :
                        (t == 11) ? 0.831469612f : (t == 12) ? 0.707106781f : (t == 13) ? 0.555570233f :
                        (t == 14) ? 0.382683432f : 0.195090322f;
    return cmul(a, mk(c, INV ? s : -s));
}

// the same with a run-time slot index that is a constant after unrolling (the switch folds away)
template <int R, bool INV>
ACIDS_HD cf mul_root_sw(cf a, int q) {
#define ACIDS_ROOT_CASE(Q) case Q: return mul_root<R, ((Q) < R / 2 ? (Q) : 0), INV>(a);
    switch (q) {
        ACIDS_ROOT_CASE(0) ACIDS_ROOT_CASE(1) ACIDS_ROOT_CASE(2) ACIDS_ROOT_CASE(3) ACIDS_ROOT_CASE(4) ACIDS_ROOT_CASE(5)
        ACIDS_ROOT_CASE(6) ACIDS_ROOT_CASE(7) ACIDS_ROOT_CASE(8) ACIDS_ROOT_CASE(9) ACIDS_ROOT_CASE(10) ACIDS_ROOT_CASE(11)
        ACIDS_ROOT_CASE(12) ACIDS_ROOT_CASE(13) ACIDS_ROOT_CASE(14) ACIDS_ROOT_CASE(15)
    }
#undef ACIDS_ROOT_CASE
    return a;
}

// In-place radix-R DFT of a[0..R), natural-order output.  b[q] = sum_r a[r] e^{-+2 pi i r q / R}.
template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<1, INV> {
    static ACIDS_HD void run(cf*) {}
};

template <bool INV>
struct Dft<2, INV> {
    static ACIDS_HD void run(cf* a) {
        cf t = a[0];
        a[0] = t + a[1];
        a[1] = t - a[1];
    }
};

template <bool INV>
struct Dft<4, INV> {
    static ACIDS_HD void run(cf* a) {
        cf t0 = a[0] + a[2], t1 = a[0] - a[2], t2 = a[1] + a[3];
        cf d = a[1] - a[3];
        cf t3 = INV ? mul_pi(d) : mul_mi(d);
        a[0] = t0 + t2;
        a[2] = t0 - t2;
        a[1] = t1 + t3;
        a[3] = t1 - t3;
    }
};

template <int R, int Q, bool INV>
struct Combine {
    static ACIDS_HD void run(cf* a, const cf* e, const cf* o) {
        cf t = mul_root<R, Q, INV>(o[Q]);
        a[Q] = e[Q] + t;
        a[Q + R / 2] = e[Q] - t;
        Combine<R, Q + 1, INV>::run(a, e, o);
    }
};
template <int R, bool INV>
struct Combine<R, R / 2, INV> {
    static ACIDS_HD void run(cf*, const cf*, const cf*) {}
};

template <int R, bool INV>
struct Dft {
    static ACIDS_HD void run(cf* a) {
        cf e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) {
            e[i] = a[2 * i];
            o[i] = a[2 * i + 1];
        }
        Dft<R / 2, INV>::run(e);
        Dft<R / 2, INV>::run(o);
        Combine<R, 0, INV>::run(a, e, o);
    }
};

// (cos, sin)(2 pi num / den), accurate (computed once per kernel, not per frame)
ACIDS_HD cf unit(int num, int den) {
    float s, c;
#if defined(__CUDA_ARCH__)
    sincospif(2.0f * (float)num / (float)den, &s, &c);
#else
    double a = 2.0 * 3.14159265358979323846 * (double)num / (double)den;
    s = (float)sin(a);
    c = (float)cos(a);
#endif
    return mk(c, s);
}

// Shared-memory index swizzle for the exchange buffers: TWO pad slots (16 bytes) every 2^PADLOG slots, so that
// pairs of slots stay 16-byte aligned (128-bit stores of the first pass) and a run of 64 slots is displaced by
// half a bank row (the strided stores of the middle passes of the 8x8x8 plan become conflict free).
// swz(base + c) == swz(base) + swz(c) whenever (base mod 2^PADLOG) + (c mod 2^PADLOG) < 2^PADLOG, which
// holds for every (pass, butterfly, slot) of every plan (checked exhaustively by tests/emu/emu_fft.cpp):
// each access is one per-thread base plus a compile-time immediate.
template <int PADLOG>
ACIDS_HD int swz(int i) {
    return i + 2 * (i >> PADLOG);
}
template <int PADLOG>
ACIDS_HD constexpr int swzc(int c) {
    return c + 2 * (c >> PADLOG);
}
#if !defined(__CUDA_ARCH__) && defined(ACIDS_EMU_CHECK)
#define ACIDS_EMU_ASSERT(cond) do { if (!(cond)) { printf("emu assert failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); abort(); } } while (0)
#else
#define ACIDS_EMU_ASSERT(cond) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------
// Plan: N real points, T threads per frame, up to 4 passes.  For the forward transform the LAST
// pass is the paired one (needs an even butterfly count per thread); for the inverse the FIRST.
// ---------------------------------------------------------------------------------------------
template <int N_, int T_, int R0_, int R1_, int R2_ = 1, int R3_ = 1>
struct Plan {
    static constexpr int N = N_;
    static constexpr int M = N_ / 2;
    static constexpr int F = N_ / 2 + 1;
    static constexpr int T = T_;
    static constexpr int V = M / T_;
    static constexpr int NP = (R3_ > 1) ? 4 : ((R2_ > 1) ? 3 : 2);
#ifndef ACIDS_T16_PADLOG
#define ACIDS_T16_PADLOG 5
#endif
    // two pad slots every 2^PADLOG slots.  The one-exchange plan (16 threads x radix 32) stores runs of 32 slots per thread:
    // with a pad every 32 the 128-bit stores of a quarter warp are 272 bytes apart (conflict free; every 16: 288 bytes, 2-way)
#ifndef ACIDS_T16I_PADLOG
#define ACIDS_T16I_PADLOG 4
#endif
#ifndef ACIDS_INV4096_PADLOG
#define ACIDS_INV4096_PADLOG 5      // pad every 32 slots: 640 instead of 704 exchange wavefronts per frame (ideal 512), 0.664 -> 0.644 ms at cfg 4
#endif
    static constexpr int PADLOG = (N_ == 1024 && T_ == 16) ? (R0_ == 32 ? ACIDS_T16_PADLOG : ACIDS_T16I_PADLOG)
                                                           : ((N_ == 4096 && R0_ == 8) ? ACIDS_INV4096_PADLOG : 4);
    // exchange buffer, complex slots (even: 16-byte rows).  Two frames share a warp in the 16-thread plan and park their |X|
    // rows in their own buffers: 2 * SMEM_CF = 16 (mod 32) floats puts the two half warps' row stores on disjoint banks
    static constexpr int SMEM_CF = M + 2 * (M >> PADLOG) + 2 + ((N_ == 1024 && T_ == 16) ? 6 : 0);
    static_assert(R0_ * R1_ * R2_ * R3_ == M, "radices must multiply to N/2");
    static_assert(M % T_ == 0, "T must divide N/2");
    static constexpr int radix(int p) { return p == 0 ? R0_ : (p == 1 ? R1_ : (p == 2 ? R2_ : R3_)); }
    static constexpr int ns(int p) {
        int s = 1;
        for (int i = 0; i < p; ++i) s *= radix(i);
        return s;
    }
    static constexpr int nb(int p) { return M / radix(p); }          // butterflies in pass p
    static constexpr int bpt(int p) { return nb(p) / T_; }           // butterflies per thread
    // twiddle storage: pass p >= 1 keeps (R-1) factors per butterfly it owns; when T is a multiple of Ns the
    // twiddle index k = (tid + T b) mod Ns does not depend on b and ONE set serves all of a thread's butterflies
    // (tw_shared is overridden to false for the paired pass by FrameFFT, whose butterflies are not tid + T b)
    static constexpr bool tw_shared(int p) { return p > 0 && (T_ % ns(p)) == 0; }
    static constexpr int tw_sets(int p, bool paired) { return (tw_shared(p) && !paired) ? 1 : bpt(p); }
    static constexpr int tw_count(int p) { return p == 0 ? 0 : bpt(p) * (radix(p) - 1); }
    static constexpr int tw_off(int p) {
        int s = 0;
        for (int i = 0; i < p; ++i) s += tw_count(i);
        return s;
    }
    static constexpr int TWN = tw_off(NP) > 0 ? tw_off(NP) : 1;
};

// Lane permutation of the paired pass.  Butterfly pi and its mirror NB - pi are taken by the same thread; with
// pi = tid the mirrored operands of a half warp are a DESCENDING run of 16 exchange slots that straddles two pad
// groups of the swizzle and folds back onto itself by one bank pair: every mirrored 8-byte access costs 2 wavefronts per
// half warp instead of 1 (ncu: 16 of the 144 exchange wavefronts of a 1024-point frame).  The bank conflicts of the
// ascending and of the mirrored accesses are two perfect matchings on the 32 lanes; their union is 2-colourable, and for
// the 32-thread forward plan the colouring differs from (lanes 0-15 | lanes 16-31) by ONE transposition: lanes 2 and 16
// trade their butterfly pairs and both access streams are conflict free (tests/emu/emu_fft.cpp counts the wavefronts:
// n_fft 1024 144 -> 128 = ideal; the larger forward plans gain their first warp's share; tests/emu/emu_search.cpp is the
// search).  The inverse plans' excess sits in the mirrored STORES of their first pass; no single transposition helps.
#if !defined(__CUDA_ARCH__) && defined(ACIDS_EMU_SEARCH)
static int g_pair_swap_a = 0, g_pair_swap_b = 0;
#endif
template <class P, int PASS>
struct PairLane {
#if !defined(__CUDA_ARCH__) && defined(ACIDS_EMU_SEARCH)
    static ACIDS_HD int map(int tid) {
        const int l = tid & 31;
        return l == g_pair_swap_a ? (tid - l + g_pair_swap_b) : (l == g_pair_swap_b ? (tid - l + g_pair_swap_a) : tid);
    }
#else
#ifdef ACIDS_PAIR_SWAP         // opt-in: measured SLOWER on B200 although the counted wavefronts drop
    static constexpr bool SWAP = P::T >= 32 && PASS == P::NP - 1 && PASS > 0;      // the forward plans of n_fft >= 1024
#else
    static constexpr bool SWAP = false;
#endif
    static ACIDS_HD int map(int tid) {
        if (!SWAP) return tid;
        const int l = tid & 31;
        return l == 2 ? tid + 14 : (l == 16 ? tid - 14 : tid);
    }
#endif
};

// Index helpers of the paired pass (radix R, nb = M/R butterflies, pair c of thread tid).
template <class P, int PASS>
struct Pairing {
    static constexpr int R = P::radix(PASS);
    static constexpr int NB = P::nb(PASS);
    static constexpr int PC = P::bpt(PASS) / 2;   // pairs per thread
    static_assert(P::bpt(PASS) % 2 == 0, "paired pass needs an even butterfly count per thread");
    // pair index of (thread, c); 0 is the self-mirrored pair and stays on thread 0
    static ACIDS_HD int pi(int tid, int c) { return PairLane<P, PASS>::map(tid) + P::T * c; }
    static ACIDS_HD int ja(int tid, int c) { return pi(tid, c); }
    static ACIDS_HD int jb(int tid, int c) {
        int p = pi(tid, c);
        return p == 0 ? NB / 2 : NB - p;
    }
    // k1(tid, c, s) = (s < R/2 ? klo : khi) + s * NB: two per-thread bases, the rest is an immediate
    static ACIDS_HD int klo(int tid, int c) { return pi(tid, c); }       // 0 for the self-mirrored pair
    static ACIDS_HD int khi(int tid, int c) {
        int p = pi(tid, c);
        return p != 0 ? p : NB / 2 - (R / 2) * NB;
    }
    // bin index of untangle slot s of pair c: X[k1] and X[M - k1] come out of it
    static ACIDS_HD int k1(int tid, int c, int s) {
        int p = pi(tid, c);
        if (p != 0) return p + s * NB;
        return s < R / 2 ? s * NB : NB / 2 + (s - R / 2) * NB;
    }
};

// ---------------------------------------------------------------------------------------------
// Per-thread state and phases.  INV = false: real -> half-complex; INV = true: the reverse.
// v[] (V complex registers) is owned by the caller so that it can be shared with the epilogue.
// ---------------------------------------------------------------------------------------------
template <class P, bool INV>
struct FrameFFT {
    static constexpr int PAIRED = INV ? 0 : P::NP - 1;
    using PR = Pairing<P, PAIRED>;
    // COMPACT_WK: the untangle twiddle of slot s of a pair is W_N^{k1}, k1 = (s < R/2 ? klo : khi) + s NB, and W_N^{s NB} =
    // W_{2R}^s is a compile-time 2R-th root of unity: keep W_N^{klo} and W_N^{khi} per pair (4 registers) instead of R
    // twiddles (16 registers for the radix-8 plans) and apply the constant root with immediates (mul_root: free for s = 0 and
    // s = R/2, two packed instructions otherwise).  Registers, not arithmetic, bound these kernels.
#ifndef ACIDS_COMPACT_WK
#define ACIDS_COMPACT_WK 1
#endif
    // (the n_fft = 1024 inverse runs at 168 registers / 3 CTAs either way — at 128 / 4 CTAs it is SLOWER even without spills,
    // 1.37 vs 1.11 ms — so it keeps its full twiddle tables and saves the ~45 instructions per frame; ACIDS_INV_COMPACT=1 and
    // ACIDS_INV_DERIVED_TW=1 switch the compact forms on for the inverse plans too)
#ifndef ACIDS_INV_COMPACT
#define ACIDS_INV_COMPACT 0
#endif
    static constexpr bool COMPACT_WK = ACIDS_COMPACT_WK && (32 % (2 * PR::R) == 0) && (!INV || ACIDS_INV_COMPACT || P::T == 16);
    cf tw[P::TWN];        // pass twiddles (already conjugated for INV)
    cf wk[COMPACT_WK ? 2 * PR::PC : P::V / 2];      // untangle twiddles e^{-2 pi i k1 / N} per (pair, slot); conj for INV
    int tid;

    ACIDS_HD cf untangle_mul(cf a, int c, int s) const {     // a * W_N^{-+ k1(c, s)}; c, s constants after unrolling
        if (COMPACT_WK) return cmul(mul_root_sw<(COMPACT_WK ? 2 * PR::R : 32), INV>(a, s), wk[2 * c + (s < PR::R / 2 ? 0 : 1)]);
        return cmul(a, wk[c * PR::R + s]);
    }

    template <int PASS>
    static constexpr bool tw_is_shared() { return PASS != PAIRED && P::tw_shared(PASS); }

    // butterfly that slot b of this thread computes in pass PASS.  Paired pass: mirrored couples (see above).
    // First pass of the forward transform: bpt(0) CONSECUTIVE butterflies, so that a thread's operands of one
    // radix slot are adjacent in memory (128-bit global / window loads) and its outputs one contiguous run
    // (128-bit exchange stores).  Elsewhere: strided by T (consecutive lanes touch consecutive slots).
    template <int PASS>
    ACIDS_HD int bfly(int b) const {
        if (PASS == PAIRED) return (b & 1) ? PR::jb(tid, b >> 1) : PR::ja(tid, b >> 1);
        if (PASS == 0) return tid * P::bpt(0) + b;
        return tid + P::T * b;
    }

    // MIRROR_TW (forward, paired last pass): NS * R = M, so butterfly j < NB carries the twiddles W_M^{r j}, and its mirror
    // NB - j carries W_M^{r (NB - j)} = W_R^r * conj(W_M^{r j}).  A factor W_R^r on input r of a radix-R DFT only rotates its
    // outputs by one bin, so the mirrored butterfly multiplies by the CONJUGATE of its partner's twiddles and relabels its
    // outputs: one twiddle set per pair instead of two (14 registers fewer per thread for the radix-8 plans), for free.
    // The self-mirrored pair (0, NB/2) of thread 0 keeps the set of NB/2 (W_{2R}^r, which the same rule maps onto itself)
    // and skips the multiplication for butterfly 0.
#ifndef ACIDS_FWD_MIRROR_TW
#define ACIDS_FWD_MIRROR_TW 1
#endif
    template <int PASS>
    static constexpr bool mirror_tw() { return ACIDS_FWD_MIRROR_TW && !INV && PASS == PAIRED && PASS > 0 && P::ns(PASS) * P::radix(PASS) == P::M; }

    // POST_TW (inverse, two-pass plans with a wide last pass, i.e. the 16-thread n_fft = 1024 plan): the twiddles of pass 1
    // (R1 - 1 = 31 per thread on the input side) are applied to the OUTPUTS of the paired pass 0 instead: output q of
    // butterfly j is element j R0 + q, which pass 1 multiplies by W_M^{j q}.  The mirrored butterfly NB0 - j needs
    // W_M^{(NB0 - j) q} = W_{R0}^q * conj(W_M^{j q}), and a factor W_{R0}^q on OUTPUT q of a radix-R0 DFT is a cyclic shift of
    // its INPUTS by one: R0 - 1 = 15 twiddles per thread serve both butterflies, the shift is a register relabelling.
    static constexpr bool POST_TW = INV && P::NP == 2 && P::N == 1024 && P::T == 16 && P::ns(1) * P::radix(1) == P::M &&
                                    P::nb(1) == P::radix(0);

    // DERIVED_TW (a non-paired last pass with two butterflies per thread whose twiddle index differs by T): butterfly b = 1
    // carries W^{r (k + T)} = W^{r k} * W_{NS R / T}^r — its partner's twiddles times a compile-time root of unity.  One set
    // instead of two (14 registers fewer at radix 8) for ~12 packed instructions per frame; with the compact untangle twiddles
    // this brings the n_fft = 1024 inverse to 128 registers with 52 bytes of spills (104-144 before) — measured, not adopted.
#ifndef ACIDS_INV_DERIVED_TW
#define ACIDS_INV_DERIVED_TW 0
#endif
    template <int PASS>
    static constexpr bool derived_tw() {
        return ACIDS_INV_DERIVED_TW && INV && PASS > 0 && PASS == P::NP - 1 && PASS != PAIRED && !P::tw_shared(PASS) && P::bpt(PASS) == 2 &&
               P::ns(PASS) % P::T == 0 && P::ns(PASS) / P::T >= 2 && (32 % ((P::ns(PASS) * P::radix(PASS)) / P::T)) == 0 &&
               P::radix(PASS) - 1 < ((P::ns(PASS) * P::radix(PASS)) / P::T) / 2;
    }

    template <int PASS>
    ACIDS_HD void init_pass() {
        constexpr int R = P::radix(PASS), NS = P::ns(PASS), B = (tw_is_shared<PASS>() || derived_tw<PASS>()) ? 1 : P::bpt(PASS);
        if (POST_TW) {
            if (PASS == 1) {
                constexpr int R0 = P::radix(0);
#pragma unroll
                for (int c = 0; c < PR::PC; ++c) {
                    const int pi = PR::pi(tid, c);
                    const int j = pi == 0 ? PR::NB / 2 : pi;
#pragma unroll
                    for (int q = 1; q < R0; ++q) tw[P::tw_off(1) + c * (R0 - 1) + (q - 1)] = unit((j * q) % P::M, P::M);
                }
            }
        } else if (mirror_tw<PASS>()) {
#pragma unroll
            for (int c = 0; c < PR::PC; ++c) {
                const int pi = PR::pi(tid, c);
                const int j = pi == 0 ? PR::NB / 2 : pi;
#pragma unroll
                for (int r = 1; r < R; ++r) tw[P::tw_off(PASS) + c * (R - 1) + (r - 1)] = cconj(unit((r * j) % (NS * R), NS * R));
            }
        } else if (PASS > 0) {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                int j = bfly<PASS>(b);
                int k = j % NS;
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    cf u = unit((r * k) % (NS * R), NS * R);
                    tw[P::tw_off(PASS) + b * (R - 1) + (r - 1)] = INV ? u : cconj(u);
                }
            }
        }
    }

    ACIDS_HD void init(int tid_) {
        tid = tid_;
        init_pass<0>();
        init_pass<1>();
        if (P::NP > 2) init_pass<(P::NP > 2 ? 2 : 0)>();
        if (P::NP > 3) init_pass<(P::NP > 3 ? 3 : 0)>();
#pragma unroll
        for (int c = 0; c < PR::PC; ++c) {
            if (COMPACT_WK) {
                // klo / khi may be negative for the self-mirrored pair (khi = NB/2 - (R/2) NB): reduce modulo N
                const cf ul = unit((PR::klo(tid, c) + P::N) % P::N, P::N), uh = unit((PR::khi(tid, c) + P::N) % P::N, P::N);
                wk[2 * c] = INV ? ul : cconj(ul);
                wk[2 * c + 1] = INV ? uh : cconj(uh);
            } else {
#pragma unroll
                for (int s = 0; s < PR::R; ++s) {
                    cf u = unit(PR::k1(tid, c, s), P::N);
                    wk[c * PR::R + s] = INV ? u : cconj(u);
                }
            }
        }
    }

    // element (complex index) that register v[b*R + r] of pass PASS holds BEFORE the butterfly
    template <int PASS>
    ACIDS_HD int in_index(int b, int r) const {
        constexpr int NB = P::nb(PASS);
        return bfly<PASS>(b) + r * NB;
    }
    // element that v[b*R + q] holds AFTER the butterfly of pass PASS
    template <int PASS>
    ACIDS_HD int out_index(int b, int q) const {
        constexpr int R = P::radix(PASS), NS = P::ns(PASS);
        int j = bfly<PASS>(b);
        int k = j % NS;
        return (j - k) * R + k + q * NS;
    }

    template <int PASS>
    ACIDS_HD void butterflies(cf* v) const {
        constexpr int R = P::radix(PASS), B = P::bpt(PASS);
        if (POST_TW) {
            if (PASS == 0) {
#pragma unroll
                for (int b = 0; b < B; ++b) {
                    const cf* t = tw + P::tw_off(1) + (b >> 1) * (R - 1);
                    if ((b & 1) == 0) {
                        const bool sp = PR::pi(tid, b >> 1) == 0;       // butterfly 0: unit twiddles
                        Dft<R, INV>::run(v + b * R);
#pragma unroll
                        for (int q = 1; q < R; ++q) v[b * R + q] = csel(sp, v[b * R + q], cmul(v[b * R + q], t[q - 1]));
                    } else {
                        const cf last = v[b * R + R - 1];                 // inputs shifted by one (the W_R^q factor on the outputs)
#pragma unroll
                        for (int r = R - 1; r > 0; --r) v[b * R + r] = v[b * R + r - 1];
                        v[b * R] = last;
                        Dft<R, INV>::run(v + b * R);
#pragma unroll
                        for (int q = 1; q < R; ++q) v[b * R + q] = cmul(v[b * R + q], cconj(t[q - 1]));
                    }
                }
            } else {
#pragma unroll
                for (int b = 0; b < B; ++b) Dft<R, INV>::run(v + b * R);
            }
            return;
        }
#pragma unroll
        for (int b = 0; b < B; ++b) {
            if (mirror_tw<PASS>()) {
                const cf* t = tw + P::tw_off(PASS) + (b >> 1) * (R - 1);
                if ((b & 1) == 0) {
                    const bool sp = PR::pi(tid, b >> 1) == 0;       // butterfly 0: unit twiddles
#ifdef ACIDS_FWD_PRED_SP      // tuning experiment: predicated multiply instead of multiply + select
                    if (!sp) {
#pragma unroll
                        for (int r = 1; r < R; ++r) v[b * R + r] = cmul(v[b * R + r], t[r - 1]);
                    }
#else
#pragma unroll
                    for (int r = 1; r < R; ++r) v[b * R + r] = csel(sp, v[b * R + r], cmul(v[b * R + r], t[r - 1]));
#endif
                    Dft<R, INV>::run(v + b * R);
                } else {
#pragma unroll
                    for (int r = 1; r < R; ++r) v[b * R + r] = cmul(v[b * R + r], cconj(t[r - 1]));
                    Dft<R, INV>::run(v + b * R);
                    const cf first = v[b * R];                       // outputs rotated by one bin (the W_R^r factor)
#pragma unroll
                    for (int q = 0; q + 1 < R; ++q) v[b * R + q] = v[b * R + q + 1];
                    v[b * R + R - 1] = first;
                }
                continue;
            }
            if (derived_tw<PASS>()) {
                constexpr int RO = (P::ns(PASS) * R) / P::T;       // order of the constant root: W_RO^r
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    const cf a = b == 0 ? v[b * R + r] : mul_root_sw<(derived_tw<PASS>() ? RO : 32), INV>(v[b * R + r], r);
                    v[b * R + r] = cmul(a, tw[P::tw_off(PASS) + (r - 1)]);
                }
            } else if (PASS > 0) {
                constexpr int bs = tw_is_shared<PASS>() ? 0 : 1;
#pragma unroll
                for (int r = 1; r < R; ++r) v[b * R + r] = cmul(v[b * R + r], tw[P::tw_off(PASS) + bs * b * (R - 1) + (r - 1)]);
            }
            Dft<R, INV>::run(v + b * R);
        }
    }

    template <int PASS>
    ACIDS_HD void store(const cf* v, cf* s) const {
        constexpr int R = P::radix(PASS), B = P::bpt(PASS), NS = P::ns(PASS);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            cf* sb = s + swz<P::PADLOG>(out_index<PASS>(b, 0));
            if (NS == 1 && R % 2 == 0) {
                // the R outputs of a first-pass butterfly are one contiguous, 16-byte aligned run: 128-bit stores
#pragma unroll
                for (int q = 0; q < R; q += 2) {
                    ACIDS_EMU_ASSERT(swz<P::PADLOG>(out_index<PASS>(b, q)) == swz<P::PADLOG>(out_index<PASS>(b, 0)) + swzc<P::PADLOG>(q));
                    ACIDS_EMU_ASSERT(swz<P::PADLOG>(out_index<PASS>(b, q + 1)) == swz<P::PADLOG>(out_index<PASS>(b, q)) + 1);
                    ACIDS_EMU_ASSERT(swz<P::PADLOG>(out_index<PASS>(b, q)) % 2 == 0);
                    cf2 w;
                    w.a = v[b * R + q];
                    w.b = v[b * R + q + 1];
                    *reinterpret_cast<cf2*>(sb + swzc<P::PADLOG>(q)) = w;
                }
            } else {
#pragma unroll
                for (int q = 0; q < R; ++q) {
                    ACIDS_EMU_ASSERT(swz<P::PADLOG>(out_index<PASS>(b, q)) == swz<P::PADLOG>(out_index<PASS>(b, 0)) + swzc<P::PADLOG>(q * NS));
                    sb[swzc<P::PADLOG>(q * NS)] = v[b * R + q];
                }
            }
        }
    }

    template <int PASS>
    ACIDS_HD void load(cf* v, const cf* s) const {
        constexpr int R = P::radix(PASS), B = P::bpt(PASS), NB = P::nb(PASS);
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const cf* sb = s + swz<P::PADLOG>(in_index<PASS>(b, 0));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                ACIDS_EMU_ASSERT(swz<P::PADLOG>(in_index<PASS>(b, r)) == swz<P::PADLOG>(in_index<PASS>(b, 0)) + swzc<P::PADLOG>(r * NB));
                v[b * R + r] = sb[swzc<P::PADLOG>(r * NB)];
            }
        }
    }

    // ---- forward untangle (after the paired last pass): v holds Z, produce X -------------------
    // o1[c*R+s] = X[k1(c,s)], o2[c*R+s] = X[M - k1(c,s)]; extra = X[M/2] (meaningful on tid 0 only).
    // Z must be the FFT of z[n] = (x[2n] + i x[2n+1]) / 2: the caller folds the 1/2 of the
    // even/odd split into the analysis window.  DC and Nyquist get an exact +0 imaginary part,
    // like the reference's r2c transform (their phase is 0 or +pi, never -pi).
    ACIDS_HD void untangle_fwd(const cf* v, cf* o1, cf* o2, cf& extra) const {
        constexpr int R = PR::R;
#pragma unroll
        for (int c = 0; c < PR::PC; ++c) {
            const cf* va = v + (2 * c) * R;
            const cf* vb = v + (2 * c + 1) * R;
            const bool sp = PR::pi(tid, c) == 0;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                cf A = va[s], Bv = vb[R - 1 - s];
                if (c == 0) {   // only pair 0 can be the self-mirrored one
                    if (s == 0) Bv = csel(sp, va[0], Bv);
                    else if (s < R / 2) Bv = csel(sp, va[R - s], Bv);
                    else {
                        A = csel(sp, vb[s - R / 2], A);
                        Bv = csel(sp, vb[3 * R / 2 - 1 - s], Bv);
                    }
                }
                cf E = A + cconj(Bv);                     // A + conj(B)
                cf O = mul_mi(A) + mk(Bv.y, Bv.x);        // -i (A - conj(B)) = (A.y + B.y, B.x - A.x)
                cf Pm = untangle_mul(O, c, s);
                cf x1 = E + Pm, x2 = cconj(E - Pm);
                if (c == 0 && s == 0) {
                    x1.y = sp ? 0.f : x1.y;
                    x2.y = sp ? 0.f : x2.y;
                }
                o1[c * R + s] = x1;
                o2[c * R + s] = x2;
            }
            if (c == 0) extra = mk(2.f * va[R / 2].x, -2.f * va[R / 2].y);
        }
    }

    // ---- inverse pre-tangle (before the paired first pass): A = X[k1], Bx = X[M-k1] ------------
    // in1/in2 laid out like o1/o2 above; extra = X[M/2].  Produces v = 2 Z (unnormalised).
    ACIDS_HD void pretangle_inv(const cf* in1, const cf* in2, cf extra, cf* v) const {
        constexpr int R = PR::R;
#pragma unroll
        for (int c = 0; c < PR::PC; ++c) {
            cf za[R], zb[R];
            const bool sp = PR::pi(tid, c) == 0;
#pragma unroll
            for (int s = 0; s < R; ++s) {
                cf A = in1[c * R + s], Bx = in2[c * R + s];
                if (c == 0 && s == 0) {   // c2r ignores the imaginary parts of DC and Nyquist
                    A.y = sp ? 0.f : A.y;
                    Bx.y = sp ? 0.f : Bx.y;
                }
                cf E2 = A + cconj(Bx);                        // A + conj(Bx)
                cf P2 = A - cconj(Bx);                        // A - conj(Bx)
                cf O2 = untangle_mul(P2, c, s);             // conj(W^k) P2
                cf iO = mul_pi(O2);
                za[s] = E2 + iO;
                zb[s] = cconj(E2 - iO);
            }
            cf* va = v + (2 * c) * R;
            cf* vb = v + (2 * c + 1) * R;
#pragma unroll
            for (int q = 0; q < R; ++q) {
                cf ga = za[q], gb = zb[R - 1 - q];
                if (c == 0) {
                    cf sa, sb;
                    if (q == 0) sa = za[0];
                    else if (q < R / 2) sa = za[q];
                    else if (q == R / 2) sa = mk(2.f * extra.x, -2.f * extra.y);
                    else sa = zb[R - q];
                    if (q < R / 2) sb = za[q + R / 2];
                    else sb = zb[3 * R / 2 - 1 - q];
                    ga = csel(sp, sa, ga);
                    gb = csel(sp, sb, gb);
                }
                va[q] = ga;
                vb[q] = gb;
            }
        }
    }
};

}  // namespace acids
