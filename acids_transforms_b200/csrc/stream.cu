// stream.cu — the block-by-block streaming step as ONE kernel per direction, or one kernel for the whole round trip
// (SURVEY.md §8f N2).  The reference's stated use (README.md:4) is real-time processing:
//
//   analysis :  OverlapAdd.forward (oadd.py:70-74: prepend the saved tail, frame with hop)  ->  RealtimeSTFT / RealtimeDGT
//               .forward (stft.py:248-253, dgt.py:284-289: rfft(frame * window))
//   synthesis:  RealtimeSTFT / RealtimeDGT.invert (stft.py:259-266, dgt.py:296-302: irfft(X) * inv_window)  ->
//               OverlapAdd.invert (oadd.py:91-104: carry-in + rectangular overlap-add + carry-out, / gain)
//
// on a few hundred samples per call — each eager kernel runs for microseconds and the step is launch bound.  Here one
// CTA owns one stream (one row of the flattened batch): it stages tail ++ block in shared memory, runs the block's frames
// through window -> rFFT [-> the spectrum goes to the caller, or stays in registers for the round trip] -> irFFT ->
// synthesis window into a shared-memory frame buffer, overlap-adds them with the carried tail in the order of
// acids_ola_stream (carry first, frames ascending) and writes the block and the next carry.  The two pieces of state
// (OverlapAdd.input_buffer / .output_buffer) live in global memory at fixed addresses and are advanced IN PLACE, so a
// step is capturable in a CUDA graph as is.  The arithmetic is that of the batch kernels (same FrameFFT code, same
// operation order): the synthesis half is bit-identical to the eager modules, the analysis half agrees to rounding (ptxas
// contracts packed mul + add pairs into FFMA2 per kernel, so the window multiply fuses into the first butterfly differently).
#include "common.cuh"
#include "plans.cuh"

namespace acids {

struct StreamParams {
    const float* x;          // [B, C] new samples (analysis / round trip)
    const cf* X_in;          // [B, n, F] spectrum (synthesis)
    cf* X_out;               // [B, n, F] spectrum (analysis; optional for the round trip)
    float* out;              // [B, n * hop] samples (synthesis / round trip)
    float* tail;             // [B, keep] OverlapAdd.input_buffer, in place
    float* carry;            // [B, keep] OverlapAdd.output_buffer, in place
    const float* window;     // analysis window [n_fft]
    const float* inv_window; // synthesis window [n_fft]
    int64_t B;
    int C, n, hop, keep;
    float gain;
};

template <class P>
struct StreamCfg {
    static constexpr int THREADS = P::T < 128 ? 128 : P::T;
    static constexpr int G = THREADS / P::T;
};

// forward passes of one frame whose windowed samples sit in v[] in first-pass operand order; leaves Z in v[] (paired layout)
template <class P, int THREADS>
__device__ __forceinline__ void forward_frame(const FrameFFT<P, false>& fft, cf* v, cf* s, int g) {
    auto gsync = [&]() { group_sync<P::T, THREADS>(g); };
    fft.template butterflies<0>(v);
    gsync();
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
}

template <class PF, class PI, bool DO_FWD, bool DO_INV>
__global__ void __launch_bounds__(StreamCfg<PF>::THREADS, 1) stream_step_kernel(const StreamParams p) {
    constexpr int THREADS = StreamCfg<PF>::THREADS, G = StreamCfg<PF>::G;
    constexpr int N = PF::N, M = PF::M, T = PF::T, V = PF::V, F = PF::F;
    static_assert(PI::N == N && PI::T == T, "forward and inverse plans of one n_fft share the thread layout");
    using FF = FrameFFT<PF, false>;
    using FI = FrameFFT<PI, true>;
    using PRF = typename FF::PR;
    using PRI = typename FI::PR;
    static_assert(PRF::R == PRI::R && PRF::NB == PRI::NB && PRF::PC == PRI::PC, "untangle and pre-tangle share the register layout of the spectrum");
    constexpr int RP = PRF::R, NBP = PRF::NB;
    constexpr int R0 = PF::radix(0), B0 = PF::bpt(0);
    constexpr int LP = PI::NP - 1, RL = PI::radix(LP), BL = PI::bpt(LP), NSL = PI::ns(LP);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    cf* const s = reinterpret_cast<cf*>(smem_raw) + (size_t)g * PF::SMEM_CF;
    float* const sbuf = reinterpret_cast<float*>(reinterpret_cast<cf*>(smem_raw) + (size_t)G * PF::SMEM_CF);   // tail ++ block (analysis)
    float* const frames = sbuf + (DO_FWD ? ((p.keep + p.C + 3) & ~3) : 0);                                        // n synthesised frames
    float* const scarry = frames + (DO_INV ? (size_t)p.n * N : 0);                                                // carried tail (synthesis)
    float* const swf = scarry + (DO_INV ? ((p.keep + 3) & ~3) : 0);                                                // analysis window
    float* const swi = swf + (DO_FWD ? N : 0);                                                                     // synthesis window
    const int64_t b = blockIdx.x;
    const int hop = p.hop, keep = p.keep, n = p.n;

    // The step is a chain of dependent latencies (a few microseconds in all), so everything that comes from global memory —
    // the block, both carried buffers, both windows — is requested up front with asynchronous copies straight into shared
    // memory, and the twiddle set-up (sincospi, ~2 us of ALU work) runs while they are in flight.
    auto cp4 = [](float* dst, const float* src) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr_u32(dst)), "l"(src) : "memory");
    };
    if (DO_FWD) {
        // OverlapAdd.forward: saved tail ++ new block; the block's last `keep` samples are the next tail (oadd.py:70-74)
        const float* __restrict__ xb = p.x + b * p.C;
        const float* __restrict__ tb = p.tail + b * keep;
        for (int i = threadIdx.x; i < keep + p.C; i += THREADS) cp4(sbuf + i, i < keep ? tb + i : xb + (i - keep));
        for (int i = threadIdx.x; i < N; i += THREADS) cp4(swf + i, p.window + i);
    }
    if (DO_INV) {
        const float* __restrict__ cb = p.carry + b * keep;
        for (int i = threadIdx.x; i < keep; i += THREADS) cp4(scarry + i, cb + i);
        for (int i = threadIdx.x; i < N; i += THREADS) cp4(swi + i, p.inv_window + i);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    FF ff;
    FI fi;
    if (DO_FWD) ff.init(tid);
    if (DO_INV) fi.init(tid);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (DO_FWD) {
        float* __restrict__ tb = p.tail + b * keep;
        for (int i = threadIdx.x; i < keep; i += THREADS) tb[i] = sbuf[p.C + i];
    }

    const int rounds = (n + G - 1) / G;
    for (int rd = 0; rd < rounds; ++rd) {
        const int f = rd * G + g;
        const bool valid = f < n;
        cf o1[V / 2], o2[V / 2], ex;
        cf v[V];
        if (DO_FWD) {
            // RealtimeSTFT.forward: rfft(frame * window); the 1/2 of the even/odd split is folded into the window like
            // the batch kernel does (stft_fwd_kernel: swin = 0.5 * window)
            const float* __restrict__ fr = sbuf + (valid ? f : 0) * hop;
#pragma unroll
            for (int b0 = 0; b0 < B0; ++b0)
#pragma unroll
                for (int r = 0; r < R0; ++r) {
                    const int e = ff.template in_index<0>(b0, r);
                    const float2 w = make_float2(0.5f * swf[2 * e], 0.5f * swf[2 * e + 1]);
                    v[b0 * R0 + r] = cmul2(mk(fr[2 * e], fr[2 * e + 1]), mk(w.x, w.y));
                }
            forward_frame<PF, THREADS>(ff, v, s, g);
            ff.untangle_fwd(v, o1, o2, ex);
            if (p.X_out && valid) {
                float2* __restrict__ row = reinterpret_cast<float2*>(p.X_out) + (b * n + f) * (int64_t)F;
#pragma unroll
                for (int c = 0; c < PRF::PC; ++c) {
                    float2* lo = row + PRF::klo(tid, c);
                    float2* hi = row + PRF::khi(tid, c);
                    float2* mlo = row + (M - PRF::klo(tid, c));
                    float2* mhi = row + (M - PRF::khi(tid, c));
#pragma unroll
                    for (int q = 0; q < RP; ++q) {
                        (q < RP / 2 ? lo : hi)[q * NBP] = make_float2(o1[c * RP + q].x, o1[c * RP + q].y);
                        (q < RP / 2 ? mlo : mhi)[-q * NBP] = make_float2(o2[c * RP + q].x, o2[c * RP + q].y);
                    }
                }
                if (tid == 0) row[M / 2] = make_float2(ex.x, ex.y);
            }
            group_sync<T, THREADS>(g);     // the inverse transform reuses the exchange buffer
        } else {
            const float2* __restrict__ row = reinterpret_cast<const float2*>(p.X_in) + (b * n + (valid ? f : 0)) * (int64_t)F;
#pragma unroll
            for (int c = 0; c < PRI::PC; ++c) {
                const float2* lo = row + PRI::klo(tid, c);
                const float2* hi = row + PRI::khi(tid, c);
                const float2* mlo = row + (M - PRI::klo(tid, c));
                const float2* mhi = row + (M - PRI::khi(tid, c));
#pragma unroll
                for (int q = 0; q < RP; ++q) {
                    const float2 a = __ldg((q < RP / 2 ? lo : hi) + q * NBP);
                    const float2 d = __ldg((q < RP / 2 ? mlo : mhi) - q * NBP);
                    o1[c * RP + q] = mk(a.x, a.y);
                    o2[c * RP + q] = mk(d.x, d.y);
                }
            }
            const float2 e = __ldg(row + M / 2);
            ex = mk(e.x, e.y);
        }
        if (DO_INV) {
            // RealtimeSTFT.invert: irfft(X) * inv_window, 1/N folded into the window like irfft_frames_kernel
            auto gsync = [&]() { group_sync<T, THREADS>(g); };
            fi.pretangle_inv(o1, o2, ex, v);
            fi.template butterflies<0>(v);
            gsync();
            fi.template store<0>(v, s);
            gsync();
            fi.template load<1>(v, s);
            fi.template butterflies<1>(v);
            if constexpr (PI::NP > 2) {
                gsync();
                fi.template store<1>(v, s);
                gsync();
                fi.template load<2>(v, s);
                fi.template butterflies<2>(v);
            }
            if constexpr (PI::NP > 3) {
                gsync();
                fi.template store<2>(v, s);
                gsync();
                fi.template load<3>(v, s);
                fi.template butterflies<3>(v);
            }
            if (valid) {
                float2* __restrict__ dstf = reinterpret_cast<float2*>(frames + (size_t)f * N);
#pragma unroll
                for (int bb = 0; bb < BL; ++bb) {
                    float2* dst = dstf + (tid + T * bb);
#pragma unroll
                    for (int q = 0; q < RL; ++q) {
                        const int e = tid + T * bb + q * NSL;
                        const float2 w = make_float2(swi[2 * e] * (1.0f / N), swi[2 * e + 1] * (1.0f / N));
                        const cf y = cmul2(v[bb * RL + q], mk(w.x, w.y));
                        dst[q * NSL] = make_float2(y.x, y.y);
                    }
                }
            }
            gsync();       // the next round's first exchange store must not overtake this round's last reads
        }
    }
    if (DO_INV) {
        __syncthreads();
        // OverlapAdd.invert (oadd.py:91-104) in the order of ola_gather_kernel: the carried tail first, then the frames in
        // ascending order; the first n * hop samples leave (/ gain), the last `keep` are the next call's carry
        const int total = (n - 1) * hop + N;
        const int out_len = total - keep;
        float* __restrict__ ob = p.out + b * out_len;
        float* __restrict__ cb = p.carry + b * keep;
        for (int np = threadIdx.x; np < total; np += THREADS) {
            const int num = np - N + hop;
            const int t_lo = num > 0 ? num / hop : 0;
            const int t_hi = min(n - 1, np / hop);
            float acc = np < keep ? scarry[np] : 0.f;
            for (int tt = t_lo; tt <= t_hi; ++tt) acc += frames[(size_t)tt * N + (np - tt * hop)];
            if (np < out_len) ob[np] = acc / p.gain;
            else cb[np - out_len] = acc;
        }
    }
}

template <class PF, class PI>
static int launch_stream(const StreamParams& p, int mode, cudaStream_t st) {
    constexpr int THREADS = StreamCfg<PF>::THREADS, G = StreamCfg<PF>::G;
    const bool fwd = mode != 1, inv = mode != 0;
    size_t smem = (size_t)G * PF::SMEM_CF * sizeof(cf);
    if (fwd) smem += (size_t)((p.keep + p.C + 3) & ~3) * sizeof(float);
    if (inv) smem += ((size_t)p.n * PF::N + ((p.keep + 3) & ~3)) * sizeof(float);
    smem += (size_t)((fwd ? 1 : 0) + (inv ? 1 : 0)) * PF::N * sizeof(float);       // the windows
    ACIDS_REQUIRE(smem <= 227 * 1024, ACIDS_ENOTSUP, "stream step: a block of %d frames of n_fft=%d needs %zu bytes of shared memory (max 232448)",
                  p.n, PF::N, smem);
    void (*kern)(const StreamParams) = mode == 0 ? stream_step_kernel<PF, PI, true, false>
                                       : (mode == 1 ? stream_step_kernel<PF, PI, false, true> : stream_step_kernel<PF, PI, true, true>);
    static PerDevice cache[3][kMaxDevices];
    size_t& reserved = per_device(cache[mode]).reserved;
    if (smem > reserved) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            set_error("stream step: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
            return ACIDS_ECUDA;
        }
        reserved = smem;
    }
    if (p.B == 0) return ACIDS_OK;
    kern<<<(unsigned)p.B, THREADS, smem, st>>>(p);
    ACIDS_CHECK_LAUNCH("stream step");
    return ACIDS_OK;
}

#define ACIDS_STREAM_SWITCH(n_fft, EXPR)                                                       \
    switch (n_fft) {                                                                           \
        case 32: { using PF_ = Fwd32; using PI_ = Inv32; EXPR; } break;                        \
        case 64: { using PF_ = Fwd64; using PI_ = Inv64; EXPR; } break;                        \
        case 128: { using PF_ = Fwd128; using PI_ = Inv128; EXPR; } break;                     \
        case 256: { using PF_ = Fwd256; using PI_ = Inv256; EXPR; } break;                     \
        case 512: { using PF_ = Fwd512; using PI_ = Inv512; EXPR; } break;                     \
        case 1024: { using PF_ = Fwd1024; using PI_ = Inv1024; EXPR; } break;                  \
        case 2048: { using PF_ = Fwd2048; using PI_ = Inv2048; EXPR; } break;                  \
        case 4096: { using PF_ = Fwd4096; using PI_ = Inv4096; EXPR; } break;                  \
        case 8192: { using PF_ = Fwd8192; using PI_ = Inv8192; EXPR; } break;                  \
        case 16384: { using PF_ = Fwd16384; using PI_ = Inv16384; EXPR; } break;               \
        default:                                                                               \
            set_error("n_fft=%d is not supported (power of two in [32, 16384])", n_fft);       \
            return ACIDS_ENOTSUP;                                                              \
    }

static int stream_common(StreamParams& p, int64_t B, int64_t n, int n_fft, int hop, int64_t keep) {
    ACIDS_REQUIRE(B >= 0 && B < ((int64_t)1 << 31) && n >= 1 && n < (1 << 20) && hop > 0 && hop <= n_fft, ACIDS_EINVAL,
                  "stream step: bad sizes B=%lld frames=%lld hop=%d", (long long)B, (long long)n, hop);
    ACIDS_REQUIRE(keep == n_fft - hop, ACIDS_EINVAL, "stream step: the carried tail must be n_fft - hop = %d samples (got %lld)", n_fft - hop, (long long)keep);
    ACIDS_REQUIRE(n * (int64_t)hop >= keep, ACIDS_ENOTSUP, "stream step: a block (%lld samples) shorter than the carried tail (%lld) is not supported",
                  (long long)(n * hop), (long long)keep);
    p.B = B; p.n = (int)n; p.hop = hop; p.keep = (int)keep; p.C = (int)(n * hop);
    return ACIDS_OK;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int acids_stream_analysis(const float* x, int64_t B, int64_t n, int n_fft, int hop, const float* window, float* tail,
                                     float* X_out, void* stream) {
    StreamParams p{};
    ACIDS_REQUIRE(x && window && tail && X_out, ACIDS_EINVAL, "stream_analysis: NULL pointer");
    int rc = stream_common(p, B, n, n_fft, hop, n_fft - hop);
    if (rc) return rc;
    p.x = x; p.window = window; p.tail = tail; p.X_out = reinterpret_cast<cf*>(X_out);
    ACIDS_STREAM_SWITCH(n_fft, return (launch_stream<PF_, PI_>(p, 0, static_cast<cudaStream_t>(stream))));
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_stream_synthesis(const float* X, int64_t B, int64_t n, int n_fft, int hop, const float* inv_window, float gain,
                                      float* carry, float* out, void* stream) {
    StreamParams p{};
    ACIDS_REQUIRE(X && inv_window && carry && out, ACIDS_EINVAL, "stream_synthesis: NULL pointer");
    ACIDS_REQUIRE(gain != 0.f, ACIDS_EINVAL, "stream_synthesis: zero gain");
    int rc = stream_common(p, B, n, n_fft, hop, n_fft - hop);
    if (rc) return rc;
    p.X_in = reinterpret_cast<const cf*>(X); p.inv_window = inv_window; p.gain = gain; p.carry = carry; p.out = out;
    ACIDS_STREAM_SWITCH(n_fft, return (launch_stream<PF_, PI_>(p, 1, static_cast<cudaStream_t>(stream))));
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_stream_roundtrip(const float* x, int64_t B, int64_t n, int n_fft, int hop, const float* window,
                                      const float* inv_window, float gain, float* tail, float* carry, float* X_out, float* out, void* stream) {
    StreamParams p{};
    ACIDS_REQUIRE(x && window && inv_window && tail && carry && out, ACIDS_EINVAL, "stream_roundtrip: NULL pointer");
    ACIDS_REQUIRE(gain != 0.f, ACIDS_EINVAL, "stream_roundtrip: zero gain");
    int rc = stream_common(p, B, n, n_fft, hop, n_fft - hop);
    if (rc) return rc;
    p.x = x; p.window = window; p.inv_window = inv_window; p.gain = gain; p.tail = tail; p.carry = carry;
    p.X_out = reinterpret_cast<cf*>(X_out); p.out = out;
    ACIDS_STREAM_SWITCH(n_fft, return (launch_stream<PF_, PI_>(p, 2, static_cast<cudaStream_t>(stream))));
    return ACIDS_OK;
}
