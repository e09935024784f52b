// istft.cu — inverse real FFT fused with synthesis-window overlap-add and envelope normalisation.
// Rows A15, A16, A17 of SURVEY.md §8(a).
//
// Atomic-free segmented accumulation: a CTA owns a contiguous span of OUTPUT samples of one clip
// (a chunk of frames [ta, tb)), transforms the frames that touch that span (ceil(n_fft/hop) - 1 halo
// frames at the chunk start are recomputed rather than exchanged), parks each windowed frame in a
// shared-memory ring slot, and after every round of G frames gathers the samples that have now seen
// all of their contributing frames: y[n] = sum_t ring[t][n - t*hop] / sum_t g^2[n - t*hop], summed in
// ascending t like torch.istft's col2im.  Nothing is scattered, no atomics, no global scratch.
//
// HBM traffic per frame: (n_fft/2+1)*8 B in, hop*4 B out.
#include <stdlib.h>
#include "common.cuh"
#include "plans.cuh"
#include <type_traits>

namespace acids {

struct InvParams {
    const cf* X;
    int64_t B;
    int n_frames;
    int hop;
    const float* window;
    float* out;
    int64_t out_len;        // hop * (n_frames - 1) for the centred istft
    int trim;               // n_fft / 2
    int chunk_frames;
    int chunks_per_clip;
    int ovc;                // ceil(n_fft / hop)
    int ring;               // G + ovc - 1 slots (1 in accumulate mode)
    int accum;              // one frame per CTA round, n_fft % hop == 0, hop % 4 == 0: the ring holds n_fft partial sums, not frames
};

// One spectrum row into registers, in the operand order of the paired first pass (bins k and M - k together).
template <class P>
__device__ __forceinline__ void load_spectrum(const FrameFFT<P, true>& fft, const cf* __restrict__ row, bool valid, cf* i1, cf* i2, cf& ex) {
    constexpr int M = P::M, V = P::V;
    using PR = typename FrameFFT<P, true>::PR;
    constexpr int RP = PR::R, NBP = PR::NB;
    if (valid) {
        const float2* __restrict__ r2 = reinterpret_cast<const float2*>(row);
#pragma unroll
        for (int c = 0; c < PR::PC; ++c) {
            // bins k = base + q*NB and M - k: two per-thread bases, compile-time offsets
            const float2* lo = r2 + PR::klo(fft.tid, c);
            const float2* hi = r2 + PR::khi(fft.tid, c);
            const float2* mlo = r2 + (M - PR::klo(fft.tid, c));
            const float2* mhi = r2 + (M - PR::khi(fft.tid, c));
#pragma unroll
            for (int q = 0; q < RP; ++q) {
                const float2 a = ldg_stream2((q < RP / 2 ? lo : hi) + q * NBP);
                const float2 d = ldg_stream2((q < RP / 2 ? mlo : mhi) - q * NBP);
                i1[c * RP + q] = mk(a.x, a.y);
                i2[c * RP + q] = mk(d.x, d.y);
            }
        }
        const float2 e = __ldg(r2 + M / 2);
        ex = mk(e.x, e.y);
    } else {
#pragma unroll
        for (int i = 0; i < V / 2; ++i) i1[i] = i2[i] = mk(0.f, 0.f);
        ex = mk(0.f, 0.f);
    }
}

// Spectrum row (already in registers) -> N time samples, left in v[] in the natural order of the last pass.
template <class P, int THREADS>
__device__ __forceinline__ void inverse_frame(const FrameFFT<P, true>& fft, const cf* i1, const cf* i2, cf ex, cf* v, cf* s, int g) {
    constexpr int T = P::T;
    auto gsync = [&]() { group_sync<T, THREADS>(g); };
    fft.pretangle_inv(i1, i2, ex, v);
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

#ifndef ACIDS_INV_MINB_SMALL
#define ACIDS_INV_MINB_SMALL 3   // 168 registers: at 4 CTAs / SM (128 registers) the prefetched spectrum spills and costs 30 %
#endif
// launch shape per plan (see FwdCfg): small frame groups run 128-thread CTAs
template <class P>
struct InvCfg {
    // n_fft = 4096 (T = 128): one frame per CTA.  With two frames the ring (5 x 16 KB) leaves a single 8-warp CTA per
    // SM whose warps all sit in the same phase; one frame per CTA (ring 4 x 16 KB, 103 KB) gives two independent CTAs,
    // and the accumulate mode of the kernel (16 KB of partial sums instead of the ring, 55 KB) three at 168 registers.
#ifndef ACIDS_INV_T128_THREADS
#define ACIDS_INV_T128_THREADS 128
#endif
#ifndef ACIDS_INV_SMALL_THREADS
#define ACIDS_INV_SMALL_THREADS 128
#endif
    // n_fft = 2048 (T = 64): 128-thread CTAs at 3 per SM (168 registers); at 256 threads x 2 the 128-register budget spills
#ifndef ACIDS_INV_T16_THREADS
#define ACIDS_INV_T16_THREADS 128
#endif
#ifndef ACIDS_INV_T16_MINB
#define ACIDS_INV_T16_MINB 3
#endif
    static constexpr int THREADS = (P::N == 1024 && P::T == 16) ? ACIDS_INV_T16_THREADS : P::T <= 32 ? ACIDS_INV_SMALL_THREADS
                                              : (P::T > 256 ? P::T : (P::T == 128 ? ACIDS_INV_T128_THREADS : (P::T == 64 ? 128 : 256)));
    static constexpr int MINB = (P::N == 1024 && P::T == 16) ? ACIDS_INV_T16_MINB : P::T <= 32 ? ACIDS_INV_MINB_SMALL * (128 / ACIDS_INV_SMALL_THREADS)
                                           : (P::T == 64 || P::T == 128 ? 3 : (P::T <= 256 ? 2 : 1));
    static constexpr int G = THREADS / P::T;
    // RX (the one-exchange n_fft = 1024 plan, two frames per warp): a ring slot doubles as the exchange buffer of the frame
    // that will be parked in it — the slots a round overwrites hold frames the previous round's gather has finished with —
    // so 8 frames per round need 11 padded slots (51 KB) instead of 8 exchange buffers + 11 slots (81 KB): 4 CTAs per SM.
    static constexpr bool RX = P::N == 1024 && P::T == 16;
    static constexpr int SLOTF = RX ? 2 * P::SMEM_CF : P::N;      // floats per ring slot
    static constexpr size_t exch_bytes() { return RX ? 0 : (((size_t)G * P::SMEM_CF * sizeof(cf) + 15) & ~(size_t)15); }
};

// plan of the overlap-add kernel of an n_fft (the per-frame kernels keep the table's plan)
template <class P> struct OlaPlan { using type = P; };
// ACIDS_INV1024_T16: the one-exchange plan for the n_fft = 1024 overlap-add kernel.  Measured on B200: 177
// instead of 258 shared-memory wavefronts and 746 instead of 851 instructions per frame, but 1.29-1.53 ms against 1.11 ms —
// 32 values per thread need 156-168 registers (12 warps / SM) and each warp issues its 32 exchange accesses back to back
// (short-scoreboard / MIO-throttle bound at 30-38 % issue utilisation); opt-in until that is solved.
#ifdef ACIDS_INV1024_T16
template <> struct OlaPlan<Inv1024> { using type = Inv1024T16; };
#endif

template <class P>
__global__ void __launch_bounds__(InvCfg<P>::THREADS, InvCfg<P>::MINB) istft_ola_kernel(const InvParams p) {
    constexpr int THREADS = InvCfg<P>::THREADS;
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V, G = THREADS / T;
    constexpr int LP = P::NP - 1, RL = P::radix(LP), BL = P::bpt(LP), NSL = P::ns(LP);
    using FFT = FrameFFT<P, true>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    constexpr bool RX = InvCfg<P>::RX;
    constexpr int SLOTF = InvCfg<P>::SLOTF;
#ifndef ACIDS_INV_T16_PREFETCH
#define ACIDS_INV_T16_PREFETCH 1
#endif
#ifndef ACIDS_INV_PREFETCH
#define ACIDS_INV_PREFETCH 1
#endif
    constexpr bool PREFETCH = RX ? ACIDS_INV_T16_PREFETCH : ACIDS_INV_PREFETCH;      // next round's spectrum rows into registers before the gather
    cf* s = reinterpret_cast<cf*>(smem_raw) + (size_t)g * P::SMEM_CF;      // !RX: the group's exchange buffer
    constexpr size_t kExch = InvCfg<P>::exch_bytes();                       // keeps the ring 16-byte aligned
    float* ring = reinterpret_cast<float*>(smem_raw + kExch);
    float2* swin = reinterpret_cast<float2*>(ring + (size_t)p.ring * SLOTF);   // synthesis window pairs with irfft's 1/N folded in
    float* inv_env = reinterpret_cast<float*>(swin + M);   // 1 / sum_i g^2[r + i hop]: the interior envelope
    // window^2 for the envelope (N is a power of two: scaling by N undoes the folded 1/N exactly)
    const float* swf = reinterpret_cast<const float*>(swin);
    auto g2 = [&](int i) { const float w = swf[i] * (float)N; return w * w; };

    FFT fft;
    fft.init(tid);
    const int hop = p.hop;
    for (int n = threadIdx.x; n < M; n += THREADS)
        swin[n] = make_float2(__ldg(p.window + 2 * n) * (1.0f / N), __ldg(p.window + 2 * n + 1) * (1.0f / N));
    __syncthreads();
    const bool aligned = (N % hop) == 0;             // integer overlap: every hop segment sees the same frames
    const int ov = N / hop;                          // (aligned) frames per hop segment
    const int out_len = (int)p.out_len;
    // Accumulate mode (G == 1 plans, i.e. n_fft >= 4096): instead of the last ov windowed frames the CTA keeps ONE
    // n_fft-sample buffer of partial overlap-add sums, sample n of frame t at (t hop + n) mod n_fft.  A frame's last
    // hop samples open a new hop segment (assigned), the rest is added to what earlier frames left; after frame t the
    // segment t is complete and is emitted.  Same summation order as the frame ring (ascending frames, like col2im),
    // a quarter of the shared memory at hop = n_fft / 4: three CTAs per SM instead of two at n_fft = 4096.
    const bool accum = G == 1 && p.accum != 0;
    // ring slot arithmetic without modulo: arguments stay within (-ring, 2 ring)
    auto wrap = [&](int sl) { return sl < 0 ? sl + p.ring : (sl >= p.ring ? sl - p.ring : sl); };
    if (aligned)
        for (int r = threadIdx.x; r < hop; r += THREADS) {
            float e = 0.f;
            for (int i = ov - 1; i >= 0; --i) e += g2(r + i * hop);     // ascending frame order, like col2im
            inv_env[r] = 1.0f / e;
        }

    // Persistent grid: CTA c owns a contiguous run of units (G frames of one clip) of the whole batch, i.e. a sequence
    // of segments (clip, [ta, tb)).  Only a segment that starts inside a clip recomputes halo frames, and there is at
    // most one such segment per CTA.
    const int nT = p.n_frames;
    const int upc = (nT + G - 1) / G;
    const int64_t total_units = p.B * upc;
    int64_t u0 = total_units * blockIdx.x / gridDim.x;
    const int64_t u1 = total_units * (blockIdx.x + 1) / gridDim.x;
    int64_t clip = u0 / upc;
    int ua = (int)(u0 - clip * upc);
    for (; u0 < u1; ++clip, ua = 0) {
    const int ub = (int)min((int64_t)upc, ua + (u1 - u0));
    u0 += ub - ua;
    const int ta = ua * G;
    const int tb = min(nT, ub * G);
    // padded-sample span owned by this CTA
    const int64_t own_lo = (int64_t)ta * hop;
    const int64_t own_hi = (tb == nT) ? (int64_t)(nT - 1) * hop + N : (int64_t)tb * hop;
    const int t_start = max(0, ta - (p.ovc - 1));
    int64_t emitted = own_lo;
    const cf* __restrict__ Xc = p.X + clip * (int64_t)nT * P::F;
    float* __restrict__ outc = p.out + clip * p.out_len;

    // slot of frame tr in the ring (frames take consecutive slots, a segment starts at slot 0); hop segments emitted
    int base_slot = 0;
    int q_next = ta;
    const int q_own_hi = (tb == nT) ? nT - 1 + ov : tb;
    cf i1[V / 2], i2[V / 2], ex;
    if (PREFETCH) load_spectrum<P>(fft, Xc + (int64_t)(t_start + g) * P::F, t_start + g < tb, i1, i2, ex);
    for (int tr = t_start; tr < tb; tr += G) {
        const int t = tr + g;
        const bool valid = t < tb;
        if (!PREFETCH) load_spectrum<P>(fft, Xc + (int64_t)t * P::F, valid, i1, i2, ex);
        cf v[V];
        if (RX) {
            // the slot this frame will be parked in is its exchange buffer; the group leaves the buffer before it parks
            s = reinterpret_cast<cf*>(ring + (size_t)wrap(base_slot + g) * SLOTF);
            inverse_frame<P, THREADS>(fft, i1, i2, ex, v, s, g);
            group_sync<T, THREADS>(g);
        } else {
            inverse_frame<P, THREADS>(fft, i1, i2, ex, v, s, g);
        }
        if (G == 1 && accum) {
            const int rot = (int)(((int64_t)t * hop) & (N - 1));
            const bool first = tr == t_start;        // nothing of this run is in the buffer yet: every sample opens its segment
            const int tail = N - hop;
#pragma unroll
            for (int b = 0; b < BL; ++b) {
                const float2* wv = swin + (tid + T * b);
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const int n = 2 * (tid + T * b + q * NSL);
                    const float2 w = wv[q * NSL];
                    const cf y = cmul2(v[b * RL + q], mk(w.x, w.y));
                    float2* d = reinterpret_cast<float2*>(ring + ((rot + n) & (N - 1)));
                    if (first || n >= tail) {
                        *d = make_float2(y.x, y.y);
                    } else {
                        const float2 o = *d;
                        *d = make_float2(o.x + y.x, o.y + y.y);
                    }
                }
            }
        } else if (valid) {
            float2* slot = reinterpret_cast<float2*>(ring + (size_t)wrap(base_slot + g) * SLOTF);
#pragma unroll
            for (int b = 0; b < BL; ++b) {
                // last-pass outputs sit at n = (tid + T b) + q * Ns: one base, compile-time offsets
                float2* dst = slot + (tid + T * b);
                const float2* wv = swin + (tid + T * b);
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const float2 w = wv[q * NSL];
                    const cf y = cmul2(v[b * RL + q], mk(w.x, w.y));
                    dst[q * NSL] = make_float2(y.x, y.y);
                }
            }
        }
        // the next round's spectrum rows land while this round is gathered
        if (PREFETCH && tr + G < tb) load_spectrum<P>(fft, Xc + (int64_t)(t + G) * P::F, t + G < tb, i1, i2, ex);
        __syncthreads();
        // gather every sample that no later frame can touch
        const int t_done = min(tr + G, tb) - 1;
        const int64_t hi = (t_done == nT - 1) ? own_hi : min(own_hi, (int64_t)(t_done + 1) * hop);
        if (aligned) {
            // whole hop segments; segment q gets frames [max(0, q - ov + 1), min(nT - 1, q)].  All bookkeeping is
            // incremental (q_next, base_slot): no division or modulo per sample.
            const int q0 = q_next;
            const int q1 = (t_done == nT - 1) ? nT - 1 + ov : min(q_own_hi, t_done + 1);
            if (q1 > q_next) q_next = q1;
            if ((hop & 3) == 0) {
                // One warp per hop segment: which frames overlap it, where they sit in the ring and whether it is an
                // interior segment are warp-uniform and computed once; a lane then takes four consecutive samples at a
                // time (16-byte shared loads, one 16-byte streaming store).
                constexpr int NW = THREADS / 32;
                const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
                // fewer segments than warps (large n_fft: few frames per round): split every segment into `parts` ranges
                int lp = 0;                                              // log2(parts)
                while (((q1 - q0) << (lp + 1)) <= NW && (hop & ((256 << lp) - 1)) == 0) ++lp;
                const int span = hop >> lp;
                for (int idx = warp; idx < ((q1 - q0) << lp); idx += NW) {
                    const int q = q0 + (idx >> lp);
                    const int part = idx & ((1 << lp) - 1);
                    const int r_lo = part * span + (lane << 2), r_hi = part * span + span;
                    const int t_lo = max(0, q - ov + 1), t_hi = min(nT - 1, q);
                    const int slot0 = wrap(base_slot + (t_lo - tr));
                    const int n0 = q * hop - p.trim;
                    if (G == 1 && accum) {
                        // the finished sums of segment q; interior segments take the precomputed inverse envelope
                        const float* seg = ring + (int)(((int64_t)q * hop) & (N - 1));
                        const bool interior = t_hi - t_lo + 1 == ov;
                        for (int r = r_lo; r < r_hi; r += 128) {
                            const float4 a = *reinterpret_cast<const float4*>(seg + r);
                            float yy[4];
                            if (interior) {
                                const float4 e = *reinterpret_cast<const float4*>(inv_env + r);
                                yy[0] = a.x * e.x; yy[1] = a.y * e.y; yy[2] = a.z * e.z; yy[3] = a.w * e.w;
                            } else {
                                float4 env = make_float4(0.f, 0.f, 0.f, 0.f);
                                int o2 = r + (q - t_lo) * hop;
                                for (int tt = t_lo; tt <= t_hi; ++tt, o2 -= hop) {
                                    env.x += g2(o2); env.y += g2(o2 + 1); env.z += g2(o2 + 2); env.w += g2(o2 + 3);
                                }
                                yy[0] = a.x / env.x; yy[1] = a.y / env.y; yy[2] = a.z / env.z; yy[3] = a.w / env.w;
                            }
                            const int n = n0 + r;
                            if (n >= 0 && n + 3 < out_len) {
                                asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(outc + n), "f"(yy[0]),
                                             "f"(yy[1]), "f"(yy[2]), "f"(yy[3]) : "memory");
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (n + j >= 0 && n + j < out_len) stg_stream1(outc + n + j, yy[j]);
                            }
                        }
                        continue;
                    }
                    if (t_hi - t_lo + 1 == ov) {
                        // interior: ov frames, ascending frame order like torch.istft's col2im; loads issued together
                        auto run = [&](auto ovc_) {
                            constexpr int OV = decltype(ovc_)::value;
                            int off[OV];
                            int slot = slot0;
#pragma unroll
                            for (int i = 0; i < OV; ++i) {
                                off[i] = slot * SLOTF + (OV - 1 - i) * hop;
                                slot = (slot + 1 == p.ring) ? 0 : slot + 1;
                            }
                            for (int r = r_lo; r < r_hi; r += 128) {
                                float4 f[OV];
#pragma unroll
                                for (int i = 0; i < OV; ++i) f[i] = *reinterpret_cast<const float4*>(ring + off[i] + r);
                                const float4 e = *reinterpret_cast<const float4*>(inv_env + r);
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                                for (int i = 0; i < OV; ++i) { acc.x += f[i].x; acc.y += f[i].y; acc.z += f[i].z; acc.w += f[i].w; }
                                const int n = n0 + r;
                                if (n >= 0 && n + 3 < out_len) {
                                    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(outc + n), "f"(acc.x * e.x),
                                                 "f"(acc.y * e.y), "f"(acc.z * e.z), "f"(acc.w * e.w) : "memory");
                                } else {
                                    const float yy[4] = {acc.x * e.x, acc.y * e.y, acc.z * e.z, acc.w * e.w};
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (n + j >= 0 && n + j < out_len) stg_stream1(outc + n + j, yy[j]);
                                }
                            }
                        };
                        if (ov == 4) run(std::integral_constant<int, 4>{});
                        else if (ov == 2) run(std::integral_constant<int, 2>{});
                        else if (ov == 8) run(std::integral_constant<int, 8>{});
                        else if (ov == 1) run(std::integral_constant<int, 1>{});
                        else if (ov == 16) run(std::integral_constant<int, 16>{});
                        else {
                            for (int r = r_lo; r < r_hi; r += 128) {
                                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                                int slot = slot0;
                                for (int i = 0; i < ov; ++i) {
                                    const float4 f = *reinterpret_cast<const float4*>(ring + slot * SLOTF + (ov - 1 - i) * hop + r);
                                    acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
                                    slot = (slot + 1 == p.ring) ? 0 : slot + 1;
                                }
                                const float4 e = *reinterpret_cast<const float4*>(inv_env + r);
                                const float yy[4] = {acc.x * e.x, acc.y * e.y, acc.z * e.z, acc.w * e.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (n0 + r + j >= 0 && n0 + r + j < out_len) stg_stream1(outc + n0 + r + j, yy[j]);
                            }
                        }
                    } else {
                        // clip edges: fewer frames, exact envelope
                        for (int r = r_lo; r < r_hi; r += 128) {
                            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f), env = make_float4(0.f, 0.f, 0.f, 0.f);
                            int slot = slot0;
                            int o2 = r + (q - t_lo) * hop;
                            for (int tt = t_lo; tt <= t_hi; ++tt, o2 -= hop) {
                                const float4 f = *reinterpret_cast<const float4*>(ring + slot * SLOTF + o2);
                                acc.x += f.x; acc.y += f.y; acc.z += f.z; acc.w += f.w;
                                env.x += g2(o2); env.y += g2(o2 + 1); env.z += g2(o2 + 2); env.w += g2(o2 + 3);
                                slot = (slot + 1 == p.ring) ? 0 : slot + 1;
                            }
                            const float yy[4] = {acc.x / env.x, acc.y / env.y, acc.z / env.z, acc.w / env.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (n0 + r + j >= 0 && n0 + r + j < out_len) stg_stream1(outc + n0 + r + j, yy[j]);
                        }
                    }
                }
            } else {
                for (int q = q0; q < q1; ++q) {
                    const int t_lo = max(0, q - ov + 1), t_hi = min(nT - 1, q);
                    const int slot0 = wrap(base_slot + (t_lo - tr));
                    for (int r = threadIdx.x; r < hop; r += THREADS) {
                        float acc = 0.f, env = 0.f;
                        int slot = slot0, off = r + (q - t_lo) * hop;
                        for (int tt = t_lo; tt <= t_hi; ++tt) {
                            acc += ring[slot * SLOTF + off];
                            env += g2(off);
                            off -= hop;
                            slot = (slot + 1 == p.ring) ? 0 : slot + 1;
                        }
                        const int64_t n = (int64_t)q * hop + r - p.trim;
                        if (n >= 0 && n < p.out_len) stg_stream1(outc + n, acc / env);
                    }
                }
            }
        } else {
            for (int64_t np = emitted + threadIdx.x; np < hi; np += THREADS) {
                const int64_t num = np - N + hop;
                const int t_lo = num > 0 ? (int)(num / hop) : 0;
                const int t_hi = (int)min((int64_t)nT - 1, np / hop);
                float acc = 0.f, env = 0.f;
                for (int tt = t_lo; tt <= t_hi; ++tt) {
                    const int m = (int)(np - (int64_t)tt * hop);
                    acc += ring[(size_t)wrap(base_slot + (tt - tr)) * SLOTF + m];
                    env += g2(m);
                }
                const int64_t n = np - p.trim;
                if (n >= 0 && n < p.out_len) stg_stream1(outc + n, acc / env);
            }
        }
        if (hi > emitted) emitted = hi;
        base_slot = wrap(base_slot + G);
        __syncthreads();
    }
    }
}

// ---- per-frame inverse without overlap-add: irfft(X) * window (RealtimeSTFT.invert) ------------
struct InvFramesParams {
    const cf* X;
    int64_t rows;
    const float* window;
    float* out;
};

template <class P>
__global__ void __launch_bounds__(InvCfg<P>::THREADS, InvCfg<P>::MINB) irfft_frames_kernel(const InvFramesParams p) {
    constexpr int THREADS = InvCfg<P>::THREADS;
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V, G = THREADS / T;
    constexpr int LP = P::NP - 1, RL = P::radix(LP), BL = P::bpt(LP), NSL = P::ns(LP);
    using FFT = FrameFFT<P, true>;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / T, tid = threadIdx.x % T;
    cf* s = reinterpret_cast<cf*>(smem_raw) + (size_t)g * P::SMEM_CF;
    float2* swin = reinterpret_cast<float2*>(reinterpret_cast<cf*>(smem_raw) + (size_t)G * P::SMEM_CF);
    FFT fft;
    fft.init(tid);
    for (int n = threadIdx.x; n < M; n += THREADS)
        swin[n] = make_float2(__ldg(p.window + 2 * n) * (1.0f / N), __ldg(p.window + 2 * n + 1) * (1.0f / N));
    __syncthreads();
    const int64_t units = (p.rows + G - 1) / G;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int64_t r = u * G + g;
        const bool valid = r < p.rows;
        cf v[V];
        {
            cf i1[V / 2], i2[V / 2], ex;
            load_spectrum<P>(fft, p.X + r * (int64_t)P::F, valid, i1, i2, ex);
            inverse_frame<P, THREADS>(fft, i1, i2, ex, v, s, g);
        }
        if (valid) {
#pragma unroll
            for (int b = 0; b < BL; ++b) {
                float2* dst = reinterpret_cast<float2*>(p.out + r * (int64_t)N) + (tid + T * b);
                const float2* wv = swin + (tid + T * b);
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    const float2 w = wv[q * NSL];
                    stg_stream2(dst + q * NSL, v[b * RL + q].x * w.x, v[b * RL + q].y * w.y);
                }
            }
        }
        group_sync<T, THREADS>(g);   // next frame's store<0> must not overtake this frame's last reads
    }
}

// ---- overlap-add from frames in global memory (fallback for huge n_fft, and the streaming OLA) ----
struct OlaParams {
    const float* frames;    // [B, n, N]
    int64_t B;
    int n;
    int N, hop;
    const float* window;    // g for the envelope, or nullptr (no envelope)
    const float* carry_in;  // [B, keep] or nullptr
    int64_t keep;           // carried tail length (streaming) or 0
    float gain;             // divide by this when window == nullptr
    int64_t trim;           // leading samples to drop
    float* out;             // [B, out_len]
    int64_t out_len;
    float* carry_out;       // [B, keep] or nullptr
};

__global__ void ola_gather_kernel(const OlaParams p) {
    const int64_t total = (int64_t)(p.n - 1) * p.hop + p.N;
    const int64_t clip = blockIdx.y;
    const float* __restrict__ fr = p.frames + clip * (int64_t)p.n * p.N;
    for (int64_t np = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; np < total; np += (int64_t)gridDim.x * blockDim.x) {
        const int64_t num = np - p.N + p.hop;
        const int t_lo = num > 0 ? (int)(num / p.hop) : 0;
        const int t_hi = (int)min((int64_t)p.n - 1, np / p.hop);
        float acc = (p.carry_in && np < p.keep) ? __ldg(p.carry_in + clip * p.keep + np) : 0.f;
        float env = 0.f;
        for (int tt = t_lo; tt <= t_hi; ++tt) {
            const int m = (int)(np - (int64_t)tt * p.hop);
            acc += __ldg(fr + (int64_t)tt * p.N + m);
            if (p.window) {
                const float w = __ldg(p.window + m);
                env += w * w;
            }
        }
        const int64_t n = np - p.trim;
        if (n >= 0 && n < p.out_len) p.out[clip * p.out_len + n] = p.window ? acc / env : acc / p.gain;
        else if (p.carry_out && n >= p.out_len && n - p.out_len < p.keep) p.carry_out[clip * p.keep + (n - p.out_len)] = acc;
    }
}

template <class P>
struct InvLaunch {
    static constexpr int THREADS = InvCfg<P>::THREADS;
    static constexpr int G = InvCfg<P>::G;
    static bool accumulates(int hop) { return G == 1 && (P::N % hop) == 0 && (hop & 3) == 0; }
    static size_t smem_ola(int ovc, int hop) {
        // exchange buffers | frame ring (or the partial sums) | window pairs | interior inverse envelope
        return InvCfg<P>::exch_bytes() + (size_t)(accumulates(hop) ? 1 : G + ovc - 1) * InvCfg<P>::SLOTF * sizeof(float) +
               (size_t)P::M * sizeof(float2) + (size_t)hop * sizeof(float);
    }
    static int ola(InvParams p, cudaStream_t st) {
        auto kern = istft_ola_kernel<P>;
        const size_t smem = smem_ola(p.ovc, p.hop);
        static PerDevice cache[kMaxDevices];
        PerDevice& pd = per_device(cache);
        size_t& reserved = pd.reserved;
        if (smem > reserved) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                set_error("istft_ola: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(cudaGetLastError()));
                return ACIDS_ECUDA;
            }
            reserved = smem;
        }
        p.accum = accumulates(p.hop) ? 1 : 0;
        p.ring = p.accum ? 1 : G + p.ovc - 1;
        int& ctas_per_sm = pd.ctas_per_sm;
        size_t& occ_smem = pd.occ_smem;
        if (ctas_per_sm == 0 || occ_smem != smem) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, smem);
            ctas_per_sm = nb > 0 ? nb : 1;
            occ_smem = smem;
        }
        // persistent grid over a contiguous partition of all (clip, unit) pairs; a CTA's run should be long compared
        // with the ovc - 1 halo frames it recomputes at its start
        const int64_t total_units = p.B * (((int64_t)p.n_frames + G - 1) / G);
        if (total_units == 0) return ACIDS_OK;
        int64_t grid = (int64_t)num_sms() * ctas_per_sm;
#ifdef ACIDS_INV_GRID_ENV      // tuning experiment: cap the persistent grid from the environment (CTAs per SM)
        if (const char* e = getenv("ACIDS_INV_CTAS_PER_SM")) grid = (int64_t)num_sms() * atoi(e);
#endif
        const int64_t min_units = 4 * ((p.ovc + G - 1) / G) + 1;
        if (grid > total_units / min_units) grid = total_units / min_units;
        if (grid < 1) grid = 1;
        kern<<<(unsigned)grid, THREADS, smem, st>>>(p);
        ACIDS_CHECK_LAUNCH("istft_ola");
        return ACIDS_OK;
    }
    static int frames(const InvFramesParams& p, cudaStream_t st) {
        auto kern = irfft_frames_kernel<P>;
        constexpr size_t smem = (size_t)G * P::SMEM_CF * sizeof(cf) + (size_t)P::M * sizeof(float2);
        static PerDevice cache[kMaxDevices];
        int& ctas_per_sm = per_device(cache).ctas_per_sm;
        if (ctas_per_sm == 0) {
            if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
                set_error("irfft_frames: cannot reserve shared memory");
                return ACIDS_ECUDA;
            }
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, THREADS, smem);
            ctas_per_sm = nb > 0 ? nb : 1;
        }
        const int64_t units = (p.rows + G - 1) / G;
        if (units == 0) return ACIDS_OK;
        int64_t grid = (int64_t)num_sms() * ctas_per_sm;
        if (grid > units) grid = units;
        kern<<<(unsigned)grid, THREADS, smem, st>>>(p);
        ACIDS_CHECK_LAUNCH("irfft_frames");
        return ACIDS_OK;
    }
};

static const size_t kMaxSmem = 227 * 1024;

#define ACIDS_INV_SWITCH(n_fft, EXPR)                                          \
    switch (n_fft) {                                                           \
        case 32: { using PL = Inv32; EXPR; } break;                            \
        case 64: { using PL = Inv64; EXPR; } break;                            \
        case 128: { using PL = Inv128; EXPR; } break;                          \
        case 256: { using PL = Inv256; EXPR; } break;                          \
        case 512: { using PL = Inv512; EXPR; } break;                          \
        case 1024: { using PL = Inv1024; EXPR; } break;                        \
        case 2048: { using PL = Inv2048; EXPR; } break;                        \
        case 4096: { using PL = Inv4096; EXPR; } break;                        \
        case 8192: { using PL = Inv8192; EXPR; } break;                        \
        case 16384: { using PL = Inv16384; EXPR; } break;                      \
        default:                                                               \
            set_error("n_fft=%d is not supported (power of two in [32, 16384])", n_fft); \
            return ACIDS_ENOTSUP;                                              \
    }

static int fused_fits(int n_fft, int hop, bool& fits) {
    const int ovc = (n_fft + hop - 1) / hop;
    size_t need = 0;
    ACIDS_INV_SWITCH(n_fft, need = InvLaunch<typename OlaPlan<PL>::type>::smem_ola(ovc, hop));
    fits = need <= kMaxSmem;
    return ACIDS_OK;
}

static int launch_ola_gather(const OlaParams& p, cudaStream_t st) {
    const int64_t total = (int64_t)(p.n - 1) * p.hop + p.N;
    if (p.B == 0 || p.n == 0) return ACIDS_OK;
    ACIDS_REQUIRE(p.B < 65536, ACIDS_EINVAL, "ola: more than 65535 clips per call");
    int64_t gx = (total + 255) / 256;
    if (gx > 4096) gx = 4096;
    ola_gather_kernel<<<dim3((unsigned)gx, (unsigned)p.B), 256, 0, st>>>(p);
    ACIDS_CHECK_LAUNCH("ola_gather");
    return ACIDS_OK;
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int64_t acids_istft_workspace_bytes(int64_t B, int64_t n_frames, int n_fft, int hop) {
    if (hop <= 0) return 0;
    bool fits = true;
    if (fused_fits(n_fft, hop, fits) != ACIDS_OK) return 0;
    return fits ? 0 : B * n_frames * (int64_t)n_fft * (int64_t)sizeof(float);
}

extern "C" ACIDS_API int acids_irfft_frames(const float* X, int64_t rows, int n_fft, const float* window, float* out, void* stream) {
    ACIDS_REQUIRE(X && window && out, ACIDS_EINVAL, "irfft_frames: NULL pointer");
    ACIDS_REQUIRE(rows >= 0, ACIDS_EINVAL, "irfft_frames: negative row count");
    InvFramesParams p{reinterpret_cast<const cf*>(X), rows, window, out};
    ACIDS_INV_SWITCH(n_fft, return InvLaunch<PL>::frames(p, static_cast<cudaStream_t>(stream)));
    return ACIDS_OK;
}

extern "C" ACIDS_API int acids_istft_ola(const float* X, int64_t B, int64_t n_frames, int n_fft, int hop, const float* window,
                               float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    ACIDS_REQUIRE(X && window && out, ACIDS_EINVAL, "istft_ola: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_frames < (1LL << 30) && hop > 0 && hop <= n_fft, ACIDS_EINVAL,
                  "istft_ola: bad sizes B=%lld frames=%lld hop=%d", (long long)B, (long long)n_frames, hop);
    bool fits = true;
    int rc = fused_fits(n_fft, hop, fits);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t out_len = (int64_t)hop * (n_frames - 1);
    ACIDS_REQUIRE(out_len + 2 * (int64_t)n_fft < ((int64_t)1 << 31), ACIDS_ENOTSUP, "istft_ola: clips of 2^31 samples or more are not supported");
    if (fits) {
        InvParams p{};
        p.X = reinterpret_cast<const cf*>(X); p.B = B; p.n_frames = (int)n_frames; p.hop = hop; p.window = window;
        p.out = out; p.out_len = out_len; p.trim = n_fft / 2; p.ovc = (n_fft + hop - 1) / hop;
        ACIDS_INV_SWITCH(n_fft, return InvLaunch<typename OlaPlan<PL>::type>::ola(p, st));
        return ACIDS_OK;
    }
    // two-step fallback: frames to the caller's workspace, then a gather
    const int64_t need = B * n_frames * (int64_t)n_fft * (int64_t)sizeof(float);
    ACIDS_REQUIRE(workspace && workspace_bytes >= need, ACIDS_EWORKSPACE,
                  "istft_ola: n_fft=%d hop=%d needs a %lld-byte workspace", n_fft, hop, (long long)need);
    rc = acids_irfft_frames(X, B * n_frames, n_fft, window, static_cast<float*>(workspace), stream);
    if (rc) return rc;
    OlaParams q{};
    q.frames = static_cast<const float*>(workspace); q.B = B; q.n = (int)n_frames; q.N = n_fft; q.hop = hop;
    q.window = window; q.trim = n_fft / 2; q.out = out; q.out_len = out_len; q.gain = 1.f;
    return launch_ola_gather(q, st);
}

extern "C" ACIDS_API int acids_ola_stream(const float* frames, int64_t B, int64_t n, int n_fft, int hop, int64_t keep,
                                const float* carry_in, float gain, float* out, float* carry_out, void* stream) {
    ACIDS_REQUIRE(frames && out, ACIDS_EINVAL, "ola_stream: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n >= 1 && hop > 0 && n_fft > 0 && keep >= 0, ACIDS_EINVAL, "ola_stream: bad sizes");
    const int64_t total = (n - 1) * hop + n_fft;
    ACIDS_REQUIRE(keep < total, ACIDS_EINVAL, "ola_stream: carry longer than the recomposed signal");
    ACIDS_REQUIRE(gain != 0.f, ACIDS_EINVAL, "ola_stream: zero gain");
    OlaParams q{};
    q.frames = frames; q.B = B; q.n = (int)n; q.N = n_fft; q.hop = hop; q.window = nullptr;
    q.carry_in = carry_in; q.keep = keep; q.gain = gain; q.trim = 0; q.out = out; q.out_len = total - keep;
    q.carry_out = carry_out;
    return launch_ola_gather(q, static_cast<cudaStream_t>(stream));
}
