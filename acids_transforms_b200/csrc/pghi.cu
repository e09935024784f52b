// pghi.cu — phase-gradient heap integration on the GPU (reference dgt.py:156-236).
//
// Průša, Balazs, Søndergaard, "A Noniterative Method for Reconstruction of Phase from STFT Magnitude" (TASLP 2017) as the
// reference parameterises it: the phase of a [T, F] magnitude is flood-filled from the loudest bin outwards, always
// continuing from the loudest bin visited so far (a priority queue), integrating the frequency derivative of log|X| along
// frames and the (negated) time derivative along bins; bins below tol * max keep phase 0; disconnected regions are seeded
// again from their own loudest bin.  The fill of ONE clip is sequential and data dependent — the reference runs it as a
// Python loop over `heapq`, ~1 s per 4-second clip — but clips are independent: one CTA per clip, all threads for the
// element-wise set-up and the arg-max of every seed, thread 0 for the heap walk.  A batch of 1024 clips then costs about
// what one clip costs on the host.
//
// Exactness: the visiting order only depends on the magnitudes and on the order of the heap keys, which are compared like
// the reference's tuples (-|X|, t, k) (linear index breaks ties: same order); every arithmetic step is a single rounded
// float32 operation in the reference's order, so the result differs from the host restatement only through logf (1 ulp).
#include "common.cuh"

namespace acids {

struct PghiParams {
    const float* mag;      // [B, T, F]
    int64_t B;
    int T, F;
    float fmul;            // gamma / (hop * n_fft)
    float kstep;           // 2 pi hop / n_fft
    double tol;            // compared as float32(peak * tol) with the product in double, like the reference's Python floats
    float abstol;
    float* logm;           // workspace [B, T, F]: log(max(mag, abstol))
    float* s;              // workspace [B, T, F]: magnitudes still to visit (abstol = visited / too quiet)
    float* hkey;           // workspace, 64-byte aligned: B heaps of pghi_heap_stride(T F) 64-bit keys (see Heap8)
    float* phase;          // out [B, T, F]
};

// Binary min-heap of 64-bit keys: (0xFFFFFFFF - bits(|X|)) << 32 | linear bin index.  |X| > 0, so its bit pattern orders like
// its value: ascending keys are descending magnitudes, ties by ascending (t, k) — the order of the reference's tuples
// (-|X|, t, k) — with ONE load and ONE compare per entry (the walk is a single thread's dependent chain: instructions count).
struct Heap {
    unsigned long long* a;
    int n;
    __device__ __forceinline__ static unsigned long long make(float mag, int i) {
        return ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(mag)) << 32) | (unsigned)i;
    }
    __device__ void push(unsigned long long k) {
        int c = n++;
        while (c > 0) {
            const int pnt = (c - 1) >> 1;
            const unsigned long long pk = a[pnt];
            if (!(k < pk)) break;
            a[c] = pk;
            c = pnt;
        }
        a[c] = k;
    }
    __device__ int pop() {
        const int top = (int)(unsigned)a[0];
        --n;
        if (n > 0) {
            const unsigned long long k = a[n];
            int c = 0;
            for (;;) {
                int l = 2 * c + 1;
                if (l >= n) break;
                unsigned long long lk = a[l];
                if (l + 1 < n) {
                    const unsigned long long rk = a[l + 1];
                    if (rk < lk) { ++l; lk = rk; }
                }
                if (!(lk < k)) break;
                a[c] = lk;
                c = l;
            }
            a[c] = k;
        }
        return top;
    }
};

// The whole-spectrogram walk keeps hundreds of thousands of entries in GLOBAL memory, where every level of a sift is a trip to
// L2: an 8-ary heap has a third of the binary heap's levels (6 instead of 18 at 354 k entries), and the eight children of a
// node are one aligned 64-byte line fetched by four independent 16-byte loads.  Node i lives at a[7 + i], so the children
// 8 i + 1 ... 8 i + 8 start at a[8 (i + 1)].  The pop order only depends on the keys, not on the heap's shape.
struct Heap8 {
    // (Keeping the top four levels — 585 nodes, 4.7 KB — in shared memory was measured: 0.65 s against 0.47 s per clip; the
    // address-space select on every access costs more than the L1-resident top of the heap saves.)
    unsigned long long* a;      // global, 64-byte aligned; node i at a[7 + i]
    int n;
    __device__ __forceinline__ unsigned long long& at(int i) { return a[7 + i]; }
    __device__ __forceinline__ void push(unsigned long long k) {
        int c = n++;
        while (c > 0) {
            const int pnt = (c - 1) >> 3;
            const unsigned long long pk = at(pnt);
            if (!(k < pk)) break;
            at(c) = pk;
            c = pnt;
        }
        at(c) = k;
    }
    __device__ __forceinline__ int peek() const { return (int)(unsigned)a[7]; }
    __device__ __forceinline__ void drop() {          // removes the root (peek() first)
        --n;
        if (n > 0) {
            const unsigned long long k = at(n);
            int c = 0;
            for (;;) {
                const int l = 8 * c + 1;
                if (l >= n) break;
                unsigned long long v[8];
                const unsigned long long* ch = a + 7 + l;
                if (l + 8 <= n) {
                    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(ch);
                    const ulonglong2 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                    v[0] = q0.x; v[1] = q0.y; v[2] = q1.x; v[3] = q1.y; v[4] = q2.x; v[5] = q2.y; v[6] = q3.x; v[7] = q3.y;
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = l + j < n ? ch[j] : ~0ull;
                }
                // smallest of the eight (keys are distinct), as a tree
                const bool b01 = v[1] < v[0], b23 = v[3] < v[2], b45 = v[5] < v[4], b67 = v[7] < v[6];
                const unsigned long long m01 = b01 ? v[1] : v[0], m23 = b23 ? v[3] : v[2], m45 = b45 ? v[5] : v[4], m67 = b67 ? v[7] : v[6];
                const int i01 = b01 ? 1 : 0, i23 = b23 ? 3 : 2, i45 = b45 ? 5 : 4, i67 = b67 ? 7 : 6;
                const bool b03 = m23 < m01, b47 = m67 < m45;
                const unsigned long long m03 = b03 ? m23 : m01, m47 = b47 ? m67 : m45;
                const int i03 = b03 ? i23 : i01, i47 = b47 ? i67 : i45;
                const bool b07 = m47 < m03;
                const unsigned long long mk = b07 ? m47 : m03;
                if (!(mk < k)) break;
                at(c) = mk;
                c = l + (b07 ? i47 : i03);
            }
            at(c) = k;
        }
    }
    __device__ __forceinline__ int pop() {
        const int top = peek();
        drop();
        return top;
    }
};

static __host__ __device__ __forceinline__ int64_t pghi_heap_stride(int64_t n) { return ((n + 15) & ~(int64_t)7) + 8; }   // entries per clip, a multiple of 8

__global__ void __launch_bounds__(256) pghi_kernel(const PghiParams p) {
    const int64_t b = blockIdx.x;
    const int T = p.T, F = p.F, n = T * F;
    const float* __restrict__ mag = p.mag + b * n;
    float* __restrict__ logm = p.logm + b * n;
    float* __restrict__ s = p.s + b * n;
    float* __restrict__ phase = p.phase + b * n;
    __shared__ float red_v[256];
    __shared__ int red_i[256];
    const float abstol = p.abstol;
    // ---- set-up (dgt.py:167-176, :226-236): clamp, log, peak, threshold ----
    float mx = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float m = fmaxf(__ldg(mag + i), abstol);
        logm[i] = logf(m);
        s[i] = m;
        phase[i] = 0.f;
        mx = fmaxf(mx, m);
    }
    red_v[threadIdx.x] = mx;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red_v[threadIdx.x] = fmaxf(red_v[threadIdx.x], red_v[threadIdx.x + o]);
        __syncthreads();
    }
    const float thr = (float)((double)red_v[0] * p.tol);
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        if (s[i] < thr) s[i] = abstol;          // too quiet to trust: never visited, phase stays 0
    __syncthreads();

    // gradients on demand, each a chain of single rounded float32 operations in the reference's order:
    //   fgrad = ((y[k+1] - y[k-1]) / 2) / fmul + (2 pi hop / n_fft) k      tgrad = -fmul ((y[t+1] - y[t-1]) / 2) + pi
    // with replicated edges (np.pad(..., mode="edge")); see the walk

    Heap8 h{reinterpret_cast<unsigned long long*>(p.hkey) + b * pghi_heap_stride(n), 0};
    for (;;) {
        // ---- seed: the loudest unvisited bin, FIRST index on ties like np.argmax ----
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const float v = s[i];
            if (v > bv) { bv = v; bi = i; }      // a thread walks ascending indices: keeps its first maximum
        }
        red_v[threadIdx.x] = bv;
        red_i[threadIdx.x] = bi;
        __syncthreads();
        for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v2 = red_v[threadIdx.x + o];
                const int i2 = red_i[threadIdx.x + o];
                if (v2 > red_v[threadIdx.x] || (v2 == red_v[threadIdx.x] && i2 < red_i[threadIdx.x])) {
                    red_v[threadIdx.x] = v2;
                    red_i[threadIdx.x] = i2;
                }
            }
            __syncthreads();
        }
        const float top = red_v[0];
        const int seed = red_i[0];
        __syncthreads();
        if (!(top > abstol)) break;
        if (threadIdx.x == 0) {
            // ---- the flood fill of this region (dgt.py:186-224) ----
            h.n = 0;
            h.push(Heap::make(top, seed));
            s[seed] = abstol;
            while (h.n > 0) {
                const int i = h.peek();
                const int t = i / F, k = i - t * F;
                // every load of this step first — the four neighbours' magnitudes, the bin's phase and the 3 x 3 patch of
                // log-magnitudes the six gradients read (edges replicated like np.pad(mode="edge")) — so that they are in
                // flight while the heap sifts; the arithmetic below is the reference's, one rounded float32 operation per step
                const bool hn = t + 1 < T, hp = t > 0, hr = k + 1 < F, hl = k > 0;
                const int tn = min(t + 1, T - 1) * F, tp = max(t - 1, 0) * F, t0 = t * F, kr = min(k + 1, F - 1), kl = max(k - 1, 0);
                const float sn = hn ? s[i + F] : 0.f, sp = hp ? s[i - F] : 0.f, sr = hr ? s[i + 1] : 0.f, sl = hl ? s[i - 1] : 0.f;
                const float ph = phase[i];
                const float y0l = logm[t0 + kl], y0r = logm[t0 + kr];
                const float ynl = logm[tn + kl], ynk = logm[tn + k], ynr = logm[tn + kr];
                const float ypl = logm[tp + kl], ypk = logm[tp + k], ypr = logm[tp + kr];
                h.drop();
                auto fgv = [&](float right, float left) {        // fgrad of a row at bin k
                    return __fadd_rn(__fdiv_rn(__fmul_rn(__fsub_rn(right, left), 0.5f), p.fmul), __fmul_rn(p.kstep, (float)k));
                };
                auto tgv = [&](float next, float prev) {         // tgrad of a column between frames t - 1 and t + 1
                    return __fadd_rn(__fmul_rn(-p.fmul, __fmul_rn(__fsub_rn(next, prev), 0.5f)), 3.14159265358979323846f);
                };
                const float f0 = fgv(y0r, y0l), t0g = tgv(ynk, ypk);
                if (hn && sn > abstol) {
                    phase[i + F] = __fadd_rn(ph, __fmul_rn(__fadd_rn(f0, fgv(ynr, ynl)), 0.5f));
                    h.push(Heap::make(sn, i + F));
                    s[i + F] = abstol;
                }
                if (hp && sp > abstol) {
                    phase[i - F] = __fsub_rn(ph, __fmul_rn(__fadd_rn(f0, fgv(ypr, ypl)), 0.5f));
                    h.push(Heap::make(sp, i - F));
                    s[i - F] = abstol;
                }
                if (hr && sr > abstol) {
                    // tgrad(t, k + 1): the column to the right, kstep-free
                    phase[i + 1] = __fadd_rn(ph, __fmul_rn(__fadd_rn(t0g, tgv(ynr, ypr)), 0.5f));
                    h.push(Heap::make(sr, i + 1));
                    s[i + 1] = abstol;
                }
                if (hl && sl > abstol) {
                    phase[i - 1] = __fsub_rn(ph, __fmul_rn(__fadd_rn(t0g, tgv(ynl, ypl)), 0.5f));
                    h.push(Heap::make(sl, i - 1));
                    s[i - 1] = abstol;
                }
            }
        }
        __syncthreads();
    }
}


// ---- frame-by-frame variant (RealtimeDGT.pghi, dgt.py:338-452) --------------------------------------------------------
// Every new frame f is filled from two sources: the previous frame's audible bins (a step along TIME with the 3-point
// stencil gradient) and the frame's own loudest bin (steps along BINS); the heap never holds more than two frames, so it
// lives in shared memory together with the frame's working rows — a pop is a chain of ~10 shared-memory loads instead of
// ~10 trips to L2.  One CTA per stream; the block's frames are sequential (frame f needs the phase of f - 1).
// Reference quirks kept (each shapes the output): the gradient rows are indexed two frames late (dgt.py:393-395), bin 0 is
// never reached from bin 1 (`> 0`, dgt.py:434), the frame's first seed is pushed without being marked (dgt.py:409), quiet
// bins take `noise` (the reference draws randn).  The stencil row before the first history frame replicates that frame
// (the reference reads torch.empty memory there, see transforms/pghi.py).
struct RtPghiParams {
    const float* mag;         // [B, n, F] new frames
    const float* hist_mag;    // [B, 2, F]
    const float* hist_phase;  // [B, F]
    const float* noise;       // [B, n, F] or NULL (quiet bins keep 0)
    int64_t B;
    int n, F;
    float fmul, kstep, tol, eps;
    float* logm;              // workspace [B, n + 2, F]
    unsigned long long* gheap;  // workspace heap [B, 2 F + 8] when it does not fit shared memory, else NULL
    float* phase;             // out [B, n, F]
    int p2;                   // > 0: rank-queue mode, the two frames' keys are sorted into p2 (power of two >= 2 F) slots
};

// Priority queue over a universe that is known in advance.  Every key a frame can ever queue — (|X|, frame, bin) of the
// previous frame's bins and of its own — exists before the walk starts, so the keys are sorted once by the whole CTA and
// the queue is a bitmap over their ranks with one summary level: push = set a bit, pop = two find-first-set steps
// instead of a ~10-level sift through the heap.  Queueing a key twice (the frame's first seed can be, dgt.py:409) sets the
// same bit: the heap would pop the duplicate right after the original, where it finds its neighbours visited — a no-op.
struct RankQueue {
    unsigned* bm;
    unsigned* sum;
    int sw;
    __device__ __forceinline__ void push(int r) {
        bm[r >> 5] |= 1u << (r & 31);
        sum[r >> 10] |= 1u << ((r >> 5) & 31);
    }
    __device__ __forceinline__ int pop() {          // the smallest queued rank, -1 when the queue is empty
        for (int i = 0; i < sw; ++i) {
            const unsigned sm = sum[i];
            if (sm) {
                const int w = (i << 5) + __ffs(sm) - 1;
                unsigned word = bm[w];
                const int bit = __ffs(word) - 1;
                word &= word - 1;
                bm[w] = word;
                if (!word) sum[i] = sm & (sm - 1);
                return (w << 5) + bit;
            }
        }
        return -1;
    }
};

__global__ void __launch_bounds__(32) rt_pghi_kernel(const RtPghiParams p) {
    extern __shared__ __align__(16) unsigned char rt_smem[];
    const int64_t b = blockIdx.x;
    const int F = p.F, n = p.n, nt = n + 2, NT = blockDim.x;
    float* s_row = reinterpret_cast<float*>(rt_smem);      // magnitudes of frame f still to visit
    float* ph_a = s_row + F;                                // phase rows of frames f - 1 / f (swapped every frame)
    float* ph_b = ph_a + F;
    float* e_row = ph_b + F;                                // 0.5 (tg[f - 1] + tg[f]): the step along time into frame f
    float* f_row = e_row + F;                               // fg[f]
    unsigned long long* hs = reinterpret_cast<unsigned long long*>(f_row + F + (F & 1));    // heap, or the sorted keys
    const int P2 = p.p2, W = max(P2 >> 5, 1), SW = (W + 31) >> 5;
    unsigned short* rank_of = reinterpret_cast<unsigned short*>(hs + P2);                    // [2 F]: (frame, bin) -> rank
    unsigned* bm = reinterpret_cast<unsigned*>(rank_of + 2 * F + (2 * F & 1));
    RankQueue q{bm, bm + W, SW};
    __shared__ float red_v[128];
    __shared__ int red_i[128];
    const float* __restrict__ mag = p.mag + b * (int64_t)n * F;
    const float* __restrict__ hmag = p.hist_mag + b * 2 * (int64_t)F;
    float* __restrict__ logm = p.logm + b * (int64_t)nt * F;
    float* __restrict__ out = p.phase + b * (int64_t)n * F;
    const float* __restrict__ noise = p.noise ? p.noise + b * (int64_t)n * F : nullptr;
    auto clamped = [&](int t, int k) { return fmaxf(t < 2 ? __ldg(hmag + t * F + k) : __ldg(mag + (int64_t)(t - 2) * F + k), p.eps); };

    // ---- log of every row, the block's peak, the threshold (dgt.py:343, :387) ----
    float mx = 0.f;
    for (int i = threadIdx.x; i < nt * F; i += NT) {
        const float m = clamped(i / F, i % F);
        logm[i] = logf(m);
        mx = fmaxf(mx, m);
    }
    red_v[threadIdx.x] = mx;
    __syncthreads();
    for (int o = NT >> 1; o > 0; o >>= 1) {
        if (threadIdx.x < o) red_v[threadIdx.x] = fmaxf(red_v[threadIdx.x], red_v[threadIdx.x + o]);
        __syncthreads();
    }
    const float abstol = fmaxf(__fmul_rn(p.tol, red_v[0]), p.eps);
    __syncthreads();
    for (int k = threadIdx.x; k < F; k += NT) ph_a[k] = __ldg(p.hist_phase + b * F + k);      // phase of frame 1

    // gradient rows as the reference indexes them: two zero rows, then the rows of the stencil (float32, one rounding per step)
    auto tg = [&](int t, int k) {
        if (t < 2) return 0.f;
        const int j = t - 2;
        const float a = __fmul_rn(3.f, logm[(j + 1) * F + k]), c = __fmul_rn(4.f, logm[j * F + k]);
        const float d = __fmul_rn(__fadd_rn(__fsub_rn(a, c), logm[max(j - 1, 0) * F + k]), 0.5f);
        return __fadd_rn(__fmul_rn(-p.fmul, d), 3.14159265358979323846f);
    };
    auto fg = [&](int t, int k) {
        if (t < 2) return 0.f;
        const float* row = logm + (t - 2) * F;
        const float d = __fmul_rn(__fsub_rn(row[min(k + 1, F - 1)], row[max(k - 1, 0)]), 0.5f);
        return __fadd_rn(__fdiv_rn(d, p.fmul), __fmul_rn(p.kstep, (float)k));
    };
    auto argmax_row = [&](float& top, int& k0) {        // first index on ties, like np.argmax / nonzero(...)[0]
        float bv = -1.f;
        int bi = 0x7fffffff;
        for (int k = threadIdx.x; k < F; k += NT)
            if (s_row[k] > bv) { bv = s_row[k]; bi = k; }
        red_v[threadIdx.x] = bv;
        red_i[threadIdx.x] = bi;
        __syncthreads();
        for (int o = NT >> 1; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v2 = red_v[threadIdx.x + o];
                const int i2 = red_i[threadIdx.x + o];
                if (v2 > red_v[threadIdx.x] || (v2 == red_v[threadIdx.x] && i2 < red_i[threadIdx.x])) {
                    red_v[threadIdx.x] = v2;
                    red_i[threadIdx.x] = i2;
                }
            }
            __syncthreads();
        }
        top = red_v[0];
        k0 = red_i[0];
        __syncthreads();
    };
    // heap keys: magnitude (complemented bits) | frame (0: f - 1, 1: f) | bin — the order of the reference's (-|X|, (t, k))
    auto key = [](float m, int trel, int k) {
        return ((unsigned long long)(0xFFFFFFFFu - __float_as_uint(m)) << 32) | (unsigned)(trel << 16) | (unsigned)k;
    };
    Heap h{p.gheap ? p.gheap + b * (2 * (int64_t)F + 8) : hs, 0};
    float* ph_prev = ph_a;
    float* ph_cur = ph_b;
    for (int f = 2; f < nt; ++f) {
        __syncthreads();
        for (int k = threadIdx.x; k < F; k += NT) {
            const float m = clamped(f, k);
            s_row[k] = m;
            ph_cur[k] = (m > abstol || !noise) ? 0.f : __ldg(noise + (int64_t)(f - 2) * F + k);
            e_row[k] = __fmul_rn(0.5f, __fadd_rn(tg(f - 1, k), tg(f, k)));
            f_row[k] = fg(f, k);
        }
        __syncthreads();
        float top;
        int k0;
        argmax_row(top, k0);
        if (top > abstol && P2 > 0) {
            // ---- rank-queue mode: sort the 2 F keys of frames f - 1 and f, queue the audible bins of f - 1 and the seed ----
            for (int i = threadIdx.x; i < P2; i += NT)
                hs[i] = i < F ? key(clamped(f - 1, i), 0, i) : (i < 2 * F ? key(s_row[i - F], 1, i - F) : ~0ull);
            __syncthreads();
            for (int k2 = 2; k2 <= P2; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < P2; i += NT) {
                        const int ixj = i ^ j;
                        if (ixj > i) {
                            const unsigned long long a = hs[i], c = hs[ixj];
                            if ((a > c) == ((i & k2) == 0)) { hs[i] = c; hs[ixj] = a; }
                        }
                    }
                    __syncthreads();
                }
            for (int w = threadIdx.x; w < W; w += NT) bm[w] = 0;
            __syncthreads();
            for (int r = threadIdx.x; r < P2; r += NT) {
                const unsigned long long kk = hs[r];
                const unsigned lo = (unsigned)kk;
                if (kk == ~0ull) continue;
                rank_of[(lo >> 16 ? F : 0) + (lo & 0xFFFFu)] = (unsigned short)r;
                if (!(lo >> 16) && __uint_as_float(0xFFFFFFFFu - (unsigned)(kk >> 32)) > abstol) atomicOr(bm + (r >> 5), 1u << (r & 31));
            }
            __syncthreads();
            for (int i = threadIdx.x; i < SW; i += NT) {
                unsigned sm = 0;
                for (int w = 0; w < 32 && (i << 5) + w < W; ++w) sm |= (bm[(i << 5) + w] ? 1u : 0u) << w;
                q.sum[i] = sm;
            }
            __syncthreads();
            if (threadIdx.x == 0) q.push(rank_of[F + k0]);               // the frame's loudest bin: queued, NOT marked
            for (;;) {
                if (threadIdx.x == 0) {
                    int r;
                    while ((r = q.pop()) >= 0) {
                        const unsigned lo = (unsigned)hs[r];
                        const int k = (int)(lo & 0xFFFFu);
                        const int kr = min(k + 1, F - 1), kl = max(k - 1, 0);
                        // every load of the step before its first store: independent shared-memory reads in flight together
                        // instead of a chain through the branches (a lane's walk is latency, not bandwidth)
                        const float sk = s_row[k], sr = s_row[kr], sl = s_row[kl];
                        const float pc = ph_cur[k], pp = ph_prev[k], ek = e_row[k];
                        const float fk = f_row[k], fr = f_row[kr], fl = f_row[kl];
                        const int rk = rank_of[F + k], rr = rank_of[F + kr], rl = rank_of[F + kl];
                        if (!(lo >> 16)) {                       // a bin of frame f - 1: the step along time
                            if (sk > abstol) {
                                ph_cur[k] = __fadd_rn(pp, ek);
                                q.push(rk);
                                s_row[k] = abstol;
                            }
                        } else {                                 // a bin of frame f: its two neighbours
                            if (k + 1 < F && sr > abstol) {
                                ph_cur[k + 1] = __fadd_rn(pc, __fmul_rn(0.5f, __fadd_rn(fk, fr)));
                                q.push(rr);
                                s_row[k + 1] = abstol;
                            }
                            if (k - 1 > 0 && sl > abstol) {
                                ph_cur[k - 1] = __fsub_rn(pc, __fmul_rn(0.5f, __fadd_rn(fk, fl)));
                                q.push(rl);
                                s_row[k - 1] = abstol;
                            }
                        }
                    }
                }
                __syncthreads();
                argmax_row(top, k0);                             // what the flood has not reached (dgt.py:447-451)
                if (!(top > abstol)) break;
                if (threadIdx.x == 0) {
                    q.push(rank_of[F + k0]);
                    s_row[k0] = abstol;
                }
                __syncthreads();
            }
        } else if (top > abstol) {
            if (threadIdx.x == 0) {
                h.n = 0;
                h.push(key(top, 1, k0));
                for (int k = 0; k < F; ++k) {
                    const float m = clamped(f - 1, k);
                    if (m > abstol) h.push(key(m, 0, k));
                }
            }
            for (;;) {
                if (threadIdx.x == 0) {
                    while (h.n > 0) {
                        const unsigned lo = (unsigned)h.a[0];
                        h.pop();
                        const int k = (int)(lo & 0xFFFFu);
                        if (!(lo >> 16)) {                       // a bin of frame f - 1: the step along time
                            const float m = s_row[k];
                            if (m > abstol) {
                                ph_cur[k] = __fadd_rn(ph_prev[k], e_row[k]);
                                h.push(key(m, 1, k));
                                s_row[k] = abstol;
                            }
                        } else {                                 // a bin of frame f: its two neighbours
                            const float ph = ph_cur[k], fk = f_row[k];
                            if (k + 1 < F && s_row[k + 1] > abstol) {
                                ph_cur[k + 1] = __fadd_rn(ph, __fmul_rn(0.5f, __fadd_rn(fk, f_row[k + 1])));
                                h.push(key(s_row[k + 1], 1, k + 1));
                                s_row[k + 1] = abstol;
                            }
                            if (k - 1 > 0 && s_row[k - 1] > abstol) {
                                ph_cur[k - 1] = __fsub_rn(ph, __fmul_rn(0.5f, __fadd_rn(fk, f_row[k - 1])));
                                h.push(key(s_row[k - 1], 1, k - 1));
                                s_row[k - 1] = abstol;
                            }
                        }
                    }
                }
                __syncthreads();
                argmax_row(top, k0);                             // what the flood has not reached (dgt.py:447-451)
                if (threadIdx.x == 0) {
                    h.push(key(top, 1, k0));
                    s_row[k0] = abstol;
                }
                if (!(top > abstol)) break;
                __syncthreads();
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < F; k += NT) out[(int64_t)(f - 2) * F + k] = ph_cur[k];
        float* t = ph_prev;
        ph_prev = ph_cur;
        ph_cur = t;
    }
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int64_t acids_pghi_workspace_bytes(int64_t B, int64_t n_frames, int n_bins) {
    if (B < 0 || n_frames < 0 || n_bins < 0) return 0;
    // log|X| and |X| to visit (4 bytes per bin each), 64 bytes of alignment, one 8-byte heap entry per bin + the heap's slack
    return B * n_frames * n_bins * 8 + 64 + B * pghi_heap_stride(n_frames * (int64_t)n_bins) * 8;
}

extern "C" ACIDS_API int acids_pghi(const float* mag, int64_t B, int64_t n_frames, int n_bins, float gamma, int n_fft, int hop,
                          double tol, float eps, void* workspace, int64_t workspace_bytes, float* phase, void* stream) {
    ACIDS_REQUIRE(mag && phase, ACIDS_EINVAL, "pghi: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_bins >= 1 && n_frames * (int64_t)n_bins < (1LL << 30), ACIDS_EINVAL, "pghi: bad sizes");
    ACIDS_REQUIRE(n_fft > 0 && hop > 0 && gamma > 0.f && eps > 0.f, ACIDS_EINVAL, "pghi: bad parameters");
    if (B == 0) return ACIDS_OK;
    ACIDS_REQUIRE(workspace && workspace_bytes >= acids_pghi_workspace_bytes(B, n_frames, n_bins), ACIDS_EINVAL,
                  "pghi: workspace of %lld bytes required", (long long)acids_pghi_workspace_bytes(B, n_frames, n_bins));
    const int64_t n = B * n_frames * n_bins;
    PghiParams p{};
    p.mag = mag; p.B = B; p.T = (int)n_frames; p.F = n_bins;
    // float32 like the reference: fmul = gamma / (hop * n_fft), the bin step 2 pi hop / n_fft (dgt.py:226-236)
    p.fmul = (float)((double)gamma / ((double)hop * (double)n_fft));
    p.kstep = (float)(2.0 * 3.14159265358979323846 * (double)hop / (double)n_fft);
    p.tol = tol; p.abstol = eps;
    float* w = static_cast<float*>(workspace);
    p.logm = w; p.s = w + n;
    p.hkey = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(w + 2 * n) + 63) & ~(uintptr_t)63);      // the heaps' 64-byte lines
    p.phase = phase;
    pghi_kernel<<<(unsigned)B, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("pghi");
    return ACIDS_OK;
}

static size_t rt_pghi_smem(int n_bins, bool heap_in_smem) {
    return (size_t)(5 * n_bins + (n_bins & 1)) * 4 + (heap_in_smem ? (size_t)(2 * n_bins + 8) * 8 : 0);
}
static int rt_pghi_p2(int n_bins) {          // slots of the sorted key array: a power of two >= max(2 n_bins, 32)
    int p2 = 32;
    while (p2 < 2 * n_bins) p2 <<= 1;
    return p2;
}
static size_t rt_pghi_smem_ranked(int n_bins) {
    const int p2 = rt_pghi_p2(n_bins), w = p2 / 32, sw = (w + 31) / 32;
    return (size_t)(5 * n_bins + (n_bins & 1)) * 4 + (size_t)p2 * 8 + (size_t)(2 * n_bins + (2 * n_bins & 1)) * 2 + (size_t)(w + sw) * 4;
}
static const size_t kRtPghiSmemMax = 200 * 1024;

extern "C" ACIDS_API int64_t acids_rt_pghi_workspace_bytes(int64_t B, int64_t n_frames, int n_bins) {
    if (B < 0 || n_frames < 0 || n_bins < 0) return 0;
    int64_t bytes = B * (n_frames + 2) * n_bins * 4;                        // log|X| of the history and the new frames
    bytes = (bytes + 7) & ~(int64_t)7;
    if (rt_pghi_smem(n_bins, true) > kRtPghiSmemMax) bytes += B * (2 * (int64_t)n_bins + 8) * 8;     // the heap, when it is too large for shared memory
    return bytes;
}

extern "C" ACIDS_API int acids_rt_pghi(const float* mag, const float* hist_mag, const float* hist_phase, const float* noise, int64_t B,
                             int64_t n_frames, int n_bins, float gamma, int n_fft, int hop, float tol, float eps, void* workspace,
                             int64_t workspace_bytes, float* phase, void* stream) {
    ACIDS_REQUIRE(mag && hist_mag && hist_phase && phase, ACIDS_EINVAL, "rt_pghi: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 1 && n_bins >= 2 && n_bins < 65536 && (n_frames + 2) * (int64_t)n_bins < (1LL << 30), ACIDS_EINVAL,
                  "rt_pghi: bad sizes");
    ACIDS_REQUIRE(n_fft > 0 && hop > 0 && gamma > 0.f && eps > 0.f, ACIDS_EINVAL, "rt_pghi: bad parameters");
    if (B == 0) return ACIDS_OK;
    ACIDS_REQUIRE(workspace && workspace_bytes >= acids_rt_pghi_workspace_bytes(B, n_frames, n_bins), ACIDS_EINVAL,
                  "rt_pghi: workspace of %lld bytes required", (long long)acids_rt_pghi_workspace_bytes(B, n_frames, n_bins));
    RtPghiParams p{};
    p.mag = mag; p.hist_mag = hist_mag; p.hist_phase = hist_phase; p.noise = noise;
    p.B = B; p.n = (int)n_frames; p.F = n_bins;
    p.fmul = (float)((double)gamma / ((double)hop * (double)n_fft));
    p.kstep = (float)(2.0 * 3.14159265358979323846 * (double)hop / (double)n_fft);
    p.tol = tol; p.eps = eps;
    p.logm = static_cast<float*>(workspace);
    const bool ranked = rt_pghi_smem_ranked(n_bins) <= kRtPghiSmemMax;        // n_fft <= 4096; larger transforms keep the heap
    const bool in_smem = ranked || rt_pghi_smem(n_bins, true) <= kRtPghiSmemMax;
    if (ranked) p.p2 = rt_pghi_p2(n_bins);
    if (!in_smem) {
        int64_t off = B * (n_frames + 2) * n_bins * 4;
        off = (off + 7) & ~(int64_t)7;
        p.gheap = reinterpret_cast<unsigned long long*>(static_cast<unsigned char*>(workspace) + off);
    }
    p.phase = phase;
    const size_t smem = ranked ? rt_pghi_smem_ranked(n_bins) : rt_pghi_smem(n_bins, in_smem);
    ACIDS_REQUIRE(smem <= kRtPghiSmemMax, ACIDS_EINVAL, "rt_pghi: %d bins do not fit shared memory", n_bins);
    ACIDS_REQUIRE(cudaFuncSetAttribute(rt_pghi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem > 48 * 1024 ? smem : 48 * 1024)) == cudaSuccess,
                  ACIDS_ECUDA, "rt_pghi: cannot reserve %zu bytes of shared memory", smem);
    // ONE warp per stream: the parallel phases are loops over at most 2 F elements (a 2048-key bitonic sort is ~40 us for 32
    // lanes) and the walk is a single lane either way, so more warps only add barriers between phases
    rt_pghi_kernel<<<(unsigned)B, 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("rt_pghi");
    return ACIDS_OK;
}
