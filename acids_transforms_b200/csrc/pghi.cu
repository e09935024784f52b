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
    float* hkey;           // workspace [B, T * F] x 8 bytes: the heap (64-bit keys, see Heap)
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
    // with replicated edges (np.pad(..., mode="edge"))
    auto fgrad = [&](int t, int k) {
        const float* row = logm + t * F;
        const float d = __fmul_rn(__fsub_rn(row[min(k + 1, F - 1)], row[max(k - 1, 0)]), 0.5f);
        return __fadd_rn(__fdiv_rn(d, p.fmul), __fmul_rn(p.kstep, (float)k));
    };
    auto tgrad = [&](int t, int k) {
        const float d = __fmul_rn(__fsub_rn(logm[min(t + 1, T - 1) * F + k], logm[max(t - 1, 0) * F + k]), 0.5f);
        return __fadd_rn(__fmul_rn(-p.fmul, d), 3.14159265358979323846f);
    };

    Heap h{reinterpret_cast<unsigned long long*>(p.hkey) + b * n, 0};
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
                const int i = h.pop();
                const int t = i / F, k = i - t * F;
                const float ph = phase[i];
                if (t + 1 < T && s[i + F] > abstol) {
                    phase[i + F] = __fadd_rn(ph, __fmul_rn(__fadd_rn(fgrad(t, k), fgrad(t + 1, k)), 0.5f));
                    h.push(Heap::make(s[i + F], i + F));
                    s[i + F] = abstol;
                }
                if (t > 0 && s[i - F] > abstol) {
                    phase[i - F] = __fsub_rn(ph, __fmul_rn(__fadd_rn(fgrad(t, k), fgrad(t - 1, k)), 0.5f));
                    h.push(Heap::make(s[i - F], i - F));
                    s[i - F] = abstol;
                }
                if (k + 1 < F && s[i + 1] > abstol) {
                    phase[i + 1] = __fadd_rn(ph, __fmul_rn(__fadd_rn(tgrad(t, k), tgrad(t, k + 1)), 0.5f));
                    h.push(Heap::make(s[i + 1], i + 1));
                    s[i + 1] = abstol;
                }
                if (k > 0 && s[i - 1] > abstol) {
                    phase[i - 1] = __fsub_rn(ph, __fmul_rn(__fadd_rn(tgrad(t, k), tgrad(t, k - 1)), 0.5f));
                    h.push(Heap::make(s[i - 1], i - 1));
                    s[i - 1] = abstol;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int64_t acids_pghi_workspace_bytes(int64_t B, int64_t n_frames, int n_bins) {
    if (B < 0 || n_frames < 0 || n_bins < 0) return 0;
    return B * n_frames * n_bins * 16;          // log|X|, |X| to visit, heap keys, heap payload: 4 x 4 bytes per bin
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
    p.logm = w; p.s = w + n; p.hkey = w + 2 * n;       // 2 n floats = n 64-bit heap entries (8-byte aligned: n floats each before)
    p.phase = phase;
    pghi_kernel<<<(unsigned)B, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
    ACIDS_CHECK_LAUNCH("pghi");
    return ACIDS_OK;
}
