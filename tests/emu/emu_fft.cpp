// CPU emulation of acids_transforms_b200/csrc/fft_core.cuh: the T threads of a frame group are
// stepped phase by phase (a phase boundary = the barrier the kernel places there), using the very
// same __host__ __device__ code the kernels run.  Checks forward (real -> half complex) and inverse
// against a naive double-precision DFT.  Build: g++ -O1 -std=c++17 emu_fft.cpp -o emu_fft
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../acids_transforms_b200/csrc/fft_core.cuh"
#include "../../acids_transforms_b200/csrc/plans.cuh"

using namespace acids;

template <class P, bool INV, int PASS>
static void mid_passes(std::vector<FrameFFT<P, INV>>& th, std::vector<std::vector<cf>>& v, std::vector<cf>& smem) {
    // store<PASS-1> | barrier | load<PASS> butterflies<PASS>
    if constexpr (PASS < P::NP) {
        for (int t = 0; t < P::T; ++t) th[t].template store<PASS - 1>(v[t].data(), smem.data());
        for (int t = 0; t < P::T; ++t) {
            th[t].template load<PASS>(v[t].data(), smem.data());
            th[t].template butterflies<PASS>(v[t].data());
        }
        mid_passes<P, INV, PASS + 1>(th, v, smem);
    }
}

template <class P>
static double check_fwd(const char* name) {
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V;
    std::vector<float> x(N);
    srand(1234 + N);
    for (int i = 0; i < N; ++i) x[i] = (float)rand() / RAND_MAX - 0.5f;
    // ---------------- forward ----------------
    std::vector<FrameFFT<P, false>> th(T);
    std::vector<std::vector<cf>> v(T, std::vector<cf>(V));
    std::vector<cf> smem(P::SMEM_CF, mk(NAN, NAN));
    for (int t = 0; t < T; ++t) th[t].init(t);
    for (int t = 0; t < T; ++t) {
        for (int b = 0; b < P::bpt(0); ++b)
            for (int r = 0; r < P::radix(0); ++r) {
                int n = th[t].template in_index<0>(b, r);
                v[t][b * P::radix(0) + r] = mk(0.5f * x[2 * n], 0.5f * x[2 * n + 1]);
            }
        th[t].template butterflies<0>(v[t].data());
    }
    mid_passes<P, false, 1>(th, v, smem);
    std::vector<cf> X(M + 1, mk(NAN, NAN));
    std::vector<int> hits(M + 1, 0);
    using PR = typename FrameFFT<P, false>::PR;
    for (int t = 0; t < T; ++t) {
        cf o1[V / 2], o2[V / 2], extra;
        th[t].untangle_fwd(v[t].data(), o1, o2, extra);
        for (int c = 0; c < PR::PC; ++c)
            for (int s = 0; s < PR::R; ++s) {
                int k = PR::k1(t, c, s);
                X[k] = o1[c * PR::R + s]; hits[k]++;
                X[M - k] = o2[c * PR::R + s]; hits[M - k]++;
            }
        if (t == 0) { X[M / 2] = extra; hits[M / 2]++; }
    }
    double maxerr = 0, peak = 0;
    for (int k = 0; k <= M; ++k) {
        double re = 0, im = 0;
        for (int n = 0; n < N; ++n) {
            double a = -2.0 * M_PI * (double)((long)k * n % N) / N;
            re += x[n] * cos(a); im += x[n] * sin(a);
        }
        if (hits[k] != 1) { printf("%s: bin %d written %d times\n", name, k, hits[k]); return 1e9; }
        maxerr = fmax(maxerr, hypot(X[k].x - re, X[k].y - im));
        peak = fmax(peak, hypot(re, im));
    }
    if (std::signbit(X[0].y) || std::signbit(X[M].y) || X[0].y != 0 || X[M].y != 0) { printf("%s: DC/Nyquist imag not +0\n", name); return 1e9; }
    double ef = maxerr / peak;
    printf("%-10s N=%5d T=%3d V=%2d NP=%d  fwd relerr %.2e\n", name, N, T, V, P::NP, ef);
    return ef;
}

template <class P>
static double check_inv(const char* name) {
    constexpr int N = P::N, M = P::M, T = P::T, V = P::V;
    srand(4321 + N);
    std::vector<std::vector<cf>> v(T, std::vector<cf>(V));
    std::vector<cf> smem(P::SMEM_CF, mk(NAN, NAN));
    std::vector<cf> Y(M + 1);
    for (int k = 0; k <= M; ++k) Y[k] = mk((float)rand() / RAND_MAX - 0.5f, (float)rand() / RAND_MAX - 0.5f);
    std::vector<FrameFFT<P, true>> ti(T);
    // the inverse uses the radices in the order given by the plan's mirror (first pass paired)
    for (int t = 0; t < T; ++t) ti[t].init(t);
    using PI = typename FrameFFT<P, true>::PR;
    for (int t = 0; t < T; ++t) {
        cf i1[V / 2], i2[V / 2];
        for (int c = 0; c < PI::PC; ++c)
            for (int s = 0; s < PI::R; ++s) {
                int k = PI::k1(t, c, s);
                i1[c * PI::R + s] = Y[k];
                i2[c * PI::R + s] = Y[M - k];
            }
        ti[t].pretangle_inv(i1, i2, Y[M / 2], v[t].data());
        ti[t].template butterflies<0>(v[t].data());
    }
    mid_passes<P, true, 1>(ti, v, smem);
    std::vector<float> y(N, NAN);
    for (int t = 0; t < T; ++t)
        for (int b = 0; b < P::bpt(P::NP - 1); ++b)
            for (int q = 0; q < P::radix(P::NP - 1); ++q) {
                int n = ti[t].template out_index<P::NP - 1>(b, q);
                cf z = v[t][b * P::radix(P::NP - 1) + q];
                y[2 * n] = z.x; y[2 * n + 1] = z.y;
            }
    double maxi = 0, peaki = 0;
    for (int n = 0; n < N; ++n) {
        double acc = Y[0].x + ((n & 1) ? -1.0 : 1.0) * Y[M].x;   // c2r ignores imag of DC / Nyquist
        for (int k = 1; k < M; ++k) {
            double a = 2.0 * M_PI * (double)((long)k * n % N) / N;
            acc += 2.0 * (Y[k].x * cos(a) - Y[k].y * sin(a));
        }
        maxi = fmax(maxi, fabs(y[n] - acc));       // kernel output is N * irfft (unnormalised)
        peaki = fmax(peaki, fabs(acc));
    }
    double ei = maxi / peaki;
    printf("%-10s N=%5d T=%3d V=%2d NP=%d  inv relerr %.2e\n", name, N, T, V, P::NP, ei);
    return ei;
}

// ---- shared-memory bank model: wavefronts a warp needs for each exchange access of a plan ----------------------
// 8-byte accesses are served per half warp (16 lanes, 16 bank pairs), 16-byte accesses per quarter warp
// (8 lanes, 8 bank quads); lanes that hit the same word are merged.  Prints actual / ideal wavefronts per frame.
static int wavefronts(const std::vector<long>& byte_addr, int width) {
    const int lanes_per = width == 16 ? 8 : 16, units = 128 / width;
    int total = 0;
    for (size_t l0 = 0; l0 < byte_addr.size(); l0 += lanes_per) {
        std::vector<std::vector<long>> bank(units);
        for (size_t l = l0; l < l0 + lanes_per && l < byte_addr.size(); ++l) {
            long u = byte_addr[l] / width;
            auto& bk = bank[u % units];
            bool seen = false;
            for (long w : bk) seen |= (w == u);
            if (!seen) bk.push_back(u);
        }
        size_t deg = 1;
        for (auto& bk : bank) deg = std::max(deg, bk.size());
        total += (int)deg;
    }
    return total;
}

template <class P, bool INV, int PASS>
static void pass_conflicts(std::vector<FrameFFT<P, INV>>& th, long& actual, long& ideal) {
    if constexpr (PASS < P::NP) {
        constexpr int R = P::radix(PASS), B = P::bpt(PASS), NS = P::ns(PASS);
        const int lanes = 32, groups = P::T >= 32 ? 1 : 32 / P::T, warps = P::T >= 32 ? P::T / 32 : 1;
        auto slot_addr = [&](int lane_global, int idx) {   // byte address of slot idx for that lane's frame group
            int g = P::T >= 32 ? 0 : lane_global / P::T;
            return ((long)g * P::SMEM_CF + swz<P::PADLOG>(idx)) * 8L;
        };
        for (int w = 0; w < warps; ++w) {
            if (PASS < P::NP - 1) {   // store<PASS>
                const bool vec = (NS == 1 && R % 2 == 0);
                for (int b = 0; b < B; ++b)
                    for (int q = 0; q < R; q += vec ? 2 : 1) {
                        std::vector<long> a;
                        for (int l = 0; l < lanes; ++l) {
                            int t = P::T >= 32 ? w * 32 + l : l % P::T;
                            a.push_back(slot_addr(l, th[t].template out_index<PASS>(b, q)));
                        }
                        actual += wavefronts(a, vec ? 16 : 8);
                        ideal += vec ? 4 : 2;
                    }
            }
            if (PASS > 0) {           // load<PASS>
                for (int b = 0; b < B; ++b)
                    for (int r = 0; r < R; ++r) {
                        std::vector<long> a;
                        for (int l = 0; l < lanes; ++l) {
                            int t = P::T >= 32 ? w * 32 + l : l % P::T;
                            a.push_back(slot_addr(l, th[t].template in_index<PASS>(b, r)));
                        }
                        actual += wavefronts(a, 8);
                        ideal += 2;
                    }
            }
        }
        (void)groups;
        pass_conflicts<P, INV, PASS + 1>(th, actual, ideal);
    }
}

template <class P, bool INV>
static void report_conflicts(const char* name) {
    std::vector<FrameFFT<P, INV>> th(P::T);
    for (int t = 0; t < P::T; ++t) th[t].init(t);
    long actual = 0, ideal = 0;
    pass_conflicts<P, INV, 0>(th, actual, ideal);
    printf("%-10s %s exchange wavefronts per warp-frame-set: %ld (ideal %ld, x%.2f)\n", name, INV ? "inv" : "fwd", actual, ideal,
           (double)actual / ideal);
}

#define CHECKF(PL) worst = fmax(worst, check_fwd<PL>(#PL))
#define CHECKI(PL) worst = fmax(worst, check_inv<PL>(#PL))

int main() {
    double worst = 0;
    ACIDS_FOR_EACH_FWD_PLAN(CHECKF);
    ACIDS_FOR_EACH_INV_PLAN(CHECKI);
#define CONFF(PL) report_conflicts<PL, false>(#PL)
#define CONFI(PL) report_conflicts<PL, true>(#PL)
    ACIDS_FOR_EACH_FWD_PLAN(CONFF);
    ACIDS_FOR_EACH_INV_PLAN(CONFI);
    if (!(worst < 5e-6)) { printf("FAIL worst %.3e\n", worst); return 1; }
    printf("OK worst %.3e\n", worst);
    return 0;
}
