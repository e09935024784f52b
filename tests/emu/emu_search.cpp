// Search for lane transpositions of the paired pass that lower the exchange wavefront count of a plan
// (see PairLane in fft_core.cuh).  Build: g++ -O1 -std=c++17 -DACIDS_EMU_SEARCH emu_search.cpp -o emu_search
#define main emu_main
#include "emu_fft.cpp"
#undef main

template <class P, bool INV>
static long count_conf() {
    std::vector<FrameFFT<P, INV>> th(P::T);
    for (int t = 0; t < P::T; ++t) th[t].init(t);
    long actual = 0, ideal = 0;
    pass_conflicts<P, INV, 0>(th, actual, ideal);
    return actual;
}

template <class P, bool INV>
static void search(const char* name) {
    g_pair_swap_a = g_pair_swap_b = 0;
    long base = count_conf<P, INV>(), best = base;
    int ba = 0, bb = 0;
    for (int a = 1; a < 32; ++a)
        for (int b = a + 1; b < 32; ++b) {
            if (a >= P::T || b >= P::T) continue;
            g_pair_swap_a = a; g_pair_swap_b = b;
            long c = count_conf<P, INV>();
            if (c < best) { best = c; ba = a; bb = b; }
        }
    g_pair_swap_a = g_pair_swap_b = 0;
    printf("%-10s %s: identity %ld, best single swap (%d,%d) -> %ld\n", name, INV ? "inv" : "fwd", base, ba, bb, best);
}

int main() {
#define SF(PL) search<PL, false>(#PL)
#define SI(PL) search<PL, true>(#PL)
    ACIDS_FOR_EACH_FWD_PLAN(SF);
    ACIDS_FOR_EACH_INV_PLAN(SI);
    return 0;
}
