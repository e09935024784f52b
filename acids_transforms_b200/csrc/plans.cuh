// plans.cuh — the (N, threads-per-frame, radices) table.  V = N/2/T complex registers per thread.
// Forward plans need an even butterfly count per thread in their LAST pass (paired untangle),
// inverse plans in their FIRST pass; with these radix orders one table serves both when the
// inverse walks it mirrored (see InvPlan).
#pragma once
#include "fft_core.cuh"

namespace acids {

// forward: last pass paired
using Fwd32    = Plan<32,    1, 2, 8>;
using Fwd64    = Plan<64,    2, 4, 8>;
using Fwd128   = Plan<128,   4, 8, 8>;
using Fwd256   = Plan<256,   8, 4, 4, 8>;
using Fwd512   = Plan<512,  16, 4, 8, 8>;
#if defined(ACIDS_FWD1024_R0)
using Fwd1024  = Plan<1024, 32, ACIDS_FWD1024_R0, ACIDS_FWD1024_R1, ACIDS_FWD1024_R2>;   // experiment: other radix orders
#else
using Fwd1024  = Plan<1024, 32, 8, 8, 8>;
#endif
// n_fft = 1024 with a real-valued (magnitude / mel) epilogue: ONE exchange — 16 threads x 32 values per frame, radix 32 then a
// paired radix 16; two frames share a warp.  64 instead of 144 exchange wavefronts per frame; affordable since the mirrored
// twiddle set and the compact untangle twiddles (fft_core.cuh) brought it from 238 to 128 registers.  ACIDS_FWD1024R_T32
// puts the real-output kernels back on the two-exchange plan (tuning experiment).
#if defined(ACIDS_FWD1024R_T32)
using Fwd1024R = Fwd1024;
#else
using Fwd1024R = Plan<1024, 16, 32, 16>;
#endif
using Fwd2048  = Plan<2048, 64, 16, 8, 8>;
using Fwd4096  = Plan<4096, 128, 16, 16, 8>;
using Fwd8192  = Plan<8192, 256, 8, 8, 8, 8>;
using Fwd16384 = Plan<16384, 512, 16, 8, 8, 8>;
// inverse: first pass paired
using Inv32    = Plan<32,    1, 8, 2>;
using Inv64    = Plan<64,    2, 8, 4>;
using Inv128   = Plan<128,   4, 8, 8>;
using Inv256   = Plan<256,   8, 8, 4, 4>;
using Inv512   = Plan<512,  16, 8, 8, 4>;
using Inv1024  = Plan<1024, 32, 8, 8, 8>;
// one-exchange inverse plan (see Fwd1024R): paired radix 16, then radix 32; the pass-1 twiddles ride on the outputs of pass 0
using Inv1024T16 = Plan<1024, 16, 16, 32>;
using Inv2048  = Plan<2048, 64, 8, 8, 16>;
using Inv4096  = Plan<4096, 128, 8, 16, 16>;
using Inv8192  = Plan<8192, 256, 8, 8, 8, 8>;
using Inv16384 = Plan<16384, 512, 8, 8, 8, 16>;

}  // namespace acids

#define ACIDS_FOR_EACH_FWD_PLAN(X) \
    X(Fwd32); X(Fwd64); X(Fwd128); X(Fwd256); X(Fwd512); X(Fwd1024); X(Fwd1024R); X(Fwd2048); X(Fwd4096); X(Fwd8192); X(Fwd16384)
#define ACIDS_FOR_EACH_INV_PLAN(X) \
    X(Inv32); X(Inv64); X(Inv128); X(Inv256); X(Inv512); X(Inv1024); X(Inv1024T16); X(Inv2048); X(Inv4096); X(Inv8192); X(Inv16384)
