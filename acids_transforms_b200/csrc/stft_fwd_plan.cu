// stft_fwd_plan.cu — instantiates the fused forward kernels of ONE FFT plan (-DACIDS_FWD_PLAN_N=<n_fft>).
// The build compiles this file once per plan so that the ten plans build in parallel.
#include "stft_fwd_kernel.cuh"

#ifndef ACIDS_FWD_PLAN_N
#error "compile with -DACIDS_FWD_PLAN_N=<n_fft>"
#endif
#define ACIDS_CAT2(a, b) a##b
#define ACIDS_CAT(a, b) ACIDS_CAT2(a, b)

namespace acids {

int ACIDS_CAT(launch_fwd_plan_, ACIDS_FWD_PLAN_N)(int variant, const FwdParams& p, cudaStream_t st) {
    return launch_any<ACIDS_CAT(Fwd, ACIDS_FWD_PLAN_N)>(variant, p, st);
}

}  // namespace acids
