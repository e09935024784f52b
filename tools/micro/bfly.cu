// Microbenchmark: what do the register-resident parts of the 1024-point forward transform cost on their own?
//   WHAT = 1: butterflies of the three passes + untangle (FP pipe only, no shared memory)
//   WHAT = 2: the two exchanges (STS / __syncwarp / LDS), no arithmetic
//   WHAT = 3: both (the complete transform without global loads / stores)
// Built twice: packed FP32x2 (default) and -DACIDS_NO_PACKED (scalar FADD / FFMA).  Prints clk per frame per SM.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../acids_transforms_b200/csrc/plans.cuh"
using namespace acids;
using P = Fwd1024;
using FFT = FrameFFT<P, false>;

template <int WHAT, int MINB>
__global__ void __launch_bounds__(128, MINB) k(float* out, int iters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int g = threadIdx.x / 32, tid = threadIdx.x % 32;
    cf* s = reinterpret_cast<cf*>(smem_raw) + g * P::SMEM_CF;
    FFT fft;
    fft.init(tid);
    cf v[P::V];
#pragma unroll
    for (int i = 0; i < P::V; ++i) v[i] = mk(0.001f * (threadIdx.x + i), 0.002f * i);
    for (int it = 0; it < iters; ++it) {
        if (WHAT & 1) fft.template butterflies<0>(v);
        if (WHAT & 2) {
            __syncwarp();
            fft.template store<0>(v, s);
            __syncwarp();
            fft.template load<1>(v, s);
        }
        if (WHAT & 1) fft.template butterflies<1>(v);
        if (WHAT & 2) {
            __syncwarp();
            fft.template store<1>(v, s);
            __syncwarp();
            fft.template load<2>(v, s);
        }
        if (WHAT & 1) {
            fft.template butterflies<2>(v);
            cf o1[P::V / 2], o2[P::V / 2], ex;
            fft.untangle_fwd(v, o1, o2, ex);
#pragma unroll
            for (int i = 0; i < P::V / 2; ++i) {
                v[i] = mk(o1[i].x * 1e-3f, o1[i].y * 1e-3f);
                v[P::V / 2 + i] = mk(o2[i].x * 1e-3f + ex.x * 1e-9f, o2[i].y * 1e-3f);
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < P::V; ++i) r += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int WHAT, int MINB>
void run(float* out, const char* name) {
    const int iters = 2000, ctas = 148 * MINB;
    size_t smem = 4 * P::SMEM_CF * sizeof(cf);
    k<WHAT, MINB><<<ctas, 128, smem>>>(out, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<WHAT, MINB><<<ctas, 128, smem>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double frames_per_sm = (double)MINB * 4 * iters;
    printf("%-28s warps/SM %2d  %.3f ms  %.1f clk/frame/SM  (%s)\n", name, MINB * 4, ms, ms * 1e-3 * 1.965e9 / frames_per_sm,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 8 * 128 * 4);
#ifdef ACIDS_NO_PACKED
    printf("scalar FP32 build\n");
#else
    printf("packed FP32x2 build\n");
#endif
    run<1, 4>(out, "butterflies + untangle");
    run<1, 3>(out, "butterflies + untangle");
    run<1, 2>(out, "butterflies + untangle");
    run<1, 1>(out, "butterflies + untangle");
    run<2, 4>(out, "exchanges only");
    run<2, 2>(out, "exchanges only");
    run<3, 4>(out, "whole transform (no gmem)");
    run<3, 3>(out, "whole transform (no gmem)");
    run<3, 2>(out, "whole transform (no gmem)");
    run<3, 1>(out, "whole transform (no gmem)");
    return 0;
}
