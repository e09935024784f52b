// Microbenchmark for the "warp-shuffle exchange" question (BASELINE north_star; DESIGN.md section 5, experiment c):
// what does it cost a warp to exchange 16 complex values per lane
//   (A) through shared memory  : 16 STS.64 + __syncwarp + 16 LDS.64 (the transposition the FFT passes do), or
//   (B) with shuffles          : a radix-2 cross-lane stage, 8 complex = 16 SHFL.32 per lane plus the selects that pick
//                                what each lane sends / keeps (the cheapest shuffle scheme: one bit of the index), or
//   (C) a full 4-bit transposition with shuffles (what replaces ONE shared-memory exchange): 4 x (B).
// Reports cycles per exchange per SM at 16 resident warps per SM.   nvcc -arch=sm_100a -O3 exch.cu -o exch && ./exch
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
__global__ void __launch_bounds__(128, 4) k_smem(float2* out) {
    __shared__ float2 buf[4][16 * 34];
    float2 v[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) buf[w][i * 34 + lane] = v[i];              // conflict-free 8-byte stores
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = buf[w][((i + lane) & 15) * 34 + ((lane + 3 * i) & 31)];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i].x += 1.0f;
    }
    float2 r = make_float2(0, 0);
    for (int i = 0; i < 16; ++i) { r.x += v[i].x; r.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int STAGES>
__global__ void __launch_bounds__(128, 4) k_shfl(float2* out) {
    float2 v[16];
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = make_float2(threadIdx.x + i, i);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            const int m = 1 << s;
            const bool up = lane & m;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                // the lane keeps one half of its values and trades the other half with its partner
                float2 send = up ? v[i] : v[i + 8];
                send.x = __shfl_xor_sync(0xffffffffu, send.x, m);
                send.y = __shfl_xor_sync(0xffffffffu, send.y, m);
                if (up) v[i] = send; else v[i + 8] = send;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i].x += 1.0f;
    }
    float2 r = make_float2(0, 0);
    for (int i = 0; i < 16; ++i) { r.x += v[i].x; r.y += v[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class K>
static void run(const char* name, K kern, float2* out, int sms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<sms * 4, 128>>>(out);
    cudaEventRecord(a);
    kern<<<sms * 4, 128>>>(out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    // 16 warps per SM, ITERS exchanges each
    printf("%-34s %.3f ms  = %.1f cycles per warp-exchange per SM (at %d MHz, 16 warps / SM)\n", name, ms, ms * 1e-3 * khz * 1e3 / (16.0 * ITERS), khz / 1000);
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float2* out; cudaMalloc(&out, sizeof(float2) * sms * 4 * 128);
    run("shared memory (16 STS.64 + 16 LDS.64)", k_smem, out, sms);
    run("shuffle, 1 index bit (16 SHFL.32)", k_shfl<1>, out, sms);
    run("shuffle, 4 index bits (64 SHFL.32)", k_shfl<4>, out, sms);
    run("shuffle, 5 index bits (80 SHFL.32)", k_shfl<5>, out, sms);
    return 0;
}
