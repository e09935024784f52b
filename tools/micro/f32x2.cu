// Microbenchmark: scalar FADD/FFMA vs packed add.f32x2 / fma.rn.f32x2 issue throughput on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__global__ void k_scalar(float* out, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], s, 1.0f);
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_packed(float* out, float s) {
    unsigned long long a[8], sv, one;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
    asm("mov.b64 %0, {%1, %1};" : "=l"(one) : "f"(1.0f));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(sv), "l"(one));
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_scalar_add(float* out, float s) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = a[i] + s;
    }
    float r = 0; for (int i = 0; i < 16; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_packed_add(float* out, float s) {
    unsigned long long a[8], sv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(sv));
    }
    float r = 0;
    for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class K> float run(K k, float* out, const char* name, double flops_per_thread) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 8, 256>>>(out, 1.0001f); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) k<<<148 * 8, 256>>>(out, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double lane_ops = 148.0 * 8 * 256 * flops_per_thread;
    printf("%-14s %.3f ms  %.2f T lane-results/s  (%.1f results/clk/SM @1.965GHz)\n", name, ms, lane_ops / ms / 1e9, lane_ops / (ms * 1e-3) / 148 / 1.965e9);
    return ms;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run(k_scalar, out, "FFMA scalar", 16.0 * ITERS);
    run(k_packed, out, "FFMA2 packed", 16.0 * ITERS);
    run(k_scalar_add, out, "FADD scalar", 16.0 * ITERS);
    run(k_packed_add, out, "FADD2 packed", 16.0 * ITERS);
    cudaError_t e = cudaGetLastError(); printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
