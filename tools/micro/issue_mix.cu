// Microbenchmark: do packed FP32 instructions (FADD2) leave issue slots free for other pipes on sm_100a?
// Each kernel runs 16 warps / SM (4 per scheduler) of straight-line independent work:
//   fp2      : 16 independent FADD2 chains
//   fp1      : 32 independent FADD chains (same FP work, scalar)
//   alu      : 16 independent IADD3-class chains (LOP3/IADD on the integer pipe)
//   fp2+alu  : both interleaved.  If time(fp2+alu) ~= max(time(fp2), time(alu)) the issue port is free while the FMA
//              pipe drains a packed op; if ~= sum, a packed op blocks dispatch for both of its cycles.
//   fp2+lds  : FADD2 interleaved with shared-memory loads.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define U 16
__global__ void k_fp2(float* out, float s) {
    unsigned long long a[U], sv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
#pragma unroll
    for (int i = 0; i < U; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(sv));
    }
    float r = 0;
    for (int i = 0; i < U; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_fp1(float* out, float s) {
    float a[2 * U];
#pragma unroll
    for (int i = 0; i < 2 * U; ++i) a[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 2 * U; ++i) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(s));
    }
    float r = 0;
    for (int i = 0; i < 2 * U; ++i) r += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_alu(float* out, int s, int s2) {
    int b[U];
#pragma unroll
    for (int i = 0; i < U; ++i) b[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) { asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(s)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(s2)); }
    }
    int r = 0;
    for (int i = 0; i < U; ++i) r += b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (float)r;
}
__global__ void k_fp2_alu(float* out, float s, int si, int s2) {
    unsigned long long a[U], sv;
    int b[U];
    asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
#pragma unroll
    for (int i = 0; i < U; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); b[i] = threadIdx.x + i; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(sv));
            { asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(si)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(s2)); }
        }
    }
    float r = 0;
    for (int i = 0; i < U; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y + b[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_fp1_alu(float* out, float s, int si, int s2) {
    float a[2 * U];
    int b[U];
#pragma unroll
    for (int i = 0; i < 2 * U; ++i) a[i] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < U; ++i) b[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[2 * i]) : "f"(s));
            asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[2 * i + 1]) : "f"(s));
            { asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(si)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(b[i]) : "r"(s2)); }
        }
    }
    float r = 0;
    for (int i = 0; i < 2 * U; ++i) r += a[i];
    for (int i = 0; i < U; ++i) r += b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_fp2_lds(float* out, float s) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    unsigned long long a[U], sv;
    float acc[4] = {0, 0, 0, 0};
    asm("mov.b64 %0, {%1, %1};" : "=l"(sv) : "f"(s));
#pragma unroll
    for (int i = 0; i < U; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(a[i]) : "f"(x)); }
    const volatile float* p = sm + threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < U; ++i) {
            asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a[i]) : "l"(sv));
            if ((i & 1) == 0) acc[(i >> 1) & 3] += p[(i * 32) & 511];
        }
    }
    float r = acc[0] + acc[1] + acc[2] + acc[3];
    for (int i = 0; i < U; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a[i])); r += x + y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <class F> void run(const char* name, F launch, double warp_instr_per_thread_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < 5; ++i) launch();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    // 16 warps / SM = 4 per scheduler; cycles per scheduler per inner iteration (one pass over the U chains) per warp
    double cyc = ms * 1e-3 * 1.965e9 / ITERS / 4.0;
    printf("%-10s %.3f ms   %.1f clk per (warp, iteration) per scheduler   [%g warp-instr per iteration]\n", name, ms, cyc, warp_instr_per_thread_iter);
}
int main() {
    float* out; cudaMalloc(&out, 148 * 4 * 128 * 4);
    dim3 g(148 * 4), b(128);
    run("fp2", [&] { k_fp2<<<g, b>>>(out, 1.0001f); }, U);
    run("fp1", [&] { k_fp1<<<g, b>>>(out, 1.0001f); }, 2 * U);
    run("alu", [&] { k_alu<<<g, b>>>(out, 3, 0x5555); }, 2 * U);
    run("fp2+alu", [&] { k_fp2_alu<<<g, b>>>(out, 1.0001f, 3, 0x5555); }, 3 * U);
    run("fp1+alu", [&] { k_fp1_alu<<<g, b>>>(out, 1.0001f, 3, 0x5555); }, 4 * U);
    run("fp2+lds", [&] { k_fp2_lds<<<g, b>>>(out, 1.0001f); }, U + U / 2 + U / 2);
    printf("status: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
