#include <cuda_runtime.h>
struct __align__(8) cf { float x, y; };
__device__ __forceinline__ unsigned long long pk(cf a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
__device__ __forceinline__ cf upk(unsigned long long r) { cf a; asm("mov.b64 {%0, %1}, %2;" : "=f"(a.x), "=f"(a.y) : "l"(r)); return a; }
__device__ __forceinline__ cf padd(cf a, cf b) { unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
__device__ __forceinline__ cf psub(cf a, cf b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a)), "l"(pk(b))); return upk(r); }
// packed complex multiply: (a.x*w.x - a.y*w.y, a.y*w.x + a.x*w.y)
__device__ __forceinline__ cf pcmul(cf a, cf w) {
    cf ww = {w.x, w.x}, sw = {a.y, a.x}, wy = {-w.y, w.y};
    unsigned long long t, r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(pk(a)), "l"(pk(ww)));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(sw)), "l"(pk(wy)), "l"(t));
    return upk(r);
}
__global__ void k(const cf* in, cf* out, cf w) {
    cf a[8];
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
    // radix-2 layers like a butterfly
    for (int s = 1; s < 8; s <<= 1)
        for (int i = 0; i < 8; ++i) if (!(i & s)) { cf u = a[i], v = a[i | s]; a[i] = padd(u, v); a[i | s] = psub(u, v); }
    for (int i = 1; i < 8; ++i) a[i] = pcmul(a[i], w);
    for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = a[i];
}
