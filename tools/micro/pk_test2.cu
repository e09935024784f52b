#include <cuda_runtime.h>
struct __align__(8) cf { float x, y; };
__device__ __forceinline__ cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }
__device__ __forceinline__ float2 f2(cf a) { return make_float2(a.x, a.y); }
__device__ __forceinline__ cf c2(float2 a) { return mk(a.x, a.y); }
__device__ __forceinline__ cf padd(cf a, cf b) { return c2(__fadd2_rn(f2(a), f2(b))); }
__device__ __forceinline__ cf psub(cf a, cf b) { return c2(__fadd2_rn(f2(a), make_float2(-b.x, -b.y))); }
__device__ __forceinline__ cf pcmul(cf a, cf w) {
    float2 t = __fmul2_rn(f2(a), make_float2(w.x, w.x));
    return c2(__ffma2_rn(make_float2(a.y, a.x), make_float2(-w.y, w.y), t));
}
__global__ void k(const cf* in, cf* out, const cf* tw) {
    cf a[8];
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
    cf w = tw[threadIdx.x];
    for (int s = 1; s < 8; s <<= 1)
        for (int i = 0; i < 8; ++i) if (!(i & s)) { cf u = a[i], v = a[i | s]; a[i] = padd(u, v); a[i | s] = psub(u, v); }
    for (int i = 1; i < 8; ++i) a[i] = pcmul(a[i], w);
    // conj-add and (-i) rotation patterns used by the untangle / radix-4
    cf e = mk(a[0].x + a[1].x, a[0].y - a[1].y);
    cf r = padd(a[2], mk(a[3].y, -a[3].x));
    a[0] = padd(e, r);
    for (int i = 0; i < 8; ++i) out[threadIdx.x * 8 + i] = a[i];
}
