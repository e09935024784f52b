// mfcc_tc.cu — the MFCC tail (dB + top_db floor + ortho DCT-II) with the DCT on the 5th-generation tensor cores.
//
// The DCT (n_mels -> n_mfcc, 128 -> 40 at BASELINE cfg 3) is the one genuinely dense contraction of the path
// (DESIGN.md §6).  Per tile of 128 frames:  D[128 frames x NP coefficients] = dB[128 x K] . DCT[K x NP], K = n_mels.
//   * operands in shared memory, K-major, no swizzle ("interleave": 8-row x 16-byte core matrices, LBO = 128 B between
//     core matrices along K, SBO = K/4 * 128 B between 8-row groups), described by UMMA shared-memory descriptors;
//   * `tcgen05.mma.cta_group::1.kind::tf32` issued by one elected thread, accumulator in tensor memory (TMEM);
//   * fp32 fidelity by operand splitting (3xTF32): x = hi + lo with hi = x truncated to TF32; D = A_hi B_hi + A_lo B_hi
//     + A_hi B_lo accumulated in the same TMEM tile (the dropped lo.lo term is 2^-22 relative): plain TF32 inputs
//     (2^-11) would break the 1e-4 parity budget on dB-scaled data;
//   * `tcgen05.commit` -> mbarrier; the four warps read their TMEM lane quadrant back with `tcgen05.ld 32x32b` and
//     store [n_mfcc, T] frequency-major like torchaudio.
// The dB conversion, not the GEMM, dominates the kernel: the point of this file is the numerically faithful tensor-core
// formulation north_star asks for, measured next to the FP32 register-tiled kernel (pointwise.cu).
#include <float.h>
#include "common.cuh"

namespace acids {

namespace tc {

constexpr int TM = 128;      // frames per tile = MMA M
constexpr int NP = 48;       // padded coefficient count = MMA N (multiple of 8 for cta_group::1)
constexpr int KMAX = 128;    // n_mels <= 128, multiple of 8
constexpr int TMEM_COLS = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), K-major, SWIZZLE_NONE
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);                 // start address, bits [0, 14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;       // leading byte offset, bits [16, 30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;       // stride byte offset, bits [32, 46)
    d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
    return d;                                               // base offset 0, lbo mode 0, layout type 0 (no swizzle)
}

// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, dense, M = 128, N = NP
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

// byte offset of element (row r, column k) in a K-major interleaved operand with K columns
__device__ __forceinline__ uint32_t op_offset(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K >> 2) * 128 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xffffe000u);     // sign + exponent + 10 mantissa bits
    lo = x - hi;                                                // exact; the tensor core reads its TF32 prefix
}

struct Params {
    const float* mel;        // [B, K, T]
    int64_t B;
    int K;
    int64_t T;
    const float* dct;        // [K, n_mfcc]
    int n_mfcc;
    float top_db;
    const float* gmax;
    int64_t clips_per_group;
    float* out;              // [B, n_mfcc, T]
};

constexpr int THREADS = 512;      // 16 warps: operand fill and dB conversion are the bulk of the work

__global__ void __launch_bounds__(THREADS, 1) mfcc_dct_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int K = p.K;
    const uint32_t a_bytes = (uint32_t)TM * K * 4, b_bytes = (uint32_t)NP * K * 4;
    unsigned char* a_hi = smem;
    unsigned char* a_lo = a_hi + a_bytes;
    unsigned char* b_hi = a_lo + a_bytes;
    unsigned char* b_lo = b_hi + b_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + b_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B operand: row n = coefficient, column k = mel band: DCT^T, zero rows for n >= n_mfcc
    for (int i = tid; i < NP * K; i += THREADS) {
        const int n = i / K, k = i - n * K;
        float hi, lo;
        split_tf32(n < p.n_mfcc ? __ldg(p.dct + (size_t)k * p.n_mfcc + n) : 0.f, hi, lo);
        const uint32_t off = op_offset(n, k, K);
        *reinterpret_cast<float*>(b_hi + off) = hi;
        *reinterpret_cast<float*>(b_lo + off) = lo;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc();
    const uint32_t sbo = (uint32_t)(K >> 2) * 128;
    uint32_t phase = 0;
    const int cores_k = K >> 2;                       // core matrices along K
    const int n_cores = (TM / 8) * cores_k;           // 8 frames x 4 mel bands each

    const int64_t tiles_per_clip = (p.T + TM - 1) / TM;
    for (int64_t tile = blockIdx.x; tile < p.B * tiles_per_clip; tile += gridDim.x) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * TM;
        float floor_db = -FLT_MAX;
        if (p.top_db >= 0.f) floor_db = 10.0f * log10f(fmaxf(__ldg(p.gmax + b / p.clips_per_group), 1e-10f)) - p.top_db;
        // A operand: row = frame, column = mel band.  Per step a warp takes 32 frames x 4 bands: four coalesced 128-byte
        // loads along the frame axis; a lane then holds one 16-byte row of a core matrix (its frame, 4 bands), so the
        // hi and lo parts go out as two conflict-free 128-bit shared-memory stores.
        // dB through lg2.approx (2^-22 relative: 1e-5 dB, four orders below the parity budget on values of +-100 dB).
        const float* src = p.mel + b * (int64_t)K * p.T;
        const int steps = (TM / 32) * cores_k;
        constexpr int NW = THREADS / 32, U = 4;      // U steps in flight per warp: 16 independent loads before the first use
        for (int c0 = warp; c0 < steps; c0 += NW * U) {
            float v[U][4];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * NW;
                const int fb = c / cores_k, kc = c - fb * cores_k;      // frame block (32 frames), band block (4 bands)
                const int64_t t = t0 + fb * 32 + lane;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[u][j] = (c < steps && t < p.T) ? __ldg(src + (int64_t)(kc * 4 + j) * p.T + t) : 1.0f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = c0 + u * NW;
                if (c >= steps) break;
                const int fb = c / cores_k, kc = c - fb * cores_k;
                const int r = fb * 32 + lane;
                const bool live = t0 + r < p.T;
                float4 hi, lo;
                float* hp = &hi.x;
                float* lp = &lo.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = live ? fmaxf(3.01029995663981195f * fast_lg2(fmaxf(v[u][j], 1e-10f)), floor_db) : 0.f;   // 10 log10(x), functional.py:390-404
                    split_tf32(d, hp[j], lp[j]);
                }
                const uint32_t off = (uint32_t)((r >> 3) * cores_k + kc) * 128 + (uint32_t)(r & 7) * 16;      // == op_offset(r, 4 kc, K)
                *reinterpret_cast<float4*>(a_hi + off) = hi;
                *reinterpret_cast<float4*>(a_lo + off) = lo;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo), sb_hi = smem_u32(b_hi), sb_lo = smem_u32(b_lo);
                uint32_t acc = 0;
                for (int ks = 0; ks < K / 8; ++ks) {              // one MMA consumes K = 8 (two core matrices of 4 TF32)
                    const uint32_t ko = (uint32_t)ks * 256;
                    mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, acc);
                    acc = 1;
                    mma_tf32(tmem_d, make_desc(sa_lo + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, 1);
                    mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_lo + ko, 128, sbo), idesc, 1);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
            }
            __syncwarp();
        }
        // wait for the accumulator
        {
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, p;\n\t"
                    "}\n"
                    : "=r"(done)
                    : "r"(smem_u32(mbar)), "r"(phase)
                    : "memory");
            }
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: a warp reads the TMEM lane quadrant (warp % 4) = 32 frames; the warp group (warp / 4) picks the
        // 16-column chunk of coefficients.  Stores are coalesced along frames ([n_mfcc, T] frequency-major).
        const int quad = warp & 3, chunk = warp >> 2;
        if (chunk * 16 < NP) {
            uint32_t r[16];
            const uint32_t taddr = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chunk * 16);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int64_t t = t0 + quad * 32 + lane;
            if (t < p.T) {
                float* o = p.out + (b * p.n_mfcc + chunk * 16) * p.T + t;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (chunk * 16 + n < p.n_mfcc) o[(int64_t)n * p.T] = __uint_as_float(r[n]);
            }
        }
        // the next tile overwrites the operands and the accumulator
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
}

}  // namespace tc
}  // namespace acids

using namespace acids;

// Same contract as acids_mfcc_dct (include/acids_b200.h); group_max must already hold the per-group maxima when
// top_db >= 0 (acids_mfcc_dct computes them; this entry point is called by it when the shape fits the tensor-core tile).
int acids_mfcc_dct_tc_launch(const float* mel, int64_t B, int n_mels, int64_t n_frames, const float* dct, int n_mfcc, float top_db,
                             int64_t clips_per_group, const float* group_max, float* out, cudaStream_t st) {
    tc::Params p{mel, B, n_mels, n_frames, dct, n_mfcc, top_db, group_max, clips_per_group > 0 ? clips_per_group : 1, out};
    const size_t smem = (size_t)2 * tc::TM * n_mels * 4 + (size_t)2 * tc::NP * n_mels * 4 + 64;
    ACIDS_REQUIRE(cudaFuncSetAttribute(tc::mfcc_dct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                  ACIDS_ECUDA, "mfcc_dct_tc: cannot reserve %zu B of shared memory", smem);
    int64_t grid = B * ((n_frames + tc::TM - 1) / tc::TM);
    if (grid > num_sms()) grid = num_sms();
    tc::mfcc_dct_tc_kernel<<<(unsigned)grid, tc::THREADS, smem, st>>>(p);
    ACIDS_CHECK_LAUNCH("mfcc_dct_tc");
    return ACIDS_OK;
}
