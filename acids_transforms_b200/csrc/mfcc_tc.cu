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
#include "tc_common.cuh"

namespace acids {

namespace tc {

constexpr int TM = 128;      // frames per tile = MMA M
constexpr int NP = 48;       // padded coefficient count = MMA N (multiple of 8 for cta_group::1)
constexpr int KMAX = 128;    // n_mels <= 128, multiple of 8
constexpr int TMEM_COLS = 64;

// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, dense, M = 128, N = NP
__host__ __device__ constexpr uint32_t make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NP >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}

// byte offset of element (row r, column k) in a K-major interleaved operand with K columns
__device__ __forceinline__ uint32_t op_offset(int r, int k, int K) {
    return (uint32_t)((r >> 3) * (K >> 2) * 128 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
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

constexpr int THREADS = 256;      // 8 warps; two CTAs per SM (112 KB each), so one CTA's drain overlaps the other's operand fill
constexpr int KC = 32;            // mel bands per pipeline stage: the A operand is filled and multiplied in K chunks of 32

// Round 2: the kernel was one CTA per SM walking fill(128 KB) -> MMA -> drain strictly in turn (tensor pipe 8.6 % active, the
// SM idle during every barrier).  Now the A operand lives in a ring of two 32-band chunk buffers (2 x 32 KB for hi + lo): while
// the tensor core multiplies chunk c (12 tcgen05.mma, committed to the chunk buffer's mbarrier) the warps convert and store
// chunk c + 1; the accumulator in TMEM sums over the chunks.  112 KB of shared memory per CTA instead of 180 KB -> two CTAs
// per SM, whose fill and drain phases interleave.
__global__ void __launch_bounds__(THREADS, 2) mfcc_dct_tc_kernel(const Params p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int K = p.K;
    const uint32_t a_chunk = (uint32_t)TM * KC * 4, b_bytes = (uint32_t)NP * K * 4;
    unsigned char* a_hi = smem;                         // [2 stages][TM x KC]
    unsigned char* a_lo = a_hi + 2 * a_chunk;
    unsigned char* b_hi = a_lo + 2 * a_chunk;
    unsigned char* b_lo = b_hi + b_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_lo + b_bytes);      // [0], [1]: chunk buffer free again; [2]: accumulator complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar + i)), "r"(1) : "memory");
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc();
    const uint32_t sbo_a = (uint32_t)(KC >> 2) * 128, sbo_b = (uint32_t)(K >> 2) * 128;
    const int n_chunks = (K + KC - 1) / KC;           // K is a multiple of 8: the last chunk may hold 8, 16 or 24 bands
    uint32_t ph_buf[2] = {0, 0}, ph_done = 0;         // mbarrier phases
    uint32_t used[2] = {0, 0};                        // has the stage's mbarrier an outstanding commit?
    auto wait = [&](uint64_t* bar, uint32_t parity) {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t"
                ".reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t"
                "}\n"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity)
                : "memory");
        }
    };

    const int64_t tiles_per_clip = (p.T + TM - 1) / TM;
    const int64_t n_tiles = p.B * tiles_per_clip;
    int stage = 0;
    constexpr int cores_k = KC >> 2;                      // 8 core matrices along K per chunk
    constexpr int steps = (TM / 32) * cores_k;            // 32
    constexpr int NW = THREADS / 32, U = 4;               // 4 steps (16 independent loads) in flight per warp
    static_assert(steps == NW * U, "one pass of the warps covers a chunk");
    // raw mel values of chunk kc of a tile, in fill order: per step a warp takes 32 frames x 4 bands (four coalesced 128-byte
    // loads along the frame axis).  Issued one chunk AHEAD of their use (ncu on the first pipelined version: 4.6 long-scoreboard
    // stall cycles per issue — the loads of a chunk were only issued after the previous chunk's barrier).
    auto load_chunk = [&](int64_t tile, int kc, float (&v)[U][4]) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * TM;
        const float* src = p.mel + b * (int64_t)K * p.T;
        const int kw = min(KC, K - kc * KC);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = warp + u * NW;
            const int fb = c / cores_k, kb = c - fb * cores_k;      // frame block (32 frames), band block (4 bands)
            const int64_t t = t0 + fb * 32 + lane;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[u][j] = (t < p.T && kb * 4 < kw) ? __ldg(src + (int64_t)(kc * KC + kb * 4 + j) * p.T + t) : 1.0f;
        }
    };
    float v[U][4];
    if ((int64_t)blockIdx.x < n_tiles) load_chunk(blockIdx.x, 0, v);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * TM;
        float floor_db = -FLT_MAX;
        if (p.top_db >= 0.f) floor_db = 10.0f * log10f(fmaxf(__ldg(p.gmax + b / p.clips_per_group), 1e-10f)) - p.top_db;
        for (int kc = 0; kc < n_chunks; ++kc, stage ^= 1) {
            // the MMAs that read this stage two chunks ago must have completed before it is overwritten
            if (used[stage]) {
                wait(mbar + stage, ph_buf[stage]);
                ph_buf[stage] ^= 1;
                used[stage] = 0;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // A chunk: row = frame, column = mel band kc * 32 + (0..31).  A lane holds one 16-byte row of a core matrix (its
            // frame, 4 bands), so the hi and lo parts go out as two conflict-free 128-bit shared-memory stores.
            // dB through lg2.approx (2^-22 relative: 1e-5 dB, four orders below the parity budget on values of +-100 dB).
            unsigned char* ah = a_hi + stage * a_chunk;
            unsigned char* al = a_lo + stage * a_chunk;
            const int kw = min(KC, K - kc * KC);                  // bands in this chunk
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int c = warp + u * NW;
                const int fb = c / cores_k, kb = c - fb * cores_k;
                const int r = fb * 32 + lane;
                if (kb * 4 >= kw) continue;                       // beyond the last band: never read by the MMAs below
                const bool live = t0 + r < p.T;
                float4 hi, lo;
                float* hp = &hi.x;
                float* lp = &lo.x;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float d = live ? fmaxf(3.01029995663981195f * fast_lg2(fmaxf(v[u][j], 1e-10f)), floor_db) : 0.f;   // 10 log10(x), functional.py:390-404
                    split_tf32(d, hp[j], lp[j]);
                }
                const uint32_t off = (uint32_t)((r >> 3) * cores_k + kb) * 128 + (uint32_t)(r & 7) * 16;      // == op_offset(r, 4 kb, KC)
                *reinterpret_cast<float4*>(ah + off) = hi;
                *reinterpret_cast<float4*>(al + off) = lo;
            }
            // the next chunk's (or the next tile's first chunk's) loads go out now and land behind the barrier and the MMA issue
            if (kc + 1 < n_chunks) load_chunk(tile, kc + 1, v);
            else if (tile + gridDim.x < n_tiles) load_chunk(tile + gridDim.x, 0, v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sa_hi = smem_u32(ah), sa_lo = smem_u32(al);
                    const uint32_t sb_hi = smem_u32(b_hi) + (uint32_t)kc * (KC >> 2) * 128, sb_lo = smem_u32(b_lo) + (uint32_t)kc * (KC >> 2) * 128;
                    for (int ks = 0; ks < kw / 8; ++ks) {              // one MMA consumes K = 8 (two core matrices of 4 TF32)
                        const uint32_t ko = (uint32_t)ks * 256;
                        mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo_a), make_desc(sb_hi + ko, 128, sbo_b), idesc, (kc | ks) ? 1u : 0u);
                        mma_tf32(tmem_d, make_desc(sa_lo + ko, 128, sbo_a), make_desc(sb_hi + ko, 128, sbo_b), idesc, 1);
                        mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo_a), make_desc(sb_lo + ko, 128, sbo_b), idesc, 1);
                    }
                    // the stage is free again when these MMAs have read it; the last chunk also signals the finished accumulator
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + stage)) : "memory");
                    if (kc == n_chunks - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + 2)) : "memory");
                }
                __syncwarp();
            }
            used[stage] = 1;
        }
        // wait for the accumulator
        wait(mbar + 2, ph_done);
        ph_done ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: a warp reads the TMEM lane quadrant (warp % 4) = 32 frames; warps 0-3 take coefficient chunks 0 and 2,
        // warps 4-7 chunk 1 (16 columns each).  Stores are coalesced along frames ([n_mfcc, T] frequency-major).
        const int quad = warp & 3;
        for (int chunk = warp >> 2; chunk * 16 < NP; chunk += THREADS / 128) {
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
        // the next tile's first MMA overwrites the accumulator: everybody has read it
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
    const size_t smem = (size_t)4 * tc::TM * tc::KC * 4 + (size_t)2 * tc::NP * n_mels * 4 + 64;
    ACIDS_REQUIRE(cudaFuncSetAttribute(tc::mfcc_dct_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                  ACIDS_ECUDA, "mfcc_dct_tc: cannot reserve %zu B of shared memory", smem);
    int64_t grid = B * ((n_frames + tc::TM - 1) / tc::TM);
    if (grid > 2 * (int64_t)num_sms()) grid = 2 * (int64_t)num_sms();
    tc::mfcc_dct_tc_kernel<<<(unsigned)grid, tc::THREADS, smem, st>>>(p);
    ACIDS_CHECK_LAUNCH("mfcc_dct_tc");
    return ACIDS_OK;
}
