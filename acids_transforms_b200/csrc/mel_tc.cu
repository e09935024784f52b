// mel_tc.cu — the dense mel projection (|X|^p rows x filterbank, mel.py:38-44 / torchaudio MelScale) on the 5th-generation
// tensor cores: SURVEY.md §7 hard part 2 / north_star "mel filterbank ... as tensor-core GEMMs", next to the banded FP32
// epilogue the fused kernels use (DESIGN.md §6 has the measured comparison: the bank is 98 % zeros, so the banded form does
// 1/60 of the multiplies and never materialises the spectrum; this kernel is the dense formulation, numerically faithful).
//
// Per tile of 128 frames:  D[128 frames x 128 mels] = S[128 x F] . BANK[F x 128], F = n_fft / 2 + 1 walked in chunks of MEL_KC.
//   * both operands K-major in shared memory, no swizzle (8-row x 16-byte core matrices), UMMA descriptors, 3xTF32 operand
//     split (hi + lo, three MMAs per K step of 8) for fp32 fidelity, accumulator in 128 TMEM columns;
//   * the bank is split and laid out in the UMMA order ONCE per call by a small pack kernel; every chunk is then one
//     contiguous block that ONE thread moves with a bulk asynchronous copy (cp.async.bulk, the TMA engine) into a ring of
//     four stages, two chunks ahead of its MMAs, completing on the stage's mbarrier — no warp touches the B operand;
//   * the A chunk (128 frames x MEL_KC bins) is loaded a chunk ahead, split and stored by the warps as conflict-free 16-byte
//     rows while the tensor core multiplies the previous chunk (two stages, tcgen05.commit -> mbarrier frees a stage);
//   * epilogue: tcgen05.ld 32x32b, stores coalesced along frames into [n_mels, T] (frequency-major like MelSpectrogram).
#include "common.cuh"
#include "tc_common.cuh"

namespace acids {
namespace tc {

constexpr int MEL_TM = 128;       // frames per tile = MMA M
constexpr int MEL_N = 128;        // mel bands = MMA N (banks with fewer bands are zero padded)
#ifndef ACIDS_MEL_TC_KC
#define ACIDS_MEL_TC_KC 16
#endif
#ifndef ACIDS_MEL_TC_CTAS
#define ACIDS_MEL_TC_CTAS 2
#endif
constexpr int MEL_KC = ACIDS_MEL_TC_KC;   // bins per pipeline stage (16: 96 KB of rings per CTA, two CTAs per SM)
constexpr int MEL_THREADS = 256;
constexpr uint32_t MEL_A_BYTES = MEL_TM * MEL_KC * 4;      // one of hi / lo of an A chunk: 16 KB
constexpr uint32_t MEL_B_BYTES = MEL_N * MEL_KC * 4;       // one of hi / lo of a B chunk
constexpr int MEL_DEPTH = 3;                               // chunks of the spectrum a thread keeps in flight
constexpr int MEL_BSTAGES = 4;                             // B ring: a chunk's copy is issued two chunks before its MMAs

struct MelParams {
    const float* spec;        // [B, T, F] non-negative rows (|X| or |X|^2), unit stride along bins
    int64_t B, T;
    int F;
    const float* bank_packed; // [n_chunks][hi, lo][MEL_N x MEL_KC] in UMMA order (mel_tc_pack_kernel)
    int n_mels;
    float* out;               // [B, n_mels, T]
};

// bank [F, n_mels] row-major -> per chunk of 32 bins: hi block then lo block, element (mel n, bin k) at the K-major
// interleaved offset ((n >> 3) * 8 + (k >> 2)) * 128 + (n & 7) * 16 + (k & 3) * 4; zeros beyond F and n_mels
__global__ void mel_tc_pack_kernel(const float* __restrict__ bank, int F, int n_mels, int n_chunks, float* __restrict__ packed) {
    const int total = n_chunks * MEL_N * MEL_KC;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i / (MEL_N * MEL_KC), r = i - c * (MEL_N * MEL_KC);
        const int kk = r / MEL_N, n = r - kk * MEL_N;          // consecutive threads: consecutive mels of one bin (coalesced reads)
        const int k = c * MEL_KC + kk;
        float hi, lo;
        split_tf32((k < F && n < n_mels) ? __ldg(bank + (size_t)k * n_mels + n) : 0.f, hi, lo);
        const uint32_t off = (uint32_t)(((n >> 3) * (MEL_KC >> 2) + (kk >> 2)) * 128 + (n & 7) * 16 + (kk & 3) * 4) / 4;
        float* dst = packed + (size_t)c * (2 * MEL_N * MEL_KC);
        dst[off] = hi;
        dst[MEL_N * MEL_KC + off] = lo;
    }
}

__global__ void __launch_bounds__(MEL_THREADS, ACIDS_MEL_TC_CTAS) mel_tc_kernel(const MelParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // A ring: 2 stages of (hi | lo); B ring: MEL_BSTAGES stages of (hi | lo), filled two chunks ahead of their use
    constexpr uint32_t A_STAGE = 2 * MEL_A_BYTES, B_STAGE = 2 * MEL_B_BYTES;
    unsigned char* b_ring = smem + 2 * A_STAGE;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(b_ring + MEL_BSTAGES * B_STAGE);   // [0], [1]: A stage free (its MMAs are done); [2]: accumulator done; [3 ...]: B stage landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3 + MEL_BSTAGES);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F = p.F;
    const int n_chunks = (F + MEL_KC - 1) / MEL_KC;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MEL_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 3 + MEL_BSTAGES; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar + i)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const uint32_t idesc = make_idesc_mn(MEL_TM, MEL_N);
    constexpr uint32_t sbo = (uint32_t)(MEL_KC >> 2) * 128;      // 8-row groups are 8 core matrices (1 KB) apart, both operands
    uint32_t ph_free[2] = {0, 0}, ph_done = 0;
    uint32_t used[2] = {0, 0};
    uint32_t ph_b = 0;            // bit s: parity the B stage s completes with next
    // this CTA's chunks, numbered across its tiles: chunk gc multiplies bank chunk gc % n_chunks from B stage gc % MEL_BSTAGES,
    // and its copy is issued while chunk gc - 2 is being prepared (the stage's last reader, chunk gc - MEL_BSTAGES, is long done:
    // the A-stage wait of chunk gc - 2 already covered chunk gc - 4)
    const int64_t my_tiles = (int64_t)blockIdx.x < (p.B * ((p.T + MEL_TM - 1) / MEL_TM)) ? ((p.B * ((p.T + MEL_TM - 1) / MEL_TM)) - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t my_chunks = my_tiles * n_chunks;
    int64_t gc = 0;
    auto issue_b = [&](int bs, int bank_chunk) {          // thread 0 only: bank chunk -> B stage bs
        const uint32_t bytes = B_STAGE;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar + 3 + bs)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(b_ring + bs * B_STAGE)),
                     "l"(p.bank_packed + (size_t)bank_chunk * (2 * MEL_N * MEL_KC)), "r"(bytes), "r"(smem_u32(mbar + 3 + bs))
                     : "memory");
    };
    if (tid == 0) {
        if (my_chunks > 0) issue_b(0, 0);
        if (my_chunks > 1) issue_b(1, 1 % n_chunks);
    }

    const int64_t tiles_per_clip = (p.T + MEL_TM - 1) / MEL_TM;
    const int64_t n_tiles = p.B * tiles_per_clip;
    // A chunk = 128 frames x 8 blocks of 4 bins.  A warp-load covers 8 frames x 4 blocks (lane = frame % 8 + 8 * block): its
    // 16-byte rows land as four contiguous 128-byte runs in shared memory (conflict free); 32 such pieces per chunk, 4 per warp.
    constexpr int HALVES = MEL_KC / 16;                       // groups of four 4-bin blocks per chunk
    constexpr int U = 16 * HALVES / (MEL_THREADS / 32);       // pieces per warp
    static_assert(U >= 1 && U * (MEL_THREADS / 32) == 16 * HALVES, "a pass of the warps covers a chunk");
    const int r8 = lane & 7, kbl = lane >> 3;
    // base = first frame of the tile, rows = frames of the clip from there on (both fixed per tile: no division per chunk)
    auto load_chunk = [&](const float* __restrict__ base, int rows, int kc, float (&v)[U][4]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int piece = warp * U + u;                     // frame group (16) x group of four blocks (HALVES)
            const int fg = piece / HALVES, half = piece % HALVES;
            const int r = fg * 8 + r8;
            const int k = kc * MEL_KC + (half * 4 + kbl) * 4;
            const float* src = base + (int64_t)r * F + k;
#pragma unroll
            for (int j = 0; j < 4; ++j) v[u][j] = (r < rows && k + j < F) ? __ldg(src + j) : 0.f;
        }
    };
    auto tile_base = [&](int64_t tile, const float*& base, int& rows) {
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * MEL_TM;
        base = p.spec + (b * p.T + t0) * (int64_t)F;
        rows = (int)((p.T - t0) < MEL_TM ? (p.T - t0) : MEL_TM);
    };
    // The kernel is bound by the bytes it keeps in flight (Little's law: with one chunk of prefetch a thread has 32 bytes out, an
    // SM 16 KB, i.e. ~2.4 TB/s at a microsecond of latency): MEL_DEPTH register sets hold the chunks gc + 1 ... gc + MEL_DEPTH
    // while chunk gc is split and stored.  The chunks of this CTA are numbered across its tiles, so the prefetch runs through
    // tile boundaries; the loop is unrolled by MEL_DEPTH to keep the register sets statically indexed.
    // two cursors over the CTA's chunk sequence, advanced without divisions: the next chunk to LOAD and the chunk being multiplied
    int64_t l_tile = blockIdx.x, s_tile = blockIdx.x;
    int l_kc = 0, s_kc = 0, l_rows = 0;
    const float* l_base = p.spec;
    int64_t loaded = 0;
    if (my_chunks > 0) tile_base(l_tile, l_base, l_rows);
    auto load_next = [&](float (&vv)[U][4]) {
        if (loaded >= my_chunks) return;
        load_chunk(l_base, l_rows, l_kc, vv);
        ++loaded;
        if (++l_kc == n_chunks) {
            l_kc = 0;
            l_tile += gridDim.x;
            if (loaded < my_chunks) tile_base(l_tile, l_base, l_rows);
        }
    };
    auto step = [&](float (&v)[U][4]) {
        const int64_t tile = s_tile;
        const int kc = s_kc;
        const int stage = (int)(gc & 1);
        {
            unsigned char* st = smem + stage * A_STAGE;
            const int bs = (int)(gc % MEL_BSTAGES);
            // the MMAs that read this stage two chunks ago have completed
            if (used[stage]) {
                mbar_wait_parity(mbar + stage, ph_free[stage]);
                ph_free[stage] ^= 1;
                used[stage] = 0;
            }
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the bank chunk two ahead: its stage's last reader (chunk gc - 2, when MEL_BSTAGES = 4) completed with the wait above
            if (tid == 0 && gc + 2 < my_chunks) issue_b((bs + 2) % MEL_BSTAGES, (kc + 2) % n_chunks);
            // A chunk: split and store
            unsigned char* ah = st;
            unsigned char* al = st + MEL_A_BYTES;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int piece = warp * U + u;
                const int fg = piece / HALVES, half = piece % HALVES;
                float4 hi, lo;
                split_tf32(v[u][0], hi.x, lo.x);
                split_tf32(v[u][1], hi.y, lo.y);
                split_tf32(v[u][2], hi.z, lo.z);
                split_tf32(v[u][3], hi.w, lo.w);
                const uint32_t off = (uint32_t)(fg * (MEL_KC >> 2) + half * 4 + kbl) * 128 + (uint32_t)r8 * 16;   // row fg * 8 + r8, block half * 4 + kbl
                *reinterpret_cast<float4*>(ah + off) = hi;
                *reinterpret_cast<float4*>(al + off) = lo;
            }
            load_next(v);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            if (warp == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    mbar_wait_parity(mbar + 3 + bs, (ph_b >> bs) & 1u);      // the bank chunk has landed
                    const uint32_t sa_hi = smem_u32(ah), sa_lo = smem_u32(al);
                    const uint32_t sb_hi = smem_u32(b_ring + bs * B_STAGE), sb_lo = sb_hi + MEL_B_BYTES;
#pragma unroll
                    for (int ks = 0; ks < MEL_KC / 8; ++ks) {         // one MMA consumes K = 8 (two core matrices of 4 TF32)
                        const uint32_t ko = (uint32_t)ks * 256;
                        mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, (kc | ks) ? 1u : 0u);
                        mma_tf32(tmem_d, make_desc(sa_lo + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, 1);
                        mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_lo + ko, 128, sbo), idesc, 1);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + stage)) : "memory");
                    if (kc == n_chunks - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar + 2)) : "memory");
                }
                __syncwarp();
            }
            ph_b ^= 1u << bs;
            used[stage] = 1;
        }
        ++gc;
        if (++s_kc < n_chunks) return;
        s_kc = 0;
        s_tile += gridDim.x;
        const int64_t b = tile / tiles_per_clip;
        const int64_t t0 = (tile - b * tiles_per_clip) * MEL_TM;
        mbar_wait_parity(mbar + 2, ph_done);
        ph_done ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: a warp reads its TMEM lane quadrant (32 frames); warps 0-3 take the even 16-column chunks, warps 4-7 the odd ones
        const int quad = warp & 3;
        for (int chunk = warp >> 2; chunk * 16 < MEL_N; chunk += MEL_THREADS / 128) {
            if (chunk * 16 >= p.n_mels) break;
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
                float* o = p.out + (b * p.n_mels + chunk * 16) * p.T + t;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (chunk * 16 + n < p.n_mels) stg_stream1(o + (int64_t)n * p.T, __uint_as_float(r[n]));
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };
    float va[U][4], vb[U][4], vc[U][4];
    static_assert(MEL_DEPTH == 3, "three register sets below");
    load_next(va);
    load_next(vb);
    load_next(vc);
    while (gc < my_chunks) {
        step(va);
        if (gc < my_chunks) step(vb);
        if (gc < my_chunks) step(vc);
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(MEL_N) : "memory");
}


// ---- warp-specialised variant: no CTA barrier inside the tile ------------------------------------------------------------
// Warps 4-7 are PRODUCERS (load, split, store A chunks into a ring of WS_ASTAGES stages, `a_full` arrive per warp), lane 0 of
// warp 0 is the MMA ISSUER (waits `a_full` + the bank chunk's `b_full`, issues the MMAs, `tcgen05.commit` -> `a_free`; it also
// keeps the bank ring two chunks ahead), warps 0-3 run the EPILOGUE of a tile (each its TMEM lane quadrant) and hand the
// accumulator back through `acc_free`.  Stages change hands through mbarriers only, so the producers run ahead of the tensor
// core by up to WS_ASTAGES chunks and across tile boundaries (the next tile's chunks are filled during the epilogue).
constexpr int WS_ASTAGES = 3;

__device__ __forceinline__ void mbar_arrive_plain(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(MEL_THREADS, 2) mel_tc_ws_kernel(const MelParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr uint32_t A_STAGE = 2 * MEL_A_BYTES, B_STAGE = 2 * MEL_B_BYTES;
    unsigned char* a_ring = smem;
    unsigned char* b_ring = smem + WS_ASTAGES * A_STAGE;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + MEL_BSTAGES * B_STAGE);
    uint64_t* a_free = a_full + WS_ASTAGES;
    uint64_t* b_full = a_free + WS_ASTAGES;
    uint64_t* acc_done = b_full + MEL_BSTAGES;
    uint64_t* acc_free = acc_done + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int F = p.F;
    const int n_chunks = (F + MEL_KC - 1) / MEL_KC;
    static_assert(MEL_KC == 16, "the producer mapping below covers a 128 x 16 chunk with four warps");

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(MEL_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        auto init = [](uint64_t* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory"); };
        for (int i = 0; i < WS_ASTAGES; ++i) { init(a_full + i, 4); init(a_free + i, 1); }
        for (int i = 0; i < MEL_BSTAGES; ++i) init(b_full + i, 1);
        init(acc_done, 1);
        init(acc_free, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *tmem_slot;
    const int64_t tiles_per_clip = (p.T + MEL_TM - 1) / MEL_TM;
    const int64_t n_tiles = p.B * tiles_per_clip;
    const int64_t my_tiles = (int64_t)blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int64_t my_chunks = my_tiles * n_chunks;

    if (warp >= 4) {
        // ---------------- producers ----------------
        const int pw = warp - 4;
        const int r8 = lane & 7, kbl = lane >> 3;
        constexpr int U = 4;                                   // 16 pieces (8 frames x 16 bins) per chunk, four per warp
        float v[U][4];
        auto load_chunk = [&](int64_t tile, int kc) {
            const int64_t b = tile / tiles_per_clip;
            const int64_t t0 = (tile - b * tiles_per_clip) * MEL_TM;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int64_t t = t0 + (pw * U + u) * 8 + r8;
                const int k = kc * MEL_KC + kbl * 4;
                const float* src = p.spec + (b * p.T + t) * (int64_t)F + k;
#pragma unroll
                for (int j = 0; j < 4; ++j) v[u][j] = (t < p.T && k + j < F) ? __ldg(src + j) : 0.f;
            }
        };
        int64_t gc = 0;
        if (my_tiles > 0) load_chunk(blockIdx.x, 0);
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < n_chunks; ++kc, ++gc) {
                const int as = (int)(gc % WS_ASTAGES);
                if (gc >= WS_ASTAGES) mbar_wait_parity(a_free + as, (uint32_t)((gc / WS_ASTAGES - 1) & 1));
                unsigned char* ah = a_ring + as * A_STAGE;
                unsigned char* al = ah + MEL_A_BYTES;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    float4 hi, lo;
                    split_tf32(v[u][0], hi.x, lo.x);
                    split_tf32(v[u][1], hi.y, lo.y);
                    split_tf32(v[u][2], hi.z, lo.z);
                    split_tf32(v[u][3], hi.w, lo.w);
                    const uint32_t off = (uint32_t)((pw * U + u) * (MEL_KC >> 2) + kbl) * 128 + (uint32_t)r8 * 16;
                    *reinterpret_cast<float4*>(ah + off) = hi;
                    *reinterpret_cast<float4*>(al + off) = lo;
                }
                if (kc + 1 < n_chunks) load_chunk(tile, kc + 1);
                else if (tile + gridDim.x < n_tiles) load_chunk(tile + gridDim.x, 0);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
                __syncwarp();
                if (lane == 0) mbar_arrive_plain(a_full + as);
            }
        }
    } else {
        // ---------------- MMA issuer (warp 0, lane 0) and epilogue (warps 0-3) ----------------
        const uint32_t idesc = make_idesc_mn(MEL_TM, MEL_N);
        constexpr uint32_t sbo = (uint32_t)(MEL_KC >> 2) * 128;
        auto issue_b = [&](int64_t g) {
            const int bs = (int)(g % MEL_BSTAGES);
            const int bank_chunk = (int)(g % n_chunks);
            const uint32_t bytes = B_STAGE;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b_full + bs)), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(b_ring + bs * B_STAGE)),
                         "l"(p.bank_packed + (size_t)bank_chunk * (2 * MEL_N * MEL_KC)), "r"(bytes), "r"(smem_u32(b_full + bs))
                         : "memory");
        };
        if (tid == 0) {
            if (my_chunks > 0) issue_b(0);
            if (my_chunks > 1) issue_b(1);
        }
        int64_t gc = 0, ti = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
            if (warp == 0) {
                if (lane == 0) {
                    if (ti > 0) mbar_wait_parity(acc_free, (uint32_t)((ti - 1) & 1));       // the epilogue has read the accumulator
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kc = 0; kc < n_chunks; ++kc, ++gc) {
                        const int as = (int)(gc % WS_ASTAGES), bs = (int)(gc % MEL_BSTAGES);
                        // bank chunk gc + 2 goes into the stage chunk gc - 2 read: those MMAs have completed when its A stage is free
                        if (gc + 2 < my_chunks) {
                            if (gc >= 2) mbar_wait_parity(a_free + (int)((gc - 2) % WS_ASTAGES), (uint32_t)(((gc - 2) / WS_ASTAGES) & 1));
                            issue_b(gc + 2);
                        }
                        mbar_wait_parity(a_full + as, (uint32_t)((gc / WS_ASTAGES) & 1));
                        mbar_wait_parity(b_full + bs, (uint32_t)((gc / MEL_BSTAGES) & 1));
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t sa_hi = smem_u32(a_ring + as * A_STAGE), sa_lo = sa_hi + MEL_A_BYTES;
                        const uint32_t sb_hi = smem_u32(b_ring + bs * B_STAGE), sb_lo = sb_hi + MEL_B_BYTES;
#pragma unroll
                        for (int ks = 0; ks < MEL_KC / 8; ++ks) {
                            const uint32_t ko = (uint32_t)ks * 256;
                            mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, (kc | ks) ? 1u : 0u);
                            mma_tf32(tmem_d, make_desc(sa_lo + ko, 128, sbo), make_desc(sb_hi + ko, 128, sbo), idesc, 1);
                            mma_tf32(tmem_d, make_desc(sa_hi + ko, 128, sbo), make_desc(sb_lo + ko, 128, sbo), idesc, 1);
                        }
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(a_free + as)) : "memory");
                        if (kc == n_chunks - 1)
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(acc_done)) : "memory");
                    }
                }
                __syncwarp();
            }
            // epilogue of this tile: warp w reads TMEM lanes 32 w ... 32 w + 31 (frames), all 128 columns in chunks of 16
            mbar_wait_parity(acc_done, (uint32_t)(ti & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t b = tile / tiles_per_clip;
            const int64_t t0 = (tile - b * tiles_per_clip) * MEL_TM;
            for (int chunk = 0; chunk * 16 < p.n_mels; ++chunk) {
                uint32_t r[16];
                const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)(chunk * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int64_t t = t0 + warp * 32 + lane;
                if (t < p.T) {
                    float* o = p.out + (b * p.n_mels + chunk * 16) * p.T + t;
#pragma unroll
                    for (int n = 0; n < 16; ++n)
                        if (chunk * 16 + n < p.n_mels) stg_stream1(o + (int64_t)n * p.T, __uint_as_float(r[n]));
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_plain(acc_free);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(MEL_N) : "memory");
}

}  // namespace tc
}  // namespace acids

using namespace acids;

extern "C" ACIDS_API int64_t acids_mel_tc_workspace_bytes(int n_bins) {
    if (n_bins <= 0) return 0;
    const int64_t n_chunks = (n_bins + tc::MEL_KC - 1) / tc::MEL_KC;
    return n_chunks * 2 * tc::MEL_N * tc::MEL_KC * 4;
}

extern "C" ACIDS_API int acids_mel_tc(const float* spec, int64_t B, int64_t n_frames, int n_bins, const float* bank, int n_mels,
                            void* workspace, int64_t workspace_bytes, float* out, void* stream) {
    ACIDS_REQUIRE(spec && bank && out, ACIDS_EINVAL, "mel_tc: NULL pointer");
    ACIDS_REQUIRE(B >= 0 && n_frames >= 0 && n_bins >= 1, ACIDS_EINVAL, "mel_tc: bad sizes");
    ACIDS_REQUIRE(n_mels >= 1 && n_mels <= tc::MEL_N, ACIDS_EINVAL, "mel_tc: 1 <= n_mels <= %d (got %d)", tc::MEL_N, n_mels);
    if (B == 0 || n_frames == 0) return ACIDS_OK;
    ACIDS_REQUIRE(workspace && workspace_bytes >= acids_mel_tc_workspace_bytes(n_bins) && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, ACIDS_EINVAL,
                  "mel_tc: a 16-byte aligned workspace of %lld bytes is required", (long long)acids_mel_tc_workspace_bytes(n_bins));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_chunks = (n_bins + tc::MEL_KC - 1) / tc::MEL_KC;
    tc::mel_tc_pack_kernel<<<(n_chunks * 16 < 1024 ? n_chunks * 16 : 1024), 256, 0, st>>>(bank, n_bins, n_mels, n_chunks, static_cast<float*>(workspace));
    ACIDS_CHECK_LAUNCH("mel_tc_pack");
    tc::MelParams p{spec, B, n_frames, n_bins, static_cast<const float*>(workspace), n_mels, out};
    // the warp-specialised kernel is correct and measured SLOWER (0.340 vs 0.305 ms at the cfg-3 shape): four producer warps
    // per CTA do not keep up — the kernel is bound by the A operand's path (scalar loads of rows that are not 16-byte aligned,
    // split, store), not by the CTA barrier.  Opt-in: -DACIDS_MEL_TC_WS=1 or ACIDS_MEL_TC_WS=1 in the environment.
#ifndef ACIDS_MEL_TC_WS
#define ACIDS_MEL_TC_WS 0
#endif
    static const bool ws = tc::MEL_KC == 16 && (ACIDS_MEL_TC_WS || getenv("ACIDS_MEL_TC_WS") != nullptr);
    const size_t smem = (size_t)(ws ? tc::WS_ASTAGES : 2) * 2 * tc::MEL_A_BYTES + (size_t)tc::MEL_BSTAGES * 2 * tc::MEL_B_BYTES + 128;
    ACIDS_REQUIRE(cudaFuncSetAttribute(ws ? tc::mel_tc_ws_kernel : tc::mel_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess,
                  ACIDS_ECUDA, "mel_tc: cannot reserve %zu B of shared memory", smem);
    int64_t grid = B * ((n_frames + tc::MEL_TM - 1) / tc::MEL_TM);
    if (grid > (int64_t)ACIDS_MEL_TC_CTAS * num_sms()) grid = (int64_t)ACIDS_MEL_TC_CTAS * num_sms();
    if (ws) tc::mel_tc_ws_kernel<<<(unsigned)grid, tc::MEL_THREADS, smem, st>>>(p);
    else tc::mel_tc_kernel<<<(unsigned)grid, tc::MEL_THREADS, smem, st>>>(p);
    ACIDS_CHECK_LAUNCH("mel_tc");
    return ACIDS_OK;
}
