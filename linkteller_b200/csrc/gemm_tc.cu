// 3xTF32 GEMM on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM).
//
//   C[M,N] = A[M,K]·B (+bias) (ReLU)     — torch.mm of GraphConvolution.forward (gcn/layers.py:31)
//
// fp32 parity on tensor cores: every operand is split into two TF32 numbers, a = a_hi + a_lo with
// a_hi = rna_tf32(a), a_lo = rna_tf32(a − a_hi), and the product is accumulated in fp32 as
//   a_hi·b_hi + a_hi·b_lo + a_lo·b_hi          (the dropped a_lo·b_lo term is ≤ 2^-22 relative),
// i.e. three tcgen05.mma per logical MMA.
//
// Structure (one CTA of 8 warps per 128-row tile, persistent over tiles):
//   * B is pre-packed once per call by k_tc_prepack_b into the exact shared-memory image the MMA wants
//     — per 32-wide K chunk, [hi | lo] x [N_pad rows x 128 B], K-major, 128-byte swizzle — so a stage's B
//     operand is ONE 1-D bulk async copy (cp.async.bulk → mbarrier complete_tx);
//   * A rows (optionally gathered through a_rows) are loaded coalesced, split into hi/lo and stored
//     into the same canonical layout by all threads (fence.proxy.async → visible to the tensor core);
//   * one elected thread issues 12 MMAs (4 K-steps x 3 products, 128 x N x 8 each) per chunk and
//     commits them to the stage's mbarrier; two stages overlap staging with the MMAs;
//   * epilogue: tcgen05.ld 32x32b (warp w reads TMEM lanes 32·(w%4)…, column half w/4), bias / ReLU,
//     16-byte stores of the row.
#include "common.cuh"

namespace lt {

constexpr int TC_BM = 128;        // rows per tile = UMMA M (cta_group::1)
constexpr int TC_BK = 32;         // tf32 elements per K chunk = one 128-byte swizzle atom
constexpr int TC_THREADS = 256;
constexpr int TC_STAGES = 2;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint32_t to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = to_tf32(x);
    lo = to_tf32(x - __uint_as_float(hi));
}

// byte offset of element (row, k) inside a [rows x 32] K-major tile with 128-byte swizzle:
// 8-row groups of 1024 B, 16-byte chunks XOR-ed with the row index inside the group
__host__ __device__ __forceinline__ uint32_t sw128_offset(int row, int k) {
    return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((((k >> 2) ^ (row & 7)) & 7) << 4) + ((k & 3) << 2));
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (SBO), version 1
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address       bits [0,14)
    d |= (uint64_t)1 << 16;                           // leading byte offset bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                 // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version  bits [46,48)
    d |= (uint64_t)2 << 61;                           // SWIZZLE_128B        bits [61,64)
    return d;
}
// instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// B → [kchunks][2 (hi,lo)][n_pad rows][32 k] swizzled image.  One thread per (chunk, n, k).
__global__ void k_tc_prepack_b(const float* __restrict__ B, int64_t ldb, int transB, int32_t N, int32_t K, int32_t n_pad,
                               int32_t kchunks, uint32_t* __restrict__ img) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)kchunks * n_pad * TC_BK;
    if (t >= total) return;
    const int kk = (int)(t % TC_BK);
    const int nn = (int)((t / TC_BK) % n_pad);
    const int kc = (int)(t / ((int64_t)TC_BK * n_pad));
    const int k = kc * TC_BK + kk;
    float v = 0.f;
    if (nn < N && k < K) v = transB ? B[(int64_t)nn * ldb + k] : B[(int64_t)k * ldb + nn];
    uint32_t hi, lo;
    split_tf32(v, hi, lo);
    const size_t tile_words = (size_t)n_pad * TC_BK;                 // one [n_pad x 32] tile
    uint32_t* base = img + (size_t)kc * 2 * tile_words;
    const uint32_t off = sw128_offset(nn, kk) >> 2;
    base[off] = hi;
    base[tile_words + off] = lo;
}

// ---------------------------------------------------------------------------------------------
struct TcSmem {                        // carved from dynamic shared memory (1024-byte aligned)
    static constexpr int A_TILE = TC_BM * TC_BK * 4;    // 16 KB
};

template <int NT>   // N tile (multiple of 16, ≤ 256) = UMMA N
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_tc(const float* __restrict__ A, int64_t lda, const int32_t* __restrict__ a_rows, const uint32_t* __restrict__ Bimg,
          const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc, int32_t M, int32_t N, int32_t K,
          int32_t n_pad_total, int32_t kchunks) {
    constexpr int B_TILE = NT * TC_BK * 4;
    constexpr int STAGE = 2 * TcSmem::A_TILE + 2 * B_TILE;
    constexpr int TMEM_COLS = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t tma_bar[TC_STAGES], mma_bar[TC_STAGES];
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&tma_bar[s], 1); mbar_init(&mma_bar[s], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t idesc = make_idesc(TC_BM, NT);

    const int n0 = blockIdx.y * NT;                          // first column of this CTA's N tile
    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    uint32_t it = 0;                                         // chunk counter across tiles (stage = it & 1)

    for (int tile = blockIdx.x; tile < tiles_m; tile += gridDim.x) {
        const int m0 = tile * TC_BM;
        // rows this warp stages: 16 rows, lane ↔ k inside the chunk
        int64_t row_off[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int m = m0 + warp * 16 + q;
            row_off[q] = m < M ? (int64_t)(a_rows ? a_rows[m] : m) * lda : -1;
        }
        for (int kc = 0; kc < kchunks; ++kc, ++it) {
            const int s = it & 1;
            const uint32_t use = it >> 1;
            uint8_t* st = smem + (size_t)s * STAGE;
            uint8_t* a_hi = st;
            uint8_t* a_lo = st + TcSmem::A_TILE;
            uint8_t* b_hi = st + 2 * TcSmem::A_TILE;        // [hi | lo] contiguous, as in the packed image
            if (use > 0) mbar_wait(&mma_bar[s], (use - 1) & 1);   // MMAs that read this stage have completed
            if (tid == 0) {
                mbar_expect_tx(&tma_bar[s], 2 * B_TILE);
                const uint32_t* src = Bimg + ((size_t)kc * 2 * n_pad_total + 0) * TC_BK;
                // hi and lo tiles of this N tile: rows [n0, n0+NT) of each [n_pad_total x 32] tile
                bulk_g2s(b_hi, src + (size_t)n0 * TC_BK, B_TILE, &tma_bar[s]);
                bulk_g2s(b_hi + B_TILE, src + ((size_t)n_pad_total + n0) * TC_BK, B_TILE, &tma_bar[s]);
            }
            const int k = kc * TC_BK + lane;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                float v = 0.f;
                if (row_off[q] >= 0 && k < K) v = __ldg(A + row_off[q] + k);
                uint32_t hi, lo;
                split_tf32(v, hi, lo);
                const uint32_t off = sw128_offset(warp * 16 + q, lane);
                *reinterpret_cast<uint32_t*>(a_hi + off) = hi;
                *reinterpret_cast<uint32_t*>(a_lo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores → async proxy
            __syncthreads();
            if (tid == 0) {
                mbar_wait(&tma_bar[s], use & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint64_t da_hi = make_smem_desc(smem_u32(a_hi)), da_lo = make_smem_desc(smem_u32(a_lo));
                const uint64_t db_hi = make_smem_desc(smem_u32(b_hi)), db_lo = make_smem_desc(smem_u32(b_hi + B_TILE));
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);     // +32 bytes per K step inside the swizzle atom
                    umma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
                    umma_tf32(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
                    umma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
                }
                umma_commit(&mma_bar[s]);
            }
        }
        // ---- epilogue: wait for the last commit (covers every MMA of the tile)
        {
            const uint32_t last = it - 1;
            mbar_wait(&mma_bar[last & 1], (last >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // warp w may only touch TMEM lanes 32·(w%4)…+31; warps 4-7 take the upper column half when the
            // tile is wide enough to split (NT ≥ 64)
            constexpr int HALVES = NT >= 64 ? 2 : 1;
            constexpr int COLS_PER_HALF = NT / HALVES;
            const int q = warp & 3, half = warp >> 2;
            const int m = m0 + q * 32 + lane;
            const bool vec4 = (ldc % 4 == 0) && (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
            if (half < HALVES) {
#pragma unroll 1
                for (int c0 = half * COLS_PER_HALF; c0 < (half + 1) * COLS_PER_HALF; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    if (m < M) {
                        float* dst = C + (int64_t)m * ldc;
                        const int lim = min((half + 1) * COLS_PER_HALF - c0, 32);   // valid columns of this load
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const int col = n0 + c0 + j;
                            if (j >= lim || col >= N) continue;
                            float x[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                x[u] = __uint_as_float(v[j + u]);
                                if (bias && col + u < N) x[u] += __ldg(bias + col + u);
                                if (relu) x[u] = fmaxf(x[u], 0.f);
                            }
                            if (vec4 && j + 4 <= lim && col + 4 <= N) {
                                *reinterpret_cast<float4*>(dst + col) = make_float4(x[0], x[1], x[2], x[3]);
                            } else {
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    if (j + u < lim && col + u < N) dst[col + u] = x[u];
                            }
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();      // TMEM fully read before the next tile's first MMA overwrites it
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

template <int NT>
static int launch_tc(const float* A, int64_t lda, const int32_t* a_rows, const uint32_t* img, const float* bias, int relu,
                     float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t n_pad, int32_t kchunks,
                     cudaStream_t s) {
    constexpr int STAGE = 2 * TcSmem::A_TILE + 2 * NT * TC_BK * 4;
    const size_t smem = (size_t)TC_STAGES * STAGE + 1024;
    LT_CUDA(cudaFuncSetAttribute(k_gemm_tc<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    const int n_tiles = n_pad / NT;
    int gx = num_sms() / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > tiles_m) gx = tiles_m;
    k_gemm_tc<NT><<<dim3((unsigned)gx, (unsigned)n_tiles), TC_THREADS, smem, s>>>(A, lda, a_rows, img, bias, relu, C, ldc, M, N,
                                                                                K, n_pad, kchunks);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

// handled = false ⇒ the caller uses the fp32 SIMT kernel (tiny problems are not worth a tensor-core launch)
int gemm_3xtf32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
                const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, cudaStream_t s,
                bool force, bool* handled) {
    *handled = false;
    if (M <= 0 || N <= 0 || K <= 0) return LT_OK;
    if (!force && ((int64_t)M * N * K < (int64_t)1 << 24 || N < 16)) return LT_OK;
    // N tile: multiples of 16 up to 256; wider outputs loop over N tiles of 256 (grid.y)
    int nt;
    if (N <= 256) {
        nt = (N + 15) / 16 * 16;
        if (nt < 32) nt = 32;
        // instantiate a few widths only
        nt = nt <= 32 ? 32 : nt <= 48 ? 48 : nt <= 64 ? 64 : nt <= 128 ? 128 : 256;
    } else {
        nt = 256;
    }
    const int32_t n_pad = (N + nt - 1) / nt * nt;
    const int32_t kchunks = (K + TC_BK - 1) / TC_BK;
    Scratch scratch(s);
    uint32_t* img;
    const size_t words = (size_t)kchunks * 2 * n_pad * TC_BK;
    LT_TRY(scratch.alloc(&img, words));
    k_tc_prepack_b<<<(unsigned)ceil_div((int64_t)kchunks * n_pad * TC_BK, 256), 256, 0, s>>>(B, ldb, transB, N, K, n_pad,
                                                                                          kchunks, img);
    LT_LAUNCH_CHECK();
    switch (nt) {
        case 32: LT_TRY(launch_tc<32>(A, lda, a_rows, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s)); break;
        case 48: LT_TRY(launch_tc<48>(A, lda, a_rows, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s)); break;
        case 64: LT_TRY(launch_tc<64>(A, lda, a_rows, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s)); break;
        case 128: LT_TRY(launch_tc<128>(A, lda, a_rows, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s)); break;
        default: LT_TRY(launch_tc<256>(A, lda, a_rows, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s)); break;
    }
    *handled = true;
    return LT_OK;
}

}  // namespace lt
