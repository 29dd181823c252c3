// tcgen05 3xTF32 GEMM building block shared by gemm_tc.cu (plain A operand) and frontier.cu (A operand
// = ΔR computed on the fly).  See gemm_tc.cu for the design notes.
#pragma once
#include "common.cuh"

#include <type_traits>

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
// The same split with the rounding done by hand: round-to-nearest (ties away from zero) to 10 mantissa bits is
// "add half an ulp, clear the low 13 bits" on the bit pattern — two integer instructions instead of the ~6 ptxas
// emits for cvt.rna.tf32.f32 (which also screens Inf/NaN).  Identical to split_tf32 for finite inputs; Inf stays
// Inf, a NaN stays a NaN through the lo term.  The A-operand staging loop is issue-bound, so this matters there.
__device__ __forceinline__ uint32_t rna_tf32_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void split_tf32_fast(float x, uint32_t& hi, uint32_t& lo) {
    hi = rna_tf32_bits(x);
    lo = rna_tf32_bits(x - __uint_as_float(hi));
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
// A-operand producers.  A producer describes how element (row m, column k) of the A operand is
// obtained: per-row metadata (two offsets and a scalar, kept in shared memory per tile), the raw
// global loads for one element (prefetched one K chunk ahead) and the value computed from them.
//   meta(m, o0, o1, f)      per-row metadata (o0 < 0 marks a row past M)
//   Raw, raw4(o0, o1, k, K) raw operands of elements k..k+3 (k % 4 == 0; elements at or past K must come back as 0)
//   zero()                  the Raw of four zero elements
//   value4(raw, f)          the four elements from their raw operands and the row scalar
//   PD                      prefetch distance in K chunks (1 or 2): how many Raw sets are kept in registers
// four consecutive floats of a row: one 16-byte load when the address allows it, bounds-checked scalars otherwise
__device__ __forceinline__ float4 tc_load4(const float* p, int k, int K, bool vec) {
    if (vec && k + 4 <= K) return __ldg(reinterpret_cast<const float4*>(p + k));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) v.x = __ldg(p + k);
    if (k + 1 < K) v.y = __ldg(p + k + 1);
    if (k + 2 < K) v.z = __ldg(p + k + 2);
    if (k + 3 < K) v.w = __ldg(p + k + 3);
    return v;
}

struct PlainA {   // A[m,k] = A[a_rows ? a_rows[m] : m][k]
    const float* A; int64_t lda; const int32_t* a_rows;
    bool vec;     // rows are 16-byte aligned (A aligned, lda % 4 == 0)
    using Raw = float4;
    static constexpr int PD = 2;
    PlainA(const float* A_, int64_t lda_, const int32_t* a_rows_)
        : A(A_), lda(lda_), a_rows(a_rows_), vec(lda_ % 4 == 0 && (reinterpret_cast<uintptr_t>(A_) & 15) == 0) {}
    __device__ __forceinline__ void meta(int m, int64_t& o0, int64_t& o1, float& f) const {
        o0 = (int64_t)(a_rows ? a_rows[m] : m) * lda; o1 = 0; f = 0.f;
    }
    __device__ __forceinline__ Raw raw4(int64_t o0, int64_t, int k, int K) const { return tc_load4(A + o0, k, K, vec); }
    __device__ __forceinline__ Raw zero() const { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ float4 value4(const Raw& r, float) const { return r; }
};

struct TcSmem {
    static constexpr int A_TILE = TC_BM * TC_BK * 4;    // 16 KB
};

template <int NT>
constexpr size_t tc_smem_bytes() { return (size_t)TC_STAGES * (2 * TcSmem::A_TILE + 2 * NT * TC_BK * 4) + 1024; }

// C[M,N] = Aop[M,K]·B (+bias) (ReLU), B given as the packed image of k_tc_prepack_b.
//
// EPI selects what the epilogue keeps of a tile: TC_EPI_STORE writes C; TC_EPI_SIGN keeps only the sign of every
// element — one bit per element in sign_words[m·words_per_row + col/32] — and the row's smallest magnitude
// (atomicMin on the float bits into minabs_bits[m], pre-set to a large value), for callers that need to know
// which ReLUs fire and how far the closest one is from switching, not the values (frontier.cu).
constexpr int TC_EPI_STORE = 0, TC_EPI_SIGN = 1;

template <int NT, int MINB, class Prod, int EPI = TC_EPI_STORE>
__global__ void __launch_bounds__(TC_THREADS, MINB)
k_gemm_tc(Prod prod, const uint32_t* __restrict__ Bimg, const float* __restrict__ bias, int relu, float* __restrict__ C,
          int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t n_pad_total, int32_t kchunks,
          uint32_t* __restrict__ sign_words = nullptr, int32_t words_per_row = 0, uint32_t* __restrict__ minabs_bits = nullptr) {
    constexpr int B_TILE = NT * TC_BK * 4;
    constexpr int STAGE = 2 * TcSmem::A_TILE + 2 * B_TILE;
    constexpr int TMEM_COLS = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t tma_bar[TC_STAGES], mma_bar[TC_STAGES];
    __shared__ uint32_t tmem_base_s;
    __shared__ int64_t meta0[2][TC_BM], meta1[2][TC_BM];
    __shared__ float meta2[2][TC_BM];
    constexpr int NT32 = (NT + 31) / 32 * 32;                // the epilogue walks 32-column groups
    __shared__ __align__(16) float bias_s[NT32];             // this CTA's slice of the bias row (zeros without bias)

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
    const int n0 = blockIdx.y * NT;                          // first column of this CTA's N tile
    for (int j = tid; j < NT32; j += TC_THREADS) bias_s[j] = (bias && j < NT && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    const int my_tiles = blockIdx.x < tiles_m ? (tiles_m - 1 - blockIdx.x) / gridDim.x + 1 : 0;

    auto load_meta = [&](int tile_idx) {                     // threads 0..127: metadata of tile #tile_idx of this CTA
        if (tid < TC_BM && tile_idx < my_tiles) {
            const int m = (blockIdx.x + tile_idx * gridDim.x) * TC_BM + tid;
            int64_t o0 = -1, o1 = 0; float f = 0.f;
            if (m < M) prod.meta(m, o0, o1, f);
            meta0[tile_idx & 1][tid] = o0; meta1[tile_idx & 1][tid] = o1; meta2[tile_idx & 1][tid] = f;
        }
    };
    load_meta(0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    constexpr uint32_t idesc = make_idesc(TC_BM, NT);

    // raw operands of the NEXT chunk, prefetched while the tensor core works on the current one.
    // This warp stages rows warp*16 .. +15; a lane owns FOUR consecutive k of four rows (row sub-index lane/8,
    // 16-byte group lane%8), so one 16-byte load per operand and one 16-byte shared store per hi / lo image
    // cover four elements — the staging loop was instruction-issue bound with one element per load and store
    // (ncu r1g: issue slots 50 % busy, tensor pipe 7 %).
    // Prod::PD sets of raw operands live in registers: with PD = 2 the loads of chunk it+2 are issued while chunk it's
    // MMAs run, so a whole iteration passes before they are consumed (with one set the load latency sat on the
    // critical path of every iteration — ncu r1i: 35 % of the issue slots used, tensor pipe 12-31 %).
    constexpr int PD = Prod::PD;
    typename Prod::Raw pre[PD][4];
    const int st_row = warp * 16 + (lane >> 3), st_k4 = lane & 7;
    auto prefetch = [&](auto set_c, int tile_i, int kc_i) {
        constexpr int SET = decltype(set_c)::value;
        const int k = kc_i * TC_BK + st_k4 * 4;
        const int b = tile_i & 1;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t o0 = meta0[b][st_row + 4 * i];
            if (o0 >= 0 && k < K) pre[SET][i] = prod.raw4(o0, meta1[b][st_row + 4 * i], k, K);
            else pre[SET][i] = prod.zero();
        }
    };
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)kchunks;   // kchunks ≥ 2 (the launcher pads): chunk it+2 is in this or the next tile
    if (total) prefetch(std::integral_constant<int, 0>{}, 0, 0);
    if (PD == 2 && total > 1) prefetch(std::integral_constant<int, PD - 1>{}, 0, 1);
    // B chunk of iteration `it` → stage it&1, issued ONE iteration ahead (right after the MMAs of it-1 are
    // queued) so that the bulk copy's latency hides behind those MMAs and the next A staging
    auto issue_b = [&](uint32_t it, int kc_b) {
        const int sb = it & 1;
        uint8_t* b_dst = smem + (size_t)sb * STAGE + 2 * TcSmem::A_TILE;
        mbar_expect_tx(&tma_bar[sb], 2 * B_TILE);
        const uint32_t* src = Bimg + (size_t)kc_b * 2 * n_pad_total * TC_BK;
        bulk_g2s(b_dst, src + (size_t)n0 * TC_BK, B_TILE, &tma_bar[sb]);
        bulk_g2s(b_dst + B_TILE, src + ((size_t)n_pad_total + n0) * TC_BK, B_TILE, &tma_bar[sb]);
    };
    if (total && tid == 0) issue_b(0, 0);

    int tile_idx = 0, kc = 0;
    auto step = [&](auto par_c, uint32_t it) {
        constexpr int SET = PD == 2 ? decltype(par_c)::value : 0;      // register set holding chunk `it`
        const int s = it & 1;
        const uint32_t use = it >> 1;
        uint8_t* st = smem + (size_t)s * STAGE;
        uint8_t* a_hi = st;
        uint8_t* a_lo = st + TcSmem::A_TILE;
        uint8_t* b_hi = st + 2 * TcSmem::A_TILE;            // [hi | lo] contiguous, as in the packed image
        if (use > 0) mbar_wait(&mma_bar[s], (use - 1) & 1);  // MMAs that read this stage have completed
        {
            const int b = tile_idx & 1;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = st_row + 4 * i;
                const float4 v = prod.value4(pre[SET][i], meta2[b][r]);
                uint4 hi, lo;
                split_tf32_fast(v.x, hi.x, lo.x);
                split_tf32_fast(v.y, hi.y, lo.y);
                split_tf32_fast(v.z, hi.z, lo.z);
                split_tf32_fast(v.w, hi.w, lo.w);
                const uint32_t off = sw128_offset(r, st_k4 * 4);      // one whole 16-byte group of the swizzled row
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
        }
        if (kc == 0) load_meta(tile_idx + 1);                // visible after this iteration's barrier
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores → async proxy
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&tma_bar[s], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t da_hi = make_smem_desc(smem_u32(a_hi)), da_lo = make_smem_desc(smem_u32(a_lo));
            const uint64_t db_hi = make_smem_desc(smem_u32(b_hi)), db_lo = make_smem_desc(smem_u32(b_hi + B_TILE));
            const int ksteps = min(TC_BK / 8, (K - kc * TC_BK + 7) / 8);   // K steps past K multiply zeros: not issued
#pragma unroll
            for (int ks = 0; ks < TC_BK / 8; ++ks) {
                if (ks >= ksteps) break;
                const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);     // +32 bytes per K step inside the swizzle atom
                umma_tf32(tmem_base, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
                umma_tf32(tmem_base, da_hi + adv, db_lo + adv, idesc, 1u);
                umma_tf32(tmem_base, da_lo + adv, db_hi + adv, idesc, 1u);
            }
            umma_commit(&mma_bar[s]);
        }
        // next chunk's global loads fly while the MMAs run (and, at a tile end, during the epilogue)
        const bool tile_end = kc + 1 == kchunks;
        if (tid == 0 && it + 1 < total) {
            const uint32_t use1 = (it + 1) >> 1;             // the other stage: free once the MMAs of it-1 are done
            if (use1 > 0) mbar_wait(&mma_bar[s ^ 1], (use1 - 1) & 1);
            issue_b(it + 1, tile_end ? 0 : kc + 1);
        }
        if (it + PD < total) {                               // chunk it+PD → the register set chunk `it` just left
            int t2 = tile_idx, k2 = kc + PD;
            if (k2 >= kchunks) { k2 -= kchunks; ++t2; }
            prefetch(std::integral_constant<int, SET>{}, t2, k2);
        }
        if (tile_end) {
            // ---- epilogue: the last commit covers every MMA of the tile
            mbar_wait(&mma_bar[s], use & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // warp w may only touch TMEM lanes 32·(w%4)…+31; warps 4-7 take the upper column half when the
            // tile is wide enough to split (NT ≥ 64)
            constexpr int HALVES = NT >= 64 ? 2 : 1;
            constexpr int COLS_PER_HALF = NT / HALVES;
            const int q = warp & 3, half = warp >> 2;
            const int m_base = (blockIdx.x + tile_idx * gridDim.x) * TC_BM + q * 32;
            const bool vec4 = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
            // A lane holds 32 columns of ONE row after the TMEM load; stored as is, a warp-wide store would touch 32
            // different 128-byte lines (ncu r1g: the epilogue was half of the kernel, one L1 tag per lane per store).
            // The 32x32 block is transposed through this warp's 4 KB of the stage-s A buffer (free: every MMA of the
            // tile has completed; 16-byte chunks XOR-swizzled by the row, conflict-free both ways) so that eight
            // lanes write one row's 128 contiguous bytes and a store instruction covers four full lines.
            uint8_t* xp = st + warp * 4096;
            if (half < HALVES) {
#pragma unroll 1
                for (int c0 = half * COLS_PER_HALF; c0 < (half + 1) * COLS_PER_HALF; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
                    const int lim = min((half + 1) * COLS_PER_HALF - c0, 32);   // valid columns of this load
                    if (EPI == TC_EPI_SIGN) {
                        // lane = row: its 32 columns make exactly one word of the row's sign mask
                        uint32_t word = 0;
                        float mn = __uint_as_float(0x7f7f7f7fu), mx = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x = __uint_as_float(v[j]) + bias_s[c0 + j];
                            if (j < lim && n0 + c0 + j < N) {
                                word |= (x > 0.f ? 1u : 0u) << j;
                                mn = fminf(mn, fabsf(x));
                                mx = fmaxf(mx, fabsf(x));
                            }
                        }
                        const int m = m_base + lane;
                        if (m < M && n0 + c0 < N) {
                            sign_words[(int64_t)m * words_per_row + ((n0 + c0) >> 5)] = word;
                            atomicMin(minabs_bits + m, __float_as_uint(mn));
                            atomicMax(minabs_bits + M + m, __float_as_uint(mx));    // [M, 2M): the row's largest magnitude
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        // bias from shared memory (one broadcast float4 per four columns)
                        const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + j);
                        float4 x = make_float4(__uint_as_float(v[j]) + b4.x, __uint_as_float(v[j + 1]) + b4.y,
                                               __uint_as_float(v[j + 2]) + b4.z, __uint_as_float(v[j + 3]) + b4.w);
                        if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
                        *reinterpret_cast<float4*>(xp + lane * 128 + ((((j >> 2) ^ lane) & 7) << 4)) = x;
                    }
                    __syncwarp();
                    const int c4 = lane & 7;
                    const int col = n0 + c0 + c4 * 4;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int rl = 4 * i + (lane >> 3);
                        const float4 x = *reinterpret_cast<const float4*>(xp + rl * 128 + (((c4 ^ rl) & 7) << 4));
                        const int m = m_base + rl;
                        if (m < M && c4 * 4 < lim && col < N) {
                            float* dst = C + (int64_t)m * ldc + col;
                            if (vec4 && c4 * 4 + 4 <= lim && col + 4 <= N) {
                                *reinterpret_cast<float4*>(dst) = x;
                            } else {
                                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                                for (int u = 0; u < 4; ++u)
                                    if (c4 * 4 + u < lim && col + u < N) dst[u] = xs[u];
                            }
                        }
                    }
                    __syncwarp();
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();      // TMEM fully read before the next tile's first MMA overwrites it
            kc = 0;
            ++tile_idx;
        } else {
            ++kc;
        }
    };
    for (uint32_t it = 0; it < total; it += 2) {            // unrolled by two: the register set of a chunk is static
        step(std::integral_constant<int, 0>{}, it);
        if (it + 1 < total) step(std::integral_constant<int, 1>{}, it + 1);
    }
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    }
}

// launch over M tiles (persistent CTAs) x N tiles; img = packed B of k_tc_prepack_b
template <int NT, int MINB, class Prod, int EPI = TC_EPI_STORE>
static int launch_tc(const Prod& prod, const uint32_t* img, const float* bias, int relu, float* C, int64_t ldc, int32_t M,
                     int32_t N, int32_t K, int32_t n_pad, int32_t kchunks, cudaStream_t s, uint32_t* sign_words = nullptr,
                     int32_t words_per_row = 0, uint32_t* minabs_bits = nullptr) {
    if (kchunks < 2) { set_error("launch_tc: kchunks must be >= 2 (tc_kchunks pads)"); return LT_ERR_INTERNAL; }
    const size_t smem = tc_smem_bytes<NT>();
    LT_CUDA(cudaFuncSetAttribute(k_gemm_tc<NT, MINB, Prod, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    const int n_tiles = n_pad / NT;
    int gx = num_sms() * MINB / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > tiles_m) gx = tiles_m;
    k_gemm_tc<NT, MINB, Prod, EPI><<<dim3((unsigned)gx, (unsigned)n_tiles), TC_THREADS, smem, s>>>(
        prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, sign_words, words_per_row, minabs_bits);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

// K chunks of a product: at least two, so that the kernel's two-chunk look-ahead never leaves the next tile
// (a padded chunk lies past K: its A and B are zeros and none of its K steps is issued)
static inline int32_t tc_kchunks(int32_t K) { const int32_t c = (K + TC_BK - 1) / TC_BK; return c < 2 ? 2 : c; }

// N tile used for an output width (a few instantiated widths only) and the padded width
static inline int tc_pick_nt(int N) {
    if (N > 256) return 256;
    int nt = (N + 15) / 16 * 16;
    return nt <= 32 ? 32 : nt <= 48 ? 48 : nt <= 64 ? 64 : nt <= 128 ? 128 : 256;
}

int tc_prepack_b(const float* B, int64_t ldb, int transB, int32_t N, int32_t K, int32_t n_pad, uint32_t* img, cudaStream_t s);
// signs of A[M,K]·B + bias: sign_words [M, (N+31)/32] and minabs_bits [2M]: [0,M) the row's smallest |value| (caller pre-sets
// 0x7f7f7f7f), [M,2M) its largest (caller pre-sets 0); gemm_tc.cu
int gemm_3xtf32_signs(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, int32_t M, int32_t N,
                      int32_t K, uint32_t* sign_words, uint32_t* minabs_bits, cudaStream_t s);

}  // namespace lt
