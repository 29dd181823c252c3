// Warp-specialised form of the tcgen05 3xTF32 GEMM (same operands, producers and epilogues as k_gemm_tc).
//
// k_gemm_tc runs every K chunk in lock step: all eight warps stage A, a CTA-wide barrier, one thread issues the MMAs,
// and the epilogue of a tile runs with the tensor core idle (ncu r1i: tensor pipe 12-31 % busy).  Here the roles are
// split and connected by mbarriers only:
//
//   warps 0-7   producers   raw A operands → value → hi/lo split → swizzled shared-memory stage; the loads of chunk
//                           c+1 are issued before chunk c is converted (two static register sets); lane 0 of warp 0
//                           also issues the bulk copy of the stage's B image
//   warp  8     MMA         one elected thread: wait full_a / full_b, issue 3 MMAs per K step, commit → empty[stage];
//                           the last chunk of a tile also commits → acc_full[accumulator]
//   warps 9-12  epilogue    tcgen05.ld of the finished accumulator (warp w owns TMEM lanes 32·(w%4)…), bias / ReLU,
//                           transposed coalesced stores (or the sign/min epilogue), arrive → acc_empty[accumulator]
//
// Two TMEM accumulators alternate between tiles, so the epilogue of tile t overlaps the main loop of tile t+1, and a
// ring of WS_STAGES shared-memory stages lets the producers run ahead of the tensor core.
#pragma once
#include "gemm_tc.cuh"

namespace lt {

constexpr int WS_PROD_WARPS = 8;
constexpr int WS_MMA_WARP = 8;
constexpr int WS_EPI_WARP0 = 9;
constexpr int WS_THREADS = (WS_PROD_WARPS + 1 + 4) * 32;   // 416

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int NT>
struct WsCfg {
    static constexpr int STAGES = NT >= 256 ? 2 : NT >= 128 ? 3 : 4;
    static constexpr size_t SMEM = (size_t)STAGES * (2 * TcSmem::A_TILE + 2 * NT * TC_BK * 4) + 4 * 4096 + 1024;
};

template <int NT, class Prod, int EPI>
__global__ void __launch_bounds__(WS_THREADS, 1)
k_gemm_ws(Prod prod, const uint32_t* __restrict__ Bimg, const float* __restrict__ bias, int relu, float* __restrict__ C,
          int64_t ldc, int32_t M, int32_t N, int32_t K, int32_t n_pad_total, int32_t kchunks,
          uint32_t* __restrict__ sign_words, int32_t words_per_row, uint32_t* __restrict__ minabs_bits) {
    constexpr int S = WsCfg<NT>::STAGES;
    constexpr int B_TILE = NT * TC_BK * 4;
    constexpr int A_STAGE = 2 * TcSmem::A_TILE;
    constexpr int NTC = NT <= 32 ? 32 : NT <= 64 ? 64 : NT <= 128 ? 128 : 256;   // TMEM columns of one accumulator
    constexpr int NT32 = (NT + 31) / 32 * 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* a_ring = smem;                                   // S x [hi 16 KB | lo 16 KB]
    uint8_t* b_ring = smem + (size_t)S * A_STAGE;             // S x [hi | lo]
    uint8_t* xpose = b_ring + (size_t)S * 2 * B_TILE;         // 4 x 4 KB, one per epilogue warp
    __shared__ uint64_t full_a[S], full_b[S], empty[S], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float bias_s[NT32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n0 = blockIdx.y * NT;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full_a[s], WS_PROD_WARPS); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == WS_MMA_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "n"(2 * NTC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int j = tid; j < NT32; j += WS_THREADS) bias_s[j] = (bias && j < NT && n0 + j < N) ? __ldg(bias + n0 + j) : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    const int my_tiles = blockIdx.x < tiles_m ? (tiles_m - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const uint32_t total = (uint32_t)my_tiles * (uint32_t)kchunks;   // kchunks ≥ 2

    if (warp < WS_PROD_WARPS) {
        // ------------------------------------------------------------------------------ producers
        const int st_row = warp * 16 + (lane >> 3), st_k4 = lane & 7;
        int64_t o0[4], o1[4];
        float fcur[4], fnxt[4];
        auto load_meta = [&](int tile_i) {
            const int m0 = (blockIdx.x + tile_i * gridDim.x) * TC_BM;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = m0 + st_row + 4 * i;
                o0[i] = -1; o1[i] = 0; fnxt[i] = 0.f;
                if (m < M) prod.meta(m, o0[i], o1[i], fnxt[i]);
            }
        };
        typename Prod::Raw pre[2][4];
        auto prefetch = [&](auto set_c, int kc_i) {
            constexpr int SET = decltype(set_c)::value;
            const int k = kc_i * TC_BK + st_k4 * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (o0[i] >= 0 && k < K) pre[SET][i] = prod.raw4(o0[i], o1[i], k, K);
                else pre[SET][i] = prod.zero();
            }
        };
        int tile_p = 0, kc_p = 0;           // position of the next chunk to prefetch
        int kc_c = 0;                       // K chunk of the chunk being converted
        int stage = 0;
        uint32_t phase = 0;                 // parity of this stage's current use
        if (total) {
            load_meta(0);
            prefetch(std::integral_constant<int, 0>{}, 0);
            kc_p = 1;
        }
        auto step = [&](auto par_c, uint32_t c) {
            constexpr int SET = decltype(par_c)::value;
            if (kc_c == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) fcur[i] = fnxt[i];
            }
            if (c + 1 < total) {
                if (kc_p == 0) load_meta(tile_p);            // the next chunk opens a tile: its rows' metadata first
                prefetch(std::integral_constant<int, SET ^ 1>{}, kc_p);
                if (++kc_p == kchunks) { kc_p = 0; ++tile_p; }
            }
            mbar_wait(&empty[stage], phase ^ 1);             // the MMAs that read this stage have completed
            if (warp == 0 && lane == 0) {
                uint8_t* b_dst = b_ring + (size_t)stage * 2 * B_TILE;
                mbar_expect_tx(&full_b[stage], 2 * B_TILE);
                const uint32_t* src = Bimg + (size_t)kc_c * 2 * n_pad_total * TC_BK;
                bulk_g2s(b_dst, src + (size_t)n0 * TC_BK, B_TILE, &full_b[stage]);
                bulk_g2s(b_dst + B_TILE, src + ((size_t)n_pad_total + n0) * TC_BK, B_TILE, &full_b[stage]);
            }
            uint8_t* a_hi = a_ring + (size_t)stage * A_STAGE;
            uint8_t* a_lo = a_hi + TcSmem::A_TILE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = st_row + 4 * i;
                const float4 v = prod.value4(pre[SET][i], fcur[i]);
                uint4 hi, lo;
                split_tf32_fast(v.x, hi.x, lo.x);
                split_tf32_fast(v.y, hi.y, lo.y);
                split_tf32_fast(v.z, hi.z, lo.z);
                split_tf32_fast(v.w, hi.w, lo.w);
                const uint32_t off = sw128_offset(r, st_k4 * 4);
                *reinterpret_cast<uint4*>(a_hi + off) = hi;
                *reinterpret_cast<uint4*>(a_lo + off) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores → async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_a[stage]);
            if (++kc_c == kchunks) kc_c = 0;
            if (++stage == S) { stage = 0; phase ^= 1; }
        };
        for (uint32_t c = 0; c < total; c += 2) {            // unrolled by two: the register set of a chunk is static
            step(std::integral_constant<int, 0>{}, c);
            if (c + 1 < total) step(std::integral_constant<int, 1>{}, c + 1);
        }
    } else if (warp == WS_MMA_WARP) {
        // ------------------------------------------------------------------------------ MMA issue
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(TC_BM, NT);
            int stage = 0, kc = 0, tile = 0;
            uint32_t phase = 0;
            for (uint32_t c = 0; c < total; ++c) {
                const int acc = tile & 1;
                if (kc == 0) mbar_wait(&acc_empty[acc], (((uint32_t)tile >> 1) & 1) ^ 1);   // epilogue done with it
                mbar_wait(&full_a[stage], phase);
                mbar_wait(&full_b[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint8_t* a_hi = a_ring + (size_t)stage * A_STAGE;
                uint8_t* b_hi = b_ring + (size_t)stage * 2 * B_TILE;
                const uint64_t da_hi = make_smem_desc(smem_u32(a_hi)), da_lo = make_smem_desc(smem_u32(a_hi + TcSmem::A_TILE));
                const uint64_t db_hi = make_smem_desc(smem_u32(b_hi)), db_lo = make_smem_desc(smem_u32(b_hi + B_TILE));
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NTC);
                const int ksteps = min(TC_BK / 8, (K - kc * TC_BK + 7) / 8);   // K steps past K multiply zeros: not issued
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    if (ks >= ksteps) break;
                    const uint64_t adv = (uint64_t)((ks * 8 * 4) >> 4);
                    umma_tf32(d_tmem, da_hi + adv, db_hi + adv, idesc, (kc | ks) ? 1u : 0u);
                    umma_tf32(d_tmem, da_hi + adv, db_lo + adv, idesc, 1u);
                    umma_tf32(d_tmem, da_lo + adv, db_hi + adv, idesc, 1u);
                }
                umma_commit(&empty[stage]);
                if (kc + 1 == kchunks) { umma_commit(&acc_full[acc]); kc = 0; ++tile; } else { ++kc; }
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------------------ epilogue
        const int q = warp & 3;                                // TMEM lane group this warp may touch
        uint8_t* xp = xpose + (warp - WS_EPI_WARP0) * 4096;
        const bool vec4 = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
        for (int tile = 0; tile < my_tiles; ++tile) {
            const int acc = tile & 1;
            mbar_wait(&acc_full[acc], ((uint32_t)tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int m_base = (blockIdx.x + tile * gridDim.x) * TC_BM + q * 32;
#pragma unroll 1
            for (int c0 = 0; c0 < NT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * NTC + c0), v);
                const int lim = min(NT - c0, 32);
                if (EPI == TC_EPI_SIGN) {
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
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[acc]);       // the accumulator may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == WS_MMA_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * NTC) : "memory");
    }
}

template <int NT, class Prod, int EPI = TC_EPI_STORE>
static int launch_ws(const Prod& prod, const uint32_t* img, const float* bias, int relu, float* C, int64_t ldc, int32_t M,
                     int32_t N, int32_t K, int32_t n_pad, int32_t kchunks, cudaStream_t s, uint32_t* sign_words = nullptr,
                     int32_t words_per_row = 0, uint32_t* minabs_bits = nullptr) {
    if (kchunks < 2) { set_error("launch_ws: kchunks must be >= 2 (tc_kchunks pads)"); return LT_ERR_INTERNAL; }
    const size_t smem = WsCfg<NT>::SMEM;
    LT_CUDA(cudaFuncSetAttribute(k_gemm_ws<NT, Prod, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_m = (M + TC_BM - 1) / TC_BM;
    const int n_tiles = n_pad / NT;
    int gx = num_sms() / n_tiles;
    if (gx < 1) gx = 1;
    if (gx > tiles_m) gx = tiles_m;
    k_gemm_ws<NT, Prod, EPI><<<dim3((unsigned)gx, (unsigned)n_tiles), WS_THREADS, smem, s>>>(
        prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, sign_words, words_per_row, minabs_bits);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

bool tc_use_ws();   // gemm_tc.cu

}  // namespace lt
