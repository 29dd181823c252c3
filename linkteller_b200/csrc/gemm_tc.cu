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
#include "gemm_ws.cuh"

#include <stdlib.h>

namespace lt {

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

int tc_prepack_b(const float* B, int64_t ldb, int transB, int32_t N, int32_t K, int32_t n_pad, uint32_t* img, cudaStream_t s) {
    const int32_t kchunks = tc_kchunks(K);
    k_tc_prepack_b<<<(unsigned)ceil_div((int64_t)kchunks * n_pad * TC_BK, 256), 256, 0, s>>>(B, ldb, transB, N, K, n_pad,
                                                                                          kchunks, img);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

// option "tc_ws" = 0 (LT_TC_WS=0) selects the lock-step kernel (k_gemm_tc) instead of the warp-specialised one (k_gemm_ws)
bool tc_use_ws() { return opt(OPT_TC_WS) != 0; }

// handled = false ⇒ the caller uses the fp32 SIMT kernel (tiny problems are not worth a tensor-core launch)
int gemm_3xtf32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
                const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, cudaStream_t s,
                bool force, bool* handled) {
    *handled = false;
    if (M <= 0 || N <= 0 || K <= 0) return LT_OK;
    if (!force && ((int64_t)M * N * K < (int64_t)1 << 24 || N < 16)) return LT_OK;
    const int nt = tc_pick_nt(N);
    const int32_t n_pad = (N + nt - 1) / nt * nt;
    const int32_t kchunks = tc_kchunks(K);
    Scratch scratch(s);
    uint32_t* img;
    LT_TRY(scratch.alloc(&img, (size_t)kchunks * 2 * n_pad * TC_BK));
    LT_TRY(tc_prepack_b(B, ldb, transB, N, K, n_pad, img, s));
    const PlainA prod{A, lda, a_rows};
    if (tc_use_ws()) {
        switch (nt) {
            case 32: LT_TRY((launch_ws<32>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
            case 48: LT_TRY((launch_ws<48>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
            case 64: LT_TRY((launch_ws<64>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
            case 128: LT_TRY((launch_ws<128>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
            default: LT_TRY((launch_ws<256>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
        }
        *handled = true;
        return LT_OK;
    }
    switch (nt) {
        case 32: LT_TRY((launch_tc<32, 2>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
        case 48: LT_TRY((launch_tc<48, 2>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
        case 64: LT_TRY((launch_tc<64, 2>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
        case 128: LT_TRY((launch_tc<128, 1>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
        default: LT_TRY((launch_tc<256, 1>(prod, img, bias, relu, C, ldc, M, N, K, n_pad, kchunks, s))); break;
    }
    *handled = true;
    return LT_OK;
}

// sign mask + smallest magnitude per row of A·B + bias (TC_EPI_SIGN); N tiles are multiples of 32 so that a
// 32-column group of the epilogue is one whole word of the mask
int gemm_3xtf32_signs(const float* A, int64_t lda, const float* B, int64_t ldb, const float* bias, int32_t M, int32_t N,
                      int32_t K, uint32_t* sign_words, uint32_t* minabs_bits, cudaStream_t s) {
    if (M <= 0 || N <= 0 || K <= 0) return LT_OK;
    const int nt = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    const int32_t n_pad = (N + nt - 1) / nt * nt;
    const int32_t kchunks = tc_kchunks(K);
    const int32_t wpr = (N + 31) / 32;
    Scratch scratch(s);
    uint32_t* img;
    LT_TRY(scratch.alloc(&img, (size_t)kchunks * 2 * n_pad * TC_BK));
    LT_TRY(tc_prepack_b(B, ldb, 0, N, K, n_pad, img, s));
    const PlainA prod{A, lda, nullptr};
    if (tc_use_ws()) {
        switch (nt) {
            case 32: return launch_ws<32, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits);
            case 64: return launch_ws<64, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits);
            case 128: return launch_ws<128, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits);
            default: return launch_ws<256, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits);
        }
    }
    switch (nt) {
        case 32: LT_TRY((launch_tc<32, 2, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits))); break;
        case 64: LT_TRY((launch_tc<64, 2, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits))); break;
        case 128: LT_TRY((launch_tc<128, 1, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits))); break;
        default: LT_TRY((launch_tc<256, 1, PlainA, TC_EPI_SIGN>(prod, img, bias, 0, nullptr, 0, M, N, K, n_pad, kchunks, s, sign_words, wpr, minabs_bits))); break;
    }
    return LT_OK;
}

}  // namespace lt
