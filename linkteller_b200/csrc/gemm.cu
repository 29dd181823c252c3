// Dense fp32 GEMM  C[M,N] = A[M,K]·B (+bias) (ReLU) — torch.mm in GraphConvolution.forward
// (gcn/layers.py:31), with an optional row gather on A (P_v = X[v]·W1 for the sampled nodes) and
// an optional transposed B (LSA2 Gram matrix).
//
// algo 1: fp32 SIMT, 128x64x32 tiles, 8x4 register blocking, register-staged double buffering.
// algo 2: 3xTF32 on the 5th-gen tensor cores (tcgen05.mma kind::tf32, TMEM accumulators) — see
//         gemm_tc.cu.  algo 0 (auto) picks it for products above 2^24 multiply-adds with N ≥ 16.
// algo 3: float64 accumulation of the float32 products (k_gemm_f64acc), for the rows whose values must be exact.
#include "common.cuh"

namespace lt {

int gemm_3xtf32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
                const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K,
                cudaStream_t s, bool force, bool* handled);

constexpr int GM_BM = 128, GM_BN = 64, GM_BK = 32, GM_THREADS = 256;
constexpr int GM_APAD = 4;

template <bool TRANSB>
__global__ void __launch_bounds__(GM_THREADS)
k_gemm_simt(const float* __restrict__ A, int64_t lda, const int32_t* __restrict__ a_rows, const float* __restrict__ B,
            int64_t ldb, const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc, int M, int N,
            int K) {
    __shared__ __align__(16) float As[GM_BK][GM_BM + GM_APAD];  // As[k][m]
    __shared__ __align__(16) float Bs[GM_BK][GM_BN];            // Bs[k][n]
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GM_BM, n0 = blockIdx.x * GM_BN;
    const int ty = tid >> 4, tx = tid & 15;  // 16 x 16 threads → 8 rows x 4 cols each

    // global → register staging.  A tile: 128 rows x 32 k; a warp reads 32 consecutive k of one row.
    const int a_k = tid & 31, a_r0 = tid >> 5;  // rows a_r0 + 8*i, i<16
    int64_t a_off[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int m = m0 + a_r0 + 8 * i;
        a_off[i] = m < M ? (int64_t)(a_rows ? a_rows[m] : m) * lda : -1;
    }
    float ra[16], rb[8];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int k = k0 + a_k;
            ra[i] = (a_off[i] >= 0 && k < K) ? __ldg(A + a_off[i] + k) : 0.f;
        }
        if (!TRANSB) {  // B[K,N]: a warp reads 32 consecutive n of one k (two k rows per 64-wide tile pass)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = tid + i * GM_THREADS;       // 0..2047
                const int k = k0 + (idx >> 6), nn = n0 + (idx & 63);
                rb[i] = (k < K && nn < N) ? __ldg(B + (int64_t)k * ldb + nn) : 0.f;
            }
        } else {        // B stored [N,K]: read 32 consecutive k of one n
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = tid + i * GM_THREADS;
                const int nn = n0 + (idx >> 5), k = k0 + (idx & 31);
                rb[i] = (k < K && nn < N) ? __ldg(B + (int64_t)nn * ldb + k) : 0.f;
            }
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < 16; ++i) As[a_k][a_r0 + 8 * i] = ra[i];
        if (!TRANSB) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = tid + i * GM_THREADS;
                Bs[idx >> 6][idx & 63] = rb[i];
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int idx = tid + i * GM_THREADS;
                Bs[idx & 31][idx >> 5] = rb[i];
            }
        }
    };

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += GM_BK) {
        store_tiles();
        __syncthreads();
        if (k0 + GM_BK < K) load_tiles(k0 + GM_BK);
#pragma unroll
        for (int k = 0; k < GM_BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int nn = n0 + tx * 4 + j;
            if (nn >= N) continue;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + nn);
            if (relu) v = fmaxf(v, 0.f);
            C[(int64_t)m * ldc + nn] = v;
        }
    }
}

// algo 3: float32 operands, products and sums carried in float64, one rounding at the end.  For the few rows whose VALUES
// enter the result directly — P_v = X[v]·W1 of the perturbed nodes (ΔR = Â[r,v]·P[v] on the active units): the 3xTF32 /
// float32-accumulate error grows linearly with K (4e-5 of the row scale at K = 3170, profiles/h_error_probe.py), which
// cancellation between frontier rows amplified to 2.7e-3 on single influence entries of the twitch-shaped config.
constexpr int GD_ROWS = 8, GD_COLS = 64, GD_KG = 4, GD_KC = 256, GD_THREADS = GD_COLS * GD_KG;

// A CTA owns 8 rows x 64 columns; its 256 threads are 64 columns x 4 k-groups (k ≡ group mod 4), whose partial sums are
// added in group order at the end — deterministic, and 4x the CTAs and a quarter of the dependent FMA chain of a plain
// thread-per-column kernel (M = 500, K = 3170: 0.72 ms → tens of µs).
__global__ void __launch_bounds__(GD_THREADS)
k_gemm_f64acc(const float* __restrict__ A, int64_t lda, const int32_t* __restrict__ a_rows, const float* __restrict__ B,
              int64_t ldb, const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc, int M, int N, int K) {
    __shared__ double As[GD_ROWS][GD_KC];
    __shared__ double Ps[GD_KG][GD_ROWS][GD_COLS];
    const int tid = threadIdx.x;
    const int col = tid % GD_COLS, kg = tid / GD_COLS;
    const int m0 = blockIdx.y * GD_ROWS, nn = blockIdx.x * GD_COLS + col;
    double acc[GD_ROWS];
#pragma unroll
    for (int r = 0; r < GD_ROWS; ++r) acc[r] = 0.0;
    for (int k0 = 0; k0 < K; k0 += GD_KC) {
        for (int idx = tid; idx < GD_ROWS * GD_KC; idx += GD_THREADS) {
            const int r = idx / GD_KC, kk = idx % GD_KC;
            const int m = m0 + r, k = k0 + kk;
            As[r][kk] = (m < M && k < K) ? (double)__ldg(A + (int64_t)(a_rows ? a_rows[m] : m) * lda + k) : 0.0;
        }
        __syncthreads();
        if (nn < N) {
            const int kend = min(GD_KC, K - k0);
#pragma unroll 4
            for (int kk = kg; kk < kend; kk += GD_KG) {
                const double b = (double)__ldg(B + (int64_t)(k0 + kk) * ldb + nn);
#pragma unroll
                for (int r = 0; r < GD_ROWS; ++r) acc[r] = fma(As[r][kk], b, acc[r]);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < GD_ROWS; ++r) Ps[kg][r][col] = acc[r];
    __syncthreads();
    if (kg != 0 || nn >= N) return;
#pragma unroll
    for (int r = 0; r < GD_ROWS; ++r) {
        const int m = m0 + r;
        if (m >= M) break;
        double sum = Ps[0][r][col];
#pragma unroll
        for (int q = 1; q < GD_KG; ++q) sum += Ps[q][r][col];
        float v = (float)(bias ? sum + (double)__ldg(bias + nn) : sum);
        if (relu) v = fmaxf(v, 0.f);
        C[(int64_t)m * ldc + nn] = v;
    }
}

int gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
             const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int algo,
             cudaStream_t s) {
    if (M == 0 || N == 0) return LT_OK;
    if (algo == 3) {
        if (transB) { set_error("lt_gemm_f32: algo=3 (float64 accumulation) does not take a transposed B"); return LT_ERR_ARG; }
        dim3 grid((unsigned)ceil_div(N, GD_COLS), (unsigned)ceil_div(M, GD_ROWS));
        k_gemm_f64acc<<<grid, GD_THREADS, 0, s>>>(A, lda, a_rows, B, ldb, bias, relu, C, ldc, M, N, K);
        LT_LAUNCH_CHECK();
        return LT_OK;
    }
    if (algo == 0 || algo == 2) {
        bool handled = false;
        LT_TRY(gemm_3xtf32(A, lda, a_rows, B, ldb, transB, bias, relu, C, ldc, M, N, K, s, algo == 2, &handled));
        if (handled) return LT_OK;
        if (algo == 2) {
            set_error("lt_gemm_f32: algo=2 (3xTF32 tensor cores) does not cover M=%d N=%d K=%d transB=%d", M, N, K,
                      transB);
            return LT_ERR_ARG;
        }
    }
    dim3 grid((unsigned)ceil_div(N, GM_BN), (unsigned)ceil_div(M, GM_BM));
    if (transB)
        k_gemm_simt<true><<<grid, GM_THREADS, 0, s>>>(A, lda, a_rows, B, ldb, bias, relu, C, ldc, M, N, K);
    else
        k_gemm_simt<false><<<grid, GM_THREADS, 0, s>>>(A, lda, a_rows, B, ldb, bias, relu, C, ldc, M, N, K);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

}  // namespace lt

extern "C" int lt_gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb,
                           int transB, const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N,
                           int32_t K, int algo, lt_stream_t stream) {
    LT_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0 && lda >= K && ldc >= N, "lt_gemm_f32");
    LT_CHECK_ARG(transB ? ldb >= K : ldb >= N, "lt_gemm_f32: ldb");
    LT_CHECK_ARG(algo >= 0 && algo <= 3, "lt_gemm_f32: algo");
    return lt::gemm_f32(A, lda, a_rows, B, ldb, transB, bias, relu, C, ldc, M, N, K, algo, (cudaStream_t)stream);
}
