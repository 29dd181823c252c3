// CSR SpMM  C[i,:] = Σ_k Â[r_i,k]·B[k,:] (+bias) (ReLU)  — torch.spmm in GraphConvolution.forward
// (gcn/layers.py:32-34) and the restricted precompute H_R = Â[R,:]·P + b1 of the attack.
//
// HBM/L2-bound gather.  One warp per row SEGMENT (≤ SEG non-zeros), so hub rows of a scale-free
// graph are split across warps and the hardware CTA scheduler balances the load; rows that span
// several segments are reduced in segment order by a second tiny kernel (deterministic, no float
// atomics).  Lanes own float4 column slices (coalesced 16 B gathers of each B row); indices and
// values are read 32 at a time and broadcast by shuffle; 4 gathers in flight per lane.
#include "common.cuh"

namespace lt {

constexpr int SPMM_SEG = 256;      // non-zeros per warp work item
constexpr int SPMM_THREADS = 256;  // 8 warps per CTA

__global__ void k_spmm_segcount(const int32_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                                int32_t n_rows, int32_t* __restrict__ nseg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int32_t r = rows ? rows[i] : (int32_t)i;
    const int32_t len = indptr[r + 1] - indptr[r];
    nseg[i] = len <= SPMM_SEG ? 1 : (len + SPMM_SEG - 1) / SPMM_SEG;
}

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
    a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y); a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
}

// NV = float4 slices per lane (d ≤ 128·NV).  seg_off: exclusive scan of nseg (n_rows+1 entries).
template <int NV>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_vec(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
           const int32_t* __restrict__ rows, int32_t n_rows, const int32_t* __restrict__ seg_off,
           const float4* __restrict__ B4, int64_t ldb4, int32_t d4, const float* __restrict__ bias, int relu,
           float* __restrict__ C, int64_t ldc, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
    const int32_t total = seg_off[n_rows];
    if (g >= total) return;
    const int32_t i = (int32_t)upper_bound(seg_off, 0, (int64_t)n_rows + 1, (int32_t)g) - 1;
    const int32_t r = rows ? rows[i] : i;
    const int32_t first = seg_off[i];
    const int32_t nseg = seg_off[i + 1] - first;
    const int32_t row_beg = indptr[r], row_end = indptr[r + 1];
    const int32_t s0 = row_beg + ((int32_t)g - first) * SPMM_SEG;
    const int32_t s1 = min(row_end, s0 + SPMM_SEG);

    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int32_t base = s0; base < s1; base += 32) {
        const int32_t my = base + lane;
        const int32_t mc = my < s1 ? indices[my] : 0;
        const float mv = my < s1 ? values[my] : 0.f;
        const int cnt = min(32, s1 - base);
        int k = 0;
        for (; k + 4 <= cnt; k += 4) {
            float4 x[4][NV];
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int32_t c = __shfl_sync(LT_FULL, mc, k + u);
                v[u] = __shfl_sync(LT_FULL, mv, k + u);
                const float4* src = B4 + (int64_t)c * ldb4;
#pragma unroll
                for (int q = 0; q < NV; ++q) {
                    const int col = lane + 32 * q;
                    x[u][q] = col < d4 ? __ldg(src + col) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int q = 0; q < NV; ++q) fma4(acc[q], v[u], x[u][q]);
        }
        for (; k < cnt; ++k) {
            const int32_t c = __shfl_sync(LT_FULL, mc, k);
            const float v = __shfl_sync(LT_FULL, mv, k);
            const float4* src = B4 + (int64_t)c * ldb4;
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                const int col = lane + 32 * q;
                if (col < d4) fma4(acc[q], v, __ldg(src + col));
            }
        }
    }

    if (nseg == 1) {
        float* dst = C + (int64_t)i * ldc;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = lane + 32 * q;
            if (col < d4) {
                float4 a = acc[q];
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + col);
                    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                }
                if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                reinterpret_cast<float4*>(dst)[col] = a;
            }
        }
    } else {
        float4* dst = reinterpret_cast<float4*>(partial) + (int64_t)g * d4;
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            const int col = lane + 32 * q;
            if (col < d4) dst[col] = acc[q];
        }
    }
}

// rows split over several segments: sum the partial rows in segment order, then bias / ReLU
__global__ void k_spmm_fixup(const int32_t* __restrict__ seg_off, int32_t n_rows, const float* __restrict__ partial,
                             int32_t d, const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_rows) return;
    const int32_t first = seg_off[i], nseg = seg_off[i + 1] - first;
    if (nseg == 1) return;
    for (int col = lane; col < d; col += 32) {
        float a = 0.f;
        for (int32_t sgm = 0; sgm < nseg; ++sgm) a += partial[(int64_t)(first + sgm) * d + col];
        if (bias) a += bias[col];
        if (relu) a = fmaxf(a, 0.f);
        C[i * ldc + col] = a;
    }
}

// Generic fallback (d not a multiple of 4, unaligned, or d > 512): one warp per row, lanes are split
// into 32/G groups of G column lanes; groups take alternating non-zeros and are combined by shuffle.
template <int G>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_scalar(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
              const int32_t* __restrict__ rows, int32_t n_rows, const float* __restrict__ B, int64_t ldb, int32_t d,
              const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
    if (i >= n_rows) return;
    const int32_t r = rows ? rows[i] : (int32_t)i;
    const int sub = lane / G, cl = lane % G;
    constexpr int NSUB = 32 / G;
    const int32_t a = indptr[r], b = indptr[r + 1];
    for (int c0 = 0; c0 < d; c0 += G) {
        const int col = c0 + cl;
        float acc = 0.f;
        if (col < d)
            for (int32_t k = a + sub; k < b; k += NSUB) acc = fmaf(values[k], B[(int64_t)indices[k] * ldb + col], acc);
#pragma unroll
        for (int o = G; o < 32; o <<= 1) acc += __shfl_xor_sync(LT_FULL, acc, o);
        if (sub == 0 && col < d) {
            if (bias) acc += bias[col];
            if (relu) acc = fmaxf(acc, 0.f);
            C[i * ldc + col] = acc;
        }
    }
}

int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, const int32_t* rows,
                 int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu, float* C,
                 int64_t ldc, cudaStream_t s) {
    if (n_rows == 0 || d == 0) return LT_OK;
    const bool vec = (d % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && d <= 512 &&
                     ((uintptr_t)B % 16 == 0) && ((uintptr_t)C % 16 == 0) && (!bias || (uintptr_t)bias % 16 == 0);
    if (!vec) {
        const unsigned bl = (unsigned)ceil_div((int64_t)n_rows * 32, SPMM_THREADS);
#define LT_SPMM_SC(G_) k_spmm_scalar<G_><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, n_rows, B, ldb, d, bias, relu, C, ldc)
        if (d <= 1) LT_SPMM_SC(1);
        else if (d <= 2) LT_SPMM_SC(2);
        else if (d <= 4) LT_SPMM_SC(4);
        else if (d <= 8) LT_SPMM_SC(8);
        else if (d <= 16) LT_SPMM_SC(16);
        else LT_SPMM_SC(32);
#undef LT_SPMM_SC
        LT_LAUNCH_CHECK();
        return LT_OK;
    }
    Scratch scratch(s);
    int32_t* seg_off;
    LT_TRY(scratch.alloc(&seg_off, (size_t)n_rows + 1));
    k_spmm_segcount<<<(unsigned)ceil_div(n_rows, 256), 256, 0, s>>>(indptr, rows, n_rows, seg_off);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(seg_off, seg_off, n_rows, s));
    // total segments ≤ n_rows + nnz/SEG; bound it on the host without a sync: the grid is sized by
    // the upper bound and surplus warps exit at once.
    int32_t nnz_bound = 0;
    // indptr lives on the device; the caller-independent bound uses the row count of the whole matrix
    // which we do not know here, so read the scanned total back (4-byte copy, one sync per SpMM).
    LT_CUDA(cudaMemcpyAsync(&nnz_bound, seg_off + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    const int64_t total_seg = nnz_bound;
    float* partial = nullptr;
    const bool multi = total_seg > n_rows;
    if (multi) LT_TRY(scratch.alloc(&partial, (size_t)total_seg * d));
    const int32_t d4 = d / 4;
    const unsigned bl = (unsigned)ceil_div(total_seg * 32, SPMM_THREADS);
    const float4* B4 = reinterpret_cast<const float4*>(B);
#define LT_SPMM_V(NV_) k_spmm_vec<NV_><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, n_rows, seg_off, B4, ldb / 4, d4, bias, relu, C, ldc, partial)
    if (d4 <= 32) LT_SPMM_V(1);
    else if (d4 <= 64) LT_SPMM_V(2);
    else if (d4 <= 96) LT_SPMM_V(3);
    else LT_SPMM_V(4);
#undef LT_SPMM_V
    LT_LAUNCH_CHECK();
    if (multi) {
        k_spmm_fixup<<<(unsigned)ceil_div((int64_t)n_rows * 32, 256), 256, 0, s>>>(seg_off, n_rows, partial, d, bias,
                                                                                 relu, C, ldc);
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

}  // namespace lt

extern "C" int lt_spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values,
                               const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d,
                               const float* bias, int relu, float* C, int64_t ldc, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && B && C && n_rows >= 0 && d >= 0 && ldb >= d && ldc >= d,
                 "lt_spmm_csr_f32");
    return lt::spmm_csr_f32(indptr, indices, values, rows, n_rows, B, ldb, d, bias, relu, C, ldc,
                            (cudaStream_t)stream);
}
