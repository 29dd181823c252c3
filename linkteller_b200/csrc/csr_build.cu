// Â construction: edge list → symmetric CSR, torch COO → CSR, row sums,
// normalisation with bit-exact indices and values.  All integer/byte work, HBM-bound.
#include "common.cuh"

namespace lt {

// ------------------------------------------------------------------ a1: edges → A = M + Mᵀ
__global__ void k_edge_keys(const int64_t* __restrict__ edges, int64_t m, int32_t n,
                            uint64_t* __restrict__ keys, int* __restrict__ err) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= m) return;
    int64_t r = edges[2 * e], c = edges[2 * e + 1];
    if (r < 0 || r >= n || c < 0 || c >= n) { *err = 1; r = 0; c = 0; }
    keys[e] = ((uint64_t)r << 32) | (uint64_t)c;
    keys[m + e] = ((uint64_t)c << 32) | (uint64_t)r;
}

__global__ void k_heads(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ head) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// pos = exclusive scan of head (n+1 entries).  For every run start write its column, its start
// index, and the row pointers of all rows that begin at or before it.
__global__ void k_unique_emit(const uint64_t* __restrict__ keys, const int32_t* __restrict__ pos, int64_t n,
                              int32_t n_rows, int32_t* __restrict__ indices, int32_t* __restrict__ starts,
                              int32_t* __restrict__ indptr, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t p = pos[i];
    if (pos[i + 1] == p) return;  // not a run start
    const uint64_t k = keys[i];
    const int32_t r = (int32_t)(k >> 32);
    if (p < cap) { indices[p] = (int32_t)(k & 0xffffffffu); starts[p] = (int32_t)i; }
    int32_t prev_r = (i == 0) ? -1 : (int32_t)(keys[i - 1] >> 32);
    for (int32_t rr = prev_r + 1; rr <= r; ++rr) indptr[rr] = p;
    if (i == n - 1 || pos[i + 1] == pos[n]) {  // last run: close the remaining rows
        // (only the final run start satisfies pos[i+1]==total)
        for (int32_t rr = r + 1; rr <= n_rows; ++rr) indptr[rr] = pos[n];
    }
}

__global__ void k_run_values(const int32_t* __restrict__ starts, int64_t nnz, int64_t n_keys, int binarize,
                             float* __restrict__ values) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int64_t end = (p + 1 < nnz) ? starts[p + 1] : n_keys;
    values[p] = binarize ? 1.0f : (float)(end - starts[p]);
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ a3⁻¹: sorted COO → CSR
__global__ void k_coo_rowptr(const int64_t* __restrict__ row, int64_t nnz, int32_t n,
                             int32_t* __restrict__ indptr) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    indptr[r] = (int32_t)lower_bound(row, 0, nnz, (int64_t)r);
}
__global__ void k_narrow_cols(const int64_t* __restrict__ col, int64_t nnz, int32_t* __restrict__ indices) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nnz) indices[i] = (int32_t)col[i];
}
__global__ void k_csr_check(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int32_t n,
                            int64_t nnz, int* __restrict__ bad) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    int32_t a = indptr[r], b = indptr[r + 1];
    if (a > b || a < 0 || b > nnz || (r == 0 && a != 0) || (r == n - 1 && b != nnz)) { *bad = 1; return; }
    for (int32_t k = a; k < b; ++k) {
        int32_t c = indices[k];
        if (c < 0 || c >= n || (k > a && indices[k - 1] >= c)) { *bad = 1; return; }
    }
}

// ------------------------------------------------------------------ a2: row sums + normalisation
template <typename T>
__global__ void k_row_sums(const int32_t* __restrict__ indptr, const T* __restrict__ values, int32_t n,
                           T* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    T s = 0;
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) s = s + values[k];  // column order, like coo @ ones
    out[r] = s;
}

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// value of output entry for input entry (r, c, a), excluding the identity (FirstOrderGCN) /
// including it (AugNormAdj).  Returned as double; 0 means "dropped".
template <typename T, int KIND>
__device__ __forceinline__ double norm_value(T di, T a, T dj, bool diag) {
    if (KIND == LT_NORM_FIRST_ORDER_GCN) {
        T v = mul_rn(mul_rn(di, a), dj);       // (D·A)·D, each product rounded in the working dtype
        return (double)v;                      // caller adds 1.0 on the diagonal in float64 (sp.eye)
    } else {
        double aa = (double)a + (diag ? 1.0 : 0.0);  // adj + sp.eye, float64
        return __dmul_rn(__dmul_rn((double)di, aa), (double)dj);
    }
}

// PASS 0: count output entries per row.  PASS 1: write them (column order, diagonal merged in).
template <typename T, int KIND, int PASS>
__global__ void k_normalize(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const T* __restrict__ values, int32_t n, const T* __restrict__ dis,
                            int32_t* __restrict__ out_indptr, int32_t* __restrict__ out_indices,
                            float* __restrict__ out_values) {
    int64_t r64 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r64 >= n) return;
    const int32_t r = (int32_t)r64;
    const T di = dis[r];
    int32_t w = PASS ? out_indptr[r] : 0;
    bool diag_done = false;
    auto emit = [&](int32_t c, double v) {
        if (v != 0.0) {
            if (PASS) { out_indices[w] = c; out_values[w] = (float)v; }
            ++w;
        }
    };
    auto diag_only = [&]() {  // identity entry for a row without a stored diagonal
        double v = (KIND == LT_NORM_FIRST_ORDER_GCN) ? 1.0 : __dmul_rn(__dmul_rn((double)di, 1.0), (double)di);
        emit(r, v);
    };
    for (int32_t k = indptr[r]; k < indptr[r + 1]; ++k) {
        const int32_t c = indices[k];
        if (!diag_done && c > r) { diag_only(); diag_done = true; }
        if (c == r) {
            double v = norm_value<T, KIND>(di, values[k], di, true);
            if (KIND == LT_NORM_FIRST_ORDER_GCN) v = 1.0 + v;   // eye + (zero products already dropped)
            emit(c, v);
            diag_done = true;
        } else {
            emit(c, norm_value<T, KIND>(di, values[k], dis[c], false));
        }
    }
    if (!diag_done) diag_only();
    if (!PASS) out_indptr[r] = w;   // counts; scanned in place afterwards
}

template <typename T>
static int normalize_impl(const int32_t* indptr, const int32_t* indices, const T* values, int32_t n, int kind,
                          const T* dis, int32_t* out_indptr, int32_t* out_indices, float* out_values,
                          int64_t cap, int64_t* nnz_host, cudaStream_t s) {
    const int th = 128;
    const unsigned bl = (unsigned)ceil_div(n, th);
    if (kind == LT_NORM_FIRST_ORDER_GCN)
        k_normalize<T, LT_NORM_FIRST_ORDER_GCN, 0><<<bl, th, 0, s>>>(indptr, indices, values, n, dis, out_indptr,
                                                                    nullptr, nullptr);
    else
        k_normalize<T, LT_NORM_AUG_NORM_ADJ, 0><<<bl, th, 0, s>>>(indptr, indices, values, n, dis, out_indptr,
                                                                 nullptr, nullptr);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(out_indptr, out_indptr, n, s));
    int32_t nnz = 0;
    LT_CUDA(cudaMemcpyAsync(&nnz, out_indptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    *nnz_host = nnz;
    if (nnz > cap) {
        set_error("lt_adj_normalize: output needs %d entries, capacity %lld", nnz, (long long)cap);
        return LT_ERR_CAPACITY;
    }
    if (kind == LT_NORM_FIRST_ORDER_GCN)
        k_normalize<T, LT_NORM_FIRST_ORDER_GCN, 1><<<bl, th, 0, s>>>(indptr, indices, values, n, dis, out_indptr,
                                                                    out_indices, out_values);
    else
        k_normalize<T, LT_NORM_AUG_NORM_ADJ, 1><<<bl, th, 0, s>>>(indptr, indices, values, n, dis, out_indptr,
                                                                 out_indices, out_values);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

}  // namespace lt

using namespace lt;

namespace lt {
int csr_from_edges(const int64_t* edges, int64_t m, int32_t n, int binarize, int32_t* indptr, int32_t* indices,
                   float* values, int64_t cap, int64_t* nnz_host, cudaStream_t s) {
    LT_CHECK_ARG((edges || m == 0) && indptr && indices && values && nnz_host && m >= 0 && n > 0, "lt_csr_from_edges");
    LT_CHECK_ARG(2 * m < ((int64_t)1 << 31), "lt_csr_from_edges: 2m must be < 2^31");
    const int64_t nk = 2 * m;
    if (nk == 0) {
        k_fill_i32<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(indptr, n + 1, 0);
        LT_LAUNCH_CHECK();
        *nnz_host = 0;
        return LT_OK;
    }
    Scratch scratch(s);
    uint64_t* keys; int32_t *head, *starts; int* err;
    LT_TRY(scratch.alloc(&keys, (size_t)nk));
    LT_TRY(scratch.alloc(&head, (size_t)nk + 1));
    LT_TRY(scratch.alloc(&starts, (size_t)nk));
    LT_TRY(scratch.alloc(&err, 1));
    LT_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
    k_edge_keys<<<(unsigned)ceil_div(m, 256), 256, 0, s>>>(edges, m, n, keys, err);
    LT_LAUNCH_CHECK();
    const int cb = bits_for((uint64_t)(n - 1));
    BitRange ranges[2] = {{0, cb}, {32, 32 + cb}};
    LT_TRY(radix_sort_u64(keys, nullptr, nk, ranges, 2, s));
    k_heads<<<(unsigned)ceil_div(nk, 256), 256, 0, s>>>(keys, nk, head);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(head, head, nk, s));
    int32_t nnz = 0; int herr = 0;
    LT_CUDA(cudaMemcpyAsync(&nnz, head + nk, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    *nnz_host = nnz;
    if (herr) { set_error("lt_csr_from_edges: node id outside [0,%d)", n); return LT_ERR_DATA; }
    if (nnz > cap) {
        set_error("lt_csr_from_edges: %d entries exceed capacity %lld", nnz, (long long)cap);
        return LT_ERR_CAPACITY;
    }
    k_unique_emit<<<(unsigned)ceil_div(nk, 256), 256, 0, s>>>(keys, head, nk, n, indices, starts, indptr, cap);
    LT_LAUNCH_CHECK();
    k_run_values<<<(unsigned)ceil_div(nnz, 256), 256, 0, s>>>(starts, nnz, nk, binarize, values);
    LT_LAUNCH_CHECK();
    return LT_OK;
}
}  // namespace lt

extern "C" int lt_csr_from_edges(const int64_t* edges, int64_t m, int32_t n, int binarize, int32_t* indptr,
                                 int32_t* indices, float* values, int64_t cap, int64_t* nnz_host,
                                 lt_stream_t stream) {
    return lt::csr_from_edges(edges, m, n, binarize, indptr, indices, values, cap, nnz_host, (cudaStream_t)stream);
}

extern "C" int lt_coo_to_csr(const int64_t* row, const int64_t* col, int64_t nnz, int32_t n, int32_t* indptr,
                             int32_t* indices, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && nnz >= 0 && n > 0 && nnz < ((int64_t)1 << 31), "lt_coo_to_csr");
    cudaStream_t s = (cudaStream_t)stream;
    k_coo_rowptr<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(row, nnz, n, indptr);
    LT_LAUNCH_CHECK();
    if (nnz) {
        k_narrow_cols<<<(unsigned)ceil_div(nnz, 256), 256, 0, s>>>(col, nnz, indices);
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

extern "C" int lt_csr_check(const int32_t* indptr, const int32_t* indices, int32_t n, int64_t nnz, int* ok_host,
                            lt_stream_t stream) {
    LT_CHECK_ARG(indptr && ok_host && n > 0, "lt_csr_check");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    int* bad;
    LT_TRY(scratch.alloc(&bad, 1));
    LT_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    k_csr_check<<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(indptr, indices, n, nnz, bad);
    LT_LAUNCH_CHECK();
    int h = 0;
    LT_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    *ok_host = !h;
    return LT_OK;
}

extern "C" int lt_csr_row_sums(const int32_t* indptr, const void* values, int32_t n, int prec, void* row_sums,
                               lt_stream_t stream) {
    LT_CHECK_ARG(indptr && values && row_sums && n > 0, "lt_csr_row_sums");
    cudaStream_t s = (cudaStream_t)stream;
    if (prec == LT_PREC_F32)
        k_row_sums<float><<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(indptr, (const float*)values, n, (float*)row_sums);
    else
        k_row_sums<double><<<(unsigned)ceil_div(n, 128), 128, 0, s>>>(indptr, (const double*)values, n,
                                                                     (double*)row_sums);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

extern "C" int lt_adj_normalize(const int32_t* indptr, const int32_t* indices, const void* values, int32_t n,
                                int kind, int prec, const void* d_inv_sqrt, int32_t* out_indptr,
                                int32_t* out_indices, float* out_values, int64_t cap, int64_t* nnz_host,
                                lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && d_inv_sqrt && out_indptr && out_indices && out_values && nnz_host,
                 "lt_adj_normalize: null pointer");
    LT_CHECK_ARG(kind == LT_NORM_FIRST_ORDER_GCN || kind == LT_NORM_AUG_NORM_ADJ, "lt_adj_normalize: kind");
    LT_CHECK_ARG(!(kind == LT_NORM_AUG_NORM_ADJ && prec == LT_PREC_F32),
                 "AugNormAdj is always float64 arithmetic (sp.eye promotes, utils/load.py:563)");
    cudaStream_t s = (cudaStream_t)stream;
    if (prec == LT_PREC_F32)
        return normalize_impl<float>(indptr, indices, (const float*)values, n, kind, (const float*)d_inv_sqrt,
                                     out_indptr, out_indices, out_values, cap, nnz_host, s);
    return normalize_impl<double>(indptr, indices, (const double*)values, n, kind, (const double*)d_inv_sqrt,
                                  out_indptr, out_indices, out_values, cap, nnz_host, s);
}
