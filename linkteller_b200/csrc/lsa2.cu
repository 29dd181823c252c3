// LSA2 baselines — Attacker.baseline_attack (attacker.py:287-334): centred cosine similarity of the
// posteriors (LSA2-post: softmax, or sigmoid for 'ppi', attacker.py:296-299) or raw attributes
// (LSA2-attr, :301) of every pair of test nodes.  The reference evaluates T(T-1)/2 tiny torch
// expressions in a Python loop; here: gather+transform rows, column mean, centre + row norms, one
// T x d x T Gram product (lt::gemm_f32, B transposed) and a scaling pass.
#include "common.cuh"

namespace lt {

int gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
             const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int algo,
             cudaStream_t s);

// one warp per test node: sub[i,:] = transform(M[test_nodes[i], :d])
__global__ void __launch_bounds__(256)
k_lsa2_rows(const float* __restrict__ M, int64_t ldm, int32_t d, const int32_t* __restrict__ test_nodes, int32_t T,
            int transform, float* __restrict__ sub) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= T) return;
    const float* src = M + (int64_t)(test_nodes ? test_nodes[i] : (int32_t)i) * ldm;   // null list: all rows in order
    float* dst = sub + i * d;
    if (transform == 1) {            // softmax(dim=1): exp(x − max) / Σ
        float mx = -INFINITY;
        for (int k = lane; k < d; k += 32) mx = fmaxf(mx, src[k]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(LT_FULL, mx, o));
        float sum = 0.f;
        for (int k = lane; k < d; k += 32) sum += expf(src[k] - mx);
        sum = warp_sum(sum);
        for (int k = lane; k < d; k += 32) dst[k] = expf(src[k] - mx) / sum;
    } else if (transform == 2) {     // sigmoid
        for (int k = lane; k < d; k += 32) dst[k] = 1.0f / (1.0f + expf(-src[k]));
    } else {
        for (int k = lane; k < d; k += 32) dst[k] = src[k];
    }
}

// mean over the T rows: one thread per column, rows in order (deterministic; coalesced across columns)
__global__ void k_lsa2_mean(const float* __restrict__ sub, int32_t T, int32_t d, float* __restrict__ mean) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d) return;
    float s = 0.f;
    for (int i = 0; i < T; ++i) s += sub[(int64_t)i * d + k];
    mean[k] = s / (float)T;
}

// one warp per row: centre in place, row 2-norm
__global__ void __launch_bounds__(256)
k_lsa2_centre(float* __restrict__ sub, int32_t T, int32_t d, const float* __restrict__ mean, float* __restrict__ nrm) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= T) return;
    float* row = sub + i * d;
    float ss = 0.f;
    for (int k = lane; k < d; k += 32) {
        const float v = row[k] - mean[k];
        row[k] = v;
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) nrm[i] = sqrtf(ss);
}

// mean over many rows in two deterministic levels: part[b,:] = Σ of rows b·chunk … (one thread per column)
__global__ void k_lsa2_partial(const float* __restrict__ sub, int32_t T, int32_t d, int32_t chunk, float* __restrict__ part) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (k >= d) return;
    const int i0 = b * chunk, i1 = min(T, i0 + chunk);
    float s = 0.f;
    for (int i = i0; i < i1; ++i) s += sub[(int64_t)i * d + k];
    part[(int64_t)b * d + k] = s;
}
__global__ void k_lsa2_mean_of_parts(const float* __restrict__ part, int32_t nb, int32_t d, int32_t T, float* __restrict__ mean) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= d) return;
    float s = 0.f;
    for (int b = 0; b < nb; ++b) s += part[(int64_t)b * d + k];
    mean[k] = s / (float)T;
}

// one warp per listed pair: <c_u, c_v> / ‖c_u‖ / ‖c_v‖ on the centred rows (attacker.py:364-368)
__global__ void __launch_bounds__(256)
k_lsa2_pairs(const float* __restrict__ cen, int32_t d, const float* __restrict__ nrm, const int32_t* __restrict__ pu,
             const int32_t* __restrict__ pv, int64_t K, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= K) return;
    const int32_t u = pu[k], v = pv[k];
    const float* a = cen + (int64_t)u * d;
    const float* b = cen + (int64_t)v * d;
    float s = 0.f;
    for (int c = lane; c < d; c += 32) s = fmaf(a[c], b[c], s);
    s = warp_sum(s);
    if (lane == 0) out[k] = (s / nrm[u]) / nrm[v];
}

// dist[i][j] = gram/‖i‖/‖j‖ for i<j, 0 elsewhere (attacker.py:316: dot / norm / norm, in that order)
__global__ void k_lsa2_scale(float* __restrict__ dist, int64_t ldd, int32_t T, const float* __restrict__ nrm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= T) return;
    float* p = dist + (int64_t)i * ldd + j;
    *p = (i < j) ? (*p / nrm[i]) / nrm[j] : 0.f;
}

}  // namespace lt

using namespace lt;

extern "C" int lt_lsa2_corr(const float* M, int64_t ldm, int32_t d, const int32_t* test_nodes, int32_t T,
                            int transform, float* dist, int64_t ldd, lt_stream_t stream) {
    LT_CHECK_ARG(M && test_nodes && dist && d > 0 && T > 0 && ldm >= d && ldd >= T, "lt_lsa2_corr");
    LT_CHECK_ARG(transform >= 0 && transform <= 2, "lt_lsa2_corr: transform");
    LT_CHECK_ARG(T <= 65535, "lt_lsa2_corr: T must be <= 65535");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    float *sub, *mean, *nrm;
    LT_TRY(scratch.alloc(&sub, (size_t)T * d));
    LT_TRY(scratch.alloc(&mean, (size_t)d));
    LT_TRY(scratch.alloc(&nrm, (size_t)T));
    const unsigned blT = (unsigned)ceil_div((int64_t)T * 32, 256);
    k_lsa2_rows<<<blT, 256, 0, s>>>(M, ldm, d, test_nodes, T, transform, sub);
    LT_LAUNCH_CHECK();
    k_lsa2_mean<<<(unsigned)ceil_div(d, 128), 128, 0, s>>>(sub, T, d, mean);
    LT_LAUNCH_CHECK();
    k_lsa2_centre<<<blT, 256, 0, s>>>(sub, T, d, mean, nrm);
    LT_LAUNCH_CHECK();
    LT_TRY(gemm_f32(sub, d, nullptr, sub, d, 1, nullptr, 0, dist, ldd, T, T, d, 1, s));
    k_lsa2_scale<<<dim3((unsigned)ceil_div(T, 256), (unsigned)T), 256, 0, s>>>(dist, ldd, T, nrm);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

extern "C" int lt_lsa2_pairs(const float* M, int64_t ldm, int32_t n_rows, int32_t d, int transform, const int32_t* pu,
                             const int32_t* pv, int64_t n_pairs, float* out, lt_stream_t stream) {
    LT_CHECK_ARG(M && pu && pv && out && d > 0 && n_rows > 0 && ldm >= d && n_pairs >= 0, "lt_lsa2_pairs");
    LT_CHECK_ARG(transform >= 0 && transform <= 2, "lt_lsa2_pairs: transform");
    if (n_pairs == 0) return LT_OK;
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    const int32_t chunk = 1024, nb = (int32_t)ceil_div(n_rows, chunk);
    float *sub, *part, *mean, *nrm;
    LT_TRY(scratch.alloc(&sub, (size_t)n_rows * d));
    LT_TRY(scratch.alloc(&part, (size_t)nb * d));
    LT_TRY(scratch.alloc(&mean, (size_t)d));
    LT_TRY(scratch.alloc(&nrm, (size_t)n_rows));
    const unsigned blT = (unsigned)ceil_div((int64_t)n_rows * 32, 256);
    k_lsa2_rows<<<blT, 256, 0, s>>>(M, ldm, d, nullptr, n_rows, transform, sub);
    LT_LAUNCH_CHECK();
    k_lsa2_partial<<<dim3((unsigned)ceil_div(d, 128), (unsigned)nb), 128, 0, s>>>(sub, n_rows, d, chunk, part);
    LT_LAUNCH_CHECK();
    k_lsa2_mean_of_parts<<<(unsigned)ceil_div(d, 128), 128, 0, s>>>(part, nb, d, n_rows, mean);
    LT_LAUNCH_CHECK();
    k_lsa2_centre<<<blT, 256, 0, s>>>(sub, n_rows, d, mean, nrm);
    LT_LAUNCH_CHECK();
    k_lsa2_pairs<<<(unsigned)ceil_div(n_pairs * 32, 256), 256, 0, s>>>(sub, d, nrm, pu, pv, n_pairs, out);
    LT_LAUNCH_CHECK();
    return LT_OK;
}
