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
    const float* src = M + (int64_t)test_nodes[i] * ldm;
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
