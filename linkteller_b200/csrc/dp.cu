// DP adjacency perturbation applied on the device from host-drawn noise.
//
//  EdgeRand  (Worker.perturb_adj_discrete, worker.py:213-278; construct_sparse_mat :178-203)
//     host : numpy legacy stream — B = binomial(1,s,(N,N)); hits = where(B) row-major; coin =
//            binomial(1,½,len(hits)); only cells r<c are used.
//     here : adj + add − minus as ONE sort-and-merge of (cell, ±1) records with adj's entries:
//            equal cells are summed, exact zeros are dropped (scipy's binop), then −1→0 and 2→1
//            (worker.py:266-267) — so a 'minus' on a non-edge stays as an explicit zero.
//  Lapgraph  (Worker.perturb_adj_continuous, worker.py:281-335; get_noise utils/load.py:27-39)
//     host : Laplace N×N matrix and n_edges_keep from the re-seeded stream.
//     here : noise += tril(A,-1) in place (worker.py:298), exact top-n_keep of the strict lower
//            triangle by an 8-pass MSD radix select on order-preserving 64-bit keys (the reference's
//            50-chunk argpartition is an exact top-k too), ordered compaction of the selected cells,
//            symmetric binary CSR through the a1 build (mat + mat.T, worker.py:334-335).
// All integer / byte work, HBM-bound; results are bit-exact with the reference.
#include "common.cuh"

#include <string.h>

namespace lt {

int csr_from_edges(const int64_t* edges, int64_t m, int32_t n, int binarize, int32_t* indptr, int32_t* indices,
                   float* values, int64_t cap, int64_t* nnz_host, cudaStream_t s);

// ------------------------------------------------------------------------------------ EdgeRand
// records: [0,nnz) adj entries, then two per hit ((r,c) and (c,r)) carrying +1 / −1
__global__ void k_er_adj_records(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                 const double* __restrict__ values, int32_t n, uint64_t* __restrict__ keys,
                                 uint64_t* __restrict__ vals) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    for (int32_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
        keys[k] = ((uint64_t)r << 32) | (uint32_t)indices[k];
        vals[k] = (uint64_t)__double_as_longlong(values[k]);
    }
}

__global__ void k_er_hit_records(const int64_t* __restrict__ hits, const uint8_t* __restrict__ coin, int64_t n_hits,
                                 int32_t n, int64_t base, uint64_t* __restrict__ keys, uint64_t* __restrict__ vals,
                                 int* __restrict__ err) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_hits) return;
    const int64_t cell = hits[i];
    int64_t r = cell / n, c = cell % n;
    if (cell < 0 || r >= n || r >= c) { atomicOr(err, 1); r = 0; c = 1; }
    if (i > 0 && hits[i - 1] >= cell) atomicOr(err, 2);
    const uint64_t v = (uint64_t)__double_as_longlong(coin[i] ? 1.0 : -1.0);
    keys[base + 2 * i] = ((uint64_t)r << 32) | (uint64_t)c;
    keys[base + 2 * i + 1] = ((uint64_t)c << 32) | (uint64_t)r;
    vals[base + 2 * i] = v;
    vals[base + 2 * i + 1] = v;
}

__global__ void k_er_heads(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// one thread per run start: sum the run (adj entry first — the sort is stable), apply the rules
__global__ void k_er_runs(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, int64_t n,
                          const int32_t* __restrict__ pos /* exclusive scan of head */, int32_t* __restrict__ keep,
                          uint64_t* __restrict__ run_key, double* __restrict__ run_val) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (pos[i + 1] == pos[i]) return;
    const uint64_t k = keys[i];
    double sum = __longlong_as_double((long long)vals[i]);
    for (int64_t q = i + 1; q < n && keys[q] == k; ++q) sum += __longlong_as_double((long long)vals[q]);
    const int32_t p = pos[i];
    keep[p] = sum != 0.0 ? 1 : 0;
    if (sum == -1.0) sum = 0.0;
    if (sum == 2.0) sum = 1.0;
    run_key[p] = k;
    run_val[p] = sum;
}

__global__ void k_er_emit(const uint64_t* __restrict__ run_key, const double* __restrict__ run_val,
                          const int32_t* __restrict__ kpos /* exclusive scan of keep */, int64_t n_runs,
                          uint64_t* __restrict__ out_key, int32_t* __restrict__ out_indices,
                          double* __restrict__ out_values, int64_t cap) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_runs) return;
    if (kpos[p + 1] == kpos[p]) return;
    const int32_t w = kpos[p];
    if (w >= cap) return;
    out_key[w] = run_key[p];
    out_indices[w] = (int32_t)(run_key[p] & 0xffffffffu);
    out_values[w] = run_val[p];
}

__global__ void k_rowptr_from_keys(const uint64_t* __restrict__ keys, int64_t nnz, int32_t n,
                                   int32_t* __restrict__ indptr) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n) return;
    indptr[r] = (int32_t)lower_bound(keys, 0, nnz, (uint64_t)r << 32);
}

// ------------------------------------------------------------------------------------ Lapgraph
__device__ __forceinline__ uint64_t f64_key(double d) {   // order-preserving: larger double ⇒ larger key
    const uint64_t u = (uint64_t)__double_as_longlong(d);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

// noise[r,c] += A[r,c] for the stored entries below the diagonal (tril(adj,-1) + noise)
__global__ void k_lap_add_tril(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                               const double* __restrict__ values, int32_t n, double* __restrict__ noise) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    for (int32_t k = indptr[r] + lane; k < indptr[r + 1]; k += 32) {
        const int32_t c = indices[k];
        if (c < r) noise[r * n + c] += values[k];
    }
}

struct SelectState {
    unsigned long long prefix, mask, k_rem;
    unsigned long long hist[256];
};

// histogram of byte `shift/8` over the strict-lower-triangle cells whose decided bytes match
__global__ void __launch_bounds__(256)
k_lap_hist(const double* __restrict__ noise, int32_t n, int shift, SelectState* __restrict__ st) {
    __shared__ unsigned int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = st->prefix, mask = st->mask;
    for (int64_t r = blockIdx.x; r < n; r += gridDim.x) {
        const double* row = noise + r * n;
        for (int64_t c = threadIdx.x; c < r; c += blockDim.x) {
            const uint64_t k = f64_key(row[c]);
            if ((k & mask) == prefix) atomicAdd(&h[(k >> shift) & 0xFF], 1u);
        }
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], (unsigned long long)h[threadIdx.x]);
}

// pick the byte whose bucket contains the k_rem-th largest remaining cell; clear the histogram
__global__ void k_lap_pick(SelectState* st, int shift) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    unsigned long long cum = 0, k = st->k_rem;
    int b = 255;
    for (; b > 0; --b) {
        if (cum + st->hist[b] >= k) break;
        cum += st->hist[b];
    }
    st->k_rem = k - cum;
    st->prefix |= (unsigned long long)b << shift;
    st->mask |= 0xFFull << shift;
    for (int i = 0; i < 256; ++i) st->hist[i] = 0;
}

// warp per row: number of cells above / equal to the threshold key
__global__ void k_lap_count(const double* __restrict__ noise, int32_t n, const SelectState* __restrict__ st,
                            int32_t* __restrict__ cnt_gt, int32_t* __restrict__ cnt_eq) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint64_t thr = st->prefix;
    const double* row = noise + r * n;
    int g = 0, e = 0;
    for (int64_t c = lane; c < r; c += 32) {
        const uint64_t k = f64_key(row[c]);
        g += k > thr;
        e += k == thr;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { g += __shfl_xor_sync(LT_FULL, g, o); e += __shfl_xor_sync(LT_FULL, e, o); }
    if (lane == 0) { cnt_gt[r] = g; cnt_eq[r] = e; }
}

// warp per row, cells in column order: emit (r,c) of every cell above the threshold and of the first
// k_rem cells equal to it (row-major rank) — deterministic.
__global__ void k_lap_emit(const double* __restrict__ noise, int32_t n, const SelectState* __restrict__ st,
                           const int32_t* __restrict__ off_gt, const int32_t* __restrict__ off_eq,
                           int64_t* __restrict__ edges) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const uint64_t thr = st->prefix;
    const long long k_rem = (long long)st->k_rem;
    const long long n_gt = off_gt[n];
    const double* row = noise + r * n;
    long long pg = off_gt[r], pe = off_eq[r];
    const unsigned below = (1u << lane) - 1u;
    for (int64_t c0 = 0; c0 < r; c0 += 32) {
        const int64_t c = c0 + lane;
        uint64_t k = 0;
        if (c < r) k = f64_key(row[c]);
        const bool g = c < r && k > thr, e = c < r && k == thr;
        const unsigned mg = __ballot_sync(LT_FULL, g), me = __ballot_sync(LT_FULL, e);
        if (g) {
            const long long p = pg + __popc(mg & below);
            edges[2 * p] = r;
            edges[2 * p + 1] = c;
        }
        if (e) {
            const long long q = pe + __popc(me & below);
            if (q < k_rem) {
                edges[2 * (n_gt + q)] = r;
                edges[2 * (n_gt + q) + 1] = c;
            }
        }
        pg += __popc(mg);
        pe += __popc(me);
    }
}

}  // namespace lt

using namespace lt;

extern "C" int lt_dp_edgerand_apply(const int32_t* indptr, const int32_t* indices, const double* values, int32_t n,
                                    const int64_t* hits, const uint8_t* coin, int64_t n_hits, int32_t* out_indptr,
                                    int32_t* out_indices, double* out_values, int64_t cap, int64_t* nnz_host,
                                    lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && out_indptr && out_indices && out_values && nnz_host && n > 0 &&
                     n_hits >= 0 && (n_hits == 0 || (hits && coin)), "lt_dp_edgerand_apply");
    cudaStream_t s = (cudaStream_t)stream;
    int32_t nnz = 0;
    LT_CUDA(cudaMemcpyAsync(&nnz, indptr + n, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    const int64_t total = (int64_t)nnz + 2 * n_hits;
    LT_CHECK_ARG(total < ((int64_t)1 << 31), "lt_dp_edgerand_apply: too many records");
    Scratch scratch(s);
    uint64_t *keys, *vals, *run_key, *out_key;
    double* run_val;
    int32_t *head, *keep;
    int* err;
    LT_TRY(scratch.alloc(&keys, (size_t)total));
    LT_TRY(scratch.alloc(&vals, (size_t)total));
    LT_TRY(scratch.alloc(&head, (size_t)total + 1));
    LT_TRY(scratch.alloc(&keep, (size_t)total + 1));
    LT_TRY(scratch.alloc(&run_key, (size_t)total));
    LT_TRY(scratch.alloc(&run_val, (size_t)total));
    LT_TRY(scratch.alloc(&out_key, (size_t)total));
    LT_TRY(scratch.alloc(&err, 1));
    LT_CUDA(cudaMemsetAsync(err, 0, 4, s));
    if (total == 0) {
        LT_CUDA(cudaMemsetAsync(out_indptr, 0, ((size_t)n + 1) * 4, s));
        *nnz_host = 0;
        return LT_OK;
    }
    k_er_adj_records<<<(unsigned)ceil_div((int64_t)n * 32, 256), 256, 0, s>>>(indptr, indices, values, n, keys, vals);
    LT_LAUNCH_CHECK();
    if (n_hits) {
        k_er_hit_records<<<(unsigned)ceil_div(n_hits, 256), 256, 0, s>>>(hits, coin, n_hits, n, nnz, keys, vals, err);
        LT_LAUNCH_CHECK();
    }
    const int cb = bits_for((uint64_t)(n - 1));
    BitRange ranges[2] = {{0, cb}, {32, 32 + cb}};
    LT_TRY(radix_sort_u64(keys, vals, total, ranges, 2, s));
    k_er_heads<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(keys, total, head);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(head, head, total, s));
    LT_CUDA(cudaMemsetAsync(keep, 0, ((size_t)total + 1) * 4, s));
    k_er_runs<<<(unsigned)ceil_div(total, 256), 256, 0, s>>>(keys, vals, total, head, keep, run_key, run_val);
    LT_LAUNCH_CHECK();
    int32_t n_runs = 0;
    int herr = 0;
    LT_CUDA(cudaMemcpyAsync(&n_runs, head + total, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (herr & 1) { set_error("lt_dp_edgerand_apply: hit cell outside the strict upper triangle"); return LT_ERR_DATA; }
    if (herr & 2) { set_error("lt_dp_edgerand_apply: hits must be strictly ascending"); return LT_ERR_DATA; }
    LT_TRY(exclusive_scan_i32(keep, keep, n_runs, s));
    int32_t out_nnz = 0;
    LT_CUDA(cudaMemcpyAsync(&out_nnz, keep + n_runs, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    *nnz_host = out_nnz;
    if (out_nnz > cap) {
        set_error("lt_dp_edgerand_apply: %d entries exceed capacity %lld", out_nnz, (long long)cap);
        return LT_ERR_CAPACITY;
    }
    k_er_emit<<<(unsigned)ceil_div(n_runs, 256), 256, 0, s>>>(run_key, run_val, keep, n_runs, out_key, out_indices,
                                                             out_values, cap);
    LT_LAUNCH_CHECK();
    k_rowptr_from_keys<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(out_key, out_nnz, n, out_indptr);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

extern "C" int lt_dp_lapgraph_select(const int32_t* indptr, const int32_t* indices, const double* values, int32_t n,
                                     double* noise, int64_t n_keep, int32_t* out_indptr, int32_t* out_indices,
                                     float* out_values, int64_t cap, int64_t* nnz_host, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && noise && out_indptr && out_indices && out_values && nnz_host && n > 1,
                 "lt_dp_lapgraph_select");
    const int64_t cells = (int64_t)n * (n - 1) / 2;
    LT_CHECK_ARG(n_keep > 0 && n_keep <= cells && 2 * n_keep < ((int64_t)1 << 31), "lt_dp_lapgraph_select: n_keep");
    LT_CHECK_ARG(cap >= 2 * n_keep, "lt_dp_lapgraph_select: capacity must be >= 2*n_keep");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    SelectState* st;
    int32_t *cnt_gt, *cnt_eq;
    int64_t* edges;
    LT_TRY(scratch.alloc(&st, 1));
    LT_TRY(scratch.alloc(&cnt_gt, (size_t)n + 1));
    LT_TRY(scratch.alloc(&cnt_eq, (size_t)n + 1));
    LT_TRY(scratch.alloc(&edges, (size_t)2 * n_keep));
    k_lap_add_tril<<<(unsigned)ceil_div((int64_t)n * 32, 256), 256, 0, s>>>(indptr, indices, values, n, noise);
    LT_LAUNCH_CHECK();
    SelectState h0;
    memset(&h0, 0, sizeof(h0));
    h0.k_rem = (unsigned long long)n_keep;
    LT_CUDA(cudaMemcpyAsync(st, &h0, sizeof(h0), cudaMemcpyHostToDevice, s));
    LT_CUDA(cudaStreamSynchronize(s));   // h0 lives on this stack frame
    const unsigned grid = (unsigned)min((int64_t)n, (int64_t)num_sms() * 8);
    for (int shift = 56; shift >= 0; shift -= 8) {
        k_lap_hist<<<grid, 256, 0, s>>>(noise, n, shift, st);
        LT_LAUNCH_CHECK();
        k_lap_pick<<<1, 32, 0, s>>>(st, shift);
        LT_LAUNCH_CHECK();
    }
    SelectState h1;
    LT_CUDA(cudaMemcpyAsync(&h1, st, sizeof(h1), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    {   // the reference ranks the whole dense N×N array, whose other cells are exactly 0: the two agree
        // only while the cut is above 0 (always the case for real ε: half the noise cells are positive)
        const unsigned long long k = h1.prefix;
        const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        double thr;
        memcpy(&thr, &u, 8);
        if (!(thr > 0.0)) {
            set_error("lt_dp_lapgraph_select: the n_keep-th largest cell is %g <= 0 (n_keep exceeds the positive cells)", thr);
            return LT_ERR_DATA;
        }
    }
    const unsigned blw = (unsigned)ceil_div((int64_t)n * 32, 256);
    k_lap_count<<<blw, 256, 0, s>>>(noise, n, st, cnt_gt, cnt_eq);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(cnt_gt, cnt_gt, n, s));
    LT_TRY(exclusive_scan_i32(cnt_eq, cnt_eq, n, s));
    k_lap_emit<<<blw, 256, 0, s>>>(noise, n, st, cnt_gt, cnt_eq, edges);
    LT_LAUNCH_CHECK();
    return csr_from_edges(edges, n_keep, n, 1, out_indptr, out_indices, out_values, cap, nnz_host, s);
}
