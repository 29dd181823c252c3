// Pair labelling, pair scores and ranking metrics — integer / sort work,
// HBM-bound.
//
//   lt_pair_labels : _get_edge_sets_among_nodes (utils/load.py:304-326) — for every pair i<j in
//                    sample order, is test_nodes[j] stored in row test_nodes[i] of adj_ori?
//   lt_pair_scores : attacker.py:233-245 — score(i<j) = influence[j][i]  (tiled transpose gather)
//   lt_auc_ap      : attacker.py:382-386 — sklearn roc_curve/auc (one threshold per distinct score
//                    = tie-aware Mann-Whitney) and average_precision_score
//   lt_topk_tp     : attack_stats_all.py:108-116 — true positives among the global top-n_pos
//                    scores, as (min,max) over the ways argpartition may break ties at the cut
//
// Pair index of (i,j), i<j, row-major:  p = i·(2T−i−1)/2 + (j−i−1).
#include "common.cuh"

namespace lt {

__host__ __device__ __forceinline__ int64_t pair_row_offset(int64_t i, int64_t T) { return i * (2 * T - i - 1) / 2; }

// ------------------------------------------------------------------------------------ labels
// grid.y = i, threads over j.  Row of u = test_nodes[i] is searched for v = test_nodes[j].
__global__ void __launch_bounds__(256)
k_pair_labels(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, int32_t n,
              const int32_t* __restrict__ test_nodes, int32_t T, uint8_t* __restrict__ labels,
              unsigned long long* __restrict__ n_pos, int* __restrict__ err) {
    const int32_t i = blockIdx.y;
    const int32_t j = i + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    int hit = 0;
    if (j < T) {
        const int32_t u = test_nodes[i], v = test_nodes[j];
        if (u < 0 || u >= n || v < 0 || v >= n) {
            *err = 1;
        } else {
            const int64_t a = indptr[u], b = indptr[u + 1];
            const int64_t p = lower_bound(indices, a, b, v);
            hit = (p < b && indices[p] == v) ? 1 : 0;
            labels[pair_row_offset(i, T) + (j - i - 1)] = (uint8_t)hit;
        }
    }
    const unsigned m = __ballot_sync(LT_FULL, hit);
    __shared__ int wsum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int s = 0;
        for (int w = 0; w < 8; ++w) s += wsum[w];
        if (s) atomicAdd(n_pos, (unsigned long long)s);
    }
}

// ------------------------------------------------------------------------------------ scores
// 32x32 tiles: read influence[j, i] with i fastest (coalesced), write scores with j fastest.
__global__ void __launch_bounds__(256)
k_pair_scores(const float* __restrict__ infl, int64_t ldi, int32_t T, float* __restrict__ scores) {
    __shared__ float tile[32][33];
    const int ib = blockIdx.x, jb = blockIdx.y;
    if (ib > jb) return;   // only tiles that contain some i<j
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int j = jb * 32 + r, i = ib * 32 + tx;
        tile[r][tx] = (j < T && i < T) ? infl[(int64_t)j * ldi + i] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int i = ib * 32 + r, j = jb * 32 + tx;
        if (i < j && j < T) scores[pair_row_offset(i, T) + (j - i - 1)] = tile[tx][r];
    }
}

// listed pairs: out[k] = block[(rows[k] − row0)·ld + cols[k]] for the pairs whose row lies in [row0, row0+n_rows)
__global__ void k_pair_gather(const float* __restrict__ block, int64_t ld, int32_t row0, int32_t n_rows,
                              const int32_t* __restrict__ rows, const int32_t* __restrict__ cols, int64_t K,
                              float* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int32_t r = rows[k] - row0;
    if (r >= 0 && r < n_rows) out[k] = block[(int64_t)r * ld + cols[k]];
}

// ------------------------------------------------------------------------------------ ranking
// descending order of float scores = ascending order of ~sortable(bits)
__device__ __forceinline__ uint32_t desc_key(float f) {
    uint32_t u = __float_as_uint(f);
    if (f == 0.f) u = 0;                                  // -0.0 == +0.0 (one tie group, like numpy)
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // ascending sortable
    return ~u;
}

__global__ void k_rank_keys(const float* __restrict__ scores, const uint8_t* __restrict__ labels, int64_t n,
                            uint64_t* __restrict__ keys, int* __restrict__ err) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const float s = scores[p];
    if (s != s) *err = 1;  // NaN scores make sklearn raise; refuse them too
    keys[p] = ((uint64_t)desc_key(s) << 32) | (uint64_t)(labels[p] ? 1u : 0u);
}

__global__ void k_rank_flags(const uint64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ head,
                             int32_t* __restrict__ lab) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint64_t k = keys[p];
    head[p] = (p == 0 || (keys[p - 1] >> 32) != (k >> 32)) ? 1 : 0;
    lab[p] = (int32_t)(k & 1u);
}

// gid = exclusive scan of head (n+1 entries).  gstart[g] = first position of tie group g; gstart[G] = n
__global__ void k_rank_gstart(const int32_t* __restrict__ gid, int64_t n, int32_t* __restrict__ gstart) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    if (p == n) { gstart[gid[n]] = (int32_t)n; return; }
    if (gid[p + 1] != gid[p]) gstart[gid[p]] = (int32_t)p;
}

// One thread per tie group g.  tp_excl = exclusive scan of labels (n+1 entries).
//   auc term  : (FP_g − FP_{g−1})·(TP_g + TP_{g−1})       (exact in int64; AUC = Σ / (2·P·N))
//   ap  term  : (TP_g − TP_{g−1})·TP_g / (TP_g + FP_g)    (AP = Σ / P)
// Block partial sums go to part_auc / part_ap and are summed in block order afterwards (deterministic).
__global__ void __launch_bounds__(256)
k_rank_terms(const int32_t* __restrict__ gstart, int32_t G, const int32_t* __restrict__ tp_excl,
             long long* __restrict__ part_auc, double* __restrict__ part_ap) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    long long ta = 0;
    double tp = 0.0;
    if (g < G) {
        const int64_t s = gstart[g], e = gstart[g + 1];
        const int64_t tp0 = tp_excl[s], tp1 = tp_excl[e];
        const int64_t fp0 = s - tp0, fp1 = e - tp1;
        ta = (fp1 - fp0) * (tp1 + tp0);
        tp = (double)(tp1 - tp0) * ((double)tp1 / (double)e);
    }
    __shared__ long long sa[256];
    __shared__ double sp[256];
    sa[threadIdx.x] = ta;
    sp[threadIdx.x] = tp;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sp[threadIdx.x] += sp[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { part_auc[blockIdx.x] = sa[0]; part_ap[blockIdx.x] = sp[0]; }
}

__global__ void __launch_bounds__(256)
k_rank_final(const long long* __restrict__ part_auc, const double* __restrict__ part_ap, int32_t nb,
             const int32_t* __restrict__ tp_excl, int64_t n, double* __restrict__ out /* auc, ap, P, N */) {
    __shared__ long long sa[256];
    __shared__ double sp[256];
    long long a = 0;
    double p = 0.0;
    for (int b = threadIdx.x; b < nb; b += 256) { a += part_auc[b]; p += part_ap[b]; }
    sa[threadIdx.x] = a;
    sp[threadIdx.x] = p;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sa[threadIdx.x] += sa[threadIdx.x + o]; sp[threadIdx.x] += sp[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double P = (double)tp_excl[n], N = (double)(n - tp_excl[n]);
        out[0] = (double)sa[0] / (2.0 * P * N);
        out[1] = sp[0] / P;
        out[2] = P;
        out[3] = N;
    }
}

// one thread per query: n_pos = number of pairs predicted positive
__global__ void k_topk_tp(const int32_t* __restrict__ gid, const int32_t* __restrict__ gstart,
                          const int32_t* __restrict__ tp_excl, int64_t n, const int64_t* __restrict__ n_pos,
                          int32_t nq, int64_t* __restrict__ tp_min, int64_t* __restrict__ tp_max) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    int64_t k = n_pos[q];
    if (k <= 0) { tp_min[q] = tp_max[q] = 0; return; }
    if (k > n) k = n;
    const int32_t g = gid[k - 1 + 1] - 1;          // group of the k-th largest score (gid is an exclusive scan)
    const int64_t s = gstart[g], e = gstart[g + 1];
    const int64_t tp_above = tp_excl[s];
    const int64_t need = k - s;
    const int64_t tie_pos = tp_excl[e] - tp_excl[s];
    const int64_t tie_neg = (e - s) - tie_pos;
    const int64_t lo = need - tie_neg;
    tp_min[q] = tp_above + (lo > 0 ? lo : 0);
    tp_max[q] = tp_above + (need < tie_pos ? need : tie_pos);
}

// one thread per tie group: the curve point after consuming the group (sklearn's _binary_clf_curve:
// one threshold per distinct score, descending) — threshold, cumulative TP and FP
__global__ void k_rank_curve(const uint64_t* __restrict__ keys, const int32_t* __restrict__ gstart, int32_t G,
                             const int32_t* __restrict__ tp_excl, float* __restrict__ thr, int64_t* __restrict__ tps,
                             int64_t* __restrict__ fps) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int64_t s = gstart[g], e = gstart[g + 1];
    uint32_t u = ~(uint32_t)(keys[s] >> 32);                       // undo desc_key
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    thr[g] = __uint_as_float(u);
    const int64_t tp = tp_excl[e];
    tps[g] = tp;
    fps[g] = e - tp;
}

struct Ranking {
    uint64_t* keys = nullptr;
    int32_t *gid = nullptr, *gstart = nullptr, *tp_excl = nullptr;
    int32_t G = 0;
};

// sort by score (descending, stable) and build the tie-group structure; synchronises once
static int build_ranking(Scratch& scratch, const float* scores, const uint8_t* labels, int64_t n, Ranking* rk,
                         cudaStream_t s) {
    if (n >= ((int64_t)1 << 31) - 1) { set_error("ranking: %lld pairs exceed 2^31", (long long)n); return LT_ERR_CAPACITY; }
    uint64_t* keys;
    int32_t* lab;
    int* err;
    LT_TRY(scratch.alloc(&keys, (size_t)n));
    rk->keys = keys;
    LT_TRY(scratch.alloc(&rk->gid, (size_t)n + 1));
    LT_TRY(scratch.alloc(&lab, (size_t)n + 1));
    LT_TRY(scratch.alloc(&rk->gstart, (size_t)n + 2));
    LT_TRY(scratch.alloc(&err, 1));
    rk->tp_excl = lab;
    LT_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
    const unsigned bl = (unsigned)ceil_div(n, 256);
    k_rank_keys<<<bl, 256, 0, s>>>(scores, labels, n, keys, err);
    LT_LAUNCH_CHECK();
    BitRange r{32, 64};
    LT_TRY(radix_sort_u64(keys, nullptr, n, &r, 1, s));
    k_rank_flags<<<bl, 256, 0, s>>>(keys, n, rk->gid, lab);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(rk->gid, rk->gid, n, s));
    LT_TRY(exclusive_scan_i32(lab, lab, n, s));
    k_rank_gstart<<<(unsigned)ceil_div(n + 1, 256), 256, 0, s>>>(rk->gid, n, rk->gstart);
    LT_LAUNCH_CHECK();
    int herr = 0;
    LT_CUDA(cudaMemcpyAsync(&rk->G, rk->gid + n, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (herr) { set_error("ranking: NaN score"); return LT_ERR_DATA; }
    return LT_OK;
}

}  // namespace lt

using namespace lt;

extern "C" int lt_pair_labels(const int32_t* indptr, const int32_t* indices, int32_t n, const int32_t* test_nodes,
                              int32_t T, uint8_t* labels, int64_t* n_pos_host, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && test_nodes && labels && n > 0 && T >= 0, "lt_pair_labels");
    cudaStream_t s = (cudaStream_t)stream;
    if (n_pos_host) *n_pos_host = 0;
    if (T < 2) return LT_OK;
    LT_CHECK_ARG(T <= 65535, "lt_pair_labels: T must be <= 65535");
    Scratch scratch(s);
    unsigned long long* cnt;
    int* err;
    LT_TRY(scratch.alloc(&cnt, 1));
    LT_TRY(scratch.alloc(&err, 1));
    LT_CUDA(cudaMemsetAsync(cnt, 0, 8, s));
    LT_CUDA(cudaMemsetAsync(err, 0, 4, s));
    dim3 grid((unsigned)ceil_div(T - 1, 256), (unsigned)(T - 1));
    k_pair_labels<<<grid, 256, 0, s>>>(indptr, indices, n, test_nodes, T, labels, cnt, err);
    LT_LAUNCH_CHECK();
    unsigned long long h = 0;
    int herr = 0;
    LT_CUDA(cudaMemcpyAsync(&h, cnt, 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (herr) { set_error("lt_pair_labels: test node outside [0,%d)", n); return LT_ERR_DATA; }
    if (n_pos_host) *n_pos_host = (int64_t)h;
    return LT_OK;
}

extern "C" int lt_pair_scores(const float* influence, int64_t ldi, int32_t T, float* scores, lt_stream_t stream) {
    LT_CHECK_ARG(influence && scores && T >= 0 && ldi >= T, "lt_pair_scores");
    if (T < 2) return LT_OK;
    const unsigned nb = (unsigned)ceil_div(T, 32);
    LT_CHECK_ARG(nb <= 65535, "lt_pair_scores: T too large");
    k_pair_scores<<<dim3(nb, nb), 256, 0, (cudaStream_t)stream>>>(influence, ldi, T, scores);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

extern "C" int lt_pair_gather(const float* block, int64_t ld, int32_t row0, int32_t n_rows, const int32_t* rows,
                              const int32_t* cols, int64_t n_pairs, float* out, lt_stream_t stream) {
    LT_CHECK_ARG(block && rows && cols && out && ld > 0 && n_rows >= 0 && n_pairs >= 0, "lt_pair_gather");
    if (n_pairs == 0 || n_rows == 0) return LT_OK;
    k_pair_gather<<<(unsigned)ceil_div(n_pairs, 256), 256, 0, (cudaStream_t)stream>>>(block, ld, row0, n_rows, rows, cols,
                                                                                  n_pairs, out);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

extern "C" int lt_auc_ap(const float* scores, const uint8_t* labels, int64_t n_pairs, double* auc_host,
                         double* ap_host, lt_stream_t stream) {
    LT_CHECK_ARG(scores && labels && n_pairs > 0 && auc_host && ap_host, "lt_auc_ap");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    Ranking rk;
    LT_TRY(build_ranking(scratch, scores, labels, n_pairs, &rk, s));
    const int32_t nb = (int32_t)ceil_div(rk.G, 256);
    long long* pa;
    double *pp, *out;
    LT_TRY(scratch.alloc(&pa, (size_t)nb));
    LT_TRY(scratch.alloc(&pp, (size_t)nb));
    LT_TRY(scratch.alloc(&out, 4));
    k_rank_terms<<<nb, 256, 0, s>>>(rk.gstart, rk.G, rk.tp_excl, pa, pp);
    LT_LAUNCH_CHECK();
    k_rank_final<<<1, 256, 0, s>>>(pa, pp, nb, rk.tp_excl, n_pairs, out);
    LT_LAUNCH_CHECK();
    double h[4];
    LT_CUDA(cudaMemcpyAsync(h, out, sizeof(h), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (h[2] == 0.0 || h[3] == 0.0) {
        set_error("lt_auc_ap: only one class present (P=%g, N=%g)", h[2], h[3]);
        return LT_ERR_DATA;
    }
    *auc_host = h[0];
    *ap_host = h[1];
    return LT_OK;
}

extern "C" int lt_topk_tp(const float* scores, const uint8_t* labels, int64_t n_pairs, const int64_t* n_pos_host,
                          int32_t n_queries, int64_t* tp_min_host, int64_t* tp_max_host, lt_stream_t stream) {
    LT_CHECK_ARG(scores && labels && n_pairs > 0 && n_pos_host && tp_min_host && tp_max_host && n_queries > 0 &&
                     n_queries <= 1024, "lt_topk_tp");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    Ranking rk;
    LT_TRY(build_ranking(scratch, scores, labels, n_pairs, &rk, s));
    int64_t *q, *lo, *hi;
    LT_TRY(scratch.alloc(&q, (size_t)n_queries));
    LT_TRY(scratch.alloc(&lo, (size_t)n_queries));
    LT_TRY(scratch.alloc(&hi, (size_t)n_queries));
    LT_CUDA(cudaMemcpyAsync(q, n_pos_host, (size_t)n_queries * 8, cudaMemcpyHostToDevice, s));
    k_topk_tp<<<(unsigned)ceil_div(n_queries, 64), 64, 0, s>>>(rk.gid, rk.gstart, rk.tp_excl, n_pairs, q, n_queries, lo,
                                                              hi);
    LT_LAUNCH_CHECK();
    LT_CUDA(cudaMemcpyAsync(tp_min_host, lo, (size_t)n_queries * 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(tp_max_host, hi, (size_t)n_queries * 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    return LT_OK;
}

extern "C" int lt_rank_curve(const float* scores, const uint8_t* labels, int64_t n_pairs, float* thresholds,
                             int64_t* tps, int64_t* fps, int64_t cap, int64_t* n_points_host, lt_stream_t stream) {
    LT_CHECK_ARG(scores && labels && n_pairs > 0 && n_points_host, "lt_rank_curve");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    Ranking rk;
    LT_TRY(build_ranking(scratch, scores, labels, n_pairs, &rk, s));
    *n_points_host = rk.G;
    if (rk.G > cap || !thresholds || !tps || !fps) {
        if (cap == 0) return LT_OK;   // size query
        set_error("lt_rank_curve: %d curve points exceed capacity %lld", rk.G, (long long)cap);
        return LT_ERR_CAPACITY;
    }
    k_rank_curve<<<(unsigned)ceil_div(rk.G, 256), 256, 0, s>>>(rk.keys, rk.gstart, rk.G, rk.tp_excl, thresholds, tps, fps);
    LT_LAUNCH_CHECK();
    LT_CUDA(cudaStreamSynchronize(s));   // scratch (keys) must outlive the kernel's reads of it: freed stream-ordered anyway
    return LT_OK;
}
