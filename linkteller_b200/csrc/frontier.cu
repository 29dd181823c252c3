// The influence hot path: Attacker.get_gradient_eps_mat (attacker.py:100-108) inside the loop of
// link_prediction_attack_efficient (attacker.py:209-229), re-derived for the GPU.
//
// Perturbing x_v ← (1+δ)x_v changes layer 1 only on N̂(v) = {r : Â[r,v] ≠ 0}:
//     H'[r] = H[r] + δ·Â[r,v]·P[v],  P = X·W1,  H = Â·P + b1.
// With ΔR[r] = (relu(H'[r]) − relu(H[r]))/δ and ΔQ[r] = ΔR[r]·W2 (b2 cancels),
//     ((f(X') − f(X))/δ)[u] = Σ_{r ∈ N̂(v) ∩ N̂(u)} Â[u,r]·ΔQ_v[r],
// and the influence entry is its 2-norm; everything outside two hops is exactly 0, as in the
// reference.  Where no ReLU changes sign ΔR[r,k] = Â[r,v]·P[v,k] exactly, so the fp32 cancellation
// of the reference's finite difference never happens.
//
// Work is organised around three sparse structures built per attack ("plan", integer work only):
//   S : incidences e = (j, r)  for v = test_nodes[lo+j], r ∈ N̂(v)      [CSR by j]  + a_e = Â[r,v]
//   B : observers  (r → (i, w)) for test node i adjacent to r, w = Â[u_i,r]        [CSR by r ∈ R]
//   G : records (j, i, e, w) grouped by output pair (j,i), contributions in neighbour order
// and two numeric kernels:
//   F1 k_frontier_dq     : ΔQ[e,:] for every incidence      (reads one H row per incidence: HBM-bound
//                          on large graphs; a [64 x h]·[h x c] register-blocked product per CTA)
//   F2 k_frontier_reduce : out[j,i] = ‖Σ_k w_k·ΔQ[e_k,:]‖₂  (segmented, deterministic)
#include "gemm_ws.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace lt {

int gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
             const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int algo,
             cudaStream_t s);
int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                 const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                 float* C, int64_t ldc, cudaStream_t s);
int spmm_csr_f32_seg(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                     const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                     float* C, int64_t ldc, const int32_t* seg_off_known, int64_t total_seg_known, cudaStream_t s,
                     const HubPanel* hub = nullptr, float hint_deg = 0.f);

}  // namespace lt

struct lt_plan {
    int32_t n = 0, T = 0, lo = 0, hi = 0, Ts = 0;
    int32_t R = 0;        // |R|
    int64_t E = 0;        // incidences
    int64_t NB = 0;       // observer entries
    int64_t M = 0;        // records
    int64_t nnzR = 0;     // Σ_{r∈R} nnz(Â[r,:])
    int32_t* s_indptr = nullptr;  // [Ts+1]
    int32_t* s_j = nullptr;       // [E] owner row j
    int32_t* s_rho = nullptr;     // [E] compact index of r in R
    float* s_a = nullptr;         // [E] Â[r,v]
    int32_t* rows = nullptr;      // [R] node ids, ascending
    int32_t* b_indptr = nullptr;  // [R+1]
    int32_t* b_i = nullptr;       // [NB]
    float* b_w = nullptr;         // [NB]
    uint64_t* rec_key = nullptr;  // [M] (j<<32 | i), sorted
    uint64_t* rec_val = nullptr;  // [M] (e<<32 | float bits of w)
    int32_t* grp_off = nullptr;   // [M+1] start of every (j,i) group; first n_groups+1 entries valid
    int32_t* n_groups = nullptr;  // device scalar
    int32_t* big = nullptr;       // [M / F2_BIG + 1] groups with more than F2_BIG records
    int32_t* n_big = nullptr;     // device scalar
    int32_t* seg_off = nullptr;   // [R+1] exclusive scan of the SpMM segment counts of the frontier rows
    int64_t total_seg = -1;       // its total (read back with the other sizes: the restricted SpMM needs no sync of its own)
    int32_t n_groups_host = -1;
    bool mark_all = false;        // frontier rows of ALL T test nodes, not of the slice (several ranks sharding the precompute)
    bool listed_only = false;     // listed-pair mode (lt_pair_influence): incidences only, no observer lists / records
    void* owned[24];
    int n_owned = 0;
};

namespace lt {

// ------------------------------------------------------------------------------------ plan kernels
// The frontier row set R = ∪_{v∈slice} N̂(v) is a bitmap over the nodes (n/32 words: 306 KB at n = 2.45 M, against the
// 10 MB int flag / rank arrays of round 1) with an exclusive popcount prefix per word: the compact index of row r is
// pref[r/32] + popc(bits[r/32] & below(r)), evaluated where it is needed instead of being stored for all n nodes.
__device__ __forceinline__ int32_t frontier_rank(const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref, int32_t r) {
    const uint32_t b = bits[r >> 5], m = 1u << (r & 31);
    return (b & m) ? pref[r >> 5] + __popc(b & (m - 1)) : -1;
}

// len[i] = nnz of row test_nodes[i] for all T; slice lens go to s_len.  A second bitmap catches duplicate test nodes.
__global__ void k_test_lens(const int32_t* __restrict__ indptr, const int32_t* __restrict__ test_nodes, int32_t T,
                            int32_t n, int32_t lo, int32_t hi, int32_t* __restrict__ all_len, int32_t* __restrict__ s_len,
                            int* __restrict__ err = nullptr, uint32_t* __restrict__ seen = nullptr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    const int32_t v = test_nodes[i];
    const bool ok = v >= 0 && v < n;
    if (!ok && err) atomicOr(err, 1);                                         // reported with the first host sync
    if (ok && seen && (atomicOr(&seen[v >> 5], 1u << (v & 31)) & (1u << (v & 31)))) atomicOr(err, 2);   // duplicate
    const int32_t len = ok ? indptr[v + 1] - indptr[v] : 0;
    all_len[i] = len;
    if (s_len && i >= lo && i < hi) s_len[i - lo] = len;
}

// The same for the attack's plan, with both exclusive scans done in place by ONE CTA (T is a few thousand: three
// launches on the serial head of every attack otherwise): all_off[T+1] over all test nodes, s_indptr[Ts+1] over the
// slice [lo,hi); sizes[0] = E (Σ slice lengths), sizes[2] = Σ all lengths, sizes[3] |= error bits.
constexpr int TL_THREADS = 1024, TL_ITEMS = 8;
__global__ void __launch_bounds__(TL_THREADS)
k_test_lens_scan(const int32_t* __restrict__ indptr, const int32_t* __restrict__ test_nodes, int32_t T, int32_t n, int32_t lo,
                 int32_t hi, int32_t* __restrict__ all_off, int32_t* __restrict__ s_indptr, int32_t* __restrict__ sizes,
                 uint32_t* __restrict__ seen) {
    __shared__ int32_t wa[32], ws[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t carry_a = 0, carry_s = 0;
    int* err = sizes + 3;
    for (int32_t tile = 0; tile < T; tile += TL_THREADS * TL_ITEMS) {
        const int32_t base = tile + (int32_t)threadIdx.x * TL_ITEMS;
        int32_t v[TL_ITEMS], len[TL_ITEMS];
#pragma unroll
        for (int k = 0; k < TL_ITEMS; ++k) v[k] = base + k < T ? test_nodes[base + k] : -1;
        int32_t sa = 0, ss = 0;
#pragma unroll
        for (int k = 0; k < TL_ITEMS; ++k) {
            const int32_t i = base + k;
            int32_t l = 0;
            if (i < T) {
                const bool ok = v[k] >= 0 && v[k] < n;
                if (!ok) atomicOr(err, 1);
                else {
                    if (atomicOr(&seen[v[k] >> 5], 1u << (v[k] & 31)) & (1u << (v[k] & 31))) atomicOr(err, 2);   // duplicate
                    l = indptr[v[k] + 1] - indptr[v[k]];
                }
            }
            len[k] = l;
            sa += l;
            if (i >= lo && i < hi) ss += l;
        }
        const int32_t ia = warp_inclusive_scan(sa), is = warp_inclusive_scan(ss);
        if (lane == 31) { wa[warp] = ia; ws[warp] = is; }
        __syncthreads();
        if (warp == 0) {
            const int32_t ta = wa[lane], ts = ws[lane];
            const int32_t xa = warp_inclusive_scan(ta), xs = warp_inclusive_scan(ts);
            wa[lane] = xa - ta;                                  // exclusive prefix of the warp
            ws[lane] = xs - ts;
        }
        __syncthreads();
        int32_t pa = carry_a + wa[warp] + ia - sa, ps = carry_s + ws[warp] + is - ss;
#pragma unroll
        for (int k = 0; k < TL_ITEMS; ++k) {
            const int32_t i = base + k;
            if (i < T) {
                all_off[i] = pa;
                if (i >= lo && i < hi) s_indptr[i - lo] = ps;
            }
            pa += len[k];
            if (i >= lo && i < hi) ps += len[k];
        }
        // totals of the tile: the last thread's running values
        __shared__ int32_t tot_a, tot_s;
        if (threadIdx.x == TL_THREADS - 1) { tot_a = pa; tot_s = ps; }
        __syncthreads();
        carry_a = tot_a;
        carry_s = tot_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        all_off[T] = carry_a;
        s_indptr[hi - lo] = carry_s;
        sizes[0] = carry_s;
        sizes[2] = carry_a;
    }
}

// one thread per incidence e = (j, k-th entry of row v_j): set the bit of r.  (A warp per row left the kernel waiting
// for the few test nodes with thousands of neighbours: 58 µs against 8.)
__global__ void k_mark_rows(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const int32_t* __restrict__ test_nodes, int32_t n, int32_t lo, int32_t Ts,
                            const int32_t* __restrict__ s_indptr, uint32_t* __restrict__ bits) {
    const int64_t E = s_indptr[Ts];                       // not known on the host yet: fixed grid, strided
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
        const int32_t j = (int32_t)upper_bound(s_indptr, 0, (int64_t)Ts + 1, (int32_t)e) - 1;
        const int32_t v = test_nodes[lo + j];             // valid: an invalid test node has length 0 and owns no incidence
        const int32_t r = indices[indptr[v] + ((int32_t)e - s_indptr[j])];
        const uint32_t m = 1u << (r & 31);
        if (!(bits[r >> 5] & m)) atomicOr(&bits[r >> 5], m);  // hubs are hit by many rows: test first
    }
}

static int mark_rows(const int32_t* indptr, const int32_t* indices, const int32_t* test_nodes, int32_t n, int32_t lo, int32_t Ts,
                     const int32_t* s_indptr, uint32_t* bits, cudaStream_t s) {
    k_mark_rows<<<(unsigned)num_sms() * 16, 256, 0, s>>>(indptr, indices, test_nodes, n, lo, Ts, s_indptr, bits);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

// thread per word: rows[pref + k] = node id of its k-th set bit; nseg = SpMM segments of that row
__global__ void k_compact_rows(const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref, int32_t n_words,
                               const int32_t* __restrict__ indptr, int32_t* __restrict__ rows, int32_t* __restrict__ nseg) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    uint32_t b = bits[w];
    int32_t p = pref[w];
    while (b) {
        const int32_t r = (int32_t)(w * 32) + __ffs(b) - 1;
        b &= b - 1;
        rows[p] = r;
        const int32_t len = indptr[r + 1] - indptr[r];
        nseg[p] = len <= SPMM_SEG ? 1 : (len + SPMM_SEG - 1) / SPMM_SEG;
        ++p;
    }
}

// one thread per incidence e: owner row j (binary search in s_indptr), frontier row r, and
// a_e = Â[r,v] looked up in row r (the value the reference's spmm multiplies P'[v] with; a missing
// entry means Â's pattern is not symmetric).  Thread-per-incidence keeps hub test nodes from
// serialising thousands of binary searches in one warp.
__global__ void k_fill_S(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                         const float* __restrict__ values, const int32_t* __restrict__ test_nodes, int32_t lo,
                         int32_t Ts, int64_t E, const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref,
                         const int32_t* __restrict__ s_indptr, int32_t* __restrict__ s_j,
                         int32_t* __restrict__ s_rho, float* __restrict__ s_a, int* __restrict__ err) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t j = (int32_t)upper_bound(s_indptr, 0, (int64_t)Ts + 1, (int32_t)e) - 1;
    const int32_t v = test_nodes[lo + j];
    const int32_t r = indices[indptr[v] + ((int32_t)e - s_indptr[j])];
    const int32_t rb = indptr[r], re = indptr[r + 1];
    const int64_t p = lower_bound(indices, (int64_t)rb, (int64_t)re, v);
    float a;
    if (p < re && indices[p] == v) a = values[p];
    else { a = 0.f; atomicOr(err, 4); }
    s_j[e] = j;
    s_rho[e] = frontier_rank(bits, pref, r);
    s_a[e] = a;
}

// observers: one thread per entry of the concatenated rows of all T test nodes (all_off = exclusive
// scan of their lengths), so a hub test node does not serialise its whole row in one warp.
// MODE 0 counts per frontier row, MODE 1 fills.
template <int MODE>
__global__ void k_observers(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                            const float* __restrict__ values, const int32_t* __restrict__ test_nodes, int32_t T,
                            const int32_t* __restrict__ all_off, const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref,
                            int32_t* __restrict__ cnt_or_cursor, int32_t* __restrict__ b_i, float* __restrict__ b_w) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= all_off[T]) return;
    const int32_t i = (int32_t)upper_bound(all_off, 0, (int64_t)T + 1, (int32_t)t) - 1;
    const int32_t k = indptr[test_nodes[i]] + ((int32_t)t - all_off[i]);
    const int32_t rho = bits ? frontier_rank(bits, pref, indices[k]) : indices[k];   // no bitmap: observer lists over all n nodes
    if (rho < 0) return;
    if (MODE == 0) {
        atomicAdd(&cnt_or_cursor[rho], 1);
    } else {
        const int32_t p = atomicAdd(&cnt_or_cursor[rho], 1);
        b_i[p] = i;
        b_w[p] = values[k];
    }
}

__global__ void k_rec_counts(const int32_t* __restrict__ s_rho, int64_t E, const int32_t* __restrict__ b_indptr,
                             int32_t* __restrict__ rc) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t rho = s_rho[e];
    rc[e] = b_indptr[rho + 1] - b_indptr[rho];
}

// one thread per incidence: emit its records (a frontier row has a handful of observers; a warp per
// incidence left most lanes idle)
__global__ void k_expand(const int32_t* __restrict__ s_j, const int32_t* __restrict__ s_rho, int64_t E,
                         const int32_t* __restrict__ b_indptr, const int32_t* __restrict__ b_i,
                         const float* __restrict__ b_w, const int64_t* __restrict__ rec_off,
                         uint64_t* __restrict__ key, uint64_t* __restrict__ val) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t rho = s_rho[e];
    const int32_t b0 = b_indptr[rho], nb = b_indptr[rho + 1] - b0;
    const int64_t dst = rec_off[e];
    const uint64_t jhi = (uint64_t)(uint32_t)s_j[e] << 32, ehi = (uint64_t)(uint32_t)e << 32;
    for (int32_t k = 0; k < nb; ++k) {
        key[dst + k] = jhi | (uint32_t)b_i[b0 + k];
        val[dst + k] = ehi | (uint64_t)__float_as_uint(b_w[b0 + k]);
    }
}

__global__ void k_rec_heads(const uint64_t* __restrict__ key, int64_t M, int32_t* __restrict__ head) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    head[m] = (m == 0 || key[m] != key[m - 1]) ? 1 : 0;
}
// pos = exclusive scan of head (M+1 entries); grp_off[pos[m]] = m for heads, grp_off[total] = M
__global__ void k_group_offsets(const int32_t* __restrict__ pos, int64_t M, int32_t* __restrict__ grp_off,
                                int32_t* __restrict__ n_groups) {
    int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m > M) return;
    if (m == M) { grp_off[pos[M]] = (int32_t)M; *n_groups = pos[M]; return; }
    if (pos[m + 1] != pos[m]) grp_off[pos[m]] = (int32_t)m;
}

// ------------------------------------------------------------------------------------ F1
constexpr int F1_TILE_E = 64;    // incidences per CTA (8 warps x 8)
constexpr int F1_HK = 128;       // k-chunk of the hidden dimension held in shared memory
constexpr int F1_THREADS = 256;
constexpr int F1_ASTRIDE = F1_HK + 4;

// ΔR chunk → shared (each warp writes and later reads only its own 8 rows), W2 chunk → shared,
// then acc[8][NC] += ΔR[8 x HK]·W2[HK x 32·NC] with lane ↔ output column.
template <int NC>
__global__ void __launch_bounds__(F1_THREADS)
k_frontier_dq(const int32_t* __restrict__ s_j, const int32_t* __restrict__ s_rho, const float* __restrict__ s_a,
              int64_t E, const float* __restrict__ H, int64_t ldh, const float* __restrict__ P, int64_t ldp,
              const int32_t* __restrict__ p_index, const float* __restrict__ W2, int32_t h, int32_t c, int32_t c0,
              float delta, float inv_delta, float* __restrict__ dq) {
    extern __shared__ __align__(16) float smem[];
    float* As = smem;                                   // [F1_TILE_E][F1_ASTRIDE]
    float* Ws = smem + F1_TILE_E * F1_ASTRIDE;          // [F1_HK][32*NC]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t e0 = (int64_t)blockIdx.x * F1_TILE_E + warp * 8;

    // per-warp incidence metadata (lane q < 8 holds incidence e0+q)
    int64_t my_h = -1, my_p = -1;
    float my_a = 0.f;
    if (lane < 8 && e0 + lane < E) {
        const int64_t e = e0 + lane;
        const int32_t j = s_j[e];
        my_h = (int64_t)s_rho[e] * ldh;
        my_p = (int64_t)(p_index ? p_index[j] : j) * ldp;
        my_a = s_a[e];
    }
    float acc[8][NC];
#pragma unroll
    for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int m = 0; m < NC; ++m) acc[q][m] = 0.f;

    for (int kc = 0; kc < h; kc += F1_HK) {
        __syncthreads();  // previous chunk fully consumed before Ws is overwritten
        // W2 chunk: rows kc..kc+HK, columns c0..c0+32*NC (zero padded)
        for (int idx = threadIdx.x; idx < F1_HK * 32 * NC; idx += F1_THREADS) {
            const int kk = idx / (32 * NC), cc = idx % (32 * NC);
            const int k = kc + kk, col = c0 + cc;
            Ws[idx] = (k < h && col < c) ? __ldg(W2 + (int64_t)k * c + col) : 0.f;
        }
        // ΔR chunk for this warp's 8 incidences; lane covers k = kc + 4*lane .. +3
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int64_t oh = __shfl_sync(LT_FULL, my_h, q);
            const int64_t op = __shfl_sync(LT_FULL, my_p, q);
            const float a = __shfl_sync(LT_FULL, my_a, q);
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            const int k = kc + 4 * lane;
            if (oh >= 0 && k < h) {   // h % 4 == 0 is a precondition of this kernel
                const float4 hv = __ldg(reinterpret_cast<const float4*>(H + oh + k));
                const float4 pv = __ldg(reinterpret_cast<const float4*>(P + op + k));
                const float da = delta * a;
                auto dr = [&](float hh, float p) -> float {
                    const float hp = fmaf(da, p, hh);           // H'[r,k]
                    if (hh > 0.f) return hp > 0.f ? a * p : -hh * inv_delta;
                    return hp > 0.f ? hp * inv_delta : 0.f;
                };
                t.x = dr(hv.x, pv.x); t.y = dr(hv.y, pv.y); t.z = dr(hv.z, pv.z); t.w = dr(hv.w, pv.w);
            }
            *reinterpret_cast<float4*>(As + (warp * 8 + q) * F1_ASTRIDE + 4 * lane) = t;
        }
        __syncthreads();
#pragma unroll 2
        for (int k = 0; k < F1_HK; k += 4) {
            float4 av[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) av[q] = *reinterpret_cast<const float4*>(As + (warp * 8 + q) * F1_ASTRIDE + k);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                float b[NC];
#pragma unroll
                for (int m = 0; m < NC; ++m) b[m] = Ws[(k + kk) * 32 * NC + lane + 32 * m];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float a = kk == 0 ? av[q].x : kk == 1 ? av[q].y : kk == 2 ? av[q].z : av[q].w;
#pragma unroll
                    for (int m = 0; m < NC; ++m) acc[q][m] = fmaf(a, b[m], acc[q][m]);
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int64_t e = e0 + q;
        if (e >= E) break;
#pragma unroll
        for (int m = 0; m < NC; ++m) {
            const int col = c0 + lane + 32 * m;
            if (col < c) dq[e * c + col] = acc[q][m];
        }
    }
}

// ------------------------------------------------------------------------------------ F2
// out[j,i] = sqrt(Σ_col (Σ_k w_k·ΔQ[e_k,col])²) per output pair (group of records).
// Small groups: G lanes per group, lanes ↔ columns, record loop unrolled for memory-level
// parallelism.  Groups above F2_BIG records (the (v,v) pair of a hub test node has deg(v)+1
// records) are listed in the plan and reduced by a whole CTA each, so no warp walks thousands of
// records alone.  Summation order is fixed per group ⇒ results are deterministic.
constexpr int F2_BIG = 64;
constexpr int F2_BIG_WARPS = 32;   // 1024-thread CTA per big group

// Where F2 writes an influence entry.  One GPU: the caller's row block.  Several GPUs (lt_attack_efficient_comm): the
// SAME entry goes into the full matrix of every rank — base[r] is rank r's matrix, mapped into this process by CUDA IPC,
// so the stores to r != rank travel over NVLink / NVSwitch from inside the reduce kernel and no collective follows it.
struct OutDst {
    float* base[LT_MAX_RANKS];
    int32_t n;          // destinations
    int64_t row0;       // row of the matrix that slice row 0 maps to (0 when base[0] is the caller's row block)
    int64_t ldo;
    __device__ __forceinline__ void store(uint64_t key, float v) const {
        const int64_t off = ((int64_t)(key >> 32) + row0) * ldo + (int64_t)(key & 0xffffffffu);
#pragma unroll
        for (int d = 0; d < LT_MAX_RANKS; ++d)
            if (d < n) base[d][off] = v;
    }
};
static inline OutDst single_dst(float* out, int64_t ldo) {
    OutDst d;
    for (int i = 0; i < LT_MAX_RANKS; ++i) d.base[i] = nullptr;
    d.base[0] = out; d.n = 1; d.row0 = 0; d.ldo = ldo;
    return d;
}

__device__ __forceinline__ float f2_column(const uint64_t* __restrict__ val, const float* __restrict__ dq, int32_t c /* row stride of dq */,
                                           int col, int32_t m0, int32_t m1) {
    float z0 = 0.f, z1 = 0.f, z2 = 0.f, z3 = 0.f;
    int32_t m = m0;
    for (; m + 4 <= m1; m += 4) {
        const uint64_t v0 = val[m], v1 = val[m + 1], v2 = val[m + 2], v3 = val[m + 3];
        const float d0 = dq[(int64_t)(v0 >> 32) * c + col], d1 = dq[(int64_t)(v1 >> 32) * c + col];
        const float d2 = dq[(int64_t)(v2 >> 32) * c + col], d3 = dq[(int64_t)(v3 >> 32) * c + col];
        z0 = fmaf(__uint_as_float((uint32_t)v0), d0, z0);
        z1 = fmaf(__uint_as_float((uint32_t)v1), d1, z1);
        z2 = fmaf(__uint_as_float((uint32_t)v2), d2, z2);
        z3 = fmaf(__uint_as_float((uint32_t)v3), d3, z3);
    }
    for (; m < m1; ++m) {
        const uint64_t v = val[m];
        z0 = fmaf(__uint_as_float((uint32_t)v), dq[(int64_t)(v >> 32) * c + col], z0);
    }
    return (z0 + z1) + (z2 + z3);
}

template <int G>
__global__ void __launch_bounds__(256)
k_frontier_reduce(const uint64_t* __restrict__ key, const uint64_t* __restrict__ val,
                  const int32_t* __restrict__ grp_off, const int32_t* __restrict__ n_groups,
                  const float* __restrict__ dq, int32_t lddq, int32_t c, OutDst out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t g = t / G;
    const int sub = (int)(t % G);
    bool valid = g < *n_groups;
    float ss = 0.f;
    uint64_t k0 = 0;
    if (valid) {
        const int32_t m0 = grp_off[g], m1 = grp_off[g + 1];
        if (m1 - m0 > F2_BIG) {
            valid = false;            // k_frontier_reduce_big owns this group
        } else {
            k0 = key[m0];
            const int32_t nrec = m1 - m0;
            if (nrec <= 4 && c <= 4 * G) {
                // the common case (a pair shares a couple of frontier rows): the records go into registers once and the
                // ΔQ loads of all of the lane's columns are issued together — one round trip instead of one per column.
                // Same operations in the same order as f2_column, so the bits do not depend on the path taken.
                const uint64_t v0 = val[m0], v1 = nrec > 1 ? val[m0 + 1] : 0ull, v2 = nrec > 2 ? val[m0 + 2] : 0ull,
                               v3 = nrec > 3 ? val[m0 + 3] : 0ull;
                const float w0 = __uint_as_float((uint32_t)v0), w1 = __uint_as_float((uint32_t)v1),
                            w2 = __uint_as_float((uint32_t)v2), w3 = __uint_as_float((uint32_t)v3);
                const float* r0 = dq + (int64_t)(v0 >> 32) * lddq;
                const float* r1 = dq + (int64_t)(v1 >> 32) * lddq;
                const float* r2 = dq + (int64_t)(v2 >> 32) * lddq;
                const float* r3 = dq + (int64_t)(v3 >> 32) * lddq;
                float d[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = min(sub + q * G, c - 1);          // clamped: the loads stay unconditional
                    d[q][0] = r0[col];
                    d[q][1] = nrec > 1 ? r1[col] : 0.f;
                    d[q][2] = nrec > 2 ? r2[col] : 0.f;
                    d[q][3] = nrec > 3 ? r3[col] : 0.f;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (sub + q * G >= c) break;
                    float z;
                    if (nrec == 4) {
                        z = (fmaf(w0, d[q][0], 0.f) + fmaf(w1, d[q][1], 0.f)) + (fmaf(w2, d[q][2], 0.f) + fmaf(w3, d[q][3], 0.f));
                    } else {
                        z = fmaf(w0, d[q][0], 0.f);
                        if (nrec > 1) z = fmaf(w1, d[q][1], z);
                        if (nrec > 2) z = fmaf(w2, d[q][2], z);
                        z = (z + 0.f) + (0.f + 0.f);
                    }
                    ss = fmaf(z, z, ss);
                }
            } else {
                for (int col = sub; col < c; col += G) {
                    const float z = f2_column(val, dq, lddq, col, m0, m1);
                    ss = fmaf(z, z, ss);
                }
            }
        }
    }
#pragma unroll
    for (int o = 1; o < G; o <<= 1) ss += __shfl_xor_sync(LT_FULL, ss, o);
    if (valid && sub == 0) out.store(k0, sqrtf(ss));
}

// thread per group: append the groups with more than F2_BIG records to big[]
__global__ void k_big_groups(const int32_t* __restrict__ grp_off, const int32_t* __restrict__ n_groups,
                             int32_t* __restrict__ big, int32_t* __restrict__ n_big) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= *n_groups) return;
    if (grp_off[g + 1] - grp_off[g] > F2_BIG) big[atomicAdd(n_big, 1)] = (int32_t)g;
}

// one CTA (32 warps) per big group: warp w reduces the w-th contiguous chunk of the records
// (lanes ↔ columns), partial sums meet in shared memory and are added in warp order.
__global__ void __launch_bounds__(32 * F2_BIG_WARPS)
k_frontier_reduce_big(const uint64_t* __restrict__ key, const uint64_t* __restrict__ val,
                      const int32_t* __restrict__ grp_off, const int32_t* __restrict__ big,
                      const int32_t* __restrict__ n_big, const float* __restrict__ dq, int32_t lddq, int32_t c, OutDst out) {
    __shared__ float part[F2_BIG_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int32_t b = blockIdx.x; b < *n_big; b += gridDim.x) {
        const int32_t g = big[b];
        const int32_t m0 = grp_off[g], m1 = grp_off[g + 1];
        const int32_t chunk = (m1 - m0 + F2_BIG_WARPS - 1) / F2_BIG_WARPS;
        const int32_t a0 = min(m1, m0 + warp * chunk), a1 = min(m1, a0 + chunk);
        float ss = 0.f;   // meaningful in warp 0 only
        for (int c0 = 0; c0 < c; c0 += 32) {
            const int col = c0 + lane;
            part[warp][lane] = col < c ? f2_column(val, dq, lddq, col, a0, a1) : 0.f;
            __syncthreads();
            if (warp == 0) {
                float z = 0.f;
#pragma unroll
                for (int w = 0; w < F2_BIG_WARPS; ++w) z += part[w][lane];
                ss = fmaf(z, z, ss);
            }
            __syncthreads();
        }
        if (warp == 0) {
            ss = warp_sum(ss);
            if (lane == 0) {
                out.store(key[m0], sqrtf(ss));
            }
        }
    }
}

// Listed pairs (link_prediction_attack_efficient_balanced / the naive attack: one entry per listed (perturbed, observed)
// pair, attacker.py:250-284, 143-160).  One warp per pair: N̂(p) ∩ N̂(o) by a merge of the two sorted CSR rows (binary
// search of the shorter row in the longer when they differ by more than 8x), Σ_r Â[o,r]·ΔQ_p[r,:] in ascending r, 2-norm.
// ΔQ_p[r] is row s_indptr[j] + (position of r in row p) of dq.  Lanes ↔ columns (c ≤ 256).
__global__ void __launch_bounds__(256)
k_pair_reduce(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
              const int32_t* __restrict__ v_nodes /* node id of slice row j */, const int32_t* __restrict__ s_indptr,
              const int32_t* __restrict__ pair_j, const int32_t* __restrict__ pair_obs, int64_t K,
              const float* __restrict__ dq, int32_t lddq, int32_t c, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= K) return;
    const int32_t j = pair_j[k], o = pair_obs[k];
    const int32_t p = v_nodes[j];
    const int32_t pa = indptr[p], pe = indptr[p + 1], oa = indptr[o], oe = indptr[o + 1];
    const int64_t e0 = s_indptr[j];
    float z[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) z[q] = 0.f;
    auto hit = [&](int32_t a /* position in row p */, int32_t b /* position in row o */) {
        const float w = values[b];
        const float* row = dq + (e0 + (a - pa)) * lddq;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int col = lane + 32 * q;
            if (col < c) z[q] = fmaf(w, row[col], z[q]);
        }
    };
    const int32_t lp = pe - pa, lo_ = oe - oa;
    if (lp > 8 * lo_) {                                   // few entries in row o: look each up in row p
        for (int32_t b = oa; b < oe; ++b) {
            const int32_t r = indices[b];
            const int64_t a = lower_bound(indices, (int64_t)pa, (int64_t)pe, r);
            if (a < pe && indices[a] == r) hit((int32_t)a, b);
        }
    } else if (lo_ > 8 * lp) {
        for (int32_t a = pa; a < pe; ++a) {
            const int32_t r = indices[a];
            const int64_t b = lower_bound(indices, (int64_t)oa, (int64_t)oe, r);
            if (b < oe && indices[b] == r) hit(a, (int32_t)b);
        }
    } else {
        int32_t a = pa, b = oa;
        while (a < pe && b < oe) {
            const int32_t ra = indices[a], rb = indices[b];
            if (ra == rb) { hit(a, b); ++a; ++b; }
            else if (ra < rb) ++a;
            else ++b;
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) ss = fmaf(z[q], z[q], ss);
    ss = warp_sum(ss);
    if (lane == 0) out[k] = sqrtf(ss);
}

// A operand of the tensor-core F1: row e of ΔR, computed from one H row, one P row and a_e
struct DeltaRProd {
    const int32_t* s_j; const int32_t* s_rho; const float* s_a; const int32_t* p_index;
    const float* H; int64_t ldh; const float* P; int64_t ldp; float delta, inv_delta;
    __device__ __forceinline__ void meta(int e, int64_t& o0, int64_t& o1, float& f) const {
        const int32_t j = s_j[e];
        o0 = (int64_t)s_rho[e] * ldh;
        o1 = (int64_t)(p_index ? p_index[j] : j) * ldp;
        f = s_a[e];
    }
    struct Raw { float4 h, p; };
    static constexpr int PD = 1;      // two Raw sets (64 registers) would not fit two CTAs per SM
    // h, ldh, ldp are multiples of 4 and H, P 16-byte aligned (checked by frontier_dq): always one 16-byte load each
    __device__ __forceinline__ Raw raw4(int64_t o0, int64_t o1, int k, int) const {
        return Raw{__ldg(reinterpret_cast<const float4*>(H + o0 + k)), __ldg(reinterpret_cast<const float4*>(P + o1 + k))};
    }
    __device__ __forceinline__ Raw zero() const { return Raw{make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)}; }
    __device__ __forceinline__ float value(float hh, float p, float a) const {
        const float hp = fmaf(delta * a, p, hh);           // H'[r,k]
        if (hh > 0.f) return hp > 0.f ? a * p : -hh * inv_delta;
        return hp > 0.f ? hp * inv_delta : 0.f;
    }
    __device__ __forceinline__ float4 value4(const Raw& r, float a) const {
        return make_float4(value(r.h.x, r.p.x, a), value(r.h.y, r.p.y, a), value(r.h.z, r.p.z, a), value(r.h.w, r.p.w, a));
    }
};

// A operand of F1 when only the SIGNS of H are kept (sign-mask path of lt_attack_efficient): where no ReLU of the
// row can change sign — the row's smallest |H| exceeds |δ·a|·max|P[v]| — ΔR[e,k] is a·P[v,k] on the active units
// and 0 elsewhere, exactly the values DeltaRProd computes from the H row, but read from 32 bytes instead of 1 KB.
// Rows that fail the test are corrected afterwards by k_flip_fix.
struct MaskProd {
    const int32_t* s_j; const int32_t* s_rho; const float* s_a; const int32_t* p_index;
    const uint32_t* sign_words; int32_t wpr; const float* P; int64_t ldp;
    __device__ __forceinline__ void meta(int e, int64_t& o0, int64_t& o1, float& f) const {
        const int32_t j = s_j[e];
        o0 = (int64_t)s_rho[e] * wpr;
        o1 = (int64_t)(p_index ? p_index[j] : j) * ldp;
        f = s_a[e];
    }
    struct Raw { float4 p; uint32_t bits; };
    static constexpr int PD = 2;
    __device__ __forceinline__ Raw raw4(int64_t o0, int64_t o1, int k, int) const {
        return Raw{__ldg(reinterpret_cast<const float4*>(P + o1 + k)), __ldg(sign_words + o0 + (k >> 5)) >> (k & 31)};
    }
    __device__ __forceinline__ Raw zero() const { return Raw{make_float4(0.f, 0.f, 0.f, 0.f), 0u}; }
    __device__ __forceinline__ float4 value4(const Raw& r, float a) const {
        return make_float4((r.bits & 1u) ? a * r.p.x : 0.f, (r.bits & 2u) ? a * r.p.y : 0.f, (r.bits & 4u) ? a * r.p.z : 0.f,
                           (r.bits & 8u) ? a * r.p.w : 0.f);
    }
};

// pmax[j] = max_k |P[j,k]|; one warp per row
__global__ void k_row_absmax(const float* __restrict__ P, int64_t ldp, int32_t rows, int32_t h, float* __restrict__ pmax) {
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= rows) return;
    float m = 0.f;
    for (int k = lane; k < h; k += 32) m = fmaxf(m, fabsf(P[j * ldp + k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(LT_FULL, m, o));
    if (lane == 0) pmax[j] = m;
}

// ------------------------------------------------------------------------------------ exact fix-up (float64)
// Where a ReLU of row r switches under the perturbation — |H[r,k]| ≤ |δ·Â[r,v]·P[v,k]| — the finite difference is
// −H/δ or (H + δ·a·P)/δ: it depends on the VALUE of H[r,k] to within δ·a·|P|·tol.  With δ = 1e-4 that is 1e-9 … 1e-6
// absolute while a float32 evaluation of H (a sum over deg(r)·f products) is only good to ~1e-7·Σ|terms|: on the
// twitch-shaped config (|H| ~ 3, f = 3170) the float32 result of such a unit is off by tens of per cent, and the
// influence entries it feeds by up to 7 % (measured against the reference run in float64).  Those units are rare
// (a few per thousand incidences), so they are re-evaluated exactly: the incidences with a unit within reach of a
// switch — or whose sign float32 cannot decide, |H| ≤ ε_H(r) = 2^-fix_eps·max|H[r,:]| — are queued, row r of Â·X is
// gathered in float64 (k_fix_gather, segments of FIX_SEG entries summed in order: deterministic), H[r,:] and P[v,:]
// are evaluated in float64 from it (layers.py:31-34 restated), ΔR is taken case by case in float64, and the
// difference to what F1 used for the unit (sign-mask or value form, recomputed here bit for bit) goes through W2
// into ΔQ[e,:].  The queue length is read back while F1 runs (pinned word + event), so no launch bubble is added.
constexpr int FIX_SEG = 128;

struct FixCtx {                       // what the exact fix-up needs beyond the plan: the graph, X and layer 1
    const int32_t* indptr; const int32_t* indices; const float* values;
    const float* X; int64_t ldx; int32_t f;
    const float* W1; const float* b1;
    const int32_t* v_nodes;           // node id of perturbed node j (test_nodes + lo)
    const int32_t* rows;              // node id of compact frontier row ρ (null: s_row already holds node ids)
};

__device__ __forceinline__ int32_t fix_nseg(int32_t len) { return len <= FIX_SEG ? 1 : (len + FIX_SEG - 1) / FIX_SEG; }

// sign-mask path: thread per incidence; queue it when the row's smallest |H| is within reach of a switch or of the
// float32 uncertainty of H.  counters[0] = queue length, counters[1] = Σ gather segments of the queued rows.
__global__ void k_slow_mark(const int32_t* __restrict__ s_j, const int32_t* __restrict__ s_rho, const float* __restrict__ s_a,
                            int64_t E, const uint32_t* __restrict__ minabs_bits, int32_t R, const float* __restrict__ pmax,
                            float delta, float eps_scale, const int32_t* __restrict__ rows, const int32_t* __restrict__ indptr,
                            int32_t* __restrict__ slow_list, int32_t* __restrict__ counters) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t rho = s_rho[e];
    const float thr = fabsf(delta * s_a[e]) * pmax[s_j[e]] * 1.001f + eps_scale * __uint_as_float(minabs_bits[R + rho]);
    if (!(__uint_as_float(minabs_bits[rho]) > thr)) {
        slow_list[atomicAdd(&counters[0], 1)] = (int32_t)e;
        const int32_t r = rows ? rows[rho] : rho;
        atomicAdd(&counters[1], fix_nseg(indptr[r + 1] - indptr[r]));
    }
}

// value path (H rows stored): warp per incidence, same test evaluated unit by unit on the stored row
__global__ void __launch_bounds__(256)
k_slow_mark_values(const int32_t* __restrict__ s_j, const int32_t* __restrict__ s_row, const float* __restrict__ s_a, int64_t E,
                   const float* __restrict__ H, int64_t ldh, const float* __restrict__ P, int64_t ldp,
                   const int32_t* __restrict__ p_index, int32_t h, float delta, float eps_scale,
                   const int32_t* __restrict__ rows, const int32_t* __restrict__ indptr, int32_t* __restrict__ slow_list,
                   int32_t* __restrict__ counters) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (e >= E) return;
    const int32_t j = s_j[e];
    const float* hr = H + (int64_t)s_row[e] * ldh;
    const float* pr = P + (int64_t)(p_index ? p_index[j] : j) * ldp;
    const float da = fabsf(delta * s_a[e]) * 1.001f;
    float mx = 0.f;
    for (int k = lane; k < h; k += 32) mx = fmaxf(mx, fabsf(hr[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(LT_FULL, mx, o));
    const float eps = eps_scale * mx;
    bool hit = false;
    for (int k = lane; k < h; k += 32) hit |= !(fabsf(hr[k]) > da * fabsf(pr[k]) + eps);
    if (__any_sync(LT_FULL, hit) && lane == 0) {
        slow_list[atomicAdd(&counters[0], 1)] = (int32_t)e;
        const int32_t r = rows ? rows[s_row[e]] : s_row[e];
        atomicAdd(&counters[1], fix_nseg(indptr[r + 1] - indptr[r]));
    }
}

__global__ void k_fix_nseg(const int32_t* __restrict__ slow_list, int32_t n_slow, const int32_t* __restrict__ s_row,
                           const int32_t* __restrict__ rows, const int32_t* __restrict__ indptr, int32_t* __restrict__ nseg) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slow) return;
    const int32_t rr = s_row[slow_list[q]];
    const int32_t r = rows ? rows[rr] : rr;
    nseg[q] = fix_nseg(indptr[r + 1] - indptr[r]);
}

// warp per gather segment g: part[g, i] = Σ_{t in segment} (double)Â[r,t]·(double)X[col_t, i], entries in CSR order.
// The (column, value) pairs of a segment are read 32 at a time and broadcast by shuffle; eight rows of X are in flight
// per lane before their FMAs — walked one dependent load at a time, a 512-entry segment took 0.5 ms (r2c).
__global__ void __launch_bounds__(256)      // launched with 128 threads at small f: 128 x 88 registers fit beside F1's persistent CTA
k_fix_gather(const int32_t* __restrict__ slow_list, int32_t n_slow, const int32_t* __restrict__ seg_off,
             const int32_t* __restrict__ s_row, FixCtx c, double* __restrict__ part) {
    const int lane = threadIdx.x & 31;
    const int nfb = (c.f + 127) / 128;                     // a warp takes one block of 128 features of one segment
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t g = wid / nfb;
    if (g >= seg_off[n_slow]) return;
    const int32_t q = (int32_t)upper_bound(seg_off, 0, (int64_t)n_slow + 1, (int32_t)g) - 1;
    const int32_t rr = s_row[slow_list[q]];
    const int32_t r = c.rows ? c.rows[rr] : rr;
    const int32_t t0 = c.indptr[r] + ((int32_t)g - seg_off[q]) * FIX_SEG;
    const int32_t t1 = min(c.indptr[r + 1], t0 + FIX_SEG);
    double* dst = part + (int64_t)g * c.f;
    {
        const int i0 = (int)(wid % nfb) * 128;             // 4 features per lane: i0 + lane + 32u
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        int ic[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ic[u] = min(i0 + lane + 32 * u, c.f - 1);    // clamped: loads stay unconditional
        for (int32_t base = t0; base < t1; base += 32) {
            const int32_t t = base + lane;
            const int64_t mo = t < t1 ? (int64_t)c.indices[t] * c.ldx : 0;
            const float mv = t < t1 ? c.values[t] : 0.f;
            const int cnt = min(32, t1 - base);
            for (int k = 0; k < cnt; k += 8) {
                float x[8][4];
#pragma unroll
                for (int e8 = 0; e8 < 8; ++e8) {
                    const float* xr = c.X + __shfl_sync(LT_FULL, mo, (k + e8) & 31);
#pragma unroll
                    for (int u = 0; u < 4; ++u) x[e8][u] = __ldg(xr + ic[u]);
                }
#pragma unroll
                for (int e8 = 0; e8 < 8; ++e8) {
                    const double a = (double)__shfl_sync(LT_FULL, mv, (k + e8) & 31);      // 0 past the end of the segment
                    if (k + e8 < cnt) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) acc[u] = fma(a, (double)x[e8][u], acc[u]);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (i0 + lane + 32 * u < c.f) dst[i0 + lane + 32 * u] = acc[u];
    }
}

// MODE 0: sign-mask form (what MaskProd fed F1), MODE 1: value form (DeltaRProd on the stored H row).
// One CTA per queued incidence (grid-stride); shared: d[h] corrections.
template <int MODE>
__global__ void __launch_bounds__(256)      // launched with 128 threads at small f (resident next to F1's one-CTA-per-SM GEMM)
k_exact_fix(const int32_t* __restrict__ slow_list, int32_t n_slow, const int32_t* __restrict__ seg_off,
            const int32_t* __restrict__ s_j, const int32_t* __restrict__ s_row, const float* __restrict__ s_a, FixCtx c,
            double* __restrict__ part, const float* __restrict__ P, int64_t ldp, const int32_t* __restrict__ p_index,
            const uint32_t* __restrict__ sign_words, int32_t wpr, const uint32_t* __restrict__ maxabs_bits,
            const float* __restrict__ H, int64_t ldh, const float* __restrict__ W2, int32_t h, int32_t cdim, float delta,
            float inv_delta, float eps_scale, float* __restrict__ corr /* [n_slow, cdim] */) {
    extern __shared__ float dsm[];                      // [h]
    for (int32_t q = blockIdx.x; q < n_slow; q += gridDim.x) {
        const int64_t e = slow_list[q];
        const int32_t j = s_j[e], rr = s_row[e];
        const float a = s_a[e];
        const int32_t g0 = seg_off[q], g1 = seg_off[q + 1];
        double* ax = part + (int64_t)g0 * c.f;          // the row of Â·X in float64: segments summed in order, in place
        for (int i = threadIdx.x; i < c.f; i += blockDim.x) {
            double sum = ax[i];
            int32_t g = g0 + 1;
            for (; g + 4 <= g1; g += 4) {                 // four loads in flight, added in segment order
                const double p0 = part[(int64_t)g * c.f + i], p1 = part[(int64_t)(g + 1) * c.f + i],
                             p2 = part[(int64_t)(g + 2) * c.f + i], p3 = part[(int64_t)(g + 3) * c.f + i];
                sum += p0; sum += p1; sum += p2; sum += p3;
            }
            for (; g < g1; ++g) sum += part[(int64_t)g * c.f + i];
            ax[i] = sum;
        }
        __syncthreads();
        const float* xv = c.X + (int64_t)c.v_nodes[j] * c.ldx;
        const float* prow = P + (int64_t)(p_index ? p_index[j] : j) * ldp;
        for (int k = threadIdx.x; k < h; k += blockDim.x) {
            double hd = (double)c.b1[k], pd = 0.0;
            for (int i = 0; i < c.f; ++i) {
                const double w = (double)c.W1[(int64_t)i * h + k];
                hd = fma(ax[i], w, hd);                  // H[r,k] = (Â·X)[r,:]·W1[:,k] + b1[k]
                pd = fma((double)xv[i], w, pd);          // P[v,k] = x_v·W1[:,k]
            }
            const double dap = (double)delta * (double)a * pd;
            float d = 0.f;
            float eps;
            float cur;
            const float p32 = prow[k];
            const float ap = __fmul_rn(a, p32);
            if (MODE == 0) {
                eps = eps_scale * __uint_as_float(maxabs_bits[rr]);
                const uint32_t bit = (sign_words[(int64_t)rr * wpr + (k >> 5)] >> (k & 31)) & 1u;
                cur = bit ? ap : 0.f;
            } else {
                const float* hrow = H + (int64_t)rr * ldh;
                float mx = 0.f;                          // every thread needs the row's scale: cheap (h ≤ a few hundred)
                for (int kk = 0; kk < h; ++kk) mx = fmaxf(mx, fabsf(hrow[kk]));
                eps = eps_scale * mx;
                const float hh = hrow[k];
                const float hp = fmaf(__fmul_rn(delta, a), p32, hh);
                if (hh > 0.f) cur = hp > 0.f ? ap : __fmul_rn(-hh, inv_delta);
                else cur = hp > 0.f ? __fmul_rn(hp, inv_delta) : 0.f;
            }
            if (fabs(hd) <= fabs(dap) * 1.01 + 2.0 * (double)eps) {
                const double hp = hd + dap;
                double exact;
                if (hd > 0.0) exact = hp > 0.0 ? (double)a * pd : -hd / (double)delta;
                else exact = hp > 0.0 ? hp / (double)delta : 0.0;
                d = (float)exact - cur;
            }
            dsm[k] = d;
        }
        __syncthreads();
        // the handful of units with a correction, in ascending k (one thread lists them: h ≤ a few hundred)
        int32_t* nzk = reinterpret_cast<int32_t*>(dsm + h);
        if (threadIdx.x == 0) {
            int32_t cnt = 0;
            for (int k = 0; k < h; ++k)
                if (dsm[k] != 0.f) nzk[1 + cnt++] = k;
            nzk[0] = cnt;
        }
        __syncthreads();
        const int32_t cnt = nzk[0];
        for (int col = threadIdx.x; col < cdim; col += blockDim.x) {
            float acc = 0.f;
            for (int t = 0; t < cnt; ++t) {
                const int k = nzk[1 + t];
                acc = fmaf(dsm[k], W2[(int64_t)k * cdim + col], acc);
            }
            corr[(int64_t)q * cdim + col] = acc;
        }
        __syncthreads();
    }
}

// host side of the fix-up.  `counters` (device: queue length, Σ segments) were filled by a marking kernel already queued
// on `s`; `queue_f1` launches F1 — it is called after the read-back of the counters has been queued, so the host
// learns the queue length while F1 runs.
struct FixPinned {
    int32_t* h = nullptr;
    ~FixPinned() { if (h) cudaFreeHost(h); }
};

// dq[e,:] += corr[q,:] for the queued incidences (after F1 has written ΔQ)
__global__ void k_fix_apply(const int32_t* __restrict__ slow_list, int32_t n_slow, const float* __restrict__ corr, int32_t c,
                            float* __restrict__ dq, int32_t lddq) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_slow * c) return;
    const int32_t q = (int32_t)(t / c), col = (int32_t)(t % c);
    dq[(int64_t)slow_list[q] * lddq + col] += corr[t];
}

static cudaStream_t side_stream();
static bool overlap_enabled();

// The fix-up needs nothing F1 produces: once the host knows the queue length (read back while F1 is being queued) the
// gather and the float64 evaluation run on the library's side stream NEXT TO F1 and leave their corrections in `corr`;
// a last small kernel adds them to ΔQ after F1.  `n_groups_dev` (nullable) rides along in the same read-back so that F2
// can size its grid by the number of output pairs instead of the number of records.
template <typename LaunchF1>
static int exact_fix_run(int mode, const lt_plan* p, const int32_t* s_row, const FixCtx& ctx, int32_t* slow_list,
                         int32_t* counters, const float* P, int64_t ldp, const int32_t* p_index,
                         const uint32_t* sign_words, int32_t wpr, const uint32_t* maxabs_bits, const float* H, int64_t ldh,
                         const float* W2, int32_t h, int32_t c, float delta, float eps_scale, float* dq, int32_t lddq,
                         int32_t* n_fixed_host, cudaStream_t s, LaunchF1 queue_f1) {
    static thread_local FixPinned pin;
    if (!pin.h) LT_CUDA(cudaHostAlloc((void**)&pin.h, 16, cudaHostAllocDefault));
    cudaEvent_t ev;
    LT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    struct EvGuard { cudaEvent_t e; ~EvGuard() { cudaEventDestroy(e); } } evg{ev};
    pin.h[2] = -1;
    LT_CUDA(cudaMemcpyAsync(pin.h, counters, 8, cudaMemcpyDeviceToHost, s));
    if (p->n_groups) LT_CUDA(cudaMemcpyAsync(pin.h + 2, p->n_groups, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaEventRecord(ev, s));
    LT_TRY(queue_f1());
    {
        HostScope hsync("fix_counters_wait");
        LT_CUDA(cudaEventSynchronize(ev));
    }
    const int32_t n_slow = pin.h[0], total_seg = pin.h[1];
    if (p->n_groups && pin.h[2] >= 0) const_cast<lt_plan*>(p)->n_groups_host = pin.h[2];
    if (n_fixed_host) *n_fixed_host = n_slow;
    if (n_slow == 0) return LT_OK;
    cudaStream_t sd = overlap_enabled() ? side_stream() : nullptr;
    cudaStream_t run = sd ? sd : s;
    if (run != s) LT_CUDA(cudaStreamWaitEvent(run, ev, 0));            // the queue is complete; F1 is not waited for
    Scratch scratch(run);                                              // freed on `run`, after the apply kernel (see below)
    int32_t* seg_off;
    double* part;
    float* corr;
    LT_TRY(scratch.alloc(&seg_off, (size_t)n_slow + 1));
    LT_TRY(scratch.alloc(&part, (size_t)total_seg * ctx.f));
    LT_TRY(scratch.alloc(&corr, (size_t)n_slow * c));
    {
        ProfScope prof(run != s ? "frontier_fix_side" : "frontier_fix", run);
        k_fix_nseg<<<(unsigned)ceil_div(n_slow, 256), 256, 0, run>>>(slow_list, n_slow, s_row, ctx.rows, ctx.indptr, seg_off);
        LT_LAUNCH_CHECK();
        LT_TRY(exclusive_scan_i32(seg_off, seg_off, n_slow, run));
        // small f: the fix-up is cheap and should run NEXT TO F1 (CTAs of 128 threads fit in the registers F1 leaves free);
        // large f (twitch-shaped, f = 3170): it is the longer of the two, full-size CTAs finish it sooner
        const int fix_threads = ctx.f <= 256 ? 128 : 256;
        k_fix_gather<<<(unsigned)ceil_div((int64_t)total_seg * ((ctx.f + 127) / 128) * 32, fix_threads), fix_threads, 0, run>>>(
            slow_list, n_slow, seg_off, s_row, ctx, part);
        LT_LAUNCH_CHECK();
        const unsigned grid = (unsigned)(n_slow < num_sms() * 8 ? n_slow : num_sms() * 8);
        const size_t smem = ((size_t)2 * h + 1) * sizeof(float);
        if (mode == 0)
            k_exact_fix<0><<<grid, fix_threads, smem, run>>>(slow_list, n_slow, seg_off, p->s_j, s_row, p->s_a, ctx, part, P, ldp, p_index,
                                                     sign_words, wpr, maxabs_bits, H, ldh, W2, h, c, delta, 1.0f / delta, eps_scale, corr);
        else
            k_exact_fix<1><<<grid, fix_threads, smem, run>>>(slow_list, n_slow, seg_off, p->s_j, s_row, p->s_a, ctx, part, P, ldp, p_index,
                                                     sign_words, wpr, maxabs_bits, H, ldh, W2, h, c, delta, 1.0f / delta, eps_scale, corr);
        LT_LAUNCH_CHECK();
    }
    cudaEvent_t ev2;
    LT_CUDA(cudaEventCreateWithFlags(&ev2, cudaEventDisableTiming));
    EvGuard evg2{ev2};
    if (run != s) {
        LT_CUDA(cudaEventRecord(ev2, run));
        LT_CUDA(cudaStreamWaitEvent(s, ev2, 0));
    }
    {
        ProfScope prof("frontier_fix", s);
        k_fix_apply<<<(unsigned)ceil_div((int64_t)n_slow * c, 256), 256, 0, s>>>(slow_list, n_slow, corr, c, dq, lddq);
        LT_LAUNCH_CHECK();
    }
    if (run != s) {                                                    // the side stream frees corr only after the apply kernel
        LT_CUDA(cudaEventRecord(ev2, s));
        LT_CUDA(cudaStreamWaitEvent(run, ev2, 0));
    }
    return LT_OK;
}

// ε_H(r) = eps_scale·max|H[r,:]|: what float32 cannot decide.  Measured on the B200 (profiles/h_error_probe.py, 3xTF32
// products with float32 accumulation, against float64): max over rows of |H32 − H64| / max|H[r,:]| = 1.1e-6 at f = 50,
// 1.7e-6 at f = 100, 6.8e-6 at f = 500, 4.1e-5 at f = 3170 — it grows linearly with the depth f of the first-layer
// product.  eps_scale = 2^-fix_eps·max(f,64)/64 (fix_eps = 19) stays 1.7-2.3x above those.
static inline float fix_eps_scale(int32_t f) { return ldexpf(1.0f, -opt(OPT_FIX_EPS)) * (float)(f > 64 ? f : 64) / 64.f; }

static int plan_own(lt_plan* p, void* ptr) {
    if (p->n_owned >= 24) { set_error("plan: too many buffers"); return LT_ERR_INTERNAL; }
    p->owned[p->n_owned++] = ptr;
    return LT_OK;
}
template <typename T>
static int plan_alloc(lt_plan* p, T** out, size_t count, cudaStream_t s) {
    void* q = nullptr;
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMallocAsync(&q, bytes, s);
    if (e != cudaSuccess) { set_error("plan: cudaMallocAsync(%zu) failed: %s", bytes, cudaGetErrorString(e)); return LT_ERR_NOMEM; }
    *out = reinterpret_cast<T*>(q);
    return plan_own(p, q);
}

// One stream-ordered allocation carved into the plan's arrays: the plan needs ~20 device arrays whose sizes become known
// at its two host syncs; allocating them one by one cost ~40 driver calls per attack, and at 8 GPUs (kernels of a few µs)
// the attack is bound by exactly that host time.
struct PlanArena {
    char* base = nullptr;
    size_t off = 0, cap = 0;
    static size_t pad(size_t b) { return (b + 255) / 256 * 256; }
    template <typename T>
    void want(size_t count) { cap += pad((count ? count : 1) * sizeof(T)); }
    int commit(lt_plan* p, cudaStream_t s) { return plan_alloc(p, &base, cap ? cap : 256, s); }
    template <typename T>
    T* take(size_t count) {
        T* q = reinterpret_cast<T*>(base + off);
        off += pad((count ? count : 1) * sizeof(T));
        return q;
    }
};

// A high-priority side stream per device (created once): the second half of the plan runs on it while the restricted
// SpMM — which only needs the compacted frontier rows — already occupies the caller's stream.
static cudaStream_t side_stream() {
    static cudaStream_t streams[64] = {};
    static std::mutex mu;                       // two host threads may make their first call at the same time
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) {
        int lo_p = 0, hi_p = 0;
        cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        if (cudaStreamCreateWithPriority(&streams[dev], cudaStreamNonBlocking, hi_p) != cudaSuccess) streams[dev] = nullptr;
    }
    return streams[dev];
}

// debugging switches (lt_set_option "overlap" / "sign_path"); the environment is read once per process, not per call
static bool overlap_enabled() { return opt(OPT_OVERLAP) != 0; }
static bool sign_path_enabled() { return opt(OPT_SIGN_PATH) != 0; }

// joins the side stream back into the caller's stream on every exit path of plan_build
struct StreamJoin {
    cudaStream_t main, side;
    cudaEvent_t ev = nullptr;
    StreamJoin(cudaStream_t m, cudaStream_t sd) : main(m), side(sd) {}
    ~StreamJoin() {
        if (side == main) return;
        if (!ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        cudaEventRecord(ev, side);
        cudaStreamWaitEvent(main, ev, 0);
        cudaEventDestroy(ev);
    }
};

// `early` (nullable) is called once the frontier rows, their SpMM segment offsets and the sizes R / E / total_seg are
// known: the caller enqueues the restricted precompute on `s` there, and the rest of the plan moves to the side stream.
typedef int (*plan_early_fn)(lt_plan* p, void* ctx, cudaStream_t side);   // side: where the rest of the plan will run

static int plan_build(lt_plan* p, const int32_t* indptr, const int32_t* indices, const float* values,
                      const int32_t* test_nodes, cudaStream_t s, plan_early_fn early = nullptr, void* early_ctx = nullptr) {
    const int32_t n = p->n, T = p->T, lo = p->lo, Ts = p->Ts;
    Scratch scratch(s);
    struct ScopeBox { ProfScope* p; ~ScopeBox() { delete p; } void end() { delete p; p = nullptr; } } prof_a{new ProfScope("plan", s)};
    HostScope* host_head = new HostScope("plan_head_enqueue");
    struct HostBox { HostScope*& p; ~HostBox() { delete p; p = nullptr; } } host_head_box{host_head};
    const int32_t n_words = (n + 31) / 32;
    // one zero-filled block: [frontier bitmap | test-node bitmap | sizes (8 words) | look-back descriptors of the word scan].
    // sizes = Σ segments, Σ nnz (2 x u64) | E | R | Σ all lens | err: what the first read-back needs, contiguous, so it
    // is ONE 32-byte copy; the head of the plan is memset + 3 kernels + that copy.
    uint32_t* zero_blk;
    int32_t *all_len, *pref;
    const size_t desc_words = 2 * frontier_word_scan_desc(n_words);
    const size_t zero_words = (size_t)2 * n_words + 8 + desc_words;
    LT_TRY(scratch.alloc(&zero_blk, zero_words));
    LT_TRY(scratch.alloc(&all_len, (size_t)T + 1));
    LT_TRY(scratch.alloc(&pref, (size_t)n_words + 1));
    uint32_t* bits = zero_blk;
    uint32_t* seen = zero_blk + n_words;
    uint32_t* sizes = zero_blk + 2 * (size_t)n_words;                                  // word offset 2·n_words: 8-byte aligned
    unsigned long long* seg_tot = reinterpret_cast<unsigned long long*>(sizes);
    unsigned long long* scan_desc = reinterpret_cast<unsigned long long*>(sizes + 8);
    int* err = reinterpret_cast<int*>(sizes + 7);
    LT_TRY(plan_alloc(p, &p->s_indptr, (size_t)Ts + 1, s));
    LT_CUDA(cudaMemsetAsync(zero_blk, 0, zero_words * 4, s));
    k_test_lens_scan<<<1, TL_THREADS, 0, s>>>(indptr, test_nodes, T, n, lo, p->hi, all_len, p->s_indptr,
                                              reinterpret_cast<int32_t*>(sizes + 4), seen);
    LT_LAUNCH_CHECK();
    if (p->mark_all) LT_TRY(mark_rows(indptr, indices, test_nodes, n, 0, T, all_len, bits, s));     // the same set on every rank
    else LT_TRY(mark_rows(indptr, indices, test_nodes, n, lo, Ts, p->s_indptr, bits, s));
    LT_TRY(frontier_word_scan(bits, n_words, indptr, pref, scan_desc, seg_tot, reinterpret_cast<int32_t*>(sizes + 5), s));
    static thread_local FixPinned pin;              // pinned landing zone of the size read-back
    if (!pin.h) LT_CUDA(cudaHostAlloc((void**)&pin.h, 64, cudaHostAllocDefault));
    int32_t* hs = pin.h;                            // [0..3] Σ seg, Σ nnz (2 x u64) | [4] E | [5] R | [6] Σ all lens | [7] err
    LT_CUDA(cudaMemcpyAsync(hs, sizes, 32, cudaMemcpyDeviceToHost, s));
    delete host_head; host_head = nullptr;
    {
        HostScope hsync("plan_sync1_wait");
        LT_CUDA(cudaStreamSynchronize(s));
    }
    HostScope host_mid("plan_after_sync1_enqueue");
    int herr = hs[7];
    if (herr & 1) { set_error("lt_plan_create: test node outside [0,%d)", n); return LT_ERR_DATA; }
    if (herr & 2) { set_error("lt_plan_create: duplicate test node"); return LT_ERR_DATA; }
    unsigned long long h_seg_tot[2];
    memcpy(h_seg_tot, hs, 16);
    p->E = hs[4];
    p->R = hs[5];
    const int64_t nb_cap = hs[6];
    const int32_t R = p->R;
    const int64_t E = p->E;

    PlanArena ar1;                                  // everything whose size the first sync settled (R, E, Σ test-node degrees)
    ar1.want<int32_t>((size_t)R); ar1.want<int32_t>((size_t)R + 1);
    ar1.want<int32_t>((size_t)E); ar1.want<int32_t>((size_t)E); ar1.want<float>((size_t)E);
    ar1.want<int32_t>((size_t)R + 1); ar1.want<int32_t>((size_t)R + 1);
    ar1.want<int32_t>((size_t)nb_cap); ar1.want<float>((size_t)nb_cap);
    ar1.want<int32_t>((size_t)E + 1); ar1.want<int64_t>((size_t)E + 1);
    LT_TRY(ar1.commit(p, s));
    p->rows = ar1.take<int32_t>((size_t)R);
    p->seg_off = ar1.take<int32_t>((size_t)R + 1);
    k_compact_rows<<<(unsigned)ceil_div(n_words, 256), 256, 0, s>>>(bits, pref, n_words, indptr, p->rows, p->seg_off);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(p->seg_off, p->seg_off, R, s));
    p->total_seg = (int64_t)h_seg_tot[0];
    p->nnzR = (int64_t)h_seg_tot[1];

    // ---- fork: the caller's precompute goes onto `s` now, the rest of the plan onto the side stream
    cudaStream_t s2 = s;
    if (early) {
        cudaStream_t sd = side_stream();
        if (sd) {
            cudaEvent_t ev;
            LT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            cudaError_t e1 = cudaEventRecord(ev, s);
            cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(sd, ev, 0) : e1;
            cudaEventDestroy(ev);
            if (e2 != cudaSuccess) { set_error("plan: fork onto the side stream failed: %s", cudaGetErrorString(e2)); return LT_ERR_CUDA; }
            s2 = sd;
        }
    }
    StreamJoin join(s, s2);                     // declared before scratch2: the join runs after scratch2's frees are queued
    prof_a.end();
    {
        HostScope he("precompute_enqueue");
        if (early) LT_TRY(early(p, early_ctx, s2));
    }
    s = s2;                                     // everything below runs on the side stream (or still on the caller's)
    Scratch scratch2(s);
    ProfScope prof_b(s2 != join.main ? "plan_side" : "plan", s);
    HostScope host_side("plan_side_total");
    p->s_j = ar1.take<int32_t>((size_t)E);
    p->s_rho = ar1.take<int32_t>((size_t)E);
    p->s_a = ar1.take<float>((size_t)E);
    if (E) {
        k_fill_S<<<(unsigned)ceil_div(E, 256), 256, 0, s>>>(indptr, indices, values, test_nodes, lo, Ts, E, bits, pref,
                                                            p->s_indptr, p->s_j, p->s_rho, p->s_a, err);
        LT_LAUNCH_CHECK();
    }

    if (p->listed_only) {                       // the listed-pair reduce intersects rows itself: no observers, no records
        LT_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, s));
        LT_CUDA(cudaStreamSynchronize(s));
        if (herr & 4) {
            set_error("lt_pair_influence: Â[r,v] missing for a stored Â[v,r] — the adjacency pattern is not symmetric");
            return LT_ERR_DATA;
        }
        return LT_OK;
    }

    // observers
    int32_t* cursor;
    p->b_indptr = ar1.take<int32_t>((size_t)R + 1);
    cursor = ar1.take<int32_t>((size_t)R + 1);
    p->b_i = ar1.take<int32_t>((size_t)nb_cap);
    p->b_w = ar1.take<float>((size_t)nb_cap);
    LT_CUDA(cudaMemsetAsync(p->b_indptr, 0, ((size_t)R + 1) * 4, s));
    const unsigned blT = (unsigned)ceil_div(nb_cap > 0 ? nb_cap : 1, 256);
    k_observers<0><<<blT, 256, 0, s>>>(indptr, indices, values, test_nodes, T, all_len, bits, pref, p->b_indptr, nullptr, nullptr);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(p->b_indptr, p->b_indptr, R, s));
    LT_CUDA(cudaMemcpyAsync(cursor, p->b_indptr, ((size_t)R + 1) * 4, cudaMemcpyDeviceToDevice, s));
    k_observers<1><<<blT, 256, 0, s>>>(indptr, indices, values, test_nodes, T, all_len, bits, pref, cursor, p->b_i, p->b_w);
    LT_LAUNCH_CHECK();

    // records
    int32_t* rc;
    int64_t* rec_off;
    rc = ar1.take<int32_t>((size_t)E + 1);
    rec_off = ar1.take<int64_t>((size_t)E + 1);
    if (E) {
        k_rec_counts<<<(unsigned)ceil_div(E, 256), 256, 0, s>>>(p->s_rho, E, p->b_indptr, rc);
        LT_LAUNCH_CHECK();
    }
    LT_TRY(exclusive_scan_i32_to_i64(rc, rec_off, E, s));
    // second read-back, into the pinned block too (a copy to pageable memory makes the host wait three times)
    LT_CUDA(cudaMemcpyAsync(hs + 8, rec_off + E, 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(hs + 10, p->b_indptr + R, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(hs + 11, err, sizeof(int), cudaMemcpyDeviceToHost, s));
    {
        HostScope hsync("plan_sync2_wait");
        LT_CUDA(cudaStreamSynchronize(s));
    }
    int64_t hM;
    memcpy(&hM, hs + 8, 8);
    const int32_t hNB = hs[10];
    herr = hs[11];
    if (herr & 1) { set_error("lt_plan_create: test node outside [0,%d)", n); return LT_ERR_DATA; }
    if (herr & 2) { set_error("lt_plan_create: duplicate test node"); return LT_ERR_DATA; }
    if (herr & 4) {
        set_error("lt_plan_create: Â[r,v] missing for a stored Â[v,r] — the adjacency pattern is not symmetric");
        return LT_ERR_DATA;
    }
    p->M = hM;
    p->NB = hNB;
    if (hM >= ((int64_t)1 << 31)) { set_error("lt_plan_create: %lld records exceed 2^31", (long long)hM); return LT_ERR_CAPACITY; }
    const int64_t M = hM;
    PlanArena ar2;                                  // everything sized by the record count
    ar2.want<uint64_t>((size_t)M); ar2.want<uint64_t>((size_t)M); ar2.want<int32_t>((size_t)M + 2); ar2.want<int32_t>(2);
    ar2.want<int32_t>((size_t)(M / F2_BIG + 1)); ar2.want<int32_t>((size_t)M + 1);
    LT_TRY(ar2.commit(p, s));
    p->rec_key = ar2.take<uint64_t>((size_t)M);
    p->rec_val = ar2.take<uint64_t>((size_t)M);
    p->grp_off = ar2.take<int32_t>((size_t)M + 2);
    p->n_groups = ar2.take<int32_t>(2);
    p->n_big = p->n_groups + 1;
    p->big = ar2.take<int32_t>((size_t)(M / F2_BIG + 1));
    int32_t* head_buf = ar2.take<int32_t>((size_t)M + 1);
    LT_CUDA(cudaMemsetAsync(p->n_groups, 0, 8, s));
    if (M) {
        k_expand<<<(unsigned)ceil_div(E, 256), 256, 0, s>>>(p->s_j, p->s_rho, E, p->b_indptr, p->b_i, p->b_w,
                                                               rec_off, p->rec_key, p->rec_val);
        LT_LAUNCH_CHECK();
        BitRange ranges[2] = {{0, bits_for((uint64_t)(T - 1))}, {32, 32 + bits_for((uint64_t)(Ts > 1 ? Ts - 1 : 1))}};
        LT_TRY(radix_sort_u64(p->rec_key, p->rec_val, M, ranges, Ts > 1 ? 2 : 1, s));
        int32_t* head = head_buf;
        k_rec_heads<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(p->rec_key, M, head);
        LT_LAUNCH_CHECK();
        LT_TRY(exclusive_scan_i32(head, head, M, s));
        k_group_offsets<<<(unsigned)ceil_div(M + 1, 256), 256, 0, s>>>(head, M, p->grp_off, p->n_groups);
        LT_LAUNCH_CHECK();
        k_big_groups<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(p->grp_off, p->n_groups, p->big, p->n_big);
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

// ΔQ[e,:] = ΔR[e,:]·W for every incidence e of the plan (F1).  H row of incidence e = H[h_index[e]]
// (the plan's compact frontier index by default; node ids when H holds every row), P row = P[p_index ? p_index[j] : j].
// dq: [E, lddq] with lddq = frontier_lddq(E, h, c).
// sharded precompute (several ranks): every rank must take the tensor-core / sign-mask path whatever the size of its own
// slice, or the ranks would disagree about the exchange of sign records — set for the duration of such a call
static thread_local bool g_force_tc = false;
static inline bool frontier_use_tc(int64_t E, int32_t h, int32_t c) {
    if (g_force_tc) return c <= 256 && c >= 8;
    return c <= 256 && c >= 8 && E * (int64_t)h * c >= ((int64_t)1 << 24);
}
static inline int32_t frontier_lddq(int64_t E, int32_t h, int32_t c) { return frontier_use_tc(E, h, c) ? (c + 3) / 4 * 4 : c; }

static int frontier_dq_launch(const lt_plan* p, const int32_t* h_index, const float* H, int64_t ldh, const float* P, int64_t ldp,
                              const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta, float* dq,
                              cudaStream_t s);

// `fix` (nullable): the graph / X / layer-1 weights the float64 fix-up needs.  Without it (lt_influence_rows: the caller
// hands in H_R only) the units where a ReLU switches keep their float32 evaluation.
static int frontier_dq(const lt_plan* p, const int32_t* h_index, const float* H, int64_t ldh, const float* P, int64_t ldp,
                       const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta, float* dq,
                       cudaStream_t s, const FixCtx* fix = nullptr, int32_t* n_fixed_host = nullptr) {
    if (n_fixed_host) *n_fixed_host = 0;
    if (!fix || p->E == 0) return frontier_dq_launch(p, h_index, H, ldh, P, ldp, p_index, W2, h, c, delta, dq, s);
    Scratch scratch(s);
    int32_t *slow_list, *counters;
    LT_TRY(scratch.alloc(&slow_list, (size_t)p->E));
    LT_TRY(scratch.alloc(&counters, 2));
    LT_CUDA(cudaMemsetAsync(counters, 0, 8, s));
    const float eps_scale = fix_eps_scale(fix->f);
    {
        ProfScope prof("frontier_dq", s);
        k_slow_mark_values<<<(unsigned)ceil_div(p->E * 32, 256), 256, 0, s>>>(p->s_j, h_index, p->s_a, p->E, H, ldh, P, ldp, p_index, h,
                                                                              delta, eps_scale, fix->rows, fix->indptr, slow_list, counters);
        LT_LAUNCH_CHECK();
    }
    return exact_fix_run(1, p, h_index, *fix, slow_list, counters, P, ldp, p_index, nullptr, 0, nullptr, H, ldh, W2, h, c, delta,
                         eps_scale, dq, frontier_lddq(p->E, h, c), n_fixed_host, s,
                         [&]() { return frontier_dq_launch(p, h_index, H, ldh, P, ldp, p_index, W2, h, c, delta, dq, s); });
}

static int frontier_dq_launch(const lt_plan* p, const int32_t* h_index, const float* H, int64_t ldh, const float* P, int64_t ldp,
                              const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta, float* dq,
                              cudaStream_t s) {
    const int64_t E = p->E;
    if (E == 0) return LT_OK;
    if (h % 4 != 0 || ldh % 4 != 0 || ldp % 4 != 0 || (uintptr_t)H % 16 || (uintptr_t)P % 16) {
        set_error("lt_influence_rows: h, ldh, ldp must be multiples of 4 and H_R, P 16-byte aligned");
        return LT_ERR_ARG;
    }
    Scratch scratch(s);
    const float inv_delta = 1.0f / delta;
    // ΔQ = ΔR·W2: on the tensor cores (3xTF32, tcgen05) when the product is large enough, else fp32 SIMT
    const bool use_tc = frontier_use_tc(E, h, c);
    const int32_t lddq = frontier_lddq(E, h, c);
    if (use_tc) {
        ProfScope prof("frontier_dq", s);
        const int nt = tc_pick_nt(c);
        const int32_t n_pad = nt, kchunks = tc_kchunks(h);
        uint32_t* img;
        LT_TRY(scratch.alloc(&img, (size_t)kchunks * 2 * n_pad * TC_BK));
        LT_TRY(tc_prepack_b(W2, c, 0, c, h, n_pad, img, s));
        const DeltaRProd prod{p->s_j, h_index, p->s_a, p_index, H, ldh, P, ldp, delta, inv_delta};
        if (tc_use_ws()) {
            switch (nt) {
                case 32: return launch_ws<32>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 48: return launch_ws<48>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 64: return launch_ws<64>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 128: return launch_ws<128>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                default: return launch_ws<256>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            }
        }
        switch (nt) {
            case 32: LT_TRY((launch_tc<32, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s))); break;
            case 48: LT_TRY((launch_tc<48, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s))); break;
            case 64: LT_TRY((launch_tc<64, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s))); break;
            case 128: LT_TRY((launch_tc<128, 1>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s))); break;
            default: LT_TRY((launch_tc<256, 1>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s))); break;
        }
        return LT_OK;
    }
    const unsigned bl = (unsigned)ceil_div(E, F1_TILE_E);
    for (int32_t c0 = 0; c0 < c; c0 += 128) {
        ProfScope prof("frontier_dq", s);
        const int32_t cw = c - c0 < 128 ? c - c0 : 128;
        const int nc = (cw + 31) / 32;
        const size_t smem = ((size_t)F1_TILE_E * F1_ASTRIDE + (size_t)F1_HK * 32 * nc) * sizeof(float);
#define LT_F1(NC_)                                                                                             \
    do {                                                                                                       \
        LT_CUDA(cudaFuncSetAttribute(k_frontier_dq<NC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_frontier_dq<NC_><<<bl, F1_THREADS, smem, s>>>(p->s_j, h_index, p->s_a, E, H, ldh, P, ldp, p_index, W2, h,  \
                                                        c, c0, delta, inv_delta, dq);                          \
    } while (0)
        if (nc == 1) LT_F1(1);
        else if (nc == 2) LT_F1(2);
        else if (nc == 3) LT_F1(3);
        else LT_F1(4);
#undef LT_F1
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

// F1 from the sign mask of H_R (see MaskProd): ΔQ = (mask ⊙ a·P[v])·W2 on the tensor cores, then the float64 fix-up
static int frontier_dq_signs(const lt_plan* p, const uint32_t* sign_words, int32_t wpr, const uint32_t* minabs_bits,
                             const FixCtx& fix, const float* P, int64_t ldp, const int32_t* p_index, const float* W2,
                             int32_t h, int32_t c, float delta, float* dq, int32_t* n_fixed_host, cudaStream_t s) {
    const int64_t E = p->E;
    if (n_fixed_host) *n_fixed_host = 0;
    if (E == 0) return LT_OK;
    Scratch scratch(s);
    const int32_t lddq = frontier_lddq(E, h, c);
    float* pmax;
    int32_t *slow_list, *counters;
    LT_TRY(scratch.alloc(&pmax, (size_t)p->Ts));
    LT_TRY(scratch.alloc(&slow_list, (size_t)E));
    LT_TRY(scratch.alloc(&counters, 2));
    const int nt = tc_pick_nt(c);
    const int32_t n_pad = nt, kchunks = tc_kchunks(h);
    uint32_t* img;
    LT_TRY(scratch.alloc(&img, (size_t)kchunks * 2 * n_pad * TC_BK));
    const float eps_scale = fix_eps_scale(fix.f);
    {
        ProfScope prof("frontier_dq", s);
        LT_CUDA(cudaMemsetAsync(counters, 0, 8, s));
        k_row_absmax<<<(unsigned)ceil_div((int64_t)p->Ts * 32, 256), 256, 0, s>>>(P, ldp, p->Ts, h, pmax);
        LT_LAUNCH_CHECK();
        LT_TRY(tc_prepack_b(W2, c, 0, c, h, n_pad, img, s));
        k_slow_mark<<<(unsigned)ceil_div(E, 256), 256, 0, s>>>(p->s_j, p->s_rho, p->s_a, E, minabs_bits, p->R, pmax, delta, eps_scale,
                                                               fix.rows, fix.indptr, slow_list, counters);
        LT_LAUNCH_CHECK();
    }
    const MaskProd prod{p->s_j, p->s_rho, p->s_a, p_index, sign_words, wpr, P, ldp};
    auto queue_f1 = [&]() -> int {
        ProfScope prof("frontier_dq", s);
        if (tc_use_ws()) {
            switch (nt) {
                case 32: return launch_ws<32>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 48: return launch_ws<48>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 64: return launch_ws<64>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                case 128: return launch_ws<128>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
                default: return launch_ws<256>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            }
        }
        switch (nt) {
            case 32: return launch_tc<32, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            case 48: return launch_tc<48, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            case 64: return launch_tc<64, 2>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            case 128: return launch_tc<128, 1>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
            default: return launch_tc<256, 1>(prod, img, nullptr, 0, dq, lddq, (int32_t)E, c, h, n_pad, kchunks, s);
        }
    };
    return exact_fix_run(0, p, p->s_rho, fix, slow_list, counters, P, ldp, p_index, sign_words, wpr, minabs_bits + p->R, nullptr, 0, W2,
                         h, c, delta, eps_scale, dq, lddq, n_fixed_host, s, queue_f1);
}

// out[j,i] = ‖Σ_k w_k·ΔQ[e_k,:]‖₂ for every group of sorted records (F2)
static int frontier_reduce(const uint64_t* rec_key, const uint64_t* rec_val, const int32_t* grp_off, const int32_t* n_groups,
                           const int32_t* big, const int32_t* n_big, int64_t M, const float* dq, int32_t lddq, int32_t c,
                           const OutDst& out, cudaStream_t s, int64_t n_groups_host = -1) {
    if (M == 0) return LT_OK;
    const int64_t n_work = n_groups_host >= 0 ? n_groups_host : M;       // groups when the host knows their number, else ≤ M
    ProfScope prof2("frontier_reduce", s);
    // groups above F2_BIG records (a CTA each) write other output pairs than the small groups: they go first, on the
    // side stream, and run next to the small-group kernel.  The join is RAII: every exit path below (a failed launch
    // included) makes the caller's stream wait for the side stream before the caller's scratch (dq, records) is freed.
    cudaStream_t side = s;
    if (M > F2_BIG) {
        cudaStream_t sd = overlap_enabled() ? side_stream() : nullptr;
        if (sd) {
            cudaEvent_t ev;
            LT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            cudaError_t e1 = cudaEventRecord(ev, s);
            cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(sd, ev, 0) : e1;
            cudaEventDestroy(ev);
            if (e2 != cudaSuccess) { set_error("frontier_reduce: fork failed: %s", cudaGetErrorString(e2)); return LT_ERR_CUDA; }
            side = sd;
        }
    }
    StreamJoin join(s, side);
    if (M > F2_BIG) {
        k_frontier_reduce_big<<<(unsigned)num_sms() * 2, 32 * F2_BIG_WARPS, 0, side>>>(rec_key, rec_val, grp_off, big, n_big, dq, lddq,
                                                                                       c, out);
        LT_LAUNCH_CHECK();
    }
#define LT_F2(G_)                                                                                              \
    k_frontier_reduce<G_><<<(unsigned)ceil_div((n_work > 0 ? n_work : 1) * G_, 256), 256, 0, s>>>(rec_key, rec_val, grp_off, n_groups, dq, lddq, c, out)
    // lanes per group: a group is a couple of records, so half a warp with up to four columns per lane keeps more
    // groups in flight than a whole warp (C5, c = 47: 0.188 → 0.171 ms; 8 lanes 0.189, 4 lanes 0.209)
    if (c <= 1) LT_F2(1);
    else if (c <= 2) LT_F2(2);
    else if (c <= 4) LT_F2(4);
    else if (c <= 8) LT_F2(8);
    else if (c <= 64) LT_F2(16);
    else LT_F2(32);
#undef LT_F2
    LT_LAUNCH_CHECK();
    return LT_OK;
}

int influence_rows(const lt_plan* p, const float* H_R, int64_t ldh, const float* P, int64_t ldp,
                   const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta, float* out,
                   int64_t ldo, cudaStream_t s, const FixCtx* fix = nullptr, int32_t* n_fixed_host = nullptr,
                   const OutDst* dst = nullptr /* multi-GPU: every rank's matrix, already zero-filled */) {
    const int64_t E = p->E, M = p->M;
    if (!dst) LT_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)p->T * 4, (size_t)p->Ts, s));
    if (E == 0 || M == 0) return LT_OK;
    Scratch scratch(s);
    float* dq;
    const int32_t lddq = frontier_lddq(E, h, c);
    LT_TRY(scratch.alloc(&dq, (size_t)E * lddq));
    LT_TRY(frontier_dq(p, p->s_rho, H_R, ldh, P, ldp, p_index, W2, h, c, delta, dq, s, fix, n_fixed_host));
    return frontier_reduce(p->rec_key, p->rec_val, p->grp_off, p->n_groups, p->big, p->n_big, M, dq, lddq, c,
                           dst ? *dst : single_dst(out, ldo), s);
}

static void plan_info(const lt_plan* p, int32_t h, int32_t c, int64_t* info) {
    info[0] = p->R;
    info[1] = p->E;
    info[2] = p->NB;
    info[3] = p->M;
    info[4] = p->n_groups_host;
    // algorithmic bytes of lt_influence_rows: per incidence one H row + (j,ρ,a) + ΔQ written and
    // read back once; P rows of the slice; per record key+payload; the dense output rows.
    info[5] = p->E * ((int64_t)4 * h + 12 + 8 * (int64_t)c) + (int64_t)p->Ts * 4 * h + p->M * 16 +
              (int64_t)p->Ts * p->T * 4;
    info[6] = p->nnzR;
    info[7] = 0;
}


// ==================================================================================== three layers (GCN3)
// --n-layer 3 (gcn/models.py:28-46, built by gcn_trainer.py:81-86) under the same efficient attack.
// With H1 = Â·X·W1 + b1, H2 = Â·relu(H1)·W2 + b2 (both kept for every node) the perturbation of x_v moves
//     layer 1 on N̂(v):    ΔR1[r] exactly as in the two-layer model,            ΔQ1[r] = ΔR1[r]·W2
//     layer 2 on N̂²(v):   ΔH2[s] = Σ_{r ∈ N̂(v) ∩ N̂(s)} Â[s,r]·ΔQ1[r],  ΔR2[s] = (relu(H2[s] + δ·ΔH2[s]) − relu(H2[s]))/δ
//                          (evaluated case by case on the signs, so nothing cancels),   ΔQ2[s] = ΔR2[s]·W3
//     logits on N̂³(v):    ΔZ[u] = Σ_{s ∈ N̂²(v) ∩ N̂(u)} Â[u,s]·ΔQ2[s],   influence = ‖ΔZ[u]‖₂ for the test nodes u.
// Stage 1 is the two-layer plan's incidence list S and F1.  Stage 2 expands every incidence (j,r) over row r
// into records keyed (j,s), radix-sorts them (stable: contributions stay in ascending r) and reduces each
// group against one H2 row.  Stage 3 is F2 over the (j,s) incidences with observer lists for all n nodes.
// The perturbed nodes are processed in chunks sized so that the stage-2 records fit the memory budget.

__global__ void k3_relu_copy(const float4* __restrict__ src, float4* __restrict__ dst, int64_t count4) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count4) return;
    float4 v = src[i];
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    dst[i] = v;
}

// cost[j] = Σ_{r ∈ N̂(v_j)} nnz(Â[r,:]) = stage-2 records of perturbed node j; one warp per node of the slice
__global__ void k3_costs(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                         const int32_t* __restrict__ test_nodes, int32_t lo, int32_t Ts, int32_t n, int64_t* __restrict__ cost) {
    const int lane = threadIdx.x & 31;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (j >= Ts) return;
    const int32_t v = test_nodes[lo + j];
    int64_t acc = 0;
    if (v >= 0 && v < n)
        for (int32_t k = indptr[v] + lane; k < indptr[v + 1]; k += 32) {
            const int32_t r = indices[k];
            acc += indptr[r + 1] - indptr[r];
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(LT_FULL, acc, o);
    if (lane == 0) cost[j] = acc;
}

// per incidence: node id of its frontier row and the length of that row
__global__ void k3_inc_rows(const int32_t* __restrict__ s_rho, const int32_t* __restrict__ rows, int64_t E,
                            const int32_t* __restrict__ indptr, int32_t* __restrict__ s_node, int32_t* __restrict__ len) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const int32_t r = rows[s_rho[e]];
    s_node[e] = r;
    len[e] = indptr[r + 1] - indptr[r];
}

// one thread per stage-2 record m: incidence e = (j,r) owning it (binary search in off), s = its (m − off[e])-th
// neighbour of r, w = Â[s,r] looked up in row s (the value the layer-2 spmm multiplies ΔQ1[r] with)
__global__ void k3_expand(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                          const float* __restrict__ values, const int32_t* __restrict__ s_j,
                          const int32_t* __restrict__ s_node, const int64_t* __restrict__ off, int64_t E, int64_t M,
                          uint64_t* __restrict__ key, uint64_t* __restrict__ val, int* __restrict__ err) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const int64_t e = upper_bound(off, 0, E + 1, m) - 1;
    const int32_t r = s_node[e];
    const int32_t sn = indices[indptr[r] + (int32_t)(m - off[e])];
    const int32_t sb = indptr[sn], se = indptr[sn + 1];
    const int64_t p = lower_bound(indices, (int64_t)sb, (int64_t)se, r);
    float w = 0.f;
    if (p < se && indices[p] == r) w = values[p];
    else atomicOr(err, 4);
    key[m] = ((uint64_t)(uint32_t)s_j[e] << 32) | (uint32_t)sn;
    val[m] = ((uint64_t)(uint32_t)e << 32) | (uint64_t)__float_as_uint(w);
}

// one warp per group g = (j,s): ΔH2 = Σ w·ΔQ1[e,:] in record order, ΔR2 against H2[s]; lanes ↔ columns
__global__ void __launch_bounds__(256)
k3_middle(const uint64_t* __restrict__ key, const uint64_t* __restrict__ val, const int32_t* __restrict__ grp_off,
          const int32_t* __restrict__ n_groups, const float* __restrict__ dq1, int32_t lddq1,
          const float* __restrict__ H2, int64_t ldh2, int32_t h2, float delta, float inv_delta,
          float* __restrict__ dr2, int64_t lddr2, int32_t* __restrict__ e2_j, int32_t* __restrict__ e2_s) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= *n_groups) return;
    const int32_t m0 = grp_off[g], m1 = grp_off[g + 1];
    const uint64_t k0 = key[m0];
    const int32_t sn = (int32_t)(k0 & 0xffffffffu);
    if (lane == 0) { e2_j[g] = (int32_t)(k0 >> 32); e2_s[g] = sn; }
    for (int col = lane; col < h2; col += 32) {
        const float z = f2_column(val, dq1, lddq1, col, m0, m1);
        const float hh = H2[(int64_t)sn * ldh2 + col];
        const float hp = fmaf(delta, z, hh);               // H2'[s,col]
        float d;
        if (hh > 0.f) d = hp > 0.f ? z : -hh * inv_delta;
        else d = hp > 0.f ? hp * inv_delta : 0.f;
        dr2[g * lddr2 + col] = d;
    }
}

struct Gcn3Info { int64_t chunks = 0, E1 = 0, M2 = 0, E2 = 0, M3 = 0, pairs = 0; };

// sorted records → group offsets (+ list of big groups when `big` is given); returns nothing on the host
static int group_records(const uint64_t* key, int64_t M, int32_t* grp_off, int32_t* n_groups, int32_t* big, int32_t* n_big,
                         Scratch& scratch, cudaStream_t s) {
    int32_t* head;
    LT_TRY(scratch.alloc(&head, (size_t)M + 1));
    LT_CUDA(cudaMemsetAsync(n_groups, 0, 4, s));
    k_rec_heads<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(key, M, head);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(head, head, M, s));
    k_group_offsets<<<(unsigned)ceil_div(M + 1, 256), 256, 0, s>>>(head, M, grp_off, n_groups);
    LT_LAUNCH_CHECK();
    if (big) {
        LT_CUDA(cudaMemsetAsync(n_big, 0, 4, s));
        k_big_groups<<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(grp_off, n_groups, big, n_big);
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

// influence rows of the perturbed nodes test_nodes[lo:hi) for one chunk; out → row lo of the caller's block
static int gcn3_chunk(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                      const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, const float* P, const float* H1,
                      int64_t ld1, const float* H2, int64_t ld2, const float* W2, const float* W3, int32_t h1, int32_t h2,
                      int32_t c, const int32_t* b_indptr, const int32_t* b_i, const float* b_w, float delta, float* out,
                      int64_t ldo, Gcn3Info* info, const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                      cudaStream_t s) {
    const int32_t Ts = hi - lo;
    // layer 1 of GCN3 is the two-layer model's layer 1: same float64 fix-up of the switching units (s_node = node ids)
    const FixCtx fix_ctx{indptr, indices, values, X, ldx, f, W1, b1, test_nodes + lo, nullptr};
    const FixCtx* fix1 = &fix_ctx;
    LT_CUDA(cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)T * 4, (size_t)Ts, s));
    lt_plan* plan = nullptr;
    LT_TRY(lt_plan_create(indptr, indices, values, n, test_nodes, T, lo, hi, (lt_stream_t)s, &plan));
    struct PlanGuard { lt_plan* p; cudaStream_t s; ~PlanGuard() { lt_plan_destroy(p, (lt_stream_t)s); } } guard{plan, s};
    const int64_t E1 = plan->E;
    if (E1 == 0) return LT_OK;
    Scratch scratch(s);
    const float inv_delta = 1.0f / delta;

    // ---- stage 1: ΔQ1 = ΔR1·W2 per incidence (j,r)
    int32_t *s_node, *len1;
    int64_t* off2;
    float* dq1;
    const int32_t lddq1 = frontier_lddq(E1, h1, h2);
    LT_TRY(scratch.alloc(&s_node, (size_t)E1));
    LT_TRY(scratch.alloc(&len1, (size_t)E1 + 1));
    LT_TRY(scratch.alloc(&off2, (size_t)E1 + 1));
    LT_TRY(scratch.alloc(&dq1, (size_t)E1 * lddq1));
    k3_inc_rows<<<(unsigned)ceil_div(E1, 256), 256, 0, s>>>(plan->s_rho, plan->rows, E1, indptr, s_node, len1);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32_to_i64(len1, off2, E1, s));
    LT_TRY(frontier_dq(plan, s_node, H1, ld1, P, ld1, test_nodes + lo, W2, h1, h2, delta, dq1, s, fix1));
    int64_t M2 = 0;
    LT_CUDA(cudaMemcpyAsync(&M2, off2 + E1, 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (M2 >= ((int64_t)1 << 31)) { set_error("lt_attack_efficient3: %lld stage-2 records in one chunk exceed 2^31", (long long)M2); return LT_ERR_CAPACITY; }
    if (M2 == 0) return LT_OK;

    // ---- stage 2: records (j,s) ← (e, Â[s,r]); sort; ΔR2 per group; ΔQ2 = ΔR2·W3
    uint64_t *key2, *val2;
    int32_t *grp2, *ng2;
    int* err;
    LT_TRY(scratch.alloc(&key2, (size_t)M2));
    LT_TRY(scratch.alloc(&val2, (size_t)M2));
    LT_TRY(scratch.alloc(&grp2, (size_t)M2 + 2));
    LT_TRY(scratch.alloc(&ng2, 1));
    LT_TRY(scratch.alloc(&err, 1));
    LT_CUDA(cudaMemsetAsync(err, 0, sizeof(int), s));
    {
        ProfScope prof("gcn3_plan2", s);
        k3_expand<<<(unsigned)ceil_div(M2, 256), 256, 0, s>>>(indptr, indices, values, plan->s_j, s_node, off2, E1, M2, key2, val2, err);
        LT_LAUNCH_CHECK();
        BitRange ranges[2] = {{0, bits_for((uint64_t)(n - 1))}, {32, 32 + bits_for((uint64_t)(Ts > 1 ? Ts - 1 : 1))}};
        LT_TRY(radix_sort_u64(key2, val2, M2, ranges, Ts > 1 ? 2 : 1, s));
        LT_TRY(group_records(key2, M2, grp2, ng2, nullptr, nullptr, scratch, s));
    }
    int32_t E2 = 0;
    int herr = 0;
    LT_CUDA(cudaMemcpyAsync(&E2, ng2, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (herr & 4) {
        set_error("lt_attack_efficient3: Â[s,r] missing for a stored Â[r,s] — the adjacency pattern is not symmetric");
        return LT_ERR_DATA;
    }
    float *dr2, *dq2;
    int32_t *e2_j, *e2_s, *rc3;
    int64_t* off3;
    const int64_t lddr2 = (h2 + 3) / 4 * 4;
    LT_TRY(scratch.alloc(&dr2, (size_t)E2 * lddr2));
    LT_TRY(scratch.alloc(&dq2, (size_t)E2 * c));
    LT_TRY(scratch.alloc(&e2_j, (size_t)E2));
    LT_TRY(scratch.alloc(&e2_s, (size_t)E2));
    LT_TRY(scratch.alloc(&rc3, (size_t)E2 + 1));
    LT_TRY(scratch.alloc(&off3, (size_t)E2 + 1));
    {
        ProfScope prof("gcn3_middle", s);
        if (lddr2 != h2) LT_CUDA(cudaMemsetAsync(dr2, 0, (size_t)E2 * lddr2 * 4, s));
        k3_middle<<<(unsigned)ceil_div((int64_t)E2 * 32, 256), 256, 0, s>>>(key2, val2, grp2, ng2, dq1, lddq1, H2, ld2, h2, delta,
                                                                           inv_delta, dr2, lddr2, e2_j, e2_s);
        LT_LAUNCH_CHECK();
    }
    {
        ProfScope prof("gcn3_gemm_q2", s);
        LT_TRY(gemm_f32(dr2, lddr2, nullptr, W3, c, 0, nullptr, 0, dq2, c, E2, c, h2, 0, s));
    }

    // ---- stage 3: records (j,i) ← (e2, Â[u_i,s]) from the observer lists; sort; 2-norm per output pair
    k_rec_counts<<<(unsigned)ceil_div(E2, 256), 256, 0, s>>>(e2_s, E2, b_indptr, rc3);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32_to_i64(rc3, off3, E2, s));
    int64_t M3 = 0;
    LT_CUDA(cudaMemcpyAsync(&M3, off3 + E2, 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    if (M3 >= ((int64_t)1 << 31)) { set_error("lt_attack_efficient3: %lld stage-3 records in one chunk exceed 2^31", (long long)M3); return LT_ERR_CAPACITY; }
    info->chunks += 1; info->E1 += E1; info->M2 += M2; info->E2 += E2; info->M3 += M3;
    if (M3 == 0) return LT_OK;
    uint64_t *key3, *val3;
    int32_t *grp3, *ng3, *big3, *nbig3;
    LT_TRY(scratch.alloc(&key3, (size_t)M3));
    LT_TRY(scratch.alloc(&val3, (size_t)M3));
    LT_TRY(scratch.alloc(&grp3, (size_t)M3 + 2));
    LT_TRY(scratch.alloc(&ng3, 1));
    LT_TRY(scratch.alloc(&big3, (size_t)(M3 / F2_BIG + 1)));
    LT_TRY(scratch.alloc(&nbig3, 1));
    {
        ProfScope prof("gcn3_plan3", s);
        k_expand<<<(unsigned)ceil_div(E2, 256), 256, 0, s>>>(e2_j, e2_s, E2, b_indptr, b_i, b_w, off3, key3, val3);
        LT_LAUNCH_CHECK();
        BitRange ranges[2] = {{0, bits_for((uint64_t)(T - 1))}, {32, 32 + bits_for((uint64_t)(Ts > 1 ? Ts - 1 : 1))}};
        LT_TRY(radix_sort_u64(key3, val3, M3, ranges, Ts > 1 ? 2 : 1, s));
        LT_TRY(group_records(key3, M3, grp3, ng3, big3, nbig3, scratch, s));
    }
    LT_TRY(frontier_reduce(key3, val3, grp3, ng3, big3, nbig3, M3, dq2, c, c, single_dst(out, ldo), s));
    if (info) {
        int32_t g = 0;
        LT_CUDA(cudaMemcpyAsync(&g, ng3, 4, cudaMemcpyDeviceToHost, s));
        LT_CUDA(cudaStreamSynchronize(s));
        info->pairs += g;
    }
    return LT_OK;
}

int attack_efficient3(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n, const float* X,
                      int64_t ldx, int32_t f, const float* W1, const float* b1, const float* W2, const float* b2,
                      const float* W3, int32_t h1, int32_t h2, int32_t c, const int32_t* test_nodes, int32_t T, int32_t lo,
                      int32_t hi, float delta, float* out, int64_t ldo, int64_t max_chunk_records, int64_t* info_host,
                      cudaStream_t s) {
    const int32_t Ts = hi - lo;
    Scratch scratch(s);
    // ---- dense precompute for every node: P = X·W1, H1 = Â·P + b1, H2 = Â·(relu(H1)·W2) + b2   (layers.py:31-34 twice)
    const int64_t ld1 = h1, ld2 = (h2 + 3) / 4 * 4;
    float *P, *H1, *R1, *G, *H2;
    LT_TRY(scratch.alloc(&P, (size_t)n * ld1));
    LT_TRY(scratch.alloc(&H1, (size_t)n * ld1));
    LT_TRY(scratch.alloc(&R1, (size_t)n * ld1));
    LT_TRY(scratch.alloc(&G, (size_t)n * ld2));
    LT_TRY(scratch.alloc(&H2, (size_t)n * ld2));
    {
        ProfScope prof("gcn3_precompute", s);
        LT_TRY(gemm_f32(X, ldx, nullptr, W1, h1, 0, nullptr, 0, P, ld1, n, h1, f, 0, s));
        LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, P, ld1, h1, b1, 0, H1, ld1, s));
        const int64_t count4 = (int64_t)n * ld1 / 4;
        k3_relu_copy<<<(unsigned)ceil_div(count4, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(H1), reinterpret_cast<float4*>(R1), count4);
        LT_LAUNCH_CHECK();
        if (ld2 != h2) LT_CUDA(cudaMemsetAsync(G, 0, (size_t)n * ld2 * 4, s));
        LT_TRY(gemm_f32(R1, ld1, nullptr, W2, h2, 0, nullptr, 0, G, ld2, n, h2, h1, 0, s));
        LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, G, ld2, h2, b2, 0, H2, ld2, s));
    }

    // ---- observer lists over all n nodes: s → (i, Â[u_i,s]) for every test node u_i adjacent to s
    int32_t *all_len, *b_indptr, *cursor, *b_i;
    float* b_w;
    int64_t* cost;
    LT_TRY(scratch.alloc(&all_len, (size_t)T + 1));
    LT_TRY(scratch.alloc(&b_indptr, (size_t)n + 1));
    LT_TRY(scratch.alloc(&cursor, (size_t)n + 1));
    LT_TRY(scratch.alloc(&cost, (size_t)Ts));
    k_test_lens<<<(unsigned)ceil_div(T, 256), 256, 0, s>>>(indptr, test_nodes, T, n, 0, 0, all_len, nullptr);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(all_len, all_len, T, s));
    k3_costs<<<(unsigned)ceil_div((int64_t)Ts * 32, 256), 256, 0, s>>>(indptr, indices, test_nodes, lo, Ts, n, cost);
    LT_LAUNCH_CHECK();
    int32_t nb_cap = 0;
    int64_t* cost_h = (int64_t*)malloc((size_t)Ts * 8);
    if (!cost_h) { set_error("lt_attack_efficient3: host allocation failed"); return LT_ERR_NOMEM; }
    struct HostGuard { void* p; ~HostGuard() { free(p); } } hguard{cost_h};
    LT_CUDA(cudaMemcpyAsync(&nb_cap, all_len + T, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(cost_h, cost, (size_t)Ts * 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    LT_TRY(scratch.alloc(&b_i, (size_t)nb_cap));
    LT_TRY(scratch.alloc(&b_w, (size_t)nb_cap));
    LT_CUDA(cudaMemsetAsync(b_indptr, 0, ((size_t)n + 1) * 4, s));
    const unsigned blT = (unsigned)ceil_div(nb_cap > 0 ? nb_cap : 1, 256);
    k_observers<0><<<blT, 256, 0, s>>>(indptr, indices, values, test_nodes, T, all_len, nullptr, nullptr, b_indptr, nullptr, nullptr);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(b_indptr, b_indptr, n, s));
    LT_CUDA(cudaMemcpyAsync(cursor, b_indptr, ((size_t)n + 1) * 4, cudaMemcpyDeviceToDevice, s));
    k_observers<1><<<blT, 256, 0, s>>>(indptr, indices, values, test_nodes, T, all_len, nullptr, nullptr, cursor, b_i, b_w);
    LT_LAUNCH_CHECK();

    // ---- chunks of perturbed nodes: stage-2 records per chunk bounded by the memory budget
    int64_t cap = max_chunk_records;
    if (cap <= 0) {
        size_t free_b = 0, total_b = 0;
        LT_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const int64_t per_rec = 48 + 4 * (int64_t)ld2 + 4 * (int64_t)c + 192;   // records ×2 (sort), ΔR2/ΔQ2 rows, stage-3 records
        cap = (int64_t)(free_b / 3) / per_rec;
        if (cap > ((int64_t)1 << 30)) cap = (int64_t)1 << 30;
        if (cap < 1) cap = 1;
    }
    Gcn3Info info;
    int32_t a = 0;
    while (a < Ts) {
        int32_t b = a;
        int64_t acc = 0;
        while (b < Ts && (b == a || acc + cost_h[b] <= cap)) acc += cost_h[b++];
        LT_TRY(gcn3_chunk(indptr, indices, values, n, test_nodes, T, lo + a, lo + b, P, H1, ld1, H2, ld2, W2, W3, h1, h2, c,
                          b_indptr, b_i, b_w, delta, out + (int64_t)a * ldo, ldo, &info, X, ldx, f, W1, b1, s));
        a = b;
    }
    if (info_host) {
        info_host[0] = info.chunks; info_host[1] = info.E1; info_host[2] = info.M2; info_host[3] = info.E2;
        info_host[4] = info.M3; info_host[5] = info.pairs; info_host[6] = 0; info_host[7] = 0;
    }
    return LT_OK;
}

}  // namespace lt

using namespace lt;

extern "C" int lt_plan_create(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                              const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, lt_stream_t stream,
                              lt_plan** plan_host) {
    LT_CHECK_ARG(indptr && indices && values && test_nodes && plan_host, "lt_plan_create: null pointer");
    LT_CHECK_ARG(n > 0 && T > 0 && lo >= 0 && hi >= lo && hi <= T, "lt_plan_create: need 0 <= lo <= hi <= T");
    cudaStream_t s = (cudaStream_t)stream;
    lt_plan* p = new lt_plan();
    p->n = n; p->T = T; p->lo = lo; p->hi = hi; p->Ts = hi - lo;
    const int st = plan_build(p, indptr, indices, values, test_nodes, s);
    if (st != LT_OK) { lt_plan_destroy(p, stream); return st; }
    *plan_host = p;
    return LT_OK;
}

extern "C" int lt_plan_destroy(lt_plan* plan, lt_stream_t stream) {
    if (!plan) return LT_OK;
    for (int i = plan->n_owned - 1; i >= 0; --i) cudaFreeAsync(plan->owned[i], (cudaStream_t)stream);
    delete plan;
    return LT_OK;
}

extern "C" int lt_plan_info(const lt_plan* plan, int32_t h, int32_t c, int64_t* info_host, lt_stream_t stream) {
    LT_CHECK_ARG(plan && info_host, "lt_plan_info");
    cudaStream_t s = (cudaStream_t)stream;
    lt_plan* p = const_cast<lt_plan*>(plan);
    if (p->n_groups_host < 0) {
        int32_t g = 0;
        LT_CUDA(cudaMemcpyAsync(&g, p->n_groups, 4, cudaMemcpyDeviceToHost, s));
        LT_CUDA(cudaStreamSynchronize(s));
        p->n_groups_host = g;
    }
    plan_info(p, h, c, info_host);
    return LT_OK;
}

extern "C" const int32_t* lt_plan_rows(const lt_plan* plan) { return plan ? plan->rows : nullptr; }

extern "C" int lt_plan_copy_rows(const lt_plan* plan, int32_t* dst, lt_stream_t stream) {
    LT_CHECK_ARG(plan && dst, "lt_plan_copy_rows");
    LT_CUDA(cudaMemcpyAsync(dst, plan->rows, (size_t)plan->R * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return LT_OK;
}

extern "C" int lt_influence_rows(const lt_plan* plan, const float* H_R, int64_t ldh, const float* P, int64_t ldp,
                                 const int32_t* p_index, const float* W2, int32_t h, int32_t c, float delta,
                                 float* out, int64_t ldo, lt_stream_t stream) {
    LT_CHECK_ARG(plan && H_R && P && W2 && out, "lt_influence_rows: null pointer");
    LT_CHECK_ARG(h > 0 && c > 0 && delta != 0.f && ldo >= plan->T && ldh >= h && ldp >= h, "lt_influence_rows: sizes");
    return influence_rows(plan, H_R, ldh, P, ldp, p_index, W2, h, c, delta, out, ldo, (cudaStream_t)stream);
}

// plan + restricted precompute + frontier in one call (the whole timed hot path of one rank).
// The restricted precompute only needs the compacted frontier rows, so it is enqueued on the caller's stream from the
// plan's `early` hook while the second half of the plan (incidences, observers, records, sort) runs on the side stream.
namespace {
struct AttackCtx {
    const int32_t *indptr, *indices; const float* values; int32_t n;
    const float* X; int64_t ldx; int32_t f;
    const float *W1, *b1; int32_t h, c;
    const int32_t* test_nodes; int32_t lo;
    cudaStream_t s; Scratch* scratch;
    // produced by the hook
    bool sign_path = false;
    float *Pt = nullptr, *HR = nullptr, *AX = nullptr;
    uint32_t *signs = nullptr, *minabs = nullptr;
    int32_t wpr = 0;
    HubPanel hub;                       // hot rows of X kept in the L2 for the restricted SpMM (built on the side stream)
    float hint_deg = 0.f;               // column degree from which a gathered X row is hinted evict_last (0: no hints)
    cudaEvent_t hub_ready = nullptr;
    cudaEvent_t pv_ready = nullptr;     // P_v = X[v]·W1 of the slice: started at entry on the side stream, needed by F1 only
    lt_comm* comm = nullptr;            // several ranks
    bool shard = false;                 // … that shard the restricted precompute by (global) frontier rows
};

// seg_rel[i] = seg_off[i] − seg_off[0] for i ≤ n: the SpMM work-item offsets of a share of the (globally scanned) rows
__global__ void k_rebase(const int32_t* __restrict__ seg_off, int32_t n, int32_t* __restrict__ seg_rel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) seg_rel[i] = seg_off[i] - seg_off[0];
}

// This rank's share [r0, r0+Rg) of the sign records → the aux region of every rank: the sign words are already in the local
// region (the GEMM wrote them there), the min / max |H| of the share come from `minmax` ([0,Rg) min, [Rg,2Rg) max).
// Layout of an aux region: sign words [R x wpr], min |H| [R], max |H| [R].
__global__ void k_push_signs(lt_comm c, int32_t R, int32_t wpr, int32_t r0, int32_t Rg, const uint32_t* __restrict__ minmax) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t n_sign = (int64_t)Rg * wpr;
    if (idx < n_sign) {
        const int64_t at = (int64_t)r0 * wpr + idx;
        const uint32_t w = static_cast<const uint32_t*>(c.aux[c.rank])[at];
        for (int r = 0; r < c.world; ++r)
            if (r != c.rank) static_cast<uint32_t*>(c.aux[r])[at] = w;
    } else if (idx < n_sign + 2 * (int64_t)Rg) {
        const int64_t i = idx - n_sign;                                   // [0,Rg): min, [Rg,2Rg): max
        const int64_t at = (int64_t)R * wpr + (i < Rg ? r0 + i : (int64_t)R + r0 + (i - Rg));
        const uint32_t v = minmax[i];
        for (int r = 0; r < c.world; ++r) static_cast<uint32_t*>(c.aux[r])[at] = v;
    }
}

int attack_early(lt_plan* plan, void* vctx, cudaStream_t side) {
    AttackCtx& a = *static_cast<AttackCtx*>(vctx);
    cudaStream_t s = a.s;
    Scratch& scratch = *a.scratch;
    const int32_t R = plan->R, Ts = plan->Ts, n = a.n, f = a.f, h = a.h, c = a.c;
    (void)side;
    const bool aggregate_first = (f < h) && (f % 4 == 0) && (a.ldx % 4 == 0) && ((uintptr_t)a.X % 16 == 0);
    if (a.shard) {
        // R is the frontier of ALL test nodes here — the same rows, in the same order, on every rank.  Every decision below
        // depends on global quantities only, so the ranks agree on it.
        lt_comm* cm = a.comm;
        const int32_t wpr = (h + 31) / 32;
        const bool fits = (size_t)R * (wpr + 2) * 4 <= cm->aux_bytes && (int64_t)R * h * f >= ((int64_t)1 << 24) && R >= cm->world;
        if (fits) {
            const int32_t r0 = (int32_t)((int64_t)R * cm->rank / cm->world), r1 = (int32_t)((int64_t)R * (cm->rank + 1) / cm->world);
            const int32_t Rg = r1 - r0;
            uint32_t* aux = static_cast<uint32_t*>(cm->aux[cm->rank]);
            uint32_t* minmax;
            int32_t* seg_rel;
            LT_TRY(scratch.alloc(&a.AX, (size_t)Rg * f));
            LT_TRY(scratch.alloc(&minmax, (size_t)2 * Rg));
            LT_TRY(scratch.alloc(&seg_rel, (size_t)Rg + 1));
            {
                // the share's work items: the global offsets rebased; their number is not on the host, so the grid and the
                // partial-sum buffer take an upper bound (the share holds at most the global count minus one item per other
                // row) and the surplus warps leave at once — cheaper than a host round trip for the exact number
                ProfScope prof("spmm_ax", s);
                k_rebase<<<(unsigned)ceil_div((int64_t)Rg + 1, 256), 256, 0, s>>>(plan->seg_off + r0, Rg, seg_rel);
                LT_LAUNCH_CHECK();
                const int64_t bound = plan->total_seg - (int64_t)(R - Rg);
                LT_TRY(spmm_csr_f32_seg(a.indptr, a.indices, a.values, n, plan->rows + r0, Rg, a.X, a.ldx, f, nullptr, 0, a.AX, f,
                                        seg_rel, bound, s, nullptr, a.hint_deg));
            }
            {
                ProfScope prof("gemm_h", s);
                LT_CUDA(cudaMemsetAsync(minmax, 0x7f, (size_t)Rg * 4, s));
                LT_CUDA(cudaMemsetAsync(minmax + Rg, 0, (size_t)Rg * 4, s));
                LT_TRY(gemm_3xtf32_signs(a.AX, f, a.W1, h, a.b1, Rg, h, f, aux + (size_t)r0 * wpr, minmax, s));
            }
            {
                ProfScope prof("comm_push_signs", s);
                const int64_t words = (int64_t)Rg * (wpr + 2);
                k_push_signs<<<(unsigned)ceil_div(words, 256), 256, 0, s>>>(*cm, R, wpr, r0, Rg, minmax);
                LT_LAUNCH_CHECK();
                LT_TRY(comm_signal(cm, 2, s));
                LT_TRY(comm_wait(cm, 2, s));
            }
            a.wpr = wpr;
            a.signs = aux;
            a.minabs = aux + (size_t)R * wpr;
            a.sign_path = true;
            return LT_OK;
        }
        // too small / does not fit: every rank evaluates all R rows itself (redundant, consistent)
    }
    // sign-mask path: H_R is never written — the GEMM keeps one bit per element and the row's smallest magnitude
    // (36 bytes per frontier row instead of 4h), F1 reads those, and the rows where a ReLU can switch are patched
    a.sign_path = aggregate_first && frontier_use_tc(plan->E, h, c) && h >= 16 && (int64_t)R * h * f >= ((int64_t)1 << 24) &&
                  (uintptr_t)a.W1 % 16 == 0 && sign_path_enabled();
    if (aggregate_first) {
        // H_R = (Â[R,:]·X)·W1 + b1 : gathers f floats per non-zero instead of h
        LT_TRY(scratch.alloc(&a.AX, (size_t)R * f));
        if (a.hub_ready) LT_CUDA(cudaStreamWaitEvent(s, a.hub_ready, 0));
        {
            ProfScope prof("spmm_ax", s);
            LT_TRY(spmm_csr_f32_seg(a.indptr, a.indices, a.values, n, plan->rows, R, a.X, a.ldx, f, nullptr, 0, a.AX, f,
                                    plan->seg_off, plan->total_seg, s, a.hub.valid ? &a.hub : nullptr, a.hint_deg));
        }
        ProfScope prof("gemm_h", s);
        if (a.sign_path) {
            a.wpr = (h + 31) / 32;
            LT_TRY(scratch.alloc(&a.signs, (size_t)R * a.wpr));
            LT_TRY(scratch.alloc(&a.minabs, (size_t)R * 2));                  // [0,R) min |H[r,:]|, [R,2R) max |H[r,:]|
            LT_CUDA(cudaMemsetAsync(a.minabs, 0x7f, (size_t)R * 4, s));
            LT_CUDA(cudaMemsetAsync(a.minabs + R, 0, (size_t)R * 4, s));
            LT_TRY(gemm_3xtf32_signs(a.AX, f, a.W1, h, a.b1, R, h, f, a.signs, a.minabs, s));
        } else {
            LT_TRY(scratch.alloc(&a.HR, (size_t)R * h));
            LT_TRY(gemm_f32(a.AX, f, nullptr, a.W1, h, 0, a.b1, 0, a.HR, h, R, h, f, 0, s));
        }
    } else {
        // H_R = Â[R,:]·(X·W1) + b1
        float* Pfull = nullptr;
        LT_TRY(scratch.alloc(&Pfull, (size_t)n * h));
        LT_TRY(scratch.alloc(&a.HR, (size_t)R * h));
        {
            ProfScope prof("gemm_p", s);
            LT_TRY(gemm_f32(a.X, a.ldx, nullptr, a.W1, h, 0, nullptr, 0, Pfull, h, n, h, f, 0, s));
        }
        ProfScope prof("spmm_h", s);
        LT_TRY(spmm_csr_f32_seg(a.indptr, a.indices, a.values, n, plan->rows, R, Pfull, h, h, a.b1, 0, a.HR, h, plan->seg_off,
                                plan->total_seg, s));
    }
    return LT_OK;
}
}  // namespace

struct ListedPairs { const int32_t* pair_j; const int32_t* pair_obs; int64_t K; float* out; };

static int pair_reduce(const lt_plan* p, const int32_t* indptr, const int32_t* indices, const float* values,
                       const int32_t* v_nodes, const ListedPairs& lp, const float* dq, int32_t lddq, int32_t c, cudaStream_t s) {
    if (lp.K == 0) return LT_OK;
    if (c > 256) { set_error("lt_pair_influence: c = %d > 256 classes is not supported", c); return LT_ERR_ARG; }
    ProfScope prof("pair_reduce", s);
    k_pair_reduce<<<(unsigned)ceil_div(lp.K * 32, 256), 256, 0, s>>>(indptr, indices, values, v_nodes, p->s_indptr, lp.pair_j,
                                                                     lp.pair_obs, lp.K, dq, lddq, c, lp.out);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

static int attack_efficient_impl(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n, const float* X,
                                 int64_t ldx, int32_t f, const float* W1, const float* b1, const float* W2, int32_t h, int32_t c,
                                 const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, float delta, float* out, int64_t ldo,
                                 lt_comm* comm, int64_t* info_host, cudaStream_t s, const ListedPairs* listed = nullptr) {
    HostScope host_all("attack_total");
    lt_plan* plan = new lt_plan();
    plan->n = n; plan->T = T; plan->lo = lo; plan->hi = hi; plan->Ts = hi - lo;
    plan->listed_only = listed != nullptr;
    int st = LT_OK;
    {
        Scratch scratch(s);
        AttackCtx a{indptr, indices, values, n, X, ldx, f, W1, b1, h, c, test_nodes, lo, s, &scratch};
        st = scratch.alloc(&a.Pt, (size_t)(plan->Ts > 0 ? plan->Ts : 1) * h);
        struct EvGuard { cudaEvent_t* e; ~EvGuard() { if (*e) cudaEventDestroy(*e); } } hub_guard{&a.hub_ready};
        cudaEvent_t zero_done = nullptr;
        EvGuard zero_guard{&zero_done};
        OutDst dst = single_dst(out, ldo);
        if (st == LT_OK && comm) {
            // every rank clears ITS matrix (next to the plan, on the side stream) and tells the others; the reduce kernel
            // further down then stores this rank's entries into all of them (comm.cu)
            ++comm->epoch;
            for (int r = 0; r < comm->world; ++r) dst.base[r] = static_cast<float*>(comm->peer[r]);
            dst.n = comm->world; dst.row0 = lo; dst.ldo = T;
            cudaStream_t sd = overlap_enabled() ? side_stream() : nullptr;
            cudaStream_t run = sd ? sd : s;
            cudaEvent_t fork = nullptr;
            if (run != s) {
                if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(fork, s) != cudaSuccess ||
                    cudaStreamWaitEvent(run, fork, 0) != cudaSuccess) { set_error("lt_attack_efficient_comm: fork failed"); st = LT_ERR_CUDA; }
                if (fork) cudaEventDestroy(fork);
            }
            if (st == LT_OK && cudaMemsetAsync(comm->peer[comm->rank], 0, (size_t)T * T * 4, run) != cudaSuccess) {
                set_error("lt_attack_efficient_comm: cudaMemsetAsync failed"); st = LT_ERR_CUDA;
            }
            if (st == LT_OK) st = comm_signal(comm, 0, run);
            if (st == LT_OK && run != s &&
                (cudaEventCreateWithFlags(&zero_done, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(zero_done, run) != cudaSuccess)) {
                set_error("lt_attack_efficient_comm: event failed"); st = LT_ERR_CUDA;
            }
        }
        // P_v = X[v]·W1 for the slice (gathered rows; float64 accumulation, algo 3: ΔR = Â[r,v]·P[v] carries these values
        // straight into the result).  Nothing but the test nodes is needed: it runs next to the plan head.
        EvGuard pv_guard{&a.pv_ready};
        if (st == LT_OK && hi > lo) {
            cudaStream_t sd = overlap_enabled() ? side_stream() : nullptr;
            cudaStream_t run = sd ? sd : s;
            if (run != s) {
                cudaEvent_t fork = nullptr;
                if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(fork, s) != cudaSuccess ||
                    cudaStreamWaitEvent(run, fork, 0) != cudaSuccess) { set_error("lt_attack_efficient: fork failed"); st = LT_ERR_CUDA; }
                if (fork) cudaEventDestroy(fork);
            }
            if (st == LT_OK) {
                ProfScope prof("gemm_pv", run);
                st = gemm_f32(X, ldx, test_nodes + lo, W1, h, 0, nullptr, 0, a.Pt, h, plan->Ts, h, f, 3, run);
            }
            if (st == LT_OK && run != s &&
                (cudaEventCreateWithFlags(&a.pv_ready, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(a.pv_ready, run) != cudaSuccess)) {
                set_error("lt_attack_efficient: event failed"); st = LT_ERR_CUDA;
            }
        }
        // eviction-priority hints of the aggregate-first SpMM: the degree threshold (remembered per graph: no device work
        // after the first attack on it)
        if (st == LT_OK && hi > lo && f < h && T >= 1024) a.hint_deg = spmm_hint_degree(indptr, n, X, ldx, f, s);
        // hub panel of X for the aggregate-first SpMM: built next to the plan head, on the side stream
        if (st == LT_OK && hi > lo && f < h && hub_applicable(n, ldx, f, X) && T >= 1024) {
            cudaStream_t sd = overlap_enabled() ? side_stream() : nullptr;
            cudaStream_t run = sd ? sd : s;
            st = hub_prepare(indptr, n, X, ldx, f, &a.hub, scratch, run);
            if (st == LT_OK && a.hub.valid && run != s) {
                if (cudaEventCreateWithFlags(&a.hub_ready, cudaEventDisableTiming) != cudaSuccess ||
                    cudaEventRecord(a.hub_ready, run) != cudaSuccess) {
                    set_error("lt_attack_efficient: hub panel event failed");
                    st = LT_ERR_CUDA;
                }
            }
        }
        int32_t n_fixed = 0;
        const bool work = hi > lo;
        // several ranks with room for the sign records: shard the restricted precompute by frontier rows.  Decided from
        // quantities every rank sees alike (shapes, alignment, options), so all ranks take the same branch.
        a.comm = comm;
        a.shard = comm && comm->world > 1 && comm->aux_bytes > 0 && opt(OPT_SHARD_ROWS) != 0 && !listed && sign_path_enabled() &&
                  f < h && f % 4 == 0 && ldx % 4 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)W1 % 16 == 0 && h >= 16 && c >= 8 &&
                  c <= 256 && T >= 8 * comm->world;
        plan->mark_all = a.shard;
        struct ForceTc { bool on; explicit ForceTc(bool v) : on(v) { if (on) g_force_tc = true; } ~ForceTc() { if (on) g_force_tc = false; } } force_tc(a.shard);
        if (st == LT_OK && work)
            st = plan_build(plan, indptr, indices, values, test_nodes, s, overlap_enabled() ? attack_early : nullptr, &a);
        if (st == LT_OK && work && !overlap_enabled()) st = attack_early(plan, &a, s);
        if (st == LT_OK && a.pv_ready && cudaStreamWaitEvent(s, a.pv_ready, 0) != cudaSuccess) {
            set_error("lt_attack_efficient: wait failed"); st = LT_ERR_CUDA;
        }
        const FixCtx fix{indptr, indices, values, X, ldx, f, W1, b1, test_nodes + lo, plan->rows};
        // several GPUs: nobody's matrix may still be clearing when the reduce kernel starts storing into it
        auto before_reduce = [&]() -> int {
            if (!comm) return LT_OK;
            if (zero_done && cudaStreamWaitEvent(s, zero_done, 0) != cudaSuccess) { set_error("lt_attack_efficient_comm: wait failed"); return LT_ERR_CUDA; }
            ProfScope prof("comm_wait_zeroed", s);
            return comm_wait(comm, 0, s);
        };
        if (st == LT_OK && work) {
            if (a.sign_path) {
                const int32_t lddq = frontier_lddq(plan->E, h, c);
                float* dq = nullptr;
                if (!comm && !listed && cudaMemset2DAsync(out, (size_t)ldo * 4, 0, (size_t)T * 4, (size_t)plan->Ts, s) != cudaSuccess) {
                    set_error("lt_attack_efficient: cudaMemset2DAsync failed");
                    st = LT_ERR_CUDA;
                }
                if (st == LT_OK && plan->E > 0 && (plan->M > 0 || listed)) {
                    st = scratch.alloc(&dq, (size_t)plan->E * lddq);
                    if (st == LT_OK)
                        st = frontier_dq_signs(plan, a.signs, a.wpr, a.minabs, fix, a.Pt, h, nullptr, W2, h, c, delta, dq, &n_fixed, s);
                    if (st == LT_OK) st = before_reduce();
                    if (st == LT_OK && !listed)
                        st = frontier_reduce(plan->rec_key, plan->rec_val, plan->grp_off, plan->n_groups, plan->big, plan->n_big,
                                             plan->M, dq, lddq, c, dst, s, plan->n_groups_host);
                    if (st == LT_OK && listed) st = pair_reduce(plan, indptr, indices, values, test_nodes + lo, *listed, dq, lddq, c, s);
                } else if (st == LT_OK) {
                    st = before_reduce();
                }
            } else if (listed) {                                    // value path: ΔQ from the stored H rows, then the pair reduce
                const int32_t lddq = frontier_lddq(plan->E, h, c);
                float* dq = nullptr;
                if (plan->E > 0) {
                    st = scratch.alloc(&dq, (size_t)plan->E * lddq);
                    if (st == LT_OK) st = frontier_dq(plan, plan->s_rho, a.HR, h, a.Pt, h, nullptr, W2, h, c, delta, dq, s, &fix, &n_fixed);
                    if (st == LT_OK) st = pair_reduce(plan, indptr, indices, values, test_nodes + lo, *listed, dq, lddq, c, s);
                }
            } else {
                st = before_reduce();
                if (st == LT_OK) st = influence_rows(plan, a.HR, h, a.Pt, h, nullptr, W2, h, c, delta, out, ldo, s, &fix, &n_fixed, comm ? &dst : nullptr);
            }
        } else if (st == LT_OK) {
            st = before_reduce();
        }
        if (st == LT_OK && comm) {
            // this rank's entries are in every matrix; the local one is complete once every peer says the same
            ProfScope prof("comm_exchange", s);
            st = comm_signal(comm, 1, s);
            if (st == LT_OK) st = comm_wait(comm, 1, s);
        }
        if (st == LT_OK && info_host) {
            if (work && !listed) st = lt_plan_info(plan, h, c, info_host, (lt_stream_t)s);
            else for (int k = 0; k < 8; ++k) info_host[k] = 0;
            if (st == LT_OK) info_host[7] = n_fixed;      // incidences re-evaluated in float64 (a ReLU within reach of a switch)
        }
    }
    lt_plan_destroy(plan, (lt_stream_t)s);
    return st;
}

extern "C" int lt_attack_efficient(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                                   const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                                   const float* W2, int32_t h, int32_t c, const int32_t* test_nodes, int32_t T,
                                   int32_t lo, int32_t hi, float delta, float* out, int64_t ldo, int64_t* info_host,
                                   lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && test_nodes, "lt_attack_efficient: null pointer");
    LT_CHECK_ARG(X && W1 && b1 && W2 && out, "lt_attack_efficient: null pointer");
    LT_CHECK_ARG(f > 0 && h > 0 && c > 0 && ldx >= f && h % 4 == 0, "lt_attack_efficient: sizes (h % 4 == 0)");
    LT_CHECK_ARG(n > 0 && T > 0 && lo >= 0 && hi >= lo && hi <= T && ldo >= T,
                 "lt_attack_efficient: need 0 <= lo <= hi <= T <= ldo");
    LT_CHECK_ARG(delta != 0.f, "lt_attack_efficient: delta must not be 0");
    if (hi == lo) return LT_OK;
    return attack_efficient_impl(indptr, indices, values, n, X, ldx, f, W1, b1, W2, h, c, test_nodes, T, lo, hi, delta, out, ldo,
                                 nullptr, info_host, (cudaStream_t)stream);
}

// Listed pairs: out[k] = ‖((f(X + δ·e_p x_pᵀ) − f(X))/δ)[o_k]‖₂ for p = test_nodes[pair_j[k]], o_k = pair_obs[k] (any node) — what
// link_prediction_attack_efficient_balanced (attacker.py:250-284) and the naive attack (attacker.py:143-160, 87-97) read one
// entry at a time.  test_nodes: the T distinct perturbed nodes.  Same restricted precompute, F1 and float64 fix-up as
// lt_attack_efficient; the reduce intersects N̂(p) and N̂(o) per pair, so no T x n row block is ever formed.
extern "C" int lt_pair_influence(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n, const float* X,
                                 int64_t ldx, int32_t f, const float* W1, const float* b1, const float* W2, int32_t h, int32_t c,
                                 const int32_t* test_nodes, int32_t T, const int32_t* pair_j, const int32_t* pair_obs, int64_t K,
                                 float delta, float* out, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && test_nodes && X && W1 && b1 && W2, "lt_pair_influence: null pointer");
    LT_CHECK_ARG(K == 0 || (pair_j && pair_obs && out), "lt_pair_influence: null pair arrays");
    LT_CHECK_ARG(f > 0 && h > 0 && c > 0 && ldx >= f && h % 4 == 0 && n > 0 && T > 0 && K >= 0 && delta != 0.f,
                 "lt_pair_influence: sizes (h % 4 == 0)");
    if (K == 0) return LT_OK;
    const ListedPairs lp{pair_j, pair_obs, K, out};
    return attack_efficient_impl(indptr, indices, values, n, X, ldx, f, W1, b1, W2, h, c, test_nodes, T, 0, T, delta, nullptr, T,
                                 nullptr, nullptr, (cudaStream_t)stream, &lp);
}

// The same attack on one rank of a node-wide run: this rank perturbs test_nodes[lo:hi) and stores its influence entries
// into the T x T matrix of EVERY rank of `comm` (peer memory over NVLink, comm.cu); on return (stream order) the local
// matrix lt_comm_buffer(comm) holds all T rows.  Every rank of the communicator must make the call, also with hi == lo.
extern "C" int lt_attack_efficient_comm(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                                        const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                                        const float* W2, int32_t h, int32_t c, const int32_t* test_nodes, int32_t T,
                                        int32_t lo, int32_t hi, float delta, lt_comm* comm, int64_t* info_host,
                                        lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && test_nodes && comm, "lt_attack_efficient_comm: null pointer");
    LT_CHECK_ARG(X && W1 && b1 && W2, "lt_attack_efficient_comm: null pointer");
    LT_CHECK_ARG(f > 0 && h > 0 && c > 0 && ldx >= f && h % 4 == 0, "lt_attack_efficient_comm: sizes (h % 4 == 0)");
    LT_CHECK_ARG(n > 0 && T > 0 && lo >= 0 && hi >= lo && hi <= T, "lt_attack_efficient_comm: need 0 <= lo <= hi <= T");
    LT_CHECK_ARG(delta != 0.f, "lt_attack_efficient_comm: delta must not be 0");
    LT_CHECK_ARG(comm->opened && (size_t)T * T * 4 <= comm->bytes, "lt_attack_efficient_comm: communicator not opened or its buffer is smaller than T*T floats");
    float* local = static_cast<float*>(comm->peer[comm->rank]);
    return attack_efficient_impl(indptr, indices, values, n, X, ldx, f, W1, b1, W2, h, c, test_nodes, T, lo, hi, delta,
                                 local + (int64_t)lo * T, T, comm, info_host, (cudaStream_t)stream);
}

// --n-layer 3: the same call for GCN3 (gcn/models.py:28-46)
extern "C" int lt_attack_efficient3(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                                    const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                                    const float* W2, const float* b2, const float* W3, int32_t h1, int32_t h2, int32_t c,
                                    const int32_t* test_nodes, int32_t T, int32_t lo, int32_t hi, float delta, float* out,
                                    int64_t ldo, int64_t max_chunk_records, int64_t* info_host, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && X && W1 && b1 && W2 && b2 && W3 && test_nodes && out,
                 "lt_attack_efficient3: null pointer");
    LT_CHECK_ARG(n > 0 && f > 0 && h1 > 0 && h2 > 0 && c > 0 && ldx >= f && h1 % 4 == 0 && T > 0 && lo >= 0 && hi >= lo &&
                     hi <= T && ldo >= T && delta != 0.f,
                 "lt_attack_efficient3: sizes (h1 % 4 == 0, 0 <= lo <= hi <= T <= ldo)");
    if (hi == lo) return LT_OK;
    return attack_efficient3(indptr, indices, values, n, X, ldx, f, W1, b1, W2, b2, W3, h1, h2, c, test_nodes, T, lo, hi,
                             delta, out, ldo, max_chunk_records, info_host, (cudaStream_t)stream);
}
