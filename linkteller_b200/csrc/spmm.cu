// CSR SpMM  C[i,:] = Σ_k Â[r_i,k]·B[k,:] (+bias) (ReLU)  — torch.spmm in GraphConvolution.forward
// (gcn/layers.py:32-34) and the restricted precompute H_R = Â[R,:]·P + b1 of the attack.
//
// HBM/L2-bound gather.  One warp per row SEGMENT (≤ SEG non-zeros), so hub rows of a scale-free
// graph are split across warps and the hardware CTA scheduler balances the load; rows that span
// several segments are reduced in segment order by a second tiny kernel (deterministic, no float
// atomics).  Lanes own float4 column slices (coalesced 16 B gathers of each B row); indices and
// values are read 32 at a time and broadcast by shuffle; 4 gathers in flight per lane.
#include "common.cuh"

#include <stdlib.h>

namespace lt {

constexpr int SPMM_SEG = 256;      // non-zeros per warp work item
constexpr int SPMM_THREADS = 256;  // 8 warps per CTA

__global__ void k_spmm_segcount(const int32_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                                int32_t n_rows, int32_t seg, int32_t* __restrict__ nseg) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int32_t r = rows ? rows[i] : (int32_t)i;
    const int32_t len = indptr[r + 1] - indptr[r];
    nseg[i] = len <= seg ? 1 : (len + seg - 1) / seg;
}

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
    a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y); a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
}

// NV = float4 slices per lane (d ≤ 128·NV).  seg_off: exclusive scan of nseg (n_rows+1 entries).
template <int NV, int GIF>   // GIF = gathers in flight per lane
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
        for (; k + GIF <= cnt; k += GIF) {
            float4 x[GIF][NV];
            float v[GIF];
#pragma unroll
            for (int u = 0; u < GIF; ++u) {
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
            for (int u = 0; u < GIF; ++u)
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

// ------------------------------------------------------------------------------------ panel-blocked variant
// When the dense operand does not fit the L2 (C5: X is 0.98 GB against 126 MB), the plain kernel
// re-fetches every gathered row from DRAM (ncu r1a: 13.8 GB of traffic for 1.4 GB of compulsory
// bytes, L2 hit rate 10 %).  Here the columns are cut into panels whose B rows fit the L2 several
// times over, and all resident warps walk the panels in the same order, kept within a window of
// SPMM_WINDOW panels of each other (per-CTA progress counters in global memory; leaders wait,
// laggards never do — the CTAs are co-resident: cooperative launch, one CTA per SM).  Each warp owns
// 32 row segments (lane ↔ cursor into its segment; CSR rows are column-sorted, so a panel is a
// contiguous run) whose accumulators live in shared memory (32 x d floats per warp).  Per panel a warp
//   1. advances the 32 cursors (4-entry probes, then a binary search for long runs),
//   2. flattens the runs into a (column | item, value) list in shared memory — load-balanced: list
//      slot → owner lane by a binary search over the scanned run lengths,
//   3. gathers the listed B rows cooperatively — lane ↔ float4 column slice, 16 gathers in flight —
//      and adds them into the item's accumulator row.
// Segments are short (SPMM_PSEG) and dealt round-robin to the CTAs so that the consecutive segments of a
// hub row (which all fall into one or two panels) do not pile up in one warp.  Rows are written out
// after the last panel.  A B panel is thus read from DRAM once per wave of resident rows instead of
// once per non-zero (ncu r1b: 13.8 GB → 5.1 GB, L2 hit rate 10 % → 55 %).
constexpr int SPMM_PW_MAX = 16;                 // warps per CTA (fewer when d is large)
constexpr int SPMM_LIST = 128;                  // flattened entries per warp per round
constexpr int SPMM_ITEM_SHIFT = 27;             // column < 2^27, item (0..31) above it
constexpr int SPMM_PSEG = 64;                   // non-zeros per row segment in this kernel
constexpr int SPMM_WINDOW = 2;                  // panels a CTA may run ahead of the slowest one
constexpr int SPMM_GIF = 16;                    // gathers in flight per warp

__global__ void __launch_bounds__(SPMM_PW_MAX * 32, 1)
k_spmm_panel(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
             const int32_t* __restrict__ rows, int32_t n_rows, const int32_t* __restrict__ seg_off,
             const float4* __restrict__ B4, int64_t ldb4, int32_t d4, const float* __restrict__ bias, int relu,
             float* __restrict__ C, int64_t ldc, float* __restrict__ partial, int32_t n_cols, int32_t panel_w,
             int32_t n_panels, int* __restrict__ progress) {
    extern __shared__ __align__(16) float smem_f[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int items_cta = nw * 32;
    const int32_t total = seg_off[n_rows];
    const int d = d4 * 4;
    float4* my_acc = reinterpret_cast<float4*>(smem_f + (size_t)warp * 32 * d);           // this warp's 32 rows
    uint2* my_list = reinterpret_cast<uint2*>(smem_f + (size_t)nw * 32 * d) + warp * SPMM_LIST;
    const int n_batches = (total + items_cta - 1) / items_cta;

    int done = 0;                                // panel phases this CTA has completed
    for (int batch = blockIdx.x; batch < n_batches; batch += gridDim.x) {
        // ---- lane ↔ one row segment; segment g of the global list goes to batch g % n_batches
        const int32_t g = (warp * 32 + lane) * n_batches + batch;
        int32_t i = 0, nseg = 0, cur = 0, s1 = 0;
        if (g < total) {
            i = (int32_t)upper_bound(seg_off, 0, (int64_t)n_rows + 1, g) - 1;
            const int32_t r = rows ? rows[i] : i;
            const int32_t first = seg_off[i];
            nseg = seg_off[i + 1] - first;
            cur = indptr[r] + (g - first) * SPMM_PSEG;
            s1 = min(indptr[r + 1], cur + SPMM_PSEG);
        }
        for (int k = lane; k < 32 * d4; k += 32) my_acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();

        for (int pi = 0; pi < n_panels; ++pi) {
            // ---- stay within SPMM_WINDOW panels of the slowest CTA
            if (warp == 0) {
                const int need = done - SPMM_WINDOW;
                if (need > 0) {
                    for (;;) {
                        int mn = INT32_MAX;
                        for (int c = lane; c < (int)gridDim.x; c += 32)
                            mn = min(mn, *reinterpret_cast<volatile int*>(progress + c));
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(LT_FULL, mn, o));
                        if (mn >= need) break;
                    }
                }
            }
            __syncthreads();
            const bool last_panel = pi == n_panels - 1;
            const int32_t lim = last_panel ? INT32_MAX : (pi + 1) * panel_w;
            // 1. advance this lane's cursor to the end of the panel
            int32_t nxt = cur;
            if (last_panel) {
                nxt = s1;
            } else {
                int probes = 0;
                while (nxt < s1) {
                    const int32_t w0 = indices[nxt];
                    const int32_t w1 = nxt + 1 < s1 ? indices[nxt + 1] : INT32_MAX;
                    const int32_t w2 = nxt + 2 < s1 ? indices[nxt + 2] : INT32_MAX;
                    const int32_t w3 = nxt + 3 < s1 ? indices[nxt + 3] : INT32_MAX;
                    const int c = (w0 < lim) + (w1 < lim) + (w2 < lim) + (w3 < lim);
                    nxt += c;
                    if (c < 4) break;
                    if (++probes == 2) {
                        nxt = (int32_t)lower_bound(indices, (int64_t)nxt, (int64_t)s1, lim);
                        break;
                    }
                }
            }
            const int32_t cnt = nxt - cur;
            // 2. exclusive warp scan of the run lengths
            int32_t inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t u = __shfl_up_sync(LT_FULL, inc, o);
                if (lane >= o) inc += u;
            }
            const int32_t off = inc - cnt;
            const int32_t tot = __shfl_sync(LT_FULL, inc, 31);
            for (int32_t r0 = 0; r0 < tot; r0 += SPMM_LIST) {     // rounds of SPMM_LIST entries
                const int32_t m = min(SPMM_LIST, tot - r0);
                for (int32_t t0 = 0; t0 < m; t0 += 32) {
                    const int32_t slot = r0 + t0 + lane;
                    int owner = 0;
#pragma unroll
                    for (int step = 16; step > 0; step >>= 1) {
                        const int32_t v = __shfl_sync(LT_FULL, inc, owner + step - 1);
                        if (v <= slot) owner += step;
                    }
                    owner = min(owner, 31);
                    const int32_t e = __shfl_sync(LT_FULL, cur, owner) + (slot - __shfl_sync(LT_FULL, off, owner));
                    if (t0 + lane < m)
                        my_list[t0 + lane] = make_uint2((uint32_t)indices[e] | ((uint32_t)owner << SPMM_ITEM_SHIFT),
                                                        __float_as_uint(values[e]));
                }
                __syncwarp();
                // 3. cooperative gathers
                int32_t t = 0;
                for (; t + SPMM_GIF <= m; t += SPMM_GIF) {
                    uint2 en[SPMM_GIF];
                    float4 x[SPMM_GIF];
#pragma unroll
                    for (int u = 0; u < SPMM_GIF; ++u) en[u] = my_list[t + u];
#pragma unroll
                    for (int u = 0; u < SPMM_GIF; ++u) {
                        const int64_t col = en[u].x & ((1u << SPMM_ITEM_SHIFT) - 1u);
                        x[u] = lane < d4 ? __ldg(B4 + col * ldb4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (lane < d4) {
#pragma unroll
                        for (int u = 0; u < SPMM_GIF; ++u) {
                            float4* p = my_acc + (size_t)(en[u].x >> SPMM_ITEM_SHIFT) * d4 + lane;
                            float4 a = *p;
                            fma4(a, __uint_as_float(en[u].y), x[u]);
                            *p = a;
                        }
                    }
                }
                for (; t < m; t += 4) {
                    uint2 en[4];
                    float4 x[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) en[u] = t + u < m ? my_list[t + u] : make_uint2(0u, 0u);
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int64_t col = en[u].x & ((1u << SPMM_ITEM_SHIFT) - 1u);
                        x[u] = (lane < d4 && t + u < m) ? __ldg(B4 + col * ldb4 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (lane < d4) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (t + u < m) {
                                float4* p = my_acc + (size_t)(en[u].x >> SPMM_ITEM_SHIFT) * d4 + lane;
                                float4 a = *p;
                                fma4(a, __uint_as_float(en[u].y), x[u]);
                                *p = a;
                            }
                        }
                    }
                }
                __syncwarp();
            }
            cur = nxt;
            // ---- publish progress (after every warp of the CTA has finished the panel)
            __syncthreads();
            ++done;
            if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(progress + blockIdx.x) = done;
        }
        // ---- write the 32 rows out (single-segment rows get bias / ReLU here, the others go to `partial`)
        for (int l0 = 0; l0 < 32; ++l0) {
            const int32_t g0 = (warp * 32 + l0) * n_batches + batch;
            if (g0 >= total) break;
            const int32_t i0 = __shfl_sync(LT_FULL, i, l0);
            const int32_t ns0 = __shfl_sync(LT_FULL, nseg, l0);
            if (lane < d4) {
                float4 a = my_acc[(size_t)l0 * d4 + lane];
                if (ns0 == 1) {
                    if (bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + lane);
                        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
                    }
                    if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                    reinterpret_cast<float4*>(C + (int64_t)i0 * ldc)[lane] = a;
                } else {
                    reinterpret_cast<float4*>(partial)[(int64_t)g0 * d4 + lane] = a;
                }
            }
        }
        __syncwarp();
    }
    // a CTA that has run out of batches must not hold the others back
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile int*>(progress + blockIdx.x) = INT32_MAX;
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

int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                 const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                 float* C, int64_t ldc, cudaStream_t s) {
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
    // dense operand far beyond the L2 and enough rows to fill the machine: panel-blocked kernel
    const int32_t d4 = d / 4;
    const int64_t b_bytes = (int64_t)n_cols * d * 4;
    const char* pm = getenv("LT_SPMM_PANEL_MB");
    const int64_t panel_bytes = (int64_t)(pm ? atoi(pm) : 32) << 20;
    int nw = (int)((232448 - 1024) / (32 * (size_t)d * 4 + SPMM_LIST * 8));
    if (nw > SPMM_PW_MAX) nw = SPMM_PW_MAX;
    const bool use_panel = d4 <= 32 && nw >= 4 && n_cols < (1 << SPMM_ITEM_SHIFT) && b_bytes >= ((int64_t)256 << 20) &&
                           n_rows >= 64 * 512 && getenv("LT_SPMM_PANEL") && !getenv("LT_SPMM_NO_PANEL");
    Scratch scratch(s);
    int32_t* seg_off;
    LT_TRY(scratch.alloc(&seg_off, (size_t)n_rows + 1));
    k_spmm_segcount<<<(unsigned)ceil_div(n_rows, 256), 256, 0, s>>>(indptr, rows, n_rows, use_panel ? SPMM_PSEG : SPMM_SEG,
                                                                   seg_off);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(seg_off, seg_off, n_rows, s));
    // the number of segments sizes the grid and the partial-sum buffer: 4-byte copy, one sync per SpMM
    int32_t nnz_bound = 0;
    LT_CUDA(cudaMemcpyAsync(&nnz_bound, seg_off + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    const int64_t total_seg = nnz_bound;
    float* partial = nullptr;
    const bool multi = total_seg > n_rows;
    if (multi) LT_TRY(scratch.alloc(&partial, (size_t)total_seg * d));
    const float4* B4 = reinterpret_cast<const float4*>(B);
    if (use_panel) {
        int32_t n_panels = (int32_t)ceil_div(b_bytes, panel_bytes);
        const int32_t panel_w = (int32_t)ceil_div(n_cols, n_panels);
        n_panels = (int32_t)ceil_div(n_cols, panel_w);
        const size_t smem = (size_t)nw * (32 * (size_t)d * 4 + SPMM_LIST * 8);
        LT_CUDA(cudaFuncSetAttribute(k_spmm_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)min((int64_t)num_sms(), ceil_div(total_seg, nw * 32));
        int* progress;
        LT_TRY(scratch.alloc(&progress, (size_t)grid));
        LT_CUDA(cudaMemsetAsync(progress, 0, sizeof(int) * grid, s));
        const int64_t ldb4 = ldb / 4;
        void* args[] = {(void*)&indptr, (void*)&indices, (void*)&values, (void*)&rows, (void*)&n_rows, (void*)&seg_off,
                        (void*)&B4, (void*)&ldb4, (void*)&d4, (void*)&bias, (void*)&relu, (void*)&C, (void*)&ldc,
                        (void*)&partial, (void*)&n_cols, (void*)&panel_w, (void*)&n_panels, (void*)&progress};
        LT_CUDA(cudaLaunchCooperativeKernel((const void*)k_spmm_panel, dim3(grid), dim3(nw * 32), args, smem, s));
        LT_LAUNCH_CHECK();
    } else {
        const unsigned bl = (unsigned)ceil_div(total_seg * 32, SPMM_THREADS);
#define LT_SPMM_V(NV_, GIF_) k_spmm_vec<NV_, GIF_><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, n_rows, seg_off, B4, ldb / 4, d4, bias, relu, C, ldc, partial)
        const char* gf = getenv("LT_SPMM_GIF");
        const int gif = gf ? atoi(gf) : 8;
        if (d4 <= 32) { if (gif >= 16) LT_SPMM_V(1, 16); else if (gif >= 8) LT_SPMM_V(1, 8); else LT_SPMM_V(1, 4); }
        else if (d4 <= 64) LT_SPMM_V(2, 4);
        else if (d4 <= 96) LT_SPMM_V(3, 4);
        else LT_SPMM_V(4, 4);
#undef LT_SPMM_V
        LT_LAUNCH_CHECK();
    }
    if (multi) {
        k_spmm_fixup<<<(unsigned)ceil_div((int64_t)n_rows * 32, 256), 256, 0, s>>>(seg_off, n_rows, partial, d, bias,
                                                                                 relu, C, ldc);
        LT_LAUNCH_CHECK();
    }
    return LT_OK;
}

}  // namespace lt

extern "C" int lt_spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                               const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d,
                               const float* bias, int relu, float* C, int64_t ldc, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && B && C && n > 0 && n_rows >= 0 && d >= 0 && ldb >= d && ldc >= d,
                 "lt_spmm_csr_f32");
    return lt::spmm_csr_f32(indptr, indices, values, n, rows, n_rows, B, ldb, d, bias, relu, C, ldc,
                            (cudaStream_t)stream);
}
