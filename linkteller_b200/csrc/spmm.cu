// CSR SpMM  C[i,:] = Σ_k Â[r_i,k]·B[k,:] (+bias) (ReLU)  — torch.spmm in GraphConvolution.forward
// (gcn/layers.py:32-34) and the restricted precompute H_R = Â[R,:]·P + b1 of the attack.
//
// HBM/L2-bound gather.  One warp per row SEGMENT (≤ SEG non-zeros), so hub rows of a scale-free
// graph are split across warps and the hardware CTA scheduler balances the load; rows that span
// several segments are reduced in segment order by a second tiny kernel (deterministic, no float
// atomics).  Lanes own float4 column slices (coalesced 16 B gathers of each B row); indices and
// values are read 32 at a time, one chunk ahead, and broadcast by shuffle; 8 gathers in flight per lane.
#include "common.cuh"

#include <stdlib.h>

namespace lt {

constexpr int SPMM_THREADS = 256;  // 8 warps per CTA

// Row lists.  A product runs over list positions t = 0..n_list-1; position t stands for output row
// il = sel ? sel[t] : t, which is matrix row r = rows ? rows[il] : il.  `sel` lets the panel path
// split one row list into a short-row and a long-row pass without copying anything.
__device__ __forceinline__ int32_t list_pos(const int32_t* __restrict__ sel, int64_t t) { return sel ? sel[t] : (int32_t)t; }

// nseg[t] for t < *n_list (n_list_dev may be null: all n_cap positions are live), 0 beyond
__global__ void k_spmm_segcount(const int32_t* __restrict__ indptr, const int32_t* __restrict__ rows,
                                const int32_t* __restrict__ sel, const int32_t* __restrict__ n_list_dev, int32_t n_cap,
                                int32_t seg, int32_t* __restrict__ nseg) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cap) return;
    if (n_list_dev && t >= *n_list_dev) { nseg[t] = 0; return; }
    const int32_t il = list_pos(sel, t);
    const int32_t r = rows ? rows[il] : il;
    const int32_t len = indptr[r + 1] - indptr[r];
    nseg[t] = len <= seg ? 1 : (len + seg - 1) / seg;
}


// NV = float4 slices per lane (d ≤ 128·NV).  seg_off: exclusive scan of nseg (n_rows+1 entries).
// The (column, value) pairs of the NEXT 32 non-zeros are loaded before the gathers of the current 32 are
// issued, so a warp never waits for an index load with nothing else in flight.  Lanes keep the float4
// index of "their" B row (32 bits — the host checks n_cols·ldb4 < 2³²), so one gather costs a shuffle,
// a wide multiply-add and the load.
template <int NV, int GIF, int MINB>   // GIF = gathers in flight per lane, MINB = resident CTAs per SM asked of ptxas
__global__ void __launch_bounds__(SPMM_THREADS, MINB)
k_spmm_vec(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
           const int32_t* __restrict__ rows, const int32_t* __restrict__ sel, int32_t n_rows, const int32_t* __restrict__ seg_off,
           const float4* __restrict__ B4, uint32_t ldb4, int32_t d4, const float* __restrict__ bias, int relu,
           float* __restrict__ C, int64_t ldc, float* __restrict__ partial) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
    const int32_t total = seg_off[n_rows];
    if (g >= total) return;
    const int32_t i = (int32_t)upper_bound(seg_off, 0, (int64_t)n_rows + 1, (int32_t)g) - 1;
    const int32_t il = list_pos(sel, i);
    const int32_t r = rows ? rows[il] : il;
    const int32_t first = seg_off[i];
    const int32_t nseg = seg_off[i + 1] - first;
    const int32_t row_beg = indptr[r], row_end = indptr[r + 1];
    const int32_t s0 = row_beg + ((int32_t)g - first) * SPMM_SEG;
    const int32_t s1 = min(row_end, s0 + SPMM_SEG);

    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    // lanes past the row width re-load the row's last float4 (same sector, result unused): a load that is
    // predicated per lane needs a zero-fill of its destination, and that write waits for the load —
    // which serialises the gathers (ncu: the select after each LDG was the top long_scoreboard stall)
    const float4* Bl[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) Bl[q] = B4 + min(lane + 32 * q, d4 - 1);

    uint32_t mo = s0 + lane < s1 ? (uint32_t)indices[s0 + lane] * ldb4 : 0u;
    float mv = s0 + lane < s1 ? values[s0 + lane] : 0.f;
    for (int32_t base = s0; base < s1; base += 32) {
        const int32_t nx = base + 32 + lane;
        const uint32_t no = nx < s1 ? (uint32_t)indices[nx] * ldb4 : 0u;
        const float nv = nx < s1 ? values[nx] : 0.f;
        const int cnt = min(32, s1 - base);
        for (int k = 0; k < cnt; k += GIF) {
            // every load is unconditional and lands in its own register: entries past the end of the row
            // carry offset 0 (row 0 of B, an L1 hit) and are skipped by the warp-uniform guard below
            float4 x[GIF][NV];
#pragma unroll
            for (int u = 0; u < GIF; ++u) {
                const uint32_t o = __shfl_sync(LT_FULL, mo, (k + u) & 31);
#pragma unroll
                for (int q = 0; q < NV; ++q) x[u][q] = __ldg(Bl[q] + o);
            }
#pragma unroll
            for (int u = 0; u < GIF; ++u) {
                const float v = __shfl_sync(LT_FULL, mv, (k + u) & 31);
                if (k + u < cnt) {
#pragma unroll
                    for (int q = 0; q < NV; ++q) fma4(acc[q], v, x[u][q]);
                }
            }
        }
        mo = no;
        mv = nv;
    }

    if (nseg == 1) {
        float* dst = C + (int64_t)il * ldc;
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

// rows split over several segments: sum the partial rows in segment order, then bias / ReLU.
// One warp per row; a lane owns one float4 column slice (d ≤ 128: one pass) and keeps 8 loads in flight —
// a hub row has hundreds of segments, and one dependent load per segment made this the tail of the SpMM.
__global__ void __launch_bounds__(256)
k_spmm_fixup(const int32_t* __restrict__ seg_off, const int32_t* __restrict__ sel, int32_t n_rows,
             const float* __restrict__ partial, int32_t d, const float* __restrict__ bias, int relu,
             float* __restrict__ C, int64_t ldc) {
    // a CTA looks at 256 list positions with one thread each and queues the few that span several segments; its
    // warps then take the queued rows (a warp per position made 45 K CTAs of warps that exit at once on C5)
    __shared__ int32_t queue[256];
    __shared__ int32_t n_queued;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) n_queued = 0;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n_rows && seg_off[t + 1] - seg_off[t] > 1) queue[atomicAdd(&n_queued, 1)] = (int32_t)t;
    __syncthreads();
    const int d4 = d >> 2;                                       // the vector path guarantees d % 4 == 0
    for (int q = warp; q < n_queued; q += 8) {
        const int64_t i = queue[q];
        const int32_t first = seg_off[i], nseg = seg_off[i + 1] - first;
        const int64_t il = list_pos(sel, i);
        const float4* P4 = reinterpret_cast<const float4*>(partial) + (int64_t)first * d4;
        for (int col = lane; col < d4; col += 32) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            int32_t sgm = 0;
            for (; sgm + 8 <= nseg; sgm += 8) {
                float4 p[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) p[u] = P4[(int64_t)(sgm + u) * d4 + col];
#pragma unroll
                for (int u = 0; u < 8; ++u) { a.x += p[u].x; a.y += p[u].y; a.z += p[u].z; a.w += p[u].w; }
            }
            for (; sgm < nseg; ++sgm) {
                const float4 p = P4[(int64_t)sgm * d4 + col];
                a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
            }
            if (bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + col);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
            reinterpret_cast<float4*>(C + il * ldc)[col] = a;
        }
    }
}

// ------------------------------------------------------------------------------------ panel-blocked variant
// When the dense operand does not fit the L2 (C5: X is 0.98 GB against 126 MB), the plain kernel
// re-fetches every gathered row from DRAM (ncu r1b: 13.8 GB of traffic for 1.4 GB of compulsory
// bytes, L2 hit rate 10 %).  Here the columns are cut into panels whose B rows fit the L2 and the
// product is taken one panel per LAUNCH: stream order is the only synchronisation, every launch
// gathers from one L2-resident slice of B, and the per-segment partial sums travel through global
// memory between launches (read-modify-write of one accumulator row per non-empty piece, skipped
// for empty pieces, plain write for the first piece).  CSR rows are column-sorted, so the piece of
// a row segment that falls into a panel is a contiguous run; its bounds come from one streaming
// pass over the indices (k_pan_offsets).  A warp owns PAN_SLOTS segments per launch:
// the eight pieces are flattened into one entry list (lane ↔ entry) so that the gathers of short
// pieces are issued back to back, and the eight accumulator rows sit in shared memory.  The
// additions into one accumulator happen in ascending column order — the plain kernel's order — so
// the two kernels agree bit for bit.
//
// Measured on C5 (profiles/r1c_*): the panels do cut DRAM traffic (13.9 GB → ~7 GB) but the graph's
// rows are short (median 48 non-zeros), so a piece is a handful of non-zeros and the per-piece
// accumulator round trip plus the extra launches cost more than the saved gathers: 3.1 ms at best
// against 2.54 ms for the plain kernel, which already streams its traffic at 84 % of the DRAM copy peak.
// The path is therefore opt-in (LT_SPMM_PANEL=1; LT_SPMM_PANEL_MB / LT_SPMM_PANEL_LMIN tune it) and
// kept for graphs with long rows; the parity test covers it.
constexpr int PAN_SLOTS = 8;                    // row segments per warp
constexpr int PAN_THREADS = 256;                // 8 warps per CTA
constexpr int PAN_GIF = 8;                      // gathers in flight per lane
constexpr uint32_t PAN_FIRST = 0x80000000u;     // offset flag: nothing of the segment lies before it

// seg_row[g] = position (in the row list) of the row that owns segment g
__global__ void k_pan_segrow(const int32_t* __restrict__ seg_off, int32_t n_rows, int32_t* __restrict__ seg_row) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    for (int32_t g = seg_off[i]; g < seg_off[i + 1]; ++g) seg_row[g] = (int32_t)i;
}

// One warp per segment: poff[q][g] = first entry of segment g whose column is in panel ≥ q (q = 0..P),
// flagged PAN_FIRST when it is the segment's first entry; seg_dst[g] = float offset of its accumulator
// row (≥ 0: row of C, single-segment rows; < 0: ~offset into `partial`).  Empty rows get their result here.
__global__ void __launch_bounds__(256)
k_pan_offsets(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const int32_t* __restrict__ rows,
              const int32_t* __restrict__ sel, const int32_t* __restrict__ seg_off, const int32_t* __restrict__ seg_row, int32_t total_seg, int32_t seg_len,
              int32_t panel_w, int32_t n_panels, int32_t d, int64_t ldc, const float* __restrict__ bias, int relu,
              uint32_t* __restrict__ poff, int64_t* __restrict__ seg_dst, float* __restrict__ C) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= total_seg) return;
    const int32_t i = seg_row[g];
    const int32_t il = list_pos(sel, i);
    const int32_t r = rows ? rows[il] : il;
    const int32_t first = seg_off[i], nseg = seg_off[i + 1] - first;
    const int32_t s0 = indptr[r] + ((int32_t)g - first) * seg_len;
    const int32_t s1 = min(indptr[r + 1], s0 + seg_len);
    if (lane == 0) seg_dst[g] = nseg == 1 ? (int64_t)il * ldc : ~((int64_t)g * d);
    int32_t carry = -1;                                   // panel of the entry before this chunk
    for (int32_t base = s0; base < s1; base += 32) {
        const int32_t k = base + lane;
        const int32_t q = k < s1 ? indices[k] / panel_w : n_panels;   // past-the-end lanes close all panels
        int32_t qp = __shfl_up_sync(LT_FULL, q, 1);
        if (lane == 0) qp = carry;
        if (k <= s1) {                                    // lane k == s1 writes the tail offsets
            const uint32_t v = (uint32_t)k | (k == s0 ? PAN_FIRST : 0u);
            for (int32_t t = qp + 1; t <= q; ++t) poff[(int64_t)t * total_seg + g] = v;
        }
        carry = __shfl_sync(LT_FULL, q, 31);
    }
    if (carry < n_panels) {                               // segment length a multiple of 32 (or empty): tail offsets
        const uint32_t v = (uint32_t)s1 | (s1 == s0 ? PAN_FIRST : 0u);
        for (int32_t t = carry + 1 + lane; t <= n_panels; t += 32) poff[(int64_t)t * total_seg + g] = v;
    }
    if (s0 == s1 && nseg == 1)                            // empty row: no piece will ever write it
        for (int col = lane; col < d; col += 32) {
            float a = bias ? bias[col] : 0.f;
            if (relu) a = fmaxf(a, 0.f);
            C[(int64_t)il * ldc + col] = a;
        }
}

__global__ void __launch_bounds__(PAN_THREADS)
k_spmm_pan(const int32_t* __restrict__ indices, const float* __restrict__ values, const uint32_t* __restrict__ poff_lo,
           const uint32_t* __restrict__ poff_hi, int32_t total_seg, const int64_t* __restrict__ seg_dst,
           const float4* __restrict__ B4, int64_t ldb4, int32_t d4, float* C, float* partial) {
    __shared__ float4 acc_s[PAN_THREADS / 32][PAN_SLOTS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // segments are dealt round-robin (slot sl of warp w is segment w + sl·W): the consecutive full
    // segments of a hub row, which all fall into one or two panels, end up in different warps
    const int64_t W = ((int64_t)total_seg + PAN_SLOTS - 1) / PAN_SLOTS;
    const int64_t w = (int64_t)blockIdx.x * (PAN_THREADS / 32) + warp;
    if (w >= W) return;
    int32_t a = 0, cnt = 0;
    int first = 1;
    int64_t dst = 0;
    const int64_t gl = w + (int64_t)lane * W;
    if (lane < PAN_SLOTS && gl < total_seg) {
        const uint32_t ra = poff_lo[gl];
        const uint32_t rb = poff_hi[gl];
        dst = seg_dst[gl];
        a = (int32_t)(ra & ~PAN_FIRST);
        first = (int)(ra >> 31);
        cnt = (int32_t)(rb & ~PAN_FIRST) - a;
    }
    const int32_t incl = warp_inclusive_scan(cnt);
    const int32_t total = __shfl_sync(LT_FULL, incl, 31);
    if (total == 0) return;

    // accumulator rows of the non-empty pieces: asynchronous global → shared copies (no registers held),
    // waited for after the first index loads have been issued
#pragma unroll
    for (int sl = 0; sl < PAN_SLOTS; ++sl) {
        const int32_t c_s = __shfl_sync(LT_FULL, cnt, sl);
        const int fs = __shfl_sync(LT_FULL, first, sl);
        const int64_t ds = __shfl_sync(LT_FULL, dst, sl);
        if (c_s > 0 && lane < d4) {
            if (fs) {
                acc_s[warp][sl][lane] = make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                const float4* src = reinterpret_cast<const float4*>(ds >= 0 ? C + ds : partial + ~ds) + lane;
                const uint32_t sm = (uint32_t)__cvta_generic_to_shared(&acc_s[warp][sl][lane]);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm), "l"(src) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    bool staged = false;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);       // the one open accumulator row (slots are visited in order)
    for (int32_t base = 0; base < total; base += 32) {
        // lane ↔ entry base+lane of the flattened list: its slot and position in the CSR arrays
        const int32_t q = base + lane;
        int slot = 0;
#pragma unroll
        for (int sl = 0; sl < PAN_SLOTS - 1; ++sl) slot += (__shfl_sync(LT_FULL, incl, sl) <= q) ? 1 : 0;
        const int32_t in_q = __shfl_sync(LT_FULL, incl, slot), cn_q = __shfl_sync(LT_FULL, cnt, slot);
        const int32_t a_q = __shfl_sync(LT_FULL, a, slot);
        const int32_t pos = a_q + (q - (in_q - cn_q));
        // float4 index of the gathered row's first element (32 bits: B up to 64 GB), so that a gather costs
        // one shuffle, one wide multiply-add and the load
        const uint32_t mo = q < total ? (uint32_t)indices[pos] * (uint32_t)ldb4 : 0u;
        const float mv = q < total ? values[pos] : 0.f;
        const uint32_t starts = __ballot_sync(LT_FULL, q < total && q == in_q - cn_q);   // entry opens its slot's run
        const uint32_t ends = __ballot_sync(LT_FULL, q < total && q == in_q - 1);        // entry closes it
        if (!staged) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            staged = true;
            __syncwarp();
        }
        const int m = min(32, total - base);
        const float4* Bl = B4 + min(lane, d4 - 1);        // unconditional loads (see k_spmm_vec)
        for (int k = 0; k < m; k += PAN_GIF) {
            float4 x[PAN_GIF];
#pragma unroll
            for (int u = 0; u < PAN_GIF; ++u)            // unconditional; entries past m carry offset 0
                x[u] = __ldg(Bl + __shfl_sync(LT_FULL, mo, k + u));
#pragma unroll
            for (int u = 0; u < PAN_GIF; ++u) {
                const float v = __shfl_sync(LT_FULL, mv, k + u);
                if (k + u >= m) break;                    // warp-uniform
                if ((starts >> (k + u)) & 1u)             // warp-uniform: zero or the staged partial sum
                    acc = acc_s[warp][__shfl_sync(LT_FULL, slot, k + u)][lane];
                fma4(acc, v, x[u]);
                if ((ends >> (k + u)) & 1u) {             // warp-uniform: the row goes back to global memory
                    const int64_t ds = __shfl_sync(LT_FULL, dst, __shfl_sync(LT_FULL, slot, k + u));
                    if (lane < d4) reinterpret_cast<float4*>(ds >= 0 ? C + ds : partial + ~ds)[lane] = acc;
                }
            }
        }
    }
}

// bias / ReLU on the single-segment rows after the last panel (multi-segment rows: k_spmm_fixup)
__global__ void k_pan_epilogue(const int32_t* __restrict__ seg_off, const int32_t* __restrict__ sel, int32_t n_rows, int32_t d,
                               const float* __restrict__ bias, int relu, float* __restrict__ C, int64_t ldc) {
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_rows || seg_off[i + 1] - seg_off[i] != 1) return;
    const int64_t il = list_pos(sel, i);
    for (int col = lane; col < d; col += 32) {
        float a = C[il * ldc + col];
        if (bias) a += bias[col];
        if (relu) a = fmaxf(a, 0.f);
        C[il * ldc + col] = a;
    }
}

// split a row list by row length: long rows (≥ lmin non-zeros) take the panel path
__global__ void k_pan_classify(const int32_t* __restrict__ indptr, const int32_t* __restrict__ rows, int32_t n_rows,
                               int32_t lmin, int32_t* __restrict__ flag) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows) return;
    const int32_t r = rows ? rows[t] : (int32_t)t;
    flag[t] = (indptr[r + 1] - indptr[r] >= lmin) ? 1 : 0;
}
// pos = exclusive scan of flag: long rows → sel_long[pos], short rows → sel_short[t - pos]; counts[0..1] = {long, short}
__global__ void k_pan_split(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int32_t n_rows,
                            int32_t* __restrict__ sel_long, int32_t* __restrict__ sel_short, int32_t* __restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows) return;
    if (flag[t]) sel_long[pos[t]] = (int32_t)t; else sel_short[t - pos[t]] = (int32_t)t;
    if (t == 0) { counts[0] = pos[n_rows]; counts[1] = n_rows - pos[n_rows]; }
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

// seg_off_known / total_seg_known (optional): exclusive scan of the per-row segment counts (n_rows+1 entries,
// SPMM_SEG non-zeros per segment, at least one per row) and its total, when the caller already has them — the
// attack plan does — so that the product needs neither the counting pass nor its host synchronisation.
int spmm_csr_f32_seg(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                     const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                     float* C, int64_t ldc, const int32_t* seg_off_known, int64_t total_seg_known, cudaStream_t s);

int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                 const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                 float* C, int64_t ldc, cudaStream_t s) {
    return spmm_csr_f32_seg(indptr, indices, values, n_cols, rows, n_rows, B, ldb, d, bias, relu, C, ldc, nullptr, -1, s);
}

int spmm_csr_f32_seg(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                     const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                     float* C, int64_t ldc, const int32_t* seg_off_known, int64_t total_seg_known, cudaStream_t s) {
    if (n_rows == 0 || d == 0) return LT_OK;
    const bool vec = (d % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0) && d <= 512 && (int64_t)n_cols * (ldb / 4) < ((int64_t)1 << 32) &&
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
    const int32_t d4 = d / 4;
    const float4* B4 = reinterpret_cast<const float4*>(B);
    Scratch scratch(s);

    // the plain pass over a row list (sel == null: the whole list); total_seg < 0: not known yet (one 4-byte sync)
    auto plain_pass = [&](const int32_t* sel, int32_t n_list, const int32_t* seg_off_in, int64_t total_seg) -> int {
        int32_t* seg_off = const_cast<int32_t*>(seg_off_in);   // written only when the counts are not known yet
        if (n_list == 0) return LT_OK;
        if (total_seg < 0) {
            k_spmm_segcount<<<(unsigned)ceil_div(n_list, 256), 256, 0, s>>>(indptr, rows, sel, nullptr, n_list, SPMM_SEG, seg_off);
            LT_LAUNCH_CHECK();
            LT_TRY(exclusive_scan_i32(seg_off, seg_off, n_list, s));
            // the number of segments sizes the grid and the partial-sum buffer: 4-byte copy, one sync per SpMM
            int32_t h = 0;
            LT_CUDA(cudaMemcpyAsync(&h, seg_off + n_list, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
            LT_CUDA(cudaStreamSynchronize(s));
            total_seg = h;
        }
        float* partial = nullptr;
        const bool multi = total_seg > n_list;
        if (multi) LT_TRY(scratch.alloc(&partial, (size_t)total_seg * d));
        const unsigned bl = (unsigned)ceil_div(total_seg * 32, SPMM_THREADS);
#define LT_SPMM_V(NV_, GIF_, MINB_) k_spmm_vec<NV_, GIF_, MINB_><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4, (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial)
        // d ≤ 128: 8 gathers in flight per lane at 3 CTAs/SM (no spills) measured best on C5 (2.54 ms; 4 CTAs 2.64,
        // 16 in flight at 2 CTAs 3.16): the kernel then moves its 13.9 GB of DRAM traffic at 84 % of the copy peak
        if (d4 <= 32) LT_SPMM_V(1, 8, 3);
        else if (d4 <= 64) LT_SPMM_V(2, 4, 4);
        else if (d4 <= 96) LT_SPMM_V(3, 4, 3);
        else LT_SPMM_V(4, 4, 2);
#undef LT_SPMM_V
        LT_LAUNCH_CHECK();
        if (multi) {
            k_spmm_fixup<<<(unsigned)ceil_div((int64_t)n_list, 256), 256, 0, s>>>(seg_off, sel, n_list, partial, d, bias,
                                                                                     relu, C, ldc);
            LT_LAUNCH_CHECK();
        }
        return LT_OK;
    };

    // dense operand far beyond the L2 and enough rows to amortise the extra launches: long rows go through
    // one launch per column panel, short rows (a piece per panel would be a couple of non-zeros) stay plain
    const int64_t b_bytes = (int64_t)n_cols * d * 4;
    const int64_t panel_bytes = (int64_t)opt(OPT_SPMM_PANEL_MB) << 20;
    const int32_t lmin = opt(OPT_SPMM_PANEL_LMIN);
    const bool use_panel = d4 <= 32 && b_bytes >= ((int64_t)256 << 20) && n_rows >= 32768 && opt(OPT_SPMM_PANEL) != 0;
    if (!use_panel) {
        if (seg_off_known && total_seg_known >= 0) return plain_pass(nullptr, n_rows, seg_off_known, total_seg_known);
        int32_t* seg_off;
        LT_TRY(scratch.alloc(&seg_off, (size_t)n_rows + 1));
        return plain_pass(nullptr, n_rows, seg_off, -1);
    }
    int32_t *flag, *pos, *sel_long, *sel_short, *counts, *off_long, *off_short;
    LT_TRY(scratch.alloc(&flag, (size_t)n_rows + 1));
    LT_TRY(scratch.alloc(&pos, (size_t)n_rows + 1));
    LT_TRY(scratch.alloc(&sel_long, (size_t)n_rows));
    LT_TRY(scratch.alloc(&sel_short, (size_t)n_rows));
    LT_TRY(scratch.alloc(&counts, 2));
    LT_TRY(scratch.alloc(&off_long, (size_t)n_rows + 1));
    LT_TRY(scratch.alloc(&off_short, (size_t)n_rows + 1));
    const unsigned blr = (unsigned)ceil_div(n_rows, 256);
    k_pan_classify<<<blr, 256, 0, s>>>(indptr, rows, n_rows, lmin, flag);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(flag, pos, n_rows, s));
    k_pan_split<<<blr, 256, 0, s>>>(flag, pos, n_rows, sel_long, sel_short, counts);
    LT_LAUNCH_CHECK();
    k_spmm_segcount<<<blr, 256, 0, s>>>(indptr, rows, sel_long, counts, n_rows, SPMM_SEG, off_long);
    LT_LAUNCH_CHECK();
    k_spmm_segcount<<<blr, 256, 0, s>>>(indptr, rows, sel_short, counts + 1, n_rows, SPMM_SEG, off_short);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(off_long, off_long, n_rows, s));
    LT_TRY(exclusive_scan_i32(off_short, off_short, n_rows, s));
    int32_t hc[4] = {0, 0, 0, 0};   // long rows, short rows, long segments, short segments: one sync per SpMM
    LT_CUDA(cudaMemcpyAsync(&hc[0], counts, 8, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&hc[2], off_long + n_rows, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaMemcpyAsync(&hc[3], off_short + n_rows, 4, cudaMemcpyDeviceToHost, s));
    LT_CUDA(cudaStreamSynchronize(s));
    const int32_t n_long = hc[0], n_short = hc[1];
    const int64_t total_seg = hc[2];
    LT_TRY(plain_pass(sel_short, n_short, off_short, hc[3]));
    if (n_long == 0) return LT_OK;
    {
        float* partial = nullptr;
        const bool multi = total_seg > n_long;
        if (multi) LT_TRY(scratch.alloc(&partial, (size_t)total_seg * d));
        int32_t n_panels = (int32_t)ceil_div(b_bytes, panel_bytes);
        const int32_t panel_w = (int32_t)ceil_div(n_cols, n_panels);
        n_panels = (int32_t)ceil_div(n_cols, panel_w);
        int32_t* seg_row;
        uint32_t* poff;
        int64_t* seg_dst;
        LT_TRY(scratch.alloc(&seg_row, (size_t)total_seg));
        LT_TRY(scratch.alloc(&poff, (size_t)(n_panels + 1) * total_seg));
        LT_TRY(scratch.alloc(&seg_dst, (size_t)total_seg));
        k_pan_segrow<<<(unsigned)ceil_div(n_long, 256), 256, 0, s>>>(off_long, n_long, seg_row);
        LT_LAUNCH_CHECK();
        k_pan_offsets<<<(unsigned)ceil_div(total_seg * 32, 256), 256, 0, s>>>(indptr, indices, rows, sel_long, off_long, seg_row,
                                                                            (int32_t)total_seg, SPMM_SEG, panel_w, n_panels,
                                                                            d, ldc, bias, relu, poff, seg_dst, C);
        LT_LAUNCH_CHECK();
        const unsigned bl = (unsigned)ceil_div(total_seg, (PAN_THREADS / 32) * PAN_SLOTS);
        for (int32_t p = 0; p < n_panels; ++p) {
            k_spmm_pan<<<bl, PAN_THREADS, 0, s>>>(indices, values, poff + (int64_t)p * total_seg,
                                                  poff + (int64_t)(p + 1) * total_seg, (int32_t)total_seg, seg_dst, B4,
                                                  ldb / 4, d4, C, partial);
            LT_LAUNCH_CHECK();
        }
        if (bias || relu) {
            k_pan_epilogue<<<(unsigned)ceil_div((int64_t)n_long * 32, 256), 256, 0, s>>>(off_long, sel_long, n_long, d, bias,
                                                                                       relu, C, ldc);
            LT_LAUNCH_CHECK();
        }
        if (multi) {
            k_spmm_fixup<<<(unsigned)ceil_div((int64_t)n_long, 256), 256, 0, s>>>(off_long, sel_long, n_long, partial, d,
                                                                                     bias, relu, C, ldc);
            LT_LAUNCH_CHECK();
        }
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
