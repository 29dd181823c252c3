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
#include <string.h>

#include <mutex>

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


// seg_meta[g] = {first entry, one past the last entry, output position il, segments of the row}: a warp of the product
// reads its work item in ONE load instead of a 19-step binary search over seg_off followed by rows[] and indptr[]
// (four dependent round trips before the first gather on a kernel that is bound by latency x bytes in flight).
__global__ void k_spmm_segmeta(const int32_t* __restrict__ indptr, const int32_t* __restrict__ rows, const int32_t* __restrict__ sel,
                               int32_t n_rows, const int32_t* __restrict__ seg_off, int4* __restrict__ seg_meta) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows) return;
    const int32_t il = list_pos(sel, i);
    const int32_t r = rows ? rows[il] : il;
    const int32_t beg = indptr[r], end = indptr[r + 1];
    const int32_t first = seg_off[i], nseg = seg_off[i + 1] - first;
    for (int32_t k = 0; k < nseg; ++k)
        seg_meta[first + k] = make_int4(beg + k * SPMM_SEG, min(end, beg + (k + 1) * SPMM_SEG), il, nseg);
}

// NV = float4 slices per lane (d ≤ 128·NV).  seg_off: exclusive scan of nseg (n_rows+1 entries).
// The (column, value) pairs of the NEXT 32 non-zeros are loaded before the gathers of the current 32 are
// issued, so a warp never waits for an index load with nothing else in flight.  Lanes keep the float4
// index of "their" B row (32 bits — the host checks n_cols·ldb4 < 2³²), so one gather costs a shuffle,
// a wide multiply-add and the load.
// HUB: the most-gathered rows of B (the highest-degree columns) live in a compact copy `Hub4` that the launch keeps
// resident in the L2 (persisting access-policy window, see HubPanel below); a column c is a hub iff bit c of hub_bits is
// set and its slot is hub_pref[c / 32] + popcount of the lower bits.  Offsets carry the choice in their top bit.
constexpr uint32_t HUB_FLAG = 0x80000000u;

template <bool HUB>
__device__ __forceinline__ uint32_t spmm_offset(int32_t c, uint32_t ldb4, const HubView& hub) {
    if (HUB) return __ldg(hub.trans + c);            // one lookup in the per-attack translation table (10 MB at n = 2.45 M)
    return (uint32_t)c * ldb4;
}

// hub offsets of every work item, translated ONCE by a streaming pass (a warp per segment, coalesced) into a table padded
// to SPMM_SEG entries per segment: the product then reads offsets exactly as it reads column indices — the lookup
// (bitmap word + prefix, a dependent round trip) leaves the gather loop's critical path.
__global__ void __launch_bounds__(256)
k_hub_translate(const int4* __restrict__ seg_meta, const int32_t* __restrict__ seg_off, int32_t n_rows,
                const int32_t* __restrict__ indices, uint32_t ldb4, HubView hub, uint32_t* __restrict__ pre) {
    // 8 lanes per segment (the median segment has 48 entries): four segments per warp, eight independent index loads and
    // eight independent lookups in flight per lane
    const int sub = threadIdx.x & 7;
    const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (g >= __ldg(seg_off + n_rows)) return;
    const int4 m = __ldg(seg_meta + g);
    const int32_t cnt = m.y - m.x;
    for (int32_t t0 = 0; t0 < cnt; t0 += 64) {
        int32_t c[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = t0 + sub + 8 * u < cnt ? indices[m.x + t0 + sub + 8 * u] : -1;
        uint32_t o[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) o[u] = c[u] >= 0 ? __ldg(hub.trans + c[u]) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (c[u] >= 0) pre[g * SPMM_SEG + t0 + sub + 8 * u] = o[u];
    }
}

__device__ __forceinline__ int32_t seg_meta_total(const int32_t* __restrict__ seg_off, int32_t n_rows) { return __ldg(seg_off + n_rows); }

// L2 eviction-priority hints chosen from the VALUE of the entry (HUB == 2): under the symmetric normalisation
// Â[r,c] = 1/sqrt(d_r·d_c) the value read next to the column index already says how often column c is gathered
// (d_c = 1/(v²·d_r)), so a gather can be marked evict_last (hot: d_c ≥ hint_deg) or left alone without any lookup.
// A hint only — wrong guesses (other normalisations) cost cache efficiency, never correctness.
__device__ __forceinline__ float4 ldg_hint(const float4* p, uint64_t pol) {
    float4 r;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}

template <int NV, int GIF, int MINB, int HUB>   // GIF = gathers in flight per lane, MINB = resident CTAs per SM asked of ptxas
__global__ void __launch_bounds__(SPMM_THREADS, MINB)
k_spmm_vec(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
           const int32_t* __restrict__ rows, const int32_t* __restrict__ sel, int32_t n_rows, const int32_t* __restrict__ seg_off,
           const float4* __restrict__ B4, uint32_t ldb4, int32_t d4, const float* __restrict__ bias, int relu,
           float* __restrict__ C, int64_t ldc, float* __restrict__ partial, HubView hub, const int4* __restrict__ seg_meta,
           const uint32_t* __restrict__ pre = nullptr /* HUB: pre-translated offsets, SPMM_SEG per segment */,
           float hint_deg = 0.f, int hint_cold = 0) {
    const int lane = threadIdx.x & 31;
    const int64_t g = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) >> 5;
    int32_t s0, s1, il, nseg;
    if (seg_meta) {
        if (g >= seg_meta_total(seg_off, n_rows)) return;
        const int4 m = __ldg(seg_meta + g);
        s0 = m.x; s1 = m.y; il = m.z; nseg = m.w;
    } else {
        const int32_t total = seg_off[n_rows];
        if (g >= total) return;
        const int32_t i = (int32_t)upper_bound(seg_off, 0, (int64_t)n_rows + 1, (int32_t)g) - 1;
        il = list_pos(sel, i);
        const int32_t r = rows ? rows[il] : il;
        const int32_t first = seg_off[i];
        nseg = seg_off[i + 1] - first;
        const int32_t row_beg = indptr[r], row_end = indptr[r + 1];
        s0 = row_beg + ((int32_t)g - first) * SPMM_SEG;
        s1 = min(row_end, s0 + SPMM_SEG);
    }

    float4 acc[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    // lanes past the row width re-load the row's last float4 (same sector, result unused): a load that is
    // predicated per lane needs a zero-fill of its destination, and that write waits for the load —
    // which serialises the gathers (ncu: the select after each LDG was the top long_scoreboard stall)
    const float4* Bl[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) Bl[q] = B4 + min(lane + 32 * q, d4 - 1);
    const float4* Hl = HUB == 1 ? hub.rows4 + min(lane, d4 - 1) : nullptr;       // HUB ⇒ NV == 1

    uint64_t pol_hot = 0, pol_cold = 0;
    float hot_scale = 0.f;                   // hot ⇔ v²·d_r·hint_deg ≤ 1 (d_r: the row's length; split rows: their segment count x SPMM_SEG)
    if (HUB == 2) {
        if (hint_cold == 3) { asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_hot)); hint_cold = 1; }
        else asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hot));
        if (hint_cold == 1) asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
        else if (hint_cold == 2) asm("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol_cold));
        else asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_cold));
        hot_scale = (float)(nseg == 1 ? max(s1 - s0 - 1, 1) : nseg * SPMM_SEG) * hint_deg;
    }
    const uint32_t* pre_seg = (HUB == 1 && pre) ? pre + g * SPMM_SEG - s0 : nullptr;        // pre_seg[pos] for pos in [s0, s1)
    float mv = s0 + lane < s1 ? values[s0 + lane] : 0.f;
    uint32_t mo = s0 + lane < s1 ? (pre_seg ? pre_seg[s0 + lane] : spmm_offset<HUB == 1>(indices[s0 + lane], ldb4, hub)) : 0u;
    if (HUB == 2 && s0 + lane < s1 && mv * mv * hot_scale <= 1.f) mo |= HUB_FLAG;
    for (int32_t base = s0; base < s1; base += 32) {
        const int32_t nx = base + 32 + lane;
        uint32_t no = nx < s1 ? (pre_seg ? pre_seg[nx] : spmm_offset<HUB == 1>(indices[nx], ldb4, hub)) : 0u;
        const float nv = nx < s1 ? values[nx] : 0.f;
        if (HUB == 2 && nx < s1 && nv * nv * hot_scale <= 1.f) no |= HUB_FLAG;
        const int cnt = min(32, s1 - base);
        for (int k = 0; k < cnt; k += GIF) {
            // every load is unconditional and lands in its own register: entries past the end of the row
            // carry offset 0 (row 0 of B, an L1 hit) and are skipped by the warp-uniform guard below
            float4 x[GIF][NV];
#pragma unroll
            for (int u = 0; u < GIF; ++u) {
                const uint32_t o = __shfl_sync(LT_FULL, mo, (k + u) & 31);
                if (HUB == 1) {
                    x[u][0] = __ldg(((o & HUB_FLAG) ? Hl : Bl[0]) + (o & ~HUB_FLAG));
                } else if (HUB == 2) {
                    x[u][0] = ldg_hint(Bl[0] + (o & ~HUB_FLAG), (o & HUB_FLAG) ? pol_hot : pol_cold);
                } else {
#pragma unroll
                    for (int q = 0; q < NV; ++q) x[u][q] = __ldg(Bl[q] + o);
                }
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

// ------------------------------------------------------------------------------------ shared-memory staged variant
// k_spmm_vec keeps its gathers in registers: 8 rows of B in flight per lane, 24 warps per SM, and every batch of 8 ends
// in a full memory round trip before the next one is issued (ncu r2d: long_scoreboard 13 stalls per issue, DRAM 72 %
// busy, and 15 % fewer DRAM bytes — the hub panel — buy no time: the kernel is bound by latency x bytes in flight, not
// by bandwidth).  Here a warp streams the rows of B it needs through a ring of STG_RING slots in shared memory with
// cp.async (16 bytes per lane, L1 bypassed): the copy of entry e + STG_RING is issued the moment entry e has been
// consumed, so 32 rows (12.8 KB at d = 100) are in flight per warp at all times and 16 warps fill the SM's shared memory
// — 2.7x the bytes in flight of the register version.  Indices are fetched two chunks ahead (the hub translation of a
// chunk needs its own round trip).  Same FMAs in the same order as k_spmm_vec: the results are bit-identical.
constexpr int STG_RING = 32;
constexpr int STG_WARPS = 4;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool HUB>
__global__ void __launch_bounds__(32 * STG_WARPS, 4)
k_spmm_stage(const int32_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
             const int32_t* __restrict__ rows, const int32_t* __restrict__ sel, int32_t n_rows, const int32_t* __restrict__ seg_off,
             const float4* __restrict__ B4, uint32_t ldb4, int32_t d4, const float* __restrict__ bias, int relu,
             float* __restrict__ C, int64_t ldc, float* __restrict__ partial, HubView hub) {
    extern __shared__ float4 stage[];                                     // [STG_WARPS][STG_RING][d4]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t g = (int64_t)blockIdx.x * STG_WARPS + warp;
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
    float4* ring = stage + (size_t)warp * STG_RING * d4;
    const bool active = lane < d4;
    const float4* Bl = B4 + lane;
    const float4* Hl = HUB ? hub.rows4 + lane : nullptr;

    auto load_off = [&](int32_t pos) -> uint32_t { return pos < s1 ? spmm_offset<HUB>(indices[pos], ldb4, hub) : 0u; };
    auto issue = [&](uint32_t o, int slot, bool live) {                  // one row of B → ring[slot]; always one group
        if (live && active) {
            const float4* src = (HUB && (o & HUB_FLAG)) ? Hl + (o & ~HUB_FLAG) : Bl + o;
            cp_async16(ring + (size_t)slot * d4 + lane, src);
        }
        cp_async_commit();
    };

    // offsets of the current chunk, the next one and the one after; values of the current and the next chunk
    uint32_t o0 = load_off(s0 + lane), o1 = load_off(s0 + 32 + lane), o2 = load_off(s0 + 64 + lane);
    float v0 = s0 + lane < s1 ? values[s0 + lane] : 0.f;
    float v1 = s0 + 32 + lane < s1 ? values[s0 + 32 + lane] : 0.f;
    // prologue: the first STG_RING entries (= the first chunk)
#pragma unroll 8
    for (int k = 0; k < STG_RING; ++k) issue(__shfl_sync(LT_FULL, o0, k), k, s0 + k < s1);

    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int32_t base = s0; base < s1; base += 32) {
        const int cnt = min(32, s1 - base);
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            if (k >= cnt) { cp_async_commit(); continue; }               // warp-uniform; keeps one group per step
            cp_async_wait<STG_RING - 1>();                                // the oldest pending row (this entry) has landed
            __syncwarp();
            const float v = __shfl_sync(LT_FULL, v0, k);
            if (active) fma4(acc, v, ring[(size_t)k * d4 + lane]);
            __syncwarp();                                                 // slot k is free again
            issue(__shfl_sync(LT_FULL, o1, k), k, base + 32 + k < s1);    // entry of the next chunk into the same slot
        }
        o0 = o1; o1 = o2; v0 = v1;
        o2 = load_off(base + 96 + lane);
        v1 = base + 64 + lane < s1 ? values[base + 64 + lane] : 0.f;
    }
    cp_async_wait<0>();

    if (!active) return;
    if (nseg == 1) {
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + lane);
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        if (relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
        reinterpret_cast<float4*>(C + (int64_t)il * ldc)[lane] = acc;
    } else {
        (reinterpret_cast<float4*>(partial) + (int64_t)g * d4)[lane] = acc;
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

// ------------------------------------------------------------------------------------ hub panel (L2-resident hot rows of B)
// A restricted product Â[R,:]·B on a graph whose B does not fit the L2 re-fetches every gathered row from DRAM: on C5
// 13.9 GB of traffic for 1.4 GB of compulsory bytes (ncu r1k).  Gathers are not uniform: column c is gathered in
// proportion to its degree, and on a scale-free graph the K highest-degree columns take ~√(K/n) of all gathers (C5:
// K = 200 K rows = 76 MB → 28.6 %, measured from the CSR).  hub_prepare() copies those rows into one contiguous panel
// (rows padded to whole 32-byte sectors) and describes membership by a bitmap + popcount prefix (612 KB for 2.45 M
// nodes); the SpMM launch then runs under a persisting access-policy window over the panel, so hub gathers hit the L2's
// set-aside region whatever the streaming tail does to the rest of the cache.  The degree threshold is chosen on the
// device from a degree histogram (no host sync); the sum of a row is taken in the same order as without the panel, so
// the result is bit-identical.
constexpr int HUB_BINS = 4096;

__global__ void __launch_bounds__(256)
k_hub_hist(const int32_t* __restrict__ indptr, int32_t n, int32_t* __restrict__ hist) {
    __shared__ int32_t sh[HUB_BINS];
    for (int i = threadIdx.x; i < HUB_BINS; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&sh[min(indptr[r + 1] - indptr[r], HUB_BINS - 1)], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < HUB_BINS; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// thr[0] = smallest degree d with #{deg ≥ d} ≤ kmax.  One CTA of 1024 threads: suffix sums of the 4096 bins in shared
// memory (a single thread walking the bins took 252 µs: one dependent global load per bin).
__global__ void __launch_bounds__(1024)
k_hub_pick(const int32_t* __restrict__ hist, int32_t kmax, int32_t* __restrict__ thr) {
    __shared__ int32_t suf[HUB_BINS + 1];            // suf[d] = #{deg ≥ d}
    __shared__ int32_t blk[1024];
    const int t = threadIdx.x;
    // thread t owns bins 4t .. 4t+3
    const int32_t h0 = hist[4 * t], h1 = hist[4 * t + 1], h2 = hist[4 * t + 2], h3 = hist[4 * t + 3];
    blk[t] = h0 + h1 + h2 + h3;
    __syncthreads();
    // inclusive suffix scan of blk (Hillis-Steele over 1024 entries)
    for (int o = 1; o < 1024; o <<= 1) {
        const int32_t add = t + o < 1024 ? blk[t + o] : 0;
        __syncthreads();
        blk[t] += add;
        __syncthreads();
    }
    const int32_t above = t + 1 < 1024 ? blk[t + 1] : 0;          // Σ bins ≥ 4(t+1)
    suf[4 * t + 3] = above + h3;
    suf[4 * t + 2] = above + h3 + h2;
    suf[4 * t + 1] = above + h3 + h2 + h1;
    suf[4 * t] = above + h3 + h2 + h1 + h0;
    if (t == 0) suf[HUB_BINS] = 0;
    __syncthreads();
    // the smallest d ≥ 1 with suf[d] ≤ kmax: suf is non-increasing in d, so exactly one boundary exists
    for (int d = 1 + t; d <= HUB_BINS; d += 1024)
        if (suf[d] <= kmax && (d == 1 || suf[d - 1] > kmax)) thr[0] = d;      // d == HUB_BINS: no hubs
}
// warp per 32 nodes: membership word + its popcount
__global__ void __launch_bounds__(256)
k_hub_bits(const int32_t* __restrict__ indptr, int32_t n, const int32_t* __restrict__ thr, uint32_t* __restrict__ bits,
           int32_t* __restrict__ cnt, int32_t n_words) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_words) return;
    const int64_t r = w * 32 + lane;
    const int32_t d0 = thr[0];
    const bool hub = r < n && d0 < HUB_BINS && indptr[r + 1] - indptr[r] >= d0;
    const uint32_t b = __ballot_sync(LT_FULL, hub);
    if (lane == 0) { bits[w] = b; cnt[w] = __popc(b); }
}
// thread per column: its offset in the product's address space (see HubView::trans)
__global__ void k_hub_trans(const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref, int32_t n, uint32_t ldb4,
                            uint32_t ldh4, uint32_t* __restrict__ trans) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint32_t b = bits[c >> 5], m = 1u << (c & 31);
    trans[c] = (b & m) ? (HUB_FLAG | ((uint32_t)(pref[c >> 5] + __popc(b & (m - 1))) * ldh4)) : (uint32_t)c * ldb4;
}

// warp per word: copy its hub rows into their slots
__global__ void __launch_bounds__(256)
k_hub_copy(const uint32_t* __restrict__ bits, const int32_t* __restrict__ pref, int32_t n_words, const float4* __restrict__ B4,
           uint32_t ldb4, int32_t d4, float4* __restrict__ hub4, uint32_t ldh4) {
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_words) return;
    uint32_t b = bits[w];
    int32_t slot = pref[w];
    while (b) {
        const int bit = __ffs(b) - 1;
        b &= b - 1;
        const float4* src = B4 + (uint64_t)(w * 32 + bit) * ldb4;
        float4* dst = hub4 + (uint64_t)slot * ldh4;
        for (int q = lane; q < d4; q += 32) dst[q] = __ldg(src + q);
        ++slot;
    }
}

namespace {
struct HubStore { void* ptr = nullptr; size_t bytes = 0; bool limit_set = false; };
HubStore g_hub[64];
std::mutex g_hub_mu;
}  // namespace

// worth it only when B is far beyond the L2 and the vector kernel with one float4 per lane applies
bool hub_applicable(int32_t n_cols, int64_t ldb, int32_t d, const float* B) {
    return opt(OPT_SPMM_HUB) != 0 && d % 4 == 0 && d <= 128 && ldb % 4 == 0 && (uintptr_t)B % 16 == 0 &&
           (int64_t)n_cols * d * 4 >= ((int64_t)256 << 20) && (int64_t)n_cols * (ldb / 4) < ((int64_t)1 << 31);
}

// The degree from which a column's row of B counts as hot for the eviction-priority hints of the restricted product:
// the smallest degree d with #{columns of degree ≥ d}·(bytes a gathered row costs) ≤ spmm_hub_mb (64 MB, half the L2).
// Â is symmetric on this path, so row lengths are column counts.  0: no hints (option off, B fits the L2, shapes the
// hinted kernel does not take).  "spmm_hint_deg" > 0 fixes the value; otherwise it is measured ONCE per (indptr, n, d)
// from a degree histogram — one 4-byte read-back — and remembered: a stale entry (another graph at a recycled address)
// can only misplace hints, never change a result.
float spmm_hint_degree(const int32_t* indptr, int32_t n, const float* B, int64_t ldb, int32_t d, cudaStream_t s) {
    if (opt(OPT_SPMM_HINT) == 0 || d % 4 != 0 || d > 128 || ldb % 4 != 0 || (uintptr_t)B % 16 != 0 ||
        (int64_t)n * d * 4 < ((int64_t)256 << 20) || (int64_t)n * (ldb / 4) >= ((int64_t)1 << 31))
        return 0.f;
    const int fixed = opt(OPT_SPMM_HINT_DEG);
    if (fixed > 0) return (float)fixed;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0.f;
    const int mb = opt(OPT_SPMM_HUB_MB);
    struct Entry { const void* indptr; int32_t n, d; int mb, dev; float deg; };
    static Entry cache[8];
    static int next = 0;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    for (const Entry& e : cache)
        if (e.indptr == indptr && e.n == n && e.d == d && e.mb == mb && e.dev == dev) return e.deg;
    float deg = 0.f;
    int32_t* hist = nullptr;
    if (cudaMallocAsync((void**)&hist, ((size_t)HUB_BINS + 1) * 4, s) == cudaSuccess) {
        const int64_t row_bytes = ((int64_t)d * 4 + 63) / 64 * 64 + 48;        // 64-byte DRAM granules, unaligned start
        const int64_t kmax = ((int64_t)mb << 20) / row_bytes;
        int32_t thr = HUB_BINS;
        cudaMemsetAsync(hist, 0, ((size_t)HUB_BINS + 1) * 4, s);
        k_hub_hist<<<(unsigned)num_sms() * 4, 256, 0, s>>>(indptr, n, hist);
        k_hub_pick<<<1, 1024, 0, s>>>(hist, (int32_t)(kmax > n ? n : kmax), hist + HUB_BINS);
        const bool ok = cudaMemcpyAsync(&thr, hist + HUB_BINS, 4, cudaMemcpyDeviceToHost, s) == cudaSuccess &&
                        cudaStreamSynchronize(s) == cudaSuccess;
        cudaFreeAsync(hist, s);
        if (ok && thr > 1 && thr < HUB_BINS) deg = (float)thr;
    }
    cudaGetLastError();
    cache[next] = Entry{indptr, n, d, mb, dev, deg};
    next = (next + 1) % 8;
    return deg;
}

// Builds the panel for B on stream s.  bits / pref live in `scratch` (they must outlive the product); the panel rows live
// in a per-device buffer the library keeps (the second persistent allocation besides lt_plan: the persisting L2 lines
// must keep their address from one attack to the next, or stale lines would occupy the set-aside region).
int hub_prepare(const int32_t* indptr, int32_t n, const float* B, int64_t ldb, int32_t d, HubPanel* out, Scratch& scratch,
                cudaStream_t s) {
    out->valid = false;
    int dev = 0;
    LT_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return LT_OK;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    size_t cap = (size_t)opt(OPT_SPMM_HUB_MB) << 20;
    if ((size_t)max_persist < cap) cap = (size_t)max_persist;
    if ((size_t)max_window < cap) cap = (size_t)max_window;
    const int32_t d4 = d / 4;
    const uint32_t ldh4 = (uint32_t)((d4 + 1) & ~1);             // whole 32-byte sectors per row
    const int64_t kmax = (int64_t)(cap / ((size_t)ldh4 * 16));
    if (kmax < 1024) return LT_OK;                                // no set-aside L2 on this device / option set to 0
    void* panel = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_hub_mu);
        HubStore& st = g_hub[dev];
        if (st.bytes < cap) {
            if (st.ptr) cudaFree(st.ptr);
            st.ptr = nullptr; st.bytes = 0;
            if (cudaMalloc(&st.ptr, cap) != cudaSuccess) { cudaGetLastError(); return LT_OK; }   // no panel: plain product
            st.bytes = cap;
            st.limit_set = false;
        }
        if (!st.limit_set) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, cap) != cudaSuccess) { cudaGetLastError(); return LT_OK; }
            st.limit_set = true;
        }
        panel = st.ptr;
    }
    const int32_t n_words = (n + 31) / 32;
    int32_t *hist, *thr, *cnt, *pref;
    uint32_t* bits;
    LT_TRY(scratch.alloc(&hist, (size_t)HUB_BINS + 1));
    LT_TRY(scratch.alloc(&bits, (size_t)n_words));
    LT_TRY(scratch.alloc(&cnt, (size_t)n_words + 1));
    LT_TRY(scratch.alloc(&pref, (size_t)n_words + 1));
    uint32_t* trans;
    LT_TRY(scratch.alloc(&trans, (size_t)n));         // every allocation BEFORE the fork below: the side stream orders after it
    thr = hist + HUB_BINS;
    if (s != scratch.stream) {        // runs on another stream than the one that owns the scratch (and the inputs): order after it
        cudaEvent_t ev;
        LT_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        const cudaError_t e1 = cudaEventRecord(ev, scratch.stream);
        const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(s, ev, 0) : e1;
        cudaEventDestroy(ev);
        if (e2 != cudaSuccess) { set_error("hub_prepare: fork failed: %s", cudaGetErrorString(e2)); return LT_ERR_CUDA; }
    }
    ProfScope prof("spmm_hub_prepare", s);
    LT_CUDA(cudaMemsetAsync(hist, 0, ((size_t)HUB_BINS + 1) * 4, s));
    k_hub_hist<<<(unsigned)num_sms() * 4, 256, 0, s>>>(indptr, n, hist);
    LT_LAUNCH_CHECK();
    k_hub_pick<<<1, 1024, 0, s>>>(hist, (int32_t)(kmax > n ? n : kmax), thr);
    LT_LAUNCH_CHECK();
    k_hub_bits<<<(unsigned)ceil_div((int64_t)n_words * 32, 256), 256, 0, s>>>(indptr, n, thr, bits, cnt, n_words);
    LT_LAUNCH_CHECK();
    LT_TRY(exclusive_scan_i32(cnt, pref, n_words, s));
    k_hub_copy<<<(unsigned)ceil_div((int64_t)n_words * 32, 256), 256, 0, s>>>(bits, pref, n_words, reinterpret_cast<const float4*>(B),
                                                                           (uint32_t)(ldb / 4), d4, reinterpret_cast<float4*>(panel), ldh4);
    LT_LAUNCH_CHECK();
    k_hub_trans<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(bits, pref, n, (uint32_t)(ldb / 4), ldh4, trans);
    LT_LAUNCH_CHECK();
    out->view = HubView{bits, pref, reinterpret_cast<const float4*>(panel), ldh4, trans};
    out->base = panel;
    out->bytes = (size_t)kmax * ldh4 * 16;
    out->valid = true;
    return LT_OK;
}

// seg_off_known / total_seg_known (optional): exclusive scan of the per-row segment counts (n_rows+1 entries,
// SPMM_SEG non-zeros per segment, at least one per row) and its total, when the caller already has them — the
// attack plan does — so that the product needs neither the counting pass nor its host synchronisation.
// hub (optional): a panel built by hub_prepare() for this B; used by the plain pass of the one-float4-per-lane kernel.
int spmm_csr_f32_seg(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                     const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                     float* C, int64_t ldc, const int32_t* seg_off_known, int64_t total_seg_known, cudaStream_t s,
                     const HubPanel* hub = nullptr, float hint_deg = 0.f);

int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                 const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                 float* C, int64_t ldc, cudaStream_t s) {
    return spmm_csr_f32_seg(indptr, indices, values, n_cols, rows, n_rows, B, ldb, d, bias, relu, C, ldc, nullptr, -1, s);
}

int spmm_csr_f32_seg(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                     const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                     float* C, int64_t ldc, const int32_t* seg_off_known, int64_t total_seg_known, cudaStream_t s,
                     const HubPanel* hub, float hint_deg) {
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
        int4* seg_meta = nullptr;
        if (opt(OPT_SPMM_SEGMETA) != 0 && total_seg >= (int64_t)num_sms() * 64) {     // big products: one load per work item
            LT_TRY(scratch.alloc(&seg_meta, (size_t)total_seg));
            k_spmm_segmeta<<<(unsigned)ceil_div(n_list, 256), 256, 0, s>>>(indptr, rows, sel, n_list, seg_off, seg_meta);
            LT_LAUNCH_CHECK();
        }
#define LT_SPMM_V(NV_, GIF_, MINB_) k_spmm_vec<NV_, GIF_, MINB_, 0><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4, (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial, HubView{nullptr, nullptr, nullptr, 0, nullptr}, seg_meta)
        // d ≤ 128: 8 gathers in flight per lane at 3 CTAs/SM (no spills) measured best on C5 (2.54 ms; 4 CTAs 2.64,
        // 16 in flight at 2 CTAs 3.16): the kernel then moves its 13.9 GB of DRAM traffic at 84 % of the copy peak
        const bool use_hub = d4 <= 32 && hub && hub->valid;
        if (d4 <= 32 && opt(OPT_SPMM_STAGE) != 0 && total_seg >= (int64_t)num_sms() * 64) {
            const size_t smem = (size_t)STG_WARPS * STG_RING * d4 * sizeof(float4);
            const unsigned bls = (unsigned)ceil_div(total_seg, STG_WARPS);
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            if (use_hub) {
                attr.accessPolicyWindow.base_ptr = hub->base;
                attr.accessPolicyWindow.num_bytes = hub->bytes;
                attr.accessPolicyWindow.hitRatio = 1.0f;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                LT_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
                LT_CUDA(cudaFuncSetAttribute(k_spmm_stage<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_spmm_stage<true><<<bls, 32 * STG_WARPS, smem, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4,
                                                                     (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial, hub->view);
                attr.accessPolicyWindow.num_bytes = 0;
                cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
            } else {
                LT_CUDA(cudaFuncSetAttribute(k_spmm_stage<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_spmm_stage<false><<<bls, 32 * STG_WARPS, smem, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4,
                                                                      (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial,
                                                                      HubView{nullptr, nullptr, nullptr, 0, nullptr});
            }
        }
        else if (use_hub) {
            // hub gathers under a persisting window over the panel; everything else streams through the rest of the L2
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = hub->base;
            attr.accessPolicyWindow.num_bytes = hub->bytes;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            LT_CUDA(cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr));
            uint32_t* pre = nullptr;
            if (seg_meta && opt(OPT_SPMM_HUB) == 1) {                     // 2: look the offsets up inside the product instead
                LT_TRY(scratch.alloc(&pre, (size_t)total_seg * SPMM_SEG));
                k_hub_translate<<<(unsigned)ceil_div(total_seg * 8, 256), 256, 0, s>>>(seg_meta, seg_off, n_list, indices,
                                                                                        (uint32_t)(ldb / 4), hub->view, pre);
                LT_LAUNCH_CHECK();
            }
            k_spmm_vec<1, 8, 3, 1><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4,
                                                                  (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial, hub->view, seg_meta, pre);
            attr.accessPolicyWindow.num_bytes = 0;                      // the caller's stream gets its default policy back
            cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        else if (d4 <= 32 && hint_deg > 0.f && opt(OPT_SPMM_HINT) != 0 && seg_meta && (int64_t)n_cols * (ldb / 4) < ((int64_t)1 << 31)) {
            // value-derived eviction-priority hints (see ldg_hint).  spmm_hint = 1: cold gathers evict_normal, 2 (default):
            // evict_first, 3: evict_unchanged, 4: evict_first with the hot gathers left evict_normal
            k_spmm_vec<1, 8, 3, 2><<<bl, SPMM_THREADS, 0, s>>>(indptr, indices, values, rows, sel, n_list, seg_off, B4,
                                                               (uint32_t)(ldb / 4), d4, bias, relu, C, ldc, partial,
                                                               HubView{nullptr, nullptr, nullptr, 0, nullptr}, seg_meta, nullptr,
                                                               hint_deg, opt(OPT_SPMM_HINT) - 1);
        }
        else if (d4 <= 32) LT_SPMM_V(1, 8, 3);
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
