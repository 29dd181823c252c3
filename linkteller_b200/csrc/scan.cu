// Device-wide exclusive prefix sum (reduce → scan of block sums → downsweep), recursive over
// block sums.  Used by the CSR build, the frontier plan, the radix sort and the ranking kernels.
#include "common.cuh"

#include <type_traits>

namespace lt {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096

// Block-wide exclusive scan of one value per thread; returns the exclusive prefix and the block
// total (in `total`).  All SCAN_THREADS threads must call.
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T& total) {
    __shared__ T warp_tot[SCAN_THREADS / 32];
    __shared__ T blk_tot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_inclusive_scan(v);
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < SCAN_THREADS / 32 ? warp_tot[lane] : T(0);
        T winc = warp_inclusive_scan(w);
        if (lane < SCAN_THREADS / 32) warp_tot[lane] = winc - w;  // exclusive prefix of the warp
        if (lane == SCAN_THREADS / 32 - 1) blk_tot = winc;
    }
    __syncthreads();
    T res = warp_tot[warp] + inc - v;
    total = blk_tot;
    __syncthreads();  // shared arrays are reused by the next call
    return res;
}

// the input is treated as having n_in valid elements; indices >= n_in read as 0
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const TI* __restrict__ in, int64_t n_in,
                                                             TO* __restrict__ block_sums) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TO s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n_in) s += (TO)in[i];
    }
    TO total;
    block_exclusive_scan<TO>(s, total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// out[i] = offsets[block] + exclusive prefix inside the tile, for i < n_out
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_down(const TI* in, int64_t n_in, TO* out,
                                                           int64_t n_out,
                                                           const TO* __restrict__ offsets) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    TO v[SCAN_ITEMS];
    TO s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        v[k] = (i < n_in) ? (TO)in[i] : TO(0);
        s += v[k];
    }
    TO total;
    TO pre = block_exclusive_scan<TO>(s, total);
    pre += offsets ? offsets[blockIdx.x] : TO(0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        int64_t i = base + k;
        if (i < n_out) out[i] = pre;
        pre += v[k];
    }
}

// one CTA walks the whole input tile by tile, carrying the running total: one launch instead of three
// for inputs of a few tiles (the plan has many scans of T-sized arrays)
template <typename TI, typename TO>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_small(const TI* in, int64_t n_in, TO* out, int64_t n_out) {
    TO carry = 0;
    for (int64_t tile = 0; tile < n_out; tile += SCAN_TILE) {
        const int64_t base = tile + (int64_t)threadIdx.x * SCAN_ITEMS;
        TO v[SCAN_ITEMS];
        TO s = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            v[k] = (i < n_in) ? (TO)in[i] : TO(0);
            s += v[k];
        }
        TO total;
        TO pre = block_exclusive_scan<TO>(s, total) + carry;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t i = base + k;
            if (i < n_out) out[i] = pre;
            pre += v[k];
        }
        carry += total;
    }
}

// Single-pass scan with decoupled look-back (int32 → int32, values ≥ 0): every tile publishes its aggregate and
// then its inclusive prefix in one 64-bit descriptor (status in the high word), and a tile obtains its exclusive
// prefix by walking back over its predecessors' descriptors.  Tiles take their index from an atomic counter, so a
// tile's predecessors have always started.  One launch (plus one small memset) instead of three for the mid-sized
// scans of the plan — about twenty per attack.
constexpr uint32_t LB_AGGREGATE = 1u, LB_INCLUSIVE = 2u;

// The value of element i comes from a loader (plain int32 input, or something computed on the fly — the plan's frontier
// bitmap words), ITEMS consecutive elements per thread.
struct LoadI32 {
    const int32_t* in;
    __device__ __forceinline__ int32_t operator()(int64_t i) { return in[i]; }
    __device__ __forceinline__ void finish() {}
};

// element w = popcount of word w of the frontier bitmap; on the side, Σ over the marked rows of their SpMM segment counts
// and of their lengths (one atomic pair per CTA) — all the first host read-back of the plan needs, in one launch
struct LoadFrontierWords {
    const uint32_t* bits;
    const int32_t* indptr;
    unsigned long long* tot;               // [0] Σ segments, [1] Σ nnz
    unsigned long long seg = 0, nnz = 0;
    __device__ __forceinline__ int32_t operator()(int64_t w) {
        uint32_t b = bits[w];
        const int32_t c = __popc(b);
        while (b) {
            const int32_t r = (int32_t)(w * 32) + __ffs(b) - 1;
            b &= b - 1;
            const int32_t len = indptr[r + 1] - indptr[r];
            seg += len <= SPMM_SEG ? 1 : (len + SPMM_SEG - 1) / SPMM_SEG;
            nnz += len;
        }
        return c;
    }
    __device__ __forceinline__ void finish() {         // every thread of the CTA calls
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { seg += __shfl_xor_sync(LT_FULL, seg, o); nnz += __shfl_xor_sync(LT_FULL, nnz, o); }
        __shared__ unsigned long long sh[2][SCAN_THREADS / 32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) { sh[0][warp] = seg; sh[1][warp] = nnz; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int k = 1; k < SCAN_THREADS / 32; ++k) { seg += sh[0][k]; nnz += sh[1][k]; }
            if (seg | nnz) { atomicAdd(&tot[0], seg); atomicAdd(&tot[1], nnz); }
        }
    }
};

template <typename Load, int ITEMS>
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_lookback(Load load, int64_t n_in, int32_t* out, int64_t n_out, unsigned long long* desc /* [0] = tile counter */,
                int32_t* total_out /* nullable: also receives out[n_in] */) {
    constexpr int TILE = SCAN_THREADS * ITEMS;
    __shared__ int32_t tile_s, prefix_s;
    if (threadIdx.x == 0) tile_s = (int32_t)atomicAdd(desc, 1ull);
    __syncthreads();
    const int32_t tile = tile_s;
    unsigned long long* d = desc + 1;
    const int64_t base = (int64_t)tile * TILE + (int64_t)threadIdx.x * ITEMS;
    int32_t v[ITEMS];
    int32_t sum = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + k;
        v[k] = (i < n_in) ? load(i) : 0;
        sum += v[k];
    }
    load.finish();
    int32_t total;
    int32_t pre = block_exclusive_scan<int32_t>(sum, total);
    if (threadIdx.x < 32) {
        // warp 0 looks back 32 descriptors at a time: lane l reads predecessor tile−1−l (spinning until it has
        // published), the nearest inclusive prefix ends the walk.  One predecessor per step made the last of ~90
        // resident tiles wait ~90 dependent reads (ncu r1k: 14 µs per mid-sized scan).
        const int lane = threadIdx.x;
        int32_t excl = 0;
        if (tile > 0) {
            if (lane == 0) atomicExch(&d[tile], ((unsigned long long)LB_AGGREGATE << 32) | (uint32_t)total);
            for (int32_t t_hi = tile - 1;; t_hi -= 32) {
                const int32_t t = t_hi - lane;
                unsigned long long w = (unsigned long long)LB_INCLUSIVE << 32;       // before tile 0: prefix 0
                if (t >= 0) { do { w = atomicAdd(&d[t], 0ull); } while ((w >> 32) == 0); }
                const uint32_t incl = __ballot_sync(LT_FULL, (w >> 32) == LB_INCLUSIVE);
                const int first = __ffs(incl) - 1;                                   // nearest inclusive prefix, if any
                int32_t val = (first < 0 || lane <= first) ? (int32_t)(uint32_t)w : 0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(LT_FULL, val, o);
                excl += val;
                if (first >= 0) break;
            }
        }
        if (lane == 0) {
            atomicExch(&d[tile], ((unsigned long long)LB_INCLUSIVE << 32) | (uint32_t)(excl + total));
            prefix_s = excl;
        }
    }
    __syncthreads();
    pre += prefix_s;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + k;
        if (i < n_out) out[i] = pre;
        if (total_out && i == n_in) *total_out = pre;
        pre += v[k];
    }
}

// Exclusive scan producing n_in+1 outputs (out[n_in] = total).  in may alias out when TI == TO
// (every tile is fully read into registers before it is written, tiles do not overlap).
template <typename TI, typename TO>
static int scan_impl(const TI* in, TO* out, int64_t n_in, cudaStream_t s) {
    const int64_t n_out = n_in + 1;
    const int64_t blocks = ceil_div(n_out, SCAN_TILE);
    if (blocks <= 16) {
        k_scan_small<TI, TO><<<1, SCAN_THREADS, 0, s>>>(in, n_in, out, n_out);
        LT_LAUNCH_CHECK();
        return LT_OK;
    }
    Scratch scratch(s);
    if (std::is_same<TI, int32_t>::value && std::is_same<TO, int32_t>::value) {
        unsigned long long* desc = nullptr;
        LT_TRY(scratch.alloc(&desc, (size_t)blocks + 1));
        LT_CUDA(cudaMemsetAsync(desc, 0, ((size_t)blocks + 1) * 8, s));
        k_scan_lookback<LoadI32, SCAN_ITEMS><<<(unsigned)blocks, SCAN_THREADS, 0, s>>>(
            LoadI32{reinterpret_cast<const int32_t*>(in)}, n_in, reinterpret_cast<int32_t*>(out), n_out, desc, nullptr);
        LT_LAUNCH_CHECK();
        return LT_OK;
    }
    TO* sums = nullptr;
    LT_TRY(scratch.alloc(&sums, (size_t)blocks + 1));
    k_scan_reduce<TI, TO><<<(unsigned)blocks, SCAN_THREADS, 0, s>>>(in, n_in, sums);
    LT_LAUNCH_CHECK();
    LT_TRY((scan_impl<TO, TO>(sums, sums, blocks, s)));
    k_scan_down<TI, TO><<<(unsigned)blocks, SCAN_THREADS, 0, s>>>(in, n_in, out, n_out, sums);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

// pref[w] = number of frontier rows in the words before w (pref[n_words] = R, also written to *R_out); tot += (Σ segments,
// Σ nnz) of the marked rows.  desc: frontier_word_scan_desc(n_words) zero-filled 64-bit words.  One word per thread: a
// thread walks up to 32 rows of its word, wider tiles would serialise that.
size_t frontier_word_scan_desc(int32_t n_words) { return (size_t)ceil_div((int64_t)n_words + 1, SCAN_THREADS) + 1; }
int frontier_word_scan(const uint32_t* bits, int32_t n_words, const int32_t* indptr, int32_t* pref, unsigned long long* desc,
                       unsigned long long* tot, int32_t* R_out, cudaStream_t s) {
    const int64_t blocks = ceil_div((int64_t)n_words + 1, SCAN_THREADS);
    k_scan_lookback<LoadFrontierWords, 1><<<(unsigned)blocks, SCAN_THREADS, 0, s>>>(LoadFrontierWords{bits, indptr, tot}, n_words,
                                                                                     pref, (int64_t)n_words + 1, desc, R_out);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, cudaStream_t s) {
    return scan_impl<int32_t, int32_t>(in, out, n, s);
}
int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, cudaStream_t s) {
    return scan_impl<int64_t, int64_t>(in, out, n, s);
}
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s) {
    return scan_impl<int32_t, int64_t>(in, out, n, s);
}

}  // namespace lt

extern "C" int lt_exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, lt_stream_t stream) {
    LT_CHECK_ARG(n >= 0 && (in || n == 0) && out, "lt_exclusive_scan_i32");
    return lt::exclusive_scan_i32(in, out, n, (cudaStream_t)stream);
}
