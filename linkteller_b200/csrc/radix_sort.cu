// Stable LSD radix sort of 64-bit keys (optional 64-bit payload), 8-bit digits.
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the digit-major count table,
// (3) stable scatter.  A tile is 8 warps x 512 consecutive keys; inside a warp the keys are taken
// 32 at a time and ranked with __match_any_sync, which keeps equal digits in input order, so the
// sort is stable — the frontier plan relies on that (contributions of a (v,u) pair stay in
// neighbour order, which makes the float sums deterministic).
//
// HBM traffic per pass: keys+payload read twice and written once (N·(2·16+16) B with payload).
#include "common.cuh"

namespace lt {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_BINS = 256;
// keys per thread: 16 (tiles of 4096 keys) for large inputs; 4 (tiles of 1024) when 4096-key tiles would
// leave most SMs without a CTA — the plan's record sort (~10^6 keys) is latency-bound, not bandwidth-bound

template <int RS_ROUNDS>
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n,
                                                       int shift, int32_t* __restrict__ counts,
                                                       int32_t n_tiles) {
    constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
    __shared__ int32_t hist[RS_BINS];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int q = 0; q < RS_ROUNDS; ++q) {
        int64_t i = base + (int64_t)q * RS_THREADS + threadIdx.x;  // order is irrelevant for counting
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 0xFF], 1);
    }
    __syncthreads();
    counts[(int64_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

template <bool HAS_VALS, int RS_ROUNDS>
__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const uint64_t* __restrict__ keys_in,
                                                          const uint64_t* __restrict__ vals_in,
                                                          uint64_t* __restrict__ keys_out,
                                                          uint64_t* __restrict__ vals_out, int64_t n,
                                                          int shift, const int32_t* __restrict__ offsets,
                                                          int32_t n_tiles) {
    constexpr int RS_WARP_KEYS = 32 * RS_ROUNDS;
    constexpr int RS_TILE = RS_THREADS * RS_ROUNDS;
    __shared__ int32_t wcnt[RS_WARPS][RS_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&wcnt[0][0])[k] = 0;
    __syncthreads();

    const int64_t wbase = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * RS_WARP_KEYS;
    uint64_t key[RS_ROUNDS];
    int32_t rank[RS_ROUNDS];
#pragma unroll
    for (int q = 0; q < RS_ROUNDS; ++q) {
        const int64_t i = wbase + q * 32 + lane;
        const bool valid = i < n;
        key[q] = valid ? keys_in[i] : 0;
        const unsigned act = __ballot_sync(LT_FULL, valid);
        rank[q] = 0;
        if (valid) {
            const int d = (int)((key[q] >> shift) & 0xFF);
            const unsigned peers = __match_any_sync(act, d);
            const int leader = __ffs(peers) - 1;
            int old = 0;
            if (lane == leader) {
                old = wcnt[warp][d];
                wcnt[warp][d] = old + __popc(peers);
            }
            old = __shfl_sync(peers, old, leader);
            rank[q] = old + __popc(peers & ((1u << lane) - 1u));
        }
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: turn per-warp counts into global base offsets
        const int d = threadIdx.x;
        int32_t run = offsets[(int64_t)d * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            int32_t c = wcnt[w][d];
            wcnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RS_ROUNDS; ++q) {
        const int64_t i = wbase + q * 32 + lane;
        if (i < n) {
            const int d = (int)((key[q] >> shift) & 0xFF);
            const int64_t pos = (int64_t)wcnt[warp][d] + rank[q];
            keys_out[pos] = key[q];
            if (HAS_VALS) vals_out[pos] = vals_in[i];
        }
    }
}

int radix_sort_u64(uint64_t* keys, uint64_t* vals, int64_t n, const BitRange* ranges, int n_ranges,
                   cudaStream_t s) {
    if (n <= 1) return LT_OK;
    if (n >= (int64_t)1 << 31) {
        set_error("radix_sort_u64: n=%lld exceeds 2^31-1", (long long)n);
        return LT_ERR_ARG;
    }
    int shifts[16];
    int n_pass = 0;
    for (int r = 0; r < n_ranges; ++r)
        for (int b = ranges[r].begin; b < ranges[r].end; b += 8) {
            if (n_pass >= 16) { set_error("radix_sort_u64: too many passes"); return LT_ERR_ARG; }
            shifts[n_pass++] = b;
        }
    if (n_pass == 0) return LT_OK;
    const bool small = ceil_div(n, RS_THREADS * 16) < 4 * (int64_t)num_sms();
    const int64_t n_tiles = ceil_div(n, RS_THREADS * (small ? 4 : 16));
    Scratch scratch(s);
    uint64_t *k2 = nullptr, *v2 = nullptr;
    int32_t* counts = nullptr;
    LT_TRY(scratch.alloc(&k2, (size_t)n));
    if (vals) LT_TRY(scratch.alloc(&v2, (size_t)n));
    LT_TRY(scratch.alloc(&counts, (size_t)(RS_BINS * n_tiles + 1)));
    uint64_t *ki = keys, *ko = k2, *vi = vals, *vo = v2;
    const unsigned g = (unsigned)n_tiles;
    const int32_t nt = (int32_t)n_tiles;
    for (int p = 0; p < n_pass; ++p) {
        const int sh = shifts[p];
        if (small) k_rs_hist<4><<<g, RS_THREADS, 0, s>>>(ki, n, sh, counts, nt);
        else k_rs_hist<16><<<g, RS_THREADS, 0, s>>>(ki, n, sh, counts, nt);
        LT_LAUNCH_CHECK();
        LT_TRY(exclusive_scan_i32(counts, counts, RS_BINS * n_tiles, s));
        if (vals) {
            if (small) k_rs_scatter<true, 4><<<g, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, sh, counts, nt);
            else k_rs_scatter<true, 16><<<g, RS_THREADS, 0, s>>>(ki, vi, ko, vo, n, sh, counts, nt);
        } else {
            if (small) k_rs_scatter<false, 4><<<g, RS_THREADS, 0, s>>>(ki, nullptr, ko, nullptr, n, sh, counts, nt);
            else k_rs_scatter<false, 16><<<g, RS_THREADS, 0, s>>>(ki, nullptr, ko, nullptr, n, sh, counts, nt);
        }
        LT_LAUNCH_CHECK();
        uint64_t* t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
    }
    if (ki != keys) {
        LT_CUDA(cudaMemcpyAsync(keys, ki, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
        if (vals) LT_CUDA(cudaMemcpyAsync(vals, vi, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    }
    return LT_OK;
}

}  // namespace lt

extern "C" int lt_radix_sort_pairs_u64(uint64_t* keys, uint64_t* vals, int64_t n, int begin_bit, int end_bit,
                                       lt_stream_t stream) {
    LT_CHECK_ARG(keys && n >= 0 && begin_bit >= 0 && end_bit <= 64 && begin_bit <= end_bit,
                 "lt_radix_sort_pairs_u64");
    lt::BitRange r{begin_bit, end_bit};
    return lt::radix_sort_u64(keys, vals, n, &r, 1, (cudaStream_t)stream);
}
