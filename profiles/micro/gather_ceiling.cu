// What can random 400-byte row gathers reach on this GPU?  A ceiling for k_spmm_vec (spmm.cu): same access pattern (a warp
// reads a whole row of B, 25 lanes x 16 B, row ids uniformly random over n rows so that the L2 cannot help), but no CSR
// walk, no per-row start-up chain and as many gathers in flight as registers allow.  Prints GB/s for several depths.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_ceiling gather_ceiling.cu && ./gather_ceiling
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int GIF>
__global__ void __launch_bounds__(256) k_gather(const uint32_t* __restrict__ idx, int64_t n_idx, const float4* __restrict__ B4,
                                                uint32_t ld4, int d4, float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float4* Bl = B4 + min(lane, d4 - 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // each warp owns a contiguous span of the index array; 32 ids per coalesced load, prefetched one chunk ahead
    const int64_t per = (n_idx / nwarps) & ~31LL;
    const int64_t beg = warp * per, end = beg + per;
    uint32_t mo = idx[beg + lane] * ld4;
    for (int64_t base = beg; base < end; base += 32) {
        const uint32_t no = base + 32 < end ? idx[base + 32 + lane] * ld4 : 0u;
#pragma unroll
        for (int k = 0; k < 32; k += GIF) {
            float4 x[GIF];
#pragma unroll
            for (int u = 0; u < GIF; ++u) x[u] = __ldg(Bl + __shfl_sync(0xffffffffu, mo, k + u));
#pragma unroll
            for (int u = 0; u < GIF; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
        }
        mo = no;
    }
    if (lane < d4) out[warp * 32 + lane] = acc;
}

template <int GIF>
static void run(const uint32_t* idx, int64_t n_idx, const float4* B4, uint32_t ld4, int d4, float4* out, int ctas_per_sm, int sms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = sms * ctas_per_sm;
    for (int it = 0; it < 3; ++it) k_gather<GIF><<<grid, 256>>>(idx, n_idx, B4, ld4, d4, out);
    cudaEventRecord(a);
    for (int it = 0; it < 5; ++it) k_gather<GIF><<<grid, 256>>>(idx, n_idx, B4, ld4, d4, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 5;
    const int64_t nw = (int64_t)grid * 8, per = (n_idx / nw) & ~31LL;
    const double bytes = (double)per * nw * (d4 * 16.0 + 4.0);
    printf("GIF=%2d ctas/SM=%d  %.3f ms  %.0f GB/s (%.2f M rows)\n", GIF, ctas_per_sm, ms, bytes / ms / 1e6, per * nw / 1e6);
}

// optional argument: the row stride of B in float4 units (25 = packed 400-byte rows; 26 = 416 B, rows start on a 32-byte
// sector boundary; 32 = 512 B) — does aligning the rows buy anything?
int main(int argc, char** argv) {
    const int64_t n = 2449029; const int d4 = 25; const uint32_t ld4 = argc > 1 ? (uint32_t)atoi(argv[1]) : 25;
    printf("row stride %u B\n", ld4 * 16);
    const int64_t n_idx = 31284977;                       // gathers of one C5 attack
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<uint32_t> h(n_idx);
    uint64_t s = 88172645463325252ULL;
    for (auto& v : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; v = (uint32_t)(s % n); }
    uint32_t* idx; float4 *B4, *out;
    cudaMalloc(&idx, n_idx * 4); cudaMalloc(&B4, n * ld4 * 16); cudaMalloc(&out, (size_t)sms * 8 * 8 * 32 * 16);
    cudaMemcpy(idx, h.data(), n_idx * 4, cudaMemcpyHostToDevice);
    cudaMemset(B4, 0, n * ld4 * 16);
    for (int c : (argc > 1 ? std::vector<int>{3, 4} : std::vector<int>{2, 3, 4, 6, 8})) { run<4>(idx, n_idx, B4, ld4, d4, out, c, sms); run<8>(idx, n_idx, B4, ld4, d4, out, c, sms); run<16>(idx, n_idx, B4, ld4, d4, out, c, sms); }
    return 0;
}
