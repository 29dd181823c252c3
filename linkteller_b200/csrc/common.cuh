// Shared helpers for the linkteller_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/linkteller_b200.h"

#define LT_WARP 32
#define LT_FULL 0xffffffffu

namespace lt {

// ---- error reporting: thread-local message, negative int codes across the C ABI ----
void set_error(const char* fmt, ...);
extern long long g_launches;

#define LT_CUDA(expr)                                                                        \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            lt::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return LT_ERR_CUDA;                                                              \
        }                                                                                    \
    } while (0)

#define LT_TRY(expr)                 \
    do {                             \
        int _s = (expr);             \
        if (_s != LT_OK) return _s;  \
    } while (0)

#define LT_CHECK_ARG(cond, msg)                                   \
    do {                                                          \
        if (!(cond)) {                                            \
            lt::set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return LT_ERR_ARG;                                    \
        }                                                         \
    } while (0)

// every kernel launch in the library is followed by one LT_LAUNCH_CHECK: it also counts launches
// (lt_launch_count) so that a bench can state how many of our kernels ran
#define LT_LAUNCH_CHECK()              \
    do {                               \
        ++lt::g_launches;              \
        LT_CUDA(cudaGetLastError());   \
    } while (0)

// ---- stream-ordered scratch memory (freed in reverse order by the destructor) ----
struct Scratch {
    cudaStream_t stream;
    void* ptrs[64];
    int n = 0;
    explicit Scratch(cudaStream_t s) : stream(s) {}
    ~Scratch() {
        for (int i = n - 1; i >= 0; --i) cudaFreeAsync(ptrs[i], stream);
    }
    template <typename T>
    int alloc(T** out, size_t count) {
        void* p = nullptr;
        size_t bytes = count * sizeof(T);
        if (bytes == 0) bytes = 16;
        if (n >= 64) {
            set_error("Scratch: too many allocations");
            return LT_ERR_INTERNAL;
        }
        cudaError_t e = cudaMallocAsync(&p, bytes, stream);
        if (e != cudaSuccess) {
            set_error("cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return LT_ERR_NOMEM;
        }
        ptrs[n++] = p;
        *out = reinterpret_cast<T*>(p);
        return LT_OK;
    }
};

int device_alloc(void** out, size_t bytes, cudaStream_t s);

// ---- process-wide options (lt_set_option / lt_get_option).  Initial values come from the environment ONCE, at the
// first query (LT_NO_OVERLAP, LT_NO_SIGN_PATH, LT_TC_WS, LT_SPMM_PANEL, LT_SPMM_PANEL_MB, LT_SPMM_PANEL_LMIN,
// LT_SPMM_HUB, LT_SPMM_HUB_MB, LT_FIX_EPS); after that no call reads the environment.
enum Opt { OPT_OVERLAP = 0, OPT_SIGN_PATH, OPT_TC_WS, OPT_SPMM_PANEL, OPT_SPMM_PANEL_MB, OPT_SPMM_PANEL_LMIN, OPT_SPMM_HUB,
           OPT_SPMM_HUB_MB, OPT_FIX_EPS, OPT_SPMM_STAGE, OPT_SPMM_SEGMETA, OPT_SHARD_ROWS, OPT_SPMM_HINT, OPT_SPMM_HINT_DEG, OPT_COUNT };
int opt(Opt which);

// ---- phase timers (lt_prof_*): CUDA events on the launching stream, only when enabled ----
struct ProfScope {
    int slot;
    cudaStream_t stream;
    ProfScope(const char* name, cudaStream_t s);
    ~ProfScope();
};
// host wall-clock of a region of library code (enqueue cost / time blocked in a sync), recorded only while the phase
// timers are on and reported by lt_prof_report as "host_<name> count total_ms"
struct HostScope {
    const char* name;
    long long t0;
    explicit HostScope(const char* n);
    ~HostScope();
};
int num_sms();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int bits_for(uint64_t max_value) {  // number of bits needed to represent max_value
    int b = 0;
    while (max_value) { ++b; max_value >>= 1; }
    return b < 1 ? 1 : b;
}

// ---- device helpers ----
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
    a.x = fmaf(v, x.x, a.x); a.y = fmaf(v, x.y, a.y); a.z = fmaf(v, x.z, a.z); a.w = fmaf(v, x.w, a.w);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(LT_FULL, v, o);
    return v;
}

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T u = __shfl_up_sync(LT_FULL, v, o);
        if (lane >= o) v += u;
    }
    return v;
}

// first index in [lo,hi) with a[idx] >= key (a ascending)
template <typename T, typename K>
__device__ __forceinline__ int64_t lower_bound(const T* __restrict__ a, int64_t lo, int64_t hi, K key) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}
// first index in [lo,hi) with a[idx] > key
template <typename T, typename K>
__device__ __forceinline__ int64_t upper_bound(const T* __restrict__ a, int64_t lo, int64_t hi, K key) {
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// non-zeros per warp work item of the CSR SpMM (spmm.cu); the attack plan sizes the segment list with it
constexpr int SPMM_SEG = 256;

// ---- hub panel of the restricted SpMM (spmm.cu): the most-gathered rows of the dense operand in one L2-resident copy ----
struct HubView {
    const uint32_t* bits;    // [ceil(n/32)] membership bitmap over the columns
    const int32_t* pref;     // [ceil(n/32) + 1] exclusive scan of the words' popcounts: slot of a hub column
    const float4* rows4;     // compact copy of the hub rows, row stride ld4 float4
    uint32_t ld4;
    const uint32_t* trans;   // [n] float4 offset of column c: HUB_FLAG | slot·ld4 into rows4 for a hub, c·ldb4 into B otherwise
};
struct HubPanel {
    HubView view{nullptr, nullptr, nullptr, 0, nullptr};
    void* base = nullptr;    // start and size of the persisting access-policy window
    size_t bytes = 0;
    bool valid = false;
};
struct Scratch;
bool hub_applicable(int32_t n_cols, int64_t ldb, int32_t d, const float* B);
float spmm_hint_degree(const int32_t* indptr, int32_t n, const float* B, int64_t ldb, int32_t d, cudaStream_t s);
int hub_prepare(const int32_t* indptr, int32_t n, const float* B, int64_t ldb, int32_t d, HubPanel* out, Scratch& scratch,
                cudaStream_t s);

// ---- lt_comm (comm.cu): one symmetric buffer per rank, mapped into every rank of the node with CUDA IPC ----
constexpr int LT_MAX_RANKS = 8;
}  // namespace lt
struct lt_comm {
    int rank = 0, world = 1;
    size_t bytes = 0;                      // payload bytes of every rank's buffer (the T x T matrix)
    size_t aux_bytes = 0;                  // second region of every rank's buffer: sign records of the frontier rows (0: none)
    void* peer[lt::LT_MAX_RANKS] = {};     // base of rank r's buffer in THIS process (peer[rank] = the local allocation)
    void* aux[lt::LT_MAX_RANKS] = {};      // rank r's aux region in THIS process
    uint32_t* flags[lt::LT_MAX_RANKS] = {};  // rank r's flag block: [3][LT_MAX_RANKS] epochs ("zeroed", "rows written", "signs pushed") by writer rank
    uint32_t epoch = 0;                    // attacks run so far (every rank calls in lock step)
    int* timeout_flag = nullptr;           // device word set by a wait that gave up
    bool opened = false;
    unsigned char handle[64] = {};         // cudaIpcMemHandle_t of the local allocation
};
namespace lt {
int comm_signal(lt_comm* c, int which, cudaStream_t s);   // which: 0 = "my matrix is zero-filled", 1 = "my rows are written", 2 = "my sign records are pushed"
int comm_wait(lt_comm* c, int which, cudaStream_t s);

// ---- primitives implemented in scan.cu / radix_sort.cu ----
// out has n+1 entries: out[0] = 0, out[k] = sum(in[0..k)), out[n] = total.  in may alias out.
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, cudaStream_t s);
int exclusive_scan_i64(const int64_t* in, int64_t* out, int64_t n, cudaStream_t s);
int exclusive_scan_i32_to_i64(const int32_t* in, int64_t* out, int64_t n, cudaStream_t s);
// the plan's frontier bitmap: popcount prefix per word + Σ SpMM segments / nnz of the marked rows, one launch
size_t frontier_word_scan_desc(int32_t n_words);
int frontier_word_scan(const uint32_t* bits, int32_t n_words, const int32_t* indptr, int32_t* pref, unsigned long long* desc,
                       unsigned long long* tot, int32_t* R_out, cudaStream_t s);

// Stable LSD radix sort of 64-bit keys over the given bit ranges (8-bit digits), optional 64-bit
// payload.  Sorted data ends in keys / vals (in place from the caller's view).
struct BitRange { int begin, end; };
int radix_sort_u64(uint64_t* keys, uint64_t* vals, int64_t n, const BitRange* ranges, int n_ranges,
                   cudaStream_t s);

}  // namespace lt
