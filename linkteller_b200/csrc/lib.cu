// Library plumbing: thread-local error message, device info, GCN forward composed from the kernels.
#include "common.cuh"

#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <mutex>

namespace lt {

// ---- options ----
namespace {
struct OptDef { const char* name; const char* env; int dflt; bool env_negates; };
// env_negates: the variable being SET (to anything) switches the option off (LT_NO_*)
const OptDef kOpts[OPT_COUNT] = {
    {"overlap", "LT_NO_OVERLAP", 1, true},          {"sign_path", "LT_NO_SIGN_PATH", 1, true},
    {"tc_ws", "LT_TC_WS", 1, false},                {"spmm_panel", "LT_SPMM_PANEL", 0, false},
    {"spmm_panel_mb", "LT_SPMM_PANEL_MB", 48, false}, {"spmm_panel_lmin", "LT_SPMM_PANEL_LMIN", 64, false},
    {"spmm_hub", "LT_SPMM_HUB", 0, false},          {"spmm_hub_mb", "LT_SPMM_HUB_MB", 64, false},
    {"fix_eps", "LT_FIX_EPS", 19, false},           {"spmm_stage", "LT_SPMM_STAGE", 0, false},
    {"spmm_segmeta", "LT_SPMM_SEGMETA", 1, false}, {"shard_rows", "LT_SHARD_ROWS", 1, false},
    {"spmm_hint", "LT_SPMM_HINT", 2, false},       {"spmm_hint_deg", "LT_SPMM_HINT_DEG", 0, false},
};
std::atomic<int> g_opt[OPT_COUNT];
std::once_flag g_opt_once;
void opt_init() {
    for (int i = 0; i < OPT_COUNT; ++i) {
        int v = kOpts[i].dflt;
        const char* e = getenv(kOpts[i].env);
        if (e) v = kOpts[i].env_negates ? 0 : atoi(e);
        g_opt[i].store(v);
    }
}
}  // namespace

int opt(Opt which) {
    std::call_once(g_opt_once, opt_init);
    return g_opt[which].load(std::memory_order_relaxed);
}

static thread_local char g_err[1024] = "";
long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- phase timers ----
namespace {
struct ProfRec { const char* name; cudaEvent_t a, b; };
constexpr int PROF_MAX = 4096;
ProfRec g_prof[PROF_MAX];
int g_prof_n = 0;
int g_prof_created = 0;      // slots [0, g_prof_created) own a pair of events
bool g_prof_on = false;
}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t s) : slot(-1), stream(s) {
    if (!g_prof_on || g_prof_n >= PROF_MAX) return;
    ProfRec& r = g_prof[g_prof_n];
    // events are created once per slot and reused by later measurements (lt_prof_reset only rewinds): creating two events
    // per scope cost ~0.1 ms of host time per attack inside the very region being measured
    if (g_prof_n >= g_prof_created) {
        if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
        g_prof_created = g_prof_n + 1;
    }
    r.name = name;
    slot = g_prof_n++;
    cudaEventRecord(r.a, s);
}
ProfScope::~ProfScope() {
    if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

namespace {
struct HostRec { const char* name; int count; double ms; };
HostRec g_host[64];
int g_host_n = 0;
}  // namespace
static long long now_ns() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
HostScope::HostScope(const char* n) : name(n), t0(g_prof_on ? now_ns() : 0) {}
HostScope::~HostScope() {
    if (!t0) return;
    const double ms = (now_ns() - t0) * 1e-6;
    for (int i = 0; i < g_host_n; ++i)
        if (g_host[i].name == name || strcmp(g_host[i].name, name) == 0) { g_host[i].count++; g_host[i].ms += ms; return; }
    if (g_host_n < 64) g_host[g_host_n++] = HostRec{name, 1, ms};
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

int gemm_f32(const float* A, int64_t lda, const int32_t* a_rows, const float* B, int64_t ldb, int transB,
             const float* bias, int relu, float* C, int64_t ldc, int32_t M, int32_t N, int32_t K, int algo,
             cudaStream_t s);
int spmm_csr_f32(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n_cols,
                 const int32_t* rows, int32_t n_rows, const float* B, int64_t ldb, int32_t d, const float* bias, int relu,
                 float* C, int64_t ldc, cudaStream_t s);

}  // namespace lt

using namespace lt;

extern "C" const char* lt_last_error(void) { return g_err; }
extern "C" int lt_version(void) { return 100; }

extern "C" int lt_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host, int64_t* mem_bytes_host) {
    int dev = 0;
    LT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    LT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count_host) *sm_count_host = prop.multiProcessorCount;
    if (cc_major_host) *cc_major_host = prop.major;
    if (cc_minor_host) *cc_minor_host = prop.minor;
    if (mem_bytes_host) *mem_bytes_host = (int64_t)prop.totalGlobalMem;
    // keep freed stream-ordered scratch in the pool instead of returning it to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    return LT_OK;
}

extern "C" int lt_set_option(const char* name, int value) {
    LT_CHECK_ARG(name, "lt_set_option: null name");
    std::call_once(g_opt_once, opt_init);
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(kOpts[i].name, name) == 0) { g_opt[i].store(value); return LT_OK; }
    set_error("lt_set_option: unknown option '%s'", name);
    return LT_ERR_ARG;
}
extern "C" int lt_get_option(const char* name, int* value_host) {
    LT_CHECK_ARG(name && value_host, "lt_get_option: null pointer");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(kOpts[i].name, name) == 0) { *value_host = opt((Opt)i); return LT_OK; }
    set_error("lt_get_option: unknown option '%s'", name);
    return LT_ERR_ARG;
}

extern "C" int64_t lt_launch_count(void) { return (int64_t)g_launches; }
extern "C" int lt_prof_enable(int on) { g_prof_on = on != 0; return LT_OK; }
extern "C" int lt_prof_reset(void) {
    g_prof_n = 0;
    g_host_n = 0;
    return LT_OK;
}
extern "C" int lt_prof_report(char* buf_host, int64_t cap) {
    LT_CHECK_ARG(buf_host && cap > 0, "lt_prof_report");
    LT_CUDA(cudaDeviceSynchronize());
    const char* names[256];
    int counts[256];
    double total[256];
    int nn = 0;
    for (int i = 0; i < g_prof_n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b) != cudaSuccess) continue;
        int k = 0;
        while (k < nn && strcmp(names[k], g_prof[i].name) != 0) ++k;
        if (k == nn) {
            if (nn == 256) continue;
            names[nn] = g_prof[i].name; counts[nn] = 0; total[nn] = 0.0; ++nn;
        }
        counts[k] += 1;
        total[k] += ms;
    }
    int64_t off = 0;
    buf_host[0] = 0;
    for (int k = 0; k < nn; ++k) {
        int w = snprintf(buf_host + off, (size_t)(cap - off), "%s %d %.6f\n", names[k], counts[k], total[k]);
        if (w < 0 || off + w >= cap) { set_error("lt_prof_report: buffer too small"); return LT_ERR_CAPACITY; }
        off += w;
    }
    for (int k = 0; k < g_host_n; ++k) {
        int w = snprintf(buf_host + off, (size_t)(cap - off), "host_%s %d %.6f\n", g_host[k].name, g_host[k].count, g_host[k].ms);
        if (w < 0 || off + w >= cap) { set_error("lt_prof_report: buffer too small"); return LT_ERR_CAPACITY; }
        off += w;
    }
    return LT_OK;
}

// GCN.forward (gcn/models.py:19-24), eval mode: Z = Â·(relu(Â·(X·W1) + b1)·W2) + b2
extern "C" int lt_gcn_forward(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                              const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                              const float* W2, const float* b2, int32_t h, int32_t c, float* logits, int64_t ldz,
                              lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && X && W1 && W2 && logits, "lt_gcn_forward: null pointer");
    LT_CHECK_ARG(n > 0 && f > 0 && h > 0 && c > 0 && ldx >= f && ldz >= c, "lt_gcn_forward: sizes");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    float *P, *Rl, *Q;
    const int64_t hp = (h + 3) / 4 * 4;   // padded leading dimension keeps the float4 SpMM path
    LT_TRY(scratch.alloc(&P, (size_t)n * hp));
    LT_TRY(scratch.alloc(&Rl, (size_t)n * hp));
    LT_TRY(scratch.alloc(&Q, (size_t)n * c));
    LT_TRY(gemm_f32(X, ldx, nullptr, W1, h, 0, nullptr, 0, P, hp, n, h, f, 0, s));            // layers.py:31
    LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, P, hp, h, b1, 1, Rl, hp, s));     // :32-34 + relu
    LT_TRY(gemm_f32(Rl, hp, nullptr, W2, c, 0, nullptr, 0, Q, c, n, c, h, 0, s));
    LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, Q, c, c, b2, 0, logits, ldz, s));
    return LT_OK;
}

// GCN3.forward (gcn/models.py:40-46), eval mode: Z = Â·(relu(Â·(relu(Â·(X·W1) + b1)·W2) + b2)·W3) + b3
extern "C" int lt_gcn3_forward(const int32_t* indptr, const int32_t* indices, const float* values, int32_t n,
                               const float* X, int64_t ldx, int32_t f, const float* W1, const float* b1,
                               const float* W2, const float* b2, const float* W3, const float* b3, int32_t h1,
                               int32_t h2, int32_t c, float* logits, int64_t ldz, lt_stream_t stream) {
    LT_CHECK_ARG(indptr && indices && values && X && W1 && W2 && W3 && logits, "lt_gcn3_forward: null pointer");
    LT_CHECK_ARG(n > 0 && f > 0 && h1 > 0 && h2 > 0 && c > 0 && ldx >= f && ldz >= c, "lt_gcn3_forward: sizes");
    cudaStream_t s = (cudaStream_t)stream;
    Scratch scratch(s);
    float *P, *R1, *G, *R2, *Q;
    const int64_t p1 = (h1 + 3) / 4 * 4, p2 = (h2 + 3) / 4 * 4;   // padded leading dimensions keep the float4 SpMM path
    LT_TRY(scratch.alloc(&P, (size_t)n * p1));
    LT_TRY(scratch.alloc(&R1, (size_t)n * p1));
    LT_TRY(scratch.alloc(&G, (size_t)n * p2));
    LT_TRY(scratch.alloc(&R2, (size_t)n * p2));
    LT_TRY(scratch.alloc(&Q, (size_t)n * c));
    LT_TRY(gemm_f32(X, ldx, nullptr, W1, h1, 0, nullptr, 0, P, p1, n, h1, f, 0, s));
    LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, P, p1, h1, b1, 1, R1, p1, s));
    LT_TRY(gemm_f32(R1, p1, nullptr, W2, h2, 0, nullptr, 0, G, p2, n, h2, h1, 0, s));
    LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, G, p2, h2, b2, 1, R2, p2, s));
    LT_TRY(gemm_f32(R2, p2, nullptr, W3, c, 0, nullptr, 0, Q, c, n, c, h2, 0, s));
    LT_TRY(spmm_csr_f32(indptr, indices, values, n, nullptr, n, Q, c, c, b3, 0, logits, ldz, s));
    return LT_OK;
}
