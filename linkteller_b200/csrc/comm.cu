// lt_comm — the multi-GPU side of the efficient attack without a collective call on the data path.
//
// The reference has no multi-device code at all (no torch.distributed / NCCL / MPI call site anywhere in the tree);
// SURVEY.md §8e defines the partitioning: the graph is replicated, rank g owns a slice of the perturbed nodes, and every
// rank must end up with the full T x T influence matrix.  Round 1 did that with one NCCL all-gather of the DENSE
// matrix (268 MB at T = 8192, of which 99.5 % are exact zeros) — 0.45 ms on top of a 1 ms attack.  Here every rank owns one
// symmetric buffer (the full matrix + a page of flags), allocated with cudaMalloc and mapped into every other rank of the
// node through CUDA IPC.  The reduce kernel of the attack (k_frontier_reduce) then stores each influence entry it
// produces into ALL ranks' matrices — plain st.global on the mapped peer pointers, carried by NVLink / NVSwitch — so
// the "all-gather" is the epilogue of the compute kernel and moves only the ~40 K support entries per rank.
//
// Ordering (per attack, epoch e; flags are written with system-scope fences and read with volatile loads):
//   1. every rank zero-fills its own matrix (side stream) and writes zeroed[rank] = e into every peer's flag block;
//   2. before its reduce kernel a rank waits until zeroed[r] ≥ e for all r (nobody's matrix is still being cleared);
//   3. after its reduce kernel it writes written[rank] = e into every peer's flag block (kernel boundary + fence.sys);
//   4. it waits until written[r] ≥ e for all r: the local matrix is complete.
// A rank cannot start clearing for epoch e+1 before step 4 of epoch e, and step 4 needs every peer's step 3, so no
// store of epoch e can land in a matrix that is already being reused.  Waits are bounded (≈ 2 s) and report a timeout
// instead of hanging the GPU.
#include "common.cuh"

#include <string.h>

namespace lt {

__global__ void k_comm_signal(lt_comm c, int which, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= c.world) return;
    __threadfence_system();                                    // everything this stream wrote before is visible first
    volatile uint32_t* f = c.flags[r] + which * LT_MAX_RANKS + c.rank;
    *f = epoch;
    __threadfence_system();
}

__global__ void k_comm_wait(lt_comm c, int which, uint32_t epoch) {
    const int r = threadIdx.x;
    if (r >= c.world) return;
    volatile const uint32_t* f = c.flags[c.rank] + which * LT_MAX_RANKS + r;
    const long long t0 = clock64();
    while ((int32_t)(*f - epoch) < 0) {
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) { atomicExch(c.timeout_flag, 1); break; }    // ≈ 2 s at 2 GHz: report, do not hang
    }
    __threadfence_system();
}

int comm_signal(lt_comm* c, int which, cudaStream_t s) {
    k_comm_signal<<<1, 32, 0, s>>>(*c, which, c->epoch);
    LT_LAUNCH_CHECK();
    return LT_OK;
}
int comm_wait(lt_comm* c, int which, cudaStream_t s) {
    k_comm_wait<<<1, 32, 0, s>>>(*c, which, c->epoch);
    LT_LAUNCH_CHECK();
    return LT_OK;
}

}  // namespace lt

using namespace lt;

static size_t comm_payload_aligned(size_t bytes) { return (bytes + 4095) / 4096 * 4096; }

extern "C" int lt_comm_create(int rank, int world, int64_t bytes, int64_t aux_bytes, lt_comm** comm_host) {
    LT_CHECK_ARG(comm_host && world >= 1 && world <= LT_MAX_RANKS && rank >= 0 && rank < world && bytes > 0 && aux_bytes >= 0,
                 "lt_comm_create: need 0 <= rank < world <= 8, bytes > 0, aux_bytes >= 0");
    lt_comm* c = new lt_comm();
    c->rank = rank; c->world = world; c->bytes = (size_t)bytes; c->aux_bytes = (size_t)aux_bytes;
    const size_t mat = comm_payload_aligned((size_t)bytes);
    const size_t pay = mat + comm_payload_aligned((size_t)aux_bytes);
    void* base = nullptr;
    cudaError_t e = cudaMalloc(&base, pay + 4096);
    if (e != cudaSuccess) { set_error("lt_comm_create: cudaMalloc(%zu) failed: %s", pay + 4096, cudaGetErrorString(e)); delete c; return LT_ERR_NOMEM; }
    cudaMemset((char*)base + pay, 0, 4096);
    c->peer[rank] = base;
    c->aux[rank] = (char*)base + mat;
    c->flags[rank] = reinterpret_cast<uint32_t*>((char*)base + pay);
    c->timeout_flag = reinterpret_cast<int*>((char*)base + pay + 2048);
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, base);
    if (e != cudaSuccess) {
        set_error("lt_comm_create: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
        cudaFree(base); delete c; return LT_ERR_CUDA;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(c->handle, &h, 64);
    if (world == 1) c->opened = true;
    *comm_host = c;
    return LT_OK;
}

extern "C" int lt_comm_get_handle(const lt_comm* comm, void* handle64_host) {
    LT_CHECK_ARG(comm && handle64_host, "lt_comm_get_handle");
    memcpy(handle64_host, comm->handle, 64);
    return LT_OK;
}

// all_handles_host: world x 64 bytes, rank order (the caller exchanges them, e.g. torch.distributed.all_gather_object)
extern "C" int lt_comm_open(lt_comm* comm, const void* all_handles_host) {
    LT_CHECK_ARG(comm && all_handles_host, "lt_comm_open");
    const size_t mat = comm_payload_aligned(comm->bytes);
    const size_t pay = mat + comm_payload_aligned(comm->aux_bytes);
    for (int r = 0; r < comm->world; ++r) {
        if (r == comm->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)all_handles_host + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("lt_comm_open: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
            return LT_ERR_CUDA;
        }
        comm->peer[r] = p;
        comm->aux[r] = (char*)p + mat;
        comm->flags[r] = reinterpret_cast<uint32_t*>((char*)p + pay);
    }
    comm->opened = true;
    return LT_OK;
}

extern "C" void* lt_comm_buffer(const lt_comm* comm) { return comm ? comm->peer[comm->rank] : nullptr; }

// 1 when a bounded wait gave up since the last call (a peer never arrived); clears the flag.  Synchronises the device.
extern "C" int lt_comm_timed_out(lt_comm* comm, int* flag_host) {
    LT_CHECK_ARG(comm && flag_host, "lt_comm_timed_out");
    int v = 0;
    LT_CUDA(cudaMemcpy(&v, comm->timeout_flag, 4, cudaMemcpyDeviceToHost));
    if (v) LT_CUDA(cudaMemset(comm->timeout_flag, 0, 4));
    *flag_host = v;
    return LT_OK;
}

extern "C" int lt_comm_destroy(lt_comm* comm) {
    if (!comm) return LT_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < comm->world; ++r)
        if (r != comm->rank && comm->peer[r]) cudaIpcCloseMemHandle(comm->peer[r]);
    if (comm->peer[comm->rank]) cudaFree(comm->peer[comm->rank]);
    delete comm;
    return LT_OK;
}
