// comm.cu -- host side of the peer-memory exchange (see comm.cuh): allocation, CUDA IPC export / import, binding.
#include <cuda_runtime.h>
#include <string.h>
#include "common.cuh"
#include "comm.cuh"

namespace rcmarl {

struct CommHost {
    int rank, world;
    int64_t max_floats;
    void* base;                          // local allocation
    void* peer_base[COMM_MAX_WORLD];     // mapped peer allocations (peer_base[rank] == base)
    uint32_t seq;
    bool connected;
};

static CommHost* g_bound = nullptr;
CommHost* comm_bound() { return g_bound; }

static size_t data_bytes(const CommHost* c) { return sizeof(uint2) * 2 * (size_t)c->world * (size_t)c->max_floats; }
static size_t total_bytes(const CommHost* c) { return data_bytes(c) + sizeof(uint32_t) * 8; }

bool comm_next(CommDev* out, int64_t need_floats) { return comm_reserve(out, need_floats, 1u); }

bool comm_reserve(CommDev* out, int64_t need_floats, uint32_t count) {
    CommHost* c = g_bound;
    if (!c || !c->connected || need_floats > c->max_floats || count < 1) return false;
    const uint32_t first = c->seq + 1;
    c->seq += count;
    for (int p = 0; p < COMM_MAX_WORLD; ++p) out->cells[p] = p < c->world ? (uint2*)c->peer_base[p] : nullptr;
    out->error = (uint32_t*)((char*)c->base + data_bytes(c));
    out->max_floats = c->max_floats;
    out->rank = c->rank;
    out->world = c->world;
    out->seq = first;
    return true;
}

}  // namespace rcmarl

using namespace rcmarl;

extern "C" {

int rcmarl_comm_create(int rank, int world, int64_t max_floats, void** comm_out) {
    if (!comm_out || world < 1 || world > COMM_MAX_WORLD || rank < 0 || rank >= world || max_floats < 1) return RCMARL_ERR_ARG;
    CommHost* c = new CommHost();
    c->rank = rank; c->world = world; c->max_floats = max_floats; c->seq = 0; c->connected = false;
    for (int p = 0; p < COMM_MAX_WORLD; ++p) c->peer_base[p] = nullptr;
    if (cudaMalloc(&c->base, total_bytes(c)) != cudaSuccess) { delete c; last_cuda_error_set((int)cudaGetLastError()); return RCMARL_ERR_CUDA; }
    RC_CUDA(cudaMemset(c->base, 0, total_bytes(c)));
    RC_CUDA(cudaDeviceSynchronize());
    c->peer_base[rank] = c->base;
    *comm_out = c;
    return RCMARL_OK;
}

int rcmarl_comm_handle_bytes(void) { return (int)sizeof(cudaIpcMemHandle_t); }

int rcmarl_comm_export(void* comm, void* handle_out) {
    CommHost* c = (CommHost*)comm;
    if (!c || !handle_out) return RCMARL_ERR_ARG;
    cudaIpcMemHandle_t h;
    RC_CUDA(cudaIpcGetMemHandle(&h, c->base));
    memcpy(handle_out, &h, sizeof(h));
    return RCMARL_OK;
}

int rcmarl_comm_connect(void* comm, const void* handles) {
    CommHost* c = (CommHost*)comm;
    if (!c || !handles) return RCMARL_ERR_ARG;
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)p * sizeof(h), sizeof(h));
        RC_CUDA(cudaIpcOpenMemHandle(&c->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->connected = true;
    return RCMARL_OK;
}

int rcmarl_comm_bind(void* comm) {
    g_bound = (CommHost*)comm;             // NULL unbinds
    return RCMARL_OK;
}

int rcmarl_comm_error(void* comm) {
    CommHost* c = (CommHost*)comm;
    if (!c) return RCMARL_ERR_ARG;
    uint32_t e = 0;
    RC_CUDA(cudaMemcpy(&e, (char*)c->base + data_bytes(c), sizeof(e), cudaMemcpyDeviceToHost));
    return (int)e;
}

int rcmarl_comm_destroy(void* comm) {
    CommHost* c = (CommHost*)comm;
    if (!c) return RCMARL_ERR_ARG;
    if (g_bound == c) g_bound = nullptr;
    for (int p = 0; p < c->world; ++p)
        if (p != c->rank && c->peer_base[p]) cudaIpcCloseMemHandle(c->peer_base[p]);
    cudaFree(c->base);
    delete c;
    return RCMARL_OK;
}

}  // extern "C"
