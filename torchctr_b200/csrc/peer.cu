// Peer memory for the row-sharded tables: one process per GPU, every rank maps the table shards, gradient
// buffers and routing lists of the other ranks into its own address space (CUDA IPC over NVLink / NVSwitch)
// and the lookup / update kernels load from those peer pointers directly -- there is no all-to-all.
//
// The reference is replicas-only (accelerate / DDP, torchctr/trainer.py:128-130); this is the B200-native
// replacement of that data-parallel layer for the embedding tables.
#include <string.h>

#include <map>
#include <mutex>

#include "common.cuh"

namespace ctr {
static std::mutex g_peer_mu;
static std::map<void *, int> g_peer_open;   // pointers obtained from cudaIpcOpenMemHandle -> device they were opened on
}  // namespace ctr

using namespace ctr;

static_assert(sizeof(cudaIpcMemHandle_t) == CTR_PEER_HANDLE_BYTES, "handle size");

// Device memory that other processes can map.  Plain cudaMalloc (not a pool / VMM allocation), zero-filled.
extern "C" int ctr_peer_alloc(int64_t bytes, void **ptr) {
    CTR_REQUIRE(ptr != nullptr && bytes > 0, "ctr_peer_alloc: bad arguments");
    void *p = nullptr;
    CTR_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaMemset");
    }
    *ptr = p;
    return CTR_OK;
}

extern "C" int ctr_peer_free(void *ptr) {
    if (ptr != nullptr) CTR_CUDA_OK(cudaFree(ptr));
    return CTR_OK;
}

extern "C" int ctr_peer_export(void *ptr, void *handle_out) {
    CTR_REQUIRE(ptr != nullptr && handle_out != nullptr, "ctr_peer_export: null pointer");
    cudaIpcMemHandle_t h;
    CTR_CUDA_OK(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle_out, &h, sizeof(h));
    return CTR_OK;
}

extern "C" int ctr_peer_open(const void *handle, void **ptr) {
    CTR_REQUIRE(handle != nullptr && ptr != nullptr, "ctr_peer_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void *p = nullptr;
    CTR_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    int dev = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> lock(g_peer_mu);
        g_peer_open[p] = dev;
    }
    *ptr = p;
    return CTR_OK;
}

extern "C" int ctr_peer_close(void *ptr) {
    if (ptr == nullptr) return CTR_OK;
    {
        std::lock_guard<std::mutex> lock(g_peer_mu);
        auto it = g_peer_open.find(ptr);
        CTR_REQUIRE(it != g_peer_open.end(), "ctr_peer_close: pointer was not opened by ctr_peer_open");
        g_peer_open.erase(it);
    }
    CTR_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return CTR_OK;
}
