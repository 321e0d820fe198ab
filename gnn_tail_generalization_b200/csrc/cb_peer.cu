// Peer-visible device buffers for the node-sliced multi-GPU path (one process per GPU, one node).
//
// The reference is single-device; this is the exchange step of the 1-D node partition.  A
// buffer allocated here can be opened by the other ranks' processes (CUDA IPC, NVLink/NVSwitch peer
// access) so that the kernel producing a row block stores the rows its peers gather straight into the
// peers' copies (cb_gemm_rows / cb_gemm_rows_grad with a cb_peer_push_t) -- the all-gather rides on the
// GEMM epilogue tile by tile instead of following it.
#include <cstring>

#include "cb_internal.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == CB_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");

extern "C" {

int cb_peer_alloc(int64_t bytes, void** ptr, void* handle_out) {
    using namespace cb;
    CB_REQUIRE(ptr != nullptr && handle_out != nullptr && bytes > 0, CB_E_INVALID, "cb_peer_alloc: bad argument");
    *ptr = nullptr;
    void* p = nullptr;
    CB_CUDA(cudaMalloc(&p, (size_t)bytes));   // a dedicated allocation: IPC handles name whole allocations
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    std::memcpy(handle_out, &h, sizeof(h));
    *ptr = p;
    return CB_OK;
}

int cb_peer_open(const void* handle, void** ptr) {
    using namespace cb;
    CB_REQUIRE(handle != nullptr && ptr != nullptr, CB_E_INVALID, "cb_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    *ptr = nullptr;
    CB_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CB_OK;
}

int cb_peer_close(void* ptr) {
    using namespace cb;
    if (!ptr) return CB_OK;
    CB_CUDA(cudaIpcCloseMemHandle(ptr));
    return CB_OK;
}

int cb_peer_free(void* ptr) {
    using namespace cb;
    if (!ptr) return CB_OK;
    CB_CUDA(cudaFree(ptr));
    return CB_OK;
}

}  // extern "C"
