// Error reporting, device facts and the optional allocator / copy helpers of the C ABI.
#include "common.cuh"
#include <string.h>

namespace fdtd {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return FDTD_ECUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace fdtd

extern "C" {

const char *fdtd_last_error(void) { return fdtd::g_err; }

int fdtd_version(void) { return 100; }

int fdtd_device_info(int *sm, size_t *free_bytes, size_t *total_bytes) {
    int dev = 0;
    FDTD_CUDA(cudaGetDevice(&dev));
    int n = 0;
    FDTD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    size_t f = 0, t = 0;
    FDTD_CUDA(cudaMemGetInfo(&f, &t));
    if (sm) *sm = n;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return FDTD_OK;
}

int fdtd_malloc(void **dptr, size_t bytes) {
    FDTD_REQUIRE(dptr != nullptr, "fdtd_malloc: null out pointer");
    FDTD_CUDA(cudaMalloc(dptr, bytes));
    return FDTD_OK;
}

int fdtd_free(void *dptr) {
    FDTD_CUDA(cudaFree(dptr));
    return FDTD_OK;
}

int fdtd_memset0(void *dptr, size_t bytes, void *stream) {
    FDTD_CUDA(cudaMemsetAsync(dptr, 0, bytes, fdtd::as_stream(stream)));
    return FDTD_OK;
}

int fdtd_upload(void *dptr, const void *hptr, size_t bytes, void *stream) {
    FDTD_CUDA(cudaMemcpyAsync(dptr, hptr, bytes, cudaMemcpyHostToDevice, fdtd::as_stream(stream)));
    return FDTD_OK;
}

int fdtd_download(void *hptr, const void *dptr, size_t bytes, void *stream) {
    FDTD_CUDA(cudaMemcpyAsync(hptr, dptr, bytes, cudaMemcpyDeviceToHost, fdtd::as_stream(stream)));
    return FDTD_OK;
}

int fdtd_ipc_export(const void *dptr, void *handle64) {
    FDTD_REQUIRE(dptr && handle64, "fdtd_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    FDTD_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), const_cast<void *>(dptr)));
    return FDTD_OK;
}

int fdtd_ipc_open(const void *handle64, void **mapped) {
    FDTD_REQUIRE(handle64 && mapped, "fdtd_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    // opened with the CONSUMER's device current: the mapping belongs to this device's context and peer access to
    // the exporting GPU is enabled on demand, so kernels of this device can store into it over NVLink
    FDTD_CUDA(cudaIpcOpenMemHandle(mapped, h, cudaIpcMemLazyEnablePeerAccess));
    return FDTD_OK;
}

int fdtd_ipc_close(void *mapped) {
    FDTD_CUDA(cudaIpcCloseMemHandle(mapped));
    return FDTD_OK;
}

int fdtd_enable_peer_access(int peer_device) {
    int dev = 0;
    FDTD_CUDA(cudaGetDevice(&dev));
    if (peer_device == dev) return FDTD_OK;
    int can = 0;
    FDTD_CUDA(cudaDeviceCanAccessPeer(&can, dev, peer_device));
    FDTD_REQUIRE(can, "device %d cannot access device %d over P2P", dev, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return FDTD_OK; }
    if (e != cudaSuccess) return fdtd::cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
    return FDTD_OK;
}

int fdtd_stream_sync(void *stream) {
    FDTD_CUDA(cudaStreamSynchronize(fdtd::as_stream(stream)));
    return FDTD_OK;
}

}  // extern "C"
