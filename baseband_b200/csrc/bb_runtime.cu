// Runtime part of the C ABI: errors, device/stream/memory helpers.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "bb_runtime.cuh"

namespace bb {

static thread_local char g_error[512] = "";

int set_error(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t err, const char *what) {
    if (err == cudaSuccess) return BB_OK;
    return set_error(BB_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(err),
                     cudaGetErrorString(err));
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)
            != cudaSuccess || n <= 0) n = 148;
        cached = n;
        cached_dev = dev;
    }
    return cached;
}

int grid_override() {
    static int value = [] {
        const char *e = getenv("BB_CTAS_PER_SM");
        return e ? atoi(e) : -1;
    }();
    return value;
}

}  // namespace bb

using bb::as_stream;
using bb::check_cuda;

extern "C" {

int bb_abi_version(void) { return BB_ABI_VERSION; }

const char *bb_last_error(void) { return bb::g_error; }

int bb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bb_set_device(int device) {
    return check_cuda(cudaSetDevice(device), "cudaSetDevice");
}

int bb_device_sm_count(int device) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device)
        != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int bb_malloc(void **ptr, int64_t nbytes) {
    return check_cuda(cudaMalloc(ptr, (size_t)nbytes), "cudaMalloc");
}
int bb_free(void *ptr) { return check_cuda(cudaFree(ptr), "cudaFree"); }
int bb_host_alloc(void **ptr, int64_t nbytes) {
    return check_cuda(cudaHostAlloc(ptr, (size_t)nbytes, cudaHostAllocDefault),
                      "cudaHostAlloc");
}
int bb_host_free(void *ptr) {
    return check_cuda(cudaFreeHost(ptr), "cudaFreeHost");
}
int bb_host_register(void *ptr, int64_t nbytes) {
    return check_cuda(cudaHostRegister(ptr, (size_t)nbytes,
                                       cudaHostRegisterDefault),
                      "cudaHostRegister");
}
int bb_host_unregister(void *ptr) {
    return check_cuda(cudaHostUnregister(ptr), "cudaHostUnregister");
}
int bb_memcpy_h2d(void *dst, const void *src, int64_t nbytes, void *stream) {
    return check_cuda(cudaMemcpyAsync(dst, src, (size_t)nbytes,
                                      cudaMemcpyHostToDevice,
                                      as_stream(stream)), "memcpy h2d");
}
int bb_memcpy_d2h(void *dst, const void *src, int64_t nbytes, void *stream) {
    return check_cuda(cudaMemcpyAsync(dst, src, (size_t)nbytes,
                                      cudaMemcpyDeviceToHost,
                                      as_stream(stream)), "memcpy d2h");
}
int bb_memset(void *dst, int value, int64_t nbytes, void *stream) {
    return check_cuda(cudaMemsetAsync(dst, value, (size_t)nbytes,
                                      as_stream(stream)), "memset");
}
int bb_stream_create(void **stream) {
    cudaStream_t s;
    int rc = check_cuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking),
                        "cudaStreamCreate");
    if (rc == BB_OK) *stream = s;
    return rc;
}
int bb_stream_destroy(void *stream) {
    return check_cuda(cudaStreamDestroy(as_stream(stream)),
                      "cudaStreamDestroy");
}
int bb_stream_synchronize(void *stream) {
    return check_cuda(cudaStreamSynchronize(as_stream(stream)),
                      "cudaStreamSynchronize");
}

}  // extern "C"
