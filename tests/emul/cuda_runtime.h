// TEST INFRASTRUCTURE ONLY (tests/emul). A minimal stand-in for <cuda_runtime.h> that lets the library's kernel
// functors and host orchestration be compiled with g++ and executed sequentially on the CPU, so that index
// arithmetic and the host control flow can be debugged in a container without a GPU. It is found only because
// tests/emul/build.sh puts this directory first on the include path and defines TROYB_HOST_EMUL; the product
// library (troy_b200/libtroy_b200.so) is always built by nvcc against the real CUDA runtime and has no CPU path.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __align__(x)
#define __launch_bounds__(x)

struct ulonglong2 { unsigned long long x, y; };
typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaMemPool_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaMemPoolAttrReleaseThreshold = 0 };

inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaSetDevice(int) { return 0; }
inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
enum { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int *v, int, int) { *v = 148; return 0; }
inline cudaError_t cudaFree(void *p) { free(p); return 0; }
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 8); return 0; }
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { *p = malloc(n ? n : 8); return 0; }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { free(p); return 0; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { memset(d, v, n); return 0; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = nullptr; return 0; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, int, void *) { return 0; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((char *)d + r * dp, (const char *)s + r * sp, w);
    return 0;
}
inline cudaError_t cudaMemset2DAsync(void *d, size_t dp, int v, size_t w, size_t h, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memset((char *)d + r * dp, v, w);
    return 0;
}
template <class T> inline T __ldg(const T *p) { return *p; }
inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
typedef void *cudaEvent_t;
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
