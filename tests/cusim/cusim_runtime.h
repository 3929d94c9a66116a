// cusim_runtime.h -- the handful of CUDA runtime calls the engine's host code makes, on the host, for TESTS ONLY.
//
// "Device" memory is host memory, streams and events are tokens, every launch runs to completion before it returns
// (cusim::launch), so stream order is program order.  With it the UNCHANGED host orchestration of osp_engine.cu --
// planning, capacity decisions, row blocks, hand-overs -- and the UNCHANGED kernels run end to end in the `-m "not gpu"`
// suite against the oracle.  Nothing under outerspace_b200/ links this; the product has no CPU path.
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

enum cudaError_t { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4, cudaMemPoolAttrReservedMemCurrent, cudaMemPoolAttrUsedMemCurrent };
constexpr unsigned cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocMapped = 2;

struct cusimStream { int id; };
struct cusimEvent { double ms; };
typedef cusimStream *cudaStream_t;
typedef cusimEvent *cudaEvent_t;
typedef int cudaMemPool_t;
struct cudaDeviceProp { int multiProcessorCount; size_t totalGlobalMem; int l2CacheSize; };

namespace cusim {
inline size_t &device_bytes() { static size_t b = 0; return b; }          // live "device" allocations
inline size_t device_total() {
    if (const char *e = std::getenv("CUSIM_DEVICE_MB")) return size_t(std::strtod(e, nullptr) * double(1 << 20));
    return size_t(4) << 30;
}
inline double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
// A "device" allocation: 256 bytes of header (its size), then the payload, 256-byte aligned like cudaMalloc's.  The block
// ends 16 bytes after the payload -- the slack a bulk copy may read, its size being rounded up to 16 bytes -- so that
// under AddressSanitizer (tools/fuzz_engine_sim.py --asan) any access of a kernel beyond that is reported.
constexpr size_t DEV_HEADER = 256, DEV_SLACK = 16;
inline cudaError_t dev_alloc(void **p, size_t bytes) {
    if (device_bytes() + bytes > device_total()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    void *raw = nullptr;
    if (posix_memalign(&raw, 256, DEV_HEADER + bytes + DEV_SLACK) != 0) { *p = nullptr; return cudaErrorMemoryAllocation; }
    static const int fill = [] { const char *e = std::getenv("CUSIM_FILL"); return e ? int(std::strtol(e, nullptr, 0)) & 0xFF : 0xCD; }();
    std::memset(raw, fill, DEV_HEADER + bytes + DEV_SLACK);                 // fresh device memory is garbage (CUSIM_FILL picks which)
    *static_cast<size_t *>(raw) = bytes;
    device_bytes() += bytes;
    *p = static_cast<char *>(raw) + DEV_HEADER;
    return cudaSuccess;
}
inline void dev_free(void *p) {
    if (!p) return;
    void *raw = static_cast<char *>(p) - DEV_HEADER;
    device_bytes() -= *static_cast<size_t *>(raw);
    std::free(raw);
}
}  // namespace cusim

inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : e == cudaErrorMemoryAllocation ? "out of memory" : "cusim error"; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = std::getenv("CUSIM_NO_DEVICE") ? 0 : 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    { const char *e = std::getenv("CUSIM_SMS"); p->multiProcessorCount = e ? std::max(1, std::atoi(e)) : 2; } p->totalGlobalMem = cusim::device_total(); p->l2CacheSize = 1 << 20;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new cusimStream{0}; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new cusimStream{0}; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = 0; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamQuery(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cusimEvent{0}; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->ms = cusim::now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = float(b->ms - a->ms); return cudaSuccess; }
inline cudaError_t cudaMalloc(void **p, size_t n) { return cusim::dev_alloc(p, n); }
template <class T> inline cudaError_t cudaMalloc(T **p, size_t n) { return cusim::dev_alloc(reinterpret_cast<void **>(p), n); }
inline cudaError_t cudaFree(void *p) { cusim::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cusim::dev_alloc(p, n); }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { cusim::dev_free(p); return cudaSuccess; }
inline cudaError_t cudaMallocHost(void **p, size_t n) { *p = std::calloc(1, n); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaHostAlloc(void **p, size_t n, unsigned) { return cudaMallocHost(p, n); }
inline cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostGetDevicePointer(void **d, void *h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *t = cusim::device_total(); *f = *t - std::min(*t, cusim::device_bytes()); return cudaSuccess; }
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) { *p = 0; return cudaSuccess; }
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
inline cudaError_t cudaMemPoolGetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *v) { *static_cast<unsigned long long *>(v) = 0; return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 1; return cudaSuccess; }
