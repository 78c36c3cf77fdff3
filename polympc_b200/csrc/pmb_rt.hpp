// pmb_rt.hpp — the few host-side runtime calls the engine needs (device buffers, copies, stream, events).
// nvcc build: thin wrappers over the CUDA runtime that record the first failure in a thread-local error string.
// -DPMB_EMU build (tests/warp_emu, test infrastructure only): "device" memory is host memory and copies are memcpy.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include "pmb_warp.hpp"

namespace pmb {

inline std::string& last_error_string() { static thread_local std::string s; return s; }

#if defined(__CUDACC__) && !defined(PMB_EMU)
typedef cudaStream_t stream_t;
typedef cudaEvent_t event_t;

inline bool rt_ok(cudaError_t e, const char* what)
{
    if (e == cudaSuccess) return true;
    last_error_string() = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define PMB_RT(call) ::pmb::rt_ok((call), #call)

inline int rt_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
inline bool rt_set_device(int dev) { return PMB_RT(cudaSetDevice(dev)); }
inline int rt_sm_count() { int dev = 0, n = 0; if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
inline void* rt_alloc(size_t bytes) { void* p = nullptr; if (bytes == 0) bytes = 16; if (!PMB_RT(cudaMalloc(&p, bytes))) return nullptr; return p; }
inline void rt_free(void* p) { if (p) cudaFree(p); }
inline bool rt_h2d(void* dst, const void* src, size_t bytes, stream_t s) { return bytes == 0 || PMB_RT(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s)); }
inline bool rt_d2h(void* dst, const void* src, size_t bytes, stream_t s) { return bytes == 0 || PMB_RT(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s)); }
inline bool rt_d2d(void* dst, const void* src, size_t bytes, stream_t s) { return bytes == 0 || PMB_RT(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s)); }
inline bool rt_memset(void* dst, int v, size_t bytes, stream_t s) { return bytes == 0 || PMB_RT(cudaMemsetAsync(dst, v, bytes, s)); }
inline bool rt_sync(stream_t s) { return PMB_RT(cudaStreamSynchronize(s)); }
inline bool rt_stream_create(stream_t* s) { return PMB_RT(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking)); }
inline void rt_stream_destroy(stream_t s) { if (s) cudaStreamDestroy(s); }
inline bool rt_event_create(event_t* e) { return PMB_RT(cudaEventCreate(e)); }
inline void rt_event_destroy(event_t e) { if (e) cudaEventDestroy(e); }
inline bool rt_event_record(event_t e, stream_t s) { return PMB_RT(cudaEventRecord(e, s)); }
inline float rt_event_ms(event_t a, event_t b) { float ms = 0; if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) { cudaGetLastError(); return 0; } return ms; }
inline void* rt_host_alloc(size_t bytes) { void* p = nullptr; if (!PMB_RT(cudaMallocHost(&p, bytes ? bytes : 16))) return nullptr; return p; }
inline void rt_host_free(void* p) { if (p) cudaFreeHost(p); }
template <class Body, class... Args>
inline bool rt_launch(int grid, size_t smem, stream_t s, Args... args) { return rt_ok(launch<Body, Args...>(grid, smem, s, args...), Body::NAME); }
#else
typedef void* stream_t;
typedef void* event_t;
inline int rt_device_count() { return 1; }
inline bool rt_set_device(int) { return true; }
inline int rt_sm_count() { return 1; }   // emulator: every CTA gets a different rotation, which exercises the rotated numbering
inline void* rt_alloc(size_t bytes) { return std::calloc(bytes ? bytes : 16, 1); }
inline void rt_free(void* p) { std::free(p); }
inline bool rt_h2d(void* dst, const void* src, size_t bytes, stream_t) { if (bytes) std::memcpy(dst, src, bytes); return true; }
inline bool rt_d2h(void* dst, const void* src, size_t bytes, stream_t) { if (bytes) std::memcpy(dst, src, bytes); return true; }
inline bool rt_d2d(void* dst, const void* src, size_t bytes, stream_t) { if (bytes) std::memmove(dst, src, bytes); return true; }
inline bool rt_memset(void* dst, int v, size_t bytes, stream_t) { if (bytes) std::memset(dst, v, bytes); return true; }
inline bool rt_sync(stream_t) { return true; }
inline bool rt_stream_create(stream_t* s) { *s = nullptr; return true; }
inline void rt_stream_destroy(stream_t) {}
inline bool rt_event_create(event_t* e) { *e = nullptr; return true; }
inline void rt_event_destroy(event_t) {}
inline bool rt_event_record(event_t, stream_t) { return true; }
inline float rt_event_ms(event_t, event_t) { return 0.f; }
inline void* rt_host_alloc(size_t bytes) { return std::calloc(bytes ? bytes : 16, 1); }
inline void rt_host_free(void* p) { std::free(p); }
template <class Body, class... Args>
inline bool rt_launch(int grid, size_t smem, stream_t s, Args... args) { launch<Body, Args...>(grid, smem, s, args...); return true; }
#endif

/** RAII device buffer */
template <class T>
struct DevBuf {
    T* p = nullptr; size_t n = 0;
    DevBuf() {}
    DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { rt_free(p); }
    bool resize(size_t count) { if (count <= n && p) return true; rt_free(p); p = (T*)rt_alloc(count * sizeof(T)); n = p ? count : 0; return p != nullptr; }
    size_t bytes() const { return n * sizeof(T); }
};

} // namespace pmb
