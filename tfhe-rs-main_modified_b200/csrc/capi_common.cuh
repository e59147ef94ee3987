// Shared pieces of the C ABI translation units.
#pragma once
#include <atomic>
#include <memory>
#include <string>
#include <vector>

#include "../../include/tfhe_ntt_b200.h"
#include "ntt_engine.cuh"

struct ntt_b200_plan64 {
    std::shared_ptr<nttb200::PrimePlan> impl;
};
struct ntt_b200_plan32 {
    std::shared_ptr<nttb200::PrimePlan> impl;
};

namespace nttb200 {
extern thread_local std::string g_last_error;

// Runs f, mapping exceptions to status codes (CUDA errors keep their text for
// ntt_b200_last_error()).
template <class F>
int guarded(F&& f) {
    try {
        return f();
    } catch (const CudaError& e) {
        g_last_error = e.what();
        cudaGetLastError();
        return NTT_B200_ERR_CUDA;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return NTT_B200_ERR_CUDA;
    }
}

// Stream-ordered allocations of the host-pointer entry points come from the device's default
// memory pool; keep freed blocks cached in the pool (the default threshold of 0 hands them back
// to the driver at every synchronisation, which costs milliseconds per staging buffer).
inline void keep_pool_cached(int device) {
    static std::atomic<bool> done[64] = {};  // idempotent work, flags shared by concurrent host threads
    if (device < 0) return;
    if (device < 64 && done[device].load(std::memory_order_acquire)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    if (device < 64) done[device].store(true, std::memory_order_release);
}

// A non-blocking stream with its stream-ordered scratch buffers, for the host-pointer entry points.
// Whatever way the scope is left -- finish(), an error status or an exception thrown by
// NTT_CUDA_CHECK -- the destructor frees the buffers in stream order, waits for every copy that may
// still touch the caller's host buffers and destroys the stream, so nothing is in flight and nothing
// leaks when the call returns.
struct ScopedStream {
    cudaStream_t st = nullptr;
    std::vector<void*> bufs;
    ScopedStream() = default;
    ScopedStream(const ScopedStream&) = delete;
    ScopedStream& operator=(const ScopedStream&) = delete;
    void open() {
        if (!st) NTT_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    }
    void* alloc(size_t bytes) {
        void* d = nullptr;
        NTT_CUDA_CHECK(cudaMallocAsync(&d, bytes ? bytes : 1, st));
        bufs.push_back(d);
        return d;
    }
    // normal exit: first error of the stream's work (the buffers are released either way)
    cudaError_t finish() {
        if (!st) return cudaSuccess;
        for (void* d : bufs) cudaFreeAsync(d, st);
        bufs.clear();
        cudaError_t e = cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        st = nullptr;
        return e;
    }
    ~ScopedStream() {
        if (finish() != cudaSuccess) cudaGetLastError();
    }
};
struct ScopedEvent {
    cudaEvent_t ev = nullptr;
    void create() { NTT_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); }
    ~ScopedEvent() {
        if (ev) cudaEventDestroy(ev);
    }
};
// One non-blocking stream per host thread and device for the host-pointer entry points of the CRT and
// product plans (creating and destroying a stream per call costs more than a small transform).
// Thread-local, so concurrent callers never share a stream; never destroyed.
inline cudaStream_t cached_stream(int device) {
    thread_local std::vector<cudaStream_t> st;
    if (device < 0) return nullptr;
    if ((size_t)device >= st.size()) st.resize((size_t)device + 1, nullptr);
    if (!st[device]) {
        int prev = 0;
        cudaGetDevice(&prev);
        if (prev != device) cudaSetDevice(device);
        cudaError_t e = cudaStreamCreateWithFlags(&st[device], cudaStreamNonBlocking);
        if (prev != device) cudaSetDevice(prev);
        if (e != cudaSuccess) {
            st[device] = nullptr;
            throw CudaError(std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(e));
        }
    }
    return st[device];
}
}  // namespace nttb200
