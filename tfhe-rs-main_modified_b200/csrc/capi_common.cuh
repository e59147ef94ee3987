// Shared pieces of the C ABI translation units.
#pragma once
#include <memory>
#include <string>

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
    static bool done[64] = {};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = true;
}
// One non-blocking stream per host thread and device for the host-pointer entry points of the CRT and
// product plans (creating and destroying a stream per call costs more than a small transform).
// Thread-local, so concurrent callers never share a stream; never destroyed.
inline cudaStream_t cached_stream(int device) {
    thread_local cudaStream_t st[16] = {};
    if (device < 0 || device >= 16) return nullptr;
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
