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
}  // namespace nttb200
