// One CTA per 2^13-point polynomial (1024 threads, 72 KiB tile for u64 / 36 KiB for u32) against the shipped
// two-pass path (TMA-staged radix-16 pass + 512-point kernels): the north star's "polynomials up to N = 16384
// in one CTA" measured.  Timing only (synthetic twiddle values; the instruction stream does not depend on them).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -I../../tfhe-rs-main_modified_b200/csrc -o one_cta_8192.bin one_cta_8192.cu
#include <cstdio>
#include <vector>

#include "ntt_fast.cuh"

using namespace nttb200;

template <class A, int LOGN, int PPT, bool INV>
void run(const char* name, size_t batch) {
    using T = typename A::T;
    const size_t n = size_t(1) << LOGN;
    T* d;
    cudaMalloc(&d, batch * n * sizeof(T));
    cudaMemset(d, 1, batch * n * sizeof(T));
    std::vector<typename A::TW> tw(n);
    unsigned char* raw = reinterpret_cast<unsigned char*>(tw.data());
    for (size_t i = 0; i < n * sizeof(typename A::TW); ++i) raw[i] = (unsigned char)(i * 37 + 11) & 0x3F;
    typename A::TW* dtw;
    cudaMalloc(&dtw, n * sizeof(typename A::TW));
    cudaMemcpy(dtw, tw.data(), n * sizeof(typename A::TW), cudaMemcpyHostToDevice);
    typename A::Ctx c{};
    c.p = (T)((sizeof(T) == 8) ? 0xFFFFFFFF00000001ull : 1073479681ull);
    constexpr size_t smem = (size_t)PPT * FastShape<LOGN>::kPaddedElems * sizeof(T);
    auto kern = INV ? ntt_fast_inv_kernel<A, LOGN, 1, PPT, false> : ntt_fast_fwd_kernel<A, LOGN, 1, PPT, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    unsigned grid = (unsigned)(batch / PPT);
    for (int w = 0; w < 3; ++w) kern<<<grid, FastShape<LOGN>::kThreadsPerPoly, smem>>>(d, batch, 0u, dtw, c);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) kern<<<grid, FastShape<LOGN>::kThreadsPerPoly, smem>>>(d, batch, 0u, dtw, c);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, FastShape<LOGN>::kThreadsPerPoly, smem);
    cudaError_t err = cudaDeviceSynchronize();
    printf("%-34s n=%zu PPT %d: %3d regs, %zu B local, %d CTA(s)/SM: %.3f ms  %.2f M NTT/s  %.1f %% of the HBM copy peak %s\n", name, n, PPT,
           fa.numRegs, (size_t)fa.localSizeBytes, occ, ms, batch / ms / 1e3, 2.0 * batch * n * sizeof(T) / (ms * 1e-3) / 6543.4e9 * 100,
           err == cudaSuccess ? "" : cudaGetErrorString(err));
    cudaFree(d);
    cudaFree(dtw);
}

int main() {
    run<Solinas64, 13, 1, false>("u64 Solinas fwd, one CTA", 16384);
    run<Solinas64, 13, 1, true>("u64 Solinas inv, one CTA", 16384);
    run<Shoup<uint32_t, true>, 13, 1, false>("u32 30-bit fwd, one CTA", 32768);
    run<Shoup<uint32_t, true>, 13, 2, false>("u32 30-bit fwd, one CTA", 32768);
    run<Shoup<uint32_t, true>, 13, 1, true>("u32 30-bit inv, one CTA", 32768);
    return 0;
}
