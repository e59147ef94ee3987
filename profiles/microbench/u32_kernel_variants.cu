// Occupancy / polynomials-per-thread sweep of the single-CTA u32 (30-bit Harvey) forward kernel, n = 2048.
// The kernel body is ntt_fast_fwd_kernel's (csrc/ntt_fast.cuh) with the launch bounds as template arguments.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -lineinfo -I../../tfhe-rs-main_modified_b200/csrc -o u32_kernel_variants.bin u32_kernel_variants.cu
#include <cstdio>
#include <vector>

#include "ntt_fast.cuh"

using namespace nttb200;
using A = Shoup<uint32_t, true>;

template <int LOGN, int POLYS, int PPT, int MINB, bool INV>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly* POLYS, MINB)
    probe_kernel(uint32_t* __restrict__ data, size_t rows, const A::TW* __restrict__ tw, A::Ctx c) {
    using T = uint32_t;
    using S = FastShape<LOGN>;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];
    T* smem = reinterpret_cast<T*>(fast_smem_raw);
    const unsigned t = threadIdx.x;
    const size_t grp = (size_t)blockIdx.x * POLYS + threadIdx.y;
    const SubPoly sub{0u, 0u};
    T* s = smem + threadIdx.y * PPT * S::kPaddedElems;
    T x[PPT][8];
    T* g[PPT];
#pragma unroll
    for (int pp = 0; pp < PPT; ++pp) {
        size_t row = grp * PPT + pp;
        g[pp] = data + ((row < rows ? row : rows - 1) << LOGN);
        if (!INV) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[pp][k] = g[pp][t + k * S::kThreadsPerPoly];
        } else {
            load8_consecutive(g[pp] + 8 * t, x[pp]);
        }
    }
    if (!INV) {
        fwd_from_regs<A, LOGN, PPT>(x, s, t, tw, c, sub);
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[pp][k] = A::fwd_fin(c, x[pp][k]);
            store8_consecutive(g[pp] + 8 * t, x[pp]);
        }
    } else {
        inv_to_regs<A, LOGN, PPT>(x, s, t, tw, c, sub);
#pragma unroll
        for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
            for (int k = 0; k < 8; ++k) g[pp][t + k * S::kThreadsPerPoly] = (k < 4) ? A::inv_fin(c, x[pp][k]) : A::inv_fin_prod(c, x[pp][k]);
    }
}

template <int LOGN, int POLYS, int PPT, int MINB, bool INV>
void run(uint32_t* d, size_t batch, const A::TW* tw, A::Ctx c) {
    constexpr size_t smem = (size_t)POLYS * PPT * FastShape<LOGN>::kPaddedElems * 4;
    auto kern = probe_kernel<LOGN, POLYS, PPT, MINB, INV>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, POLYS);
    unsigned grid = (unsigned)(batch / (POLYS * PPT));
    for (int w = 0; w < 3; ++w) kern<<<grid, block, smem>>>(d, batch, tw, c);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) kern<<<grid, block, smem>>>(d, batch, tw, c);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, block.x * block.y, smem);
    cudaError_t err = cudaDeviceSynchronize();
    printf("%s n=%d polys/CTA %d PPT %d min CTAs/SM %d: %3d regs, %3zu B local, %d CTAs/SM resident (%4d threads): %.3f ms  %.1f M NTT/s  %.1f %% of HBM copy peak  %s\n",
           INV ? "inv" : "fwd", 1 << LOGN, POLYS, PPT, MINB, fa.numRegs, (size_t)fa.localSizeBytes, occ, occ * block.x * block.y, ms,
           batch / ms / 1e3, 2.0 * batch * (4 << LOGN) / (ms * 1e-3) / 6543.4e9 * 100, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

template <int LOGN, int PPT, bool INV>
void run_shipped(uint32_t* d, size_t batch, const A::TW* tw, A::Ctx c) {
    constexpr int POLYS = (1 << LOGN) >= 2048 ? 1 : 2048 / (1 << LOGN);  // FastPolys
    constexpr size_t smem = (size_t)POLYS * PPT * FastShape<LOGN>::kPaddedElems * 4;
    auto kern = INV ? ntt_fast_inv_kernel<A, LOGN, POLYS, PPT, false> : ntt_fast_fwd_kernel<A, LOGN, POLYS, PPT, false>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    batch = (size_t)131072 * 2048 >> LOGN;  // 1 GiB of u32 whatever the length
    unsigned grid = (unsigned)(batch / (PPT * POLYS));
    dim3 block(FastShape<LOGN>::kThreadsPerPoly, POLYS);
    for (int w = 0; w < 3; ++w) kern<<<grid, block, smem>>>(d, batch, 0u, tw, c);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) kern<<<grid, block, smem>>>(d, batch, 0u, tw, c);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 10;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    printf("shipped %s kernel n=%d PPT %d: %3d regs, %3zu B local: %.3f ms  %.1f M NTT/s  %.1f %% of HBM copy peak\n", INV ? "inv" : "fwd",
           1 << LOGN, PPT, fa.numRegs, (size_t)fa.localSizeBytes, ms, batch / ms / 1e3, 2.0 * batch * (4 << LOGN) / (ms * 1e-3) / 6543.4e9 * 100);
}

int main() {
    constexpr int LOGN = 11;
    const size_t n = 1 << LOGN, batch = 131072;
    const uint32_t p = 1073479681u;
    uint32_t* d;
    cudaMalloc(&d, batch * n * 4);
    std::vector<uint32_t> h(batch * n);
    uint64_t s = 88172645463325252ull;
    for (auto& v : h) {
        s ^= s << 13, s ^= s >> 7, s ^= s << 17;
        v = (uint32_t)(s % p);
    }
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    // timing only: any table of values below p exercises the same instructions
    std::vector<A::TW> tw(n);
    for (size_t i = 0; i < n; ++i) {
        uint32_t w = (uint32_t)((i * 2654435761ull + 12345) % p);
        tw[i] = A::TW{w, (uint32_t)(((uint64_t)w << 32) / p)};
    }
    A::TW* dtw;
    cudaMalloc(&dtw, n * sizeof(A::TW));
    cudaMemcpy(dtw, tw.data(), n * sizeof(A::TW), cudaMemcpyHostToDevice);
    A::Ctx c{};
    c.p = p, c.two_p = 2 * p;
    run<LOGN, 1, 2, 6, false>(d, batch, dtw, c);
    run<LOGN, 1, 2, 5, false>(d, batch, dtw, c);
    run<LOGN, 1, 2, 4, false>(d, batch, dtw, c);
    run<LOGN, 1, 2, 8, false>(d, batch, dtw, c);
    run<LOGN, 2, 2, 3, false>(d, batch, dtw, c);
    run<LOGN, 2, 2, 2, false>(d, batch, dtw, c);
    run<LOGN, 1, 4, 4, false>(d, batch, dtw, c);
    run<LOGN, 1, 4, 5, false>(d, batch, dtw, c);
    run<LOGN, 1, 4, 3, false>(d, batch, dtw, c);
    run<LOGN, 2, 4, 2, false>(d, batch, dtw, c);
    run<LOGN, 1, 1, 8, false>(d, batch, dtw, c);
    run<LOGN, 1, 2, 6, true>(d, batch, dtw, c);
    run<LOGN, 1, 2, 5, true>(d, batch, dtw, c);
    run<LOGN, 1, 4, 4, true>(d, batch, dtw, c);
    run<LOGN, 1, 4, 5, true>(d, batch, dtw, c);
    run_shipped<LOGN, 2, false>(d, batch, dtw, c);
    run_shipped<LOGN, 4, false>(d, batch, dtw, c);
    run_shipped<LOGN, 2, true>(d, batch, dtw, c);
    run_shipped<LOGN, 4, true>(d, batch, dtw, c);
    // other lengths (the twiddle table above covers 2048 entries; 4096 reuses it cyclically through the same
    // pointer range only for timing, so give it its own)
    std::vector<A::TW> tw2(4096);
    for (size_t i = 0; i < 4096; ++i) tw2[i] = tw[i & 2047];
    A::TW* dtw2;
    cudaMalloc(&dtw2, 4096 * sizeof(A::TW));
    cudaMemcpy(dtw2, tw2.data(), 4096 * sizeof(A::TW), cudaMemcpyHostToDevice);
    run_shipped<10, 2, false>(d, batch, dtw2, c);
    run_shipped<10, 4, false>(d, batch, dtw2, c);
    run_shipped<10, 2, true>(d, batch, dtw2, c);
    run_shipped<10, 4, true>(d, batch, dtw2, c);
    run_shipped<12, 2, false>(d, batch, dtw2, c);
    run_shipped<12, 4, false>(d, batch, dtw2, c);
    run_shipped<12, 2, true>(d, batch, dtw2, c);
    run_shipped<12, 4, true>(d, batch, dtw2, c);
    run_shipped<9, 1, false>(d, batch, dtw2, c);
    run_shipped<9, 2, false>(d, batch, dtw2, c);
    run_shipped<9, 4, false>(d, batch, dtw2, c);
    run_shipped<8, 1, false>(d, batch, dtw2, c);
    run_shipped<8, 2, false>(d, batch, dtw2, c);
    return 0;
}
