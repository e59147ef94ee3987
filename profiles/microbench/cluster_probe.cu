// Timing probe for the cluster-of-eight large-N kernels (developer tool, not product code):
// variants of the forward kernel with the DSMEM scatter or the cluster barriers removed show
// where the time goes.  Results are NOT checked here (the variants are wrong on purpose).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I tfhe-rs-main_modified_b200/csrc -I include -o cluster_probe profiles/microbench/cluster_probe.cu
#include <cstdio>
#include <vector>
#include "ntt_fast.cuh"
using namespace nttb200;

template <class A, int LOGSUB, int VARIANT>
__global__ void __cluster_dims__(8, 1, 1)
    __launch_bounds__(FastShape<LOGSUB>::kThreadsPerPoly, FastMinBlocks<A, FastShape<LOGSUB>::kThreadsPerPoly>::value)
        probe_fwd(typename A::T* __restrict__ data, const typename A::TW* __restrict__ tw, typename A::Ctx c) {
    using T = typename A::T;
    using S = FastShape<LOGSUB>;
    constexpr unsigned TPP = S::kThreadsPerPoly, SUB = 1u << LOGSUB;
    extern __shared__ __align__(16) unsigned char fast_smem_raw[];
    T* tile = reinterpret_cast<T*>(fast_smem_raw);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank(), t = threadIdx.x;
    T* g = data + ((size_t)(blockIdx.x >> 3) << (LOGSUB + 3));
    T x[1][8];
    const unsigned j = r * TPP + t;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = g[j + k * SUB];
    if (VARIANT != 2) cluster.sync(); else __syncthreads();
    tuple_ro<A, 3, 0, false, false, 1, false>(x, tw, 1u, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (VARIANT == 0) cluster.map_shared_rank(tile, k)[pad_index<T>(j)] = x[0][k];
        else tile[pad_index<T>((j + k * TPP) & (SUB - 1))] = x[0][k];  // local stores instead
    }
    if (VARIANT != 2) cluster.sync(); else __syncthreads();
    const SubPoly sub{3u, r};
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = tile[pad_index<T>(t + k * TPP)];
    fwd_from_regs<A, LOGSUB, 1>(x, tile, t, tw, c, sub);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[0][k] = A::fwd_fin(c, x[0][k]);
    store8_consecutive(g + r * SUB + 8 * t, x[0]);
}

template <int LOGSUB, int VARIANT>
void run(uint64_t* d, const uint64_t* tw, size_t polys, const char* name) {
    using A = Solinas64;
    auto kern = probe_fwd<A, LOGSUB, VARIANT>;
    size_t smem = FastShape<LOGSUB>::kPaddedElems * 8;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) kern<<<polys * 8, FastShape<LOGSUB>::kThreadsPerPoly, smem>>>(d, tw, A::Ctx{A::P});
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<polys * 8, FastShape<LOGSUB>::kThreadsPerPoly, smem>>>(d, tw, A::Ctx{A::P});
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int ncl = 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(polys * 8);
    cfg.blockDim = dim3(FastShape<LOGSUB>::kThreadsPerPoly);
    cfg.dynamicSmemBytes = smem;
    cudaOccupancyMaxActiveClusters(&ncl, kern, &cfg);
    printf("%-28s logsub=%d polys=%zu  %.3f ms  (%s, max active clusters %d)\n", name, LOGSUB, polys, ms / 10,
           cudaGetErrorString(cudaGetLastError()), ncl);
}

int main() {
    size_t bytes = size_t(1) << 30;
    uint64_t *d, *tw;
    cudaMalloc(&d, bytes);
    cudaMalloc(&tw, 65536 * 8);
    cudaMemset(d, 1, bytes);
    cudaMemset(tw, 3, 65536 * 8);
#define ALL(L)                                                   \
    run<L, 0>(d, tw, bytes / 8 >> (L + 3), "dsmem scatter + cluster.sync"); \
    run<L, 1>(d, tw, bytes / 8 >> (L + 3), "local stores + cluster.sync");  \
    run<L, 2>(d, tw, bytes / 8 >> (L + 3), "local stores + syncthreads");
    ALL(10) ALL(11) ALL(12) ALL(13)
    return 0;
}
