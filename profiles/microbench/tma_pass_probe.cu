// Probe for the TMA-staged strided pass (developer tool): the bulk-copy kernel against the
// register-staged ntt_global_pass_kernel on the same data -- identical bits required -- and timing.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
//        -I tfhe-rs-main_modified_b200/csrc -o tma_pass_probe.bin profiles/microbench/tma_pass_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ntt_kernels.cuh"
using namespace nttb200;
using A = Solinas64;

template <int R, bool INV>
void run(int logn, size_t polys, uint64_t* d_a, uint64_t* d_b, const uint64_t* tw, int ctas_per_sm, int finalize) {
    const size_t n = size_t(1) << logn;
    A::Ctx c{A::P};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms_old, ms_new;
    // old
    {
        size_t gy = polys < 32768 ? polys : 32768;
        dim3 grid((unsigned)((n >> R) / 256), (unsigned)gy, (unsigned)((polys + gy - 1) / gy));
        for (int i = 0; i < 2; ++i) ntt_global_pass_kernel<A, R, INV><<<grid, 256>>>(d_a, polys, logn, 0, tw, c, finalize);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) ntt_global_pass_kernel<A, R, INV><<<grid, 256>>>(d_a, polys, logn, 0, tw, c, finalize);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_old, e0, e1);
    }
    {
        auto kern = ntt_global_pass_tma_kernel<A, R, INV>;
        size_t smem = global_pass_tma_smem<A, R>();
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        unsigned ltp = logn - R - 8;
        size_t tiles = polys << ltp;
        unsigned grid = 148 * ctas_per_sm;
        if (grid > tiles) grid = (unsigned)tiles;
        for (int i = 0; i < 2; ++i) kern<<<grid, 256, smem>>>(d_b, tiles, ltp, logn, 0, tw, c, finalize);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) kern<<<grid, 256, smem>>>(d_b, tiles, ltp, logn, 0, tw, c, finalize);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms_new, e0, e1);
    }
    cudaError_t err = cudaDeviceSynchronize();
    // both buffers went through 7 identical passes: compare
    std::vector<uint64_t> ha(1 << 20), hb(1 << 20);
    size_t total = polys * n, mism = 0;
    for (size_t off = 0; off < total; off += total / 4) {
        size_t cnt = ha.size() < total - off ? ha.size() : total - off;
        cudaMemcpy(ha.data(), d_a + off, cnt * 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(hb.data(), d_b + off, cnt * 8, cudaMemcpyDeviceToHost);
        for (size_t i = 0; i < cnt; ++i) mism += ha[i] != hb[i];
    }
    printf("logn=%d R=%d inv=%d fin=%d ctas/sm=%d polys=%zu: register-staged %.3f ms, TMA-staged %.3f ms, mismatches %zu (%s)\n", logn,
           R, (int)INV, finalize, ctas_per_sm, polys, ms_old / 5, ms_new / 5, mism, cudaGetErrorString(err));
}

int main() {
    const size_t bytes = size_t(1) << 30;
    uint64_t *d_a, *d_b, *tw;
    cudaMalloc(&d_a, bytes);
    cudaMalloc(&d_b, bytes);
    cudaMalloc(&tw, 65536 * 8);
    std::vector<uint64_t> h(bytes / 8);
    uint64_t s = 88172645463325252ull;
    for (auto& v : h) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        v = s % A::P;
    }
    std::vector<uint64_t> htw(65536);
    for (auto& v : htw) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        v = s % A::P;
    }
    cudaMemcpy(tw, htw.data(), 65536 * 8, cudaMemcpyHostToDevice);
    for (int cps = 2; cps <= 6; ++cps) {
        cudaMemcpy(d_a, h.data(), bytes, cudaMemcpyHostToDevice);
        cudaMemcpy(d_b, h.data(), bytes, cudaMemcpyHostToDevice);
        if (cps <= 3) {
            run<4, false>(16, bytes / 8 >> 16, d_a, d_b, tw, cps, 0);
            run<4, true>(16, bytes / 8 >> 16, d_a, d_b, tw, cps, 1);
        }
        run<3, false>(15, bytes / 8 >> 15, d_a, d_b, tw, cps, 0);
        run<3, true>(15, bytes / 8 >> 15, d_a, d_b, tw, cps, 1);
        run<2, false>(14, bytes / 8 >> 14, d_a, d_b, tw, cps, 0);
        run<1, false>(13, bytes / 8 >> 13, d_a, d_b, tw, cps, 0);
    }
    return 0;
}
