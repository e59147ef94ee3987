// Floor of a "small host call" on this box: what one kernel launch plus one wait costs, and whether a
// completion flag the kernel itself stores into mapped pinned memory (host spins on it) beats
// cudaStreamSynchronize.  Shapes the zero-copy path of capi_prime.cu (host_transform_small).
// Build: nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a profiles/microbench/launch_floor.cu -o profiles/microbench/launch_floor.bin
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e = (x);                                                       \
        if (e != cudaSuccess) {                                                    \
            printf("%s: %s\n", #x, cudaGetErrorString(e));                         \
            return 1;                                                              \
        }                                                                          \
    } while (0)

__global__ void empty_kernel() {}

// one CTA: every thread moves `per` 8-byte words of the mapped buffer in place (read, +1, write back),
// like the first-pass loads / last-pass stores of the one-kernel transform
template <bool FLAG>
__global__ void touch_kernel(uint64_t* buf, int per, volatile uint32_t* flag, uint32_t seq) {
    for (int k = 0; k < per; ++k) {
        size_t i = threadIdx.x + (size_t)k * blockDim.x;
        buf[i] = buf[i] + 1;
    }
    if (FLAG) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) *flag = seq;
    }
}

template <class F>
double per_call_us(F&& f, int reps = 5000, int warm = 500) {
    for (int i = 0; i < warm; ++i) f();
    auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < reps; ++i) f();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::micro>(t1 - t0).count() / reps;
}

int main() {
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint64_t* h;
    CK(cudaMallocHost(&h, 1 << 20));
    std::memset(h, 0, 1 << 20);
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(h + (1 << 16));
    uint64_t* d;
    CK(cudaMalloc(&d, 1 << 20));
    std::vector<uint64_t> user(4096, 1);

    double e = per_call_us([&] {
        empty_kernel<<<1, 256, 0, st>>>();
        cudaStreamSynchronize(st);
    });
    printf("empty kernel + cudaStreamSynchronize                      %6.2f us\n", e);
    for (int n : {1024, 2048, 4096}) {
        const int per = n / 256;
        const size_t bytes = (size_t)n * 8;
        double dev = per_call_us([&] {
            touch_kernel<false><<<1, 256, 0, st>>>(d, per, nullptr, 0);
            cudaStreamSynchronize(st);
        });
        double staged = per_call_us([&] {
            std::memcpy(h, user.data(), bytes);
            cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st);
            touch_kernel<false><<<1, 256, 0, st>>>(d, per, nullptr, 0);
            cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            std::memcpy(user.data(), h, bytes);
        });
        double mapped = per_call_us([&] {
            std::memcpy(h, user.data(), bytes);
            touch_kernel<false><<<1, 256, 0, st>>>(h, per, nullptr, 0);
            cudaStreamSynchronize(st);
            std::memcpy(user.data(), h, bytes);
        });
        uint32_t seq = 0;
        double flagged = per_call_us([&] {
            std::memcpy(h, user.data(), bytes);
            ++seq;
            touch_kernel<true><<<1, 256, 0, st>>>(h, per, flag, seq);
            while (*flag != seq) {
            }
            std::memcpy(user.data(), h, bytes);
        });
        CK(cudaStreamSynchronize(st));
        printf("n=%d u64: device-resident kernel + sync %6.2f us | staged H2D+kernel+D2H+sync %6.2f us | mapped kernel + sync %6.2f us | mapped kernel + own flag, host spin %6.2f us\n",
               n, dev, staged, mapped, flagged);
    }
    CK(cudaGetLastError());
    return 0;
}
