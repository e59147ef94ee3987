// Integer-pipe microbenchmark for B200 (sm_100a): measures warp-instruction issue rates of the
// instructions the NTT butterflies are made of.  Prints lane-ops/clk/SM for each mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe int_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void bench(uint32_t* out, uint32_t a, uint32_t b, long long* cycles) {
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        x[i] = threadIdx.x + i;
        y[i] = blockIdx.x + 3 * i;
    }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0) {  // IMAD (32-bit)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            } else if (MODE == 1) {  // IMAD.WIDE.U32
                uint64_t w;
                asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x[i]), "r"(a), "l"((uint64_t)y[i]));
                x[i] = (uint32_t)w;
                y[i] = (uint32_t)(w >> 32);
            } else if (MODE == 2) {  // IADD3 (3-input add)
                asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            } else if (MODE == 3) {  // 64-bit add via carry chain (2 IADD3)
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
            } else if (MODE == 4) {  // 1 IMAD + 1 IADD3 interleaved
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            } else if (MODE == 5) {  // 1 IMAD.WIDE + 2 IADD3
                uint64_t w;
                asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(x[i]), "r"(a), "l"((uint64_t)b));
                x[i] = (uint32_t)w;
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(y[i]), "+r"(x[i]) : "r"(a), "r"((uint32_t)(w >> 32)));
            } else if (MODE == 6) {  // LOP3 / SHF mix (alu)
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(y[i]) : "r"(x[i]));
            } else if (MODE == 7) {  // mul.hi.u32 (IMAD.HI)
                asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
            } else if (MODE == 8) {  // min.u32
                asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(a));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_chain_iter) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int threads = 1024, blocks = sms * 2;
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, blocks * sizeof(long long));
    bench<MODE><<<blocks, threads>>>(out, 12345u, 678u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<blocks, threads>>>(out, 12345u, 678u, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[4096];
    cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += h[i];
    avg /= blocks;
    // per SM: 2 resident blocks of 1024 threads run concurrently for ~avg cycles
    double lane_ops_per_sm = 2.0 * threads * (double)ITERS * CHAINS * ops_per_chain_iter;
    double per_clk = lane_ops_per_sm / avg;
    double total_ops = lane_ops_per_sm * sms;
    printf("%-34s %8.1f lane-ops/clk/SM  (%.2f warp-instr/clk/SM)  %7.2f Tops/s  %.3f ms  clk~%.0f MHz\n", name,
           per_clk, per_clk / 32, total_ops / (ms * 1e-3) / 1e12, ms, avg / (ms * 1e-3) / 1e6);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.WIDE.U32", 1);
    run<2>("IADD (add.u32 x2)", 2);
    run<3>("64-bit add (add.cc+addc)", 2);
    run<4>("IMAD + IADD interleaved", 2);
    run<5>("IMAD.WIDE + 2 IADD3 (carry)", 3);
    run<6>("LOP3 + SHF", 2);
    run<7>("IMAD.HI (mul.hi.u32)", 1);
    run<8>("IMNMX + IADD", 2);
    return 0;
}
