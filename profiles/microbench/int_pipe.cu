// Integer-pipe microbenchmark for B200 (sm_100a): measures warp-instruction issue rates of the
// instructions the NTT butterflies are made of.  Prints lane-ops/clk/SM for each mix.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe int_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

#define ITERS 4096
#define CHAINS 8

template <int MODE>
__global__ void bench(uint32_t* out, uint32_t a, uint32_t b, long long* cycles) {
    // cycles[4*block + {0: min start clock, 1: max end clock, 2: smid, 3: max end globaltimer - min start}]
    uint32_t x[CHAINS], y[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        x[i] = threadIdx.x + i;
        y[i] = blockIdx.x + 3 * i;
    }
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
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
            } else if (MODE == 9) {  // pure IMAD.WIDE.U32: 64-bit accumulator chain
                uint64_t w = ((uint64_t)y[i] << 32) | x[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w) : "r"(x[i]), "r"(a));
                x[i] = (uint32_t)w;
                y[i] = (uint32_t)(w >> 32);
            } else if (MODE == 10) {  // mul.wide.u32, both halves consumed (IMAD.WIDE.U32 + LOP3)
                uint64_t w;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(a));
                x[i] = (uint32_t)(w >> 32) ^ (uint32_t)w;
            } else if (MODE == 11) {  // mul.hi.u32 + LOP3 (same shape with IMAD.HI)
                uint32_t h;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(h) : "r"(x[i]), "r"(a));
                x[i] = h ^ y[i];
            } else if (MODE == 12) {  // Shoup-32 product via IMAD.HI: q=hi(b*ws); t=b*w-q*p
                uint32_t q;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(x[i]), "r"(a));
                x[i] = x[i] * b - q * y[i];
            } else if (MODE == 13) {  // Shoup-32 product via mul.wide: (lo,hi)=b*ws; t=b*w-hi*p+ (lo&0)
                uint64_t w;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(x[i]), "r"(a));
                uint32_t q = (uint32_t)(w >> 32);
                x[i] = x[i] * b - q * y[i] + ((uint32_t)w & 1u);
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
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if ((threadIdx.x & 31) == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicMin((unsigned long long*)&cycles[4 * blockIdx.x + 0], (unsigned long long)t0);
        atomicMax((unsigned long long*)&cycles[4 * blockIdx.x + 1], (unsigned long long)t1);
        cycles[4 * blockIdx.x + 2] = smid;
        atomicMin((unsigned long long*)&cycles[4 * blockIdx.x + 3], g0);
        atomicMax((unsigned long long*)&cycles[4 * blockIdx.x + 3 + 4 * gridDim.x], g1);
    }
}

template <int MODE>
void run(const char* name, int instr_per_chain_iter) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int threads = 1024, blocks = sms * 2;
    uint32_t* out;
    long long* cyc;
    size_t cyc_n = (size_t)blocks * 8;
    cudaMalloc(&out, (size_t)blocks * threads * 4);
    cudaMalloc(&cyc, cyc_n * sizeof(long long));
    std::vector<long long> init(cyc_n, 0);
    for (int b = 0; b < blocks; ++b) { init[4 * b + 0] = -1; init[4 * b + 3] = -1; }  // min slots = UINT64_MAX
    // spin the clocks up first: ~50 ms of the same kernel
    for (int w = 0; w < 40; ++w) {
        cudaMemcpy(cyc, init.data(), cyc_n * sizeof(long long), cudaMemcpyHostToDevice);
        bench<MODE><<<blocks, threads>>>(out, 12345u, 678u, cyc);
    }
    cudaDeviceSynchronize();
    std::vector<long long> h(cyc_n);
    cudaMemcpy(h.data(), cyc, cyc_n * sizeof(long long), cudaMemcpyDeviceToHost);
    // per SM: span between the earliest start and the latest end of the blocks that ran there
    std::vector<unsigned long long> s0(1024, ~0ull), s1(1024, 0), nb(1024, 0);
    unsigned long long g0 = ~0ull, g1 = 0;
    for (int b = 0; b < blocks; ++b) {
        int sm = (int)h[4 * b + 2];
        s0[sm] = std::min<unsigned long long>(s0[sm], (unsigned long long)h[4 * b + 0]);
        s1[sm] = std::max<unsigned long long>(s1[sm], (unsigned long long)h[4 * b + 1]);
        nb[sm]++;
        g0 = std::min<unsigned long long>(g0, (unsigned long long)h[4 * b + 3]);
        g1 = std::max<unsigned long long>(g1, (unsigned long long)h[4 * b + 3 + 4 * blocks]);
    }
    double rate_sum = 0, span_sum = 0;
    int used = 0;
    for (int sm = 0; sm < 1024; ++sm)
        if (nb[sm]) {
            double span = (double)(s1[sm] - s0[sm]);
            double work = (double)nb[sm] * threads * (double)ITERS * CHAINS * instr_per_chain_iter;
            rate_sum += work / span;
            span_sum += span;
            ++used;
        }
    double per_clk = rate_sum / used;  // lane-instr / clk / SM
    double mhz = (span_sum / used) / (double)(g1 - g0) * 1e3;
    printf("%-34s %7.1f lane-instr/clk/SM = %.2f warp-instr/clk/SM (%.2f per SMSP)   SM clock %.0f MHz  -> %.2f T lane-instr/s\n",
           name, per_clk, per_clk / 32, per_clk / 128, mhz, per_clk * sms * mhz * 1e6 / 1e12);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    // second argument = SASS instructions per chain per iteration (checked with cuobjdump)
    run<0>("IMAD (mad.lo.u32)", 1);
    run<1>("IMAD.WIDE+IADD3+IMAD.X (3/chain)", 3);
    run<2>("IADD3 / IMAD.IADD mix (add.u32 x2)", 2);
    run<3>("IADD3 + IADD3.X (64-bit add)", 2);
    run<4>("IMAD + IADD interleaved", 2);
    run<5>("IMAD.WIDE + 3 IADD3", 4);
    run<6>("LOP3 + SHF", 2);
    run<7>("IMAD.HI (mul.hi.u32)", 1);
    run<8>("VIMNMX + IMAD.IADD", 2);
    run<9>("IMAD.WIDE+IADD3+IMAD.X (b)", 3);
    run<10>("IMAD.WIDE.U32 + LOP3", 2);
    run<11>("IMAD.HI.U32 + LOP3", 2);
    run<12>("Shoup32 via IMAD.HI (3 instr)", 3);
    run<13>("Shoup32 via IMAD.WIDE (5 instr)", 5);
    return 0;
}
