// Fused blind rotation of the NTT-PBS: one persistent CTA per LWE ciphertext runs all n_lwe
// CMUXes (tfhe ntt64_pbs.rs:213-286 classic, ntt64_bnf_pbs.rs:208-276 bnf) with the GLWE accumulator
// resident in shared memory.  Per CMUX the CTA
//   rotates / subtracts / gadget-decomposes the accumulator straight into registers,
//   runs the (k+1)*l forward NTTs (fwd_from_regs), multiply-accumulates against the GGSW of this
//   mask element (read once per CTA from L2, where the whole key stays resident) into register
//   accumulators, runs the (k+1) inverse NTTs (inv_to_regs) and adds the result into shared memory.
// HBM traffic per ciphertext: the look-up table in, the accumulator out; the unfused sequence
// moves ~(k+1)(l+2) polynomials per CMUX through HBM.
// The level loop runs inside the polynomial loop (the reference nests them the other way round,
// ntt64_pbs.rs:598-643); modular sums do not depend on the order.
#pragma once
#include "ntt_fast.cuh"
#include "pbs_math.cuh"

namespace nttb200 {

constexpr unsigned kPbsSkip = 0x80000000u;  // flag in a switched mask element: CMUX not executed

template <int LOGN, int GS>
struct PbsSmem {
    static constexpr size_t kTile = FastShape<LOGN>::kPaddedElems;       // NTT exchange tile
    static constexpr size_t kAcc = (size_t)GS << LOGN;                   // GLWE accumulator
    static size_t bytes(size_t n_lwe) { return (kTile + kAcc) * 8 + (n_lwe + 1) * 4; }
};

template <class A, int LOGN, int GS, bool BNF>
__global__ void __launch_bounds__(FastShape<LOGN>::kThreadsPerPoly)
    ntt_fast_blind_rotate_kernel(uint64_t* __restrict__ acc_out, const uint64_t* __restrict__ lut,
                                 size_t lut_count, const unsigned* __restrict__ switched,
                                 const uint64_t* __restrict__ bsk, unsigned n_lwe, unsigned base_log,
                                 unsigned level, unsigned width,
                                 const typename A::TW* __restrict__ tw_fwd,
                                 const typename A::TW* __restrict__ tw_inv, typename A::Ctx c,
                                 typename A::TW n_inv) {
    using S = FastShape<LOGN>;
    constexpr unsigned N = 1u << LOGN, TPP = S::kThreadsPerPoly;
    extern __shared__ __align__(16) uint64_t pbs_smem[];
    uint64_t* tile = pbs_smem;
    uint64_t* accS = pbs_smem + PbsSmem<LOGN, GS>::kTile;
    unsigned* sw = reinterpret_cast<unsigned*>(accS + PbsSmem<LOGN, GS>::kAcc);
    const unsigned t = threadIdx.x;
    const size_t b = blockIdx.x;
    const uint64_t p = c.p;
    const SubPoly sub{0u, 0u};

    for (unsigned i = t; i <= n_lwe; i += TPP) sw[i] = switched[b * (n_lwe + 1) + i];
    __syncthreads();
    const unsigned body = sw[n_lwe] & ~kPbsSkip;
    {
        const uint64_t* l = lut + (b % lut_count) * (size_t)(GS * N);
        for (unsigned idx = t; idx < GS * N; idx += TPP) {
            unsigned cc = idx >> LOGN, j = idx & (N - 1);
            // classic: lut / X^body first (ntt64_pbs.rs:247-255); bnf rotates at the end
            accS[idx] = BNF ? l[idx] : pbs::monomial_div_coeff(l + cc * N, j, body, LOGN, p);
        }
    }
    __syncthreads();

    for (unsigned i = 0; i < n_lwe; ++i) {
        const unsigned a = sw[i];
        if (a & kPbsSkip) continue;  // uniform over the CTA
        const uint64_t* ggsw = bsk + (size_t)i * level * GS * GS * N;
        uint64_t accr[GS][8];
#pragma unroll
        for (int cc = 0; cc < GS; ++cc)
#pragma unroll
            for (int k = 0; k < 8; ++k) accr[cc][k] = 0;

#pragma unroll 1
        for (int r = 0; r < GS; ++r) {
            // ct1 - ct0 = acc * X^a - acc at this thread's 8 positions, then the decomposer state
            const uint64_t* poly = accS + r * N;
            uint64_t state[8];
            unsigned negmask = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                unsigned j = t + k * TPP;
                uint64_t rot = pbs::monomial_mul_coeff(poly, j, a, LOGN, BNF ? 0 : p);
                if (BNF) {
                    state[k] = pbs::init_decomposer_state_native(rot - poly[j], base_log, level);
                } else {
                    bool neg;
                    state[k] = pbs::init_state_non_native(pbs::sub_mod(rot, poly[j], p), base_log, level, p, neg);
                    negmask |= (unsigned)neg << k;
                }
            }
#pragma unroll 1
            for (unsigned lv = 0; lv < level; ++lv) {
                uint64_t x[1][8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (BNF) {
                        uint64_t d = pbs::decompose_one_level(base_log, state[k]);
                        x[0][k] = (int64_t)d < 0 ? d + p : d;  // forward_from_decomp, ntt64.rs:229-236
                    } else {
                        x[0][k] = pbs::next_term_non_native(base_log, state[k], (negmask >> k) & 1u, p);
                    }
                }
                fwd_from_regs<A, LOGN, 1>(x, tile, t, tw_fwd, c, sub);
#pragma unroll
                for (int k = 0; k < 8; ++k) x[0][k] = A::fwd_fin(c, x[0][k]);
                const uint64_t* row = ggsw + ((size_t)(lv * GS + r) * GS << LOGN) + 8 * t;
#pragma unroll
                for (int cc = 0; cc < GS; ++cc) {
                    uint64_t g[8];
                    load8_consecutive(row + ((size_t)cc << LOGN), g);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        accr[cc][k] = A::acc_add(c, accr[cc][k], A::mul_full(c, x[0][k], g[k]));
                }
                __syncthreads();  // the tile is reused by the next transform
            }
        }
#pragma unroll
        for (int cc = 0; cc < GS; ++cc) {
            uint64_t x[1][8];
#pragma unroll
            for (int k = 0; k < 8; ++k) x[0][k] = A::acc_fin(c, accr[cc][k]);
            inv_to_regs<A, LOGN, 1>(x, tile, t, tw_inv, c, sub);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                uint64_t v = (k < 4) ? A::inv_fin(c, x[0][k]) : A::inv_fin_prod(c, x[0][k]);
                uint64_t* dst = accS + cc * N + t + k * TPP;
                if (BNF)  // normalize, modswitch p -> 2^width, wrapping add (ntt64_bnf_pbs.rs:669-673)
                    *dst += pbs::modswitch_prime_to_pow2(A::mul_const(c, v, n_inv), width, p);
                else  // add_backward, ntt64.rs:110-131
                    *dst = pbs::add_mod(*dst, v, p);
            }
            __syncthreads();
        }
    }

    uint64_t* o = acc_out + b * (size_t)(GS * N);
    for (unsigned idx = t; idx < GS * N; idx += TPP) {
        unsigned cc = idx >> LOGN, j = idx & (N - 1);
        o[idx] = BNF ? pbs::monomial_div_coeff(accS + cc * N, j, body, LOGN, 0) : accS[idx];
    }
}

template <class A>
bool fast_blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                       const uint64_t* bsk, size_t n_lwe, size_t glwe_size, unsigned base_log,
                       unsigned level, size_t batch, int bnf, unsigned width, int logn,
                       const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                       const typename A::Ctx& c, typename A::TW n_inv, cudaStream_t st);

}  // namespace nttb200
