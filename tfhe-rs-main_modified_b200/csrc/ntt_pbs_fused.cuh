// Fused blind rotation of the NTT-PBS: one persistent CTA per LWE ciphertext runs all n_lwe
// CMUXes (tfhe ntt64_pbs.rs:213-286 classic, ntt64_bnf_pbs.rs:208-276 bnf) with the GLWE accumulator
// resident in shared memory.  Per CMUX the CTA
//   rotates / subtracts / gadget-decomposes the accumulator straight into registers,
//   runs the (k+1)*l forward NTTs (fwd_from_regs, two polynomials per thread when k+1 is even),
//   multiply-accumulates against the GGSW of this mask element (read once per CTA from L2, where
//   the whole key stays resident; stored in Montgomery form, permuted so that a thread's operands
//   are contiguous) in registers -- or, with more than one transform group, in an L2 scratch --,
//   runs the (k+1) inverse NTTs (inv_to_regs) and adds the result into the GLWE accumulator.
// HBM traffic per ciphertext: the look-up table in, the accumulator out; the unfused sequence
// moves ~(k+1)(l+2) polynomials per CMUX through HBM.
// Modular sums do not depend on the order of the (level, polynomial) loop.  For level > 1 the
// decomposer state of a coefficient is recomputed per level (cheap next to a transform) instead
// of being kept in registers across the transforms.
#pragma once
#include <cooperative_groups.h>

#include "ntt_fast.cuh"
#include "pbs_math.cuh"

namespace nttb200 {

constexpr unsigned kPbsSkip = 0x80000000u;  // flag in a switched mask element: CMUX not executed

// Families whose modulus is a compile-time constant let the decomposition shifts fold.
template <class A>
struct FixedModulus {
    static constexpr uint64_t value = 0;
};
template <>
struct FixedModulus<Solinas64> {
    static constexpr uint64_t value = Solinas64::P;
};

template <int LOGN, int GS>
struct PbsShape {
    static constexpr int kPPT = (GS % 2 == 0) ? 2 : 1;                      // polynomials per transform
    static constexpr int kGroups = GS / kPPT;                                // transforms per level
    static constexpr size_t kTile = (size_t)kPPT * FastShape<LOGN>::kPaddedElems;  // NTT exchange tiles
    static constexpr size_t kAcc = (size_t)GS << LOGN;                       // GLWE accumulator
    static constexpr int kThreads = FastShape<LOGN>::kThreadsPerPoly;
    // shared memory (71 KiB at N = 2048, k = 1) allows three CTAs per SM: ask for <= 80 registers
    static constexpr int kMinBlocks = 768 / kThreads > 0 ? 768 / kThreads : 1;
    // REGACC keeps GS*8 more 64-bit accumulators per thread: 128 registers, two CTAs per SM
    static constexpr int kMinBlocksRegAcc = 512 / kThreads > 0 ? 512 / kThreads : 1;
    static size_t bytes(size_t n_lwe) { return (kTile + kAcc) * 8 + (n_lwe + 1) * 4; }
};

// Key layout read by the fused kernel ("twiddle form"): values in the form A::mul_const
// multiplies by, permuted so that a thread's operands are contiguous.  Per mask element and level
// the (k+1) x (k+1) matrix of polynomials is stored as
//   [row group rg][q < 4][thread t < N/8][pp < PPT][column cc < GS][e < 2]
// holding coefficient 8t + 2q + e of polynomial (row rg*PPT + pp, column cc).
template <int LOGN, int GS>
__host__ __device__ inline size_t pbs_key_index(unsigned rg, unsigned q, unsigned t, unsigned pp, unsigned cc,
                                                unsigned e) {
    constexpr unsigned PPT = PbsShape<LOGN, GS>::kPPT, TPP = PbsShape<LOGN, GS>::kThreads;
    return (((((size_t)rg * 4 + q) * TPP + t) * PPT + pp) * GS + cc) * 2 + e;
}

// SINGLE: k + 1 == PPT and level == 1, i.e. one forward and one inverse transform group per CMUX
// (the reference's parameter set): the multiply-accumulate runs in place in registers.  Otherwise
// the NTT-domain accumulators live in `scratch` ([batch][GS][4][N/8] 16-byte vectors, private per
// thread, L2 resident).
// REGACC: k + 1 == PPT and level > 1 (e.g. k = 1, l = 2, the shape of BASELINE config C3): still one
// transform group per level, so the NTT-domain accumulators of the GS output polynomials stay in
// registers across the levels instead of travelling through the L2 scratch.
template <class A, int LOGN, int GS, bool BNF, bool SINGLE, bool REGACC = false>
__global__ void __launch_bounds__(PbsShape<LOGN, GS>::kThreads,
                                  REGACC ? PbsShape<LOGN, GS>::kMinBlocksRegAcc : PbsShape<LOGN, GS>::kMinBlocks)
    ntt_fast_blind_rotate_kernel(uint64_t* __restrict__ acc_out, const uint64_t* __restrict__ lut,
                                 size_t lut_count, const unsigned* __restrict__ switched,
                                 const uint64_t* __restrict__ bsk_tw, ulonglong2* __restrict__ scratch,
                                 unsigned n_lwe, unsigned base_log, unsigned level, unsigned width,
                                 const typename A::TW* __restrict__ tw_fwd,
                                 const typename A::TW* __restrict__ tw_inv, typename A::Ctx c,
                                 typename A::TW n_inv) {
    using S = FastShape<LOGN>;
    using PS = PbsShape<LOGN, GS>;
    constexpr int PPT = PS::kPPT;
    constexpr unsigned N = 1u << LOGN, TPP = S::kThreadsPerPoly;
    static_assert(!(SINGLE || REGACC) || GS == PPT, "SINGLE / REGACC need one transform group per level");
    static_assert(!(SINGLE && REGACC), "SINGLE is the level == 1 case");
    extern __shared__ __align__(16) uint64_t pbs_smem[];
    uint64_t* tile = pbs_smem;
    uint64_t* accS = pbs_smem + PS::kTile;
    unsigned* sw = reinterpret_cast<unsigned*>(accS + PS::kAcc);
    const unsigned t = threadIdx.x;
    const size_t b = blockIdx.x;
    const uint64_t p = FixedModulus<A>::value ? FixedModulus<A>::value : c.p;
    const SubPoly sub{0u, 0u};
    ulonglong2* accN = (SINGLE || REGACC) ? nullptr : scratch + b * (size_t)(GS * 4 * TPP);

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
        const ulonglong2* ggsw =
            reinterpret_cast<const ulonglong2*>(bsk_tw + (size_t)i * level * GS * GS * N);
        uint64_t x[PPT][8];
        uint64_t o[REGACC ? GS : 1][8];  // REGACC: lazy sums of canonical products, one per output polynomial
        bool first = true;
#pragma unroll 1
        for (unsigned lv = 0; lv < level; ++lv) {
#pragma unroll 1
            for (int rg = 0; rg < PS::kGroups; ++rg) {
                // digit `lv` of decompose(acc * X^a - acc) at this thread's 8 positions
#pragma unroll
                for (int pp = 0; pp < PPT; ++pp) {
                    const uint64_t* poly = accS + (rg * PPT + pp) * N;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        unsigned j = t + k * TPP;
                        uint64_t rot = pbs::monomial_mul_coeff(poly, j, a, LOGN, BNF ? 0 : p);
                        uint64_t d;
                        if (BNF) {
                            uint64_t state = pbs::init_decomposer_state_native(rot - poly[j], base_log, level);
                            for (unsigned s = 0; s < lv; ++s) pbs::decompose_one_level(base_log, state);
                            d = pbs::decompose_one_level(base_log, state);
                            d = (int64_t)d < 0 ? d + p : d;  // forward_from_decomp, ntt64.rs:229-236
                        } else if (SINGLE) {  // level == 1
                            d = pbs::single_level_term_non_native(pbs::sub_mod(rot, poly[j], p), base_log, p);
                        } else {
                            bool neg;
                            uint64_t state =
                                pbs::init_state_non_native(pbs::sub_mod(rot, poly[j], p), base_log, level, p, neg);
                            for (unsigned s = 0; s < lv; ++s) pbs::decompose_one_level(base_log, state);
                            d = pbs::next_term_non_native(base_log, state, neg, p);
                        }
                        x[pp][k] = d;
                    }
                }
                fwd_from_regs<A, LOGN, PPT>(x, tile, t, tw_fwd, c, sub);
                // multiply-accumulate against rows rg*PPT .. rg*PPT+PPT-1 of this level's matrix;
                // this thread's PPT*GS operand vectors for coefficient pair q are contiguous
                const ulonglong2* gq =
                    ggsw + ((size_t)lv * GS * GS * N) / 2 + pbs_key_index<LOGN, GS>(rg, 0, t, 0, 0, 0) / 2;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    ulonglong2 g[PPT][GS];
#pragma unroll
                    for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
                        for (int cc = 0; cc < GS; ++cc)  // read once per CTA: keep it out of L1 (twiddles live there)
                            g[pp][cc] = __ldcg(gq + (size_t)q * TPP * PPT * GS + pp * GS + cc);
                    if (SINGLE) {
                        // out[cc] = sum_pp x[pp] * g[pp][cc], written over x (PPT == GS)
                        uint64_t o[GS][2];
#pragma unroll
                        for (int cc = 0; cc < GS; ++cc) {
                            o[cc][0] = A::mul_const(c, x[0][2 * q], g[0][cc].x);
                            o[cc][1] = A::mul_const(c, x[0][2 * q + 1], g[0][cc].y);
#pragma unroll
                            for (int pp = 1; pp < PPT; ++pp) {
                                o[cc][0] = A::acc_add(c, o[cc][0], A::mul_const(c, x[pp][2 * q], g[pp][cc].x));
                                o[cc][1] = A::acc_add(c, o[cc][1], A::mul_const(c, x[pp][2 * q + 1], g[pp][cc].y));
                            }
                        }
#pragma unroll
                        for (int cc = 0; cc < GS; ++cc) {
                            x[cc % PPT][2 * q] = A::acc_fin(c, o[cc][0]);
                            x[cc % PPT][2 * q + 1] = A::acc_fin(c, o[cc][1]);
                        }
                    } else if (REGACC) {
#pragma unroll
                        for (int cc = 0; cc < GS; ++cc) {
                            uint64_t s0 = first ? 0 : o[cc][2 * q], s1 = first ? 0 : o[cc][2 * q + 1];
#pragma unroll
                            for (int pp = 0; pp < PPT; ++pp) {
                                s0 = A::acc_add(c, s0, A::mul_const(c, x[pp][2 * q], g[pp][cc].x));
                                s1 = A::acc_add(c, s1, A::mul_const(c, x[pp][2 * q + 1], g[pp][cc].y));
                            }
                            o[cc][2 * q] = s0;
                            o[cc][2 * q + 1] = s1;
                        }
                    } else {
#pragma unroll
                        for (int cc = 0; cc < GS; ++cc) {
                            ulonglong2 acc = first ? make_ulonglong2(0, 0) : accN[(cc * 4 + q) * TPP + t];
#pragma unroll
                            for (int pp = 0; pp < PPT; ++pp) {
                                acc.x = A::acc_add(c, acc.x, A::mul_const(c, x[pp][2 * q], g[pp][cc].x));
                                acc.y = A::acc_add(c, acc.y, A::mul_const(c, x[pp][2 * q + 1], g[pp][cc].y));
                            }
                            accN[(cc * 4 + q) * TPP + t] = acc;
                        }
                    }
                }
                first = false;
                // SINGLE: the inverse starts by writing the 8 tile positions this thread just read
                if (!SINGLE) __syncthreads();
            }
        }
#pragma unroll 1
        for (int cg = 0; cg < PS::kGroups; ++cg) {
            if (REGACC) {
#pragma unroll
                for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
                    for (int k = 0; k < 8; ++k) x[pp][k] = A::acc_fin(c, o[pp][k]);  // one group: cg == 0
            } else if (!SINGLE) {
#pragma unroll
                for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        ulonglong2 v = accN[((cg * PPT + pp) * 4 + q) * TPP + t];
                        x[pp][2 * q] = A::acc_fin(c, v.x);
                        x[pp][2 * q + 1] = A::acc_fin(c, v.y);
                    }
            }
            inv_to_regs<A, LOGN, PPT>(x, tile, t, tw_inv, c, sub);
#pragma unroll
            for (int pp = 0; pp < PPT; ++pp)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint64_t v = (k < 4) ? A::inv_fin(c, x[pp][k]) : A::inv_fin_prod(c, x[pp][k]);
                    uint64_t* dst = accS + (cg * PPT + pp) * N + t + k * TPP;
                    if (BNF) {  // normalize, modswitch p -> 2^width, wrapping add (ntt64_bnf_pbs.rs:669-673)
                        uint64_t nv = A::mul_const(c, v, n_inv);
                        *dst += FixedModulus<A>::value == Solinas64::P ? pbs::modswitch_solinas_to_pow2(nv, width)
                                                                       : pbs::modswitch_prime_to_pow2(nv, width, p);
                    } else {  // add_backward, ntt64.rs:110-131
                        *dst = pbs::add_mod(*dst, v, p);
                    }
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

// Latency variant for small batches (k = 1, level = 1): a thread-block cluster of two CTAs shares
// one ciphertext, CTA r owning accumulator polynomial r.  Per CMUX each CTA decomposes and
// forward-transforms its own polynomial, multiplies it by row r of the GGSW, keeps the product
// for its own column and writes the product for the other column into the peer's shared memory
// (distributed shared memory, ping-pong buffers), one cluster barrier, then adds the peer's
// contribution, inverse-transforms and updates its polynomial.  Same arithmetic, same bits as the
// one-CTA kernel; half the transforms per SM per CMUX.
template <int LOGN>
struct PbsClusterShape {
    static constexpr size_t kTile = FastShape<LOGN>::kPaddedElems;
    static constexpr size_t kAcc = (size_t)1 << LOGN;
    static constexpr size_t kXbuf = (size_t)2 << LOGN;  // two exchange buffers of one polynomial (ping-pong)
    static constexpr int kThreads = FastShape<LOGN>::kThreadsPerPoly;
    // 69 KiB of shared memory at N = 2048 and 80 registers: 3 CTAs per SM.  (A single exchange buffer
    // with a second, split-phase barrier fits 4 CTAs at 64 registers: same peak throughput, +10 %
    // latency -- measured, not kept.)
    static constexpr int kMinBlocks = 768 / kThreads > 0 ? 768 / kThreads : 1;
    static size_t bytes(size_t n_lwe) { return (kTile + kAcc + kXbuf + 2) * 8 + (n_lwe + 1) * 4; }  // + two mbarriers
};

// ---- distributed-shared-memory exchange without a cluster barrier per round -----------------------
// barrier.cluster.arrive.release / wait.acquire (cg::cluster_group::sync) costs a GPU-scope MEMBAR and an
// L1 invalidation (CCTL.IVALL) each time, which also throws the twiddle tables out of L1 once per CMUX.
// The per-round exchange instead uses asynchronous remote stores that complete a transaction count on
// an mbarrier in the RECEIVER's shared memory (st.async ... mbarrier::complete_tx::bytes, SASS STAS +
// SYNCS): the receiver arms the barrier with the byte count it expects and waits on it locally.
NTT_DEVINL uint32_t smem_addr_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
NTT_DEVINL uint32_t map_to_cta(uint32_t local_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
    return r;
}
NTT_DEVINL void mbar_init(uint32_t mbar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(arrivals) : "memory");
}
NTT_DEVINL void mbar_fence_init_cluster() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
NTT_DEVINL void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
NTT_DEVINL void mbar_wait_parity(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{ .reg .pred p;\n"
        "MBAR_WAIT_LOOP: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_WAIT_DONE;\n"
        "bra MBAR_WAIT_LOOP;\n"
        "MBAR_WAIT_DONE: }" ::"r"(mbar), "r"(parity)
        : "memory");
}
// 16 bytes into the peer CTA's shared memory, counted on the peer's mbarrier
NTT_DEVINL void st_async_remote_v2(uint32_t remote_addr, uint64_t a, uint64_t b, uint32_t remote_mbar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.v2.b64 [%0], {%1, %2}, [%3];" ::"r"(remote_addr),
                 "l"(a), "l"(b), "r"(remote_mbar)
                 : "memory");
}

template <class A, int LOGN, bool BNF>
__global__ void __cluster_dims__(2, 1, 1)
    __launch_bounds__(PbsClusterShape<LOGN>::kThreads, PbsClusterShape<LOGN>::kMinBlocks)
        ntt_fast_blind_rotate_cluster2_kernel(uint64_t* __restrict__ acc_out, const uint64_t* __restrict__ lut,
                                          size_t lut_count, const unsigned* __restrict__ switched,
                                          const uint64_t* __restrict__ bsk_tw, unsigned n_lwe,
                                          unsigned base_log, unsigned width,
                                          const typename A::TW* __restrict__ tw_fwd,
                                          const typename A::TW* __restrict__ tw_inv, typename A::Ctx c,
                                          typename A::TW n_inv) {
    namespace cg = cooperative_groups;
    using S = FastShape<LOGN>;
    using CS = PbsClusterShape<LOGN>;
    constexpr unsigned N = 1u << LOGN, TPP = S::kThreadsPerPoly, GS = 2;
    extern __shared__ __align__(16) uint64_t pbs_smem[];
    uint64_t* tile = pbs_smem;
    uint64_t* accS = pbs_smem + CS::kTile;                                  // polynomial r of the accumulator
    ulonglong2* xbuf = reinterpret_cast<ulonglong2*>(accS + CS::kAcc);      // [2][4][TPP] vectors
    uint64_t* mbar = accS + CS::kAcc + CS::kXbuf;                           // one mbarrier per exchange buffer
    unsigned* sw = reinterpret_cast<unsigned*>(mbar + 2);
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned r = cluster.block_rank();
    const unsigned t = threadIdx.x;
    // (addresses of the second barrier = first + 8: a CTA's window of the cluster address space keeps offsets)
    const uint32_t my_mbar0 = smem_addr_u32(mbar), peer_mbar0 = map_to_cta(my_mbar0, r ^ 1u);
    const uint32_t peer_xbuf = map_to_cta(smem_addr_u32(xbuf), r ^ 1u);
    if (t == 0) {
        mbar_init(my_mbar0, 1);
        mbar_init(my_mbar0 + 8u, 1);
        mbar_fence_init_cluster();
    }
    const size_t b = blockIdx.x >> 1;
    const uint64_t p = FixedModulus<A>::value ? FixedModulus<A>::value : c.p;
    const SubPoly sub{0u, 0u};

    for (unsigned i = t; i <= n_lwe; i += TPP) sw[i] = switched[b * (n_lwe + 1) + i];
    __syncthreads();
    const unsigned body = sw[n_lwe] & ~kPbsSkip;
    {
        const uint64_t* l = lut + (b % lut_count) * (size_t)(GS * N) + (size_t)r * N;
        for (unsigned j = t; j < N; j += TPP) accS[j] = BNF ? l[j] : pbs::monomial_div_coeff(l, j, body, LOGN, p);
    }
    cluster.sync();  // both CTAs are resident and their mbarriers initialised before any remote store

    // Ping-pong exchange buffers indexed by the number of executed CMUXes, one mbarrier each.  The peer's
    // stores of round e + 2 into the buffer of round e are issued after it has received this CTA's round
    // e + 1 products, which this CTA sends only after it has read round e: two buffers suffice without any
    // cluster-wide barrier inside the loop.  A buffer's mbarrier is armed (one arrival + the byte count of
    // one polynomial of products) by thread 0 at the start of the round that uses it; remote bytes that
    // land before the arming only drive the transaction count negative while that arrival is pending.
    unsigned executed = 0;
    for (unsigned i = 0; i < n_lwe; ++i) {
        const unsigned a = sw[i];
        if (a & kPbsSkip) continue;  // uniform over the cluster
        const ulonglong2* gq = reinterpret_cast<const ulonglong2*>(bsk_tw + (size_t)i * GS * GS * N) +
                               pbs_key_index<LOGN, GS>(0, 0, t, r, 0, 0) / 2;
        const unsigned slot = executed & 1u, parity = (executed >> 1) & 1u;
        if (t == 0) mbar_arrive_expect_tx(my_mbar0 + 8u * slot, 8u << LOGN);
        uint64_t x[1][8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned j = t + k * TPP;
            uint64_t rot = pbs::monomial_mul_coeff(accS, j, a, LOGN, BNF ? 0 : p);
            uint64_t d;
            if (BNF) {
                uint64_t state = pbs::init_decomposer_state_native(rot - accS[j], base_log, 1);
                d = pbs::decompose_one_level(base_log, state);
                d = (int64_t)d < 0 ? d + p : d;
            } else {
                d = pbs::single_level_term_non_native(pbs::sub_mod(rot, accS[j], p), base_log, p);
            }
            x[0][k] = d;
        }
        fwd_from_regs<A, LOGN, 1>(x, tile, t, tw_fwd, c, sub);
        const uint32_t send = peer_xbuf + (slot * 4 * TPP + t) * 16u;
        const ulonglong2* recv = xbuf + (size_t)slot * 4 * TPP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ulonglong2 g_own = __ldcg(gq + (size_t)q * TPP * 4 + r);         // column r
            ulonglong2 g_peer = __ldcg(gq + (size_t)q * TPP * 4 + (r ^ 1u));  // column 1 - r
            st_async_remote_v2(send + q * TPP * 16u, A::mul_const(c, x[0][2 * q], g_peer.x),
                               A::mul_const(c, x[0][2 * q + 1], g_peer.y), peer_mbar0 + 8u * slot);
            x[0][2 * q] = A::mul_const(c, x[0][2 * q], g_own.x);
            x[0][2 * q + 1] = A::mul_const(c, x[0][2 * q + 1], g_own.y);
        }
        mbar_wait_parity(my_mbar0 + 8u * slot, parity);  // the peer's products for my column have landed
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ulonglong2 v = recv[q * TPP + t];
            x[0][2 * q] = A::acc_fin(c, A::acc_add(c, x[0][2 * q], v.x));
            x[0][2 * q + 1] = A::acc_fin(c, A::acc_add(c, x[0][2 * q + 1], v.y));
        }
        ++executed;
        inv_to_regs<A, LOGN, 1>(x, tile, t, tw_inv, c, sub);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            uint64_t v = (k < 4) ? A::inv_fin(c, x[0][k]) : A::inv_fin_prod(c, x[0][k]);
            uint64_t* dst = accS + t + k * TPP;
            if (BNF) {
                uint64_t nv = A::mul_const(c, v, n_inv);
                *dst += FixedModulus<A>::value == Solinas64::P ? pbs::modswitch_solinas_to_pow2(nv, width)
                                                               : pbs::modswitch_prime_to_pow2(nv, width, p);
            } else {
                *dst = pbs::add_mod(*dst, v, p);
            }
        }
        __syncthreads();
    }
    uint64_t* o = acc_out + b * (size_t)(GS * N) + (size_t)r * N;
    for (unsigned j = t; j < N; j += TPP) o[j] = BNF ? pbs::monomial_div_coeff(accS, j, body, LOGN, 0) : accS[j];
    cluster.sync();  // no CTA leaves while its peer could still address its shared memory
}

// Fused blind rotation; bsk_tw is the key in twiddle form (fast_key_to_twiddle_form).  false: no
// kernel for this shape.
template <class A>
bool fast_blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                       const uint64_t* bsk, const uint64_t* bsk_tw, size_t n_lwe, size_t glwe_size,
                       unsigned base_log, unsigned level, size_t batch, int bnf, unsigned width, int logn,
                       const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                       const typename A::Ctx& c, typename A::TW n_inv, cudaStream_t st);
// The two-CTA cluster (latency) variant: k = 1, level = 1 only.
template <class A>
bool fast_blind_rotate_cluster(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                               const uint64_t* bsk_tw, size_t n_lwe, size_t glwe_size, unsigned base_log,
                               unsigned level, size_t batch, int bnf, unsigned width, int logn,
                               const typename A::TW* tw_fwd, const typename A::TW* tw_inv,
                               const typename A::Ctx& c, typename A::TW n_inv, cudaStream_t st);
// out = the key `in` ([matrices][glwe_size][glwe_size][N], NTT domain) in the layout and form the
// fused kernel reads (pbs_key_index); false: this family / shape has no fused kernel
template <class A>
bool fast_key_to_twiddle_form(uint64_t* out, const uint64_t* in, size_t matrices, size_t glwe_size, int logn,
                              const typename A::Ctx& c, cudaStream_t st);

}  // namespace nttb200
