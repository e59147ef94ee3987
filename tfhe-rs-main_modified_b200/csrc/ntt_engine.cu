// Plan construction (host) + kernel launch logic for the prime plans.
#include "ntt_engine.cuh"

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "ntt_fast.cuh"
#include "ntt_pbs_fused.cuh"
#include "ntt_kernels.cuh"
#include "plan_math.hpp"

namespace nttb200 {

namespace {

using pm::u128;

template <class A>
__global__ void mul_add_kernel(typename A::T* __restrict__ out, const typename A::T* __restrict__ rhs,
                               const typename A::T* __restrict__ acc, size_t total,
                               size_t rhs_period, size_t acc_period, typename A::Ctx c) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        typename A::T r = rhs[rhs_period == total ? i : i % rhs_period];
        typename A::T v = A::mul_full(c, out[i], r);
        if (acc) v = A::add_full(c, v, acc[acc_period == total ? i : i % acc_period]);
        out[i] = v;
    }
}

uint64_t inv_mod_2_64(uint64_t p) {  // Newton iteration, p odd
    uint64_t x = p;                  // correct to 3 bits
    for (int i = 0; i < 6; ++i) x *= 2 - p * x;
    return x;
}

// ---- per-family table / constant builders ------------------------------------------------
template <class A>
struct Build;

template <class T, bool H>
struct Build<Shoup<T, H>> {
    using A = Shoup<T, H>;
    static constexpr unsigned W = sizeof(T) * 8;
    static typename A::TW tw(uint64_t w, uint64_t p) {
        return {(T)w, (T)((((u128)w) << W) / p)};
    }
    static typename A::Ctx ctx(uint64_t p) {
        typename A::Ctx c{};
        c.p = (T)p;
        c.two_p = (T)(2 * p);
        if (W == 64) {
            c.pinv = (T)inv_mod_2_64(p);
            uint64_t r = (uint64_t)((((u128)1) << 64) % p);
            c.r2 = (T)pm::mulmod(r, r, p);
        } else {
            c.barrett64 = ~uint64_t(0) / p;
            int k = 64 - __builtin_clzll(p);  // bit length
            c.bar_shift = (uint32_t)(k - 1);
            c.bar_mu = (uint32_t)((((u128)1) << (k + 31)) / p);
        }
        return c;
    }
    static const char* name() {
        return W == 64 ? (H ? "shoup64-harvey" : "shoup64") : (H ? "shoup32-harvey" : "shoup32");
    }
};
template <>
struct Build<Wide32> {
    using A = Wide32;
    static A::TW tw(uint64_t w, uint64_t p) { return (uint32_t)((w << 32) % p); }  // Montgomery form
    static A::Ctx ctx(uint64_t p) {
        A::Ctx c{};
        c.p = (uint32_t)p;
        c.two_p = (uint32_t)(2 * p);
        c.pinv = (uint32_t)inv_mod_2_64(p);  // the low word of p^-1 mod 2^64 is p^-1 mod 2^32
        c.r2 = (uint32_t)((((u128)1) << 64) % p);
        c.barrett64 = ~uint64_t(0) / p;
        return c;
    }
    static const char* name() { return "wide32"; }
};
template <>
struct Build<Solinas64> {
    using A = Solinas64;
    static A::TW tw(uint64_t w, uint64_t p) { return (uint64_t)((((u128)w) << 64) % p); }
    static A::Ctx ctx(uint64_t p) { return A::Ctx{p}; }
    static const char* name() { return "solinas64"; }
};
template <>
struct Build<Mont64> {
    using A = Mont64;
    static A::TW tw(uint64_t w, uint64_t p) { return (uint64_t)((((u128)w) << 64) % p); }
    static A::Ctx ctx(uint64_t p) {
        A::Ctx c{};
        c.p = p;
        c.pinv = inv_mod_2_64(p);
        uint64_t r = (uint64_t)((((u128)1) << 64) % p);
        c.r2 = pm::mulmod(r, r, p);
        return c;
    }
    static const char* name() { return "mont64"; }
};

// NTT_B200_CLUSTER=1 routes u64 polynomials of 2^13 / 2^14 coefficients through the cluster-of-eight
// single-pass kernels instead of the (since measured faster) TMA-staged pass + single-CTA kernel
bool cluster_path_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("NTT_B200_CLUSTER");
        return e && e[0] == '1';
    }();
    return on;
}

// NTT_B200_NO_TMA=1 keeps the radix-16 strided passes on the register-staged kernel (A/B measurements)
bool tma_pass_enabled() {
    static const bool on = [] {
        const char* e = std::getenv("NTT_B200_NO_TMA");
        return !(e && e[0] == '1');
    }();
    return on;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count(int device) {
    static std::mutex mu;
    static int cache[64];
    std::lock_guard<std::mutex> lk(mu);
    if (device < 64 && cache[device]) return cache[device];
    int v = 148;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
    if (device < 64) cache[device] = v;
    return v;
}

template <class A>
struct PlanImpl final : PrimePlan {
    using T = typename A::T;
    using TW = typename A::TW;
    typename A::Ctx ctx{};
    TW n_inv{};
    // device tables in the reference's index order (twid[m + i]); owned through shared_ptr so
    // that clones share them
    std::shared_ptr<TW> d_fwd, d_inv;

    // Longest row one CTA keeps in shared memory: 2^14 u64 / 2^15 u32 = 128 KiB
    static constexpr int kMaxLogRow = sizeof(T) == 8 ? 14 : 15;

    static std::shared_ptr<TW> upload(const std::vector<TW>& h) {
        TW* d = nullptr;
        NTT_CUDA_CHECK(cudaMalloc(&d, h.size() * sizeof(TW)));
        std::shared_ptr<TW> sp(d, [](TW* q) { cudaFree(q); });
        NTT_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(TW), cudaMemcpyHostToDevice));
        return sp;
    }

    static std::shared_ptr<PlanImpl> create(size_t n, uint64_t p, const pm::Twiddles& t) {
        auto pl = std::make_shared<PlanImpl>();
        pl->n = n;
        pl->logn = __builtin_ctzll((uint64_t)n);
        pl->elem_bytes = sizeof(T);
        pl->p = p;
        pl->family = Build<A>::name();
        NTT_CUDA_CHECK(cudaGetDevice(&pl->device));
        pl->ctx = Build<A>::ctx(p);
        std::vector<TW> f(n), i(n);
        for (size_t k = 0; k < n; ++k) {
            f[k] = Build<A>::tw(t.fwd[k], p);
            i[k] = Build<A>::tw(t.inv[k], p);
        }
        pl->d_fwd = upload(f);
        pl->d_inv = upload(i);
        pl->n_inv = Build<A>::tw(pm::powmod((uint64_t)n, p - 2, p), p);
        auto bi = pm::barrett_info(p, sizeof(T) * 8);
        if (sizeof(T) == 8)  // prime64.rs:815-817 (use_ifma = false)
            pl->can_use_fast_reduction_code =
                p < 6148914691236517206ull || (bi.single_step && p < (uint64_t(1) << 63));
        else  // prime32.rs:748-749
            pl->can_use_fast_reduction_code =
                p < 1431655766ull || (bi.single_step && p <= (uint64_t(1) << 31));
        return pl;
    }

    std::shared_ptr<PrimePlan> clone() const override { return std::make_shared<PlanImpl>(*this); }
    bool blind_rotate(uint64_t* acc_out, const uint64_t* lut, size_t lut_count, const unsigned* switched,
                      const uint64_t* bsk, const uint64_t* bsk_tw, size_t n_lwe, size_t glwe_size,
                      unsigned base_log, unsigned level, size_t batch, int bnf, unsigned width, int latency,
                      cudaStream_t st) const override {
        if constexpr (std::is_same<A, Solinas64>::value) {
            DeviceGuard g(device);
            if (latency)
                return fast_blind_rotate_cluster<A>(acc_out, lut, lut_count, switched, bsk_tw, n_lwe, glwe_size,
                                                    base_log, level, batch, bnf, width, logn, d_fwd.get(),
                                                    d_inv.get(), ctx, n_inv, st);
            return fast_blind_rotate<A>(acc_out, lut, lut_count, switched, bsk, bsk_tw, n_lwe, glwe_size,
                                        base_log, level, batch, bnf, width, logn, d_fwd.get(), d_inv.get(), ctx,
                                        n_inv, st);
        } else {
            return false;
        }
    }
    bool key_to_twiddle_form(uint64_t* out, const uint64_t* in, size_t matrices, size_t glwe_size,
                             cudaStream_t st) const override {
        if constexpr (std::is_same<A, Solinas64>::value) {
            DeviceGuard g(device);
            return fast_key_to_twiddle_form<A>(out, in, matrices, glwe_size, logn, ctx, st);
        } else {
            return false;
        }
    }
    bool raw_shoup32h(RawShoup32H* out) const override {
        if constexpr (std::is_same<A, Shoup<uint32_t, true>>::value) {
            *out = RawShoup32H{d_fwd.get(), d_inv.get(), ctx, n_inv};
            return true;
        } else {
            return false;
        }
    }

    bool raw_shoup64h(RawShoup64H* out) const override {
        if constexpr (std::is_same<A, Shoup<uint64_t, true>>::value) {
            *out = RawShoup64H{d_fwd.get(), d_inv.get(), ctx, n_inv};
            return true;
        } else {
            return false;
        }
    }

    template <bool INV>
    void launch_rows(T* data, size_t num_rows, int log_row, int depth, int finalize,
                     cudaStream_t st) const {
        size_t row = size_t(1) << log_row;
        unsigned rows_per_cta = row >= 2048 ? 1u : (unsigned)(2048 / row);
        if (num_rows < rows_per_cta) rows_per_cta = (unsigned)num_rows;
        size_t elems = row * rows_per_cta;
        unsigned threads = (unsigned)std::min<size_t>(512, std::max<size_t>(64, elems / 8));
        size_t smem = elems * sizeof(T);
        auto kern = ntt_rows_kernel<A, INV>;
        if (smem > 48 * 1024)  // the attribute is per function and per device; setting it is cheap
            NTT_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(size_t(1) << kMaxLogRow) * (int)sizeof(T)));
        size_t ctas = (num_rows + rows_per_cta - 1) / rows_per_cta;
        kern<<<(unsigned)ctas, threads, smem, st>>>(data, num_rows, log_row, depth,
                                                    INV ? d_inv.get() : d_fwd.get(), ctx,
                                                    rows_per_cta, finalize);
        NTT_CUDA_CHECK(cudaGetLastError());
    }

    template <bool INV, int R>
    void launch_global(T* data, size_t batch, int stage, int finalize, cudaStream_t st) const {
        // n >> R tuples per polynomial, 256 per CTA (n >= 2^13 and R <= 4 here, so at least 512)
        const size_t gy = std::min<size_t>(batch, 32768);
        dim3 grid((unsigned)((n >> R) / 256), (unsigned)gy, (unsigned)((batch + gy - 1) / gy));
        ntt_global_pass_kernel<A, R, INV><<<grid, 256, 0, st>>>(
            data, batch, logn, stage, INV ? d_inv.get() : d_fwd.get(), ctx, finalize);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    // Radix-16 passes move their sixteen strided rows with the bulk-copy engine (double-buffered
    // shared-memory tiles, ntt_global_pass_tma_kernel): 0.38 ms against 0.50 ms for 1 GiB of u64
    // (profiles/r01_tma_pass_probe.txt); narrower passes already run at the HBM copy rate from registers.
    template <bool INV>
    bool launch_global_tma(T* data, size_t batch, int stage, int finalize, cudaStream_t st) const {
        constexpr int R = 4;
        if (!aligned16(data) || logn - stage - R < 8 || !tma_pass_enabled()) return false;
        auto kern = ntt_global_pass_tma_kernel<A, R, INV>;
        constexpr size_t smem = global_pass_tma_smem<A, R>();
        allow_dynamic_smem<ntt_global_pass_tma_kernel<A, R, INV>>(smem);
        const unsigned log_tiles_per_poly = (unsigned)(logn - R - 8);
        const size_t tiles = batch << log_tiles_per_poly;
        const size_t per_sm = std::min<size_t>(3, (size_t(200) << 10) / smem);
        const unsigned grid = (unsigned)std::min<size_t>(tiles, (size_t)sm_count(device) * per_sm);
        kern<<<grid, 256, smem, st>>>(data, tiles, log_tiles_per_poly, logn, stage,
                                      INV ? d_inv.get() : d_fwd.get(), ctx, finalize);
        NTT_CUDA_CHECK(cudaGetLastError());
        return true;
    }
    template <bool INV>
    void launch_global_r(int r, T* data, size_t batch, int stage, int finalize,
                         cudaStream_t st) const {
        if (r == 4 && launch_global_tma<INV>(data, batch, stage, finalize, st)) return;
        if (r == 1) launch_global<INV, 1>(data, batch, stage, finalize, st);
        if (r == 2) launch_global<INV, 2>(data, batch, stage, finalize, st);
        if (r == 3) launch_global<INV, 3>(data, batch, stage, finalize, st);
        if (r == 4) launch_global<INV, 4>(data, batch, stage, finalize, st);
    }

    // Longest polynomial the fast single-CTA kernels take; longer ones first run their top
    // `logn - kFastMaxLog` stages as strided global-memory passes (<= 4 stages per pass).
    static constexpr int kFastMaxLog = 12, kFastMinLog = 8;
    // Sizes served by the cluster-of-eight kernels (opt-in, see cluster_path_enabled).  Measured
    // (profiles/r01_large_n_cluster.txt): against a radix-2/4 pass + 4096-point kernel they gain 21 % at 2^13
    // and 4..15 % at 2^14 (u64), but the TMA-staged radix-16 pass + 512/1024-point kernels is faster still
    // (profiles/r01_large_n_depth.txt); with 512- and 1024-thread CTAs (2^15, 2^16) too few clusters are resident.
    static bool fast_sized(int log_sub) { return log_sub >= kFastMinLog && log_sub <= kFastMaxLog; }
    static bool cluster_sized(int logn) { return sizeof(T) == 8 && (logn == 13 || logn == 14); }
    // Number of top stages run as strided passes before the single-CTA kernel takes the 2^(logn - depth)
    // point sub-blocks.  NTT_B200_DEPTH=k overrides it (A/B measurements).
    int fast_depth() const {
        int lo = std::max(0, logn - kFastMaxLog), hi = std::max(lo, logn - kFastMinLog);
        static const int forced = [] {
            const char* e = std::getenv("NTT_B200_DEPTH");
            return e ? std::atoi(e) : -1;
        }();
        if (forced >= 0) return std::min(hi, std::max(lo, forced));
        // Measured (profiles/r01_large_n_depth.txt): as soon as a strided pass is needed, four stages in one
        // TMA-staged radix-16 pass plus the smaller (more efficient) single-CTA kernels beat fewer stages
        // plus the 4096-point kernel; the one exception is u32 at 2^14 (radix-8 pass + 2048-point kernel).
        if (lo == 0) return 0;
        if (sizeof(T) == 4 && logn == 14) return 3;
        return std::min(hi, std::max(lo, 4));
    }
    static std::vector<std::pair<int, int>> global_groups(int depth) {
        std::vector<std::pair<int, int>> g;  // (first stage, number of stages)
        for (int s = 0; s < depth;) {
            int r = std::min(4, depth - s);
            if (depth - s == 5) r = 3;  // 3+2 rather than 4+1
            g.push_back({s, r});
            s += r;
        }
        return g;
    }

    void fwd(void* data, size_t batch, cudaStream_t st) const override {
        if (!batch) return;
        NTT_NVTX("ntt_b200::Plan::fwd");
        DeviceGuard g(device);
        T* d = static_cast<T*>(data);
        const bool fast_ok = aligned16(d) && logn >= kFastMinLog;
        // opt-in: 2^13 / 2^14 u64 coefficients as one cluster of eight CTAs per polynomial, a single pass over HBM
        if (fast_ok && cluster_sized(logn) && cluster_path_enabled() &&
            fast_cluster_fwd<A>(d, batch, logn, d_fwd.get(), ctx, st))
            return;
        if (fast_ok && fast_sized(logn - fast_depth())) {
            // decided before any pass is launched: once the strided passes have run, only the fast kernel
            // for the remaining sub-blocks can finish the transform
            int depth = fast_depth();
            for (auto [s, r] : global_groups(depth)) launch_global_r<false>(r, d, batch, s, 0, st);
            if (!fast_fwd<A>(d, batch << depth, logn - depth, (unsigned)depth, d_fwd.get(), ctx, st))
                throw CudaError("no fast forward kernel for this size");
            return;
        }
        // small or unaligned: generic rows kernel (n <= 2^14 / 2^15 in one CTA)
        int depth = std::max(0, logn - kMaxLogRow);
        for (auto [s, r] : global_groups(depth)) launch_global_r<false>(r, d, batch, s, 0, st);
        launch_rows<false>(d, batch << depth, logn - depth, depth, 1, st);
    }
    void inv(void* data, size_t batch, cudaStream_t st) const override {
        if (!batch) return;
        NTT_NVTX("ntt_b200::Plan::inv");
        DeviceGuard g(device);
        T* d = static_cast<T*>(data);
        const bool fast_ok = aligned16(d) && logn >= kFastMinLog;
        if (fast_ok && cluster_sized(logn) && cluster_path_enabled() &&
            fast_cluster_inv<A>(d, batch, logn, d_inv.get(), ctx, st))
            return;
        const bool use_fast = fast_ok && fast_sized(logn - fast_depth());
        int depth = use_fast ? fast_depth() : std::max(0, logn - kMaxLogRow);
        if (use_fast) {
            bool ok = fast_inv<A>(d, batch << depth, logn - depth, (unsigned)depth, d_inv.get(), ctx, st);
            if (!ok) throw CudaError("no fast inverse kernel for this size");
        } else {
            launch_rows<true>(d, batch << depth, logn - depth, depth, depth == 0, st);
        }
        // mirror of fwd: the same stage groups in reverse order, the last one canonicalises
        auto groups = global_groups(depth);
        for (size_t k = groups.size(); k-- > 0;)
            launch_global_r<true>(groups[k].second, d, batch, groups[k].first, k == 0, st);
    }

    unsigned pointwise_blocks(size_t total) const {
        return (unsigned)std::min<size_t>((total + 255) / 256, (size_t)sm_count(device) * 16);
    }
    // 128-bit path: every pointer 16-byte aligned, every length / period a multiple of 16 / sizeof(T)
    static bool vec_ok(size_t a, size_t b = 0, size_t c = 0) {
        constexpr size_t V = 16 / sizeof(T);
        return a % V == 0 && b % V == 0 && c % V == 0;
    }
    void normalize(void* v, size_t total, cudaStream_t st) const override {
        if (!total) return;
        DeviceGuard g(device);
        if (aligned16(v) && vec_ok(total))
            normalize_vec_kernel<A><<<pointwise_blocks(total / (16 / sizeof(T))), 256, 0, st>>>(static_cast<T*>(v), total, ctx, n_inv);
        else
            normalize_kernel<A><<<pointwise_blocks(total), 256, 0, st>>>(static_cast<T*>(v), total, ctx, n_inv);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    void mul_assign_normalize(void* lhs, const void* rhs, size_t total, size_t rhs_period,
                              cudaStream_t st) const override {
        if (!total) return;
        DeviceGuard g(device);
        if (aligned16(lhs) && aligned16(rhs) && vec_ok(total, rhs_period))
            mul_assign_normalize_vec_kernel<A><<<pointwise_blocks(total / (16 / sizeof(T))), 256, 0, st>>>(
                static_cast<T*>(lhs), static_cast<const T*>(rhs), total, rhs_period, ctx, n_inv);
        else
            mul_assign_normalize_kernel<A><<<pointwise_blocks(total), 256, 0, st>>>(
                static_cast<T*>(lhs), static_cast<const T*>(rhs), total, rhs_period, ctx, n_inv);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    void mul_accumulate(void* acc, const void* lhs, const void* rhs, size_t total,
                        size_t lhs_period, size_t rhs_period, cudaStream_t st) const override {
        if (!total) return;
        DeviceGuard g(device);
        if (aligned16(acc) && aligned16(lhs) && aligned16(rhs) && vec_ok(total, lhs_period, rhs_period))
            mul_accumulate_vec_kernel<A><<<pointwise_blocks(total / (16 / sizeof(T))), 256, 0, st>>>(
                static_cast<T*>(acc), static_cast<const T*>(lhs), static_cast<const T*>(rhs), total,
                lhs_period, rhs_period, ctx);
        else
            mul_accumulate_kernel<A><<<pointwise_blocks(total), 256, 0, st>>>(
                static_cast<T*>(acc), static_cast<const T*>(lhs), static_cast<const T*>(rhs), total,
                lhs_period, rhs_period, ctx);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    void ext_product(void* out, const void* in, const void* ggsw, unsigned rows, unsigned cols,
                     size_t batch, cudaStream_t st) const override {
        if (!batch || !rows || !cols) return;
        NTT_NVTX("ntt_b200::Plan::ext_product");
        DeviceGuard g(device);
        if (aligned16(out) && aligned16(in) && aligned16(ggsw) &&
            fast_ext_product<A>(static_cast<T*>(out), static_cast<const T*>(in),
                                static_cast<const T*>(ggsw), rows, cols, batch, logn, d_fwd.get(),
                                d_inv.get(), ctx, st))
            return;
        // generic composition: transform a copy of the inputs, accumulate row by row, invert
        size_t in_total = batch * rows * n;
        T *tmp = nullptr, *row = nullptr, *acc = nullptr;
        NTT_CUDA_CHECK(cudaMallocAsync(&tmp, in_total * sizeof(T), st));
        NTT_CUDA_CHECK(cudaMallocAsync(&row, batch * n * sizeof(T), st));
        NTT_CUDA_CHECK(cudaMallocAsync(&acc, batch * n * sizeof(T), st));
        NTT_CUDA_CHECK(cudaMemcpyAsync(tmp, in, in_total * sizeof(T), cudaMemcpyDeviceToDevice, st));
        fwd(tmp, batch * rows, st);
        for (unsigned cc = 0; cc < cols; ++cc) {
            NTT_CUDA_CHECK(cudaMemsetAsync(acc, 0, batch * n * sizeof(T), st));
            for (unsigned r = 0; r < rows; ++r) {
                NTT_CUDA_CHECK(cudaMemcpy2DAsync(row, n * sizeof(T), tmp + (size_t)r * n, rows * n * sizeof(T),
                                                 n * sizeof(T), batch, cudaMemcpyDeviceToDevice, st));
                mul_accumulate(acc, row, static_cast<const T*>(ggsw) + ((size_t)r * cols + cc) * n, batch * n,
                               batch * n, n, st);
            }
            inv(acc, batch, st);
            NTT_CUDA_CHECK(cudaMemcpy2DAsync(static_cast<T*>(out) + (size_t)cc * n, cols * n * sizeof(T), acc,
                                             n * sizeof(T), n * sizeof(T), batch, cudaMemcpyDeviceToDevice, st));
        }
        NTT_CUDA_CHECK(cudaFreeAsync(tmp, st));
        NTT_CUDA_CHECK(cudaFreeAsync(row, st));
        NTT_CUDA_CHECK(cudaFreeAsync(acc, st));
    }
    void fwd_mac_inv(void* out, const void* lhs, const void* rhs, size_t rhs_polys,
                     const void* acc, size_t acc_polys, size_t batch,
                     cudaStream_t st) const override {
        if (!batch) return;
        NTT_NVTX("ntt_b200::Plan::fwd_mac_inv");
        DeviceGuard g(device);
        size_t total = batch * n;
        if (aligned16(out) && aligned16(lhs) && aligned16(rhs) && aligned16(acc) &&
            fast_fwd_mac_inv<A>(static_cast<T*>(out), static_cast<const T*>(lhs),
                                static_cast<const T*>(rhs), rhs_polys, static_cast<const T*>(acc),
                                acc_polys, batch, logn, d_fwd.get(), d_inv.get(), ctx, st))
            return;
        if (out != lhs)
            NTT_CUDA_CHECK(cudaMemcpyAsync(out, lhs, total * sizeof(T), cudaMemcpyDeviceToDevice, st));
        fwd(out, batch, st);
        mul_add_kernel<A><<<pointwise_blocks(total), 256, 0, st>>>(
            static_cast<T*>(out), static_cast<const T*>(rhs), static_cast<const T*>(acc), total,
            rhs_polys * n, acc ? acc_polys * n : total, ctx);
        NTT_CUDA_CHECK(cudaGetLastError());
        inv(out, batch, st);
    }
};

template <class A>
std::shared_ptr<PrimePlan> make_impl(size_t n, uint64_t p, const pm::Twiddles& t) {
    return PlanImpl<A>::create(n, p, t);
}

}  // namespace

std::shared_ptr<PrimePlan> make_plan64(size_t n, uint64_t p) {
    auto t = pm::build_twiddles(n, p, 16);
    if (!t) return nullptr;
    if (p == pm::kSolinas) return make_impl<Solinas64>(n, p, *t);
    if (p < (uint64_t(1) << 62)) return make_impl<Shoup<uint64_t, true>>(n, p, *t);
    if (p < (uint64_t(1) << 63)) return make_impl<Shoup<uint64_t, false>>(n, p, *t);
    return make_impl<Mont64>(n, p, *t);
}

std::shared_ptr<PrimePlan> make_plan32(size_t n, uint32_t p) {
    auto t = pm::build_twiddles(n, p, 32);
    if (!t) return nullptr;
    if (p < (uint32_t(1) << 30)) return make_impl<Shoup<uint32_t, true>>(n, p, *t);
    if (p < (uint32_t(1) << 31)) return make_impl<Shoup<uint32_t, false>>(n, p, *t);
    return make_impl<Wide32>(n, p, *t);
}

}  // namespace nttb200
