// product::Plan -- negacyclic NTT modulo a product of distinct primes
// (reference: tfhe-ntt/src/product.rs:139-967).
//
// Layout of the NTT domain.  For one polynomial it is the reference's (product.rs:261-283): the
// residue arrays of the 32-bit primes (n u32 each, i.e. n/2 words) followed by those of the 64-bit
// primes.  For a batch the arrays are prime-major -- prime j owns `batch` contiguous residue
// polynomials -- so every per-prime transform is one batched kernel; batch = 1 is exactly the
// reference layout.
#include <algorithm>
#include <vector>

#include "capi_common.cuh"
#include "ntt_arith.cuh"
#include "plan_math.hpp"

using namespace nttb200;
using pm::u128;

namespace {

constexpr int kMaxPrimes = 16;

struct ProductConsts {
    int n32, n64;
    uint64_t modulus;
    uint64_t p[kMaxPrimes];      // sorted: 32-bit primes first
    uint64_t b64[kMaxPrimes];    // floor(2^64 / p) for the 32-bit primes
    uint64_t inv[kMaxPrimes * (kMaxPrimes - 1) / 2];  // p_i^-1 mod p_j, i < j (product.rs:203-225)
};

struct Regions {
    void* r[kMaxPrimes];
};

NTT_DEVINL uint64_t mulmod_any(uint64_t a, uint64_t b, uint64_t p) {
    return (uint64_t)(((unsigned __int128)a * b) % p);
}

// standard -> residues (product.rs:273-357).  `reduce` is false for the single-prime plans, which
// copy / truncate without reducing (:282-293).
__global__ void product_split_kernel(const uint64_t* __restrict__ standard, Regions reg, size_t total,
                                     int reduce, ProductConsts k) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = standard[i];
        for (int j = 0; j < k.n32; ++j) {
            uint32_t r;
            if (reduce) {
                uint64_t q = __umul64hi(v, k.b64[j]);
                uint64_t t = v - q * k.p[j];
                r = (uint32_t)(t >= k.p[j] ? t - k.p[j] : t);
            } else {
                r = (uint32_t)v;
            }
            static_cast<uint32_t*>(reg.r[j])[i] = r;
        }
        for (int j = 0; j < k.n64; ++j)
            static_cast<uint64_t*>(reg.r[k.n32 + j])[i] = reduce ? v % k.p[k.n32 + j] : v;
    }
}

// residues -> standard: mixed-radix recombination (Knuth 4.3.2; product.rs:792-879), then
// Replace or Accumulate (add modulo the product, product.rs:876-877).
__global__ void product_merge_kernel(uint64_t* __restrict__ standard, Regions reg, size_t total,
                                     int accumulate, ProductConsts k) {
    const int np = k.n32 + k.n64;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        uint64_t v[kMaxPrimes];
        int offset = 0;
        for (int j = 0; j < np; ++j) {
            uint64_t pj = k.p[j];
            uint64_t x = j < k.n32 ? (uint64_t) static_cast<const uint32_t*>(reg.r[j])[idx]
                                   : static_cast<const uint64_t*>(reg.r[j])[idx];
            for (int i = 0; i < j; ++i) {
                uint64_t diff = x >= v[i] ? x - v[i] : x - v[i] + pj;
                x = mulmod_any(diff, k.inv[offset + i], pj);
            }
            offset += j;
            v[j] = x;
        }
        uint64_t acc = 0;
        for (int j = np - 1; j >= 0; --j) acc = acc * k.p[j] + v[j];
        if (accumulate) {
            uint64_t a = standard[idx], sum = a + acc;
            acc = (sum >= k.modulus || sum < a) ? sum - k.modulus : sum;
        }
        standard[idx] = acc;
    }
}

uint64_t modular_inverse(uint64_t x, uint64_t p) {  // p prime: x^(p-2)
    return pm::powmod(x % p, p - 2, p);
}

}  // namespace

struct ntt_b200_product_plan {
    size_t n = 0;
    uint64_t modulus = 0;
    int device = 0;
    std::vector<std::shared_ptr<PrimePlan>> plans;  // 32-bit primes first
    ProductConsts k{};

    size_t domain_len() const { return (n / 2) * (size_t)k.n32 + n * (size_t)k.n64; }
    int np() const { return k.n32 + k.n64; }
    unsigned blocks(size_t total) const { return (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16); }

    // prime-major regions inside a device NTT-domain buffer holding `batch` polynomials
    Regions regions(void* ntt, size_t batch) const {
        Regions g{};
        char* base = static_cast<char*>(ntt);
        size_t off = 0;
        for (int j = 0; j < np(); ++j) {
            g.r[j] = base + off;
            off += batch * n * (j < k.n32 ? 4 : 8);
        }
        return g;
    }
    void fwd_dev(void* ntt, const void* standard, size_t batch, cudaStream_t st) const {
        if (!batch || np() == 0) return;
        Regions g = regions(ntt, batch);
        size_t total = batch * n;
        product_split_kernel<<<blocks(total), 256, 0, st>>>(static_cast<const uint64_t*>(standard), g, total,
                                                            np() > 1, k);
        NTT_CUDA_CHECK(cudaGetLastError());
        for (int j = 0; j < np(); ++j) plans[j]->fwd(g.r[j], batch, st);
    }
    void inv_dev(void* standard, void* ntt, size_t batch, bool accumulate, cudaStream_t st) const {
        if (!batch) return;
        size_t total = batch * n;
        if (np() == 0) {  // product.rs:378-384
            if (!accumulate) NTT_CUDA_CHECK(cudaMemsetAsync(standard, 0, total * 8, st));
            return;
        }
        Regions g = regions(ntt, batch);
        for (int j = 0; j < np(); ++j) plans[j]->inv(g.r[j], batch, st);
        product_merge_kernel<<<blocks(total), 256, 0, st>>>(static_cast<uint64_t*>(standard), g, total,
                                                            accumulate ? 1 : 0, k);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    // op: 0 normalize(a), 1 mul_assign_normalize(a, b), 2 mul_accumulate(a, b, c)
    void pointwise_dev(int op, void* a, const void* b, const void* c, size_t batch, cudaStream_t st) const {
        Regions ga = regions(a, batch), gb = regions(const_cast<void*>(b ? b : a), batch),
                gc = regions(const_cast<void*>(c ? c : a), batch);
        size_t total = batch * n;
        for (int j = 0; j < np(); ++j) {
            if (op == 0) plans[j]->normalize(ga.r[j], total, st);
            if (op == 1) plans[j]->mul_assign_normalize(ga.r[j], gb.r[j], total, total, st);
            if (op == 2) plans[j]->mul_accumulate(ga.r[j], gb.r[j], gc.r[j], total, total, total, st);
        }
    }
};

namespace {

struct HostStage2 {
    cudaStream_t st = nullptr;
    std::vector<void*> bufs;
    bool owned = false;
    explicit HostStage2(int device) {
        keep_pool_cached(device);
        st = cached_stream(device);
        if (!st) {
            NTT_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            owned = true;
        }
    }
    void* alloc(size_t bytes) {
        void* d = nullptr;
        NTT_CUDA_CHECK(cudaMallocAsync(&d, std::max<size_t>(bytes, 16), st));
        bufs.push_back(d);
        return d;
    }
    void* upload(const void* h, size_t bytes) {
        void* d = alloc(bytes);
        if (bytes) NTT_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, st));
        return d;
    }
    void download(void* h, const void* d, size_t bytes) {
        if (bytes) NTT_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, st));
    }
    void finish() {
        for (void* d : bufs) cudaFreeAsync(d, st);
        bufs.clear();
        NTT_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    ~HostStage2() {
        for (void* d : bufs) cudaFreeAsync(d, st);
        if (st) {
            cudaStreamSynchronize(st);
            if (owned) cudaStreamDestroy(st);
        }
    }
};

}  // namespace

extern "C" {

// product::Plan::try_new(polynomial_size, modulus, factors) -> Option<Plan>   product.rs:153-246
int ntt_b200_product_try_new(size_t n, uint64_t modulus, const uint64_t* factors, size_t nfactors,
                             ntt_b200_product_plan** out) {
    if (!out || (!factors && nfactors)) return NTT_B200_ERR_ARG;
    *out = nullptr;
    return guarded([&] {
        if (n % 2 != 0) return NTT_B200_NONE;
        std::vector<uint64_t> primes(factors, factors + nfactors);
        std::sort(primes.begin(), primes.end());
        uint64_t prev = 0;  // zeros / duplicates
        for (uint64_t f : primes) {
            if (f == prev) return NTT_B200_NONE;
            prev = f;
        }
        primes.erase(primes.begin(), std::find_if(primes.begin(), primes.end(), [](uint64_t f) { return f != 1; }));
        u128 prod = 1;
        for (uint64_t f : primes) {
            prod *= f;
            if (prod >> 64) return NTT_B200_NONE;
        }
        if ((uint64_t)prod != modulus) return NTT_B200_NONE;
        if ((int)primes.size() > kMaxPrimes) return NTT_B200_ERR_ARG;
        // every per-prime try_new must succeed (the `?` of product.rs:183-191); decided on the
        // host before touching CUDA
        for (uint64_t f : primes) {
            bool ok = f < (uint64_t(1) << 32) ? pm::build_twiddles(n, f, 32).has_value()
                                              : pm::build_twiddles(n, f, 16).has_value();
            if (!ok) return NTT_B200_NONE;
        }
        auto pl = std::make_unique<ntt_b200_product_plan>();
        pl->n = n;
        pl->modulus = modulus;
        NTT_CUDA_CHECK(cudaGetDevice(&pl->device));
        ProductConsts& k = pl->k;
        k.modulus = modulus;
        for (size_t j = 0; j < primes.size(); ++j) {
            uint64_t f = primes[j];
            k.p[j] = f;
            if (f < (uint64_t(1) << 32)) {
                k.n32++;
                k.b64[j] = ~uint64_t(0) / f;
                pl->plans.push_back(make_plan32(n, (uint32_t)f));
            } else {
                k.n64++;
                pl->plans.push_back(make_plan64(n, f));
            }
            if (!pl->plans.back()) return NTT_B200_NONE;
        }
        int offset = 0;
        for (size_t j = 0; j < primes.size(); ++j) {
            for (size_t i = 0; i < j; ++i) k.inv[offset + i] = modular_inverse(primes[i], primes[j]);
            offset += (int)j;
        }
        *out = pl.release();
        return NTT_B200_OK;
    });
}
void ntt_b200_product_free(ntt_b200_product_plan* plan) { delete plan; }
size_t ntt_b200_product_ntt_size(const ntt_b200_product_plan* plan) { return plan ? plan->n : 0; }
uint64_t ntt_b200_product_modulus(const ntt_b200_product_plan* plan) { return plan ? plan->modulus : 0; }
size_t ntt_b200_product_ntt_domain_len(const ntt_b200_product_plan* plan) { return plan->domain_len(); }

int ntt_b200_product_fwd_device(const ntt_b200_product_plan* plan, uint64_t* ntt, const uint64_t* standard,
                                size_t batch, void* stream) {
    if (!plan || (batch && (!ntt || !standard))) return NTT_B200_ERR_ARG;
    return guarded([&] {
        DeviceGuard g(plan->device);
        plan->fwd_dev(ntt, standard, batch, (cudaStream_t)stream);
        return NTT_B200_OK;
    });
}
int ntt_b200_product_inv_device(const ntt_b200_product_plan* plan, uint64_t* standard, uint64_t* ntt,
                                size_t batch, int accumulate, void* stream) {
    if (!plan || (batch && (!ntt || !standard))) return NTT_B200_ERR_ARG;
    return guarded([&] {
        DeviceGuard g(plan->device);
        plan->inv_dev(standard, ntt, batch, accumulate != 0, (cudaStream_t)stream);
        return NTT_B200_OK;
    });
}

// Plan::fwd(&self, ntt, standard, mode)   product.rs:273-357.  FwdMode::Bounded yields the same
// residues as Generic for every input that honours its bound, so one exact path serves both.
int ntt_b200_product_fwd(const ntt_b200_product_plan* plan, uint64_t* ntt, size_t ntt_len,
                         const uint64_t* standard, size_t standard_len) {
    if (!plan || !ntt || !standard) return NTT_B200_ERR_ARG;
    if (standard_len != plan->n || ntt_len != plan->domain_len()) return NTT_B200_ERR_LEN;
    return guarded([&] {
        DeviceGuard g(plan->device);
        HostStage2 hs(plan->device);
        void* ds = hs.upload(standard, standard_len * 8);
        void* dn = hs.alloc(ntt_len * 8);
        plan->fwd_dev(dn, ds, 1, hs.st);
        hs.download(ntt, dn, ntt_len * 8);
        hs.finish();
        return NTT_B200_OK;
    });
}
// Plan::inv(&self, standard, ntt, mode)   product.rs:360-880; ntt is transformed in place
int ntt_b200_product_inv(const ntt_b200_product_plan* plan, uint64_t* standard, size_t standard_len,
                         uint64_t* ntt, size_t ntt_len, int accumulate) {
    if (!plan || !ntt || !standard) return NTT_B200_ERR_ARG;
    if (standard_len != plan->n || ntt_len != plan->domain_len()) return NTT_B200_ERR_LEN;
    return guarded([&] {
        DeviceGuard g(plan->device);
        HostStage2 hs(plan->device);
        void* ds = hs.upload(standard, standard_len * 8);
        void* dn = hs.upload(ntt, ntt_len * 8);
        plan->inv_dev(ds, dn, 1, accumulate != 0, hs.st);
        hs.download(standard, ds, standard_len * 8);
        hs.download(ntt, dn, ntt_len * 8);
        hs.finish();
        return NTT_B200_OK;
    });
}
// Plan::normalize / mul_assign_normalize / mul_accumulate   product.rs:885-967
static int product_pointwise(const ntt_b200_product_plan* plan, int op, uint64_t* a, size_t a_len,
                             const uint64_t* b, size_t b_len, const uint64_t* c, size_t c_len) {
    if (!plan || !a) return NTT_B200_ERR_ARG;
    size_t len = plan->domain_len();
    if (a_len != len || (b && b_len != len) || (c && c_len != len)) return NTT_B200_ERR_LEN;
    return guarded([&] {
        DeviceGuard g(plan->device);
        HostStage2 hs(plan->device);
        void* da = hs.upload(a, len * 8);
        void* db = b ? hs.upload(b, len * 8) : nullptr;
        void* dc = c ? hs.upload(c, len * 8) : nullptr;
        plan->pointwise_dev(op, da, db, dc, 1, hs.st);
        hs.download(a, da, len * 8);
        hs.finish();
        return NTT_B200_OK;
    });
}
int ntt_b200_product_normalize(const ntt_b200_product_plan* plan, uint64_t* values, size_t len) {
    return product_pointwise(plan, 0, values, len, nullptr, 0, nullptr, 0);
}
int ntt_b200_product_mul_assign_normalize(const ntt_b200_product_plan* plan, uint64_t* lhs, size_t lhs_len,
                                          const uint64_t* rhs, size_t rhs_len) {
    if (!rhs) return NTT_B200_ERR_ARG;
    return product_pointwise(plan, 1, lhs, lhs_len, rhs, rhs_len, nullptr, 0);
}
int ntt_b200_product_mul_accumulate(const ntt_b200_product_plan* plan, uint64_t* acc, size_t acc_len,
                                    const uint64_t* lhs, size_t lhs_len, const uint64_t* rhs,
                                    size_t rhs_len) {
    if (!lhs || !rhs) return NTT_B200_ERR_ARG;
    return product_pointwise(plan, 2, acc, acc_len, lhs, lhs_len, rhs, rhs_len);
}

}  // extern "C"
