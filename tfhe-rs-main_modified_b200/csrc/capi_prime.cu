// C ABI for the prime plans (include/tfhe_ntt_b200.h).  No torch, no oracle, no CPU fallback:
// every compute entry point runs CUDA kernels or fails with NTT_B200_ERR_CUDA.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "capi_common.cuh"
#include "plan_math.hpp"

using namespace nttb200;

namespace nttb200 {
thread_local std::string g_last_error;
}

extern "C" {

const char* ntt_b200_last_error(void) { return g_last_error.c_str(); }

int ntt_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
int ntt_b200_set_device(int device) {
    return guarded([&] {
        NTT_CUDA_CHECK(cudaSetDevice(device));
        return NTT_B200_OK;
    });
}

int ntt_b200_is_prime64(uint64_t n) { return pm::is_prime(n) ? 1 : 0; }
int ntt_b200_largest_prime_in_arithmetic_progression64(uint64_t factor, uint64_t offset,
                                                       uint64_t lo, uint64_t hi, uint64_t* out) {
    auto r = pm::largest_prime_in_arithmetic_progression64(factor, offset, lo, hi);
    if (!r) return 0;
    if (out) *out = *r;
    return 1;
}

}  // extern "C"

namespace {

// ---- host-pointer paths: stage through device memory, chunked and double-buffered ---------
constexpr int kMaxStreams = 8;
// Staging chunk and the number of streams the H2D / kernel / D2H stages of successive chunks
// overlap across; tunable for experiments through NTT_B200_CHUNK_MIB / NTT_B200_STREAMS.
size_t chunk_bytes() {
    static const size_t v = [] {
        const char* e = std::getenv("NTT_B200_CHUNK_MIB");
        long m = e ? std::atol(e) : 0;
        return size_t(m > 0 && m <= 1024 ? m : 32) << 20;  // tools/e2e_sweep.py: 32 MiB x 4 streams
    }();
    return v;
}
int stream_count() {
    static const int v = [] {
        const char* e = std::getenv("NTT_B200_STREAMS");
        long m = e ? std::atol(e) : 0;
        return (int)(m > 0 && m <= kMaxStreams ? m : 4);
    }();
    return v;
}
#define kChunkBytes chunk_bytes()
#define kStreams stream_count()

// Per-thread, per-device resources for the small host calls (the reference's per-polynomial Plan::fwd /
// inv / normalize / mul_* on one slice): a stream, a device buffer and a pinned staging buffer that are
// created once and reused.  Thread-local, so calls stay re-entrant on a shared plan; never freed (a
// thread-exit destructor could run after the CUDA context is gone).
//
// Calls of at most kZeroCopyBytes whose transform is ONE kernel (n <= 4096) skip the device buffer and
// both DMA copies: the pinned staging buffer is mapped into the device's address space (cudaMallocHost
// under unified addressing), so the kernel reads the polynomial over PCIe with its first-pass loads and
// writes the result back with its last-pass stores -- one launch and one wait per call
// (profiles/r02_latency.txt).  Larger small calls (<= kSmallBytes) stage through the cached device buffer.
struct SmallCtx {
    cudaStream_t st = nullptr;
    char *d = nullptr, *h = nullptr;
    size_t cap = 0;
    // (a stream-ordered flag write + host spin instead of cudaStreamSynchronize was measured slower:
    // 20.3 against 17.5 us per call, profiles/r02_latency.txt)
    void wait_done() { NTT_CUDA_CHECK(cudaStreamSynchronize(st)); }
    void ensure(size_t bytes) {
        if (!st) NTT_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (bytes <= cap) return;
        size_t want = std::max<size_t>(bytes, size_t(256) << 10);
        if (d) cudaFree(d);
        if (h) cudaFreeHost(h);
        d = h = nullptr;
        cap = 0;
        NTT_CUDA_CHECK(cudaMalloc(&d, want));
        NTT_CUDA_CHECK(cudaMallocHost(&h, want));
        cap = want;
    }
};
constexpr size_t kSmallBytes = size_t(2) << 20;      // calls up to this size take the cached path
constexpr size_t kZeroCopyBytes = size_t(64) << 10;  // ... and up to this size the mapped-pinned path
bool zero_copy_enabled() {
    static const bool v = [] {
        const char* e = std::getenv("NTT_B200_ZERO_COPY");
        return !(e && e[0] == '0');
    }();
    return v;
}
SmallCtx& small_ctx(int device) {
    thread_local std::vector<SmallCtx> ctx;
    if ((size_t)device >= ctx.size()) ctx.resize((size_t)device + 1);
    return ctx[device];
}

int host_transform_small(const PrimePlan* pl, void* host, size_t batch, bool inverse) {
    NTT_NVTX("ntt_b200::host_transform_small");
    return guarded([&] {
        DeviceGuard g(pl->device);
        const size_t bytes = batch * pl->n * (size_t)pl->elem_bytes;
        SmallCtx& c = small_ctx(pl->device);
        c.ensure(bytes);
        std::memcpy(c.h, host, bytes);
        const bool mapped = zero_copy_enabled() && bytes <= kZeroCopyBytes && pl->logn <= 12;
        char* data = mapped ? c.h : c.d;
        if (!mapped) NTT_CUDA_CHECK(cudaMemcpyAsync(c.d, c.h, bytes, cudaMemcpyHostToDevice, c.st));
        if (inverse)
            pl->inv(data, batch, c.st);
        else
            pl->fwd(data, batch, c.st);
        if (!mapped) NTT_CUDA_CHECK(cudaMemcpyAsync(c.h, c.d, bytes, cudaMemcpyDeviceToHost, c.st));
        c.wait_done();
        std::memcpy(host, c.h, bytes);
        return NTT_B200_OK;
    });
}

int host_transform(const PrimePlan* pl, void* host, size_t batch, bool inverse) {
    if (batch && batch * pl->n * (size_t)pl->elem_bytes <= kSmallBytes)
        return host_transform_small(pl, host, batch, inverse);
    NTT_NVTX("ntt_b200::host_transform (chunked H2D / kernel / D2H pipeline)");
    return guarded([&] {
        if (!batch) return NTT_B200_OK;
        DeviceGuard g(pl->device);
        keep_pool_cached(pl->device);
        size_t poly_bytes = pl->n * (size_t)pl->elem_bytes;
        size_t chunk_polys = std::max<size_t>(1, kChunkBytes / poly_bytes);
        chunk_polys = std::min(chunk_polys, batch);
        size_t nchunks = (batch + chunk_polys - 1) / chunk_polys;
        int ns = (int)std::min<size_t>(kStreams, nchunks);
        ScopedStream sg[kMaxStreams];  // synchronised and destroyed on every exit path
        void* dbuf[kMaxStreams] = {};
        for (int i = 0; i < ns; ++i) {
            sg[i].open();
            dbuf[i] = sg[i].alloc(chunk_polys * poly_bytes);
        }
        for (size_t c = 0; c < nchunks; ++c) {
            int i = (int)(c % ns);
            size_t b0 = c * chunk_polys, nb = std::min(chunk_polys, batch - b0);
            char* h = static_cast<char*>(host) + b0 * poly_bytes;
            NTT_CUDA_CHECK(cudaMemcpyAsync(dbuf[i], h, nb * poly_bytes, cudaMemcpyHostToDevice, sg[i].st));
            if (inverse)
                pl->inv(dbuf[i], nb, sg[i].st);
            else
                pl->fwd(dbuf[i], nb, sg[i].st);
            NTT_CUDA_CHECK(cudaMemcpyAsync(h, dbuf[i], nb * poly_bytes, cudaMemcpyDeviceToHost, sg[i].st));
        }
        cudaError_t first = cudaSuccess;
        for (int i = 0; i < ns; ++i) {
            cudaError_t e = sg[i].finish();
            if (first == cudaSuccess) first = e;
        }
        NTT_CUDA_CHECK(first);
        return NTT_B200_OK;
    });
}

// The same host batch split over several GPUs of one box: contiguous slices, batch / G each with the
// remainder one by one to the first GPUs (the reference CUDA backend's rule,
// backends/tfhe-cuda-backend/cuda/src/utils/helper_multi_gpu.cu:57-88), one host thread and one set
// of staging streams per GPU, no exchange between GPUs (the polynomials are independent).
int host_transform_multi(const PrimePlan* const* plans, size_t n_plans, void* host, size_t batch, bool inverse) {
    if (n_plans == 1) return host_transform(plans[0], host, batch, inverse);
    const size_t poly_bytes = plans[0]->n * (size_t)plans[0]->elem_bytes;
    std::vector<int> status(n_plans, NTT_B200_OK);
    std::vector<std::string> message(n_plans);
    std::vector<std::thread> workers;
    const size_t base = batch / n_plans, rem = batch % n_plans;
    size_t begin = 0;
    for (size_t g = 0; g < n_plans; ++g) {
        const size_t count = base + (g < rem ? 1 : 0);
        char* slice = static_cast<char*>(host) + begin * poly_bytes;
        begin += count;
        if (!count) continue;
        workers.emplace_back([=, &status, &message] {
            status[g] = host_transform(plans[g], slice, count, inverse);
            if (status[g] != NTT_B200_OK) message[g] = g_last_error;  // thread-local in the worker
        });
    }
    for (auto& w : workers) w.join();
    for (size_t g = 0; g < n_plans; ++g)
        if (status[g] != NTT_B200_OK) {
            g_last_error = "GPU slice " + std::to_string(g) + ": " + message[g];
            return status[g];
        }
    return NTT_B200_OK;
}

// dst[0..len) op= ...; operands uploaded whole (pointwise calls are per-polynomial sized)
int host_pointwise_small(const PrimePlan* pl, int op, void* dst, size_t len, const void* a, size_t a_len,
                         const void* b, size_t b_len) {
    return guarded([&] {
        DeviceGuard g(pl->device);
        const size_t eb = (size_t)pl->elem_bytes;
        auto up16 = [](size_t x) { return (x + 15) & ~size_t(15); };
        const size_t o_a = up16(len * eb), o_b = o_a + up16(a_len * eb), total = o_b + up16(b_len * eb);
        SmallCtx& c = small_ctx(pl->device);
        c.ensure(total);
        std::memcpy(c.h, dst, len * eb);
        if (a) std::memcpy(c.h + o_a, a, a_len * eb);
        if (b) std::memcpy(c.h + o_b, b, b_len * eb);
        const bool mapped = zero_copy_enabled() && total <= kZeroCopyBytes;
        char* w = mapped ? c.h : c.d;  // the kernels read and write the mapped pinned buffer directly
        if (!mapped) NTT_CUDA_CHECK(cudaMemcpyAsync(c.d, c.h, total, cudaMemcpyHostToDevice, c.st));
        if (op == 0) pl->normalize(w, len, c.st);
        if (op == 1) pl->mul_assign_normalize(w, w + o_a, len, a_len, c.st);
        if (op == 2) pl->mul_accumulate(w, w + o_a, w + o_b, len, a_len, b_len, c.st);
        if (!mapped) NTT_CUDA_CHECK(cudaMemcpyAsync(c.h, c.d, len * eb, cudaMemcpyDeviceToHost, c.st));
        c.wait_done();
        std::memcpy(dst, c.h, len * eb);
        return NTT_B200_OK;
    });
}

int host_pointwise(const PrimePlan* pl, int op, void* dst, size_t len, const void* a, size_t a_len,
                   const void* b, size_t b_len) {
    if (len && (len + a_len + b_len) * (size_t)pl->elem_bytes + 64 <= kSmallBytes)
        return host_pointwise_small(pl, op, dst, len, a, a ? a_len : 0, b, b ? b_len : 0);
    return guarded([&] {
        if (!len) return NTT_B200_OK;
        DeviceGuard g(pl->device);
        keep_pool_cached(pl->device);
        size_t eb = (size_t)pl->elem_bytes;
        ScopedStream s;
        s.open();
        auto up = [&](const void* h, size_t n) {
            void* d = s.alloc(n * eb);
            NTT_CUDA_CHECK(cudaMemcpyAsync(d, h, n * eb, cudaMemcpyHostToDevice, s.st));
            return d;
        };
        void* dd = up(dst, len);
        void* da = a ? up(a, a_len) : nullptr;
        void* db = b ? up(b, b_len) : nullptr;
        if (op == 0) pl->normalize(dd, len, s.st);
        if (op == 1) pl->mul_assign_normalize(dd, da, len, a_len, s.st);
        if (op == 2) pl->mul_accumulate(dd, da, db, len, a_len, b_len, s.st);
        NTT_CUDA_CHECK(cudaMemcpyAsync(dst, dd, len * eb, cudaMemcpyDeviceToHost, s.st));
        NTT_CUDA_CHECK(s.finish());
        return NTT_B200_OK;
    });
}

// out = inv(acc + fwd(lhs) * rhs) for host operands, chunked and pipelined on several streams
int host_fwd_mac_inv(const PrimePlan* pl, void* out, const void* lhs, const void* rhs, size_t rhs_polys,
                     const void* acc, size_t acc_polys, size_t batch) {
    NTT_NVTX("ntt_b200::host_fwd_mac_inv (chunked pipeline)");
    return guarded([&] {
        if (!batch) return NTT_B200_OK;
        DeviceGuard g(pl->device);
        keep_pool_cached(pl->device);
        const size_t pb = pl->n * (size_t)pl->elem_bytes;
        const bool rhs_shared = rhs_polys < batch, acc_shared = acc && acc_polys < batch;
        // chunk boundaries must fall on multiples of the shared operands' periods
        size_t period = 1;
        if (rhs_shared) period = rhs_polys;
        if (acc_shared) {
            size_t a = period, b = acc_polys;
            while (b) {
                size_t t = a % b;
                a = b;
                b = t;
            }
            period = period / a * acc_polys;
        }
        size_t chunk = std::max<size_t>(1, kChunkBytes / pb);
        chunk = std::max(period, chunk / period * period);
        chunk = std::min(chunk, batch);
        size_t nchunks = (batch + chunk - 1) / chunk;
        int ns = (int)std::min<size_t>(kStreams, nchunks);
        ScopedEvent shared_ready;      // declared before the streams: destroyed after they are drained
        ScopedStream sg[kMaxStreams];  // synchronised and destroyed on every exit path
        void *d_io[kMaxStreams] = {}, *d_rhs[kMaxStreams] = {}, *d_acc[kMaxStreams] = {};
        void *d_rhs_shared = nullptr, *d_acc_shared = nullptr;
        for (int i = 0; i < ns; ++i) {
            sg[i].open();
            d_io[i] = sg[i].alloc(chunk * pb);
            if (!rhs_shared) d_rhs[i] = sg[i].alloc(chunk * pb);
            if (acc && !acc_shared) d_acc[i] = sg[i].alloc(chunk * pb);
        }
        if (rhs_shared || acc_shared) {
            shared_ready.create();
            if (rhs_shared) {
                d_rhs_shared = sg[0].alloc(rhs_polys * pb);
                NTT_CUDA_CHECK(cudaMemcpyAsync(d_rhs_shared, rhs, rhs_polys * pb, cudaMemcpyHostToDevice, sg[0].st));
            }
            if (acc_shared) {
                d_acc_shared = sg[0].alloc(acc_polys * pb);
                NTT_CUDA_CHECK(cudaMemcpyAsync(d_acc_shared, acc, acc_polys * pb, cudaMemcpyHostToDevice, sg[0].st));
            }
            NTT_CUDA_CHECK(cudaEventRecord(shared_ready.ev, sg[0].st));
            for (int i = 1; i < ns; ++i) NTT_CUDA_CHECK(cudaStreamWaitEvent(sg[i].st, shared_ready.ev, 0));
        }
        for (size_t c = 0; c < nchunks; ++c) {
            int i = (int)(c % ns);
            cudaStream_t st = sg[i].st;
            size_t b0 = c * chunk, nb = std::min(chunk, batch - b0);
            const char* hl = static_cast<const char*>(lhs) + b0 * pb;
            char* ho = static_cast<char*>(out) + b0 * pb;
            NTT_CUDA_CHECK(cudaMemcpyAsync(d_io[i], hl, nb * pb, cudaMemcpyHostToDevice, st));
            const void* r = d_rhs_shared;
            size_t rp = rhs_polys;
            if (!rhs_shared) {
                NTT_CUDA_CHECK(cudaMemcpyAsync(d_rhs[i], static_cast<const char*>(rhs) + b0 * pb, nb * pb,
                                               cudaMemcpyHostToDevice, st));
                r = d_rhs[i];
                rp = nb;
            }
            const void* a = nullptr;
            size_t ap = 0;
            if (acc) {
                a = d_acc_shared;
                ap = acc_polys;
                if (!acc_shared) {
                    NTT_CUDA_CHECK(cudaMemcpyAsync(d_acc[i], static_cast<const char*>(acc) + b0 * pb, nb * pb,
                                                   cudaMemcpyHostToDevice, st));
                    a = d_acc[i];
                    ap = nb;
                }
            }
            pl->fwd_mac_inv(d_io[i], d_io[i], r, rp, a, ap, nb, st);
            NTT_CUDA_CHECK(cudaMemcpyAsync(ho, d_io[i], nb * pb, cudaMemcpyDeviceToHost, st));
        }
        // the shared operands live in stream 0's pool: every other stream finishes first
        cudaError_t first = cudaSuccess;
        for (int i = ns - 1; i >= 0; --i) {
            cudaError_t e = sg[i].finish();
            if (first == cudaSuccess) first = e;
        }
        NTT_CUDA_CHECK(first);
        return NTT_B200_OK;
    });
}

int dev_pointwise_check(size_t len, size_t sub_len) {
    if (sub_len == 0 || sub_len > len || len % sub_len != 0) return NTT_B200_ERR_LEN;
    return NTT_B200_OK;
}


// argument checks shared by the *_batch_multi_gpu entry points: same (n, p) everywhere, one plan per GPU slice
template <class PlanT, class ELEM>
int multi_entry(const PlanT* const* plans, size_t n_plans, ELEM* host, size_t batch, bool inverse) {
    if (!plans || !n_plans || n_plans > 64 || (!host && batch)) return NTT_B200_ERR_ARG;
    std::vector<const PrimePlan*> impl(n_plans);
    for (size_t g = 0; g < n_plans; ++g) {
        if (!plans[g]) return NTT_B200_ERR_ARG;
        impl[g] = plans[g]->impl.get();
        if (impl[g]->n != impl[0]->n || impl[g]->p != impl[0]->p) return NTT_B200_ERR_ARG;
    }
    return host_transform_multi(impl.data(), n_plans, host, batch, inverse);
}
}  // namespace

// ---- tfhe Ntt64View on top of a prime64 plan (tfhe .../math/ntt/ntt64.rs:89-266) ------------
// Element-wise pre / post maps around the batched transforms.
// forward modes: 0 forward, 1 forward_normalized, 2 forward_from_decomp,
//                3 forward_from_power_of_two_modulus(width)
__global__ void ntt64_premap_kernel(uint64_t* __restrict__ ntt, const uint64_t* __restrict__ standard,
                                    size_t total, int mode, unsigned width, uint64_t p) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = standard[i];
        if (mode == 2) {  // small signed digits -> [0,p)   (ntt64.rs:229-236)
            if ((int64_t)v < 0) v += p;
        } else if (mode == 3) {  // modswitch 2^width -> p, rounded   (ntt64.rs:165-177)
            unsigned __int128 x = (unsigned __int128)(v >> (64 - width)) * p +
                                  ((unsigned __int128)1 << (width - 1));
            v = (uint64_t)(x >> width);
        }
        ntt[i] = v;
    }
}
// add_backward modes: 0 add modulo p (ntt64.rs:110-131), 1 modswitch p -> 2^width, then wrapping
// add (ntt64.rs:184-196, :242-266; the modswitched values are left in `ntt` like the reference)
__global__ void ntt64_postmap_kernel(uint64_t* __restrict__ standard, uint64_t* __restrict__ ntt,
                                     size_t total, int mode, unsigned width, uint64_t p) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        uint64_t a = standard[i], b = ntt[i];
        if (mode == 0) {
            // wrapping_add_custom_mod: a - neg(b) mod p (tfhe .../numeric/unsigned.rs:174-190)
            uint64_t nb = b == 0 ? 0 : p - b;
            standard[i] = a >= nb ? a - nb : a - nb + p;
        } else {
            unsigned __int128 x = ((unsigned __int128)b << width) | ((unsigned __int128)p >> 1);
            uint64_t m = (uint64_t)(x / p) << (64 - width);
            ntt[i] = m;
            standard[i] = a + m;
        }
    }
}

int ntt64_forward_dev(const PrimePlan* pl, uint64_t* ntt, const uint64_t* standard, size_t batch, int mode,
                      unsigned width, cudaStream_t st) {
    if (mode < 0 || mode > 3 || (mode == 3 && (width == 0 || width > 64))) return NTT_B200_ERR_ARG;
    return guarded([&] {
        if (!batch) return NTT_B200_OK;
        DeviceGuard g(pl->device);
        size_t total = batch * pl->n;
        if (mode <= 1) {
            if (ntt != standard)
                NTT_CUDA_CHECK(cudaMemcpyAsync(ntt, standard, total * 8, cudaMemcpyDeviceToDevice, st));
        } else {
            unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
            ntt64_premap_kernel<<<blocks, 256, 0, st>>>(ntt, standard, total, mode, width, pl->p);
            NTT_CUDA_CHECK(cudaGetLastError());
        }
        pl->fwd(ntt, batch, st);
        if (mode == 1) pl->normalize(ntt, total, st);
        return NTT_B200_OK;
    });
}
int ntt64_add_backward_dev(const PrimePlan* pl, uint64_t* standard, uint64_t* ntt, size_t batch, int mode,
                           unsigned width, cudaStream_t st) {
    if (mode < 0 || mode > 1 || (mode == 1 && (width == 0 || width > 64))) return NTT_B200_ERR_ARG;
    return guarded([&] {
        if (!batch) return NTT_B200_OK;
        DeviceGuard g(pl->device);
        size_t total = batch * pl->n;
        pl->inv(ntt, batch, st);
        unsigned blocks = (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16);
        ntt64_postmap_kernel<<<blocks, 256, 0, st>>>(standard, ntt, total, mode, width, pl->p);
        NTT_CUDA_CHECK(cudaGetLastError());
        return NTT_B200_OK;
    });
}

// The 32- and 64-bit families share every code path; only the handle and element types differ.
#define NTT_DEFINE_PRIME_API(SFX, ELEM, MAKE)                                                      \
    extern "C" {                                                                                   \
    int ntt_b200_plan##SFX##_try_new(size_t n, ELEM p, ntt_b200_plan##SFX** out) {                 \
        if (!out) return NTT_B200_ERR_ARG;                                                         \
        *out = nullptr;                                                                            \
        return guarded([&] {                                                                       \
            auto impl = MAKE(n, p);                                                                \
            if (!impl) return NTT_B200_NONE;                                                       \
            *out = new ntt_b200_plan##SFX{impl};                                                   \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_clone(const ntt_b200_plan##SFX* plan, ntt_b200_plan##SFX** out) {     \
        if (!plan || !out) return NTT_B200_ERR_ARG;                                                \
        *out = nullptr;                                                                            \
        return guarded([&] {                                                                       \
            *out = new ntt_b200_plan##SFX{plan->impl->clone()};                                    \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    void ntt_b200_plan##SFX##_free(ntt_b200_plan##SFX* plan) { delete plan; }                      \
    size_t ntt_b200_plan##SFX##_ntt_size(const ntt_b200_plan##SFX* plan) {                         \
        return plan ? plan->impl->n : 0; /* a null handle reads as the empty plan */               \
    }                                                                                              \
    ELEM ntt_b200_plan##SFX##_modulus(const ntt_b200_plan##SFX* plan) {                            \
        return plan ? (ELEM)plan->impl->p : (ELEM)0;                                               \
    }                                                                                              \
    int ntt_b200_plan##SFX##_can_use_fast_reduction_code(const ntt_b200_plan##SFX* plan) {         \
        return plan && plan->impl->can_use_fast_reduction_code ? 1 : 0;                            \
    }                                                                                              \
    int ntt_b200_plan##SFX##_device(const ntt_b200_plan##SFX* plan) {                              \
        return plan ? plan->impl->device : -1;                                                     \
    }                                                                                              \
    int ntt_b200_plan##SFX##_fwd(const ntt_b200_plan##SFX* plan, ELEM* buf, size_t len) {          \
        if (!plan || !buf) return NTT_B200_ERR_ARG;                                                \
        if (len != plan->impl->n) return NTT_B200_ERR_LEN;                                         \
        return host_transform(plan->impl.get(), buf, 1, false);                                    \
    }                                                                                              \
    int ntt_b200_plan##SFX##_inv(const ntt_b200_plan##SFX* plan, ELEM* buf, size_t len) {          \
        if (!plan || !buf) return NTT_B200_ERR_ARG;                                                \
        if (len != plan->impl->n) return NTT_B200_ERR_LEN;                                         \
        return host_transform(plan->impl.get(), buf, 1, true);                                     \
    }                                                                                              \
    int ntt_b200_plan##SFX##_fwd_batch(const ntt_b200_plan##SFX* plan, ELEM* host, size_t batch) { \
        if (!plan || (!host && batch)) return NTT_B200_ERR_ARG;                                    \
        return host_transform(plan->impl.get(), host, batch, false);                               \
    }                                                                                              \
    int ntt_b200_plan##SFX##_inv_batch(const ntt_b200_plan##SFX* plan, ELEM* host, size_t batch) { \
        if (!plan || (!host && batch)) return NTT_B200_ERR_ARG;                                    \
        return host_transform(plan->impl.get(), host, batch, true);                                \
    }                                                                                              \
    int ntt_b200_plan##SFX##_fwd_batch_multi_gpu(const ntt_b200_plan##SFX* const* plans,           \
                                                 size_t n_plans, ELEM* host, size_t batch) {       \
        return multi_entry(plans, n_plans, host, batch, false);                                    \
    }                                                                                              \
    int ntt_b200_plan##SFX##_inv_batch_multi_gpu(const ntt_b200_plan##SFX* const* plans,           \
                                                 size_t n_plans, ELEM* host, size_t batch) {       \
        return multi_entry(plans, n_plans, host, batch, true);                                     \
    }                                                                                              \
    int ntt_b200_plan##SFX##_normalize(const ntt_b200_plan##SFX* plan, ELEM* v, size_t len) {      \
        if (!plan || (!v && len)) return NTT_B200_ERR_ARG;                                         \
        return host_pointwise(plan->impl.get(), 0, v, len, nullptr, 0, nullptr, 0);                \
    }                                                                                              \
    int ntt_b200_plan##SFX##_mul_assign_normalize(const ntt_b200_plan##SFX* plan, ELEM* lhs,       \
                                                  size_t lhs_len, const ELEM* rhs,                 \
                                                  size_t rhs_len) {                                \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        size_t len = std::min(lhs_len, rhs_len); /* izip! truncation, lib.rs:658-688 */            \
        if (len && (!lhs || !rhs)) return NTT_B200_ERR_ARG;                                        \
        return host_pointwise(plan->impl.get(), 1, lhs, len, rhs, len, nullptr, 0);                \
    }                                                                                              \
    int ntt_b200_plan##SFX##_mul_accumulate(const ntt_b200_plan##SFX* plan, ELEM* acc,             \
                                            size_t acc_len, const ELEM* lhs, size_t lhs_len,       \
                                            const ELEM* rhs, size_t rhs_len) {                     \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        size_t len = std::min(acc_len, std::min(lhs_len, rhs_len));                                \
        if (len && (!acc || !lhs || !rhs)) return NTT_B200_ERR_ARG;                                \
        return host_pointwise(plan->impl.get(), 2, acc, len, lhs, len, rhs, len);                  \
    }                                                                                              \
    int ntt_b200_plan##SFX##_fwd_device(const ntt_b200_plan##SFX* plan, ELEM* dev, size_t batch,   \
                                        void* stream) {                                            \
        if (!plan || (!dev && batch)) return NTT_B200_ERR_ARG;                                     \
        return guarded([&] {                                                                       \
            plan->impl->fwd(dev, batch, (cudaStream_t)stream);                                     \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_inv_device(const ntt_b200_plan##SFX* plan, ELEM* dev, size_t batch,   \
                                        void* stream) {                                            \
        if (!plan || (!dev && batch)) return NTT_B200_ERR_ARG;                                     \
        return guarded([&] {                                                                       \
            plan->impl->inv(dev, batch, (cudaStream_t)stream);                                     \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_normalize_device(const ntt_b200_plan##SFX* plan, ELEM* dev,           \
                                              size_t len, void* stream) {                          \
        if (!plan || (!dev && len)) return NTT_B200_ERR_ARG;                                       \
        return guarded([&] {                                                                       \
            plan->impl->normalize(dev, len, (cudaStream_t)stream);                                 \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_mul_assign_normalize_device(const ntt_b200_plan##SFX* plan,           \
                                                         ELEM* lhs, size_t len, const ELEM* rhs,   \
                                                         size_t rhs_len, void* stream) {           \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        if (!len) return NTT_B200_OK;                                                              \
        if (!lhs || !rhs) return NTT_B200_ERR_ARG;                                                 \
        if (int e = dev_pointwise_check(len, rhs_len)) return e;                                   \
        return guarded([&] {                                                                       \
            plan->impl->mul_assign_normalize(lhs, rhs, len, rhs_len, (cudaStream_t)stream);        \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_mul_accumulate_device(const ntt_b200_plan##SFX* plan, ELEM* acc,      \
                                                   size_t len, const ELEM* lhs, size_t lhs_len,    \
                                                   const ELEM* rhs, size_t rhs_len,                \
                                                   void* stream) {                                 \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        if (!len) return NTT_B200_OK;                                                              \
        if (!acc || !lhs || !rhs) return NTT_B200_ERR_ARG;                                         \
        if (int e = dev_pointwise_check(len, lhs_len)) return e;                                   \
        if (int e = dev_pointwise_check(len, rhs_len)) return e;                                   \
        return guarded([&] {                                                                       \
            plan->impl->mul_accumulate(acc, lhs, rhs, len, lhs_len, rhs_len,                       \
                                       (cudaStream_t)stream);                                      \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    int ntt_b200_plan##SFX##_fwd_mac_inv_device(const ntt_b200_plan##SFX* plan, ELEM* out,         \
                                                const ELEM* lhs, const ELEM* rhs,                  \
                                                size_t rhs_polys, const ELEM* acc,                 \
                                                size_t acc_polys, size_t batch, void* stream) {    \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        if (!batch) return NTT_B200_OK;                                                            \
        if (!out || !lhs || !rhs) return NTT_B200_ERR_ARG;                                         \
        if (rhs_polys == 0 || batch % rhs_polys) return NTT_B200_ERR_LEN;                          \
        if (acc && (acc_polys == 0 || batch % acc_polys)) return NTT_B200_ERR_LEN;                 \
        return guarded([&] {                                                                       \
            plan->impl->fwd_mac_inv(out, lhs, rhs, rhs_polys, acc, acc_polys, batch,               \
                                    (cudaStream_t)stream);                                         \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }                                                                                              \
    }

#define NTT_DEFINE_FMI_BATCH(SFX, ELEM)                                                            \
    extern "C" int ntt_b200_plan##SFX##_fwd_mac_inv_batch(                                         \
        const ntt_b200_plan##SFX* plan, ELEM* out, const ELEM* lhs, const ELEM* rhs,               \
        size_t rhs_polys, const ELEM* acc, size_t acc_polys, size_t batch) {                       \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        if (!batch) return NTT_B200_OK;                                                            \
        if (!out || !lhs || !rhs) return NTT_B200_ERR_ARG;                                         \
        if (rhs_polys == 0 || rhs_polys > batch || batch % rhs_polys) return NTT_B200_ERR_LEN;     \
        if (acc && (acc_polys == 0 || acc_polys > batch || batch % acc_polys))                     \
            return NTT_B200_ERR_LEN;                                                               \
        return host_fwd_mac_inv(plan->impl.get(), out, lhs, rhs, rhs_polys, acc, acc_polys, batch);\
    }
#define NTT_DEFINE_EXT(SFX, ELEM)                                                                  \
    extern "C" int ntt_b200_plan##SFX##_ext_product_device(                                        \
        const ntt_b200_plan##SFX* plan, ELEM* out, const ELEM* in, const ELEM* ggsw, size_t rows,  \
        size_t cols, size_t batch, void* stream) {                                                 \
        if (!plan) return NTT_B200_ERR_ARG;                                                        \
        if (!batch) return NTT_B200_OK;                                                            \
        if (!out || !in || !ggsw) return NTT_B200_ERR_ARG;                                         \
        if (rows == 0 || cols == 0 || rows > 65535 || cols > 65535) return NTT_B200_ERR_LEN;       \
        return guarded([&] {                                                                       \
            keep_pool_cached(plan->impl->device); /* stream-ordered scratch of the launcher */     \
            plan->impl->ext_product(out, in, ggsw, (unsigned)rows, (unsigned)cols, batch,          \
                                    (cudaStream_t)stream);                                         \
            return NTT_B200_OK;                                                                    \
        });                                                                                        \
    }
extern "C" {
int ntt_b200_ntt64_forward_device(const ntt_b200_plan64* plan, uint64_t* ntt, const uint64_t* standard,
                                  size_t batch, int mode, uint32_t width, void* stream) {
    if (!plan || (batch && (!ntt || !standard))) return NTT_B200_ERR_ARG;
    return ntt64_forward_dev(plan->impl.get(), ntt, standard, batch, mode, width, (cudaStream_t)stream);
}
int ntt_b200_ntt64_add_backward_device(const ntt_b200_plan64* plan, uint64_t* standard, uint64_t* ntt,
                                       size_t batch, int mode, uint32_t width, void* stream) {
    if (!plan || (batch && (!ntt || !standard))) return NTT_B200_ERR_ARG;
    return ntt64_add_backward_dev(plan->impl.get(), standard, ntt, batch, mode, width, (cudaStream_t)stream);
}
// host forms: `batch` polynomials each
int ntt_b200_ntt64_forward(const ntt_b200_plan64* plan, uint64_t* ntt, const uint64_t* standard, size_t len,
                           int mode, uint32_t width) {
    if (!plan || !ntt || !standard) return NTT_B200_ERR_ARG;
    size_t n = plan->impl->n;
    if (len == 0 || len % n) return NTT_B200_ERR_LEN;  // izip_eq! / copy_from_slice panic on a mismatch
    int rc = NTT_B200_OK;
    int g = guarded([&] {
        DeviceGuard dg(plan->impl->device);
        keep_pool_cached(plan->impl->device);
        ScopedStream s;  // drained and destroyed on every exit path
        s.open();
        cudaStream_t st = s.st;
        uint64_t* ds = static_cast<uint64_t*>(s.alloc(len * 8));
        uint64_t* dn = static_cast<uint64_t*>(s.alloc(len * 8));
        NTT_CUDA_CHECK(cudaMemcpyAsync(ds, standard, len * 8, cudaMemcpyHostToDevice, st));
        rc = ntt64_forward_dev(plan->impl.get(), dn, ds, len / n, mode, width, st);
        NTT_CUDA_CHECK(cudaMemcpyAsync(ntt, dn, len * 8, cudaMemcpyDeviceToHost, st));
        NTT_CUDA_CHECK(s.finish());
        return NTT_B200_OK;
    });
    return g != NTT_B200_OK ? g : rc;
}
int ntt_b200_ntt64_add_backward(const ntt_b200_plan64* plan, uint64_t* standard, uint64_t* ntt, size_t len,
                                int mode, uint32_t width) {
    if (!plan || !ntt || !standard) return NTT_B200_ERR_ARG;
    size_t n = plan->impl->n;
    if (len == 0 || len % n) return NTT_B200_ERR_LEN;
    int rc = NTT_B200_OK;
    int g = guarded([&] {
        DeviceGuard dg(plan->impl->device);
        keep_pool_cached(plan->impl->device);
        ScopedStream s;  // drained and destroyed on every exit path
        s.open();
        cudaStream_t st = s.st;
        uint64_t* ds = static_cast<uint64_t*>(s.alloc(len * 8));
        uint64_t* dn = static_cast<uint64_t*>(s.alloc(len * 8));
        NTT_CUDA_CHECK(cudaMemcpyAsync(ds, standard, len * 8, cudaMemcpyHostToDevice, st));
        NTT_CUDA_CHECK(cudaMemcpyAsync(dn, ntt, len * 8, cudaMemcpyHostToDevice, st));
        rc = ntt64_add_backward_dev(plan->impl.get(), ds, dn, len / n, mode, width, st);
        NTT_CUDA_CHECK(cudaMemcpyAsync(standard, ds, len * 8, cudaMemcpyDeviceToHost, st));
        NTT_CUDA_CHECK(cudaMemcpyAsync(ntt, dn, len * 8, cudaMemcpyDeviceToHost, st));
        NTT_CUDA_CHECK(s.finish());
        return NTT_B200_OK;
    });
    return g != NTT_B200_OK ? g : rc;
}
}

NTT_DEFINE_EXT(64, uint64_t)
NTT_DEFINE_EXT(32, uint32_t)

NTT_DEFINE_FMI_BATCH(64, uint64_t)
NTT_DEFINE_FMI_BATCH(32, uint32_t)

NTT_DEFINE_PRIME_API(64, uint64_t, make_plan64)
NTT_DEFINE_PRIME_API(32, uint32_t, make_plan32)

extern "C" int ntt_b200_plan64_use_ifma(const ntt_b200_plan64*) { return 0; }
