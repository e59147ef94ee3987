// C ABI for the NTT programmable bootstrap built on the prime64 plans:
// tfhe/src/core_crypto/algorithms/lwe_programmable_bootstrapping/ntt64_pbs.rs (classic, ciphertext
// modulus = NTT prime) and ntt64_bnf_pbs.rs (power-of-two ciphertext modulus), plus the bootstrap
// key conversion of algorithms/lwe_bootstrap_key_conversion.rs:294-447.
//
// Device paths, all batched over LWE ciphertexts:
//  * cluster / fused: one persistent two-CTA cluster (k = 1, level = 1) or one persistent CTA per
//    ciphertext runs the whole blind rotation with the accumulator in shared memory
//    (ntt_pbs_fused.cuh, PrimePlan::blind_rotate) -- Solinas prime, n = 256 ... 4096;
//  * composed: per mask element, rotate/subtract/decompose kernel -> PrimePlan::ext_product ->
//    add kernel.  Any n, any prime64 plan; also the cross-check of the fused path.
// No torch, no oracle, no CPU fallback.
#include <algorithm>
#include <vector>

#include "capi_common.cuh"
#include "pbs_math.cuh"

using namespace nttb200;

struct ntt_b200_bsk {
    std::shared_ptr<PrimePlan> plan;
    size_t n_lwe = 0, glwe_size = 0;
    unsigned base_log = 0, level = 0;
    uint64_t* d_bsk = nullptr;  // [n_lwe][level][glwe_size][glwe_size][n], NTT domain
    uint64_t* d_bsk_tw = nullptr;  // the same key in the plan's twiddle form (fused kernels), or null
    size_t ggsw_len() const { return (size_t)level * glwe_size * glwe_size * plan->n; }
    size_t total_len() const { return n_lwe * ggsw_len(); }
    ~ntt_b200_bsk() {
        if (d_bsk || d_bsk_tw) {
            DeviceGuard g(plan->device);
            cudaFree(d_bsk);
            cudaFree(d_bsk_tw);
        }
    }
};

namespace {

constexpr unsigned kSkip = 0x80000000u;  // flag in the switched mask element: CMUX not executed

unsigned grid_for(size_t total) { return (unsigned)std::min<size_t>((total + 255) / 256, 148 * 16); }

// switched[b][i]: monomial degree of mask element i (i < n_lwe) and of the body (i == n_lwe).
// classic: pbs_modulus_switch_non_native (ntt64_pbs.rs:540-550), skip when the RAW element is 0
// (:257); bnf: `lwe` already holds switched values, skip when 0 (ntt64_bnf_pbs.rs:244).
__global__ void pbs_switch_kernel(unsigned* __restrict__ switched, const uint64_t* __restrict__ lwe,
                                  size_t total, size_t lwe_size, unsigned log2n, uint64_t p, int bnf,
                                  int raw_bnf_input) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        uint64_t v = lwe[i];
        bool body = (i % lwe_size) == lwe_size - 1;
        unsigned d;
        if (bnf) {
            d = (unsigned)(raw_bnf_input ? pbs::modulus_switch(v, log2n + 1) : v);
            if (!body && d == 0) d |= kSkip;
        } else {
            d = (unsigned)pbs::modulus_switch_non_native(v, log2n, p);
            if (!body && v == 0) d |= kSkip;
        }
        switched[i] = d;
    }
}

// acc[b] = lut[b % lut_count], divided by X^body for the classic variant (ntt64_pbs.rs:247-255);
// the bnf variant rotates at the end instead (ntt64_bnf_pbs.rs:264-272).
__global__ void pbs_init_acc_kernel(uint64_t* __restrict__ acc, const uint64_t* __restrict__ lut,
                                    size_t lut_count, const unsigned* __restrict__ switched,
                                    size_t batch, size_t lwe_size, size_t glwe_size, unsigned log2n,
                                    uint64_t p, int rotate) {
    size_t n = (size_t)1 << log2n, per = glwe_size * n, total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t b = i / per, r = i % per, c = r >> log2n, j = r & (n - 1);
        const uint64_t* src = lut + (b % lut_count) * per + c * n;
        if (rotate) {
            unsigned d = switched[b * lwe_size + lwe_size - 1] & ~kSkip;
            acc[i] = pbs::monomial_div_coeff(src, j, d, log2n, p);
        } else {
            acc[i] = src[j];
        }
    }
}

// out[b] = acc[b] / X^body, native arithmetic (ntt64_bnf_pbs.rs:264-272)
__global__ void pbs_final_rotate_kernel(uint64_t* __restrict__ out, const uint64_t* __restrict__ acc,
                                        const unsigned* __restrict__ switched, size_t batch,
                                        size_t lwe_size, size_t glwe_size, unsigned log2n) {
    size_t n = (size_t)1 << log2n, per = glwe_size * n, total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t b = i / per, r = i % per, c = r >> log2n, j = r & (n - 1);
        unsigned d = switched[b * lwe_size + lwe_size - 1] & ~kSkip;
        out[i] = pbs::monomial_div_coeff(acc + b * per + c * n, j, d, log2n, 0);
    }
}

// digits[b][lv * glwe_size + c][j] = level (l - lv) of decompose(acc[b][c] * X^a - acc[b][c])
// (cmux :669-680 + TensorSignedDecompositionLendingIterNonNative iter.rs:640-737; bnf: native
// decomposer, ntt64_bnf_pbs.rs:591-599, digits mapped to [0,p) as forward_from_decomp does)
__global__ void pbs_cmux_decompose_kernel(uint64_t* __restrict__ digits, const uint64_t* __restrict__ acc,
                                          const unsigned* __restrict__ switched, size_t elem,
                                          size_t batch, size_t lwe_size, size_t glwe_size,
                                          unsigned log2n, uint64_t p, unsigned base_log,
                                          unsigned level, int bnf) {
    size_t n = (size_t)1 << log2n, per = glwe_size * n, total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t b = i / per, r = i % per, c = r >> log2n, j = r & (n - 1);
        unsigned a = switched[b * lwe_size + elem];
        uint64_t* dst = digits + b * level * per + c * n + j;
        if (a & kSkip) {
            for (unsigned lv = 0; lv < level; ++lv) dst[lv * per] = 0;
            continue;
        }
        const uint64_t* poly = acc + b * per + c * n;
        uint64_t rot = pbs::monomial_mul_coeff(poly, j, a, log2n, bnf ? 0 : p);
        uint64_t diff = bnf ? rot - poly[j] : pbs::sub_mod(rot, poly[j], p);
        if (bnf) {
            uint64_t state = pbs::init_decomposer_state_native(diff, base_log, level);
            for (unsigned lv = 0; lv < level; ++lv) {
                uint64_t t = pbs::decompose_one_level(base_log, state);
                dst[lv * per] = (int64_t)t < 0 ? t + p : t;
            }
        } else {
            bool neg;
            uint64_t state = pbs::init_state_non_native(diff, base_log, level, p, neg);
            for (unsigned lv = 0; lv < level; ++lv)
                dst[lv * per] = pbs::next_term_non_native(base_log, state, neg, p);
        }
    }
}

// acc += ext (ntt64.rs:110-131 wrapping_add_custom_mod) / acc += modswitch(ext) (ntt64.rs:184-196,
// :242-266; ext was normalised before).  Skipped CMUXes leave the accumulator untouched.
__global__ void pbs_add_kernel(uint64_t* __restrict__ acc, const uint64_t* __restrict__ ext,
                               const unsigned* __restrict__ switched, size_t elem, size_t batch,
                               size_t lwe_size, size_t per, uint64_t p, int bnf, unsigned width) {
    size_t total = batch * per;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t b = i / per;
        if (switched[b * lwe_size + elem] & kSkip) continue;
        acc[i] = bnf ? acc[i] + pbs::modswitch_prime_to_pow2(ext[i], width, p)
                     : pbs::add_mod(acc[i], ext[i], p);
    }
}

// extract_lwe_sample_from_glwe_ciphertext with nth = 0 (glwe_sample_extraction.rs:89-164)
__global__ void pbs_sample_extract_kernel(uint64_t* __restrict__ lwe_out, const uint64_t* __restrict__ glwe,
                                          size_t batch, size_t glwe_size, unsigned log2n, uint64_t modulus) {
    size_t n = (size_t)1 << log2n, k = glwe_size - 1, out_size = k * n + 1, total = batch * out_size;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t b = i / out_size, r = i % out_size;
        const uint64_t* g = glwe + b * glwe_size * n;
        if (r == k * n) {
            lwe_out[i] = g[k * n];
            continue;
        }
        size_t c = r >> log2n, j = r & (n - 1);
        // reverse, negate the first n-1, rotate left by n-1: out[0] = P[0], out[j] = -P[n-j]
        lwe_out[i] = j == 0 ? g[c * n] : pbs::neg_mod(g[c * n + n - j], modulus);
    }
}

struct Scratch {
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s) : st(s) {}
    template <class T>
    T* get(size_t count) {
        void* p = nullptr;
        NTT_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(count, 1) * sizeof(T), st));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, st);
    }
};

// Blind rotation of `batch` accumulators on the device.  `lwe`: [batch][n_lwe+1] raw ciphertexts
// (classic, or bnf with raw_bnf_input) or switched values (bnf).  acc_out: [batch][glwe_size][n].
void blind_rotate_dev(const ntt_b200_bsk* key, const uint64_t* lwe, const uint64_t* lut, size_t lut_count,
                      uint64_t* acc_out, size_t batch, int bnf, unsigned width, int raw_bnf_input,
                      int path, cudaStream_t st) {
    const PrimePlan* pl = key->plan.get();
    const size_t n = pl->n, gs = key->glwe_size, per = gs * n, lwe_size = key->n_lwe + 1;
    const unsigned log2n = (unsigned)pl->logn;
    const uint64_t p = pl->p;
    Scratch sc(st);
    unsigned* switched = sc.get<unsigned>(batch * lwe_size);
    pbs_switch_kernel<<<grid_for(batch * lwe_size), 256, 0, st>>>(switched, lwe, batch * lwe_size, lwe_size,
                                                                 log2n, p, bnf, raw_bnf_input);
    NTT_CUDA_CHECK(cudaGetLastError());
    // path 0: the two-CTA cluster kernel when the shape has one (k = 1, level = 1: lower latency AND
    // higher throughput than one CTA per ciphertext, profiles/r01_pbs_bench.jsonl), else the one-CTA
    // fused kernel, else composed; 1: one-CTA fused only; 2: composed only; 3: cluster only
    if (path == 3 || path == 0) {
        if (pl->blind_rotate(acc_out, lut, lut_count, switched, key->d_bsk, key->d_bsk_tw, key->n_lwe, gs,
                             key->base_log, key->level, batch, bnf, width, 1, st))
            return;
        if (path == 3) throw std::runtime_error("no cluster blind-rotation kernel for this shape");
    }
    if (path != 2 &&
        pl->blind_rotate(acc_out, lut, lut_count, switched, key->d_bsk, key->d_bsk_tw, key->n_lwe, gs,
                         key->base_log, key->level, batch, bnf, width, 0, st))
        return;
    if (path == 1) throw std::runtime_error("no fused blind-rotation kernel for this shape");
    uint64_t* acc = bnf ? sc.get<uint64_t>(batch * per) : acc_out;
    uint64_t* digits = sc.get<uint64_t>(batch * key->level * per);
    uint64_t* ext = sc.get<uint64_t>(batch * per);
    pbs_init_acc_kernel<<<grid_for(batch * per), 256, 0, st>>>(acc, lut, lut_count, switched, batch, lwe_size,
                                                              gs, log2n, p, bnf ? 0 : 1);
    NTT_CUDA_CHECK(cudaGetLastError());
    for (size_t i = 0; i < key->n_lwe; ++i) {
        pbs_cmux_decompose_kernel<<<grid_for(batch * per), 256, 0, st>>>(
            digits, acc, switched, i, batch, lwe_size, gs, log2n, p, key->base_log, key->level, bnf);
        NTT_CUDA_CHECK(cudaGetLastError());
        pl->ext_product(ext, digits, key->d_bsk + i * key->ggsw_len(), (unsigned)(key->level * gs),
                        (unsigned)gs, batch, st);
        if (bnf) pl->normalize(ext, batch * per, st);
        pbs_add_kernel<<<grid_for(batch * per), 256, 0, st>>>(acc, ext, switched, i, batch, lwe_size, per, p,
                                                             bnf, width);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
    if (bnf) {
        pbs_final_rotate_kernel<<<grid_for(batch * per), 256, 0, st>>>(acc_out, acc, switched, batch, lwe_size,
                                                                      gs, log2n);
        NTT_CUDA_CHECK(cudaGetLastError());
    }
}

// second copy of the key in the form the fused kernels multiply by (when the family has one)
void make_twiddle_form(ntt_b200_bsk* key) {
    size_t total = key->total_len();
    NTT_CUDA_CHECK(cudaMalloc(&key->d_bsk_tw, total * 8));
    bool ok = key->plan->key_to_twiddle_form(key->d_bsk_tw, key->d_bsk, key->n_lwe * key->level, key->glwe_size,
                                             nullptr);
    NTT_CUDA_CHECK(cudaDeviceSynchronize());
    if (!ok) {
        cudaFree(key->d_bsk_tw);
        key->d_bsk_tw = nullptr;
    }
}

int check_key(const ntt_b200_bsk* key, int bnf, unsigned width) {
    if (!key) return NTT_B200_ERR_ARG;
    if (bnf && (width == 0 || width > 64)) return NTT_B200_ERR_ARG;
    if (!bnf) {
        // the classic decomposer keeps base_log * level bits of a ceil(log2 p)-bit value
        // (init_state_non_native shifts by their difference; the reference's
        // SignedDecomposerNonNative::new asserts the same bound, decomposer.rs:487-520)
        const uint64_t p = key->plan->p;
        const unsigned ceil_log2_p = 64 - (unsigned)__builtin_clzll(p - 1);
        if ((uint64_t)key->base_log * key->level > ceil_log2_p) return NTT_B200_ERR_ARG;
    }
    return NTT_B200_OK;
}

// Host entry: stage everything through device memory in chunks of ciphertexts.
constexpr size_t kChunkCts = 2048;

int host_blind_rotate(const ntt_b200_bsk* key, const uint64_t* lwe, uint64_t* lut_inout,
                      const uint64_t* accumulator, size_t acc_count, uint64_t* lwe_out, size_t batch, int bnf,
                      unsigned width, int raw_bnf_input, int path) {
    // lut_inout != null: blind_rotate_*_assign (per-ciphertext lut, rotated in place);
    // else the PBS: accumulator[acc_count] -> lwe_out
    NTT_NVTX("ntt_b200::host_blind_rotate / programmable_bootstrap");
    return guarded([&] {
        if (!batch) return NTT_B200_OK;
        const PrimePlan* pl = key->plan.get();
        DeviceGuard g(pl->device);
        keep_pool_cached(pl->device);
        const size_t n = pl->n, gs = key->glwe_size, per = gs * n, lwe_size = key->n_lwe + 1;
        const size_t out_size = (gs - 1) * n + 1;
        cudaStream_t st;
        NTT_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        int rc = NTT_B200_OK;
        try {
            Scratch sc(st);
            size_t chunk = std::min(batch, kChunkCts);
            uint64_t* d_lwe = sc.get<uint64_t>(chunk * lwe_size);
            uint64_t* d_acc = sc.get<uint64_t>(chunk * per);
            uint64_t* d_lut = nullptr;
            uint64_t* d_out = nullptr;
            bool shared_lut = !lut_inout && acc_count == 1;
            if (lut_inout) {
                d_lut = sc.get<uint64_t>(chunk * per);
            } else {
                d_out = sc.get<uint64_t>(chunk * out_size);
                if (shared_lut) {
                    d_lut = sc.get<uint64_t>(acc_count * per);
                    NTT_CUDA_CHECK(cudaMemcpyAsync(d_lut, accumulator, acc_count * per * 8,
                                                   cudaMemcpyHostToDevice, st));
                } else {
                    d_lut = sc.get<uint64_t>(chunk * per);
                }
            }
            for (size_t b0 = 0; b0 < batch; b0 += chunk) {
                size_t nb = std::min(chunk, batch - b0);
                NTT_CUDA_CHECK(cudaMemcpyAsync(d_lwe, lwe + b0 * lwe_size, nb * lwe_size * 8,
                                               cudaMemcpyHostToDevice, st));
                size_t lut_count = nb;
                if (lut_inout) {
                    NTT_CUDA_CHECK(cudaMemcpyAsync(d_lut, lut_inout + b0 * per, nb * per * 8,
                                                   cudaMemcpyHostToDevice, st));
                } else if (shared_lut) {
                    lut_count = 1;
                } else {
                    NTT_CUDA_CHECK(cudaMemcpyAsync(d_lut, accumulator + b0 * per, nb * per * 8,
                                                   cudaMemcpyHostToDevice, st));
                }
                blind_rotate_dev(key, d_lwe, d_lut, lut_count, d_acc, nb, bnf, width, raw_bnf_input, path, st);
                if (lut_inout) {
                    NTT_CUDA_CHECK(cudaMemcpyAsync(lut_inout + b0 * per, d_acc, nb * per * 8,
                                                   cudaMemcpyDeviceToHost, st));
                } else {
                    pbs_sample_extract_kernel<<<grid_for(nb * out_size), 256, 0, st>>>(
                        d_out, d_acc, nb, gs, (unsigned)pl->logn, bnf ? 0 : pl->p);
                    NTT_CUDA_CHECK(cudaGetLastError());
                    NTT_CUDA_CHECK(cudaMemcpyAsync(lwe_out + b0 * out_size, d_out, nb * out_size * 8,
                                                   cudaMemcpyDeviceToHost, st));
                }
                NTT_CUDA_CHECK(cudaStreamSynchronize(st));
            }
        } catch (...) {
            cudaStreamSynchronize(st);
            cudaStreamDestroy(st);
            throw;
        }
        NTT_CUDA_CHECK(cudaStreamSynchronize(st));
        cudaStreamDestroy(st);
        return rc;
    });
}

}  // namespace

extern "C" {

// NttLweBootstrapKey::from_container (entities/ntt_lwe_bootstrap_key.rs:68-110): `ntt_bsk` is the
// reference's flat container, already in the NTT domain.
int ntt_b200_bsk_new(const ntt_b200_plan64* plan, const uint64_t* ntt_bsk, size_t n_lwe, size_t glwe_size,
                     uint32_t base_log, uint32_t level, ntt_b200_bsk** out) {
    if (!out) return NTT_B200_ERR_ARG;
    *out = nullptr;
    if (!plan || !ntt_bsk || !n_lwe || glwe_size < 2 || !base_log || !level) return NTT_B200_ERR_ARG;
    if ((uint64_t)base_log * level >= 64) return NTT_B200_ERR_ARG;
    return guarded([&] {
        auto key = std::make_unique<ntt_b200_bsk>();
        key->plan = plan->impl;
        key->n_lwe = n_lwe;
        key->glwe_size = glwe_size;
        key->base_log = base_log;
        key->level = level;
        DeviceGuard g(key->plan->device);
        NTT_CUDA_CHECK(cudaMalloc(&key->d_bsk, key->total_len() * 8));
        NTT_CUDA_CHECK(cudaMemcpy(key->d_bsk, ntt_bsk, key->total_len() * 8, cudaMemcpyHostToDevice));
        make_twiddle_form(key.get());
        *out = key.release();
        return NTT_B200_OK;
    });
}

// convert_standard_lwe_bootstrap_key_to_ntt64 (lwe_bootstrap_key_conversion.rs:294-363) into a
// device-resident key: input_width = 0 when the standard key lives modulo the NTT prime, else the
// log2 of its power-of-two modulus; normalize != 0 is NttLweBootstrapKeyOption::Normalize.
int ntt_b200_bsk_convert_new(const ntt_b200_plan64* plan, const uint64_t* standard_bsk, size_t n_lwe,
                             size_t glwe_size, uint32_t base_log, uint32_t level, uint32_t input_width,
                             int normalize, ntt_b200_bsk** out) {
    if (!out) return NTT_B200_ERR_ARG;
    *out = nullptr;
    if (!plan || !standard_bsk || !n_lwe || glwe_size < 2 || !base_log || !level || input_width > 64)
        return NTT_B200_ERR_ARG;
    if ((uint64_t)base_log * level >= 64) return NTT_B200_ERR_ARG;
    return guarded([&] {
        auto key = std::make_unique<ntt_b200_bsk>();
        key->plan = plan->impl;
        key->n_lwe = n_lwe;
        key->glwe_size = glwe_size;
        key->base_log = base_log;
        key->level = level;
        const PrimePlan* pl = key->plan.get();
        DeviceGuard g(pl->device);
        size_t total = key->total_len();
        NTT_CUDA_CHECK(cudaMalloc(&key->d_bsk, total * 8));
        uint64_t* staging = nullptr;
        NTT_CUDA_CHECK(cudaMalloc(&staging, total * 8));
        cudaError_t e = cudaMemcpy(staging, standard_bsk, total * 8, cudaMemcpyHostToDevice);
        int rc = NTT_B200_OK;
        if (e == cudaSuccess) {
            rc = ntt_b200_ntt64_forward_device(plan, key->d_bsk, staging, total / pl->n, input_width ? 3 : 0,
                                               input_width, nullptr);
            if (rc == NTT_B200_OK && normalize) pl->normalize(key->d_bsk, total, nullptr);
            e = cudaDeviceSynchronize();
        }
        cudaFree(staging);
        NTT_CUDA_CHECK(e);
        if (rc != NTT_B200_OK) return rc;
        make_twiddle_form(key.get());
        *out = key.release();
        return NTT_B200_OK;
    });
}

// host-to-host form of the same conversion: `output` receives the NttLweBootstrapKey container
int ntt_b200_convert_standard_lwe_bootstrap_key_to_ntt64(const ntt_b200_plan64* plan, const uint64_t* input,
                                                         uint64_t* output, size_t len, uint32_t input_width,
                                                         int normalize) {
    if (!plan || (len && (!input || !output)) || input_width > 64) return NTT_B200_ERR_ARG;
    if (len % plan->impl->n) return NTT_B200_ERR_LEN;
    int rc = ntt_b200_ntt64_forward(plan, output, input, len, input_width ? 3 : 0, input_width);
    if (rc == NTT_B200_OK && normalize) rc = ntt_b200_plan64_normalize(plan, output, len);
    return rc;
}

void ntt_b200_bsk_free(ntt_b200_bsk* key) { delete key; }
size_t ntt_b200_bsk_input_lwe_dimension(const ntt_b200_bsk* key) { return key ? key->n_lwe : 0; }
size_t ntt_b200_bsk_glwe_size(const ntt_b200_bsk* key) { return key ? key->glwe_size : 0; }
size_t ntt_b200_bsk_polynomial_size(const ntt_b200_bsk* key) { return key ? key->plan->n : 0; }
uint32_t ntt_b200_bsk_decomposition_base_log(const ntt_b200_bsk* key) { return key ? key->base_log : 0; }
uint32_t ntt_b200_bsk_decomposition_level_count(const ntt_b200_bsk* key) { return key ? key->level : 0; }
const uint64_t* ntt_b200_bsk_device_data(const ntt_b200_bsk* key) { return key ? key->d_bsk : nullptr; }

int ntt_b200_bsk_read(const ntt_b200_bsk* key, uint64_t* out, size_t len) {
    if (!key || !out) return NTT_B200_ERR_ARG;
    if (len != key->total_len()) return NTT_B200_ERR_LEN;
    return guarded([&] {
        DeviceGuard g(key->plan->device);
        NTT_CUDA_CHECK(cudaMemcpy(out, key->d_bsk, len * 8, cudaMemcpyDeviceToHost));
        return NTT_B200_OK;
    });
}

// blind_rotate_ntt64_assign (ntt64_pbs.rs:175-286): lwe [batch][n_lwe+1], lut [batch][(k+1)N] in/out
int ntt_b200_blind_rotate_ntt64_assign(const ntt_b200_bsk* key, const uint64_t* lwe, uint64_t* lut,
                                       size_t batch, int path) {
    if (int e = check_key(key, 0, 0)) return e;
    if (batch && (!lwe || !lut)) return NTT_B200_ERR_ARG;
    return host_blind_rotate(key, lwe, lut, nullptr, 0, nullptr, batch, 0, 0, 0, path);
}
// blind_rotate_ntt64_bnf_assign (ntt64_bnf_pbs.rs:174-276): msed [batch][n_lwe+1] switched values
int ntt_b200_blind_rotate_ntt64_bnf_assign(const ntt_b200_bsk* key, uint32_t width, const uint64_t* msed,
                                           uint64_t* lut, size_t batch, int path) {
    if (int e = check_key(key, 1, width)) return e;
    if (batch && (!msed || !lut)) return NTT_B200_ERR_ARG;
    return host_blind_rotate(key, msed, lut, nullptr, 0, nullptr, batch, 1, width, 0, path);
}
// programmable_bootstrap_ntt64_lwe_ciphertext (ntt64_pbs.rs:439-538): accumulator is
// [acc_count][(k+1)N] with acc_count = 1 (one LUT for the batch) or batch
int ntt_b200_programmable_bootstrap_ntt64(const ntt_b200_bsk* key, const uint64_t* lwe_in, uint64_t* lwe_out,
                                          const uint64_t* accumulator, size_t acc_count, size_t batch,
                                          int path) {
    if (int e = check_key(key, 0, 0)) return e;
    if (batch && (!lwe_in || !lwe_out || !accumulator)) return NTT_B200_ERR_ARG;
    if (batch && acc_count != 1 && acc_count != batch) return NTT_B200_ERR_LEN;
    return host_blind_rotate(key, lwe_in, nullptr, accumulator, acc_count, lwe_out, batch, 0, 0, 0, path);
}
// programmable_bootstrap_ntt64_bnf_lwe_ciphertext (ntt64_bnf_pbs.rs:428-539)
int ntt_b200_programmable_bootstrap_ntt64_bnf(const ntt_b200_bsk* key, uint32_t width, const uint64_t* lwe_in,
                                              uint64_t* lwe_out, const uint64_t* accumulator, size_t acc_count,
                                              size_t batch, int path) {
    if (int e = check_key(key, 1, width)) return e;
    if (batch && (!lwe_in || !lwe_out || !accumulator)) return NTT_B200_ERR_ARG;
    if (batch && acc_count != 1 && acc_count != batch) return NTT_B200_ERR_LEN;
    return host_blind_rotate(key, lwe_in, nullptr, accumulator, acc_count, lwe_out, batch, 1, width, 1, path);
}

// device-resident forms: all pointers on the key's device, asynchronous on `stream`
int ntt_b200_blind_rotate_ntt64_device(const ntt_b200_bsk* key, int bnf, uint32_t width, const uint64_t* lwe,
                                       int lwe_is_switched, const uint64_t* lut, size_t lut_count,
                                       uint64_t* acc_out, size_t batch, int path, void* stream) {
    if (int e = check_key(key, bnf, width)) return e;
    if (!batch) return NTT_B200_OK;
    if (!lwe || !lut || !acc_out || !lut_count) return NTT_B200_ERR_ARG;
    if (!bnf && lwe_is_switched) return NTT_B200_ERR_ARG;
    return guarded([&] {
        DeviceGuard g(key->plan->device);
        blind_rotate_dev(key, lwe, lut, lut_count, acc_out, batch, bnf, width, bnf && !lwe_is_switched, path,
                         (cudaStream_t)stream);
        return NTT_B200_OK;
    });
}
int ntt_b200_extract_lwe_sample_device(const ntt_b200_bsk* key, int bnf, const uint64_t* glwe, uint64_t* lwe_out,
                                       size_t batch, void* stream) {
    if (!key) return NTT_B200_ERR_ARG;
    if (!batch) return NTT_B200_OK;
    if (!glwe || !lwe_out) return NTT_B200_ERR_ARG;
    return guarded([&] {
        DeviceGuard g(key->plan->device);
        size_t out_size = (key->glwe_size - 1) * key->plan->n + 1;
        pbs_sample_extract_kernel<<<grid_for(batch * out_size), 256, 0, (cudaStream_t)stream>>>(
            lwe_out, glwe, batch, key->glwe_size, (unsigned)key->plan->logn, bnf ? 0 : key->plan->p);
        NTT_CUDA_CHECK(cudaGetLastError());
        return NTT_B200_OK;
    });
}

}  // extern "C"
