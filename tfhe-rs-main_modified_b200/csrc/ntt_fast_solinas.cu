// Fast-kernel instantiations for one modulus family (see ntt_fast_impl.cuh).
#include "ntt_fast_impl.cuh"
namespace nttb200 {
using S64H = Shoup<uint64_t, true>;
using S64L = Shoup<uint64_t, false>;
using S32H = Shoup<uint32_t, true>;
using S32L = Shoup<uint32_t, false>;
NTT_DEFINE_FAST(Solinas64)
}  // namespace nttb200
