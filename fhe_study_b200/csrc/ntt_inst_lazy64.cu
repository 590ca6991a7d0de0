// ntt_inst_lazy64.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy64 modular policy, u64 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(lazy64, Lazy64, u64)
}  // namespace fhe
