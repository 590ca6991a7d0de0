// ntt_inst_lazy64_u32.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy64 modular policy, u32 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(lazy64_u32, Lazy64, u32)
}  // namespace fhe
