// ntt_inst_lazy32_u32.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy32 modular policy, u32 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(lazy32_u32, Lazy32, u32)
}  // namespace fhe
