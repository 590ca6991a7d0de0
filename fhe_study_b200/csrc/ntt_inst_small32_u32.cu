// ntt_inst_small32_u32.cu -- instantiates the NTT / INTT / polymul kernels for the Small32 modular policy, u32 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(small32_u32, Small32, u32)
}  // namespace fhe
