// ntt_inst_small32_pk.cu -- instantiates the NTT / INTT / polymul kernels for the Small32 modular policy, bit-packed global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(small32_pk, Small32, pk32)
}  // namespace fhe
