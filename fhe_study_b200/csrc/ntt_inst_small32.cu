// ntt_inst_small32.cu -- instantiates the NTT / INTT / polymul kernels for the Small32 modular policy, u64 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(small32, Small32, u64)
}  // namespace fhe
