// ntt_inst_strict64.cu -- instantiates the NTT / INTT / polymul kernels for the Strict64 modular policy, u64 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(strict64, Strict64, u64)
}  // namespace fhe
