// ntt_inst_lazy32.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy32 modular policy, u64 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(lazy32, Lazy32, u64)
}  // namespace fhe
