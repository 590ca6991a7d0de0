// ntt_inst_lazy32_pk.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy32 modular policy, bit-packed global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(lazy32_pk, Lazy32, pk32)
}  // namespace fhe
