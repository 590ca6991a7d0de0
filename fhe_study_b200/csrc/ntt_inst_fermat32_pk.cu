// ntt_inst_fermat32_pk.cu -- instantiates the NTT / INTT / polymul kernels for the Fermat32 modular policy (q = 65537, radix-4 butterflies), bit-packed global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(fermat32_pk, Fermat32, pk32)
}  // namespace fhe
