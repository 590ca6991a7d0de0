// ntt_inst_fermat32.cu -- instantiates the NTT / INTT / polymul kernels for the Fermat32 modular policy (q = 65537, radix-4 butterflies), u64 global words.
#include "ntt_kernels.cuh"

namespace fhe {
FHE_NTT_INSTANTIATE(fermat32, Fermat32, u64)
}  // namespace fhe
