// ntt_inst_lazy32.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy32 modular policy.
#include "ntt_kernels.cuh"

namespace fhe {
int ntt_launch_lazy32(int logn, int loge, int mode, const NttParams<Lazy32> &P, const u64 *a, const u64 *b, u64 *c,
                  u64 *c_evals, size_t batch, int flags, cudaStream_t st) {
    return launch_ntt<Lazy32>(logn, loge, mode, P, a, b, c, c_evals, batch, flags, st);
}
bool ntt_loge_ok_lazy32(int logn, int loge) { return ntt_loge_supported<Lazy32>(logn, loge); }
}  // namespace fhe
