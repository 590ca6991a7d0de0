// ntt_inst_lazy64.cu -- instantiates the NTT / INTT / polymul kernels for the Lazy64 modular policy.
#include "ntt_kernels.cuh"

namespace fhe {
int ntt_launch_lazy64(int logn, int loge, int mode, const NttParams<Lazy64> &P, const u64 *a, const u64 *b, u64 *c,
                  u64 *c_evals, size_t batch, int flags, cudaStream_t st) {
    return launch_ntt<Lazy64>(logn, loge, mode, P, a, b, c, c_evals, batch, flags, st);
}
bool ntt_loge_ok_lazy64(int logn, int loge) { return ntt_loge_supported<Lazy64>(logn, loge); }
}  // namespace fhe
