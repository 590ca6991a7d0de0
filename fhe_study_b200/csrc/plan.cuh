// plan.cuh -- the NTT plan object behind fhe_ntt_plan (shared by the translation units of the library).
#pragma once
#include "ntt_kernels.cuh"
#include "plan_host.hpp"

namespace fhe {
#define FHE_NTT_DECLARE(NAME, POLICY, IOW)                                                                           \
    int ntt_launch_##NAME(int, int, int, const NttParams<POLICY> &, const IOW *, const IOW *, IOW *, IOW *, size_t, int, \
                          cudaStream_t);                                                                             \
    bool ntt_loge_ok_##NAME(int, int);
FHE_NTT_DECLARE(lazy32, Lazy32, u64)
FHE_NTT_DECLARE(lazy64, Lazy64, u64)
FHE_NTT_DECLARE(strict64, Strict64, u64)
FHE_NTT_DECLARE(small32, Small32, u64)
// packed 32-bit words in global memory (q <= 2^32: Small32 / Lazy32, and Lazy64 for 2^30 <= q <= 2^32)
FHE_NTT_DECLARE(lazy32_u32, Lazy32, u32)
FHE_NTT_DECLARE(lazy64_u32, Lazy64, u32)
FHE_NTT_DECLARE(small32_u32, Small32, u32)
// bit-packed words (q < 2^30, n >= 1024): the PCIe wire of the host-buffer path
FHE_NTT_DECLARE(small32_pk, Small32, pk32)
FHE_NTT_DECLARE(lazy32_pk, Lazy32, pk32)
// q = 65537: radix-4 butterflies (modarith.cuh: Fermat32), all three word formats
FHE_NTT_DECLARE(fermat32, Fermat32, u64)
FHE_NTT_DECLARE(fermat32_u32, Fermat32, u32)
FHE_NTT_DECLARE(fermat32_pk, Fermat32, pk32)
#undef FHE_NTT_DECLARE
}  // namespace fhe

// One plan per (device, q, n); owned by the cache in lib_core.cu, reference-counted by create/destroy.
struct fhe_ntt_plan {
    int device = 0;
    int kind = 0;  // fhe::modulus_kind(q)
    int logn = 0;
    int loge = 0;  // log2(coefficients per thread) the device tables were laid out for
    int refs = 0;
    int gpark = 0;   // polymul with NTT(a) parked in the output row (MODE_MULG; 32-bit words, degrees >= 2^13)
    int staged = 0;  // polymul through the persistent staged kernel (MODE_MULS; degrees >= 2^13)
    int fermat = 0;  // q = 65537 (kind 3; kind 0 at n = 2^15): the transforms run under the Fermat32 policy (pfm, radix-4 tables d_fwd4 / d_inv4);
                     // psm and d_fwd / d_inv stay valid for the fused consumers that bring their own Small32 code
    int dual = 0;  // polymul of two coefficient-form operands through the dual-operand kernel (MODE_MUL2)
    fhe::HostTables host;
    void *d_fwd = nullptr, *d_inv = nullptr, *d_fwd4 = nullptr, *d_inv4 = nullptr, *d_fwd4w = nullptr, *d_inv4w = nullptr;
    fhe::NttParams<fhe::Lazy32> p32;
    fhe::NttParams<fhe::Lazy64> p64;
    fhe::NttParams<fhe::Strict64> ps64;
    fhe::NttParams<fhe::Small32> psm;
    fhe::NttParams<fhe::Fermat32> pfm;
};
