// plan.cuh -- the NTT plan object behind fhe_ntt_plan (shared by the translation units of the library).
#pragma once
#include "ntt_kernels.cuh"
#include "plan_host.hpp"

namespace fhe {
int ntt_launch_lazy32(int, int, int, const NttParams<Lazy32> &, const u64 *, const u64 *, u64 *, u64 *, size_t, int, cudaStream_t);
int ntt_launch_lazy64(int, int, int, const NttParams<Lazy64> &, const u64 *, const u64 *, u64 *, u64 *, size_t, int, cudaStream_t);
int ntt_launch_strict64(int, int, int, const NttParams<Strict64> &, const u64 *, const u64 *, u64 *, u64 *, size_t, int, cudaStream_t);
int ntt_launch_small32(int, int, int, const NttParams<Small32> &, const u64 *, const u64 *, u64 *, u64 *, size_t, int, cudaStream_t);
bool ntt_loge_ok_lazy32(int, int);
bool ntt_loge_ok_lazy64(int, int);
bool ntt_loge_ok_strict64(int, int);
bool ntt_loge_ok_small32(int, int);
}  // namespace fhe

// One plan per (device, q, n); owned by the cache in lib_core.cu, reference-counted by create/destroy.
struct fhe_ntt_plan {
    int device = 0;
    int kind = 0;  // fhe::modulus_kind(q)
    int logn = 0;
    int loge = 0;  // log2(coefficients per thread) the device tables were laid out for
    int refs = 0;
    fhe::HostTables host;
    void *d_fwd = nullptr, *d_inv = nullptr;
    fhe::NttParams<fhe::Lazy32> p32;
    fhe::NttParams<fhe::Lazy64> p64;
    fhe::NttParams<fhe::Strict64> ps64;
    fhe::NttParams<fhe::Small32> psm;
};
