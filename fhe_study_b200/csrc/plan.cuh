// plan.cuh -- the NTT plan object behind fhe_ntt_plan (shared by the translation units of the library).
#pragma once
#include "ntt_kernels.cuh"
#include "plan_host.hpp"

// One plan per (device, q, n); owned by the cache in lib_core.cu, reference-counted by create/destroy.
struct fhe_ntt_plan {
    int device = 0;
    int kind = 0;  // fhe::modulus_kind(q)
    int logn = 0;
    int refs = 0;
    fhe::HostTables host;
    void *d_fwd = nullptr, *d_inv = nullptr;
    fhe::NttParams<fhe::Lazy32> p32;
    fhe::NttParams<fhe::Lazy64> p64;
    fhe::NttParams<fhe::Strict64> ps64;
};
