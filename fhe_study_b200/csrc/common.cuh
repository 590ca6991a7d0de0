// common.cuh -- shared host/device helpers for the fhe_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "modarith.cuh"

#include <string.h>

#include <mutex>
#include <string>

typedef unsigned __int128 u128;

namespace fhe {

// ---- error plumbing ---------------------------------------------------------------------------
void set_error(const std::string &msg);
#define FHE_CUDA_OK(expr)                                                                           \
    do {                                                                                            \
        cudaError_t _e = (expr);                                                                    \
        if (_e != cudaSuccess) {                                                                    \
            ::fhe::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                   \
            return -2;                                                                              \
        }                                                                                           \
    } while (0)
#define FHE_REQUIRE(cond, msg)                                                                      \
    do {                                                                                            \
        if (!(cond)) {                                                                              \
            ::fhe::set_error(std::string(msg));                                                     \
            return -1;                                                                              \
        }                                                                                           \
    } while (0)

cudaStream_t current_stream();
int num_sms();
void device_init_once();  // per-device runtime configuration (memory pool)
void count_launch(unsigned long long n);  // feeds fhe_launch_count()

// ---- host modular helpers (plan construction) ---------------------------------------------------
inline u64 h_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }
inline u64 h_powmod(u64 x, u64 k, u64 q) {
    u64 r = 1;
    x %= q;
    while (k) {
        if (k & 1) r = h_mulmod(r, x, q);
        x = h_mulmod(x, x, q);
        k >>= 1;
    }
    return r;
}
inline u64 h_invmod(u64 x, u64 q) { return h_powmod(x, q - 2, q); }
inline bool h_is_prime(u64 n) {  // deterministic Miller-Rabin for 64-bit
    if (n < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = h_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < s; i++) {
            x = h_mulmod(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}
inline u32 h_bitrev(u32 i, int bits) {
    u32 r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}
inline int h_ilog2(u64 n) {
    int l = 0;
    while (n > 1) { n >>= 1; l++; }
    return l;
}

}  // namespace fhe
