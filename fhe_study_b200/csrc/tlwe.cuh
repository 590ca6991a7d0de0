// tlwe.cuh -- declarations of the TLWE key-switch / bootstrapping path.
#pragma once
#include "common.cuh"

namespace fhe {

// device-resident key-switching key (tfhe/src/tlwe.rs:84-100): rows[(i*l + j)*(kn_out+1) + x]
struct Ksk {
    u64 kn_in = 0, kn_out = 0, l = 0;
    u64 *rows = nullptr;
    // tensor-core layout (ks_mma.cu), built when l == 64 and kn_in is even: byte planes of the key in
    // swizzled 32 KB blocks [n_tile][k_tile]
    unsigned char *mma_blocks = nullptr;
    u32 mma_n_tiles = 0;
    // when kn_out is a multiple of 32 the body column (index kn_out) would cost a 33rd, almost empty column tile
    // (1056 instead of 1024 CTAs at n=1024, batch 8192: an eighth wave on 148 SMs): the GEMM then covers the
    // kn_out mask columns only and the body column is a separate bit-vector x column product (ks_bcol_kernel)
    u64 *bcol = nullptr;  // [kn_in*l] = rows[r][kn_out], contiguous
};

int ksk_build_mma_layout(Ksk &k, cudaStream_t st);
int key_switch_tc_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st);
int key_switch_mma_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st);
int ksk_generate_device(u64 *rows, const u64 *sk, const u64 *new_sk, u64 seed, u32 kn_in, u32 kn_out, u32 l, double sigma,
                        int uniform_mask, cudaStream_t st);
int key_switch_bcol_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st);
int key_switch_device(const Ksk &k, const u64 *ct, u64 *out, size_t batch, cudaStream_t st);
int rotate_extract_device(const u64 *table, const u64 *ct, u64 *ext, u64 *acc_out, size_t batch, u32 n, u32 k, u32 c_kn,
                          cudaStream_t st);
int tlwe_encrypt_device(u64 *out, const u64 *sk, const u64 *msgs, size_t batch, u64 seed, u32 kn, double sigma, int uniform_mask,
                        cudaStream_t st);
int tlwe_decrypt_device(const u64 *sk, const u64 *ct, u64 *out, size_t batch, u32 kn, cudaStream_t st);
int chain_prepare_device(const u64 *table, const u64 *ct, u64 *acc0, u64 *hs, size_t batch, u32 n, u32 k, u32 c_kn,
                         u32 steps, int mode, cudaStream_t st);
int sample_extract_device(const u64 *ct, u64 *out, size_t batch, u32 n, u32 k, u32 h, cudaStream_t st);
int shift_right_device(const u64 *a, u64 *out, size_t len, u32 shift, cudaStream_t st);

}  // namespace fhe
