// torus.cuh -- shared declarations of the torus (q = 2^64) path: the two CRT primes, the CRT lift, and the
// device-resident key objects.
#pragma once
#include <algorithm>

#include "plan.cuh"

namespace fhe {

// Largest two primes below 2^27 with 2^16 | p-1 (every ring degree up to 2^15 has a 2N-th root; P = p1*p2 ~ 2^53.96).
// Below 2^27 a forward NTT of up to 15 stages needs no conditional subtraction in 32-bit words (extprod_fused.cu).
static constexpr u64 TORUS_P1 = 0x7E90001ull;
static constexpr u64 TORUS_P2 = 0x7E00001ull;

// centred CRT lift of (r1 mod p1, r2 mod p2) to the representative in (-P/2, P/2], as a wrapping u64
FHE_HD u64 crt_centered(u32 r1, u32 r2, u32 p1, u32 p2, u32 p1_inv_mod_p2, u64 P, u64 halfP, const Lazy32 &m2) {
    const u32 r1m = r1 >= p2 ? r1 - p2 : r1;           // p2 < p1 < 2*p2
    const u32 d = r2 >= r1m ? r2 - r1m : r2 + p2 - r1m;
    const u32 t = m2.mul(d, p1_inv_mod_p2);
    u64 v = (u64)r1 + (u64)p1 * t;                     // in [0, P)
    if (v > halfP) v -= P;                             // wraps to the two's-complement negative
    return v;
}

struct CrtParams {
    u32 p1, p2;
    u32 p1_inv_mod_p2;  // p1^-1 mod p2
    u64 P, halfP;
    int W;
    int shift[4];
    Lazy32 m2;
};

inline const void *plan_params32(const fhe_ntt_plan *p) { return &p->p32; }

// per-(device, n) context of the torus path: the two NTT plans and the CRT constants
struct TorusCtx {
    u64 n = 0;
    int logn = 0;
    fhe_ntt_plan *plan1 = nullptr, *plan2 = nullptr;
    Lazy32 m1, m2;
    CrtParams cp;
    int init(u64 n);
    ~TorusCtx();
    int ntt(int r, int mode, const u64 *in, u64 *out, size_t polys, cudaStream_t st) const;
};

// device-resident TGGSW (tfhe/src/tggsw.rs:14): the rows' 32-bit limb planes in the NTT domain of p1 and p2
struct Tggsw {
    TorusCtx *tc = nullptr;
    u64 k = 0;
    u64 *R1 = nullptr, *R2 = nullptr;  // [(k+1)*64 rows][k+1 comps][2 limbs][n]
    u32 *R1f = nullptr, *R2f = nullptr;  // fused-kernel layout (extprod_fused.cu), when the shape is instantiated
};

}  // namespace fhe

// C-ABI handle of a loaded TGGSW
struct fhe_tggsw {
    fhe::Tggsw g;
    u64 n = 0;
};

namespace fhe {

int tn_mul_device(const TorusCtx &tc, const u64 *a, const u64 *b, u64 *c, size_t batch, cudaStream_t st);
bool tn_mul_fused_supported(int logn);
int tn_mul_fused_device(const TorusCtx &tc, const u64 *a, const u64 *b, u64 *c, size_t batch, cudaStream_t st);
// out = g (x) ct1 (ct2 == nullptr)  or  ct1 + g (x) (ct2 - ct1)  (CMux)
int extprod_device(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, cudaStream_t st);
bool extprod_fused_supported(int logn, int k1);
int tggsw_build_fused_layout(Tggsw &g, cudaStream_t st);
int extprod_fused_device(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, int cmux, cudaStream_t st);
int tggsw_precompute(Tggsw &g, const u64 *rows_dev, cudaStream_t st);
int tglwe_encrypt_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *msgs, size_t batch, double sigma, u64 seed,
                         int uniform_mask, u64 *out, cudaStream_t st);
int tglwe_decrypt_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *ct, u64 *out, size_t batch, cudaStream_t st);
int tggsw_generate_device(const TorusCtx &tc, u64 k, const u64 *sk, const u64 *m, double sigma, u64 seed, int uniform_mask,
                          u64 *rows_out, cudaStream_t st);
int tn_addsub_device(const u64 *a, const u64 *b, u64 *c, size_t len, int op, cudaStream_t st);
int tn_left_rotate_device(const u64 *a, u64 *out, size_t polys, u32 n, const u64 *hs, u64 h_const, u32 group,
                          cudaStream_t st, size_t hs_stride = 1, int negacyclic = 0);
int cmux_chain_fused_device(const TorusCtx &tc, int k1, const u32 *const *keys_dev, const u64 *h_dev, int steps,
                            int negacyclic, const u64 *acc_in, u64 *acc_out, size_t batch, cudaStream_t st);
// acc_out[b] = cmux(gs[steps-1], .., cmux(gs[0], acc_in[b], X^{-h[b][0]} acc_in[b]) ..): fused persistent kernel when
// every TGGSW has the fused layout, otherwise one rotate + CMux launch pair per step.  h = [batch][steps].
int cmux_chain_device(const Tggsw *const *gs, size_t steps, const u64 *acc_in, const u64 *h, int negacyclic, u64 *acc_out,
                      size_t batch, cudaStream_t st);

}  // namespace fhe
