// bfv_kernels.cu -- BFV ciphertext x ciphertext multiplication with relinearisation, exactly as the reference
// computes it (bfv/src/lib.rs:59-90,251-271): NOT an NTT.  Each product is ring_n::naive_mul, the
// exact LINEAR convolution (2n-1 outputs) in i128 truncated to i64 (arith/src/ring_n.rs:307-320) -- i.e. the
// low 64 bits of the sum, which 64-bit wrapping multiply-adds reproduce bit for bit -- followed, per UNFOLDED
// coefficient, by f64 round((num*v)/den) (ring_n.rs:130-138), Zq::from_f64 (zq.rs:32-40), and only then the
// X^n+1 fold as a Zq subtraction (ring_nq.rs:132-141).  The f64 steps use explicit IEEE round-to-nearest
// intrinsics (no FMA contraction) so they match the CPU's arithmetic.
#include <algorithm>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "scheme_common.cuh"

namespace fhe {

// fold of the scaled pair (ring_nq.rs:132-141): res[c] = f(conv[c]) - f(conv[c+n]); index c+n exists for c <= n-2
__device__ __forceinline__ u64 scale_fold(u64 q, u64 mu, u32 n, u32 c, u64 lo, u64 hi, double num, double den) {
    const u64 x = scale_round(q, mu, (i64)lo, num, den);
    if (c + 1 >= n) return x;
    return zq_sub(q, x, scale_round(q, mu, (i64)hi, num, den));
}

// mode 0: RLWE::tensor only (out = c0|c1|c2, 3n words) ; 1: RLWE::mul (tensor + relinearize_204, out 2n words) ;
// 2: relinearize_204 only (a = c0|c1|c2, out 2n words).  n threads per ciphertext product, thread c owns the
// unfolded coefficients c and c+n of every product: one uniform loop over i with j = (c - i) mod n, the term
// going to the low sum when i <= c and to the high sum otherwise (wrapping 64-bit == `i128 as i64`).
// CT = u32 when q < 2^32 (ciphertext coefficients are < q): the products are then single IMAD.WIDE.U32.
template <typename CT>
__global__ void bfv_mul_kernel(const u64 *__restrict__ a, const u64 *__restrict__ b, const u64 *__restrict__ rlk,
                               u64 *__restrict__ out, size_t batch, u32 n, u64 q, u64 t, u64 pq, int mode, u32 per_cta) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const u32 slot = threadIdx.x / n, c = threadIdx.x % n;
    const size_t ct = (size_t)blockIdx.x * per_cta + slot;
    const bool valid = ct < batch && slot < per_cta;
    // per CTA: rlk0 rlk1 (u64, shared by every slot) ; per slot: a0 a1 b0 b1 c2 (CT)
    u64 *k0 = reinterpret_cast<u64 *>(sm_raw), *k1 = k0 + n;
    CT *s = reinterpret_cast<CT *>(k1 + n) + (size_t)slot * 5 * n;
    CT *a0 = s, *a1 = s + n, *b0 = s + 2 * n, *b1 = s + 3 * n, *c2s = s + 4 * n;
    const double dq = __ull2double_rn(q), dt = __ull2double_rn(t);
    const u64 mu = ~0ull / q;  // floor((2^64 - 1) / q)
    u64 c0 = 0, c1 = 0, c2 = 0;
    if (mode != 0) {
        for (u32 i = threadIdx.x; i < 2 * n; i += blockDim.x) k0[i] = rlk[i];
    }
    if (valid) {
        if (mode != 2) {
            const u64 *pa = a + ct * 2 * n, *pb = b + ct * 2 * n;
            a0[c] = (CT)pa[c]; a1[c] = (CT)pa[n + c]; b0[c] = (CT)pb[c]; b1[c] = (CT)pb[n + c];
        } else {
            const u64 *pa = a + ct * 3 * n;
            c0 = pa[c]; c1 = pa[n + c]; c2 = pa[2 * n + c];
        }
    }
    __syncthreads();
    if (valid && mode != 2) {  // RLWE::tensor (lib.rs:59-85)
        u64 l00 = 0, h00 = 0, l01 = 0, h01 = 0, l11 = 0, h11 = 0;
        u32 j = c;  // j = (c - i) mod n
#pragma unroll 4
        for (u32 i = 0; i < n; i++) {
            const u64 x0 = a0[i], x1 = a1[i], y0 = b0[j], y1 = b1[j];
            if (i <= c) {
                l00 += x0 * y0; l01 += x0 * y1 + x1 * y0; l11 += x1 * y1;  // i64 `+` of the cross terms (lib.rs:76)
            } else {
                h00 += x0 * y0; h01 += x0 * y1 + x1 * y0; h11 += x1 * y1;
            }
            j = j == 0 ? n - 1 : j - 1;
        }
        c0 = scale_fold(q, mu, n, c, l00, h00, dt, dq);
        c1 = scale_fold(q, mu, n, c, l01, h01, dt, dq);
        c2 = scale_fold(q, mu, n, c, l11, h11, dt, dq);
    }
    if (mode == 0) {
        if (valid) {
            u64 *po = out + ct * 3 * n;
            po[c] = c0; po[n + c] = c1; po[2 * n + c] = c2;
        }
        return;
    }
    if (valid) c2s[c] = (CT)c2;
    __syncthreads();
    if (valid) {  // relinearize_204 (lib.rs:251-271)
        const double dp = __ull2double_rn(pq / q);
        u64 l0 = 0, h0 = 0, l1 = 0, h1 = 0;
        u32 j = c;
#pragma unroll 4
        for (u32 i = 0; i < n; i++) {
            const u64 x = c2s[i], y0 = k0[j], y1 = k1[j];
            if (i <= c) { l0 += x * y0; l1 += x * y1; } else { h0 += x * y0; h1 += x * y1; }
            j = j == 0 ? n - 1 : j - 1;
        }
        const u64 r0 = scale_fold(q, mu, n, c, l0, h0, 1.0, dp);
        const u64 r1 = scale_fold(q, mu, n, c, l1, h1, 1.0, dp);
        u64 *po = out + ct * 2 * n;
        po[c] = zq_add(q, c0, r0);
        po[n + c] = zq_add(q, c1, r1);
    }
}

static int bfv_launch(int mode, u64 q, u64 n, u64 t, u64 pq, const u64 *rlk, const u64 *a, const u64 *b, u64 *out,
                      size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(n >= 1 && n <= 1024, "bfv: n must be in 1..1024 (one thread per coefficient)");
    FHE_REQUIRE(q >= 2 && q < (1ull << 63), "bfv: q must be < 2^63");
    FHE_REQUIRE(a && out && (mode == 2 || b) && (mode == 0 || rlk), "bfv: null pointer");
    FHE_REQUIRE(mode == 0 || pq / q >= 1, "bfv: rlk modulus must be a multiple p*q with p >= 1");
    cudaStream_t st = current_stream();
    const size_t in_words = (mode == 2 ? 3 : 2) * n, out_words = (mode == 0 ? 3 : 2) * n;
    IoBuf ba, bb, bk, bo;
    int rc;
    if ((rc = ba.init(a, batch * in_words * 8, true, false, st))) return rc;
    if ((rc = bb.init(mode == 2 ? nullptr : b, batch * in_words * 8, true, false, st))) return rc;
    if ((rc = bk.init(mode == 0 ? nullptr : rlk, 2 * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * out_words * 8, false, true, st))) return rc;
    const u32 per_cta = n >= 128 ? 1 : (u32)(128 / n);
    const u32 threads = per_cta * (u32)n;
    const bool narrow = q <= 0xffffffffull;  // ciphertext coefficients (< q) fit 32 bits
    const size_t smem = 2 * n * sizeof(u64) + (size_t)per_cta * 5 * n * (narrow ? sizeof(u32) : sizeof(u64));
    const size_t grid = (batch + per_cta - 1) / per_cta;
    FHE_REQUIRE(grid <= 0x7fffffffull, "bfv: batch too large");
    auto k32 = bfv_mul_kernel<u32>;
    auto k64 = bfv_mul_kernel<u64>;
    if (smem > 48 * 1024)
        FHE_CUDA_OK(cudaFuncSetAttribute(narrow ? k32 : k64, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (narrow)
        k32<<<(unsigned)grid, threads, smem, st>>>(ba.ptr<u64>(), bb.ptr<u64>(), bk.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, q, t,
                                                 pq, mode, per_cta);
    else
        k64<<<(unsigned)grid, threads, smem, st>>>(ba.ptr<u64>(), bb.ptr<u64>(), bk.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, q, t,
                                                 pq, mode, per_cta);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&ba, &bb, &bk, &bo}, st);
}

// ---- BFV::decrypt (bfv/src/lib.rs:164-178): m = ((c0 + c1*s) . mul_div_round(t, q)) . remodule(t) -----------------
// c1 polynomials of `batch` RLWEs gathered contiguously (the transform kernels take dense batches)
__global__ void bfv_gather_c1_kernel(const u64 *__restrict__ ct, u64 *__restrict__ c1, size_t batch, u32 n) {
    const size_t total = batch * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        c1[i] = ct[(i / n) * 2 * n + n + i % n];
}
// cs = c0 + c1s (Zq::add, zq.rs:219-231); r = Zq::from_f64(round((t * cs) / q)) (ring_nq.rs:106-113); m = from_u64(t, r)
__global__ void bfv_decrypt_finish_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ c1s, u64 *__restrict__ m,
                                          size_t batch, u32 n, u64 q, u64 t) {
    const size_t total = batch * n;
    const double dq = __ull2double_rn(q), dt = __ull2double_rn(t);
    const u64 mu = ~0ull / q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const u64 cs = zq_add(q, ct[(i / n) * 2 * n + i % n], c1s[i]);
        const u64 r = zq_from_f64(q, mu, __ddiv_rn(__dmul_rn(dt, __ull2double_rn(cs)), dq));
        m[i] = r >= t ? r % t : r;
    }
}

// ---- BFV::encrypt (bfv/src/lib.rs:142-160) with a counter-based sampler (the CPU restatement orc_bfv_encrypt_ctr) -------
// draw p of ciphertext r is SplitMix64 output r*25n + p + 1: p < n -> u_x = from_f64(-1 + 2 unit) (Uniform(-1,1), lib.rs:149);
// n + 12x + t -> e1_x, 13n + 12x + t -> e2_x, each from_f64(sigma * (sum of 12 units - 6)) (Normal(0, sigma) stand-in)
__global__ void bfv_sample_u_kernel(u64 *__restrict__ U, size_t batch, u32 n, u64 q, u64 seed) {
    const size_t total = batch * n;
    const u64 mu = ~0ull / q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n;
        const u32 x = (u32)(i % n);
        U[i] = zq_from_f64(q, mu, __dadd_rn(-1.0, __dmul_rn(2.0, bfv_unit(bfv_draw(seed, r * 25 * (size_t)n + x)))));
    }
}
__global__ void bfv_encrypt_finish_kernel(const u64 *__restrict__ P0, const u64 *__restrict__ P1, const u64 *__restrict__ m,
                                          u64 *__restrict__ ct, size_t batch, u32 n, u64 q, u64 t, double sigma, u64 seed) {
    const size_t total = batch * n;
    const u64 mu = ~0ull / q, delta = q / t, dq = delta >= q ? delta % q : delta;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n, base = r * 25 * (size_t)n;
        const u32 x = (u32)(i % n);
        double a1 = 0.0, a2 = 0.0;
        for (u32 k = 0; k < 12; k++) {
            a1 = __dadd_rn(a1, bfv_unit(bfv_draw(seed, base + n + 12 * (size_t)x + k)));
            a2 = __dadd_rn(a2, bfv_unit(bfv_draw(seed, base + 13 * (size_t)n + 12 * (size_t)x + k)));
        }
        const u64 e1 = zq_from_f64(q, mu, __dmul_rn(sigma, __dadd_rn(a1, -6.0)));
        const u64 e2 = zq_from_f64(q, mu, __dmul_rn(sigma, __dadd_rn(a2, -6.0)));
        const u64 mv = m[i] >= q ? m[i] % q : m[i];                                     // m.remodule(q), ring_nq.rs:82-88
        const u64 md = (u64)(((unsigned __int128)mv * dq) % q);                         // m * floor(q/t), ring_nq.rs:274-281
        ct[r * 2 * n + x] = zq_add(q, zq_add(q, P0[i], e1), md);                        // &pk.0 * &u + e_1 + m*delta
        ct[r * 2 * n + n + x] = zq_add(q, P1[i], e2);                                   // &pk.1 * &u + e_2
    }
}

// ---- BFV::new_key (bfv/src/lib.rs:120-140), counter-based sampler (orc_bfv_keygen_ctr) ---------------------------------
// draws: p < n: s_p = draw % 2 ; n + x: a_x = draw % q ; 2n + 12x + t: e_x.  Writes s, -a (for the product), a and e.
__global__ void bfv_keygen_sample_kernel(u64 *__restrict__ sk, u64 *__restrict__ neg_a, u64 *__restrict__ a, u64 *__restrict__ e,
                                         u32 n, u64 q, double sigma, u64 seed) {
    const u64 mu = ~0ull / q;
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
        const u64 s = bfv_draw(seed, x) % 2, av = bfv_draw(seed, (u64)n + x) % q;
        sk[x] = s >= q ? s % q : s;  // Zq::from_u64 (zq.rs:21-31)
        a[x] = av;
        neg_a[x] = av == 0 ? 0 : q - av;  // Neg (zq.rs:302-314)
        e[x] = zq_from_f64(q, mu, ctr_gauss(seed, 2 * (u64)n + 12 * (u64)x, sigma));
    }
}
__global__ void rq_add_inplace_kernel(u64 *__restrict__ acc, const u64 *__restrict__ b, u32 n, u64 q) {
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) acc[x] = zq_add(q, acc[x], b[x]);
}

// ---- BFV::rlk_key (bfv/src/lib.rs:202-225) in the ring mod pq, counter-based sampler (orc_bfv_rlk_key_ctr) ------------------
// One CTA, thread c owns coefficient c: a*s and s*s through tmp_naive_mul (lib.rs:93-98) = ring_n::naive_mul (wrapping
// 64-bit sums == `i128 as i64`), Rq::from_vec_i64 (ring_nq.rs:164-170: Zq::from_f64(c as f64) per UNFOLDED coefficient),
// then the X^n+1 fold as a Zq subtraction; rlk.0 = -(a*s + e) + (s*s)*p, rlk.1 = a.
__global__ void bfv_rlk_kernel(const u64 *__restrict__ sk, u64 *__restrict__ rlk, u32 n, u64 q, u64 p, double sigma, u64 seed) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    u64 *s = reinterpret_cast<u64 *>(sm_raw), *a = s + n;
    const u64 pq = p * q, mu = ~0ull / pq;
    const u32 c = threadIdx.x;
    s[c] = sk[c] >= pq ? sk[c] % pq : sk[c];  // s.0.remodule(pq)
    a[c] = bfv_draw(seed, c) % pq;
    __syncthreads();
    u64 las = 0, has = 0, lss = 0, hss = 0;
    u32 j = c;
    for (u32 i = 0; i < n; i++) {
        const u64 as = a[i] * s[j], ss = s[i] * s[j];
        if (i <= c) { las += as; lss += ss; } else { has += as; hss += ss; }
        j = j == 0 ? n - 1 : j - 1;
    }
    auto from_i64 = [&](u64 v) { return zq_from_f64(pq, mu, __ll2double_rn((i64)v)); };
    u64 as = from_i64(las), ss = from_i64(lss);
    if (c + 1 < n) {
        as = zq_sub(pq, as, from_i64(has));
        ss = zq_sub(pq, ss, from_i64(hss));
    }
    const u64 e = zq_from_f64(pq, mu, ctr_gauss(seed, (u64)n + 12 * (u64)c, sigma));
    const u64 t0 = zq_add(pq, as, e);
    const u64 neg = t0 == 0 ? 0 : pq - t0;
    const u64 pm = p >= pq ? p % pq : p;                                              // Zq::from_u64(pq, p), ring_nq.rs:274-281
    const u64 ssp = (u64)(((unsigned __int128)ss * pm) % pq);
    rlk[c] = zq_add(pq, neg, ssp);
    rlk[n + c] = a[c];
}

// md rows of BFV::mul_const (lib.rs:189-200): (m.remodule(q) * floor(q/t), 0)
__global__ void bfv_mul_const_md_kernel(const u64 *__restrict__ m, u64 *__restrict__ md, size_t batch, u32 n, u64 q, u64 t) {
    const size_t total = batch * 2 * n;
    const u64 delta = q / t, dq = delta >= q ? delta % q : delta;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (2 * n);
        const u32 x = (u32)(i % (2 * n));
        u64 v = 0;
        if (x < n) {
            const u64 mv = m[r * n + x] >= q ? m[r * n + x] % q : m[r * n + x];
            v = (u64)(((unsigned __int128)mv * dq) % q);
        }
        md[i] = v;
    }
}

}  // namespace fhe

struct fhe_ntt_plan;
namespace fhe {
int plan_launch(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch, int flags,
                cudaStream_t st);  // lib_core.cu
}

using namespace fhe;
extern "C" {
int fhe_bfv_encrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, uint64_t t, const uint64_t *pk, const uint64_t *m,
                    double sigma, uint64_t seed, uint64_t *ct, size_t batch) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(pk && m && ct, "fhe_bfv_encrypt: null pointer");
    FHE_REQUIRE(t >= 1 && q >= 2 && n >= 1, "fhe_bfv_encrypt: need t >= 1");
    cudaStream_t st = current_stream();
    IoBuf bp, bm, bc;
    Scratch U, P0, P1;
    int rc;
    if ((rc = bp.init(pk, 2 * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(m, batch * n * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * 2 * n * 8, false, true, st))) return rc;
    if ((rc = U.alloc(batch * n * 8, st)) || (rc = P0.alloc(batch * n * 8, st)) || (rc = P1.alloc(batch * n * 8, st))) return rc;
    const unsigned grid = (unsigned)std::min<size_t>((batch * n + 255) / 256, (size_t)num_sms() * 16);
    bfv_sample_u_kernel<<<grid, 256, 0, st>>>(U.ptr<u64>(), batch, (u32)n, q, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    // pk.0 * u and pk.1 * u: one public-key polynomial against every u (Rq products are commutative: exact in Z_q)
    if ((rc = plan_launch(plan, 2, U.ptr<u64>(), bp.ptr<u64>(), P0.ptr<u64>(), nullptr, batch, 4, st))) return rc;
    if ((rc = plan_launch(plan, 2, U.ptr<u64>(), bp.ptr<u64>() + n, P1.ptr<u64>(), nullptr, batch, 4, st))) return rc;
    bfv_encrypt_finish_kernel<<<grid, 256, 0, st>>>(P0.ptr<u64>(), P1.ptr<u64>(), bm.ptr<u64>(), bc.ptr<u64>(), batch, (u32)n, q, t,
                                                  sigma, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bp, &bm, &bc}, st);
}
int fhe_bfv_decrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, uint64_t t, const uint64_t *sk, const uint64_t *ct,
                    uint64_t *m, size_t batch) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && ct && m, "fhe_bfv_decrypt: null pointer");
    FHE_REQUIRE(t >= 1 && q >= 2 && n >= 1, "fhe_bfv_decrypt: need t >= 1");
    cudaStream_t st = current_stream();
    IoBuf bs, bc, bm;
    Scratch c1, c1s;
    int rc;
    if ((rc = bs.init(sk, n * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * 2 * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(m, batch * n * 8, false, true, st))) return rc;
    if ((rc = c1.alloc(batch * n * 8, st))) return rc;
    if ((rc = c1s.alloc(batch * n * 8, st))) return rc;
    const unsigned grid = (unsigned)std::min<size_t>((batch * n + 255) / 256, (size_t)num_sms() * 16);
    bfv_gather_c1_kernel<<<grid, 256, 0, st>>>(bc.ptr<u64>(), c1.ptr<u64>(), batch, (u32)n);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    // c1 * s: every product shares the one secret-key polynomial (its transform is recomputed per CTA, n <= 2^15 words)
    if ((rc = plan_launch(plan, 2 /* MODE_MUL */, c1.ptr<u64>(), bs.ptr<u64>(), c1s.ptr<u64>(), nullptr, batch, 4 /* B_BROADCAST */, st)))
        return rc;
    bfv_decrypt_finish_kernel<<<grid, 256, 0, st>>>(bc.ptr<u64>(), c1s.ptr<u64>(), bm.ptr<u64>(), batch, (u32)n, q, t);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bs, &bc, &bm}, st);
}
int fhe_bfv_keygen(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, double sigma, uint64_t seed, uint64_t *sk, uint64_t *pk) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    FHE_REQUIRE(sk && pk, "fhe_bfv_keygen: null pointer");
    FHE_REQUIRE(q >= 2 && n >= 1, "fhe_bfv_keygen: bad parameters");
    cudaStream_t st = current_stream();
    IoBuf bs, bp;
    Scratch tmp;  // -a | e
    int rc;
    if ((rc = bs.init(sk, n * 8, false, true, st))) return rc;
    if ((rc = bp.init(pk, 2 * n * 8, false, true, st))) return rc;
    if ((rc = tmp.alloc(2 * n * 8, st))) return rc;
    const unsigned grid = (unsigned)((n + 255) / 256);
    bfv_keygen_sample_kernel<<<grid, 256, 0, st>>>(bs.ptr<u64>(), tmp.ptr<u64>(), bp.ptr<u64>() + n, tmp.ptr<u64>() + n, (u32)n, q,
                                                 sigma, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = plan_launch(plan, 2, tmp.ptr<u64>(), bs.ptr<u64>(), bp.ptr<u64>(), nullptr, 1, 0, st))) return rc;  // &(-a) * &s
    rq_add_inplace_kernel<<<grid, 256, 0, st>>>(bp.ptr<u64>(), tmp.ptr<u64>() + n, (u32)n, q);                     // + e
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bs, &bp}, st);
}
int fhe_bfv_rlk_generate(uint64_t q, uint64_t n, uint64_t p, double sigma, uint64_t seed, const uint64_t *sk, uint64_t *rlk) {
    FHE_REQUIRE(sk && rlk, "fhe_bfv_rlk_generate: null pointer");
    FHE_REQUIRE(n >= 1 && n <= 1024, "fhe_bfv_rlk_generate: n must be in 1..1024 (one thread per coefficient)");
    FHE_REQUIRE(q >= 2 && p >= 1 && (unsigned __int128)p * q < ((unsigned __int128)1 << 63),
                "fhe_bfv_rlk_generate: p*q must be < 2^63 (Zq arithmetic mod p*q, zq.rs:225)");
    cudaStream_t st = current_stream();
    IoBuf bs, bk;
    int rc;
    if ((rc = bs.init(sk, n * 8, true, false, st))) return rc;
    if ((rc = bk.init(rlk, 2 * n * 8, false, true, st))) return rc;
    bfv_rlk_kernel<<<1, (unsigned)n, 2 * n * sizeof(u64), st>>>(bs.ptr<u64>(), bk.ptr<u64>(), (u32)n, q, p, sigma, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bs, &bk}, st);
}
int fhe_bfv_mul_const(uint64_t q, uint64_t n, uint64_t t, uint64_t pq, const uint64_t *rlk, const uint64_t *c, const uint64_t *m,
                      uint64_t *out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(rlk && c && m && out, "fhe_bfv_mul_const: null pointer");
    FHE_REQUIRE(t >= 1 && q >= 2 && n >= 1, "fhe_bfv_mul_const: bad parameters");
    cudaStream_t st = current_stream();
    IoBuf bm;
    Scratch md;
    int rc;
    if ((rc = bm.init(m, batch * n * 8, true, false, st))) return rc;
    if ((rc = md.alloc(batch * 2 * n * 8, st))) return rc;
    const unsigned grid = (unsigned)std::min<size_t>((batch * 2 * n + 255) / 256, (size_t)num_sms() * 16);
    bfv_mul_const_md_kernel<<<grid, 256, 0, st>>>(bm.ptr<u64>(), md.ptr<u64>(), batch, (u32)n, q, t);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return bfv_launch(1, q, n, t, pq, rlk, c, md.ptr<u64>(), out, batch);  // RLWE::mul(t, rlk, c, md)
}
int fhe_bfv_tensor(uint64_t q, uint64_t n, uint64_t t, const uint64_t *a, const uint64_t *b, uint64_t *c012, size_t batch) {
    return bfv_launch(0, q, n, t, 0, nullptr, a, b, c012, batch);
}
int fhe_bfv_relinearize(uint64_t q, uint64_t n, uint64_t pq, const uint64_t *rlk, const uint64_t *c012, uint64_t *out,
                        size_t batch) {
    return bfv_launch(2, q, n, 0, pq, rlk, c012, nullptr, out, batch);
}
int fhe_bfv_mul_relin(uint64_t q, uint64_t n, uint64_t t, uint64_t pq, const uint64_t *rlk, const uint64_t *a,
                      const uint64_t *b, uint64_t *out, size_t batch) {
    return bfv_launch(1, q, n, t, pq, rlk, a, b, out, batch);
}
}
