// ckks_kernels.cu -- the Rq paths of CKKS (ckks/src/lib.rs:46-119) on the NTT polymul kernels:
// new_key, encrypt, decrypt, add, sub.  The canonical-embedding encoder (ckks/src/encoder.rs: a dense complex
// Vandermonde solve on <= 32 points) is out of scope; plaintexts enter and leave as elements of
// R = Z[X]/(X^n+1) (i64 coefficients), exactly what CKKS::encrypt / decrypt take and return.  Sampling uses the
// counter-based sampler specified in the oracle (orc_ckks_*_ctr).
#include <algorithm>

#include "../../include/fhe_b200.h"
#include "runtime.cuh"
#include "scheme_common.cuh"

struct fhe_ntt_plan;
namespace fhe {
int plan_launch(const fhe_ntt_plan *plan, int mode, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, size_t batch, int flags,
                cudaStream_t st);  // lib_core.cu

// new_key (lib.rs:46-63): s, a <- Uniform(-1,1) -> Zq::from_f64, e <- Normal.  draws: p < n: s_p ; n + x: a_x ; 2n + 12x + t: e_x
__global__ void ckks_keygen_sample_kernel(u64 *__restrict__ sk, u64 *__restrict__ neg_a, u64 *__restrict__ a, u64 *__restrict__ e,
                                          u32 n, u64 q, double sigma, u64 seed) {
    const u64 mu = ~0ull / q;
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
        sk[x] = zq_from_f64(q, mu, __dadd_rn(-1.0, __dmul_rn(2.0, bfv_unit(bfv_draw(seed, x)))));
        const u64 av = zq_from_f64(q, mu, __dadd_rn(-1.0, __dmul_rn(2.0, bfv_unit(bfv_draw(seed, (u64)n + x)))));
        a[x] = av;
        neg_a[x] = av == 0 ? 0 : q - av;
        e[x] = zq_from_f64(q, mu, ctr_gauss(seed, 2 * (u64)n + 12 * (u64)x, sigma));
    }
}
__global__ void ckks_add_inplace_kernel(u64 *__restrict__ acc, const u64 *__restrict__ b, u32 n, u64 q) {
    for (u32 x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) acc[x] = zq_add(q, acc[x], b[x]);
}
// encrypt (lib.rs:66-84): v of ciphertext r (draws r*25n + x)
__global__ void ckks_sample_v_kernel(u64 *__restrict__ V, size_t batch, u32 n, u64 q, u64 seed) {
    const size_t total = batch * n;
    const u64 mu = ~0ull / q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n;
        const u32 x = (u32)(i % n);
        V[i] = zq_from_f64(q, mu, __dadd_rn(-1.0, __dmul_rn(2.0, bfv_unit(bfv_draw(seed, r * 25 * (size_t)n + x)))));
    }
}
// ct = (m.to_rq(q) + e_0 + v*pk.0, v*pk.1 + e_1); m.to_rq: Zq::from_f64(c as f64) (ring_nq.rs:116-129)
__global__ void ckks_encrypt_finish_kernel(const u64 *__restrict__ P0, const u64 *__restrict__ P1, const i64 *__restrict__ m,
                                           u64 *__restrict__ ct, size_t batch, u32 n, u64 q, double sigma, u64 seed) {
    const size_t total = batch * n;
    const u64 mu = ~0ull / q;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / n, base = r * 25 * (size_t)n;
        const u32 x = (u32)(i % n);
        const u64 e0 = zq_from_f64(q, mu, ctr_gauss(seed, base + n + 12 * (size_t)x, sigma));
        const u64 e1 = zq_from_f64(q, mu, ctr_gauss(seed, base + 13 * (size_t)n + 12 * (size_t)x, sigma));
        const u64 mq = zq_from_f64(q, mu, __ll2double_rn(m[i]));
        ct[r * 2 * n + x] = zq_add(q, zq_add(q, P0[i], e0), mq);
        ct[r * 2 * n + n + x] = zq_add(q, P1[i], e1);
    }
}
__global__ void ckks_gather_c1_kernel(const u64 *__restrict__ ct, u64 *__restrict__ c1, size_t batch, u32 n) {
    const size_t total = batch * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        c1[i] = ct[(i / n) * 2 * n + n + i % n];
}
// decrypt (lib.rs:86-94): m = c.0 + c.1*s, then mod_centered_q (ring_n.rs:113-127): res = v % q; if res > q/2 { res - q }
__global__ void ckks_decrypt_finish_kernel(const u64 *__restrict__ ct, const u64 *__restrict__ c1s, i64 *__restrict__ m, size_t batch,
                                           u32 n, u64 q) {
    const size_t total = batch * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const u64 cs = zq_add(q, ct[(i / n) * 2 * n + i % n], c1s[i]);
        i64 res = (i64)cs % (i64)q;
        if (res > (i64)q / 2) res -= (i64)q;
        m[i] = res;
    }
}
// add (lib.rs:113-115): component-wise; sub (lib.rs:116-118) AS WRITTEN: (c0.0 - c1.0, c0.1 + c1.1)
__global__ void ckks_addsub_kernel(const u64 *__restrict__ c0, const u64 *__restrict__ c1, u64 *__restrict__ out, size_t batch, u32 n,
                                   u64 q, int sub) {
    const size_t total = batch * 2 * n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const bool first = (i % (2 * n)) < n;
        out[i] = (sub && first) ? zq_sub(q, c0[i], c1[i]) : zq_add(q, c0[i], c1[i]);
    }
}
}  // namespace fhe

using namespace fhe;
static unsigned flat_grid(size_t work) { return (unsigned)std::min<size_t>((work + 255) / 256, (size_t)num_sms() * 16); }

extern "C" {
int fhe_ckks_keygen(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, double sigma, uint64_t seed, uint64_t *sk, uint64_t *pk) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    FHE_REQUIRE(sk && pk, "fhe_ckks_keygen: null pointer");
    cudaStream_t st = current_stream();
    IoBuf bs, bp;
    Scratch tmp;
    int rc;
    if ((rc = bs.init(sk, n * 8, false, true, st))) return rc;
    if ((rc = bp.init(pk, 2 * n * 8, false, true, st))) return rc;
    if ((rc = tmp.alloc(2 * n * 8, st))) return rc;
    ckks_keygen_sample_kernel<<<flat_grid(n), 256, 0, st>>>(bs.ptr<u64>(), tmp.ptr<u64>(), bp.ptr<u64>() + n, tmp.ptr<u64>() + n, (u32)n,
                                                          q, sigma, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = plan_launch(plan, 2, tmp.ptr<u64>(), bs.ptr<u64>(), bp.ptr<u64>(), nullptr, 1, 0, st))) return rc;  // (&(-a) * &s)
    ckks_add_inplace_kernel<<<flat_grid(n), 256, 0, st>>>(bp.ptr<u64>(), tmp.ptr<u64>() + n, (u32)n, q);          // + e
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bs, &bp}, st);
}
int fhe_ckks_encrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, const uint64_t *pk, const int64_t *m, double sigma,
                     uint64_t seed, uint64_t *ct, size_t batch) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(pk && m && ct, "fhe_ckks_encrypt: null pointer");
    cudaStream_t st = current_stream();
    IoBuf bp, bm, bc;
    Scratch V, P0, P1;
    int rc;
    if ((rc = bp.init(pk, 2 * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(m, batch * n * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * 2 * n * 8, false, true, st))) return rc;
    if ((rc = V.alloc(batch * n * 8, st)) || (rc = P0.alloc(batch * n * 8, st)) || (rc = P1.alloc(batch * n * 8, st))) return rc;
    ckks_sample_v_kernel<<<flat_grid(batch * n), 256, 0, st>>>(V.ptr<u64>(), batch, (u32)n, q, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = plan_launch(plan, 2, V.ptr<u64>(), bp.ptr<u64>(), P0.ptr<u64>(), nullptr, batch, 4, st))) return rc;      // &v * &pk.0
    if ((rc = plan_launch(plan, 2, V.ptr<u64>(), bp.ptr<u64>() + n, P1.ptr<u64>(), nullptr, batch, 4, st))) return rc;  // &v * &pk.1
    ckks_encrypt_finish_kernel<<<flat_grid(batch * n), 256, 0, st>>>(P0.ptr<u64>(), P1.ptr<u64>(), bm.ptr<i64>(), bc.ptr<u64>(), batch,
                                                                   (u32)n, q, sigma, seed);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bp, &bm, &bc}, st);
}
int fhe_ckks_decrypt(const fhe_ntt_plan *plan, uint64_t q, uint64_t n, const uint64_t *sk, const uint64_t *ct, int64_t *m,
                     size_t batch) {
    FHE_REQUIRE(plan != nullptr, "null plan");
    if (batch == 0) return 0;
    FHE_REQUIRE(sk && ct && m, "fhe_ckks_decrypt: null pointer");
    cudaStream_t st = current_stream();
    IoBuf bs, bc, bm;
    Scratch c1, c1s;
    int rc;
    if ((rc = bs.init(sk, n * 8, true, false, st))) return rc;
    if ((rc = bc.init(ct, batch * 2 * n * 8, true, false, st))) return rc;
    if ((rc = bm.init(m, batch * n * 8, false, true, st))) return rc;
    if ((rc = c1.alloc(batch * n * 8, st)) || (rc = c1s.alloc(batch * n * 8, st))) return rc;
    ckks_gather_c1_kernel<<<flat_grid(batch * n), 256, 0, st>>>(bc.ptr<u64>(), c1.ptr<u64>(), batch, (u32)n);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    if ((rc = plan_launch(plan, 2, c1.ptr<u64>(), bs.ptr<u64>(), c1s.ptr<u64>(), nullptr, batch, 4, st))) return rc;  // &c.1 * &sk.0
    ckks_decrypt_finish_kernel<<<flat_grid(batch * n), 256, 0, st>>>(bc.ptr<u64>(), c1s.ptr<u64>(), bm.ptr<i64>(), batch, (u32)n, q);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bs, &bc, &bm}, st);
}
static int ckks_addsub(uint64_t q, uint64_t n, const uint64_t *c0, const uint64_t *c1, uint64_t *out, size_t batch, int sub) {
    if (batch == 0) return 0;
    FHE_REQUIRE(c0 && c1 && out, "fhe_ckks_add/sub: null pointer");
    cudaStream_t st = current_stream();
    IoBuf b0, b1, bo;
    int rc;
    if ((rc = b0.init(c0, batch * 2 * n * 8, true, false, st))) return rc;
    if ((rc = b1.init(c1, batch * 2 * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * 2 * n * 8, false, true, st))) return rc;
    ckks_addsub_kernel<<<flat_grid(batch * 2 * n), 256, 0, st>>>(b0.ptr<u64>(), b1.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, q, sub);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&b0, &b1, &bo}, st);
}
int fhe_ckks_add(uint64_t q, uint64_t n, const uint64_t *c0, const uint64_t *c1, uint64_t *out, size_t batch) {
    return ckks_addsub(q, n, c0, c1, out, batch, 0);
}
int fhe_ckks_sub(uint64_t q, uint64_t n, const uint64_t *c0, const uint64_t *c1, uint64_t *out, size_t batch) {
    return ckks_addsub(q, n, c0, c1, out, batch, 1);
}
}
