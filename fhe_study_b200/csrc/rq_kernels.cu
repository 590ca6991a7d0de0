// rq_kernels.cu -- coefficient-wise operations of Rq / Tn that sit either side of the transforms
// (SURVEY 8a rows a6-a9, a11): Add/Sub/Neg/scalar-mul (arith/src/ring_nq.rs:267-281,406-561), remodule /
// mod_switch / mul_div_round / decompose (ring_nq.rs:67-113, zq.rs:133-186), the X^n+1 fold of Rq::from_vec
// (ring_nq.rs:55-63,132-141), and the Tn counterparts (ring_torus.rs:67-113,300-327; torus.rs:43-70).
// All are HBM-bound maps; f64 steps use explicit round-to-nearest intrinsics.
#include "../../include/fhe_b200.h"
#include "runtime.cuh"

namespace fhe {

typedef unsigned __int128 u128d;

__device__ __forceinline__ u64 zq_from_u64(u64 q, u64 v) { return v >= q ? v % q : v; }                 // zq.rs:21-31
__device__ __forceinline__ u64 f64_as_u64(double x) { return __double2ull_rz(x); }                       // saturating, NaN -> 0
__device__ __forceinline__ i64 f64_as_i64_(double x) { return __double2ll_rz(x); }
__device__ __forceinline__ u64 zq_from_f64_(u64 q, double e) {                                           // zq.rs:32-40
    const i64 ei = f64_as_i64_(round(e)), qi = (i64)q;
    if (ei < 0 || ei >= qi) return zq_from_u64(q, (u64)(((ei % qi) + qi) % qi));
    return (u64)ei;
}
__device__ __forceinline__ u32 pow_u32_wrapping(u32 b, u32 e) {
    u32 r = 1;
    for (u32 i = 0; i < e; i++) r *= b;
    return r;
}

enum MapOp {
    OP_RQ_ADD = 0, OP_RQ_SUB, OP_RQ_NEG, OP_RQ_MUL_U64, OP_RQ_REMODULE, OP_RQ_MOD_SWITCH, OP_RQ_MUL_DIV_ROUND,
    OP_TN_MOD_SWITCH, OP_TN_MUL_U64, OP_TN_MUL_DIV_ROUND
};

// c[i] = op(a[i], b[i]); scalar parameters: q (ring modulus), s1, s2
__global__ void map_kernel(int op, const u64 *a, const u64 *b, u64 *c, size_t len, u64 q, u64 s1, u64 s2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
        const u64 x = a[i];
        u64 r = 0;
        switch (op) {
            case OP_RQ_ADD: { u64 v = x + b[i]; r = v >= q ? v - q : v; break; }                          // zq.rs:219-231
            case OP_RQ_SUB: { u64 y = b[i]; r = x >= y ? x - y : (q + x) - y; break; }                     // zq.rs:259-277
            case OP_RQ_NEG: r = x == 0 ? 0 : q - x; break;                                                 // zq.rs:302-314
            case OP_RQ_MUL_U64: r = (u64)(((u128d)x * (u128d)s1) % (u128d)q); break;                       // s1 = from_u64(q, s)
            case OP_RQ_REMODULE: r = zq_from_u64(s1, x); break;                                            // ring_nq.rs:82-88
            case OP_RQ_MOD_SWITCH:                                                                         // zq.rs:133-138
                r = zq_from_u64(s1, f64_as_u64(round(__ddiv_rn(__dmul_rn(__ull2double_rn(x), __ull2double_rn(s1)),
                                                               __ull2double_rn(q)))));
                break;
            case OP_RQ_MUL_DIV_ROUND:                                                                      // ring_nq.rs:106-113
                r = zq_from_f64_(q, round(__ddiv_rn(__dmul_rn(__ull2double_rn(s1), __ull2double_rn(x)), __ull2double_rn(s2))));
                break;
            case OP_TN_MOD_SWITCH:                                                                         // ring_torus.rs:85-101
                r = zq_from_u64(s1, s2 >= 64 ? x : x >> s2);                                               // s2 = 64 - log2(p)
                break;
            case OP_TN_MUL_U64: r = x * s1; break;                                                         // ring_torus.rs:300-327
            case OP_TN_MUL_DIV_ROUND:                                                                      // torus.rs:68-70
                r = f64_as_u64(round(__ddiv_rn(__dmul_rn(__ull2double_rn(s1), __ull2double_rn(x)), __ull2double_rn(s2))));
                break;
        }
        c[i] = r;
    }
}

// Rq::from_vec fold (ring_nq.rs:55-63,132-141): in = `batch` vectors of in_len >= n coefficients, already
// reduced or not (reduce != 0 applies Zq::from_u64 first, as from_vec_u64 does); out = `batch` x n.
__global__ void rq_fold_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, size_t batch, u32 n, u32 in_len, u64 q,
                               int reduce) {
    const size_t total = batch * (size_t)n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t b = idx / n;
        const u32 c = (u32)(idx % n);
        const u64 *v = in + b * (size_t)in_len;
        u64 r = reduce ? zq_from_u64(q, v[c]) : v[c];
        // the reference loop runs p[i-n] -= p[i] for i = n..len-1 (len <= 2n here): p[i] is still the original
        // coefficient when it is subtracted, so position c receives exactly one subtraction, by v[c+n].
        if (c + n < in_len) {
            const u64 y = reduce ? zq_from_u64(q, v[c + n]) : v[c + n];
            r = r >= y ? r - y : (q + r) - y;
        }
        out[idx] = r;
    }
}

// Rq::decompose (ring_nq.rs:67-77 over zq.rs:140-186): out[(poly*l + j)*n + c] = digit j of a[poly*n + c]
__global__ void rq_decompose_kernel(const u64 *__restrict__ a, u64 *__restrict__ out, size_t polys, u32 n, u64 q, u32 beta,
                                    u32 l) {
    const size_t total = polys * (size_t)n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = idx / n;
        const u32 c = (u32)(idx % n);
        const u64 v = a[idx];
        u64 *o = out + poly * (size_t)l * n + c;
        if (beta == 2) {  // zq.rs:174-186
            const bool sat = v >= ((u64)1 << (l & 63));
            for (u32 j = 0; j < l; j++) {
                const u32 sh = l - 1 - j;
                o[(size_t)j * n] = sat ? 1 : zq_from_u64(q, sh < 64 ? ((v >> sh) & 1) : 0);
            }
        } else {  // zq.rs:147-172
            u64 rem = v;
            const bool sat = rem >= (u64)pow_u32_wrapping(beta, l);
            for (u32 i = 1; i <= l; i++) {
                if (sat) { o[(size_t)(i - 1) * n] = (u64)beta - 1; continue; }
                const u64 den = q / (u64)pow_u32_wrapping(beta, i);
                const u64 x_i = rem / den;
                o[(size_t)(i - 1) * n] = zq_from_u64(q, x_i);
                if (x_i != 0) rem = rem % den;
            }
        }
    }
}
// Tn::decompose(2, l) (ring_torus.rs:67-77 + torus.rs:43-52): out[(poly*l + j)*n + c] = bit (l-1-j)
__global__ void tn_decompose_kernel(const u64 *__restrict__ a, u64 *__restrict__ out, size_t polys, u32 n, u32 l) {
    const size_t total = polys * (size_t)n;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t poly = idx / n;
        const u32 c = (u32)(idx % n);
        const u64 v = a[idx];
        for (u32 j = 0; j < l; j++) out[(poly * l + j) * (size_t)n + c] = (v >> (l - 1 - j)) & 1ull;
    }
}

static inline unsigned grid_for(size_t work, int threads = 256) {
    size_t g = (work + threads - 1) / threads;
    const size_t cap = (size_t)num_sms() * 16;
    return (unsigned)(g < 1 ? 1 : g > cap ? cap : g);
}

static int run_map(int op, const u64 *a, const u64 *b, u64 *c, size_t len, u64 q, u64 s1, u64 s2) {
    if (len == 0) return 0;
    FHE_REQUIRE(a && c, "null pointer");
    cudaStream_t st = current_stream();
    IoBuf ba, bb, bc;
    int rc;
    if ((rc = ba.init(a, len * 8, true, false, st))) return rc;
    if ((rc = bb.init(b, len * 8, true, false, st))) return rc;
    if ((rc = bc.init(c, len * 8, false, true, st))) return rc;
    map_kernel<<<grid_for(len), 256, 0, st>>>(op, ba.ptr<u64>(), bb.ptr<u64>(), bc.ptr<u64>(), len, q, s1, s2);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&ba, &bb, &bc}, st);
}

}  // namespace fhe

using namespace fhe;
extern "C" {
int fhe_rq_add(uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len) {
    FHE_REQUIRE(q >= 1 && q < (1ull << 63) && b, "fhe_rq_add: bad modulus or null pointer");
    return run_map(OP_RQ_ADD, a, b, c, len, q, 0, 0);
}
int fhe_rq_sub(uint64_t q, const uint64_t *a, const uint64_t *b, uint64_t *c, size_t len) {
    FHE_REQUIRE(q >= 1 && q < (1ull << 63) && b, "fhe_rq_sub: bad modulus or null pointer");
    return run_map(OP_RQ_SUB, a, b, c, len, q, 0, 0);
}
int fhe_rq_neg(uint64_t q, const uint64_t *a, uint64_t *c, size_t len) { return run_map(OP_RQ_NEG, a, nullptr, c, len, q, 0, 0); }
int fhe_rq_mul_u64(uint64_t q, const uint64_t *a, uint64_t s, uint64_t *c, size_t len) {
    FHE_REQUIRE(q >= 1, "fhe_rq_mul_u64: q must be >= 1");
    return run_map(OP_RQ_MUL_U64, a, nullptr, c, len, q, s >= q ? s % q : s, 0);  // Zq::from_u64(q, s), ring_nq.rs:274-281
}
int fhe_rq_remodule(const uint64_t *a, uint64_t p, uint64_t *c, size_t len) {
    FHE_REQUIRE(p >= 1, "fhe_rq_remodule: p must be >= 1");
    return run_map(OP_RQ_REMODULE, a, nullptr, c, len, 0, p, 0);
}
int fhe_rq_mod_switch(uint64_t q, const uint64_t *a, uint64_t p, uint64_t *c, size_t len) {
    FHE_REQUIRE(q >= 1 && p >= 1, "fhe_rq_mod_switch: moduli must be >= 1");
    return run_map(OP_RQ_MOD_SWITCH, a, nullptr, c, len, q, p, 0);
}
int fhe_rq_mul_div_round(uint64_t q, const uint64_t *a, uint64_t num, uint64_t den, uint64_t *c, size_t len) {
    FHE_REQUIRE(q >= 1 && q < (1ull << 63), "fhe_rq_mul_div_round: q must be in 1..2^63");
    return run_map(OP_RQ_MUL_DIV_ROUND, a, nullptr, c, len, q, num, den);
}
int fhe_tn_mod_switch(const uint64_t *a, uint64_t p, uint64_t *c, size_t len) {
    FHE_REQUIRE(p >= 1 && (p & (p - 1)) == 0, "fhe_tn_mod_switch: p must be a power of two (torus.rs:58-66)");
    return run_map(OP_TN_MOD_SWITCH, a, nullptr, c, len, 0, p, 64 - (63 - __builtin_clzll(p)));
}
int fhe_tn_mul_u64(const uint64_t *a, uint64_t s, uint64_t *c, size_t len) {
    return run_map(OP_TN_MUL_U64, a, nullptr, c, len, 0, s, 0);
}
int fhe_tn_mul_div_round(const uint64_t *a, uint64_t num, uint64_t den, uint64_t *c, size_t len) {
    return run_map(OP_TN_MUL_DIV_ROUND, a, nullptr, c, len, 0, num, den);
}
int fhe_rq_from_vec(uint64_t q, uint64_t n, const uint64_t *in, uint64_t in_len, uint64_t *out, size_t batch) {
    if (batch == 0) return 0;
    FHE_REQUIRE(in && out, "null pointer");
    FHE_REQUIRE(q >= 1 && n >= 1 && in_len >= n, "fhe_rq_from_vec: need in_len >= n (shorter vectors are kept short by the reference)");
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(in, batch * in_len * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, batch * n * 8, false, true, st))) return rc;
    rq_fold_kernel<<<grid_for(batch * n), 256, 0, st>>>(bi.ptr<u64>(), bo.ptr<u64>(), batch, (u32)n, (u32)in_len, q, 1);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bi, &bo}, st);
}
int fhe_rq_decompose(uint64_t q, uint64_t n, const uint64_t *a, uint32_t beta, uint32_t l, uint64_t *out, size_t polys) {
    if (polys == 0) return 0;
    FHE_REQUIRE(a && out, "null pointer");
    FHE_REQUIRE(q >= 1 && n >= 1 && beta >= 2 && l >= 1, "fhe_rq_decompose: need beta >= 2, l >= 1");
    if (beta != 2)
        for (uint32_t i = 1, pw = 1; i <= l; i++) {
            pw *= beta;
            FHE_REQUIRE(pw != 0 && q / pw != 0, "fhe_rq_decompose: q / beta^i is zero (the reference divides by zero here)");
        }
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(a, polys * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, polys * l * n * 8, false, true, st))) return rc;
    rq_decompose_kernel<<<grid_for(polys * n), 256, 0, st>>>(bi.ptr<u64>(), bo.ptr<u64>(), polys, (u32)n, q, beta, l);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bi, &bo}, st);
}
int fhe_tn_decompose(uint64_t n, const uint64_t *a, uint32_t l, uint64_t *out, size_t polys) {
    if (polys == 0) return 0;
    FHE_REQUIRE(a && out, "null pointer");
    FHE_REQUIRE(n >= 1 && l >= 1 && l <= 64, "fhe_tn_decompose: need 1 <= l <= 64");
    cudaStream_t st = current_stream();
    IoBuf bi, bo;
    int rc;
    if ((rc = bi.init(a, polys * n * 8, true, false, st))) return rc;
    if ((rc = bo.init(out, polys * l * n * 8, false, true, st))) return rc;
    tn_decompose_kernel<<<grid_for(polys * n), 256, 0, st>>>(bi.ptr<u64>(), bo.ptr<u64>(), polys, (u32)n, l);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return finish_all({&bi, &bo}, st);
}
}
