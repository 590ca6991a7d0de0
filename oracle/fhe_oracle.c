/*
 * fhe_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C restatement of the ring-arithmetic hot path of arnaucube/fhe-study, written from the
 * reference's Rust sources function by function.  Every function cites the reference file:line it
 * follows (paths relative to the reference tree).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library.
 *
 * PARITY PINNING: the reference has no Rust toolchain available in the build container, so the
 * reference itself cannot be executed.  This restatement is pinned against every known-answer vector
 * the reference's own tests hold for the path (tests/test_oracle_kats.py): arith/src/ring_nq.rs:627-729,
 * arith/src/ring_n.rs:454-483, arith/src/ring_torus.rs:334-366, arith/src/zq.rs:356-435,
 * arith/src/torus.rs:163-190, plus the functional property tests of tfhe/bfv re-run inside the oracle.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fopenmp fhe_oracle.c -o _build/libfhe_oracle.so -lm
 *
 * Rust cast semantics reproduced here:
 *   f64 as i64 / as u64 : saturating, NaN -> 0 (negative -> 0 for u64)
 *   i128 as i64, u128 as u64 : truncation
 *   f64::round : half away from zero (C round())
 *   release-mode integer overflow : wrapping
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef int64_t i64;
typedef unsigned __int128 u128;
typedef __int128 i128;

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * Rust cast helpers
 * ---------------------------------------------------------------------------------------------- */
static inline i64 f64_as_i64(double x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0) return INT64_MAX;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    return (i64)x;
}
static inline u64 f64_as_u64(double x) {
    if (x != x) return 0;
    if (x <= 0.0) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (u64)x;
}

/* ------------------------------------------------------------------------------------------------
 * Zq  (arith/src/zq.rs)
 * ---------------------------------------------------------------------------------------------- */
/* arith/src/zq.rs:12-14 */
static inline u64 modulus_u64(u64 q, u64 e) { return (e % q + q) % q; }
/* arith/src/zq.rs:21-31 */
API u64 orc_zq_from_u64(u64 q, u64 v) { return v >= q ? modulus_u64(q, v) : v; }
/* arith/src/zq.rs:32-40 */
API u64 orc_zq_from_f64(u64 q, double e) {
    i64 ei = f64_as_i64(round(e));
    i64 qi = (i64)q;
    if (ei < 0 || ei >= qi) return orc_zq_from_u64(q, (u64)(((ei % qi) + qi) % qi));
    return (u64)ei;
}
/* arith/src/zq.rs:219-231 */
API u64 orc_zq_add(u64 q, u64 a, u64 b) {
    u64 v = a + b;
    if (v >= q) v -= q;
    return v;
}
/* arith/src/zq.rs:259-277 (by-value Sub, the one the Rq paths use) */
API u64 orc_zq_sub(u64 q, u64 a, u64 b) { return a >= b ? a - b : (q + a) - b; }
/* arith/src/zq.rs:302-314 */
API u64 orc_zq_neg(u64 q, u64 a) { return a == 0 ? 0 : q - a; }
/* arith/src/zq.rs:315-328 */
API u64 orc_zq_mul(u64 q, u64 a, u64 b) { return (u64)(((u128)a * (u128)b) % (u128)q); }
/* arith/src/zq.rs:68-87 */
API u64 orc_zq_exp(u64 q, u64 base, u64 e) {
    u64 res = 1, x = base;
    while (e != 0) {
        if (e & 1) res = orc_zq_mul(q, res, x);
        x = orc_zq_mul(q, x, x);
        e >>= 1;
    }
    return res;
}
/* arith/src/zq.rs:133-138 */
API u64 orc_zq_mod_switch(u64 q, u64 v, u64 q2) {
    return orc_zq_from_u64(q2, f64_as_u64(round(((double)v * (double)q2) / (double)q)));
}
static inline uint32_t u32_pow_wrapping(uint32_t b, uint32_t e) {
    uint32_t r = 1;
    for (uint32_t i = 0; i < e; i++) r *= b;
    return r;
}
/* arith/src/zq.rs:140-186 ; out has l entries, most significant digit first */
API void orc_zq_decompose(u64 q, u64 v, uint32_t beta, uint32_t l, u64 *out) {
    if (beta == 2) {
        /* arith/src/zq.rs:174-186 ; `1 << l as u64` on u64, release-mode shift wraps mod 64 */
        if (v >= ((u64)1 << (l & 63))) {
            for (uint32_t i = 0; i < l; i++) out[i] = 1;
            return;
        }
        for (uint32_t i = 0; i < l; i++) {
            uint32_t sh = l - 1 - i;
            out[i] = orc_zq_from_u64(q, sh < 64 ? ((v >> sh) & 1) : 0);
        }
        return;
    }
    /* arith/src/zq.rs:147-172 */
    u64 rem = v;
    if (rem >= (u64)u32_pow_wrapping(beta, l)) {
        for (uint32_t i = 0; i < l; i++) out[i] = (u64)beta - 1;
        return;
    }
    for (uint32_t i = 1; i <= l; i++) {
        u64 den = q / (u64)u32_pow_wrapping(beta, i);
        u64 x_i = rem / den;
        out[i - 1] = orc_zq_from_u64(q, x_i);
        if (x_i != 0) rem = rem % den;
    }
}

/* ------------------------------------------------------------------------------------------------
 * NTT plan  (arith/src/ntt.rs)
 * ---------------------------------------------------------------------------------------------- */
/* arith/src/ntt.rs:164-179 */
API u64 orc_exp_mod(u64 q, u64 x, u64 k) {
    u128 r = 1, xx = (u128)x % (u128)q, kk = k;
    while (kk > 0) {
        if (kk % 2 == 1) r = (r * xx) % (u128)q;
        xx = (xx * xx) % (u128)q;
        kk /= 2;
    }
    return (u64)r;
}
/* arith/src/ntt.rs:182-185 */
API u64 orc_inv_mod(u64 q, u64 x) { return orc_exp_mod(q, x, q - 2); }
/* arith/src/ntt.rs:115-131 ; n here is the ORDER of the root (the caller passes 2*N). returns 0 when
 * the reference would panic. */
API u64 orc_primitive_root_of_unity(u64 q, u64 n) {
    if (n == 0 || (n & (n - 1)) != 0) return 0;
    if ((q - 1) % n != 0) return 0;
    for (u64 k = 1; k < q; k++) {
        u64 w = orc_exp_mod(q, k, (q - 1) / n);
        if (orc_exp_mod(q, w, n / 2) != 1) return w;
    }
    return 0;
}
static inline u64 bitrev(u64 i, unsigned log_n) { /* (i as u64).reverse_bits() >> (64 - log_n) */
    u64 r = 0;
    for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
    return r;
}
static unsigned ilog2_u64(u64 n) {
    unsigned l = 0;
    while (n > 1) { n >>= 1; l++; }
    return l;
}
/* arith/src/ntt.rs:20-38,133-161 ; fills roots[n], roots_inv[n], returns n_inv (0 on failure) */
API u64 orc_ntt_tables(u64 q, u64 n, u64 *roots, u64 *roots_inv) {
    u64 w = orc_primitive_root_of_unity(q, 2 * n);
    if (w == 0) return 0;
    unsigned log_n = ilog2_u64(n);
    for (u64 i = 0; i < n; i++) {
        roots[i] = orc_exp_mod(q, w, bitrev(i, log_n));
        roots_inv[i] = orc_inv_mod(q, roots[i]);
    }
    return orc_inv_mod(q, n);
}

/* plan cache (the reference's CACHE, arith/src/ntt.rs:18-38) -- single-threaded test infra, tiny */
typedef struct { u64 q, n, n_inv; u64 *roots, *roots_inv; } orc_plan;
static orc_plan g_plans[4096];
static int g_nplans = 0;
static const orc_plan *get_plan(u64 q, u64 n) {
    const orc_plan *found = NULL;
#pragma omp critical(orc_plan_cache)
    {
        for (int i = 0; i < g_nplans; i++)
            if (g_plans[i].q == q && g_plans[i].n == n) found = &g_plans[i];
        if (!found && g_nplans < 4096) {
            orc_plan *p = &g_plans[g_nplans];
            p->q = q; p->n = n;
            p->roots = (u64 *)malloc(sizeof(u64) * n);
            p->roots_inv = (u64 *)malloc(sizeof(u64) * n);
            p->n_inv = orc_ntt_tables(q, n, p->roots, p->roots_inv);
            g_nplans++;
            found = p;
        }
    }
    return found;
}

/* arith/src/ntt.rs:44-73 ; natural order in -> bit-reversed order out (left as the loop leaves it) */
API void orc_ntt(u64 q, u64 n, const u64 *a, u64 *r) {
    const orc_plan *p = get_plan(q, n);
    if (r != a) memcpy(r, a, sizeof(u64) * n);
    u64 t = n / 2, m = 1;
    while (m < n) {
        u64 k = 0;
        for (u64 i = 0; i < m; i++) {
            u64 S = p->roots[m + i];
            for (u64 j = k; j < k + t; j++) {
                u64 U = r[j];
                u64 V = orc_zq_mul(q, r[j + t], S);
                r[j] = orc_zq_add(q, U, V);
                r[j + t] = orc_zq_sub(q, U, V);
            }
            k += 2 * t;
        }
        t /= 2;
        m *= 2;
    }
}
/* arith/src/ntt.rs:78-110 */
API void orc_intt(u64 q, u64 n, const u64 *a, u64 *r) {
    const orc_plan *p = get_plan(q, n);
    if (r != a) memcpy(r, a, sizeof(u64) * n);
    u64 t = 1, m = n / 2;
    while (m > 0) {
        u64 k = 0;
        for (u64 i = 0; i < m; i++) {
            u64 S = p->roots_inv[m + i];
            for (u64 j = k; j < k + t; j++) {
                u64 U = r[j];
                u64 V = r[j + t];
                r[j] = orc_zq_add(q, U, V);
                r[j + t] = orc_zq_mul(q, orc_zq_sub(q, U, V), S);
            }
            k += 2 * t;
        }
        t *= 2;
        m /= 2;
    }
    for (u64 i = 0; i < n; i++) r[i] = orc_zq_mul(q, r[i], p->n_inv);
}

/* ------------------------------------------------------------------------------------------------
 * Rq  (arith/src/ring_nq.rs)
 * ---------------------------------------------------------------------------------------------- */
/* arith/src/ring_nq.rs:132-141 (modulus) ; in place on p[len]; returns the resulting length */
API u64 orc_rq_fold(u64 q, u64 n, u64 *p, u64 len) {
    if (len < n) return len;
    for (u64 i = n; i < len; i++) {
        p[i - n] = orc_zq_sub(q, p[i - n], p[i]);
        p[i] = 0;
    }
    return n;
}
/* arith/src/ring_nq.rs:156-159 ; coefficients reduced with Zq::from_u64, then folded */
API u64 orc_rq_from_vec_u64(u64 q, u64 n, const u64 *in, u64 len, u64 *out) {
    for (u64 i = 0; i < len; i++) out[i] = orc_zq_from_u64(q, in[i]);
    return orc_rq_fold(q, n, out, len);
}
/* arith/src/ring_nq.rs:160-163 */
API u64 orc_rq_from_vec_f64(u64 q, u64 n, const double *in, u64 len, u64 *out) {
    for (u64 i = 0; i < len; i++) out[i] = orc_zq_from_f64(q, in[i]);
    return orc_rq_fold(q, n, out, len);
}
/* arith/src/ring_nq.rs:164-170 */
API u64 orc_rq_from_vec_i64(u64 q, u64 n, const i64 *in, u64 len, u64 *out) {
    for (u64 i = 0; i < len; i++) out[i] = orc_zq_from_f64(q, (double)in[i]);
    return orc_rq_fold(q, n, out, len);
}
/* arith/src/ring_nq.rs:586-607 (mul) / :564-583 (mul_mut).  a_is_evals / b_is_evals say that the
 * operand already is the cached `evals` vector.  c_evals (may be NULL) receives the product's evals. */
API void orc_rq_mul(u64 q, u64 n, const u64 *a, const u64 *b, u64 *c, int a_is_evals, int b_is_evals,
                    u64 *c_evals) {
    u64 *A = (u64 *)malloc(sizeof(u64) * n * 3), *B = A + n, *C = B + n;
    if (a_is_evals) memcpy(A, a, sizeof(u64) * n); else orc_ntt(q, n, a, A);
    if (b_is_evals) memcpy(B, b, sizeof(u64) * n); else orc_ntt(q, n, b, B);
    for (u64 i = 0; i < n; i++) C[i] = orc_zq_mul(q, A[i], B[i]);
    if (c_evals) memcpy(c_evals, C, sizeof(u64) * n);
    orc_intt(q, n, C, c);
    free(A);
}
API void orc_rq_mul_batch(u64 q, u64 n, const u64 *a, const u64 *b, u64 *c, u64 batch, int threads) {
    get_plan(q, n);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++)
        orc_rq_mul(q, n, a + (u64)i * n, b + (u64)i * n, c + (u64)i * n, 0, 0, NULL);
}
API void orc_ntt_batch(u64 q, u64 n, const u64 *a, u64 *r, u64 batch, int inverse, int threads) {
    get_plan(q, n);
#pragma omp parallel for schedule(static) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++) {
        if (inverse) orc_intt(q, n, a + (u64)i * n, r + (u64)i * n);
        else orc_ntt(q, n, a + (u64)i * n, r + (u64)i * n);
    }
}
/* arith/src/ring_nq.rs:406-488 (Add/Sub), :549-561 (Neg) ; op: 0 add, 1 sub, 2 neg */
API void orc_rq_addsub(u64 q, u64 n, const u64 *a, const u64 *b, u64 *c, int op) {
    for (u64 i = 0; i < n; i++)
        c[i] = op == 0 ? orc_zq_add(q, a[i], b[i]) : op == 1 ? orc_zq_sub(q, a[i], b[i]) : orc_zq_neg(q, a[i]);
}
/* arith/src/ring_nq.rs:274-281 (mul_by_u64) */
API void orc_rq_mul_u64(u64 q, u64 n, const u64 *a, u64 s, u64 *c) {
    u64 sq = orc_zq_from_u64(q, s);
    for (u64 i = 0; i < n; i++) c[i] = orc_zq_mul(q, a[i], sq);
}
/* arith/src/ring_nq.rs:82-88 */
API void orc_rq_remodule(u64 n, const u64 *a, u64 p, u64 *c) {
    for (u64 i = 0; i < n; i++) c[i] = orc_zq_from_u64(p, a[i]);
}
/* arith/src/ring_nq.rs:91-101 */
API void orc_rq_mod_switch(u64 q, u64 n, const u64 *a, u64 p, u64 *c) {
    for (u64 i = 0; i < n; i++) c[i] = orc_zq_mod_switch(q, a[i], p);
}
/* arith/src/ring_nq.rs:106-113 */
API void orc_rq_mul_div_round(u64 q, u64 n, const u64 *a, u64 num, u64 den, u64 *c) {
    for (u64 i = 0; i < n; i++)
        c[i] = orc_zq_from_f64(q, round(((double)num * (double)a[i]) / (double)den));
}
/* arith/src/ring_nq.rs:67-77 ; out is l polys of n, out[j*n + c] = digit j of coefficient c */
API void orc_rq_decompose(u64 q, u64 n, const u64 *a, uint32_t beta, uint32_t l, u64 *out) {
    u64 *d = (u64 *)malloc(sizeof(u64) * l);
    for (u64 c = 0; c < n; c++) {
        orc_zq_decompose(q, a[c], beta, l, d);
        for (uint32_t j = 0; j < l; j++) out[(u64)j * n + c] = d[j];
    }
    free(d);
}

/* ------------------------------------------------------------------------------------------------
 * T64 / Tn  (arith/src/torus.rs, arith/src/ring_torus.rs) ; all arithmetic wraps mod 2^64
 * ---------------------------------------------------------------------------------------------- */
/* arith/src/torus.rs:58-66 */
API u64 orc_t64_mod_switch(u64 x, u64 q2) {
    unsigned log2_q2 = 63 - (unsigned)__builtin_clzll(q2);
    unsigned sh = 64 - log2_q2;
    return sh >= 64 ? x : x >> sh; /* release-mode shift wraps; q2>=2 in every caller */
}
/* arith/src/torus.rs:68-70 */
API u64 orc_t64_mul_div_round(u64 x, u64 num, u64 den) {
    return f64_as_u64(round(((double)num * (double)x) / (double)den));
}
/* arith/src/ring_torus.rs:266-298 : exact negacyclic product, O(n^2) u128 schoolbook */
API void orc_tn_mul(u64 n, const u64 *a, const u64 *b, u64 *c) {
    u128 *res = (u128 *)calloc(2 * n - 1, sizeof(u128));
    for (u64 i = 0; i < n; i++)
        for (u64 j = 0; j < n; j++) res[i + j] = res[i + j] + (u128)a[i] * (u128)b[j];
    for (u64 i = n; i < 2 * n - 1; i++) res[i - n] = res[i - n] - res[i]; /* wrapping_sub */
    for (u64 i = 0; i < n; i++) c[i] = (u64)res[i];
    free(res);
}
/* same values, 64-bit wrapping accumulators (exactly equal mod 2^64): used only to make the big
 * functional oracles (TGGSW keygen at n=1024) finish in seconds. Verified against orc_tn_mul in tests. */
static void tn_mul_fast(u64 n, const u64 *a, const u64 *b, u64 *c) {
    u64 *res = (u64 *)calloc(2 * n, sizeof(u64));
    for (u64 i = 0; i < n; i++) {
        u64 ai = a[i];
        if (ai == 0) continue;
        for (u64 j = 0; j < n; j++) res[i + j] += ai * b[j];
    }
    for (u64 i = 0; i < n; i++) c[i] = res[i] - res[i + n];
    free(res);
}
API void orc_tn_mul_fast(u64 n, const u64 *a, const u64 *b, u64 *c) { tn_mul_fast(n, a, b, c); }
API void orc_tn_mul_batch(u64 n, const u64 *a, const u64 *b, u64 *c, u64 batch, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++) orc_tn_mul(n, a + (u64)i * n, b + (u64)i * n, c + (u64)i * n);
}
/* arith/src/ring_torus.rs:118-132 : multiply by X^-h */
API void orc_tn_left_rotate(u64 n, const u64 *a, u64 h, u64 *c) {
    h = h % n;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    for (u64 i = 0; i < n - h; i++) tmp[i] = a[h + i];
    for (u64 i = 0; i < h; i++) tmp[n - h + i] = (u64)0 - a[i];
    memcpy(c, tmp, sizeof(u64) * n);
    free(tmp);
}
/* arith/src/ring_torus.rs:67-77 + arith/src/torus.rs:43-52 ; out[j*n+c] = bit (l-1-j) of a[c] */
API void orc_tn_decompose(u64 n, const u64 *a, uint32_t l, u64 *out) {
    for (uint32_t j = 0; j < l; j++) {
        uint32_t sh = l - 1 - j;
        for (u64 c = 0; c < n; c++) out[(u64)j * n + c] = (a[c] >> sh) & 1;
    }
}
/* arith/src/ring_torus.rs:153-249 ; op 0 add, 1 sub, 2 neg */
API void orc_tn_addsub(u64 n, const u64 *a, const u64 *b, u64 *c, int op) {
    for (u64 i = 0; i < n; i++) c[i] = op == 0 ? a[i] + b[i] : op == 1 ? a[i] - b[i] : (u64)0 - a[i];
}
/* arith/src/ring_torus.rs:300-327 (Mul<T64>, Mul<u64>) */
API void orc_tn_mul_u64(u64 n, const u64 *a, u64 s, u64 *c) {
    for (u64 i = 0; i < n; i++) c[i] = a[i] * s;
}
/* arith/src/ring_torus.rs:85-101 : Tn -> Rq_p */
API void orc_tn_mod_switch(u64 n, const u64 *a, u64 p, u64 *c) {
    for (u64 i = 0; i < n; i++) c[i] = orc_zq_from_u64(p, orc_t64_mod_switch(a[i], p));
}
/* arith/src/ring_torus.rs:106-113 */
API void orc_tn_mul_div_round(u64 n, const u64 *a, u64 num, u64 den, u64 *c) {
    for (u64 i = 0; i < n; i++) c[i] = orc_t64_mul_div_round(a[i], num, den);
}

/* ------------------------------------------------------------------------------------------------
 * R = Z[X]/(X^N+1) over i64  (arith/src/ring_n.rs) -- only what BFV mul/relin uses
 * ---------------------------------------------------------------------------------------------- */
/* arith/src/ring_n.rs:307-320 : LINEAR product, 2n-1 outputs, i128 accumulate then `as i64` */
API void orc_r_naive_mul(u64 n, const i64 *a, const i64 *b, i64 *out) {
    i128 *res = (i128 *)calloc(2 * n - 1, sizeof(i128));
    for (u64 i = 0; i < n; i++)
        for (u64 j = 0; j < n; j++)
            res[i + j] = (i128)((u128)res[i + j] + (u128)((i128)a[i] * (i128)b[j]));
    for (u64 i = 0; i < 2 * n - 1; i++) out[i] = (i64)res[i];
    free(res);
}
/* arith/src/ring_n.rs:142-151 : fold i64 vector mod X^N+1 (wrapping) */
API u64 orc_r_fold(u64 n, i64 *p, u64 len) {
    if (len < n) return len;
    for (u64 i = n; i < len; i++) {
        p[i - n] = (i64)((u64)p[i - n] - (u64)p[i]);
        p[i] = 0;
    }
    return n;
}
/* arith/src/ring_n.rs:130-138 : v has len entries (2n-1 from naive_mul); result is an Rq (n coeffs) */
API void orc_r_mul_div_round(u64 q, u64 n, const i64 *v, u64 len, u64 num, u64 den, u64 *out) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * (len > n ? len : n));
    for (u64 i = 0; i < len; i++)
        tmp[i] = orc_zq_from_f64(q, round(((double)num * (double)v[i]) / (double)den));
    u64 rl = orc_rq_fold(q, n, tmp, len);
    memcpy(out, tmp, sizeof(u64) * rl);
    free(tmp);
}
/* arith/src/ring_n.rs:265-292 (R*R negacyclic over Z) then :81-83 + ring_nq.rs:115-129 (to_rq) */
API void orc_r_mul_to_rq(u64 n, const i64 *a, const i64 *b, u64 q, u64 *out) {
    i128 *res = (i128 *)calloc(2 * n - 1, sizeof(i128));
    for (u64 i = 0; i < n; i++)
        for (u64 j = 0; j < n; j++) res[i + j] += (i128)a[i] * (i128)b[j];
    for (u64 i = n; i < 2 * n - 1; i++) res[i - n] -= res[i];
    for (u64 i = 0; i < n; i++) out[i] = orc_zq_from_f64(q, (double)(i64)res[i]);
    free(res);
}

/* ------------------------------------------------------------------------------------------------
 * BFV tensor / relinearize_204 / mul  (bfv/src/lib.rs)
 * Flat layout: RLWE = [c0 (n), c1 (n)] ; RLK = [rlk0 (n), rlk1 (n)] with coefficients mod p*q.
 * ---------------------------------------------------------------------------------------------- */
/* bfv/src/lib.rs:59-85 */
API void orc_bfv_tensor(u64 q, u64 n, u64 t, const u64 *a, const u64 *b, u64 *c0, u64 *c1, u64 *c2) {
    u64 len = 2 * n - 1;
    i64 *buf = (i64 *)malloc(sizeof(i64) * (4 * n + 4 * len));
    i64 *a0 = buf, *a1 = a0 + n, *b0 = a1 + n, *b1 = b0 + n;
    i64 *r0 = b1 + n, *r1l = r0 + len, *r1r = r1l + len, *r2 = r1r + len;
    for (u64 i = 0; i < n; i++) { /* to_r: ring_n.rs:72-90 */
        a0[i] = (i64)a[i]; a1[i] = (i64)a[n + i];
        b0[i] = (i64)b[i]; b1[i] = (i64)b[n + i];
    }
    orc_r_naive_mul(n, a0, b0, r0);
    orc_r_naive_mul(n, a0, b1, r1l);
    orc_r_naive_mul(n, a1, b0, r1r);
    for (u64 i = 0; i < len; i++) r1l[i] = (i64)((u64)r1l[i] + (u64)r1r[i]);
    orc_r_naive_mul(n, a1, b1, r2);
    orc_r_mul_div_round(q, n, r0, len, t, q, c0);
    orc_r_mul_div_round(q, n, r1l, len, t, q, c1);
    orc_r_mul_div_round(q, n, r2, len, t, q, c2);
    free(buf);
}
/* bfv/src/lib.rs:251-271 ; pq = modulus of the rlk ring */
API void orc_bfv_relinearize_204(u64 q, u64 n, u64 pq, const u64 *rlk, const u64 *c0, const u64 *c1,
                                 const u64 *c2, u64 *out) {
    u64 p = pq / q;
    u64 len = 2 * n - 1;
    i64 *buf = (i64 *)malloc(sizeof(i64) * (3 * n + 2 * len));
    i64 *c2r = buf, *k0 = c2r + n, *k1 = k0 + n, *m0 = k1 + n, *m1 = m0 + len;
    u64 *r = (u64 *)malloc(sizeof(u64) * 2 * n);
    for (u64 i = 0; i < n; i++) { c2r[i] = (i64)c2[i]; k0[i] = (i64)rlk[i]; k1[i] = (i64)rlk[n + i]; }
    orc_r_naive_mul(n, c2r, k0, m0);
    orc_r_naive_mul(n, c2r, k1, m1);
    orc_r_mul_div_round(q, n, m0, len, 1, p, r);
    orc_r_mul_div_round(q, n, m1, len, 1, p, r + n);
    for (u64 i = 0; i < n; i++) {
        out[i] = orc_zq_add(q, c0[i], r[i]);
        out[n + i] = orc_zq_add(q, c1[i], r[n + i]);
    }
    free(buf);
    free(r);
}
/* bfv/src/lib.rs:87-90 */
API void orc_bfv_mul(u64 q, u64 n, u64 t, u64 pq, const u64 *rlk, const u64 *a, const u64 *b, u64 *out) {
    u64 *c = (u64 *)malloc(sizeof(u64) * 3 * n);
    orc_bfv_tensor(q, n, t, a, b, c, c + n, c + 2 * n);
    orc_bfv_relinearize_204(q, n, pq, rlk, c, c + n, c + 2 * n, out);
    free(c);
}
API void orc_bfv_mul_batch(u64 q, u64 n, u64 t, u64 pq, const u64 *rlk, const u64 *a, const u64 *b, u64 *out,
                           u64 batch, int threads) {
#pragma omp parallel for schedule(static) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++)
        orc_bfv_mul(q, n, t, pq, rlk, a + (u64)i * 2 * n, b + (u64)i * 2 * n, out + (u64)i * 2 * n);
}

/* ------------------------------------------------------------------------------------------------
 * TFHE: TGLWE / TGGSW  (tfhe/src/tglwe.rs, tfhe/src/tggsw.rs, gfhe/src/glwe.rs)
 * Flat layouts: TGLWE = (k+1)*n u64, mask polys 0..k-1 then body.
 *   TGLev = l TGLWEs (level j=0 first, gadget 2^63-1).  TGGSW = (k+1) TGLevs (mask rows first, body last).
 * ---------------------------------------------------------------------------------------------- */
/* tfhe/src/tglwe.rs:182-194 : every component times the plaintext polynomial (plaintext on the right) */
static void tglwe_mul_tn(u64 n, u64 k, const u64 *ct, const u64 *p, u64 *out, int fast) {
    for (u64 c = 0; c <= k; c++) {
        if (fast) tn_mul_fast(n, ct + c * n, p, out + c * n);
        else orc_tn_mul(n, ct + c * n, p, out + c * n);
    }
}
API void orc_tglwe_mul_tn(u64 n, u64 k, const u64 *ct, const u64 *p, u64 *out) { tglwe_mul_tn(n, k, ct, p, out, 0); }

/* tfhe/src/tggsw.rs:45-62 (external product) with tggsw.rs:139-149 (TGLev * Vec<Tn>) and the Sum folds
 * of gfhe/src/glwe.rs:236-244.  beta=2, l=64 hard-coded by the reference.  `fast` only switches the
 * Tn*Tn implementation (identical values). */
static void tggsw_extprod(u64 n, u64 k, const u64 *tggsw, const u64 *ct, u64 *out, int fast) {
    const uint32_t l = 64;
    u64 glwe = (k + 1) * n;
    u64 *dec = (u64 *)malloc(sizeof(u64) * l * n);
    u64 *term = (u64 *)malloc(sizeof(u64) * glwe);
    u64 *lev = (u64 *)malloc(sizeof(u64) * glwe);
    for (u64 i = 0; i <= k; i++) {
        orc_tn_decompose(n, ct + i * n, l, dec);
        const u64 *tglev = tggsw + i * l * glwe;
        for (uint32_t j = 0; j < l; j++) {
            tglwe_mul_tn(n, k, tglev + (u64)j * glwe, dec + (u64)j * n, term, fast);
            if (j == 0) memcpy(lev, term, sizeof(u64) * glwe);
            else for (u64 x = 0; x < glwe; x++) lev[x] += term[x];
        }
        if (i == 0) memcpy(out, lev, sizeof(u64) * glwe);
        else for (u64 x = 0; x < glwe; x++) out[x] += lev[x];
    }
    free(dec); free(term); free(lev);
}
API void orc_tggsw_extprod(u64 n, u64 k, const u64 *tggsw, const u64 *ct, u64 *out) {
    tggsw_extprod(n, k, tggsw, ct, out, 0);
}
API void orc_tggsw_extprod_fast(u64 n, u64 k, const u64 *tggsw, const u64 *ct, u64 *out) {
    tggsw_extprod(n, k, tggsw, ct, out, 1);
}
/* tfhe/src/tggsw.rs:39-41 : ct1 + bit (x) (ct2 - ct1) */
static void tggsw_cmux(u64 n, u64 k, const u64 *tggsw, const u64 *ct1, const u64 *ct2, u64 *out, int fast) {
    u64 glwe = (k + 1) * n;
    u64 *d = (u64 *)malloc(sizeof(u64) * 2 * glwe), *e = d + glwe;
    for (u64 x = 0; x < glwe; x++) d[x] = ct2[x] - ct1[x];
    tggsw_extprod(n, k, tggsw, d, e, fast);
    for (u64 x = 0; x < glwe; x++) out[x] = ct1[x] + e[x];
    free(d);
}
API void orc_tggsw_cmux(u64 n, u64 k, const u64 *tggsw, const u64 *ct1, const u64 *ct2, u64 *out) {
    tggsw_cmux(n, k, tggsw, ct1, ct2, out, 0);
}
API void orc_tggsw_cmux_fast(u64 n, u64 k, const u64 *tggsw, const u64 *ct1, const u64 *ct2, u64 *out) {
    tggsw_cmux(n, k, tggsw, ct1, ct2, out, 1);
}
API void orc_extprod_batch(u64 n, u64 k, const u64 *tggsw, const u64 *ct, u64 *out, u64 batch, int threads) {
    u64 glwe = (k + 1) * n;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++) tggsw_extprod(n, k, tggsw, ct + (u64)i * glwe, out + (u64)i * glwe, 0);
}
/* tfhe/src/tglwe.rs:116-119 */
API void orc_tglwe_left_rotate(u64 n, u64 k, const u64 *ct, u64 h, u64 *out) {
    for (u64 c = 0; c <= k; c++) orc_tn_left_rotate(n, ct + c * n, h, out + c * n);
}
/* tfhe/src/tglwe.rs:89-115 ; out = TLWE of dimension k*n (mask k*n, then b) */
API void orc_tglwe_sample_extraction(u64 n, u64 k, const u64 *ct, u64 h, u64 *out) {
    for (u64 i = 0; i < k; i++)
        for (u64 j = 0; j < n; j++)
            out[i * n + j] = j <= h ? ct[i * n + (h - j)] : (u64)0 - ct[i * n + (n + h - j)];
    out[k * n] = ct[k * n + h];
}

/* ------------------------------------------------------------------------------------------------
 * TFHE: TLWE key switch / mod switch / blind rotation / bootstrapping (tfhe/src/tlwe.rs, tlev.rs)
 * KSK flat layout: [i < kn_in][j < l][kn_out + 1].  TLWE = kn+1 (mask then b).
 * ---------------------------------------------------------------------------------------------- */
/* tfhe/src/tlwe.rs:101-112 + tfhe/src/tlev.rs:95-105 + tfhe/src/tlwe.rs:269-279 */
API void orc_tlwe_key_switch(u64 kn_in, u64 kn_out, uint32_t l, const u64 *ksk, const u64 *ct, u64 *out) {
    u64 w = kn_out + 1;
    u64 *rhs = (u64 *)calloc(w, sizeof(u64));
    u64 *lev = (u64 *)malloc(sizeof(u64) * w);
    for (u64 i = 0; i < kn_in; i++) {
        u64 a = ct[i];
        for (uint32_t j = 0; j < l; j++) {
            u64 d = (a >> (l - 1 - j)) & 1; /* T64::decompose, torus.rs:43-52 */
            const u64 *row = ksk + (i * l + j) * w;
            /* TLWE * T64 (wrapping mul by the digit), Sum fold */
            if (j == 0) for (u64 x = 0; x < w; x++) lev[x] = row[x] * d;
            else for (u64 x = 0; x < w; x++) lev[x] += row[x] * d;
        }
        if (i == 0) memcpy(rhs, lev, sizeof(u64) * w);
        else for (u64 x = 0; x < w; x++) rhs[x] += lev[x];
    }
    for (u64 x = 0; x < kn_out; x++) out[x] = (u64)0 - rhs[x];
    out[kn_out] = ct[kn_in] - rhs[kn_out];
    free(rhs); free(lev);
}
/* tfhe/src/tlwe.rs:114-118 */
API void orc_tlwe_mod_switch(u64 kn, const u64 *ct, u64 q2, u64 *out) {
    for (u64 i = 0; i <= kn; i++) out[i] = orc_t64_mod_switch(ct[i], q2);
}
/* tfhe/src/tlwe.rs:196-214 + tfhe/src/tglwe.rs:49-58 ; out = trivial TGLWE (k zero polys, then v) */
API void orc_compute_lookup_table(u64 n, u64 k, u64 t, u64 *out) {
    u64 delta_n = n / t;
    u64 delta = UINT64_MAX / t;
    memset(out, 0, sizeof(u64) * k * n);
    u64 *tmp = (u64 *)malloc(sizeof(u64) * (t * delta_n > n ? t * delta_n : n));
    u64 len = 0;
    for (u64 i = 0; i < t; i++)
        for (u64 d = 0; d < delta_n; d++) tmp[len++] = orc_zq_from_u64(t, i);
    len = orc_rq_fold(t, n, tmp, len);
    for (u64 c = 0; c < n; c++) out[k * n + c] = (c < len ? tmp[c] : 0) * delta;
    free(tmp);
}
/* tfhe/src/tlwe.rs:121-148 AS EXECUTED: the cmux closure is a lazy iterator that is dropped, so the
 * result is table.left_rotate(mod_switch(c).b).  c is a TLWE of dimension c_kn. */
API void orc_blind_rotation_as_executed(u64 n, u64 k, const u64 *c, u64 c_kn, const u64 *table, u64 *out) {
    u64 b = orc_t64_mod_switch(c[c_kn], (u64)(k * n));
    orc_tglwe_left_rotate(n, k, table, b, out);
}
/* tfhe/src/tlwe.rs:138-147 AS WRITTEN (if the map were consumed): for j in 1..k:
 *   c_j = cmux(btk[j], c_j, c_j.left_rotate(a[j])) ; bsk = k TGGSWs.  Extension oracle: no reference
 * execution ever runs this loop. */
API void orc_blind_rotation_as_written(u64 n, u64 k, const u64 *c, u64 c_kn, const u64 *bsk, const u64 *table,
                                       u64 *out) {
    u64 glwe = (k + 1) * n;
    u64 tggsw_sz = (k + 1) * 64 * glwe;
    u64 q2 = (u64)(k * n);
    u64 *rot = (u64 *)malloc(sizeof(u64) * 2 * glwe), *nxt = rot + glwe;
    orc_blind_rotation_as_executed(n, k, c, c_kn, table, out);
    for (u64 j = 1; j < k; j++) {
        u64 aj = orc_t64_mod_switch(c[j], q2);
        orc_tglwe_left_rotate(n, k, out, aj, rot);
        tggsw_cmux(n, k, bsk + j * tggsw_sz, out, rot, nxt, 1);
        memcpy(out, nxt, sizeof(u64) * glwe);
    }
    free(rot);
}
/* X^{-h} acting on a TGLWE.  negacyclic = 0: exactly TGLWE::left_rotate (tglwe.rs:116-119 -> ring_torus.rs:118-132,
 * h reduced mod n, so the sign never flips twice).  negacyclic = 1: h is taken mod 2n and h >= n means
 * -(left_rotate(h - n)), i.e. the true multiplication by X^{-h} in T[X]/(X^n+1) that a working blind rotation
 * needs.  The reference has no such function; mode 1 is an extension. */
static void tglwe_rotate_mode(u64 n, u64 k, const u64 *ct, u64 h, int negacyclic, u64 *out) {
    if (!negacyclic) { orc_tglwe_left_rotate(n, k, ct, h, out); return; }
    h %= 2 * n;
    orc_tglwe_left_rotate(n, k, ct, h % n, out);
    if (h >= n) for (u64 x = 0; x < (k + 1) * n; x++) out[x] = (u64)0 - out[x];
}
/* CMux chain -- the loop blind_rotation spells out (tlwe.rs:138-147: c_j = cmux(btk[j], c_j, c_j.left_rotate(a_j)))
 * composed for `steps` TGGSWs (bsk = steps flat TGGSWs), rotation amounts h[0..steps).  Extension oracle:
 * composition of the reference's own cmux (tggsw.rs:39-41) and left_rotate; no reference execution runs it. */
API void orc_cmux_chain(u64 n, u64 k, u64 steps, const u64 *bsk, const u64 *acc_in, const u64 *h, int negacyclic,
                        u64 *out) {
    u64 glwe = (k + 1) * n;
    u64 tggsw_sz = (k + 1) * 64 * glwe;
    u64 *rot = (u64 *)malloc(sizeof(u64) * 2 * glwe), *nxt = rot + glwe;
    memcpy(out, acc_in, sizeof(u64) * glwe);
    for (u64 j = 0; j < steps; j++) {
        tglwe_rotate_mode(n, k, out, h[j], negacyclic, rot);
        tggsw_cmux(n, k, bsk + j * tggsw_sz, out, rot, nxt, 1);
        memcpy(out, nxt, sizeof(u64) * glwe);
    }
    free(rot);
}
/* Bootstrapping with one TGGSW per mask element of the input TLWE (bsk = steps TGGSWs, steps <= c_kn), i.e. the
 * blind rotation the reference's loop is evidently meant to be.  Extension; composition of reference primitives.
 *  mode 0 ("as written"): c' = mod_switch(c, k*n); acc = table.left_rotate(c'.b); h_j = c'.a[j]; left_rotate.
 *  mode 1 ("working PBS"): c' = mod_switch(c, 2n); acc = X^{-c'.b} * table; h_j = (2n - c'.a[j]) mod 2n, true
 *          negacyclic rotation, so that acc ends as X^{-(b - <a,s>)} * table when bsk[j] encrypts the bit s_j.
 * then sample_extraction(0) and, when ksk != NULL, key_switch(2, 64, ksk) (kn -> kn). */
API void orc_bootstrap_chain(u64 n, u64 k, u64 steps, const u64 *bsk, const u64 *ksk, const u64 *table, const u64 *c,
                             u64 c_kn, int mode, u64 *out) {
    u64 glwe = (k + 1) * n, kn = k * n;
    u64 q2 = mode ? 2 * n : kn;
    u64 *acc = (u64 *)malloc(sizeof(u64) * (2 * glwe + kn + 1)), *acc2 = acc + glwe, *ext = acc2 + glwe;
    u64 *h = (u64 *)malloc(sizeof(u64) * (steps ? steps : 1));
    tglwe_rotate_mode(n, k, table, orc_t64_mod_switch(c[c_kn], q2), mode, acc);
    for (u64 j = 0; j < steps; j++) {
        u64 a = orc_t64_mod_switch(c[j], q2);
        h[j] = mode ? (2 * n - a) % (2 * n) : a;
    }
    orc_cmux_chain(n, k, steps, bsk, acc, h, mode, acc2);
    orc_tglwe_sample_extraction(n, k, acc2, 0, ext);
    if (ksk) orc_tlwe_key_switch(kn, kn, 64, ksk, ext, out);
    else memcpy(out, ext, sizeof(u64) * (kn + 1));
    free(acc); free(h);
}
/* tfhe/src/tlwe.rs:150-161 as executed: blind_rotation -> sample_extraction(0) -> key_switch(2,64) */
API void orc_bootstrapping(u64 n, u64 k, const u64 *ksk, const u64 *table, const u64 *c, u64 c_kn, u64 *out) {
    u64 glwe = (k + 1) * n, kn = k * n;
    u64 *rot = (u64 *)malloc(sizeof(u64) * (glwe + kn + 1)), *ext = rot + glwe;
    orc_blind_rotation_as_executed(n, k, c, c_kn, table, rot);
    orc_tglwe_sample_extraction(n, k, rot, 0, ext);
    orc_tlwe_key_switch(kn, kn, 64, ksk, ext, out);
    free(rot);
}
API void orc_bootstrapping_batch(u64 n, u64 k, const u64 *ksk, const u64 *table, const u64 *c, u64 c_kn, u64 *out,
                                 u64 batch, int threads) {
    u64 kn = k * n;
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++)
        orc_bootstrapping(n, k, ksk, table, c + (u64)i * (c_kn + 1), c_kn, out + (u64)i * (kn + 1));
}
API void orc_key_switch_batch(u64 kn_in, u64 kn_out, uint32_t l, const u64 *ksk, const u64 *ct, u64 *out, u64 batch,
                              int threads) {
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (i64 i = 0; i < (i64)batch; i++)
        orc_tlwe_key_switch(kn_in, kn_out, l, ksk, ct + (u64)i * (kn_in + 1), out + (u64)i * (kn_out + 1));
}

/* ------------------------------------------------------------------------------------------------
 * Deterministic samplers + keygen / encrypt / decrypt restatements used by the functional checks.
 * The reference samples with an unseeded thread_rng, so streams are not reproducible anyway; these
 * follow the same distributions and formulas with a SplitMix64 stream.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { u64 s; } orc_rng;
static inline u64 rng_next(orc_rng *r) {
    u64 z = (r->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double rng_unit(orc_rng *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline double rng_uniform(orc_rng *r, double lo, double hi) { return lo + (hi - lo) * rng_unit(r); }
static inline double rng_normal(orc_rng *r, double sigma) {
    double u1 = rng_unit(r), u2 = rng_unit(r);
    if (u1 < 1e-300) u1 = 1e-300;
    return sigma * sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
API void orc_fill_uniform_u64(u64 seed, u64 *out, u64 len, u64 modulus /* 0 = full 64-bit */) {
    orc_rng r = { seed };
    for (u64 i = 0; i < len; i++) { u64 v = rng_next(&r); out[i] = modulus ? v % modulus : v; }
}
/* T64::rand (arith/src/torus.rs:32-35): round(sample) as u64, saturating */
static inline u64 t64_rand_key(orc_rng *r) { return f64_as_u64(round(rng_uniform(r, 0.0, 2.0))); }
static inline u64 t64_rand_err(orc_rng *r, double sigma) { return f64_as_u64(round(rng_normal(r, sigma))); }

/* TGLWE secret key: k polys with coefficients from Xi_key (gfhe/src/glwe.rs:76-92 via tglwe.rs:39-46) */
API void orc_tglwe_keygen(u64 seed, u64 n, u64 k, u64 *sk) {
    orc_rng r = { seed };
    for (u64 i = 0; i < k * n; i++) sk[i] = t64_rand_key(&r);
}
/* gfhe/src/glwe.rs:140-156 (encrypt_s) for R=Tn.  uniform_mask=0 follows the reference (mask drawn from
 * Xi_key); uniform_mask=1 draws the mask uniformly from Z_2^64 (denser digits for kernel tests). */
static void tglwe_encrypt_s(orc_rng *r, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, int uniform_mask,
                            u64 *ct) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    u64 *b = ct + k * n;
    for (u64 i = 0; i < k * n; i++) ct[i] = uniform_mask ? rng_next(r) : t64_rand_key(r);
    memset(b, 0, sizeof(u64) * n);
    for (u64 i = 0; i < k; i++) { /* TR dot product, tuple_ring.rs:117-134 */
        tn_mul_fast(n, ct + i * n, sk + i * n, tmp);
        for (u64 x = 0; x < n; x++) b[x] += tmp[x];
    }
    for (u64 x = 0; x < n; x++) b[x] += m[x];
    for (u64 x = 0; x < n; x++) b[x] += t64_rand_err(r, sigma);
    free(tmp);
}
API void orc_tglwe_encrypt_s(u64 seed, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, int uniform_mask,
                             u64 *ct) {
    orc_rng r = { seed };
    tglwe_encrypt_s(&r, n, k, sigma, sk, m, uniform_mask, ct);
}
/* gfhe/src/glwe.rs:175-179 : b - <a, s> */
API void orc_tglwe_decrypt(u64 n, u64 k, const u64 *sk, const u64 *ct, u64 *p) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * 2 * n), *acc = tmp + n;
    memset(acc, 0, sizeof(u64) * n);
    for (u64 i = 0; i < k; i++) {
        tn_mul_fast(n, ct + i * n, sk + i * n, tmp);
        for (u64 x = 0; x < n; x++) acc[x] += tmp[x];
    }
    for (u64 x = 0; x < n; x++) p[x] = ct[k * n + x] - acc[x];
    free(tmp);
}
/* tfhe/src/tglwe.rs:49-58 / :59-63 */
API void orc_tglwe_encode(u64 n, u64 t, const u64 *m, u64 *p) {
    u64 delta = UINT64_MAX / t;
    for (u64 i = 0; i < n; i++) p[i] = m[i] * delta;
}
API void orc_tglwe_decode(u64 n, u64 t, const u64 *p, u64 *m) {
    for (u64 i = 0; i < n; i++) m[i] = orc_zq_from_u64(t, orc_t64_mul_div_round(p[i], t, UINT64_MAX));
}
/* tfhe/src/tggsw.rs:17-33,100-122 : TGGSW encryption of the polynomial m under sk */
API void orc_tggsw_encrypt_s(u64 seed, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, int uniform_mask,
                             u64 *out) {
    orc_rng r = { seed };
    const uint32_t l = 64;
    u64 glwe = (k + 1) * n;
    u64 *mi = (u64 *)malloc(sizeof(u64) * 3 * n), *aux = mi + n, *negs = aux + n;
    for (u64 i = 0; i <= k; i++) {
        if (i < k) { /* &-sk_i * m */
            for (u64 x = 0; x < n; x++) negs[x] = (u64)0 - sk[i * n + x];
            tn_mul_fast(n, negs, m, mi);
        } else memcpy(mi, m, sizeof(u64) * n);
        for (uint32_t lv = 1; lv <= l; lv++) {
            if (lv < 64) { u64 g = UINT64_MAX / ((u64)1 << lv); for (u64 x = 0; x < n; x++) aux[x] = mi[x] * g; }
            else memcpy(aux, mi, sizeof(u64) * n);
            tglwe_encrypt_s(&r, n, k, sigma, sk, aux, uniform_mask, out + (i * l + (lv - 1)) * glwe);
        }
    }
    free(mi);
}
/* TLWE keygen / encrypt_s / decrypt (tfhe/src/tlwe.rs:47-82 over gfhe/src/glwe.rs with R=T64) */
API void orc_tlwe_keygen(u64 seed, u64 kn, u64 *sk) {
    orc_rng r = { seed };
    for (u64 i = 0; i < kn; i++) sk[i] = t64_rand_key(&r);
}
static void tlwe_encrypt_s(orc_rng *r, u64 kn, double sigma, const u64 *sk, u64 m, int uniform_mask, u64 *ct) {
    u64 b = 0;
    for (u64 i = 0; i < kn; i++) {
        ct[i] = uniform_mask ? rng_next(r) : t64_rand_key(r);
        b += ct[i] * sk[i];
    }
    b += m;
    b += t64_rand_err(r, sigma);
    ct[kn] = b;
}
API void orc_tlwe_encrypt_s(u64 seed, u64 kn, double sigma, const u64 *sk, u64 m, int uniform_mask, u64 *ct) {
    orc_rng r = { seed };
    tlwe_encrypt_s(&r, kn, sigma, sk, m, uniform_mask, ct);
}
API u64 orc_tlwe_decrypt(u64 kn, const u64 *sk, const u64 *ct) {
    u64 acc = 0;
    for (u64 i = 0; i < kn; i++) acc += ct[i] * sk[i];
    return ct[kn] - acc;
}
/* tfhe/src/tlwe.rs:84-100 + tfhe/src/tlev.rs:53-77 : KSK from sk (dim kn_in) to new_sk (dim kn_out) */
API void orc_tlwe_new_ksk(u64 seed, u64 kn_in, u64 kn_out, uint32_t l, double sigma, const u64 *sk, const u64 *new_sk,
                          int uniform_mask, u64 *ksk) {
    u64 w = kn_out + 1;
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < (i64)kn_in; i++) {
        orc_rng r = { seed + 0x1234567ull * (u64)(i + 1) };
        for (uint32_t lv = 1; lv <= l; lv++) {
            u64 aux = lv < 64 ? sk[i] * (UINT64_MAX / ((u64)1 << lv)) : sk[i];
            tlwe_encrypt_s(&r, kn_out, sigma, new_sk, aux, uniform_mask, ksk + ((u64)i * l + (lv - 1)) * w);
        }
    }
}

/* KSK generation with a COUNTER-BASED sampler (what the device-side generator fhe_ksk_generate reproduces bit for
 * bit).  The reference's RNG stream is an unseeded thread_rng and cannot be reproduced, so the
 * sampler is ours; the STRUCTURE is the reference's (tlwe.rs:84-100 -> tlev.rs:53-77 -> glwe.rs:140-156 with R = T64):
 * row r = i*l + (lv-1) is TLWE_{new_sk}(sk_i * g_lv), g_lv = u64::MAX / 2^lv (lv < 64), 1 (lv = 64).
 * Draw p of row r is output number r*(kn_out+12) + p + 1 of SplitMix64(seed):
 *   p < kn_out : mask word (uniform_mask: the 64 random bits; else Xi_key as the reference does: round(2u) in {0,1,2})
 *   then 12 uniforms for the error: e = round(sigma * (u_0 + .. + u_11 - 6)) cast like T64::rand (negative -> 0). */
static inline u64 ctr_draw(u64 seed, u64 pos) {
    u64 z = seed + (pos + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline double ctr_unit(u64 v) { return (double)(v >> 11) * (1.0 / 9007199254740992.0); }
API void orc_tlwe_new_ksk_ctr(u64 seed, u64 kn_in, u64 kn_out, uint32_t l, double sigma, const u64 *sk, const u64 *new_sk,
                              int uniform_mask, u64 *ksk) {
    u64 w = kn_out + 1, per_row = kn_out + 12;
#pragma omp parallel for schedule(static)
    for (i64 r = 0; r < (i64)(kn_in * l); r++) {
        u64 i = (u64)r / l, lv = (u64)r % l + 1;
        u64 g = lv < 64 ? UINT64_MAX / ((u64)1 << lv) : 1;
        u64 *row = ksk + (u64)r * w, base = (u64)r * per_row;
        u64 b = 0;
        for (u64 x = 0; x < kn_out; x++) {
            u64 v = ctr_draw(seed, base + x);
            u64 a = uniform_mask ? v : f64_as_u64(round(2.0 * ctr_unit(v)));
            row[x] = a;
            b += a * new_sk[x];
        }
        double acc = 0.0;
        for (u64 t = 0; t < 12; t++) acc += ctr_unit(ctr_draw(seed, base + kn_out + t));
        b += sk[i] * g;
        b += f64_as_u64(round(sigma * (acc - 6.0)));
        row[kn_out] = b;
    }
}

/* TLWE::encrypt_s (tlwe.rs:71-74) of `batch` encoded messages with the same counter-based sampler: ciphertext b uses
 * the draws of row b (what fhe_tlwe_encrypt reproduces). */
API void orc_tlwe_encrypt_ctr(u64 seed, u64 kn, double sigma, const u64 *sk, const u64 *msgs, u64 batch, int uniform_mask, u64 *ct) {
    u64 w = kn + 1, per_row = kn + 12;
    for (u64 r = 0; r < batch; r++) {
        u64 *row = ct + r * w, base = r * per_row, b = 0;
        for (u64 x = 0; x < kn; x++) {
            u64 v = ctr_draw(seed, base + x);
            u64 a = uniform_mask ? v : f64_as_u64(round(2.0 * ctr_unit(v)));
            row[x] = a;
            b += a * sk[x];
        }
        double acc = 0.0;
        for (u64 t = 0; t < 12; t++) acc += ctr_unit(ctr_draw(seed, base + kn + t));
        row[kn] = b + msgs[r] + f64_as_u64(round(sigma * (acc - 6.0)));
    }
}

/* TGGSW::encrypt_s (tggsw.rs:17-33 -> tggsw.rs:100-122 -> glwe.rs:140-156 with R = Tn) with the counter-based sampler of
 * orc_tlwe_new_ksk_ctr: row r = i*64 + (lv-1) (i = 0..k: TGLev of -s_i*m for i < k, of m for i = k) is
 * TGLWE_sk(mi * g_lv).  Draw p of row r is output number r*(k*n + 12*n) + p + 1 of SplitMix64(seed):
 *   p = c*n + x (c < k)        : mask coefficient x of component c (uniform_mask, or Xi_key as the reference does)
 *   p = k*n + 12*x + t (t < 12) : the 12 uniforms of error coefficient x.
 * What fhe_tggsw_generate reproduces bit for bit. */
API void orc_tggsw_encrypt_s_ctr(u64 seed, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, int uniform_mask, u64 *out) {
    const uint32_t l = 64;
    u64 glwe = (k + 1) * n, per_row = k * n + 12 * n;
    u64 *mi = (u64 *)malloc(sizeof(u64) * (k + 1) * n), *negs = (u64 *)malloc(sizeof(u64) * n);
    for (u64 i = 0; i <= k; i++) {
        if (i < k) {
            for (u64 x = 0; x < n; x++) negs[x] = (u64)0 - sk[i * n + x];
            tn_mul_fast(n, negs, m, mi + i * n);
        } else memcpy(mi + i * n, m, sizeof(u64) * n);
    }
#pragma omp parallel for schedule(dynamic)
    for (i64 r = 0; r < (i64)((k + 1) * l); r++) {
        u64 i = (u64)r / l, lv = (u64)r % l + 1, base = (u64)r * per_row;
        u64 g = lv < 64 ? UINT64_MAX / ((u64)1 << lv) : 1;
        u64 *row = out + (u64)r * glwe, *b = row + k * n;
        u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
        for (u64 p = 0; p < k * n; p++) {
            u64 v = ctr_draw(seed, base + p);
            row[p] = uniform_mask ? v : f64_as_u64(round(2.0 * ctr_unit(v)));
        }
        memset(b, 0, sizeof(u64) * n);
        for (u64 c = 0; c < k; c++) { /* TR dot product, tuple_ring.rs:117-134 */
            tn_mul_fast(n, row + c * n, sk + c * n, tmp);
            for (u64 x = 0; x < n; x++) b[x] += tmp[x];
        }
        for (u64 x = 0; x < n; x++) {
            double acc = 0.0;
            for (u64 t = 0; t < 12; t++) acc += ctr_unit(ctr_draw(seed, base + k * n + 12 * x + t));
            b[x] += mi[i * n + x] * g;
            b[x] += f64_as_u64(round(sigma * (acc - 6.0)));
        }
        free(tmp);
    }
    free(mi); free(negs);
}

/* TGLWE::encrypt_s (tglwe.rs:76-79) of `batch` encoded message polynomials, sampler and draw addressing of
 * orc_tggsw_encrypt_s_ctr with row = ciphertext index (what fhe_tglwe_encrypt reproduces). */
API void orc_tglwe_encrypt_ctr(u64 seed, u64 n, u64 k, double sigma, const u64 *sk, const u64 *msgs, u64 batch, int uniform_mask,
                               u64 *ct) {
    u64 glwe = (k + 1) * n, per_row = k * n + 12 * n;
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n);
    for (u64 r = 0; r < batch; r++) {
        u64 *row = ct + r * glwe, *b = row + k * n, base = r * per_row;
        for (u64 p = 0; p < k * n; p++) {
            u64 v = ctr_draw(seed, base + p);
            row[p] = uniform_mask ? v : f64_as_u64(round(2.0 * ctr_unit(v)));
        }
        memset(b, 0, sizeof(u64) * n);
        for (u64 c = 0; c < k; c++) {
            tn_mul_fast(n, row + c * n, sk + c * n, tmp);
            for (u64 x = 0; x < n; x++) b[x] += tmp[x];
        }
        for (u64 x = 0; x < n; x++) {
            double acc = 0.0;
            for (u64 t = 0; t < 12; t++) acc += ctr_unit(ctr_draw(seed, base + k * n + 12 * x + t));
            b[x] += msgs[r * n + x] + f64_as_u64(round(sigma * (acc - 6.0)));
        }
    }
    free(tmp);
}

/* ------------------------------------------------------------------------------------------------
 * gfhe: GLWE<Rq> / GLev<Rq> (gfhe/src/glwe.rs, gfhe/src/glev.rs) -- SURVEY 8f rank 2.
 * GLWE<Rq> flat layout: (k+1) polys of n (mask a_0..a_{k-1}, then body b).  KSK = k GLevs of l GLWEs:
 * ksk[(i*l + j)*(k+1)*n ...].
 * ---------------------------------------------------------------------------------------------- */
/* impl Mul<Vec<R>> for GLev<R> (glev.rs:67-80) with GLWE * R (glwe.rs:263-280) and the Sum fold
 * (glwe.rs:236-244): out = sum_j glev[j] * v[j], every component an Rq product (ring_nq.rs:586-607). */
API void orc_glev_rq_mul(u64 q, u64 n, u64 k, u64 l, const u64 *glev, const u64 *v, u64 *out) {
    u64 glwe = (k + 1) * n;
    u64 *term = (u64 *)malloc(sizeof(u64) * n);
    for (u64 j = 0; j < l; j++)
        for (u64 c = 0; c <= k; c++) {
            orc_rq_mul(q, n, glev + j * glwe + c * n, v + j * n, term, 0, 0, NULL);
            if (j == 0) memcpy(out + c * n, term, sizeof(u64) * n);
            else orc_rq_addsub(q, n, out + c * n, term, out + c * n, 0);
        }
    free(term);
}
/* GLWE<R>::key_switch (glwe.rs:126-137) for R = Rq: (0, b) - sum_i ksk_i * a_i.decompose(beta, l) */
API void orc_glwe_rq_key_switch(u64 q, u64 n, u64 k, uint32_t beta, uint32_t l, const u64 *ksk, const u64 *ct, u64 *out) {
    u64 glwe = (k + 1) * n;
    u64 *dec = (u64 *)malloc(sizeof(u64) * l * n);
    u64 *lev = (u64 *)malloc(sizeof(u64) * 2 * glwe), *rhs = lev + glwe;
    for (u64 i = 0; i < k; i++) {
        orc_rq_decompose(q, n, ct + i * n, beta, l, dec);
        orc_glev_rq_mul(q, n, k, l, ksk + i * l * glwe, dec, lev);
        if (i == 0) memcpy(rhs, lev, sizeof(u64) * glwe);
        else for (u64 c = 0; c <= k; c++) orc_rq_addsub(q, n, rhs + c * n, lev + c * n, rhs + c * n, 0);
    }
    for (u64 c = 0; c < k; c++) orc_rq_addsub(q, n, rhs + c * n, NULL, out + c * n, 2); /* 0 - rhs.a */
    orc_rq_addsub(q, n, ct + k * n, rhs + k * n, out + k * n, 1);                        /* b - rhs.b */
    free(dec); free(lev);
}
/* GLWE<Rq>::mod_switch (glwe.rs:197-204): every coefficient through Zq::mod_switch */
API void orc_glwe_rq_mod_switch(u64 q, u64 n, u64 k, const u64 *ct, u64 p, u64 *out) {
    orc_rq_mod_switch(q, (k + 1) * n, ct, p, out);
}
/* keygen / encrypt_s / decrypt / new_ksk for GLWE<Rq> (glwe.rs:76-92,99-125,140-156,175-179; glev.rs:36-54) */
static void rq_rand_f64(orc_rng *r, u64 q, u64 n, int kind, double sigma, u64 *out) {
    for (u64 i = 0; i < n; i++)
        out[i] = orc_zq_from_f64(q, kind == 0 ? rng_uniform(r, 0.0, 2.0) : kind == 1 ? rng_normal(r, sigma) : rng_uniform(r, 0.0, (double)q));
}
API void orc_glwe_rq_keygen(u64 seed, u64 q, u64 n, u64 k, u64 *sk) {
    orc_rng r = { seed };
    rq_rand_f64(&r, q, k * n, 0, 0.0, sk);
}
static void glwe_rq_encrypt_s(orc_rng *r, u64 q, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, u64 *ct) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * n), *b = ct + k * n;
    rq_rand_f64(r, q, k * n, 0, 0.0, ct);                 /* a <- Xi_key (glwe.rs:146-149) */
    for (u64 i = 0; i < k; i++) {                          /* TR dot product */
        orc_rq_mul(q, n, ct + i * n, sk + i * n, tmp, 0, 0, NULL);
        if (i == 0) memcpy(b, tmp, sizeof(u64) * n);
        else orc_rq_addsub(q, n, b, tmp, b, 0);
    }
    orc_rq_addsub(q, n, b, m, b, 0);
    rq_rand_f64(r, q, n, 1, sigma, tmp);
    orc_rq_addsub(q, n, b, tmp, b, 0);
    free(tmp);
}
API void orc_glwe_rq_encrypt_s(u64 seed, u64 q, u64 n, u64 k, double sigma, const u64 *sk, const u64 *m, u64 *ct) {
    orc_rng r = { seed };
    glwe_rq_encrypt_s(&r, q, n, k, sigma, sk, m, ct);
}
API void orc_glwe_rq_decrypt(u64 q, u64 n, u64 k, const u64 *sk, const u64 *ct, u64 *p) {
    u64 *tmp = (u64 *)malloc(sizeof(u64) * 2 * n), *acc = tmp + n;
    for (u64 i = 0; i < k; i++) {
        orc_rq_mul(q, n, ct + i * n, sk + i * n, tmp, 0, 0, NULL);
        if (i == 0) memcpy(acc, tmp, sizeof(u64) * n);
        else orc_rq_addsub(q, n, acc, tmp, acc, 0);
    }
    orc_rq_addsub(q, n, ct + k * n, acc, p, 1);
    free(tmp);
}
API void orc_glwe_rq_new_ksk(u64 seed, u64 q, u64 n, u64 k, uint32_t beta, uint32_t l, double sigma, const u64 *sk,
                             const u64 *new_sk, u64 *ksk) {
    orc_rng r = { seed };
    u64 glwe = (k + 1) * n;
    u64 *aux = (u64 *)malloc(sizeof(u64) * n);
    for (u64 i = 0; i < k; i++) {
        u64 bp = 1;
        for (uint32_t lv = 1; lv <= l; lv++) {
            bp *= beta;                                  /* beta.pow(i) is u32 arithmetic (glev.rs:49) */
            orc_rq_mul_u64(q, n, sk + i * n, q / (u64)(uint32_t)bp, aux);
            glwe_rq_encrypt_s(&r, q, n, k, sigma, new_sk, aux, ksk + (i * l + (lv - 1)) * glwe);
        }
    }
    free(aux);
}

/* BFV keygen / encrypt / decrypt / rlk (bfv/src/lib.rs:120-178,202-225), Rq ops through orc_rq_mul */
API void orc_bfv_keygen(u64 seed, u64 q, u64 n, u64 *sk, u64 *pk /* 2n */) {
    orc_rng r = { seed };
    u64 *a = pk + n, *e = (u64 *)malloc(sizeof(u64) * 2 * n), *na = e + n;
    for (u64 i = 0; i < n; i++) sk[i] = orc_zq_from_u64(q, rng_next(&r) % 2);
    for (u64 i = 0; i < n; i++) a[i] = orc_zq_from_u64(q, rng_next(&r) % q);
    for (u64 i = 0; i < n; i++) e[i] = orc_zq_from_f64(q, rng_normal(&r, 3.2));
    orc_rq_addsub(q, n, a, NULL, na, 2);
    orc_rq_mul(q, n, na, sk, pk, 0, 0, NULL); /* -a * s */
    orc_rq_addsub(q, n, pk, e, pk, 0);
    free(e);
}
API void orc_bfv_encrypt(u64 seed, u64 q, u64 n, u64 t, const u64 *pk, const u64 *m, u64 *ct /* 2n */) {
    orc_rng r = { seed };
    u64 *u = (u64 *)malloc(sizeof(u64) * 4 * n), *e1 = u + n, *e2 = e1 + n, *md = e2 + n;
    for (u64 i = 0; i < n; i++) u[i] = orc_zq_from_f64(q, rng_uniform(&r, -1.0, 1.0));
    for (u64 i = 0; i < n; i++) e1[i] = orc_zq_from_f64(q, rng_normal(&r, 3.2));
    for (u64 i = 0; i < n; i++) e2[i] = orc_zq_from_f64(q, rng_normal(&r, 3.2));
    orc_rq_remodule(n, m, q, md);
    orc_rq_mul_u64(q, n, md, q / t, md);
    orc_rq_mul(q, n, pk, u, ct, 0, 0, NULL);
    orc_rq_addsub(q, n, ct, e1, ct, 0);
    orc_rq_addsub(q, n, ct, md, ct, 0);
    orc_rq_mul(q, n, pk + n, u, ct + n, 0, 0, NULL);
    orc_rq_addsub(q, n, ct + n, e2, ct + n, 0);
    free(u);
}
/* BFV::encrypt (lib.rs:142-160) with the counter-based sampler the device reproduces (fhe_bfv_encrypt): draw p of ciphertext r
 * is SplitMix64 output r*25n + p + 1; p < n: u_x; n + 12x + t: e1_x; 13n + 12x + t: e2_x (ctr_draw / ctr_unit above). */
API void orc_bfv_encrypt_ctr(u64 seed, u64 q, u64 n, u64 t, double sigma, const u64 *pk, const u64 *msgs, u64 batch, u64 *ct) {
    u64 *u = (u64 *)malloc(sizeof(u64) * 4 * n), *e1 = u + n, *e2 = e1 + n, *md = e2 + n;
    for (u64 r = 0; r < batch; r++) {
        u64 base = r * 25 * n, *c = ct + r * 2 * n;
        for (u64 x = 0; x < n; x++) {
            u[x] = orc_zq_from_f64(q, -1.0 + 2.0 * ctr_unit(ctr_draw(seed, base + x)));
            double a1 = 0.0, a2 = 0.0;
            for (u64 k = 0; k < 12; k++) {
                a1 += ctr_unit(ctr_draw(seed, base + n + 12 * x + k));
                a2 += ctr_unit(ctr_draw(seed, base + 13 * n + 12 * x + k));
            }
            e1[x] = orc_zq_from_f64(q, sigma * (a1 - 6.0));
            e2[x] = orc_zq_from_f64(q, sigma * (a2 - 6.0));
        }
        orc_rq_remodule(n, msgs + r * n, q, md);
        orc_rq_mul_u64(q, n, md, q / t, md);
        orc_rq_mul(q, n, pk, u, c, 0, 0, NULL);
        orc_rq_addsub(q, n, c, e1, c, 0);
        orc_rq_addsub(q, n, c, md, c, 0);
        orc_rq_mul(q, n, pk + n, u, c + n, 0, 0, NULL);
        orc_rq_addsub(q, n, c + n, e2, c + n, 0);
    }
    free(u);
}
API void orc_bfv_decrypt(u64 q, u64 n, u64 t, const u64 *sk, const u64 *ct, u64 *m) {
    u64 *cs = (u64 *)malloc(sizeof(u64) * n);
    orc_rq_mul(q, n, ct + n, sk, cs, 0, 0, NULL);
    orc_rq_addsub(q, n, ct, cs, cs, 0);
    orc_rq_mul_div_round(q, n, cs, t, q, cs);
    orc_rq_remodule(n, cs, t, m);
    free(cs);
}
/* bfv/src/lib.rs:93-98 (tmp_naive_mul) in the ring mod pq */
static void tmp_naive_mul(u64 pq, u64 n, const u64 *a, const u64 *b, u64 *out) {
    u64 len = 2 * n - 1;
    i64 *buf = (i64 *)malloc(sizeof(i64) * (2 * n + len));
    u64 *tmp = (u64 *)malloc(sizeof(u64) * len);
    for (u64 i = 0; i < n; i++) { buf[i] = (i64)a[i]; buf[n + i] = (i64)b[i]; }
    orc_r_naive_mul(n, buf, buf + n, buf + 2 * n);
    orc_rq_from_vec_i64(pq, n, buf + 2 * n, len, tmp);
    memcpy(out, tmp, sizeof(u64) * n);
    free(buf); free(tmp);
}
API void orc_bfv_rlk_key(u64 seed, u64 q, u64 n, u64 p, const u64 *sk, u64 *rlk /* 2n, mod p*q */) {
    orc_rng r = { seed };
    u64 pq = p * q;
    u64 *s = (u64 *)malloc(sizeof(u64) * 4 * n), *e = s + n, *as = e + n, *ss = as + n;
    u64 *a = rlk + n;
    orc_rq_remodule(n, sk, pq, s);
    for (u64 i = 0; i < n; i++) a[i] = orc_zq_from_u64(pq, rng_next(&r) % pq);
    for (u64 i = 0; i < n; i++) e[i] = orc_zq_from_f64(pq, rng_normal(&r, 3.2));
    tmp_naive_mul(pq, n, a, s, as);
    orc_rq_addsub(pq, n, as, e, as, 0);
    orc_rq_addsub(pq, n, as, NULL, as, 2);
    tmp_naive_mul(pq, n, s, s, ss);
    orc_rq_mul_u64(pq, n, ss, p, ss);
    orc_rq_addsub(pq, n, as, ss, rlk, 0);
    free(s);
}

/* ---- device-reproducible (counter-based) key generation for BFV, and the CKKS Rq paths ----------
 * Same sampler as orc_bfv_encrypt_ctr: draw p = SplitMix64 output p + 1 (ctr_draw), units = top 53 bits (ctr_unit),
 * Normal(0, sigma) stand-in = sigma * (sum of 12 units - 6). */
static double ctr_gauss(u64 seed, u64 pos0, double sigma) {
    double acc = 0.0;
    for (u64 k = 0; k < 12; k++) acc += ctr_unit(ctr_draw(seed, pos0 + k));
    return sigma * (acc - 6.0);
}
/* BFV::new_key (bfv/src/lib.rs:120-140): s <- Uniform(0,2) (u64), a <- Uniform(0,q) (u64), e <- Normal; pk = (-a*s + e, a).
 * draws: p < n: s_p = draw % 2 ; n + x: a_x = draw % q ; 2n + 12x + t: e_x. */
API void orc_bfv_keygen_ctr(u64 seed, u64 q, u64 n, double sigma, u64 *sk, u64 *pk /* 2n */) {
    u64 *a = pk + n, *e = (u64 *)malloc(sizeof(u64) * 2 * n), *na = e + n;
    for (u64 x = 0; x < n; x++) {
        sk[x] = orc_zq_from_u64(q, ctr_draw(seed, x) % 2);
        a[x] = orc_zq_from_u64(q, ctr_draw(seed, n + x) % q);
        e[x] = orc_zq_from_f64(q, ctr_gauss(seed, 2 * n + 12 * x, sigma));
    }
    orc_rq_addsub(q, n, a, NULL, na, 2);
    orc_rq_mul(q, n, na, sk, pk, 0, 0, NULL); /* &(-a) * &s */
    orc_rq_addsub(q, n, pk, e, pk, 0);
    free(e);
}
/* BFV::rlk_key (bfv/src/lib.rs:202-225) in the ring mod pq = p*q, products through tmp_naive_mul (lib.rs:93-98):
 * rlk = ( -(a*s + e) + (s*s)*p , a ).  draws: p < n: a_x = draw % pq ; n + 12x + t: e_x. */
API void orc_bfv_rlk_key_ctr(u64 seed, u64 q, u64 n, u64 p, double sigma, const u64 *sk, u64 *rlk /* 2n */) {
    u64 pq = p * q;
    u64 *s = (u64 *)malloc(sizeof(u64) * 4 * n), *e = s + n, *as = e + n, *ss = as + n;
    u64 *a = rlk + n;
    orc_rq_remodule(n, sk, pq, s);
    for (u64 x = 0; x < n; x++) {
        a[x] = orc_zq_from_u64(pq, ctr_draw(seed, x) % pq);
        e[x] = orc_zq_from_f64(pq, ctr_gauss(seed, n + 12 * x, sigma));
    }
    tmp_naive_mul(pq, n, a, s, as);
    orc_rq_addsub(pq, n, as, e, as, 0);
    orc_rq_addsub(pq, n, as, NULL, as, 2);
    tmp_naive_mul(pq, n, s, s, ss);
    orc_rq_mul_u64(pq, n, ss, p, ss);
    orc_rq_addsub(pq, n, as, ss, rlk, 0);
    free(s);
}
/* BFV::mul_const (bfv/src/lib.rs:189-200): md = (m.remodule(q) * floor(q/t), 0); RLWE::mul(t, rlk, c, md) */
API void orc_bfv_mul_const(u64 q, u64 n, u64 t, u64 pq, const u64 *rlk, const u64 *c, const u64 *m, u64 *out) {
    u64 *md = (u64 *)calloc(2 * n, sizeof(u64));
    orc_rq_remodule(n, m, q, md);
    orc_rq_mul_u64(q, n, md, q / t, md);
    orc_bfv_mul(q, n, t, pq, rlk, c, md, out);
    free(md);
}

/* CKKS over Rq (ckks/src/lib.rs:46-119).  new_key (:46-63): e <- Normal, s <- Uniform(-1,1) (f64 -> Zq::from_f64),
 * a <- Uniform(-1,1); pk = (-a*s + e, a).  draws: p < n: s_p = from_f64(-1 + 2u) ; n + x: a_x ; 2n + 12x + t: e_x. */
API void orc_ckks_keygen_ctr(u64 seed, u64 q, u64 n, double sigma, u64 *sk, u64 *pk /* 2n */) {
    u64 *a = pk + n, *e = (u64 *)malloc(sizeof(u64) * 2 * n), *na = e + n;
    for (u64 x = 0; x < n; x++) {
        sk[x] = orc_zq_from_f64(q, -1.0 + 2.0 * ctr_unit(ctr_draw(seed, x)));
        a[x] = orc_zq_from_f64(q, -1.0 + 2.0 * ctr_unit(ctr_draw(seed, n + x)));
        e[x] = orc_zq_from_f64(q, ctr_gauss(seed, 2 * n + 12 * x, sigma));
    }
    orc_rq_addsub(q, n, a, NULL, na, 2);
    orc_rq_mul(q, n, na, sk, pk, 0, 0, NULL);
    orc_rq_addsub(q, n, pk, e, pk, 0);
    free(e);
}
/* CKKS::encrypt (:66-84): m in R (i64 coefficients) -> m.to_rq(q) (ring_nq.rs:116-129: Zq::from_f64(c as f64));
 * ct = (m + e_0 + v*pk.0, v*pk.1 + e_1).  draws of ciphertext r (base r*25n): p < n: v_p ; n + 12x + t: e0_x ; 13n + 12x + t: e1_x. */
API void orc_ckks_encrypt_ctr(u64 seed, u64 q, u64 n, double sigma, const u64 *pk, const i64 *msgs, u64 batch, u64 *ct) {
    u64 *v = (u64 *)malloc(sizeof(u64) * 4 * n), *e0 = v + n, *e1 = e0 + n, *mq = e1 + n;
    for (u64 r = 0; r < batch; r++) {
        u64 base = r * 25 * n, *c = ct + r * 2 * n;
        for (u64 x = 0; x < n; x++) {
            v[x] = orc_zq_from_f64(q, -1.0 + 2.0 * ctr_unit(ctr_draw(seed, base + x)));
            e0[x] = orc_zq_from_f64(q, ctr_gauss(seed, base + n + 12 * x, sigma));
            e1[x] = orc_zq_from_f64(q, ctr_gauss(seed, base + 13 * n + 12 * x, sigma));
            mq[x] = orc_zq_from_f64(q, (double)msgs[r * n + x]);
        }
        orc_rq_mul(q, n, v, pk, c, 0, 0, NULL);
        orc_rq_addsub(q, n, c, e0, c, 0);
        orc_rq_addsub(q, n, c, mq, c, 0);
        orc_rq_mul(q, n, v, pk + n, c + n, 0, 0, NULL);
        orc_rq_addsub(q, n, c + n, e1, c + n, 0);
    }
    free(v);
}
/* CKKS::decrypt (:86-94): m = c.0 + c.1*s, then Rq::mod_centered_q (ring_nq.rs:359-361 -> ring_n.rs:113-127):
 * res = v % q; if res > q/2 { res - q } on i64 */
API void orc_ckks_decrypt(u64 q, u64 n, const u64 *sk, const u64 *ct, u64 batch, i64 *m) {
    u64 *cs = (u64 *)malloc(sizeof(u64) * n);
    for (u64 r = 0; r < batch; r++) {
        const u64 *c = ct + r * 2 * n;
        orc_rq_mul(q, n, c + n, sk, cs, 0, 0, NULL);
        orc_rq_addsub(q, n, c, cs, cs, 0);
        for (u64 x = 0; x < n; x++) {
            i64 res = (i64)cs[x] % (i64)q;
            if (res > (i64)q / 2) res -= (i64)q;
            m[r * n + x] = res;
        }
    }
    free(cs);
}
/* CKKS::add (:113-115) and CKKS::sub (:116-118).  As written, sub subtracts the first components and ADDS the
 * second ones ((&c0.0 - &c1.0, &c0.1 + &c1.1)); reproduced as is. */
API void orc_ckks_addsub(u64 q, u64 n, const u64 *c0, const u64 *c1, u64 batch, int sub, u64 *out) {
    for (u64 r = 0; r < batch; r++) {
        orc_rq_addsub(q, n, c0 + r * 2 * n, c1 + r * 2 * n, out + r * 2 * n, sub ? 1 : 0);
        orc_rq_addsub(q, n, c0 + r * 2 * n + n, c1 + r * 2 * n + n, out + r * 2 * n + n, 0);
    }
}
