// scheme_common.cuh -- device helpers shared by the BFV and CKKS scheme kernels: the reference's Zq conversions
// (arith/src/zq.rs) and the counter-based sampler the device key generation / encryption uses (specified in
// the CPU restatement under oracle/: ctr_draw / ctr_unit; the reference samples from an unseeded thread_rng, which nothing can
// reproduce, so the sampler is ours and the oracle is its written specification).
#pragma once
#include "common.cuh"

namespace fhe {

// Rust `f64 as i64`: saturating, NaN -> 0 (cvt.rzi.s64.f64 has exactly these semantics)
__device__ __forceinline__ i64 f64_as_i64(double x) { return __double2ll_rz(x); }
// Zq::from_f64 (zq.rs:32-40): r = round(e) as i64; out of [0, q): ((r % q) + q) % q with the signed remainder,
// i.e. the mathematical r mod q.  Nearly every scaled coefficient takes that branch (t*v/q >> q), and two software
// 64-bit divisions per coefficient were the bulk of the kernel: |r| mod q is one Barrett step with
// mu = floor((2^64 - 1) / q) (quotient estimate low by at most two for any 64-bit operand), the sign is applied after.
__device__ __forceinline__ u64 zq_from_f64(u64 q, u64 mu, double e) {
    const i64 ei = f64_as_i64(round(e));
    if (ei >= 0 && (u64)ei < q) return (u64)ei;
    const u64 mag = ei < 0 ? (u64)0 - (u64)ei : (u64)ei;  // |r| (2^63 for i64::MIN)
    u64 r = mag - __umul64hi(mag, mu) * q;                 // quotient estimate low by at most 2: r in [0, 3q)
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    return (ei < 0 && r != 0) ? q - r : r;
}
// one coefficient of ring_n::mul_div_round (ring_n.rs:130-138): round((num as f64 * v as f64) / den as f64) -> Zq
__device__ __forceinline__ u64 scale_round(u64 q, u64 mu, i64 v, double num, double den) {
    return zq_from_f64(q, mu, __ddiv_rn(__dmul_rn(num, __ll2double_rn(v)), den));
}
__device__ __forceinline__ u64 zq_sub(u64 q, u64 a, u64 b) { return a >= b ? a - b : (q + a) - b; }  // zq.rs:259-277
__device__ __forceinline__ u64 zq_add(u64 q, u64 a, u64 b) { u64 v = a + b; return v >= q ? v - q : v; }  // zq.rs:219-231

// SplitMix64 output number pos+1 of the stream `seed`, and its top 53 bits as a double in [0, 1)
__device__ __forceinline__ u64 bfv_draw(u64 seed, u64 pos) {
    u64 z = seed + (pos + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double bfv_unit(u64 v) { return __dmul_rn(__ull2double_rn(v >> 11), 1.0 / 9007199254740992.0); }
// sigma * (sum of 12 units - 6): the Irwin-Hall stand-in for Normal(0, sigma) (IEEE adds and one multiply: bit-reproducible)
__device__ __forceinline__ double ctr_gauss(u64 seed, u64 pos0, double sigma) {
    double acc = 0.0;
    for (u32 k = 0; k < 12; k++) acc = __dadd_rn(acc, bfv_unit(bfv_draw(seed, pos0 + k)));
    return __dmul_rn(sigma, __dadd_rn(acc, -6.0));
}

}  // namespace fhe
