// modarith.cuh -- modular-arithmetic policies shared by every NTT kernel (host+device so that the
// per-thread logic can be replayed on the CPU by tests/emu, which is test infrastructure only).
//
// The reference's scalar is Zq{q,v} with mul = (u128)a*b % q, add = cond-subtract, sub = cond-add
// (arith/src/zq.rs:219-328).  All of those are exact operations in Z_q, so any exact evaluation
// strategy gives bit-identical canonical results; what must match is the *sequence of ring
// operations* (which twiddle multiplies which element), not the reduction technique.  Here:
//   * Lazy32  : q < 2^30, values kept in [0,4q) (forward) / [0,2q) (inverse), Shoup twiddles
//   * Lazy64  : q < 2^62, same scheme on 64-bit words (64x64->128 via __umul64hi)
//   * Strict64: 2^62 <= q < 2^63 (the reference's limit: Zq::add is an un-widened u64 add)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FHE_HD __host__ __device__ __forceinline__
#else
#define FHE_HD inline
#endif

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;

namespace fhe {

FHE_HD u32 mulhi_u32(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (u32)(((u64)a * b) >> 32);
#endif
}
FHE_HD u64 mulhi_u64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}
// Low 64 bits of y*w + h*nq, the tail of a 64-bit Shoup product (nq = 2^64 - q: the quotient estimate's multiple of q
// is ADDED).  Written out as two 32x32->64 multiply-adds and four 32-bit ones that chain through the accumulator: left
// to the compiler, y*w - h*q became two separate 64-bit products plus a negation and the additions that join their
// partial sums -- IMAD.IADD / IMAD.X, which execute on the same fmaheavy pipe as the multiplies (ncu, 62-bit polymul:
// fmaheavy 80 % active, the binding unit; 335 of its 2035 slots per thread were such additions and moves).
FHE_HD u64 shoup_tail64(u64 y, u64 w, u64 h, u64 nq) {
#if defined(__CUDA_ARCH__)
    u32 lo, hi;
    asm("{\n\t"
        ".reg .u64 t;\n\t"
        "mul.wide.u32 t, %2, %4;\n\t"
        "mad.wide.u32 t, %6, %8, t;\n\t"
        "mov.b64 {%0, %1}, t;\n\t"
        "mad.lo.u32 %1, %2, %5, %1;\n\t"
        "mad.lo.u32 %1, %3, %4, %1;\n\t"
        "mad.lo.u32 %1, %6, %9, %1;\n\t"
        "mad.lo.u32 %1, %7, %8, %1;\n\t"
        "}"
        : "=&r"(lo), "=&r"(hi)
        : "r"((u32)y), "r"((u32)(y >> 32)), "r"((u32)w), "r"((u32)(w >> 32)), "r"((u32)h), "r"((u32)(h >> 32)),
          "r"((u32)nq), "r"((u32)(nq >> 32)));
    return ((u64)hi << 32) | lo;
#else
    return y * w + h * nq;
#endif
}
FHE_HD u32 umin(u32 a, u32 b) { return a < b ? a : b; }
FHE_HD u64 umin(u64 a, u64 b) { return a < b ? a : b; }

// A twiddle w with its Shoup companion wp = floor(w * 2^wordbits / q).
// aligned to its size so that one vector load (LDG.64 / LDG.128) fetches both words; unaligned, the compiler
// issued two scalar loads per twiddle (124 of the 188 LDG of the N=1024 polymul)
#ifndef FHE_TW64_ALIGN
#define FHE_TW64_ALIGN 16
#endif
template <typename W> struct alignas(sizeof(W) == 8 ? FHE_TW64_ALIGN : 2 * sizeof(W)) Tw { W w, wp; };
typedef Tw<u32> Tw32;
typedef Tw<u64> Tw64;

// ---------------------------------------------------------------------------------------------------
// Lazy32: q < 2^30.
// ---------------------------------------------------------------------------------------------------
struct Lazy32 {
    typedef u32 W;
    typedef Tw32 T;
    static constexpr bool RADIX4 = false;  // see Fermat32
    u32 q, q2;       // q, 2q
    u32 bk_shift;    // k-1, with 2^(k-1) <= q < 2^k
    u32 bk_mu;       // floor(2^(2k) / q)  (< 2^(k+1))

    // w*y mod q in [0,2q) for ANY 32-bit y (Harvey / Shoup).
    FHE_HD u32 mul_tw(u32 y, T t) const { return y * t.w - mulhi_u32(y, t.wp) * q; }
    FHE_HD u32 csub(u32 x, u32 m) const { return umin(x, x - m); }  // [0,2m) -> [0,m)
    // forward (Cooley-Tukey) butterfly, arith/src/ntt.rs:56-60: U=x, V=y*S, x=U+V, y=U-V.
    // x in [0,4q), y any 32-bit word congruent to its value; outputs in [0,4q).
    FHE_HD void fwd(u32 &x, u32 &y, T t) const {
        u32 X = csub(x, q2);
        u32 V = mul_tw(y, t);
        x = X + V;
        y = X - V + q2;
    }
    // inverse (Gentleman-Sande) butterfly, arith/src/ntt.rs:92-97: x=U+V, y=(U-V)*S.  in/out [0,2q).
    FHE_HD void inv(u32 &x, u32 &y, T t) const {
        u32 s = csub(x + y, q2);
        u32 d = x - y + q2;
        x = s;
        y = mul_tw(d, t);
    }
    // last inverse stage with the n^-1 scaling (arith/src/ntt.rs:100-102) folded into both outputs:
    // x = (U+V)*ninv, y = (U-V)*(S*ninv).  in [0,2q), out [0,2q).
    FHE_HD void inv_last(u32 &x, u32 &y, T ninv, T s_ninv) const {
        u32 s = x + y;
        u32 d = x - y + q2;
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    // K = number of inverse stages already executed (only the csub-free Small32 policy needs it)
    template <int K> FHE_HD void inv_k(u32 &x, u32 &y, T t) const { inv(x, y, t); }
    template <int K> FHE_HD void inv_last_k(u32 &x, u32 &y, T ninv, T s_ninv) const { inv_last(x, y, ninv, s_ninv); }
    FHE_HD u32 canon4(u32 x) const { return csub(csub(x, q2), q); }  // [0,4q) -> [0,q)
    FHE_HD u32 canon2(u32 x) const { return csub(x, q); }            // [0,2q) -> [0,q)
    // a*b mod q for canonical a,b (Barrett on the 2k-bit product), result in [0,q).
    FHE_HD u32 mul(u32 a, u32 b) const {
        u64 x = (u64)a * b;
        u64 qh = ((x >> bk_shift) * bk_mu) >> (bk_shift + 2);
        u32 r = (u32)x - (u32)qh * q;  // in [0,3q)
        r = csub(r, q2);
        return csub(r, q);
    }
    // interface used by the polymul kernel (see Small32 for the non-trivial case)
    static constexpr bool PW_SCALED = false;           // pw_mul returns the true product
    FHE_HD u32 fwd_out(u32 x) const { return canon4(x); }   // forward output -> pointwise operand
    FHE_HD u32 fwd_canon(u32 x) const { return canon4(x); } // forward output -> canonical NTT value
    FHE_HD u32 pw_mul(u32 a, u32 b) const { return mul(a, b); }
    FHE_HD u32 pw_evals(u32 t) const { return t; }
    FHE_HD static u32 load(u64 v) { return (u32)v; }
    FHE_HD static u64 store(u32 v) { return (u64)v; }
};

// ---------------------------------------------------------------------------------------------------
// Lazy64: q < 2^62.
// ---------------------------------------------------------------------------------------------------
struct Lazy64 {
    typedef u64 W;
    typedef Tw64 T;
    static constexpr bool RADIX4 = false;
    u64 q, q2;
    u64 qinv_neg;  // -q^-1 mod 2^64 (kept for the plan's layout; the device uses qinv)
    u64 qinv;      //  q^-1 mod 2^64
    u64 r2;        // 2^128 mod q
    u64 nq;        // 2^64 - q

    FHE_HD u64 mul_tw(u64 y, T t) const { return shoup_tail64(y, t.w, mulhi_u64(y, t.wp), nq); }
    FHE_HD u64 csub(u64 x, u64 m) const { return umin(x, x - m); }
    FHE_HD void fwd(u64 &x, u64 &y, T t) const {
        u64 X = csub(x, q2);
        u64 V = mul_tw(y, t);
        x = X + V;
        y = X - V + q2;
    }
    FHE_HD void inv(u64 &x, u64 &y, T t) const {
        u64 s = csub(x + y, q2);
        u64 d = x - y + q2;
        x = s;
        y = mul_tw(d, t);
    }
    FHE_HD void inv_last(u64 &x, u64 &y, T ninv, T s_ninv) const {
        u64 s = x + y;
        u64 d = x - y + q2;
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    // K = number of inverse stages already executed (only the csub-free Small32 policy needs it)
    template <int K> FHE_HD void inv_k(u64 &x, u64 &y, T t) const { inv(x, y, t); }
    template <int K> FHE_HD void inv_last_k(u64 &x, u64 &y, T ninv, T s_ninv) const { inv_last(x, y, ninv, s_ninv); }
    FHE_HD u64 canon4(u64 x) const { return csub(csub(x, q2), q); }
    FHE_HD u64 canon2(u64 x) const { return csub(x, q); }
    // Montgomery product a*b*2^-64 mod q in [0,q) (a*b < q*2^64).
    // subtractive form: m*q has the low word of a*b, so a*b - m*q = (hi - mulhi(m,q)) * 2^64 exactly; hi, mulhi < q
    FHE_HD u64 mont(u64 a, u64 b) const {
        u64 lo = a * b, hi = mulhi_u64(a, b);
        u64 m = lo * qinv;
        u64 t = hi - mulhi_u64(m, q) + q;  // (0, 2q)
        return csub(t, q);
    }
    FHE_HD u64 mul(u64 a, u64 b) const { return mont(mont(a, b), r2); }
    // Pointwise product of the polymul: ONE Montgomery product; its 2^-64 is folded into the n^-1 constants of the
    // inverse transform (ninv_pw / s_ninv_pw), as Small32 does.  Operands only need to be below 2q (4q^2 < q * 2^64
    // for q < 2^62), so the forward output takes one conditional subtraction instead of two, and the product stays in
    // (0, 2q), which is what the inverse butterflies accept: no final subtraction either.
    static constexpr bool PW_SCALED = true;
    FHE_HD u64 fwd_out(u64 x) const { return csub(x, q2); }
    FHE_HD u64 fwd_canon(u64 x) const { return canon4(x); }
    FHE_HD u64 pw_mul(u64 a, u64 b) const {
        u64 lo = a * b, hi = mulhi_u64(a, b);
        u64 m = lo * qinv;
        return hi - mulhi_u64(m, q) + q;  // (0, 2q), congruent to a*b*2^-64
    }
    FHE_HD u64 pw_evals(u64 t) const { return mont(t, r2); }  // canonical a*b (t < 2q, r2 < q: t*r2 < q * 2^64)
    FHE_HD static u64 load(u64 v) { return v; }
    FHE_HD static u64 store(u64 v) { return v; }
};

// ---------------------------------------------------------------------------------------------------
// Strict64: 2^62 <= q < 2^63; everything canonical after every operation.
// ---------------------------------------------------------------------------------------------------
struct Strict64 {
    typedef u64 W;
    typedef Tw64 T;
    static constexpr bool RADIX4 = false;
    u64 q, q2;  // q2 unused
    u64 qinv_neg;
    u64 qinv;   // q^-1 mod 2^64
    u64 r2;
    u64 nq;     // 2^64 - q

    FHE_HD u64 mul_tw(u64 y, T t) const {  // canonical result
        u64 r = shoup_tail64(y, t.w, mulhi_u64(y, t.wp), nq);  // [0,2q), 2q < 2^64
        return r >= q ? r - q : r;
    }
    FHE_HD u64 add(u64 a, u64 b) const { u64 s = a + b; return s >= q ? s - q : s; }
    FHE_HD u64 sub(u64 a, u64 b) const { return a >= b ? a - b : a + q - b; }
    FHE_HD void fwd(u64 &x, u64 &y, T t) const {
        u64 V = mul_tw(y, t);
        u64 X = x;
        x = add(X, V);
        y = sub(X, V);
    }
    FHE_HD void inv(u64 &x, u64 &y, T t) const {
        u64 s = add(x, y);
        u64 d = sub(x, y);
        x = s;
        y = mul_tw(d, t);
    }
    FHE_HD void inv_last(u64 &x, u64 &y, T ninv, T s_ninv) const {
        u64 s = add(x, y);
        u64 d = sub(x, y);
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    // K = number of inverse stages already executed (only the csub-free Small32 policy needs it)
    template <int K> FHE_HD void inv_k(u64 &x, u64 &y, T t) const { inv(x, y, t); }
    template <int K> FHE_HD void inv_last_k(u64 &x, u64 &y, T ninv, T s_ninv) const { inv_last(x, y, ninv, s_ninv); }
    FHE_HD u64 canon4(u64 x) const { return x; }
    FHE_HD u64 canon2(u64 x) const { return x; }
    FHE_HD u64 mont(u64 a, u64 b) const {
        u64 lo = a * b, hi = mulhi_u64(a, b);
        u64 m = lo * qinv;
        u64 t = hi - mulhi_u64(m, q) + q;  // subtractive form (see Lazy64::mont): (0, 2q), 2q < 2^64
        return t >= q ? t - q : t;
    }
    FHE_HD u64 mul(u64 a, u64 b) const { return mont(mont(a, b), r2); }
    static constexpr bool PW_SCALED = true;  // one Montgomery product; 2^-64 folded into ninv_pw (see Lazy64)
    FHE_HD u64 fwd_out(u64 x) const { return x; }
    FHE_HD u64 fwd_canon(u64 x) const { return x; }
    FHE_HD u64 pw_mul(u64 a, u64 b) const { return mont(a, b); }
    FHE_HD u64 pw_evals(u64 t) const { return mont(t, r2); }
    FHE_HD static u64 load(u64 v) { return v; }
    FHE_HD static u64 store(u64 v) { return v; }
};

// ---------------------------------------------------------------------------------------------------
// Small32: q < 2^22 and 2q * n <= 2^32 (the reference's only NTT modulus, 65537, lives here for every n <= 2^14).  With so much headroom in a
// 32-bit word the forward butterfly needs no conditional subtraction at all: x' = x + V, y' = x - V + 2q
// grows by at most 2q per stage, so after LOGN <= 15 stages values stay below 31q < 2^27.  The pointwise
// product of two such lazy values (a*b < 961 q^2 < q*2^32) is one Montgomery reduction, whose 2^-32 factor
// is folded into the n^-1 constants of the inverse transform (ninv_pw / s_ninv_pw in the plan).
// ---------------------------------------------------------------------------------------------------
struct Small32 {
    typedef u32 W;
    typedef Tw32 T;
    static constexpr bool RADIX4 = false;
    u32 q, q2;
    u32 qinv_neg;  // -q^-1 mod 2^32
    u32 qinv;      //  q^-1 mod 2^32
    Tw32 one;      // w = 1           : mul_tw(x, one) = x mod q in [0,2q) for any 32-bit x
    Tw32 r;        // w = 2^32 mod q  : undoes the Montgomery factor
    u32 qk[16];    // 2q << K: offset of the K-th executed inverse stage (read from the constant bank by IADD3)
    static constexpr bool PW_SCALED = true;

    FHE_HD u32 mul_tw(u32 y, T t) const { return y * t.w - mulhi_u32(y, t.wp) * q; }
    FHE_HD u32 csub(u32 x, u32 m) const { return umin(x, x - m); }
    FHE_HD void fwd(u32 &x, u32 &y, T t) const {
        u32 V = mul_tw(y, t);
        y = x - V + q2;
        x = x + V;
    }
    FHE_HD void inv(u32 &x, u32 &y, T t) const {
        u32 s = csub(x + y, q2);
        u32 d = x - y + q2;
        x = s;
        y = mul_tw(d, t);
    }
    FHE_HD void inv_last(u32 &x, u32 &y, T ninv, T s_ninv) const {
        u32 s = x + y;
        u32 d = x - y + q2;
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    // csub-free inverse butterflies: the sum path doubles its bound every stage (values < 2q * 2^K before the
    // (K+1)-th executed stage), the difference is offset by that bound (a multiple of q) and goes through mul_tw,
    // which takes any 32-bit word.  Needs 2q * 2^LOGN <= 2^32: plan_host.hpp: modulus_kind() only selects this
    // policy when that holds.
    template <int K> FHE_HD void inv_k(u32 &x, u32 &y, T t) const {
        u32 s = x + y;
        u32 d = x - y + qk[K];
        x = s;
        y = mul_tw(d, t);
    }
    template <int K> FHE_HD void inv_last_k(u32 &x, u32 &y, T ninv, T s_ninv) const {
        u32 s = x + y;
        u32 d = x - y + qk[K];
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    FHE_HD u32 canon2(u32 x) const { return csub(x, q); }
    FHE_HD u32 fwd_out(u32 x) const { return x; }
    FHE_HD u32 fwd_canon(u32 x) const { return csub(mul_tw(x, one), q); }
    // a*b*2^-32 mod q in (0,2q); needs a*b < q*2^32.  Subtractive Montgomery form: m = lo * q^-1, so m*q has the
    // same low word as a*b and a*b - m*q = (hi - mulhi(m,q)) * 2^32 exactly; hi, mulhi(m,q) < q, and adding q makes
    // the difference positive -- one IADD3 instead of the carry test (lo != 0) of the additive form.
    FHE_HD u32 pw_mul(u32 a, u32 b) const {
        u64 p = (u64)a * b;
        u32 lo = (u32)p, hi = (u32)(p >> 32);
        u32 m = lo * qinv;
        return hi - mulhi_u32(m, q) + q;
    }
    FHE_HD u32 pw_evals(u32 t) const { return csub(mul_tw(t, r), q); }
    FHE_HD u32 mul(u32 a, u32 b) const { return pw_evals(pw_mul(a, b)); }  // canonical a*b (a*b < q*2^32)
    FHE_HD static u32 load(u64 v) { return (u32)v; }
    FHE_HD static u64 store(u32 v) { return (u64)v; }
};

// ---------------------------------------------------------------------------------------------------
// Fermat32: q = 2^16 + 1 = 65537, the reference's only NTT modulus (every parameter set of SURVEY appendix B).
// Small32 with RADIX-4 butterflies.  The twiddle tables satisfy roots[2i+1] = roots[2i] * I with
// I = psi^(n/2), I^2 = -1 (bit reversal: brev(2i+1) = brev(2i) | n/2), and the reference's root search gives
// psi = 3^(32768/n) for this modulus, hence I = 3^16384 = -2^8 for every n.  Two consecutive Cooley-Tukey stages on the
// four positions (e0, e1 = e0+h, e2 = e0+2h, e3 = e0+3h) -- parent twiddle S1 = roots[i], children S2a = roots[2i],
// S2b = roots[2i+1] = I*S2a -- are, in the reference's order of operations,
//      a0 = x0 + S1 x2, a2 = x0 - S1 x2, a1 = x1 + S1 x3, a3 = x1 - S1 x3,
//      y0 = a0 + S2a a1, y1 = a0 - S2a a1, y2 = a2 + S2b a3, y3 = a2 - S2b a3.
// With p1 = S2a x1, p2 = S1 x2, p3 = (S1 S2a) x3:  S2a a1 = p1 + p3 and S2b a3 = I (p1 - p3) = (p3 - p1) << 8:
// THREE twiddle products instead of four, the fourth being a shift.  All of it is exact arithmetic in Z_q, so the
// canonical results are those of the reference's loop (arith/src/ntt.rs:44-110) bit for bit; what changes is the number of
// multiplier-pipe (fmaheavy) slots: 9 instead of 12 per four butterflies (ncu: that pipe is the kernel's binding unit at
// 81 % active), for 9 instead of 8 additions on the half-idle ALU pipe.  The device tables keep S1*S2a in the slot of the
// now unused roots[2i+1] (plan_host.hpp: radix4_patch).
//
// Ranges.  Forward: the shifted term is below 4q * 256 = 1024q, so the never-multiplied x path grows by at most 1026q per
// radix-4 layer: < 2^29 after the 7 layers of n = 2^15; every product takes any 32-bit word.  The pointwise product folds
// ONE operand below 2q first (fold: v = lo16 - hi16 + q, three ALU instructions), the other stays lazy: 2^17 * 2^29 < q * 2^32.
// Inverse: inputs below 2q * 2^KB; the shifted difference is below q << (KB + 10), which needs KB <= 4 to stay inside a
// word; the only output that grows is the plain sum y0 (below 2q * 2^(KB+2)), and InvSched (ntt_core.cuh) folds it where
// the next radix-4 layer would otherwise be entered with KB > 4.
// ---------------------------------------------------------------------------------------------------
struct Fermat32 : Small32 {
    static constexpr bool RADIX4 = true;
    static constexpr int INV_KB_MAX = 4;
    u32 q4;      // 4q
    u32 c10;     // 1024q = (4q) << 8
    u32 okb[8];  // q << (KB + 10): offset of the shifted difference of an inverse radix-4 block entered at KB
    u32 c8, c14; // 512q, 16384q: offsets of the shift-only first layer (fwd4_first)
#ifndef FHE_FERMAT_FIRST_SHIFT
#define FHE_FERMAT_FIRST_SHIFT 1
#endif
    static constexpr bool FIRST_SHIFT = FHE_FERMAT_FIRST_SHIFT != 0;
#ifndef FHE_FERMAT_COMPACT_MIN   // smallest log2 n whose later passes read 4-byte twiddles (ntt_core.cuh: tw_load)
#define FHE_FERMAT_COMPACT_MIN 12
#endif
    FHE_HD static constexpr bool compact(int logn) { return logn >= FHE_FERMAT_COMPACT_MIN; }

    FHE_HD u32 fold(u32 v) const { return (v & 0xffffu) - (v >> 16) + q; }  // any 32-bit word -> [2, 2^17], same residue
#ifndef FHE_FERMAT_SHIFTQ
#define FHE_FERMAT_SHIFTQ 1
#endif
    // Shoup product with the multiple of q = 2^16 + 1 taken off by a shift and a three-input add instead of a second
    // low product (ptxas folds the subtraction of h << 16 into a LEA: same instruction count, 6 % fewer multiplier-pipe cycles; measured 229.3 -> 233.0 M polymul/s at N = 1024, +1..2 % at every degree; -DFHE_FERMAT_SHIFTQ=0 restores the two-product form).
    FHE_HD u32 mul_tw(u32 y, T t) const {
#if FHE_FERMAT_SHIFTQ
        const u32 h = mulhi_u32(y, t.wp);
        return y * t.w - h - (h << 16);
#else
        return Small32::mul_tw(y, t);
#endif
    }
    FHE_HD void fwd(u32 &x, u32 &y, T t) const {
        const u32 V = mul_tw(y, t);
        y = x - V + q2;
        x = x + V;
    }
    template <int K> FHE_HD void inv_k(u32 &x, u32 &y, T t) const {
        const u32 s = x + y, d = x - y + qk[K];
        x = s;
        y = mul_tw(d, t);
    }
    template <int K> FHE_HD void inv_last_k(u32 &x, u32 &y, T ninv, T s_ninv) const {
        const u32 s = x + y, d = x - y + qk[K];
        x = mul_tw(s, ninv);
        y = mul_tw(d, s_ninv);
    }
    FHE_HD void fwd4(u32 &x0, u32 &x1, u32 &x2, u32 &x3, T s1, T s2a, T s12) const {
        const u32 p2 = mul_tw(x2, s1), p1 = mul_tw(x1, s2a), p3 = mul_tw(x3, s12);  // each in [0, 2q)
        const u32 t0 = x0 + p2, t2 = x0 - p2 + q2;
        const u32 s = p1 + p3;        // S2a*a1            in [0, 4q)
        const u32 dn = p3 - p1 + q2;  // -(S2a*a3) + 2q    in (0, 4q)
        const u32 D = dn << 8;        // I*S2a*a3 = -256 (p1 - p3) = 256 dn (mod q), below 1024q
        x0 = t0 + s;
        x1 = t0 - s + q4;
        x2 = t2 + D;
        x3 = t2 - D + c10;
    }
    // The FIRST radix-4 layer of a forward transform (stages 1 and 2 of arith/src/ntt.rs:44-73, one block: H = hi = 0).
    // Its twiddles are the same powers of two for every n -- roots[1] = I = -2^8, roots[2] = 2^12, roots[3] = I*roots[2] =
    // 2^4 (psi^(n/4) = 3^8192; plan_host.hpp: fermat_ok checks all three) -- and its inputs are canonical (<= 2^16, the
    // precondition of every Rq entry point; the offsets below are sized for any 17-bit word, so that an unreduced
    // 65537..131071 still gives the right residue), so the three products are shifts that need no reduction:
    //      y0 = x0 - (x2<<8) + (x1<<12) + (x3<<4)        y2 = x0 + (x2<<8) + (x1<<4) + (x3<<12)
    //      y1 = x0 - (x2<<8) - (x1<<12) - (x3<<4)        y3 = x0 + (x2<<8) - (x1<<4) - (x3<<12)
    // (I*(p1 - p3) = -2^8 (x1<<12 - x3<<4) = x1<<4 + x3<<12 because 2^20 = -2^4.)  Eleven ALU instructions and no
    // multiplier-pipe slot instead of nine of each.  Outputs stay below 2^17 + 2^25 + c14 < 2^30.1; the later layers add
    // at most 1026q each to the never-multiplied path (< 2^30.6 after the 6 of n = 2^15), and the pointwise product takes
    // fold(a) <= 2^17 times that: below q * 2^32.
    FHE_HD void fwd4_first(u32 &x0, u32 &x1, u32 &x2, u32 &x3) const {
        const u32 a = x2 << 8;
        const u32 t0 = x0 - a + c8, t2 = x0 + a;  // c8 = 512q >= 2^25
        const u32 s = (x1 << 12) + (x3 << 4);
        const u32 r = (x3 << 12) + (x1 << 4);
        x0 = t0 + s;
        x1 = t0 - s + c14;  // c14 = 16384q >= 2^29 + 2^21
        x2 = t2 + r;
        x3 = t2 - r + c14;
    }
    // Two Gentleman-Sande stages (children first: Sa = roots_inv[2i], Sb = roots_inv[2i+1] = Sa / I = 256 Sa; then the
    // parent S1 = roots_inv[i]):  y0 = x0+x1+x2+x3, y1 = Sa (d01 + 256 d23), y2 = S1 (x0+x1-x2-x3), y3 = S1 Sa (d01 - 256 d23).
    // Inputs below 2q << KB (qk[KB]).
    template <int KB, bool FOLD> FHE_HD void inv4(u32 &x0, u32 &x1, u32 &x2, u32 &x3, T sa, T s1sa, T s1) const {
        static_assert(KB >= 0 && KB <= INV_KB_MAX, "Fermat32::inv4: the shifted difference would leave the word");
        const u32 s01 = x0 + x1, s23 = x2 + x3;
        const u32 d01 = x0 - x1 + qk[KB], d23 = x2 - x3 + qk[KB];  // (0, 2 * 2q << KB)
        const u32 E = d23 << 8;                                    // below q << (KB + 10)
        const u32 u = d01 + E, v = d01 - E + okb[KB];
        const u32 w = s01 - s23 + qk[KB + 1];
        const u32 y0 = s01 + s23;
        x0 = FOLD ? fold(y0) : y0;
        x1 = mul_tw(u, sa);
        x2 = mul_tw(w, s1);
        x3 = mul_tw(v, s1sa);
    }
    FHE_HD u32 fwd_canon(u32 x) const { return csub(fold(x), q); }
    FHE_HD u32 pw_mul(u32 a, u32 b) const { return Small32::pw_mul(fold(a), b); }
    FHE_HD u32 mul(u32 a, u32 b) const { return pw_evals(Small32::pw_mul(a, b)); }  // canonical operands
};

}  // namespace fhe
