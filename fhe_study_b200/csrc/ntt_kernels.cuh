// ntt_kernels.cuh -- batched negacyclic NTT / INTT / R_q polymul kernels (one polynomial per group of
// T = N/E threads, E coefficients per thread in registers, one shared-memory exchange between
// register-local passes).  Included by one translation unit per modular policy (ntt_inst_*.cu).
//
// Follows arith/src/ntt.rs:44-110 (NTT::ntt / NTT::intt) and arith/src/ring_nq.rs:564-607
// (mul / mul_mut: A = evals or ntt(a), B = evals or ntt(b), C = A.B pointwise, c = intt(C), the
// result keeps C as its cached evals).
#pragma once
#include <stdlib.h>

#include <atomic>

#include "common.cuh"
#include "ntt_core.cuh"
#include "tc_common.cuh"   // mbarrier / cp.async.bulk wrappers (the TMA-staged polymul below)

namespace fhe {

// Device-visible plan for one (q, n).  Passed BY VALUE (__grid_constant__): c_fwd / c_inv (the first
// 2^g0 <= 64 table entries, all that pass 0 needs) are then read straight from the constant bank.
template <class M> struct NttParams {
    M mod;
    const typename M::T *fwd;   // n entries: roots[i]     = psi^bitrev(i)      (arith/src/ntt.rs:133-147)
    const typename M::T *inv;   // n entries: roots_inv[i] = roots[i]^-1        (arith/src/ntt.rs:149-161)
    typename M::T ninv;         // n^-1                                        (arith/src/ntt.rs:27-30)
    typename M::T s_ninv;       // roots_inv[1] * n^-1
    typename M::T ninv_pw, s_ninv_pw;  // same, times the scale factor M::pw_mul leaves behind (polymul only)
    typename M::T c_fwd[64], c_inv[64];
    const u32 *fwdw, *invw;     // radix-4 policies: 4-byte twiddle tables of the large degrees (ntt_core.cuh: tw_load), else null
};

// MODE_MUL : polymul with run-time evals flags, both operands through ONE copy of the forward code (one after the other)
// MODE_MUL2: polymul of two operands in coefficient form, transformed TOGETHER (fwd_pass2: shared twiddle fetches,
//            two independent butterfly streams, both operands' global loads in flight at once)
// MODE_MULG: MODE_MUL with NTT(a) parked in the product's own output row (global memory, L2-resident for the few
//            microseconds until the pointwise product reads it back) instead of E live registers: at N = 16384 that
//            is the difference between one 512-thread CTA of ~116 registers and two of 64 per SM
enum NttMode { MODE_FWD = 0, MODE_INV = 1, MODE_MUL = 2, MODE_MUL2 = 3, MODE_MULS = 4, MODE_MULG = 5 };
__host__ __device__ constexpr bool is_mul_mode(int mode) { return mode >= MODE_MUL; }
enum MulFlags { A_IS_EVALS = 1, B_IS_EVALS = 2, B_BROADCAST = 4, STAGE_TMA = 64 /* internal: MODE_MULS with bulk copies */ };  // B_BROADCAST: b is ONE polynomial, used for every product

template <int LOGN, int LOGE> struct KernelGeom {
    typedef NttShape<LOGN, LOGE> S;
#ifndef FHE_NTT_MIN_CT
#define FHE_NTT_MIN_CT 128
#endif
    static constexpr int CT = S::T > FHE_NTT_MIN_CT ? S::T : FHE_NTT_MIN_CT;   // threads per CTA
    static constexpr int PPC = CT / S::T;                // polynomials per CTA
};

// Shared-memory padding of the exchanges.  Default: one pad word per 128 bytes.  32-bit words with 32 coefficients per
// thread (and a last pass of five stages) pad FOUR words per 32 instead: in the layout of the last pass (nL = 0) a thread then owns 32 consecutive words
// starting at a multiple of 36 -- 16-byte aligned, and eight such rows land on 32 distinct banks (36 t mod 32 = 4 t) -- so
// that side of an exchange is eight 128-bit accesses instead of thirty-two 32-bit ones.  The other layouts keep their 32
// lanes inside one 32-word row (nL >= 5), where any per-row constant pad is conflict-free.
#ifndef FHE_NTT_VEC_EXCH
#define FHE_NTT_VEC_EXCH 1
#endif
template <class M, int LOGN, int LOGE> struct PadRule {
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int WB = (int)sizeof(typename M::W), SH = WB == 4 ? 5 : 4;
    static constexpr bool VEC = FHE_NTT_VEC_EXCH != 0 && WB == 4 && LOGE == 5 && S::P >= 2 && S::g(S::P - 1) == LOGE &&
                                S::nL(S::P - 2) >= 5;
    static constexpr int K = VEC ? 4 : 1;
    __host__ __device__ static constexpr int idx(int i) { return i + K * (i >> SH); }
    static constexpr int padn = S::N + K * (S::N >> SH);
    // pass whose layout gives every thread E consecutive words
    __host__ __device__ static constexpr bool row_layout(int p) { return VEC && S::nL(p) == 0 && S::g(p) == LOGE; }
};

// the rule for 4-byte words, for the fused kernels that keep transform outputs in the exchange slots (extprod_fused.cu,
// tn_fused.cu, glwe_rq.cu): the policy only contributes its word size
template <int LOGN, int LOGE> using Pad32 = PadRule<Lazy32, LOGN, LOGE>;

template <int T> __device__ __forceinline__ void group_sync() {
    if (T <= 32) __syncwarp(); else __syncthreads();
}

// An exchange between ADJACENT passes a < b touches, per thread, a closed set of partners: the threads that share
// the s0(a) high bits H and the nL(b) low bits of the butterfly-group index.  When g(a) == g(b) the g bits that
// vary sit at [nL(b), nL(a)) in both layouts; with nL(a) <= 5 they are lane bits, so every partner is in the same
// warp and __syncwarp orders the exchange (always the case for the LAST two passes once they are equal).
template <int LOGN, int LOGE, int FROM, int TO> struct ExchScope {
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int A = FROM < TO ? FROM : TO, B = FROM < TO ? TO : FROM;
    static constexpr bool in_warp = S::T <= 32 || (B == A + 1 && S::g(A) == S::g(B) && S::nL(A) <= 5 && S::T % 32 == 0);
};

// one side of an exchange: the thread's E registers <-> its words of layout P (base = the thread's first word)
template <class M, int LOGN, int LOGE, int P>
__device__ __forceinline__ void exch_put(const typename M::W (&x)[1 << LOGE], typename M::W *base) {
    typedef NttShape<LOGN, LOGE> S;
    typedef PadRule<M, LOGN, LOGE> PR;
    if constexpr (PR::row_layout(P)) {
        uint4 *v = reinterpret_cast<uint4 *>(base);
#pragma unroll
        for (int e = 0; e < S::E; e += 4) v[e >> 2] = make_uint4(x[e], x[e + 1], x[e + 2], x[e + 3]);
    } else {
#pragma unroll
        for (int e = 0; e < S::E; e++) base[PR::idx(S::pos(P, 0, e))] = x[e];
    }
}
template <class M, int LOGN, int LOGE, int P>
__device__ __forceinline__ void exch_get(typename M::W (&x)[1 << LOGE], const typename M::W *base) {
    typedef NttShape<LOGN, LOGE> S;
    typedef PadRule<M, LOGN, LOGE> PR;
    if constexpr (PR::row_layout(P)) {
        const uint4 *v = reinterpret_cast<const uint4 *>(base);
#pragma unroll
        for (int e = 0; e < S::E; e += 4) {
            const uint4 q = v[e >> 2];
            x[e] = q.x; x[e + 1] = q.y; x[e + 2] = q.z; x[e + 3] = q.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = base[PR::idx(S::pos(P, 0, e))];
    }
}

// registers (layout of pass FROM) -> shared -> registers (layout of pass TO).
// LEAD = false drops the barrier in front of the writes: allowed when this thread's previous access to `sm` was
// the read side of an exchange INTO layout FROM (it then overwrites exactly the words it read itself, so there is
// no other reader to wait for) -- i.e. for every exchange of a chain but the first.
template <class M, int LOGN, int LOGE, int FROM, int TO, bool LEAD = true>
__device__ __forceinline__ void exchange(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
#ifdef FHE_NTT_FULL_SYNC  // the conservative scheme (two CTA-wide barriers per exchange), kept for A/B measurements
    group_sync<S::T>();
#else
    if constexpr (LEAD) group_sync<S::T>();  // earlier readers of sm are done
#endif
    // pos(p, tid, e) = pos(p, tid, 0) | pos(p, 0, e) on disjoint bits (tid < T = 2^k and e >> g select different bits
    // of the group index), and i + (i >> 5) is additive over disjoint bits: thread base + compile-time offset, so
    // every STS/LDS below takes an immediate offset instead of per-element LOP3/LEA address arithmetic.
    typedef PadRule<M, LOGN, LOGE> PR;
    const int bf = PR::idx(S::pos(FROM, tid, 0)), bt = PR::idx(S::pos(TO, tid, 0));
    exch_put<M, LOGN, LOGE, FROM>(x, sm + bf);
#ifdef FHE_NTT_FULL_SYNC
    group_sync<S::T>();
#else
    if constexpr (ExchScope<LOGN, LOGE, FROM, TO>::in_warp) __syncwarp(); else __syncthreads();
#endif
    exch_get<M, LOGN, LOGE, TO>(x, sm + bt);
}

// the same for two register sets through two buffers: one barrier orders both
template <class M, int LOGN, int LOGE, int FROM, int TO, bool LEAD = true>
__device__ __forceinline__ void exchange2(typename M::W (&x)[1 << LOGE], typename M::W (&y)[1 << LOGE], typename M::W *smx,
                                          typename M::W *smy, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (LEAD) group_sync<S::T>();
    typedef PadRule<M, LOGN, LOGE> PR;
    const int bf = PR::idx(S::pos(FROM, tid, 0)), bt = PR::idx(S::pos(TO, tid, 0));
    exch_put<M, LOGN, LOGE, FROM>(x, smx + bf);
    exch_put<M, LOGN, LOGE, FROM>(y, smy + bf);
    if constexpr (ExchScope<LOGN, LOGE, FROM, TO>::in_warp) __syncwarp(); else __syncthreads();
    exch_get<M, LOGN, LOGE, TO>(x, smx + bt);
    exch_get<M, LOGN, LOGE, TO>(y, smy + bt);
}
template <class M, int LOGN, int LOGE, int PASS = 0, bool FIRST = true>
__device__ __forceinline__ void fwd_chain2(typename M::W (&x)[1 << LOGE], typename M::W (&y)[1 << LOGE], typename M::W *smx,
                                           typename M::W *smy, int tid, const M &m, const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (PASS > 0) exchange2<M, LOGN, LOGE, PASS - 1, PASS, FIRST>(x, y, smx, smy, tid);
    fwd_pass2<M, LOGN, LOGE, PASS>(x, y, tid, m, tw);
    if constexpr (PASS + 1 < S::P) fwd_chain2<M, LOGN, LOGE, PASS + 1, false>(x, y, smx, smy, tid, m, tw);
}

template <class M, int LOGN, int LOGE, int PASS = 0, bool FIRST = true>
__device__ __forceinline__ void fwd_chain(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid, const M &m,
                                          const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (PASS > 0) exchange<M, LOGN, LOGE, PASS - 1, PASS, FIRST>(x, sm, tid);
    fwd_pass<M, LOGN, LOGE, PASS>(x, tid, m, tw);
    if constexpr (PASS + 1 < S::P) fwd_chain<M, LOGN, LOGE, PASS + 1, (FIRST && PASS == 0)>(x, sm, tid, m, tw);
}
template <class M, int LOGN, int LOGE, int PASS, bool FIRST = true>
__device__ __forceinline__ void inv_chain(typename M::W (&x)[1 << LOGE], typename M::W *sm, int tid, const M &m,
                                          const TwSrc<M> &tw, typename M::T ninv, typename M::T s_ninv) {
    inv_pass<M, LOGN, LOGE, PASS>(x, tid, m, tw, ninv, s_ninv);
    if constexpr (PASS > 0) {
        exchange<M, LOGN, LOGE, PASS, PASS - 1, FIRST>(x, sm, tid);
        inv_chain<M, LOGN, LOGE, PASS - 1, false>(x, sm, tid, m, tw, ninv, s_ninv);
    }
}

// Polynomial data is touched once.  STREAM reads it without allocating in L1 and stores it evict-first, so that L1
// keeps the twiddle tables (at N >= 4096 the per-thread twiddles of the last pass are as many bytes as the
// polynomial).  Measured (q = 65537): polymul N=4096/8192/16384 +1/+2/+3 %, but the transform-only kernels LOSE
// 2..7 % -- hence on for the polymul at N >= 4096 only (ntt_stream()).
#ifndef FHE_NTT_STREAM
#define FHE_NTT_STREAM 1
#endif
__host__ __device__ constexpr bool ntt_stream(int logn, int mode) { return FHE_NTT_STREAM && is_mul_mode(mode) && logn >= 12; }
// Global I/O word: u64 (the layout of SURVEY 8b: one 8-byte word per coefficient) or u32 (packed wire / device
// format for q <= 2^32: half the HBM and PCIe bytes).  A warp-wide access is lane-contiguous in both (128 or 256 B).
template <bool STREAM> __device__ __forceinline__ u64 ld_poly(const u64 *p) {
    if constexpr (STREAM) {
        u64 v;
        asm("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
        return v;
    } else {
        return __ldg(p);
    }
}
template <bool STREAM> __device__ __forceinline__ u32 ld_poly(const u32 *p) {
    if constexpr (STREAM) {
        u32 v;
        asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
    } else {
        return __ldg(p);
    }
}
template <bool STREAM> __device__ __forceinline__ void st_poly(u64 *p, u64 v) {
    if constexpr (STREAM) __stcs(reinterpret_cast<unsigned long long *>(p), (unsigned long long)v);
    else *p = v;
}
template <bool STREAM> __device__ __forceinline__ void st_poly(u32 *p, u32 v) {
    if constexpr (STREAM) __stcs(p, v);
    else *p = v;
}

// global (coalesced, layout of pass 0) -> registers in the layout of pass TO
template <class M, int LOGN, int LOGE, int TO, bool STREAM = false, typename IOW = u64>
__device__ __forceinline__ void load_poly(typename M::W (&x)[1 << LOGE], const IOW *__restrict__ g, bool valid,
                                          typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    const IOW *gt = g + S::pos(0, tid, 0);
#pragma unroll
    for (int e = 0; e < S::E; e++) x[e] = valid ? (typename M::W)ld_poly<STREAM>(gt + S::pos(0, 0, e)) : (typename M::W)0;
    if constexpr (TO != 0) exchange<M, LOGN, LOGE, 0, TO>(x, sm, tid);
}
// registers in the layout of pass FROM -> global (coalesced).  LEAD as in exchange().
template <class M, int LOGN, int LOGE, int FROM, bool LEAD = true, bool STREAM = false, typename IOW = u64>
__device__ __forceinline__ void store_poly(typename M::W (&x)[1 << LOGE], IOW *__restrict__ g, bool valid,
                                           typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    if constexpr (FROM != 0) exchange<M, LOGN, LOGE, FROM, 0, LEAD>(x, sm, tid);
    if (valid) {
        IOW *gt = g + S::pos(0, tid, 0);
#pragma unroll
        for (int e = 0; e < S::E; e++) st_poly<STREAM>(gt + S::pos(0, 0, e), (IOW)x[e]);
    }
}

// ---- bit-packed global format (the PCIe wire of the host-buffer path) ----------------------------------------------
// A polynomial is n * bits / 32 consecutive u32 words, coefficient i in bits [i*bits, (i+1)*bits), little endian,
// 16 <= bits <= 31 (q = 65537: 17 bits instead of 32 or 64 per coefficient over PCIe).  Only for shapes whose
// pass-0 layout gives the 32 lanes of a warp 32 consecutive coefficients per register slot (T a multiple of 32 and
// nL(0) >= 5: every n >= 1024): then a warp's slot-e fields form ONE word-aligned block of `bits` words, the field
// shift (bits * lane) & 31 is the same for all of a thread's slots, a load is two lane-contiguous LDG.32 and a funnel
// shift, and a store is assembled from <= 3 neighbouring lanes with warp shuffles.
struct pk32 { u32 v; };  // tag type: u32 words of the packed format
template <typename IOW> struct IoTraits { static constexpr bool packed = false; };
template <> struct IoTraits<pk32> { static constexpr bool packed = true; };
// words from the start of the batch to polynomial `poly`
template <typename IOW, int N> __device__ __forceinline__ size_t row_offset(size_t poly, int bits) {
    if constexpr (IoTraits<IOW>::packed) return poly * (size_t)((N >> 5) * bits);
    else return poly * (size_t)N;
}
template <int LOGN, int LOGE> __host__ __device__ constexpr bool packed_shape_ok() {
    typedef NttShape<LOGN, LOGE> S;
    return S::T % 32 == 0 && S::nL(0) >= 5;
}
template <class M, int LOGN, int LOGE, int TO, bool STREAM>
__device__ __forceinline__ void load_poly_packed(typename M::W (&x)[1 << LOGE], const pk32 *__restrict__ g, int bits,
                                                 typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    static_assert(packed_shape_ok<LOGN, LOGE>(), "packed format: unsupported shape");
    const u32 bo = (u32)bits * (u32)tid, sh = bo & 31u, mask = (1u << bits) - 1u;
    const u32 *gw = reinterpret_cast<const u32 *>(g) + (bo >> 5);
    const bool two = sh + (u32)bits > 32u;  // the field straddles a word boundary (thread constant)
#pragma unroll
    for (int e = 0; e < S::E; e++) {
        const u32 w = (u32)(S::pos(0, 0, e) >> 5) * (u32)bits;  // pos(0,0,e) is a multiple of 32
        const u32 lo = __ldg(gw + w), hi = two ? __ldg(gw + w + 1) : 0u;
        x[e] = (typename M::W)(__funnelshift_r(lo, hi, sh) & mask);
    }
    if constexpr (TO != 0) exchange<M, LOGN, LOGE, 0, TO>(x, sm, tid);
}
template <class M, int LOGN, int LOGE, int FROM, bool LEAD, bool STREAM>
__device__ __forceinline__ void store_poly_packed(typename M::W (&x)[1 << LOGE], pk32 *__restrict__ g, int bits, bool valid,
                                                  typename M::W *sm, int tid) {
    typedef NttShape<LOGN, LOGE> S;
    static_assert(packed_shape_ok<LOGN, LOGE>(), "packed format: unsupported shape");
    if constexpr (FROM != 0) exchange<M, LOGN, LOGE, FROM, 0, LEAD>(x, sm, tid);
    const u32 lane = (u32)tid & 31u;
    // word `lane` of the warp's block covers block bits [32*lane, 32*lane+32): fields f0, f0+1, f0+2
    const u32 f0 = (32u * lane) / (u32)bits, a0 = 32u * lane - f0 * (u32)bits, a1 = (u32)bits - a0, a2 = a1 + (u32)bits;
    u32 *gw = reinterpret_cast<u32 *>(g) + (u32)(tid >> 5) * (u32)bits + lane;
#pragma unroll
    for (int e = 0; e < S::E; e++) {
        const u32 v = (u32)x[e];
        const u32 v0 = __shfl_sync(0xffffffffu, v, f0 & 31u), v1 = __shfl_sync(0xffffffffu, v, (f0 + 1u) & 31u),
                  v2 = __shfl_sync(0xffffffffu, v, (f0 + 2u) & 31u);
        const u32 word = (v0 >> a0) | (a1 < 32u ? v1 << a1 : 0u) | (a2 < 32u ? v2 << a2 : 0u);
        if (valid && lane < (u32)bits) gw[(u32)(S::pos(0, 0, e) >> 5) * (u32)bits] = word;
    }
}

// Polymul with NTT(a) parked in shared memory while b is transformed (instead of E more live registers): lets more
// CTAs be resident.  A thread reads back exactly the words it wrote, so no barrier is involved.
//   32-bit words: FHE_A_SMEM_MINB = resident CTAs asked of ptxas (N=1024 only; measured slower, off).
//   64-bit words: FHE_A_SMEM64 (degrees up to 2^FHE_A_SMEM64_MAXLOGN): the 64-bit polymul holds 2 x 32 registers of
//   coefficients and sits at 112-126 registers = 16 warps per SM otherwise.
#ifndef FHE_A_SMEM_MINB
#define FHE_A_SMEM_MINB 0
#endif
#ifndef FHE_A_SMEM64
#define FHE_A_SMEM64 0
#endif
#ifndef FHE_A_SMEM64_MAXLOGN
#define FHE_A_SMEM64_MAXLOGN 12
#endif
template <class M, int LOGN, int LOGE, int MODE> struct ASmem {
    static constexpr bool W32 = sizeof(typename M::W) == 4;
    static constexpr bool on = MODE == MODE_MUL &&
                               (W32 ? (FHE_A_SMEM_MINB > 0 && NttShape<LOGN, LOGE>::T <= 32 && LOGN == 10)
                                    : (FHE_A_SMEM64 > 0 && LOGN >= 6 && LOGN <= FHE_A_SMEM64_MAXLOGN));
#ifndef FHE_MUL_MINB
#define FHE_MUL_MINB 5
#endif
    // 0 = no minimum (ptxas' own choice; an explicit 1 makes it spend up to 65536/CT registers and lose occupancy).
    // The 32-bit polymul on 128-thread CTAs asks for five resident CTAs (<= 102 registers, no spills): measured
    // 0.775 of HBM peak at N=1024 against 0.74 with ptxas' unconstrained 139 registers (three CTAs).
    // measured (q = 65537, hbm fraction before -> after): polymul N=1024 0.74 -> 0.78 (CT=128, five CTAs); NTT N=16384
    // 0.47 -> 0.57 (CT=512, two CTAs of 64 registers instead of one of ~100).  CT=256 (N=8192): with immediate-offset
    // exchanges ptxas settles for 48 registers on the forward transform and serialises its loads (0.71); a bound of
    // three CTAs lets it spend 80 and hoist them (0.78) -- the opposite of what the bound did to the older code
    static constexpr int CT_ = KernelGeom<LOGN, LOGE>::CT;
    static constexpr bool MUL = is_mul_mode(MODE);
#ifndef FHE_NTT_MINB_256
#define FHE_NTT_MINB_256 3
#endif
#ifndef FHE_NTT_MINB_512
#define FHE_NTT_MINB_512 2
#endif
#ifndef FHE_INV_MINB_256
#define FHE_INV_MINB_256 3
#endif
#ifndef FHE_MUL_MINB_256
#define FHE_MUL_MINB_256 3
#endif
#ifndef FHE_MUL_MINB_E16   // 128-thread polymul CTAs with 16 coefficients per thread (N=2048)
#define FHE_MUL_MINB_E16 8  // 63 registers, no spills: N=2048 polymul 0.652 (5 CTAs) -> 0.677 (6) -> 0.690 (7) -> 0.697 (8) -> 0.701 (9, spills)
#endif
#ifndef FHE_MUL64_MINB
#define FHE_MUL64_MINB 0
#endif
#ifndef FHE_NTT64_MINB
#define FHE_NTT64_MINB 0
#endif
#ifndef FHE_NTT64_MINB_512   // 64-bit transforms on 512-thread CTAs (N = 8192): two CTAs of <= 64 registers per SM
#define FHE_NTT64_MINB_512 2   // measured, 62-bit q, N=8192: NTT 10.0 -> 11.5 M/s, INTT 10.4 -> 11.9 M/s
#endif
#ifndef FHE_MUL2_MINB      // dual-operand polymul, 128-thread CTAs, 32 coefficients per thread
#define FHE_MUL2_MINB FHE_MUL_MINB
#endif
#ifndef FHE_MUL2_MINB_E16
#define FHE_MUL2_MINB_E16 FHE_MUL_MINB_E16
#endif
#ifndef FHE_MUL2_MINB_256
#define FHE_MUL2_MINB_256 FHE_MUL_MINB_256
#endif
    static constexpr int mul_minb = MODE == MODE_MUL2
        ? (CT_ == 128 ? (LOGE == 4 ? FHE_MUL2_MINB_E16 : FHE_MUL2_MINB) : CT_ == 64 ? 2 * FHE_MUL2_MINB : CT_ == 256 ? FHE_MUL2_MINB_256 : 0)
        : (CT_ == 128 ? (LOGE == 4 ? FHE_MUL_MINB_E16 : FHE_MUL_MINB) : CT_ == 64 ? 2 * FHE_MUL_MINB : CT_ == 256 ? FHE_MUL_MINB_256 : 0);
#ifndef FHE_MULG_MINB_512
#define FHE_MULG_MINB_512 2
#endif
#ifndef FHE_MULG_MINB_256   // N = 8192: three CTAs of 80 registers (no spills) against four of 64 (16 spilled words):
#define FHE_MULG_MINB_256 3  // q = 65537 polymul 21.4 -> 21.8 M/s (A/B/A in one session, tools/run_mulg_ab.sh)
#endif
#ifndef FHE_MULG_MINB_128
#define FHE_MULG_MINB_128 8
#endif
    static constexpr int minb = MODE == MODE_MULG ? (CT_ == 512 ? FHE_MULG_MINB_512 : CT_ == 256 ? FHE_MULG_MINB_256 : CT_ == 128 ? FHE_MULG_MINB_128 : 0)
                                : (on && W32) ? FHE_A_SMEM_MINB
                                : !W32 ? (CT_ == 128 ? (MUL ? FHE_MUL64_MINB : FHE_NTT64_MINB)
                                          : (CT_ == 512 && !MUL) ? FHE_NTT64_MINB_512 : 0)
                                : MUL ? mul_minb
                                : (CT_ == 256 ? (MODE == MODE_INV ? FHE_INV_MINB_256 : FHE_NTT_MINB_256)
                                   : CT_ == 512 ? FHE_NTT_MINB_512 : 0);
    // dynamic shared memory of one CTA, in words of M::W
    static constexpr size_t words = (size_t)KernelGeom<LOGN, LOGE>::PPC *
        ((MODE == MODE_MUL2 ? 2 : 1) * PadRule<M, LOGN, LOGE>::padn + (on ? NttShape<LOGN, LOGE>::N : 0));
};

// I/O dispatch of the kernels below: plain words or the packed format (`bits` is only read by the latter)
template <class M, int LOGN, int LOGE, int TO, bool STREAM, typename IOW>
__device__ __forceinline__ void load_any(typename M::W (&x)[1 << LOGE], const IOW *__restrict__ g, int bits, typename M::W *sm,
                                         int tid) {
    if constexpr (IoTraits<IOW>::packed) load_poly_packed<M, LOGN, LOGE, TO, STREAM>(x, g, bits, sm, tid);
    else load_poly<M, LOGN, LOGE, TO, STREAM, IOW>(x, g, true, sm, tid);
}
template <class M, int LOGN, int LOGE, int FROM, bool LEAD, bool STREAM, typename IOW>
__device__ __forceinline__ void store_any(typename M::W (&x)[1 << LOGE], IOW *__restrict__ g, int bits, bool valid,
                                          typename M::W *sm, int tid) {
    if constexpr (IoTraits<IOW>::packed) store_poly_packed<M, LOGN, LOGE, FROM, LEAD, STREAM>(x, g, bits, valid, sm, tid);
    else store_poly<M, LOGN, LOGE, FROM, LEAD, STREAM, IOW>(x, g, valid, sm, tid);
}

template <class M, int LOGN, int LOGE, int MODE, typename IOW>
__global__ void __launch_bounds__(KernelGeom<LOGN, LOGE>::CT, ASmem<M, LOGN, LOGE, MODE>::minb)
ntt_kernel(const __grid_constant__ NttParams<M> P, const IOW *__restrict__ a, const IOW *__restrict__ b,
           IOW *__restrict__ c, IOW *__restrict__ c_evals, size_t batch, int flags) {
    typedef NttShape<LOGN, LOGE> S;
    typedef KernelGeom<LOGN, LOGE> G;
    typedef typename M::W W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PADW = PadRule<M, LOGN, LOGE>::padn;
    const int slot = threadIdx.x / S::T, tid = threadIdx.x % S::T;
    W *sm = reinterpret_cast<W *>(smem_raw) + (size_t)slot * PADW;
    const size_t poly = (size_t)blockIdx.x * G::PPC + slot;
    const bool valid = poly < batch;
    // slots past the end of a ragged batch read the last polynomial (and store nothing): unconditional loads, no
    // zero-filled registers
    const int bits = flags >> 8;  // packed format only (FHE_PACKED_BITS_SHIFT)
    const size_t off_ld = row_offset<IOW, S::N>(valid ? poly : batch - 1, bits);
    const size_t off = row_offset<IOW, S::N>(poly, bits);
    const M &m = P.mod;
    constexpr int LAST = S::P - 1;
    W x[S::E];

    if constexpr (MODE == MODE_FWD) {
        const TwSrc<M> tw = {P.c_fwd, P.fwd, P.fwdw};
        load_any<M, LOGN, LOGE, 0, false, IOW>(x, a + off_ld, bits, sm, tid);
        fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, tw);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.fwd_canon(x[e]);
        store_any<M, LOGN, LOGE, LAST, false, false, IOW>(x, c + off, bits, valid, sm, tid);  // last smem access: own reads in layout LAST
    } else if constexpr (MODE == MODE_INV) {
        const TwSrc<M> tw = {P.c_inv, P.inv, P.invw};
        load_any<M, LOGN, LOGE, LAST, false, IOW>(x, a + off_ld, bits, sm, tid);
        inv_chain<M, LOGN, LOGE, LAST>(x, sm, tid, m, tw, P.ninv, P.s_ninv);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.canon2(x[e]);
        store_any<M, LOGN, LOGE, 0, true, false, IOW>(x, c + off, bits, valid, sm, tid);
    } else {
        const TwSrc<M> twf = {P.c_fwd, P.fwd, P.fwdw};
        const TwSrc<M> twi = {P.c_inv, P.inv, P.invw};
        constexpr bool ST = ntt_stream(LOGN, MODE);
        if constexpr (MODE == MODE_MUL2) {
            W y[S::E];
            W *smy = reinterpret_cast<W *>(smem_raw) + (size_t)(G::PPC + slot) * PADW;
            load_any<M, LOGN, LOGE, 0, ST, IOW>(x, a + off_ld, bits, sm, tid);
            load_any<M, LOGN, LOGE, 0, ST, IOW>(y, (flags & B_BROADCAST) ? b : b + off_ld, bits, smy, tid);
            fwd_chain2<M, LOGN, LOGE, 0, false>(x, y, sm, smy, tid, m, twf);  // fresh shared memory: no lead barrier
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(m.fwd_out(x[e]), m.fwd_out(y[e]));
        } else {
            constexpr bool PARK = ASmem<M, LOGN, LOGE, MODE>::on;
            constexpr bool GPARK = MODE == MODE_MULG;
            W A[(PARK || GPARK) ? 1 : S::E];
            // GPARK: this thread's E words of NTT(a) rest in the first N * sizeof(W) bytes of its own output row
            // (the launcher rules out c aliasing an operand); lane-contiguous, read back by the thread that wrote them
            W *gA = reinterpret_cast<W *>(c + off) + tid;
            W *sA = reinterpret_cast<W *>(smem_raw) + (size_t)G::PPC * PADW + (size_t)slot * S::N;  // PARK only
            // both operands run through ONE copy of the forward-transform code (the fully unrolled transform is
            // the bulk of the kernel's instruction footprint; see profiles/: no_instruction stalls)
#ifndef FHE_NTT_PREFETCH_B   // experiment: pull b's lines into L2 while a is being transformed
#define FHE_NTT_PREFETCH_B 0
#endif
            if constexpr (FHE_NTT_PREFETCH_B != 0 && !IoTraits<IOW>::packed) {
                if (!(flags & B_BROADCAST)) {
                    const char *pb = reinterpret_cast<const char *>(b + off_ld);
                    constexpr int LINES = (int)(S::N * sizeof(IOW) / 128);
#pragma unroll
                    for (int l = tid; l < LINES; l += S::T) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (size_t)l * 128));
                }
            }
#pragma unroll 1
            for (int op = 0; op < 2; op++) {
                const IOW *src = op == 0 ? a + off_ld : (flags & B_BROADCAST) ? b : b + off_ld;
                if (flags & (op == 0 ? A_IS_EVALS : B_IS_EVALS)) {
                    load_any<M, LOGN, LOGE, LAST, ST, IOW>(x, src, bits, sm, tid);
                } else {
                    load_any<M, LOGN, LOGE, 0, ST, IOW>(x, src, bits, sm, tid);
                    fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, twf);
#pragma unroll
                    for (int e = 0; e < S::E; e++) x[e] = m.fwd_out(x[e]);
                }
                if (op == 0) {
                    if constexpr (GPARK) {
                        if (valid) {
#pragma unroll
                            for (int e = 0; e < S::E; e++) gA[e * S::T] = x[e];
                        }
                    } else if constexpr (PARK) {
#pragma unroll
                        for (int e = 0; e < S::E; e++) sA[e * S::T + tid] = x[e];
                    } else {
#pragma unroll
                        for (int e = 0; e < S::E; e++) A[e] = x[e];
                    }
                }
            }
            if constexpr (GPARK) {
                if (valid) {
#pragma unroll
                    for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(__ldcg(gA + e * S::T), x[e]);
                }
            } else if constexpr (PARK) {
#pragma unroll
                for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(sA[e * S::T + tid], x[e]);
            } else {
#pragma unroll
                for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(A[e], x[e]);
            }
        }
        if (c_evals != nullptr) {  // ring_nq.rs:606 -- the product keeps its evals
            W ev[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) ev[e] = m.pw_evals(x[e]);
            // the last smem access (end of the forward chain, or load_poly<LAST>) was this thread's reads in layout LAST
            store_any<M, LOGN, LOGE, LAST, false, ST, IOW>(ev, c_evals + off, bits, valid, sm, tid);
            if constexpr (S::P > 1) group_sync<S::T>();  // the evals store read layout 0: foreign words
        }
        inv_chain<M, LOGN, LOGE, LAST, false>(x, sm, tid, m, twi, P.ninv_pw, P.s_ninv_pw);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.canon2(x[e]);
        store_any<M, LOGN, LOGE, 0, true, ST, IOW>(x, c + off, bits, valid, sm, tid);
    }
}

// ---- large degrees: persistent CTAs with the operands staged in shared memory by asynchronous copies --------------
// At N >= 8192 a polynomial is one CTA of 256-512 threads with ~100 registers, so one to three CTAs are resident and
// nothing else on the SM covers the HBM round trip of an operand load.  Here a CTA loops over polynomials; every
// thread copies ITS OWN words of the next operand into a staging buffer with cp.async (LDGSTS: no destination
// register, no scoreboard wait) while it transforms the current one: b arrives during fwd(a), the next polynomial's a
// during fwd(b) / pointwise / inverse / store.  A thread only ever reads the staging words it copied itself, so
// cp.async.wait_group is the only synchronisation the staging needs.
template <int BYTES> __device__ __forceinline__ void cp_async_word(void *smem_dst, const void *gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gsrc), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

#ifndef FHE_STAGED_MINB_256   // resident CTAs asked of ptxas for the 256-thread shape (N = 8192): <= 128 registers
#define FHE_STAGED_MINB_256 2
#endif
template <class M, int LOGN, int LOGE, typename IOW> struct StagedGeom {
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int PADW = PadRule<M, LOGN, LOGE>::padn;
    static constexpr size_t exch_bytes = ((size_t)PADW * sizeof(typename M::W) + 15) / 16 * 16;
    static constexpr size_t smem = exch_bytes + (size_t)S::N * sizeof(IOW);
    static constexpr bool fits = S::T >= 128 && S::T <= 1024 && smem <= 227 * 1024;
    // resident CTAs asked of ptxas: two where two fit shared memory and 64 registers per thread can hold the shape
    static constexpr int minb = S::T <= 256 ? FHE_STAGED_MINB_256 : (S::T <= 512 && 2 * (smem + 1024) <= 227 * 1024) ? 2 : 1;
};
// TMA = true (FHE_NTT_STAGED=2): the same loop with the staging done by the TMA unit instead -- ONE thread posts the
// whole operand row as bulk copies (cp.async.bulk, UBLKCP) that complete on an mbarrier; no thread issues a per-word
// copy, and the waits are mbarrier.try_wait on the phase bit.  The row is read by every thread, so a CTA barrier
// separates the read-out of the staging buffer from the next bulk copy into it.
template <class M, int LOGN, int LOGE, typename IOW, bool TMA>
__global__ void __launch_bounds__(NttShape<LOGN, LOGE>::T, StagedGeom<M, LOGN, LOGE, IOW>::minb)
ntt_mul_staged_kernel(const __grid_constant__ NttParams<M> P, const IOW *__restrict__ a, const IOW *__restrict__ b,
                      IOW *__restrict__ c, IOW *__restrict__ c_evals, size_t batch, int flags) {
    typedef NttShape<LOGN, LOGE> S;
    typedef typename M::W W;
    typedef StagedGeom<M, LOGN, LOGE, IOW> SG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W *sm = reinterpret_cast<W *>(smem_raw);
    const int tid = threadIdx.x;
    const int t0 = S::pos(0, tid, 0);
    IOW *stg = reinterpret_cast<IOW *>(smem_raw + SG::exch_bytes) + t0;  // this thread's staging words: stg[pos(0,0,e)]
    const M &m = P.mod;
    const TwSrc<M> twf = {P.c_fwd, P.fwd, P.fwdw};
    const TwSrc<M> twi = {P.c_inv, P.inv, P.invw};
    constexpr int LAST = S::P - 1;
    __shared__ __align__(8) unsigned long long stage_bar;  // TMA only
    const u32 bar = tc_smem_u32(&stage_bar);
    u32 phase = 0;
    if constexpr (TMA) {
        if (tid == 0) {
            tc_mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    auto stage = [&](const IOW *src) {
        if constexpr (TMA) {
            if (tid == 0) {
                constexpr u32 BYTES = (u32)(S::N * sizeof(IOW)), CH = BYTES < 32768u ? BYTES : 32768u;
                const u32 dst = tc_smem_u32(smem_raw + SG::exch_bytes);
                tc_mbar_expect_tx(bar, BYTES);
#pragma unroll
                for (u32 o = 0; o < BYTES; o += CH) tc_bulk_g2s(dst + o, reinterpret_cast<const char *>(src) + o, CH, bar);
            }
        } else {
#pragma unroll
            for (int e = 0; e < S::E; e++) cp_async_word<sizeof(IOW)>(stg + S::pos(0, 0, e), src + t0 + S::pos(0, 0, e));
            cp_async_commit();
        }
    };
    auto staged = [&]() {   // the operand posted last has landed
        if constexpr (TMA) {
            tc_mbar_wait(bar, phase);
            phase ^= 1;
        } else {
            cp_async_wait_all();
        }
    };
    auto drained = [&]() {  // every thread has read its words out of the staging buffer
        if constexpr (TMA) __syncthreads();
    };
    size_t poly = blockIdx.x;
    if (poly < batch) stage(a + poly * S::N);
    for (; poly < batch; poly += gridDim.x) {
        const size_t off = poly * S::N;
        W x[S::E], A[S::E];
        staged();
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = (W)stg[S::pos(0, 0, e)];
        drained();
        stage((flags & B_BROADCAST) ? b : b + off);
        fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, twf);
#pragma unroll
        for (int e = 0; e < S::E; e++) A[e] = m.fwd_out(x[e]);
        staged();
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = (W)stg[S::pos(0, 0, e)];
        drained();
        if (poly + gridDim.x < batch) stage(a + (poly + gridDim.x) * S::N);
        fwd_chain<M, LOGN, LOGE>(x, sm, tid, m, twf);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.pw_mul(A[e], m.fwd_out(x[e]));
        if (c_evals != nullptr) {
            W ev[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) ev[e] = m.pw_evals(x[e]);
            store_poly<M, LOGN, LOGE, LAST, false, true, IOW>(ev, c_evals + off, true, sm, tid);
            if constexpr (S::P > 1) group_sync<S::T>();
        }
        inv_chain<M, LOGN, LOGE, LAST, false>(x, sm, tid, m, twi, P.ninv_pw, P.s_ninv_pw);
#pragma unroll
        for (int e = 0; e < S::E; e++) x[e] = m.canon2(x[e]);
        store_poly<M, LOGN, LOGE, 0, true, true, IOW>(x, c + off, true, sm, tid);
    }
}
template <class M, int LOGN, int LOGE, typename IOW, bool TMA>
int launch_staged(const NttParams<M> &P, const IOW *a, const IOW *b, IOW *c, IOW *c_evals, size_t batch, int flags,
                  cudaStream_t st) {
    typedef StagedGeom<M, LOGN, LOGE, IOW> SG;
    auto kern = ntt_mul_staged_kernel<M, LOGN, LOGE, IOW, TMA>;
    static std::atomic<int> resident[64];
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    int per_sm = resident[dev & 63].load(std::memory_order_acquire);
    if (per_sm == 0) {
        FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SG::smem));
        FHE_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SG::S::T, SG::smem));
        if (per_sm < 1) per_sm = 1;
        resident[dev & 63].store(per_sm, std::memory_order_release);
    }
    const size_t cap = (size_t)per_sm * num_sms();
    const unsigned grid = (unsigned)(batch < cap ? batch : cap);
    kern<<<grid, SG::S::T, SG::smem, st>>>(P, a, b, c, c_evals, batch, flags);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}
// MODE_MULS: polymul through ntt_mul_staged_kernel (two coefficient-form operands, degrees >= 2^13)
constexpr bool staged_instantiated(int logn) { return logn >= 13; }
#ifndef FHE_MULG_SMALL   // experiment: the output-row park at N = 1024 .. 4096 as well (FHE_NTT_GPARK=1 selects it)
#define FHE_MULG_SMALL 0
#endif
constexpr bool mulg_instantiated(int logn) { return logn >= 13 || (FHE_MULG_SMALL && logn >= 10); }

template <class M, int LOGN, int LOGE, int MODE, typename IOW>
int launch_one(const NttParams<M> &P, const IOW *a, const IOW *b, IOW *c, IOW *c_evals, size_t batch, int flags,
               cudaStream_t st) {
    typedef KernelGeom<LOGN, LOGE> G;
    auto kern = ntt_kernel<M, LOGN, LOGE, MODE, IOW>;
    const size_t smem = ASmem<M, LOGN, LOGE, MODE>::words * sizeof(typename M::W);
    if (smem > 48 * 1024) {  // opt in once per device (and per instantiation)
        static std::atomic<unsigned long long> done_mask{0};
        int dev = 0;
        FHE_CUDA_OK(cudaGetDevice(&dev));
        if (!((done_mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
            FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            done_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
        }
    }
    const size_t grid = (batch + G::PPC - 1) / G::PPC;
    FHE_REQUIRE(grid <= 0x7fffffffull, "batch too large for one launch");
    kern<<<(unsigned)grid, G::CT, smem, st>>>(P, a, b, c, c_evals, batch, flags);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

// does the dual-operand polymul fit one CTA's shared memory at this shape?
template <class M, int LOGN, int LOGE> constexpr bool mul2_fits() {
    return ASmem<M, LOGN, LOGE, MODE_MUL2>::words * sizeof(typename M::W) <= 227 * 1024;
}

template <class M, int LOGN, int LE, typename IOW>
int launch_modes(int mode, const NttParams<M> &P, const IOW *a, const IOW *b, IOW *c, IOW *c_evals, size_t batch,
                 int flags, cudaStream_t st) {
    switch (mode) {
        case MODE_FWD: return launch_one<M, LOGN, LE, MODE_FWD, IOW>(P, a, b, c, c_evals, batch, flags, st);
        case MODE_INV: return launch_one<M, LOGN, LE, MODE_INV, IOW>(P, a, b, c, c_evals, batch, flags, st);
        case MODE_MULG:
            if constexpr (mulg_instantiated(LOGN) && sizeof(typename M::W) <= sizeof(IOW) && !IoTraits<IOW>::packed) {
                if (c != a && c != b && c_evals == nullptr)
                    return launch_one<M, LOGN, LE, MODE_MULG, IOW>(P, a, b, c, c_evals, batch, flags, st);
            }
            return launch_one<M, LOGN, LE, MODE_MUL, IOW>(P, a, b, c, c_evals, batch, flags, st);
        case MODE_MULS:
            if constexpr (staged_instantiated(LOGN) && !IoTraits<IOW>::packed && StagedGeom<M, LOGN, LE, IOW>::fits) {
                if ((flags & (A_IS_EVALS | B_IS_EVALS)) == 0) {
                    // bulk copies need 16-byte aligned rows (a row is a multiple of 32 KB: only the bases matter)
                    const bool tma = (flags & STAGE_TMA) && (((uintptr_t)a | (uintptr_t)b) & 15) == 0;
                    return tma ? launch_staged<M, LOGN, LE, IOW, true>(P, a, b, c, c_evals, batch, flags, st)
                               : launch_staged<M, LOGN, LE, IOW, false>(P, a, b, c, c_evals, batch, flags, st);
                }
            }
            return launch_one<M, LOGN, LE, MODE_MUL, IOW>(P, a, b, c, c_evals, batch, flags, st);
        case MODE_MUL2:
            if constexpr (mul2_fits<M, LOGN, LE>()) {
                if ((flags & (A_IS_EVALS | B_IS_EVALS)) == 0)
                    return launch_one<M, LOGN, LE, MODE_MUL2, IOW>(P, a, b, c, c_evals, batch, flags, st);
            }
            // fall through: evals flags (or a shape that does not fit) take the generic polymul
        default: return launch_one<M, LOGN, LE, MODE_MUL, IOW>(P, a, b, c, c_evals, batch, flags, st);
    }
}

// `loge` must be a value ntt_loge_supported() accepts for (M, logn): the default LogE<M>::of(logn), or, for
// the tunable degrees, one of the alternatives instantiated below (u64 I/O only).
constexpr bool ntt_alt_loge(int logn, int loge, int wbytes, int iobytes) {
    if (iobytes != 8) return false;
    if (wbytes == 4)
        return (logn >= 10 && logn <= 12 && loge >= 3 && loge <= 5) || (logn >= 13 && logn <= 14 && loge >= 4 && loge <= 5);
    // 64-bit words: 8 or 16 coefficients per thread (4 was measured too: slower at every degree)
    return logn >= 10 && logn <= 13 && loge >= 3 && loge <= 4;
}
template <class M, int LOGN, typename IOW>
int launch_logn(int loge, int mode, const NttParams<M> &P, const IOW *a, const IOW *b, IOW *c, IOW *c_evals,
                size_t batch, int flags, cudaStream_t st) {
    constexpr int DEF = LogE<M>::of(LOGN);
    constexpr int WB = (int)sizeof(typename M::W), IB = (int)sizeof(IOW);
    if constexpr (ntt_alt_loge(LOGN, 3, WB, IB) && DEF != 3) {
        if (loge == 3) return launch_modes<M, LOGN, 3, IOW>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if constexpr (ntt_alt_loge(LOGN, 4, WB, IB) && DEF != 4) {
        if (loge == 4) return launch_modes<M, LOGN, 4, IOW>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if constexpr (ntt_alt_loge(LOGN, 5, WB, IB) && DEF != 5) {
        if (loge == 5) return launch_modes<M, LOGN, 5, IOW>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
    if (loge != DEF) {
        set_error("internal: unsupported coefficients-per-thread setting");
        return -1;
    }
    if constexpr (IoTraits<IOW>::packed && !packed_shape_ok<LOGN, DEF>()) {
        set_error("the bit-packed format needs n >= 1024");
        return -1;
    } else {
        return launch_modes<M, LOGN, DEF, IOW>(mode, P, a, b, c, c_evals, batch, flags, st);
    }
}
template <class M, typename IOW> bool ntt_loge_supported(int logn, int loge) {
    if (loge == LogE<M>::of(logn)) return true;
    return ntt_alt_loge(logn, loge, (int)sizeof(typename M::W), (int)sizeof(IOW));
}

template <class M, typename IOW>
int launch_ntt(int logn, int loge, int mode, const NttParams<M> &P, const IOW *a, const IOW *b, IOW *c, IOW *c_evals,
               size_t batch, int flags, cudaStream_t st) {
    constexpr int MAXLOGN = sizeof(typename M::W) == 4 ? 15 : 14;  // one CTA's shared memory holds the polynomial
    if (logn < 1 || logn > MAXLOGN) {
        set_error("unsupported ring degree for this modulus width (n must be 2..2^15 for q<2^30, 2..2^14 otherwise)");
        return -1;
    }
    switch (logn) {
#define FHE_CASE(L) case L: return launch_logn<M, L, IOW>(loge, mode, P, a, b, c, c_evals, batch, flags, st);
        FHE_CASE(1) FHE_CASE(2) FHE_CASE(3) FHE_CASE(4) FHE_CASE(5) FHE_CASE(6) FHE_CASE(7) FHE_CASE(8)
        FHE_CASE(9) FHE_CASE(10) FHE_CASE(11) FHE_CASE(12) FHE_CASE(13) FHE_CASE(14)
#undef FHE_CASE
        case 15:
            if constexpr (sizeof(typename M::W) == 4)
                return launch_logn<M, 15, IOW>(loge, mode, P, a, b, c, c_evals, batch, flags, st);
    }
    set_error("unsupported ring degree");
    return -1;
}

// one launcher per (policy, I/O word), each instantiated in its own translation unit (ntt_inst_*.cu)
#define FHE_NTT_INSTANTIATE(NAME, POLICY, IOW)                                                                        \
    int ntt_launch_##NAME(int logn, int loge, int mode, const NttParams<POLICY> &P, const IOW *a, const IOW *b, IOW *c, \
                          IOW *c_evals, size_t batch, int flags, cudaStream_t st) {                                    \
        return launch_ntt<POLICY, IOW>(logn, loge, mode, P, a, b, c, c_evals, batch, flags, st);                       \
    }                                                                                                                  \
    bool ntt_loge_ok_##NAME(int logn, int loge) { return ntt_loge_supported<POLICY, IOW>(logn, loge); }

}  // namespace fhe
