// xp_octet.cuh -- table-driven first stages of the external product's digit transforms (host+device per-thread logic,
// replayed on the CPU by tests/emu; the kernel is extprod_fused.cu).
//
// A digit of Tn::decompose with beta = 2 (tfhe/src/tggsw.rs:49-50, arith/src/torus.rs:43-52) is a polynomial of BITS, and
// in pass 0 of the register-blocked forward NTT (ntt_core.cuh) the twiddle of a butterfly depends on the register slot
// only (H = 0).  Stage LS of pass 0 pairs the slots ru, ru + (G >> (LS+1)) of a G = 2^g(0) group with twiddle
// roots[2^LS + (ru >> (g-LS))] (arith/src/ntt.rs:56-60): on the 8 slots  base + jj * (G/8)  of an OCTET the first three
// stages are therefore one fixed linear map of the 8 input bits -- the same for every octet, thread and digit.  A
// 256-entry table of its images replaces 12 butterflies by two 128-bit loads.  The entries are produced by running the
// butterfly code itself on the 256 bit patterns, so the lazy representatives are exactly the ones the stages produce.
#pragma once
#include "ntt_core.cuh"

namespace fhe {

#if defined(__CUDACC__)
typedef uint4 XpQuad;
#else
struct XpQuad { u32 x, y, z, w; };
#endif

// Small32 with its forward butterfly spelled so that every addition is a THREE-input one (`zero` is a kernel-parameter
// word holding 0, opaque to ptxas): a two-input add may be emitted as IMAD.IADD on the fmaheavy pipe, the pipe the
// butterflies' multiplies already saturate (ncu, n = 1024, k = 1: fmaheavy 68 % active against ALU 24 %, 72 of the 442
// fmaheavy instructions of a digit transform being IMAD.IADD / IMAD.MOV); IADD3 only exists on the ALU pipe.  Same
// for the final fold, whose negation (IMAD.MOV) moves into the constant negq = -q.
// Z = false keeps the two-input add (the small rings, whose kernels are ALU- rather than fmaheavy-heavy).
template <bool Z> struct XpSmallT : Small32 {
    u32 zero, negq;
    FHE_HD void fwd(u32 &x, u32 &y, Tw32 t) const {
        const u32 V = mul_tw(y, t);
        y = x - V + q2;
        if constexpr (Z) x = x + V + zero;
        else x = x + V;
    }
    // x mod 2^27 + (x >> 27) * (2^27 - q) < 2^28 for q < 2^27 close to it, congruent to x
    FHE_HD u32 fold27(u32 x) const { return (x >> 27) * negq + x; }
};
typedef XpSmallT<true> XpSmall;

template <int LOGN> struct XpOct {
    static_assert(LOGN >= 6, "the table-driven first stages need a first pass of at least three stages on 32 coefficients");
    static constexpr int LOGE = 5;
    typedef NttShape<LOGN, LOGE> S;
    static_assert(S::g(0) >= 3, "pass 0 must hold at least three stages (LOGN = 6..10, 13..15 at 32 coefficients per thread)");
    static constexpr int G = 1 << S::g(0), STRIDE = G >> 3;
    // register slot of input bit jj of octet o (o = 0..3): bit 8*o + jj of a thread's plane word
    FHE_HD static constexpr int slot(int o, int jj) { return (o / STRIDE) * G + jj * STRIDE + (o % STRIDE); }
};

// Table entry b of one prime: stages 0-2 on the bit pattern b.  Stage 0 works on bits (U, V in {0,1}: V = b*S is a
// select, not a multiplication); stages 1, 2 are the butterflies of the 8-point shape, whose twiddle indices 2 + hi and
// 4 + hi are those of every octet.
template <class M> FHE_HD void octet_table_entry(const M &ms, const TwSrc<M> &twf, int b, XpQuad &lo, XpQuad &hi) {
    u32 y[8];
    for (int j = 0; j < 8; j++) y[j] = ((u32)b >> j) & 1u;
    const u32 S1 = twf.c0[1].w;
    for (int j = 0; j < 4; j++) {
        const u32 U = y[j], V = (0u - y[j + 4]) & S1;
        y[j] = U + V;
        y[j + 4] = U + ms.q2 - V;
    }
    fwd_pass<M, 3, 3, 0, 1>(y, 0, ms, twf);
    lo.x = y[0]; lo.y = y[1]; lo.z = y[2]; lo.w = y[3];
    hi.x = y[4]; hi.y = y[5]; hi.z = y[6]; hi.w = y[7];
}

// Pass 0 of a digit transform: the thread's 32 coefficients are the bits of w (bit 8*o + jj = slot(o, jj)).
template <int LOGN, class M>
FHE_HD void digit_pass0(u32 (&x)[32], u32 w, const XpQuad *tab_lo, const XpQuad *tab_hi, int tid, const M &ms,
                        const TwSrc<M> &twf) {
    typedef XpOct<LOGN> O;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int o = 0; o < 4; o++) {
        const u32 b = (w >> (8 * o)) & 255u;
        const XpQuad lo4 = tab_lo[b], hi4 = tab_hi[b];
        x[O::slot(o, 0)] = lo4.x; x[O::slot(o, 1)] = lo4.y; x[O::slot(o, 2)] = lo4.z; x[O::slot(o, 3)] = lo4.w;
        x[O::slot(o, 4)] = hi4.x; x[O::slot(o, 5)] = hi4.y; x[O::slot(o, 6)] = hi4.z; x[O::slot(o, 7)] = hi4.w;
    }
    if constexpr (O::S::g(0) > 3) fwd_pass<M, LOGN, O::LOGE, 0, 3>(x, tid, ms, twf);
}

}  // namespace fhe
