// ntt_core.cuh -- register-blocked negacyclic NTT passes (host+device per-thread logic).
//
// Reference semantics (arith/src/ntt.rs:44-110): in-place Cooley-Tukey forward (natural in,
// bit-reversed out, twiddle roots[m+i] at stage m=2^s for block i) and Gentleman-Sande inverse
// (twiddle roots_inv[m+i]) followed by a multiplication by n^-1.  The array index of an element
// never changes, so "position p" below is the reference's `r[p]`.
//
// Decomposition used here.  A polynomial of N=2^LOGN coefficients is owned by T=N/E threads with
// E=2^LOGE coefficients in registers each.  The LOGN stages are cut into P=ceil(LOGN/LOGE) passes of
// g_p stages (sum g_p = LOGN).  In pass p (stages s0..s0+g-1, s0 = g_0+..+g_{p-1}) a butterfly group
// is the 2^g positions   (H << (LOGN-s0)) | (r << nL) | L ,  r = 0..2^g-1,  nL = LOGN-s0-g,
// i.e. H = the s0 high bits (selects the twiddle block), L = the nL low bits.  A thread owns the
// E>>g groups  u = tid + qi*T  (H = u >> nL, L = u & (2^nL-1)), so all g stages of the pass are
// register-local, and the twiddle of local stage ls for the pair (ru, ru|half) is
//   roots[(1 << s) + (H << ls) + (ru >> (g-ls))] ,  s = s0+ls, half = 2^(g-1-ls).
// Between passes the E registers go through shared memory once (tools/ntt_model.py is the executable
// model of this index algebra, checked against the reference loop).
#pragma once
#include "modarith.cuh"

namespace fhe {

// Pass split.  Up to two passes: stages divided evenly, the odd one in front.  Three or more: every pass after the
// first takes LOGE stages and pass 0 the remainder.  That (a) makes every layout the exchanges touch conflict-free
// in shared memory under the i + (i >> 5) padding of 4-byte words (i + (i >> 4) for 8-byte words) -- a pass whose
// butterfly groups keep fewer than 5 (4) low index bits puts two group rows of a warp on the same banks -- and
// (b) leaves the last two passes equal, so the exchange between them stays inside one warp (ntt_kernels.cuh:
// ExchScope).  ncu before the change (N=8192, split 5/4/4): 13.0 M bank conflicts in 38.5 M shared wavefronts.
FHE_HD constexpr bool ntt_front_rem(int P) { return P >= 3; }

template <int LOGN, int LOGE> struct NttShape {
    static_assert(LOGE >= 1 && LOGE <= LOGN, "need 1 <= LOGE <= LOGN");
    static constexpr int N = 1 << LOGN;
    static constexpr int E = 1 << LOGE;
    static constexpr int T = N / E;                         // threads per polynomial
    static constexpr int P = (LOGN + LOGE - 1) / LOGE;      // passes
    static constexpr int BASE = LOGN / P, REM = LOGN % P;
    static constexpr bool FRONT = ntt_front_rem(P);
    static constexpr int G0 = LOGN - (P - 1) * LOGE;  // pass 0 of the FRONT rule
    FHE_HD static constexpr int g(int p) { return FRONT ? (p == 0 ? G0 : LOGE) : BASE + (p < REM ? 1 : 0); }
    FHE_HD static constexpr int s0(int p) {
        return FRONT ? (p == 0 ? 0 : G0 + (p - 1) * LOGE) : p * BASE + (p < REM ? p : REM);
    }
    FHE_HD static constexpr int nL(int p) { return LOGN - s0(p) - g(p); }
    // position of register slot e of thread tid in the layout of pass p
    FHE_HD static constexpr int pos(int p, int tid, int e) {
        return (((tid + (e >> g(p)) * T) >> nL(p)) << (LOGN - s0(p))) | ((e & ((1 << g(p)) - 1)) << nL(p)) |
               ((tid + (e >> g(p)) * T) & ((1 << nL(p)) - 1));
    }
};

// Device twiddle-table order.  The reference stores stage s (m = 2^s) at roots[2^s + j], j = (H << ls) + hi
// for the butterfly block H of pass p (ls = s - s0(p)).  Threads of a warp differ in H, so reading that
// order would touch a different cache line per lane.  The device tables therefore keep stage s at
//   2^s + (hi << s0(p)) + H                                  (pass 0, the first stage of every pass, 64-bit policies)
//   2^s + ((hi >> 1) << (s0(p) + 1)) + (H << 1) + (hi & 1)   (later stages of later passes, 32-bit policies)
// (a rotation of the index bits inside each stage block, identity for pass 0), which makes every warp load
// lane-contiguous AND puts the twiddles of hi = 2m, 2m+1 side by side: a thread always needs both (the two children of
// a radix-4 block, two neighbouring blocks of a radix-2 stage), so 32-bit policies fetch them with ONE 128-bit load
// (tw_pair).  Measured with a throw-away build that simply skipped every second twiddle load: N = 1024 polymul +4 %,
// N = 4096 +6.5 % -- the loads' issue slots and LSU queue entries, not their bytes, were what cost.  (A pair of 64-bit
// twiddles is 32 bytes: two loads either way, and side by side each of them would use half of every cache line it
// touches -- measured -3.5 % at N = 4096 -- so the 64-bit policies keep the first order.)
// tw_slot() is the host-side map reference index -> device slot; tw_idx() the device-side one.
inline int ntt_num_passes(int logn, int loge) { return (logn + loge - 1) / loge; }
inline int ntt_pass_s0(int logn, int loge, int p) {
    const int P = ntt_num_passes(logn, loge), base = logn / P, rem = logn % P;
    if (ntt_front_rem(P)) return p == 0 ? 0 : (logn - (P - 1) * loge) + (p - 1) * loge;
    return p * base + (p < rem ? p : rem);
}
inline u64 tw_slot(int logn, int loge, u64 ref_index, bool paired) {
    if (ref_index == 0) return 0;
    int s = 0;
    while ((2ull << s) <= ref_index) s++;  // 2^s <= ref_index < 2^(s+1)
    const int P = ntt_num_passes(logn, loge);
    int p = 0;
    while (p + 1 < P && ntt_pass_s0(logn, loge, p + 1) <= s) p++;
    const int s0 = ntt_pass_s0(logn, loge, p), ls = s - s0;
    const u64 j = ref_index - (1ull << s), H = j >> ls, hi = j & ((1ull << ls) - 1);
    if (!paired || p == 0 || ls == 0) return (1ull << s) + (hi << s0) + H;
    return (1ull << s) + ((hi >> 1) << (s0 + 1)) + (H << 1) + (hi & 1);   // pairs (2m, 2m+1) of hi side by side (tw_idx)
}
// coefficients per thread (log2): 32 for 32-bit words, 16 for 64-bit words (register budget of the
// polymul kernel, which keeps NTT(a) in registers while transforming b)
// Measured on B200 (q = 65537, fraction of HBM peak, 16 vs 32 coefficients per thread, conflict-free pass split):
// N=2048 NTT 0.97 vs 0.85, polymul 0.62 vs 0.61 -- hence 16 per thread there; N=4096 polymul 0.55 vs 0.60, INTT
// 0.86 vs 0.90, NTT 0.90 vs 0.88 -- 32 per thread (with the old 4/4/4 split, two-way bank conflicts in the middle
// pass, 16 per thread was the faster one at N=4096 too).
// 64-bit words: the fully unrolled polymul at 16 coefficients per thread is 92-114 KB of SASS and 112+ registers; ncu
// showed it bound by instruction fetch (SM instruction-cache hit rate 62 %, GPC-level instruction requests at 100 % of
// their peak, no_instruction the top stall: profiles/r2_polymul_q62_n1024_*).  8 coefficients per thread halve code
// and registers: N=1024 polymul 36.4 -> 43.9 M/s, N=2048 17.2 -> 19.5 M/s; from N=4096 on the extra exchange pass
// costs more than the smaller code gives back (8.74 vs 7.83 M/s), so 16 per thread stays there.
template <class M> struct LogE {
    static constexpr bool W32 = sizeof(typename M::W) == 4;
    static constexpr int MAXE = W32 ? 5 : 4;
    static constexpr int of(int logn) {
        return (W32 && logn == 11) ? 4 : (!W32 && logn >= 3 && logn <= 11) ? 3 : logn < MAXE ? logn : MAXE;
    }
};

// Twiddle source: pass 0 twiddles are identical for every thread (H = 0), so they are taken from a
// small by-value table living in the kernel-parameter constant bank; later passes index the global
// table (L1/L2 resident, n entries).
template <class M> struct TwSrc {
    const typename M::T *c0;    // first 2^g(0) entries (constant bank / host array)
    const typename M::T *tab;   // full table, n entries, reference order roots[m+i]
    const u32 *tabw;            // radix-4 policies, large degrees: the twiddles alone (4 bytes each), see tw_load
};
// Twiddle of a later pass (table slot i).  Fermat32 at the degrees whose tables outgrow L1 (M::compact) reads the 4-byte
// twiddle and computes its Shoup companion: 2^32 = q (2^16 - 1) + 1, so floor(w 2^32 / q) = w (2^16 - 1) for w < q.
template <class M, int LOGN> FHE_HD typename M::T tw_load(const TwSrc<M> &tw, int i) {
    if constexpr (M::RADIX4) {
        if constexpr (M::compact(LOGN)) {
            const u32 w = tw.tabw[i];
            return typename M::T{w, (w << 16) - w};
        } else {
            return tw.tab[i];
        }
    } else {
        return tw.tab[i];
    }
}

// device slot of the twiddle of local stage LS (global stage s0 + LS), block index hi, butterfly-group row H
template <class M> FHE_HD constexpr bool tw_paired() { return sizeof(typename M::T) == 8; }
template <class M, int PASS, int S0, int LS> FHE_HD constexpr int tw_idx(int hi, int H) {
    return (!tw_paired<M>() || PASS == 0 || LS == 0) ? (1 << (S0 + LS)) + (hi << S0) + H
                                                     : (1 << (S0 + LS)) + ((hi >> 1) << (S0 + 1)) + (H << 1) + (hi & 1);
}
// the twiddles of blocks hi (even) and hi + 1 of local stage LS >= 1: one vector load where they sit side by side
template <class M, int LOGN, int PASS, int S0, int LS>
FHE_HD void tw_pair(const TwSrc<M> &tw, int hi, int H, typename M::T &t0, typename M::T &t1) {
    static_assert(LS >= 1, "a stage with one twiddle per row has no pairs");
    const int i = tw_idx<M, PASS, S0, LS>(hi, H), j = tw_idx<M, PASS, S0, LS>(hi + 1, H);
    if constexpr (PASS == 0) {
        t0 = tw.c0[i];
        t1 = tw.c0[j];
    } else {
#if defined(__CUDA_ARCH__)
        // j == i + 1, i even.  Below n = 1024 a transform is a few threads of a warp (the digit transforms of the fused
        // torus kernels): the loads are broadcasts, and unpacking a 128-bit load costs IMAD.MOVs on the pipe that binds
        // there (n = 64, k = 4 external product: 6.93 -> 6.59 M/s with vector loads) -- two adjacent 64-bit loads instead.
        if constexpr (tw_paired<M>() && LOGN >= 10) {
            bool done = false;
            if constexpr (M::RADIX4) {
                if constexpr (M::compact(LOGN)) {
                    const uint2 w = *reinterpret_cast<const uint2 *>(tw.tabw + i);
                    t0 = typename M::T{w.x, (w.x << 16) - w.x};
                    t1 = typename M::T{w.y, (w.y << 16) - w.y};
                    done = true;
                }
            }
            if (!done) {
                const uint4 q = *reinterpret_cast<const uint4 *>(tw.tab + i);
                t0 = typename M::T{q.x, q.y};
                t1 = typename M::T{q.z, q.w};
            }
            return;
        }
#endif
        t0 = tw_load<M, LOGN>(tw, i);
        t1 = tw_load<M, LOGN>(tw, j);
    }
}
// twiddle of block hi of a stage: fetched as the pair (hi, hi + 1) when hi is even, remembered in tp for hi + 1
template <class M, int LOGN, int PASS, int S0, int LS>
FHE_HD typename M::T stage_tw(const TwSrc<M> &tw, int hi, int H, typename M::T (&tp)[2]) {
    if constexpr (PASS == 0) {   // constant bank: the operand is read in place, nothing to pair up
        return tw.c0[tw_idx<M, 0, S0, LS>(hi, H)];
    } else if constexpr (LS == 0) {
        return tw_load<M, LOGN>(tw, tw_idx<M, PASS, S0, 0>(0, H));
    } else {
        if ((hi & 1) == 0) tw_pair<M, LOGN, PASS, S0, LS>(tw, hi, H, tp[0], tp[1]);
        return tp[hi & 1];
    }
}

// One butterfly stage (local stage LS of pass PASS) over the thread's registers.  All trip counts are
// compile-time constants so the register array never gets dynamically indexed.
template <class M, int LOGN, int LOGE, int PASS, int LS>
FHE_HD void fwd_stage(typename M::W (&x)[1 << LOGE], int tid, const M &m, const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int g = S::g(PASS), s0 = S::s0(PASS), nL = S::nL(PASS), G = 1 << g;
    constexpr int half = 1 << (g - 1 - LS);
#pragma unroll
    for (int qi = 0; qi < (S::E >> g); qi++) {
        const int H = (PASS == 0) ? 0 : ((tid + qi * S::T) >> nL);  // pass 0: u < 2^nL
        typename M::T tp[2];
#pragma unroll
        for (int hi = 0; hi < (1 << LS); hi++) {
            // reference index (1<<s) + (H<<LS) + hi, stored at the lane-contiguous slot (see tw_slot)
            const typename M::T t = stage_tw<M, LOGN, PASS, s0, LS>(tw, hi, H, tp);
#pragma unroll
            for (int lo = 0; lo < half; lo++) {
                const int ru = (hi << (g - LS)) | lo;
                m.fwd(x[qi * G + ru], x[qi * G + ru + half], t);
            }
        }
    }
}
// Radix-4 policies (M::RADIX4, modarith.cuh: Fermat32): local stages LS and LS+1 of the pass in one sweep over blocks
// of four registers.  Parent twiddle roots[i] (stage LS), even child roots[2i] (stage LS+1), and -- in the device slot of
// the odd child roots[2i+1], which the radix-4 form does not use -- the product roots[i]*roots[2i] (radix4_patch).
template <class M, int LOGN, int LOGE, int PASS, int LS, bool DUAL>
FHE_HD void fwd_stage4(typename M::W (&x)[1 << LOGE], typename M::W (&y)[1 << LOGE], int tid, const M &m,
                       const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int g = S::g(PASS), s0 = S::s0(PASS), nL = S::nL(PASS), G = 1 << g;
    static_assert(LS + 1 < g, "radix-4 needs two stages of the same pass");
    constexpr int h = 1 << (g - 2 - LS);
    if constexpr (PASS == 0 && LS == 0 && M::FIRST_SHIFT) {  // stages 1 and 2 of the transform: twiddles known, inputs canonical
#pragma unroll
        for (int qi = 0; qi < (S::E >> g); qi++) {
#pragma unroll
            for (int lo = 0; lo < h; lo++) {
                const int b = qi * G + lo;
                m.fwd4_first(x[b], x[b + h], x[b + 2 * h], x[b + 3 * h]);
                if constexpr (DUAL) m.fwd4_first(y[b], y[b + h], y[b + 2 * h], y[b + 3 * h]);
            }
        }
        return;
    }
#pragma unroll
    for (int qi = 0; qi < (S::E >> g); qi++) {
        const int H = (PASS == 0) ? 0 : ((tid + qi * S::T) >> nL);
        typename M::T tp[2];
#pragma unroll
        for (int hi = 0; hi < (1 << LS); hi++) {
            const typename M::T t1 = stage_tw<M, LOGN, PASS, s0, LS>(tw, hi, H, tp);
            typename M::T t2, t3;   // the children 2 hi and 2 hi + 1 (the latter's slot holds parent * child)
            tw_pair<M, LOGN, PASS, s0, LS + 1>(tw, 2 * hi, H, t2, t3);
#pragma unroll
            for (int lo = 0; lo < h; lo++) {
                const int b = qi * G + (hi << (g - LS)) + lo;
                m.fwd4(x[b], x[b + h], x[b + 2 * h], x[b + 3 * h], t1, t2, t3);
                if constexpr (DUAL) m.fwd4(y[b], y[b + h], y[b + 2 * h], y[b + 3 * h], t1, t2, t3);
            }
        }
    }
}
// forward pairing: local stages (0,1), (2,3), ... ; an odd stage count leaves the last one radix-2
template <class M, int LOGN, int LOGE, int PASS, int LS = 0>
FHE_HD void fwd_pass(typename M::W (&x)[1 << LOGE], int tid, const M &m, const TwSrc<M> &tw) {
    constexpr int g = NttShape<LOGN, LOGE>::g(PASS);
    if constexpr (M::RADIX4 && LS + 1 < g) {
        fwd_stage4<M, LOGN, LOGE, PASS, LS, false>(x, x, tid, m, tw);
        if constexpr (LS + 2 < g) fwd_pass<M, LOGN, LOGE, PASS, LS + 2>(x, tid, m, tw);
    } else {
        fwd_stage<M, LOGN, LOGE, PASS, LS>(x, tid, m, tw);
        if constexpr (LS + 1 < g) fwd_pass<M, LOGN, LOGE, PASS, LS + 1>(x, tid, m, tw);
    }
}

// Two polynomials through the same forward pass at once (the two operands of a polymul): every twiddle is fetched
// once and applied to both register sets, and the two butterfly streams are independent of each other (twice the
// instruction-level parallelism per warp).
template <class M, int LOGN, int LOGE, int PASS, int LS>
FHE_HD void fwd_stage2(typename M::W (&x)[1 << LOGE], typename M::W (&y)[1 << LOGE], int tid, const M &m,
                       const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int g = S::g(PASS), s0 = S::s0(PASS), nL = S::nL(PASS), G = 1 << g;
    constexpr int half = 1 << (g - 1 - LS);
#pragma unroll
    for (int qi = 0; qi < (S::E >> g); qi++) {
        const int H = (PASS == 0) ? 0 : ((tid + qi * S::T) >> nL);
        typename M::T tp[2];
#pragma unroll
        for (int hi = 0; hi < (1 << LS); hi++) {
            const typename M::T t = stage_tw<M, LOGN, PASS, s0, LS>(tw, hi, H, tp);
#pragma unroll
            for (int lo = 0; lo < half; lo++) {
                const int ru = (hi << (g - LS)) | lo;
                m.fwd(x[qi * G + ru], x[qi * G + ru + half], t);
                m.fwd(y[qi * G + ru], y[qi * G + ru + half], t);
            }
        }
    }
}
template <class M, int LOGN, int LOGE, int PASS, int LS = 0>
FHE_HD void fwd_pass2(typename M::W (&x)[1 << LOGE], typename M::W (&y)[1 << LOGE], int tid, const M &m,
                      const TwSrc<M> &tw) {
    constexpr int g = NttShape<LOGN, LOGE>::g(PASS);
    if constexpr (M::RADIX4 && LS + 1 < g) {
        fwd_stage4<M, LOGN, LOGE, PASS, LS, true>(x, y, tid, m, tw);
        if constexpr (LS + 2 < g) fwd_pass2<M, LOGN, LOGE, PASS, LS + 2>(x, y, tid, m, tw);
    } else {
        fwd_stage2<M, LOGN, LOGE, PASS, LS>(x, y, tid, m, tw);
        if constexpr (LS + 1 < g) fwd_pass2<M, LOGN, LOGE, PASS, LS + 1>(x, y, tid, m, tw);
    }
}

// Inverse: the same groups, local stages in descending order.  When PASS == 0 the final stage
// (s = 0, the single twiddle roots_inv[1]) also applies n^-1 (M::inv_last).
// KOVR >= 0 replaces K (radix-4 policies track the bound themselves, see InvSched)
template <class M, int LOGN, int LOGE, int PASS, int LS, int KOVR = -1>
FHE_HD void inv_stage(typename M::W (&x)[1 << LOGE], int tid, const M &m, const TwSrc<M> &tw,
                      typename M::T ninv, typename M::T s_ninv) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int g = S::g(PASS), s0 = S::s0(PASS), nL = S::nL(PASS), G = 1 << g;
    constexpr int half = 1 << (g - 1 - LS);
    // inverse stages already executed (stages run from LOGN-1 down to 0): inputs are below 2q * 2^K
    constexpr int K = KOVR >= 0 ? KOVR : LOGN - 1 - (s0 + LS);
#pragma unroll
    for (int qi = 0; qi < (S::E >> g); qi++) {
        const int H = (PASS == 0) ? 0 : ((tid + qi * S::T) >> nL);
        typename M::T tp[2];
#pragma unroll
        for (int hi = 0; hi < (1 << LS); hi++) {
            if constexpr (PASS == 0 && LS == 0) {
#pragma unroll
                for (int lo = 0; lo < half; lo++)
                    m.template inv_last_k<K>(x[qi * G + lo], x[qi * G + lo + half], ninv, s_ninv);
            } else {
                const typename M::T t = stage_tw<M, LOGN, PASS, s0, LS>(tw, hi, H, tp);
#pragma unroll
                for (int lo = 0; lo < half; lo++) {
                    const int ru = (hi << (g - LS)) | lo;
                    m.template inv_k<K>(x[qi * G + ru], x[qi * G + ru + half], t);
                }
            }
        }
    }
}

// Inverse schedule of a radix-4 policy.  Stages run from LOGN-1 down to 0, pass by pass; inside a pass the local stages
// are paired from the top -- (g-1, g-2), (g-3, g-4), ... -- the last stage of the whole transform (pass 0, local stage
// 0: the n^-1 stage) always stays on its own, and a stage left over is radix-2.  KB = log2(bound of the inputs / 2q):
// +1 per radix-2 stage, +2 per radix-4 layer; a radix-4 layer must be entered with KB <= M::INV_KB_MAX, so the layer in
// front of one that would not be folds its sum output (KB back to 0: every other output is a twiddle product, below 2q).
struct InvStep {
    int kind;  // 0 = lower stage of a pair (nothing to do), 1 = radix-2, 2 = upper stage of a pair, 3 = last stage
    int kb;    // KB of the inputs
    bool fold;
};
template <int LOGN, int LOGE, int KBMAX> struct InvSched {
    typedef NttShape<LOGN, LOGE> S;
    FHE_HD static constexpr int lo_of(int p) { return p == 0 ? 1 : 0; }
    // KB at which the next radix-4 layer after (p, ls) [exclusive] would be entered if the current KB is kb; -1 if none
    FHE_HD static constexpr int next_r4_kb(int p, int ls, int kb) {
        for (;;) {
            if (ls < lo_of(p)) {
                if (p == 0) return -1;
                p--;
                ls = S::g(p) - 1;
                continue;
            }
            if (ls - 1 >= lo_of(p)) return kb;
            kb++;  // a single radix-2 stage
            ls--;
        }
    }
    FHE_HD static constexpr InvStep at(int PASS, int LS) {
        int kb = 0;
        for (int p = S::P - 1; p >= 0; p--) {
            int ls = S::g(p) - 1;
            while (ls >= lo_of(p)) {
                if (ls - 1 >= lo_of(p)) {
                    const int nxt = next_r4_kb(p, ls - 2, kb + 2);
                    const bool fold = nxt > KBMAX;
                    if (p == PASS && ls == LS) return InvStep{2, kb, fold};
                    if (p == PASS && ls - 1 == LS) return InvStep{0, kb, fold};
                    kb = fold ? 0 : kb + 2;
                    ls -= 2;
                } else {
                    if (p == PASS && ls == LS) return InvStep{1, kb, false};
                    kb++;
                    ls--;
                }
            }
            if (p == 0 && PASS == 0 && LS == 0) return InvStep{3, kb, false};
        }
        return InvStep{-1, 0, false};
    }
};

// local stages LS (children, executed first) and LS-1 (parent) of the pass in one sweep; table slots as in fwd_stage4
template <class M, int LOGN, int LOGE, int PASS, int LS, int KB, bool FOLD>
FHE_HD void inv_stage4(typename M::W (&x)[1 << LOGE], int tid, const M &m, const TwSrc<M> &tw) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int g = S::g(PASS), s0 = S::s0(PASS), nL = S::nL(PASS), G = 1 << g;
    constexpr int LP = LS - 1;
    static_assert(LP >= 0 && LS < g, "radix-4 needs two stages of the same pass");
    constexpr int h = 1 << (g - 1 - LS);
#pragma unroll
    for (int qi = 0; qi < (S::E >> g); qi++) {
        const int H = (PASS == 0) ? 0 : ((tid + qi * S::T) >> nL);
        typename M::T tp[2];
#pragma unroll
        for (int hi = 0; hi < (1 << LP); hi++) {
            const typename M::T t1 = stage_tw<M, LOGN, PASS, s0, LP>(tw, hi, H, tp);
            typename M::T t2, t3;
            tw_pair<M, LOGN, PASS, s0, LS>(tw, 2 * hi, H, t2, t3);
#pragma unroll
            for (int lo = 0; lo < h; lo++) {
                const int b = qi * G + (hi << (g - LP)) + lo;
                m.template inv4<KB, FOLD>(x[b], x[b + h], x[b + 2 * h], x[b + 3 * h], t2, t3, t1);
            }
        }
    }
}

template <class M, int LOGN, int LOGE, int PASS, int LS = -1>
FHE_HD void inv_pass(typename M::W (&x)[1 << LOGE], int tid, const M &m, const TwSrc<M> &tw,
                     typename M::T ninv, typename M::T s_ninv) {
    constexpr int ls = LS < 0 ? NttShape<LOGN, LOGE>::g(PASS) - 1 : LS;
    if constexpr (M::RADIX4) {
        constexpr InvStep st = InvSched<LOGN, LOGE, M::INV_KB_MAX>::at(PASS, ls);
        static_assert(st.kind >= 1, "inverse schedule: entered at the lower stage of a pair");
        if constexpr (st.kind == 2) {
            inv_stage4<M, LOGN, LOGE, PASS, ls, st.kb, st.fold>(x, tid, m, tw);
            if constexpr (ls >= 2) inv_pass<M, LOGN, LOGE, PASS, ls - 2>(x, tid, m, tw, ninv, s_ninv);
        } else {
            inv_stage<M, LOGN, LOGE, PASS, ls, st.kb>(x, tid, m, tw, ninv, s_ninv);
            if constexpr (ls > 0) inv_pass<M, LOGN, LOGE, PASS, ls - 1>(x, tid, m, tw, ninv, s_ninv);
        }
    } else {
        inv_stage<M, LOGN, LOGE, PASS, ls>(x, tid, m, tw, ninv, s_ninv);
        if constexpr (ls > 0) inv_pass<M, LOGN, LOGE, PASS, ls - 1>(x, tid, m, tw, ninv, s_ninv);
    }
}

}  // namespace fhe
