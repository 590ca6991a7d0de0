// extprod_fused.cu -- fused TGGSW x TGLWE external product / CMux (tfhe/src/tggsw.rs:39-62): one CTA per
// accumulator, ONE HBM round trip (read the TGLWE, write the TGLWE); decompose -> NTT -> MAC -> INTT -> CRT
// all stay on chip.  Same exact arithmetic as the unfused path in torus_kernels.cu (two NTT primes, TGGSW
// rows as two 32-bit limbs transformed once at load, centred CRT lift, recombination mod 2^64).
//
// Per CTA (256 threads), for each prime r in {p1, p2}:
//   rounds of SLOTS = 256/T concurrent digit NTTs (T = N/32 threads, 32 coefficients per thread in registers):
//     - slot s takes the bit polynomial of digit d = (i, j): bit 63-j of x_i (Tn::decompose, torus.rs:43-52) -- ONE
//       32-bit word per thread, the inputs being kept as bit planes -- looks the first three butterfly stages up in
//       a 256-entry table per byte (the stages are linear and their twiddles the same for every thread), runs the rest
//       of the register-blocked forward NTT (csub-free butterflies: primes < 2^27 leave the headroom) and leaves
//       NTT(digit) in its shared-memory slot, reduced to [0, 2p);
//     - all threads then MAC the round's digits against the resident TGGSW: thread t owns the items
//       I = t + 256 m of the (component, limb, position) space; 64-bit accumulators in registers
//       (products < 2^55, at most (k+1)*64 <= 512 of them), one IMAD.WIDE per MAC, 128-bit key loads;
//   then 2(k+1) inverse NTTs and the residues go to shared memory; after both primes the CRT lift, the
//   recombination lo + (hi << 32) and the CMux addend produce the output coefficients.
#include "../../include/fhe_b200.h"
#include "ntt_kernels.cuh"
#include "runtime.cuh"
#include "torus.cuh"
#include "xp_octet.cuh"

#include <stdlib.h>

namespace fhe {

// Threads per CTA: 256 with one accumulator, or -- n = 1024, k = 1 only -- 512 with TWO accumulators that share every key
// load (same occupancy: 16 warps, 128 registers, one 512-thread CTA per SM instead of two of 256).  The pair halves the
// key bytes through the L1 data pipe (ncu: 69 % busy, the busiest unit of the 256-thread kernel): extprod 1.83 -> 1.86 M/s,
// CMux chain 1.72 -> 1.82 M CMux/s.  A batch too small to give every SM a pair keeps the 256-thread kernel (xp_use_pair).
// Both read ONE fused key layout, the one written for xp_layout_ct() threads.
__host__ __device__ constexpr int xp_layout_ct(int logn, int k1) { return (logn == 10 && k1 == 2) ? 512 : 256; }
__host__ __device__ constexpr bool xp_has_pair(int logn, int k1) { return logn == 10 && k1 == 2; }

template <int LOGN, int K1, int CT_ = 256> struct XpGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGE = LOGN < 5 ? LOGN : 5;
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int CT = CT_, LCT = xp_layout_ct(LOGN, K1);
    // forward-butterfly policy: three-input adds (xp_octet.cuh) except at n = 64, k = 4, the one shape they slow down
    // (6.99 -> 6.34 M/s; n = 128, k = 1: 15.2 -> 16.0 with them, n = 1024, k = 1: +1 %, the others unchanged)
    typedef XpSmallT<!(LOGN <= 6 && K1 > 2)> Mod;
    static_assert(CT == 256 || (CT == 512 && xp_has_pair(LOGN, K1)), "512 threads: n = 1024, k = 1 only");
    static constexpr int SLOTS = CT / S::T;             // concurrent NTTs
    static constexpr int ND = K1 * 64;                  // digit polynomials per accumulator
    static constexpr int PADN = Pad32<LOGN, LOGE>::padn;   // ntt_kernels.cuh: PadRule
    // The decomposed inputs are kept as BIT PLANES: row (accumulator, component, thread tn of a digit transform) holds
    // 64 words, word j = the 32 coefficients that thread owns in digit (component, j), bit 8*o + jj = register slot
    // oct_slot(o, jj) (xp_octet.cuh).  One pad word per row: rows are written along j and read along tn.
    static constexpr int PLANE_ROW = 65;
    __host__ __device__ static constexpr int oct_slot(int o, int jj) { return XpOct<LOGN>::slot(o, jj); }
    // first three stages of a digit transform by table (xp_octet.cuh)
    static constexpr size_t TAB_BYTES = (size_t)2 * 2 * 256 * 16;   // [prime][outputs 0-3 | 4-7][byte] uint4
    static constexpr int UNITS = K1 * 2;                // (component, limb)
    static constexpr int ITEMS = UNITS * N;
    static constexpr int IPT = (ITEMS + CT - 1) / CT;   // MAC items per thread (per accumulator)
    static constexpr int IPT4 = (IPT + 3) / 4 * 4;      // padded to whole uint4 loads
    static constexpr int Q = IPT4 / 4;                  // key quads per thread and digit
    // MAC item m of thread t.  With the layout's own thread count: t + CT * m.  A 256-thread CTA on the 512-thread
    // layout plays the layout's threads t and t + 256: m < IPT/2 are the items of the first, the rest of the second.
    __host__ __device__ static constexpr int item(int t, int m) {
        return CT == LCT ? t + CT * m : t + CT * (m / (IPT / 2)) + LCT * (m % (IPT / 2));
    }
    // uint4 index of key quad v of thread t inside a digit's Q * CT quads ([v][thread] in the layout's geometry)
    __host__ __device__ static constexpr int quad(int t, int v) {
        return CT == LCT ? v * CT + t : (v % (Q / 2)) * LCT + t + CT * (v / (Q / 2));
    }
    // the items of a thread sit on NPOS distinct coefficient positions (CT * 4 >= N in every instantiated shape)
    static constexpr int NPOS = IPT < 4 ? IPT : (N >= CT ? (N / CT < 4 ? N / CT : 4) : 1);
    __host__ __device__ static constexpr int pos_slot(int m) {
        return CT == LCT ? (N >= CT ? m % (N / CT) : 0) : 2 * (m / (IPT / 2)) + (m & 1);
    }
    __host__ __device__ static constexpr int slot_pos(int t, int sl) {
        return CT == LCT ? (t + CT * sl) & (N - 1) : t + CT * (sl >> 1) + LCT * (sl & 1);
    }
    // Accumulators per CTA.  The MAC streams the whole transformed TGGSW (2 * ND * ITEMS * 4 B) from L2 once per
    // CTA; small rings leave registers and shared memory for several accumulators, which then share every key load
    // (n = 64, k = 4: 1.6 MB of key per 2.5 KB accumulator -- L2 bandwidth, not arithmetic, was the limit).
#ifndef FHE_XP_A10
#define FHE_XP_A10 1
#endif
#ifndef FHE_XP_A_SMALL
#define FHE_XP_A_SMALL 4
#endif
    static constexpr int A = CT == 512 ? 2 : LOGN <= 7 ? FHE_XP_A_SMALL : (LOGN <= 9 && ITEMS <= 2048) ? 2 : (LOGN == 10 && K1 == 2) ? FHE_XP_A10 : 1;
    // resident CTAs asked of ptxas: three 80-register CTAs help the small rings (n=64,k=4: 6.0 -> 6.6 M/s), while at
    // n=1024 two 128-register CTAs are faster (1.44 vs 1.27 M/s; the chain even 1.28 vs 0.91 M CMux/s)
    static constexpr int MINB = CT == 512 ? 1 : LOGN <= 7 ? (FHE_XP_A_SMALL > 4 ? 2 : 3) : (LOGN == 10 && A > 1) ? 1 : 2;
    static constexpr int DPR = SLOTS / A;               // digits per round (each for all A accumulators)
    // digits whose key quads are in flight in the MAC (registers: KDEPTH * IPT4 words)
#ifndef FHE_XP_KDEPTH10
#define FHE_XP_KDEPTH10 2
#endif
    static constexpr int KDEPTH = (LOGN == 10 && K1 == 2) ? FHE_XP_KDEPTH10 : 2;
    static constexpr int ROUNDS = (ND + DPR - 1) / DPR;
    // slots whose threads run an inverse transform (whole warps do): the slots behind them are free for the
    // residues of the second prime
    static constexpr int LIVE_SLOTS = S::T >= 32 ? A * UNITS : ((A * UNITS * S::T + 31) / 32) * (32 / S::T);
    // the second prime's residues live in the free exchange slots when there are enough of them
    static constexpr bool RES2_IN_XCH = (SLOTS - LIVE_SLOTS) * PADN >= A * UNITS * N;
    static constexpr size_t PLANES_BYTES = ((size_t)A * K1 * S::T * PLANE_ROW * 4 + 15) / 16 * 16;
    static constexpr size_t XCH_BYTES = (size_t)SLOTS * PADN * 4, RES_BYTES = (size_t)(RES2_IN_XCH ? 1 : 2) * A * UNITS * N * 4;
    static constexpr size_t SMEM = PLANES_BYTES + XCH_BYTES + RES_BYTES + TAB_BYTES;
    static_assert(XCH_BYTES % 16 == 0 && RES_BYTES % 16 == 0, "the tables behind them hold uint4 words");
    static constexpr size_t SMEM_CHAIN = SMEM;
    static_assert(SLOTS % A == 0 && A * UNITS <= SLOTS, "need a slot per inverse transform");
    static_assert(ND % DPR == 0, "every round must be full: the digit transforms synchronise whole warps");
    static_assert(S::T <= 32, "one digit NTT must fit a warp (N <= 1024) in this kernel");
    static_assert(S::E == 32 && CT >= 256, "bit planes: 32 coefficients per thread; table build: one entry per thread");
};

// 128-bit read-only key load that stays where the source puts it: as a plain __ldg ptxas sank the loads issued in
// front of the barrier back behind it
__device__ __forceinline__ uint4 ldg_key(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

static void init_xp_mod(XpSmall &m, u64 p) {
    init_mod(static_cast<Small32 &>(m), p);
    m.zero = 0;
    m.negq = (u32)0 - (u32)p;
}

struct XpParams {
    NttParams<Lazy32> P[2];   // plans of p1, p2 (device-order tables, n^-1 constants)
    XpSmall ms[2];            // same moduli, csub-free forward butterflies (xp_octet.cuh)
    u64 mu[2];                // floor(2^64 / p_r): Barrett constant for the 64-bit accumulators
    const u32 *R[2];          // fused key layout: R[r][d][v][t][j] = NTT value of item t + LCT (4v + j) of digit d, t < LCT = xp_layout_ct()
                              // (one uint4 per thread and v: every warp load is 512 contiguous bytes)
    CrtParams cp;
};

// CMux chain (blind rotation): acc <- cmux(key[j], acc, X^{-h[j]} acc) for j < steps, the accumulator staying in
// shared memory for the whole chain (ONE HBM round trip per chain instead of one per CMux).
struct XpChain {
    const u32 *const *R;   // [steps][2] fused key layouts of the TGGSW of every step (device array)
    const u64 *h;          // [batch][steps] rotation amounts
    int steps;
    int negacyclic;        // 0: TGLWE::left_rotate (h mod n, ring_torus.rs:118-132); 1: true X^{-h}, h mod 2n
};

// acc mod p for acc < 2^63 (result canonical)
__device__ __forceinline__ u32 reduce64(u64 acc, u32 p, u64 mu) {
    const u64 qh = __umul64hi(acc, mu);
    u64 r = acc - qh * p;  // in [0, 2p)
    return (u32)(r >= p ? r - p : r);
}

// 32 x 32 bit-matrix transpose across a warp: lane L gives row L, lane j returns column j (bit L = bit j of lane L's
// word).  Five block-swap steps (distance 16 ... 1), one shuffle each.
__device__ __forceinline__ u32 warp_transpose32(u32 a, int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const u32 m = s == 16 ? 0x0000FFFFu : s == 8 ? 0x00FF00FFu : s == 4 ? 0x0F0F0F0Fu : s == 2 ? 0x33333333u : 0x55555555u;
        const u32 o = __shfl_xor_sync(0xffffffffu, a, s);
        a = (lane & s) ? ((a & ~m) | ((o >> s) & m)) : ((a & m) | ((o << s) & ~m));
    }
    return a;
}

// NTT of a digit polynomial under one prime, left in `sm` (padded position order, values < 2^28).  `w` holds the
// thread's 32 coefficients (bits).
//  * stages 0-2 come out of the octet table (tab_lo / tab_hi of this prime);
//  * the csub-free butterflies leave values < (2*LOGN+1)*p < 2^32; the final partial reduction
//    x - (x >> 27)*p = (x mod 2^27) + (x >> 27)*(2^27 - p) < 2^27 + 21*2^21 < 2^28 costs one shift and one IMAD
//    (the MAC then adds at most (k+1)*64 <= 320 products < 2^28 * 2^27: below 2^64).
template <int LOGN, int K1, int CT>
__device__ __forceinline__ void digit_ntt(const typename XpGeom<LOGN, K1, CT>::Mod &ms, const TwSrc<typename XpGeom<LOGN, K1, CT>::Mod> &twf, u32 w, const uint4 *tab_lo,
                                          const uint4 *tab_hi, u32 *sm, int tid) {
    typedef XpGeom<LOGN, K1, CT> G;
    constexpr int LOGE = G::LOGE;
    typedef NttShape<LOGN, LOGE> S;
    constexpr int LAST = S::P - 1;
    u32 x[S::E];
    digit_pass0<LOGN>(x, w, tab_lo, tab_hi, tid, ms, twf);
    if constexpr (S::P > 1) fwd_chain<typename G::Mod, LOGN, LOGE, 1>(x, sm, tid, ms, twf);
#pragma unroll
    for (int e = 0; e < S::E; e++) x[e] = ms.fold27(x[e]);
    exch_put<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));
}

template <int LOGN, int K1, bool CHAIN, int CT>
__global__ void __launch_bounds__(CT, XpGeom<LOGN, K1, CT>::MINB)
extprod_fused_kernel(const __grid_constant__ XpParams X, const u64 *__restrict__ ct1, const u64 *__restrict__ ct2,
                     u64 *out, int cmux, const XpChain ch, size_t batch) {
    typedef XpGeom<LOGN, K1, CT> G;
    typedef typename G::S S;
    constexpr int LOGE = G::LOGE, N = G::N, LAST = S::P - 1, A = G::A, GLWE = K1 * N;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    u32 *planes = reinterpret_cast<u32 *>(smem_raw);                      // [A][K1][T][PLANE_ROW] bit planes of the inputs
    u32 *xch = reinterpret_cast<u32 *>(smem_raw + G::PLANES_BYTES);       // [SLOTS][PADN] exchange / NTT(digit)
    u32 *res1 = xch + (size_t)G::SLOTS * G::PADN;                         // [A][UNITS][N] residues mod p1
    u32 *res2 = G::RES2_IN_XCH ? xch + (size_t)G::LIVE_SLOTS * G::PADN    // [A][UNITS][N] residues mod p2 (free slots of xch)
                               : res1 + (size_t)A * G::UNITS * N;
    const int t = threadIdx.x;
    const int slot = t / S::T, tid = t % S::T;
    const int s_acc = slot % A, s_dig = slot / A;   // forward phase: which accumulator, which digit of the round
    u32 *sm = xch + (size_t)slot * G::PADN;
    const size_t acc0 = (size_t)blockIdx.x * A;     // first accumulator of this CTA
    const int na = (int)(batch - acc0 < (size_t)A ? batch - acc0 : (size_t)A);
    const size_t base = acc0 * GLWE;
    uint4 *tab = reinterpret_cast<uint4 *>(smem_raw + G::PLANES_BYTES + G::XCH_BYTES + G::RES_BYTES);   // [2][2][256]
    if (t < 256) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
            const TwSrc<typename G::Mod> twf = {X.P[r].c_fwd, X.P[r].fwd};
            octet_table_entry(reinterpret_cast<const typename G::Mod &>(X.ms[r]), twf, t, tab[r * 512 + t], tab[r * 512 + 256 + t]);
        }
    }   // ordered before the first digit transform by the barrier behind the bit planes
    const int lane = t & 31;
    const int my_slot = G::oct_slot(lane >> 3, lane & 7);   // register slot of plane bit `lane`

    const int steps = CHAIN ? ch.steps : 1;
#pragma unroll 1
    for (int step = 0; step < steps; step++) {
    // input of the external product: ct (extprod), ct2 - ct1 (TGGSW::cmux, tggsw.rs:39-41), or for the chain
    // X^{-h} acc - acc (the CMux of tlwe.rs:140-146 with ct2 = acc.left_rotate(h)).  The chain's accumulator
    // lives in this CTA's own output rows between steps (L2-resident; every HBM line is written once).
    // Bit planes: a warp takes (accumulator, component, transform thread tn) rows, lane L loads the coefficient of
    // register slot my_slot, two warp transposes turn the 32 values of a row into its 64 plane words.  The scattered
    // loads of up to PB rows are issued together (one row at a time the phase was 4.6 % of the stall samples).
    constexpr int NGRP = A * K1 * S::T, NW = G::CT / 32, GPW = (NGRP + NW - 1) / NW, PB = GPW < 8 ? GPW : 8;
    auto plane_input = [&](int grp) -> u64 {
        const int tn = grp % S::T, c = (grp / S::T) % K1, aa = grp / (S::T * K1);
        const int p = S::pos(0, tn, my_slot), rem = (c << LOGN) + p;
        const size_t i = (size_t)aa * GLWE + rem;
        if (aa >= na) return 0;
        if (CHAIN) {
            const u64 *acc_g = (step == 0 ? ct1 : out) + base + (size_t)aa * GLWE;
            const u64 hraw = ch.h[(acc0 + aa) * steps + step];
            const u32 h = (u32)(hraw & (N - 1));
            const bool flip = ch.negacyclic && ((hraw >> LOGN) & 1);
            const u32 src = (u32)p + h;
            u64 v = src < (u32)N ? acc_g[(c << LOGN) + src] : (u64)0 - acc_g[(c << LOGN) + src - N];
            if (flip) v = (u64)0 - v;
            return v - acc_g[rem];
        }
        return cmux ? ct2[base + i] - ct1[base + i] : ct1[base + i];
    };
#pragma unroll 1
    for (int g0 = 0; g0 < GPW; g0 += PB) {
        u64 v[PB];
#pragma unroll
        for (int u = 0; u < PB; u++) {
            const int grp = (t >> 5) + (g0 + u) * NW;
            v[u] = (g0 + u < GPW && grp < NGRP) ? plane_input(grp) : 0;
        }
#pragma unroll
        for (int u = 0; u < PB; u++) {
            const int grp = (t >> 5) + (g0 + u) * NW;
            if (g0 + u < GPW && grp < NGRP) {   // warp-uniform
                const u32 whi = warp_transpose32((u32)(v[u] >> 32), lane), wlo = warp_transpose32((u32)v[u], lane);
                u32 *row = planes + (size_t)grp * G::PLANE_ROW;
                row[31 - lane] = whi;   // digit j reads bit 63 - j (Tn::decompose, torus.rs:43-52): bit 32 + lane is digit 31 - lane
                row[63 - lane] = wlo;
            }
        }
    }
    __syncthreads();

#pragma unroll 1
    for (int r = 0; r < 2; r++) {
        const u32 *Rr = CHAIN ? ch.R[2 * step + r] : X.R[r];
        const Lazy32 &ml = X.P[r].mod;
        u64 acc[A][G::IPT];
#pragma unroll
        for (int aa = 0; aa < A; aa++)
#pragma unroll
            for (int m = 0; m < G::IPT; m++) acc[aa][m] = 0;
#pragma unroll 1
        for (int round = 0; round < G::ROUNDS; round++) {
            const int d = round * G::DPR + s_dig;
            if (d < G::ND) {
                const TwSrc<typename G::Mod> twf = {X.P[r].c_fwd, X.P[r].fwd};
                const u32 w = planes[((size_t)(s_acc * K1 + (d >> 6)) * S::T + tid) * G::PLANE_ROW + (d & 63)];
                digit_ntt<LOGN, K1, CT>(reinterpret_cast<const typename G::Mod &>(X.ms[r]), twf, w, tab + r * 512, tab + r * 512 + 256, sm, tid);
            }
            // MAC of the round's digits against the resident TGGSW, two digits per iteration.  The 128-bit key loads are
            // software-pipelined through the registers themselves: the loads of the first pair are issued BEFORE the
            // barrier that ends the transforms (the transform registers are dead by then), and every key quad is
            // re-loaded for the next pair right behind the four MACs that consumed it.  Before, an iteration issued
            // its loads and sat on their L2 latency: 17 % of all stall samples on the first IMAD.WIDE of the loop
            // (profiles/r2_extprod_fused_n1024_k1_ncu_full_b.csv).
            constexpr int nd = G::DPR, Q = G::Q;   // ND % DPR == 0: every round is full
            constexpr int KD = G::KDEPTH;                 // digits whose key quads are in flight
            static_assert(nd % KD == 0, "the MAC takes the digits KD at a time");
            const uint4 *Rt = reinterpret_cast<const uint4 *>(Rr) + (size_t)round * G::DPR * Q * G::CT;
            uint4 kq[KD][Q];
#pragma unroll
            for (int h2 = 0; h2 < KD; h2++)
#pragma unroll
                for (int v = 0; v < Q; v++) kq[h2][v] = ldg_key(Rt + (size_t)h2 * Q * G::CT + G::quad(t, v));
            __syncthreads();
#pragma unroll(nd <= 8 ? nd / KD : 1)
            for (int it = 0; it < nd / KD; it++) {
                const bool more = it + 1 < nd / KD;
#pragma unroll
                for (int h2 = 0; h2 < KD; h2++) {
                    const int dd = KD * it + h2;
                    u32 dv[A][G::NPOS];   // the items of a thread share NPOS coefficient positions
#pragma unroll
                    for (int aa = 0; aa < A; aa++)
#pragma unroll
                        for (int sl = 0; sl < G::NPOS; sl++)
                            dv[aa][sl] = xch[(size_t)(dd * A + aa) * G::PADN + Pad32<LOGN, LOGE>::idx(G::slot_pos(t, sl))];
#pragma unroll
                    for (int v = 0; v < Q; v++) {
                        const u32 k4[4] = {kq[h2][v].x, kq[h2][v].y, kq[h2][v].z, kq[h2][v].w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int m = 4 * v + j;
                            if (m < G::IPT) {   // item >= ITEMS only when ITEMS % CT != 0: key padding is zero
#pragma unroll
                                for (int aa = 0; aa < A; aa++) acc[aa][m] += (u64)dv[aa][G::pos_slot(m)] * k4[j];
                            }
                        }
                        if (more) kq[h2][v] = ldg_key(Rt + (size_t)(dd + KD) * Q * G::CT + G::quad(t, v));
                    }
                }
            }
            __syncthreads();
        }
        // accumulators -> inverse-transform inputs (slot aa*UNITS + u, padded position order).  (Parking the
        // accumulators in shared memory during the transforms removes the MOVs ptxas spends on re-pairing them in
        // the MAC loop, but measured slower: the MAC phase is latency-, not issue-bound.)
#pragma unroll
        for (int aa = 0; aa < A; aa++)
#pragma unroll
            for (int m = 0; m < G::IPT; m++) {
                const int item = G::item(t, m);
                if (item < G::ITEMS)
                    xch[(size_t)(aa * G::UNITS + (item >> LOGN)) * G::PADN + Pad32<LOGN, LOGE>::idx(item & (N - 1))] = reduce64(acc[aa][m], ml.q, X.mu[r]);
            }
        __syncthreads();
        // warp-uniform condition: every lane of a warp that owns at least one live slot runs the transform
        // (the exchanges inside synchronise whole warps); lanes of dead slots compute on scratch and store nothing
        if ((t & ~31) / S::T < A * G::UNITS) {
            const TwSrc<Lazy32> twi = {X.P[r].c_inv, X.P[r].inv};
            u32 x[S::E];
            exch_get<Lazy32, LOGN, LOGE, LAST>(x, sm + Pad32<LOGN, LOGE>::idx(S::pos(LAST, tid, 0)));
            inv_chain<Lazy32, LOGN, LOGE, LAST>(x, sm, tid, ml, twi, X.P[r].ninv, X.P[r].s_ninv);
            if (slot < A * G::UNITS) {
                u32 *R = (r == 0 ? res1 : res2) + (size_t)slot * N;
#pragma unroll
                for (int e = 0; e < S::E; e++) R[S::pos(0, tid, e)] = ml.canon2(x[e]);
            }
        }
        __syncthreads();
    }
    // CRT lift, recombination, addend
    for (int i = t; i < A * GLWE; i += G::CT) {
        const int aa = i / GLWE, rem = i % GLWE;
        if (aa >= na) continue;
        const int c = rem >> LOGN, p = rem & (N - 1);
        const int u0 = (aa * G::UNITS + c * 2) * N + p, u1 = u0 + N;
        const u64 lo = crt_centered(res1[u0], res2[u0], X.cp.p1, X.cp.p2, X.cp.p1_inv_mod_p2, X.cp.P, X.cp.halfP, X.cp.m2);
        const u64 hi = crt_centered(res1[u1], res2[u1], X.cp.p1, X.cp.p2, X.cp.p1_inv_mod_p2, X.cp.P, X.cp.halfP, X.cp.m2);
        const u64 addend = CHAIN ? (step == 0 ? ct1[base + i] : out[base + i]) : (cmux ? ct1[base + i] : 0);
        out[base + i] = addend + lo + (hi << 32);
    }
    if (CHAIN) __syncthreads();  // the next step reads this CTA's output rows (and reuses xch / res1)
    }  // step
}

// unfused key layout (u64, [d][u][x]) -> fused layout (u32, [d][v][t][j], item = t + ct (4v + j), ct = xp_layout_ct()), zero padded
__global__ void tggsw_fused_layout_kernel(const u64 *__restrict__ R, u32 *__restrict__ Rf, int nd, int items, int ipt4, int ct) {
    const size_t total = (size_t)nd * ct * ipt4;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(idx & 3), t = (int)((idx >> 2) % ct), v = (int)((idx / (4 * (size_t)ct)) % (ipt4 / 4));
        const int d = (int)(idx / ((size_t)ipt4 * ct));
        const int m = 4 * v + j;
        const int item = t + ct * m;
        Rf[idx] = item < items ? (u32)R[(size_t)d * items + item] : 0u;
    }
}

// Which kernel for a batch (FHE_XP_CT=256|512 forces one: tests, A/B runs).  Pair CTAs run one per SM and a wave of
// them takes as long as a wave of two 256-thread CTAs per SM minus ~2 %, but their last wave is all-or-nothing, while
// a last wave of at most one 256-thread CTA per SM finishes in ~0.6 of a wave (the CTA has the SM to itself).  Measured
// on 148 SMs, M extprod/s (256 | 512): batch 148: 1.50 | 0.90, 296: 1.71 | 1.79, 1024: 1.75 | 1.60, 1184: 1.80 | 1.85,
// 4144: 1.83 | 1.86.  The estimate below reproduces those orderings.
static bool xp_use_pair(size_t batch) {
    const char *e = getenv("FHE_XP_CT");
    const int forced = e ? atoi(e) : 0;
    if (forced == 256 || forced == 512) return forced == 512;
    const size_t S = (size_t)num_sms(), pairs = (batch + 1) / 2;
    const double t_pair = (double)((pairs + S - 1) / S);
    const size_t r = batch % (2 * S);
    const double t_single = (double)(batch / (2 * S)) * 1.016 + (r == 0 ? 0.0 : r <= S ? 0.62 : 1.04);
    return t_pair <= t_single;
}

template <int LOGN, int K1, int CT>
static int launch_fused_ct(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, int cmux, cudaStream_t st) {
    typedef XpGeom<LOGN, K1, CT> G;
    const TorusCtx &tc = *g.tc;
    XpParams X;
    X.P[0] = tc.plan1->p32;
    X.P[1] = tc.plan2->p32;
    init_xp_mod(X.ms[0], TORUS_P1);
    init_xp_mod(X.ms[1], TORUS_P2);
    X.mu[0] = ~0ull / TORUS_P1;
    X.mu[1] = ~0ull / TORUS_P2;
    X.R[0] = g.R1f;
    X.R[1] = g.R2f;
    X.cp = tc.cp;
    FHE_REQUIRE(batch <= 0x7fffffffull, "extprod: batch too large");
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    auto kern = extprod_fused_kernel<LOGN, K1, false, CT>;
    const int threads = G::CT;
    const size_t smem = G::SMEM;
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done_mask |= 1ull << (dev & 63);
    }
    kern<<<(unsigned)((batch + G::A - 1) / G::A), threads, smem, st>>>(X, ct1, ct2, out, cmux, XpChain{nullptr, nullptr, 1, 0}, batch);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int LOGN, int K1>
static int launch_fused(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, int cmux, cudaStream_t st) {
    if constexpr (xp_has_pair(LOGN, K1)) {
        if (xp_use_pair(batch)) return launch_fused_ct<LOGN, K1, 512>(g, ct1, ct2, out, batch, cmux, st);
    }
    return launch_fused_ct<LOGN, K1, 256>(g, ct1, ct2, out, batch, cmux, st);
}

template <int LOGN, int K1, int CT>
static int launch_chain_ct(const TorusCtx &tc, const u32 *const *keys_dev, const u64 *h_dev, int steps, int negacyclic,
                           const u64 *acc_in, u64 *acc_out, size_t batch, cudaStream_t st) {
    typedef XpGeom<LOGN, K1, CT> G;
    XpParams X;
    X.P[0] = tc.plan1->p32;
    X.P[1] = tc.plan2->p32;
    init_xp_mod(X.ms[0], TORUS_P1);
    init_xp_mod(X.ms[1], TORUS_P2);
    X.mu[0] = ~0ull / TORUS_P1;
    X.mu[1] = ~0ull / TORUS_P2;
    X.R[0] = X.R[1] = nullptr;
    X.cp = tc.cp;
    FHE_REQUIRE(batch <= 0x7fffffffull, "cmux chain: batch too large");
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    auto kern = extprod_fused_kernel<LOGN, K1, true, CT>;
    if (!((done_mask >> (dev & 63)) & 1ull)) {
        FHE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM_CHAIN));
        done_mask |= 1ull << (dev & 63);
    }
    kern<<<(unsigned)((batch + G::A - 1) / G::A), G::CT, G::SMEM_CHAIN, st>>>(X, acc_in, nullptr, acc_out, 1,
                                                                              XpChain{keys_dev, h_dev, steps, negacyclic}, batch);
    count_launch(1);
    FHE_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int LOGN, int K1>
static int launch_chain(const TorusCtx &tc, const u32 *const *keys_dev, const u64 *h_dev, int steps, int negacyclic,
                        const u64 *acc_in, u64 *acc_out, size_t batch, cudaStream_t st) {
    if constexpr (xp_has_pair(LOGN, K1)) {
        if (xp_use_pair(batch)) return launch_chain_ct<LOGN, K1, 512>(tc, keys_dev, h_dev, steps, negacyclic, acc_in, acc_out, batch, st);
    }
    return launch_chain_ct<LOGN, K1, 256>(tc, keys_dev, h_dev, steps, negacyclic, acc_in, acc_out, batch, st);
}

#define FHE_XP_SHAPES(F) F(10, 2) F(9, 2) F(8, 2) F(6, 5) F(6, 2) F(7, 2) F(8, 3) F(9, 3)

bool extprod_fused_supported(int logn, int k1) {
#define F(L, K) if (logn == L && k1 == K) return true;
    FHE_XP_SHAPES(F)
#undef F
    return false;
}
static int fused_ipt4(int logn, int k1) {
    const int items = k1 * 2 * (1 << logn), ct = xp_layout_ct(logn, k1), ipt = (items + ct - 1) / ct;
    return (ipt + 3) / 4 * 4;
}

// builds the fused key layout from the unfused one (called once at load when the shape is supported)
int tggsw_build_fused_layout(Tggsw &g, cudaStream_t st) {
    const int logn = g.tc->logn, k1 = (int)g.k + 1;
    if (!extprod_fused_supported(logn, k1)) return 0;
    // the fused kernels read the plans' twiddle tables in the device order of 32 coefficients per thread
    if (g.tc->plan1->loge != (logn < 5 ? logn : 5) || g.tc->plan2->loge != g.tc->plan1->loge) return 0;
    const int nd = k1 * 64, items = k1 * 2 * (1 << logn), ipt4 = fused_ipt4(logn, k1);
    const int ct = xp_layout_ct(logn, k1);
    const size_t words = (size_t)nd * ct * ipt4;
    FHE_CUDA_OK(cudaMalloc((void **)&g.R1f, words * sizeof(u32)));
    FHE_CUDA_OK(cudaMalloc((void **)&g.R2f, words * sizeof(u32)));
    size_t grid = (words + 255) / 256;
    if (grid > (size_t)num_sms() * 32) grid = (size_t)num_sms() * 32;
    tggsw_fused_layout_kernel<<<(unsigned)grid, 256, 0, st>>>(g.R1, g.R1f, nd, items, ipt4, ct);
    tggsw_fused_layout_kernel<<<(unsigned)grid, 256, 0, st>>>(g.R2, g.R2f, nd, items, ipt4, ct);
    count_launch(2);
    FHE_CUDA_OK(cudaGetLastError());
    FHE_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

int extprod_fused_device(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, int cmux,
                         cudaStream_t st) {
    const int logn = g.tc->logn, k1 = (int)g.k + 1;
#define F(L, K) if (logn == L && k1 == K) return launch_fused<L, K>(g, ct1, ct2, out, batch, cmux, st);
    FHE_XP_SHAPES(F)
#undef F
    set_error("internal: fused external product called for an unsupported shape");
    return -1;
}

// acc_out[b] = chain of `steps` CMuxes over acc_in[b]; keys_dev = device array [steps][2] of fused key layouts
int cmux_chain_fused_device(const TorusCtx &tc, int k1, const u32 *const *keys_dev, const u64 *h_dev, int steps,
                            int negacyclic, const u64 *acc_in, u64 *acc_out, size_t batch, cudaStream_t st) {
    const int logn = tc.logn;
#define F(L, K) if (logn == L && k1 == K) return launch_chain<L, K>(tc, keys_dev, h_dev, steps, negacyclic, acc_in, acc_out, batch, st);
    FHE_XP_SHAPES(F)
#undef F
    set_error("internal: fused CMux chain called for an unsupported shape");
    return -1;
}

}  // namespace fhe
