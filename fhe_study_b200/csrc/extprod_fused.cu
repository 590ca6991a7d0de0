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

// threads per CTA: 256, or (tuning experiment FHE_XP_CT512) 512 at n = 1024, k = 1 with two accumulators per CTA that
// share every key load at unchanged occupancy (16 warps, 128 registers, 132 KB shared memory per SM)
#ifndef FHE_XP_CT512
#define FHE_XP_CT512 0
#endif
__host__ __device__ constexpr int xp_ct(int logn, int k1) { return (logn == 10 && k1 == 2 && FHE_XP_CT512) ? 512 : 256; }

template <int LOGN, int K1> struct XpGeom {
    static constexpr int N = 1 << LOGN;
    static constexpr int LOGE = LOGN < 5 ? LOGN : 5;
    typedef NttShape<LOGN, LOGE> S;
    static constexpr int CT = xp_ct(LOGN, K1);
    static constexpr int SLOTS = CT / S::T;             // concurrent NTTs
    static constexpr int ND = K1 * 64;                  // digit polynomials per accumulator
    static constexpr int PADN = N + (N >> 5);
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

// items m and m' of a thread sit on the same coefficient position when CT * (m - m') is a multiple of N: with
// CT * 4 >= N (every instantiated shape) there are at most 4 distinct positions, selected by m mod (N / CT)
template <int LOGN, int CT> __host__ __device__ constexpr int dv_slot(int m) {
    return (1 << LOGN) >= CT ? m % ((1 << LOGN) / CT) : 0;
}

// 128-bit read-only key load that stays where the source puts it: as a plain __ldg ptxas sank the loads issued in
// front of the barrier back behind it
__device__ __forceinline__ uint4 ldg_key(const uint4 *p) {
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

struct XpParams {
    NttParams<Lazy32> P[2];   // plans of p1, p2 (device-order tables, n^-1 constants)
    Small32 ms[2];            // same moduli, csub-free forward butterflies
    u64 mu[2];                // floor(2^64 / p_r): Barrett constant for the 64-bit accumulators
    const u32 *R[2];          // fused key layout: R[r][d][v][t][j] = NTT value of item t + 256 (4v + j) of digit d
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
template <int LOGN, int K1>
__device__ __forceinline__ void digit_ntt(const Small32 &ms, const TwSrc<Small32> &twf, u32 w, const uint4 *tab_lo,
                                          const uint4 *tab_hi, u32 *sm, int tid) {
    typedef XpGeom<LOGN, K1> G;
    constexpr int LOGE = G::LOGE;
    typedef NttShape<LOGN, LOGE> S;
    constexpr int LAST = S::P - 1;
    u32 x[S::E];
    digit_pass0<LOGN>(x, w, tab_lo, tab_hi, tid, ms, twf);
    if constexpr (S::P > 1) fwd_chain<Small32, LOGN, LOGE, 1>(x, sm, tid, ms, twf);
#pragma unroll
    for (int e = 0; e < S::E; e++) sm[pad_idx(S::pos(LAST, tid, e))] = x[e] - (x[e] >> 27) * ms.q;
}

template <int LOGN, int K1, bool CHAIN>
__global__ void __launch_bounds__(XpGeom<LOGN, K1>::CT, XpGeom<LOGN, K1>::MINB)
extprod_fused_kernel(const __grid_constant__ XpParams X, const u64 *__restrict__ ct1, const u64 *__restrict__ ct2,
                     u64 *out, int cmux, const XpChain ch, size_t batch) {
    typedef XpGeom<LOGN, K1> G;
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
            const TwSrc<Small32> twf = {X.P[r].c_fwd, X.P[r].fwd};
            octet_table_entry(X.ms[r], twf, t, tab[r * 512 + t], tab[r * 512 + 256 + t]);
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
                const TwSrc<Small32> twf = {X.P[r].c_fwd, X.P[r].fwd};
                const u32 w = planes[((size_t)(s_acc * K1 + (d >> 6)) * S::T + tid) * G::PLANE_ROW + (d & 63)];
                digit_ntt<LOGN, K1>(X.ms[r], twf, w, tab + r * 512, tab + r * 512 + 256, sm, tid);
            }
            // MAC of the round's digits against the resident TGGSW, two digits per iteration.  The 128-bit key loads are
            // software-pipelined through the registers themselves: the loads of the first pair are issued BEFORE the
            // barrier that ends the transforms (the transform registers are dead by then), and every key quad is
            // re-loaded for the next pair right behind the four MACs that consumed it.  Before, an iteration issued
            // its loads and sat on their L2 latency: 17 % of all stall samples on the first IMAD.WIDE of the loop
            // (profiles/r2_extprod_fused_n1024_k1_ncu_full_b.csv).
            constexpr int nd = G::DPR, Q = G::IPT4 / 4;   // ND % DPR == 0: every round is full
            static_assert(nd % 2 == 0, "the MAC takes the digits in pairs");
            const uint4 *Rt = reinterpret_cast<const uint4 *>(Rr) + (size_t)round * G::DPR * Q * G::CT + t;
            uint4 kq[2][Q];
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++)
#pragma unroll
                for (int v = 0; v < Q; v++) kq[h2][v] = ldg_key(Rt + (size_t)(h2 * Q + v) * G::CT);
            __syncthreads();
#pragma unroll(nd <= 8 ? nd / 2 : 1)
            for (int it = 0; it < nd / 2; it++) {
                const bool more = it + 1 < nd / 2;
#pragma unroll
                for (int h2 = 0; h2 < 2; h2++) {
                    const int dd = 2 * it + h2;
                    u32 dv[A][G::IPT < 4 ? G::IPT : 4];   // items m, m + 4, .. of a thread share the position
#pragma unroll
                    for (int aa = 0; aa < A; aa++)
#pragma unroll
                        for (int m = 0; m < (G::IPT < 4 ? G::IPT : 4); m++)
                            dv[aa][m] = xch[(size_t)(dd * A + aa) * G::PADN + pad_idx((t + G::CT * m) & (N - 1))];
#pragma unroll
                    for (int v = 0; v < Q; v++) {
                        const u32 k4[4] = {kq[h2][v].x, kq[h2][v].y, kq[h2][v].z, kq[h2][v].w};
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const int m = 4 * v + j;
                            if (m < G::IPT) {   // item >= ITEMS only when ITEMS % CT != 0: key padding is zero
#pragma unroll
                                for (int aa = 0; aa < A; aa++) acc[aa][m] += (u64)dv[aa][dv_slot<LOGN, G::CT>(m)] * k4[j];
                            }
                        }
                        if (more) kq[h2][v] = ldg_key(Rt + (size_t)((dd + 2) * Q + v) * G::CT);
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
                const int item = t + G::CT * m;
                if (item < G::ITEMS)
                    xch[(size_t)(aa * G::UNITS + (item >> LOGN)) * G::PADN + pad_idx(item & (N - 1))] = reduce64(acc[aa][m], ml.q, X.mu[r]);
            }
        __syncthreads();
        // warp-uniform condition: every lane of a warp that owns at least one live slot runs the transform
        // (the exchanges inside synchronise whole warps); lanes of dead slots compute on scratch and store nothing
        if ((t & ~31) / S::T < A * G::UNITS) {
            const TwSrc<Lazy32> twi = {X.P[r].c_inv, X.P[r].inv};
            u32 x[S::E];
#pragma unroll
            for (int e = 0; e < S::E; e++) x[e] = sm[pad_idx(S::pos(LAST, tid, e))];
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

// unfused key layout (u64, [d][u][x]) -> fused layout (u32, [d][v][t][j], item = t + 256 (4v + j)), zero padded
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

template <int LOGN, int K1>
static int launch_fused(const Tggsw &g, const u64 *ct1, const u64 *ct2, u64 *out, size_t batch, int cmux, cudaStream_t st) {
    typedef XpGeom<LOGN, K1> G;
    const TorusCtx &tc = *g.tc;
    XpParams X;
    X.P[0] = tc.plan1->p32;
    X.P[1] = tc.plan2->p32;
    init_mod(X.ms[0], TORUS_P1);
    init_mod(X.ms[1], TORUS_P2);
    X.mu[0] = ~0ull / TORUS_P1;
    X.mu[1] = ~0ull / TORUS_P2;
    X.R[0] = g.R1f;
    X.R[1] = g.R2f;
    X.cp = tc.cp;
    FHE_REQUIRE(batch <= 0x7fffffffull, "extprod: batch too large");
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    auto kern = extprod_fused_kernel<LOGN, K1, false>;
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
static int launch_chain(const TorusCtx &tc, const u32 *const *keys_dev, const u64 *h_dev, int steps, int negacyclic,
                        const u64 *acc_in, u64 *acc_out, size_t batch, cudaStream_t st) {
    typedef XpGeom<LOGN, K1> G;
    XpParams X;
    X.P[0] = tc.plan1->p32;
    X.P[1] = tc.plan2->p32;
    init_mod(X.ms[0], TORUS_P1);
    init_mod(X.ms[1], TORUS_P2);
    X.mu[0] = ~0ull / TORUS_P1;
    X.mu[1] = ~0ull / TORUS_P2;
    X.R[0] = X.R[1] = nullptr;
    X.cp = tc.cp;
    FHE_REQUIRE(batch <= 0x7fffffffull, "cmux chain: batch too large");
    static unsigned long long done_mask = 0;
    int dev = 0;
    FHE_CUDA_OK(cudaGetDevice(&dev));
    auto kern = extprod_fused_kernel<LOGN, K1, true>;
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

#define FHE_XP_SHAPES(F) F(10, 2) F(9, 2) F(8, 2) F(6, 5) F(6, 2) F(7, 2) F(8, 3) F(9, 3)

bool extprod_fused_supported(int logn, int k1) {
#define F(L, K) if (logn == L && k1 == K) return true;
    FHE_XP_SHAPES(F)
#undef F
    return false;
}
static int fused_ipt4(int logn, int k1) {
    const int items = k1 * 2 * (1 << logn), ct = xp_ct(logn, k1), ipt = (items + ct - 1) / ct;
    return (ipt + 3) / 4 * 4;
}

// builds the fused key layout from the unfused one (called once at load when the shape is supported)
int tggsw_build_fused_layout(Tggsw &g, cudaStream_t st) {
    const int logn = g.tc->logn, k1 = (int)g.k + 1;
    if (!extprod_fused_supported(logn, k1)) return 0;
    // the fused kernels read the plans' twiddle tables in the device order of 32 coefficients per thread
    if (g.tc->plan1->loge != (logn < 5 ? logn : 5) || g.tc->plan2->loge != g.tc->plan1->loge) return 0;
    const int nd = k1 * 64, items = k1 * 2 * (1 << logn), ipt4 = fused_ipt4(logn, k1);
    const int ct = xp_ct(logn, k1);
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
