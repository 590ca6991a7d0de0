// plan_host.hpp -- host-side construction of an NTT plan for one (q, n): the reference's deterministic
// root search and twiddle tables (arith/src/ntt.rs:18-38,115-185) plus the Shoup / Barrett / Montgomery
// companions the kernels use.  Pure C++ (also compiled into tests/emu).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "ntt_core.cuh"

namespace fhe {

typedef unsigned __int128 u128_t;

inline u64 hp_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128_t)a * b) % q); }
inline u64 hp_powmod(u64 x, u64 k, u64 q) {
    u64 r = 1 % q;
    x %= q;
    while (k) {
        if (k & 1) r = hp_mulmod(r, x, q);
        x = hp_mulmod(x, x, q);
        k >>= 1;
    }
    return r;
}
inline bool hp_is_prime(u64 n) {  // deterministic Miller-Rabin for 64-bit integers
    if (n < 2) return false;
    const u64 bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : bases)
        if (n % p == 0) return n == p;
    u64 d = n - 1;
    int s = 0;
    while ((d & 1) == 0) { d >>= 1; s++; }
    for (u64 a : bases) {
        u64 x = hp_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (int i = 1; i < s; i++) {
            x = hp_mulmod(x, x, n);
            if (x == n - 1) { comp = false; break; }
        }
        if (comp) return false;
    }
    return true;
}
inline int hp_ilog2(u64 n) {
    int l = 0;
    while (n > 1) { n >>= 1; l++; }
    return l;
}
inline u32 hp_bitrev(u32 i, int bits) {
    u32 r = 0;
    for (int b = 0; b < bits; b++) r |= ((i >> b) & 1u) << (bits - 1 - b);
    return r;
}

struct HostTables {
    u64 q = 0, n = 0, psi = 0, n_inv = 0;
    std::vector<u64> roots, roots_inv;  // reference order: roots[i] = psi^bitrev(i)
};

// Returns "" on success, else the reason (the reference panics in these cases, arith/src/ntt.rs:116-130).
inline std::string build_host_tables(u64 q, u64 n, HostTables &t) {
    if (n < 2 || (n & (n - 1)) != 0) return "n must be a power of two >= 2";
    if (q < 3 || q >= (1ull << 63)) return "q must satisfy 3 <= q < 2^63 (Zq::add is an un-widened u64 add)";
    if ((q - 1) % (2 * n) != 0) return "2n must divide q-1";
    if (!hp_is_prime(q)) return "q must be prime";
    // arith/src/ntt.rs:115-131: smallest k >= 1 with w = k^((q-1)/2n) and w^n != 1
    u64 psi = 0;
    for (u64 k = 1; k < q; k++) {
        u64 w = hp_powmod(k, (q - 1) / (2 * n), q);
        if (hp_powmod(w, n, q) != 1) { psi = w; break; }
    }
    if (psi == 0) return "no primitive 2n-th root of unity";
    const int logn = hp_ilog2(n);
    const u64 psi_inv = hp_powmod(psi, q - 2, q);
    std::vector<u64> pw(n), pwi(n);
    pw[0] = pwi[0] = 1;
    for (u64 i = 1; i < n; i++) {
        pw[i] = hp_mulmod(pw[i - 1], psi, q);
        pwi[i] = hp_mulmod(pwi[i - 1], psi_inv, q);
    }
    t.q = q; t.n = n; t.psi = psi;
    t.roots.resize(n);
    t.roots_inv.resize(n);
    for (u64 i = 0; i < n; i++) {
        u32 r = hp_bitrev((u32)i, logn);
        t.roots[i] = pw[r];       // psi^bitrev(i)            (ntt.rs:133-147)
        t.roots_inv[i] = pwi[r];  // (psi^bitrev(i))^-1       (ntt.rs:149-161, same value as the Fermat inverse)
    }
    t.n_inv = hp_powmod(n % q, q - 2, q);  // ntt.rs:27-30
    return "";
}

inline Tw32 make_tw(u32 w, u32 q, Tw32 *) { return Tw32{w, (u32)((((u64)w) << 32) / q)}; }
inline Tw64 make_tw(u64 w, u64 q, Tw64 *) { return Tw64{w, (u64)((((u128_t)w) << 64) / q)}; }

inline void init_mod(Lazy32 &m, u64 q) {
    m.q = (u32)q;
    m.q2 = (u32)(2 * q);
    int k = hp_ilog2(q) + 1;  // 2^(k-1) <= q < 2^k
    m.bk_shift = (u32)(k - 1);
    m.bk_mu = (u32)((((u128_t)1) << (2 * k)) / q);
}
inline u64 neg_inv64(u64 q) {  // -q^-1 mod 2^64 (q odd)
    u64 x = q;                 // correct to 3 bits
    for (int i = 0; i < 6; i++) x *= 2 - q * x;
    return (u64)0 - x;
}
inline void init_mod(Lazy64 &m, u64 q) {
    m.q = q;
    m.q2 = 2 * q;
    m.qinv_neg = neg_inv64(q);
    m.qinv = (u64)0 - m.qinv_neg;
    u64 r = (u64)((((u128_t)1) << 64) % q);
    m.r2 = hp_mulmod(r, r, q);
    m.nq = (u64)0 - q;
}
inline void init_mod(Strict64 &m, u64 q) {
    m.q = q;
    m.q2 = 0;
    m.qinv_neg = neg_inv64(q);
    m.qinv = (u64)0 - m.qinv_neg;
    u64 r = (u64)((((u128_t)1) << 64) % q);
    m.r2 = hp_mulmod(r, r, q);
    m.nq = (u64)0 - q;
}

inline u32 neg_inv32(u32 q) {
    u32 x = q;
    for (int i = 0; i < 5; i++) x *= 2 - q * x;
    return (u32)0 - x;
}
inline void init_mod(Small32 &m, u64 q) {
    m.q = (u32)q;
    m.q2 = (u32)(2 * q);
    m.qinv_neg = neg_inv32((u32)q);
    m.qinv = (u32)0 - m.qinv_neg;
    m.one = make_tw((u32)1, (u32)q, (Tw32 *)nullptr);
    m.r = make_tw((u32)((1ull << 32) % q), (u32)q, (Tw32 *)nullptr);
    for (int k = 0; k < 16; k++) m.qk[k] = (u32)((2 * q) << k);  // wraps only where the policy is not selected
}

inline void init_mod(Fermat32 &m, u64 q) {
    init_mod(static_cast<Small32 &>(m), q);
    m.q4 = (u32)(4 * q);
    m.c10 = (u32)(1024 * q);
    for (int k = 0; k < 8; k++) m.okb[k] = (u32)(q << (k + 10));  // used for k <= Fermat32::INV_KB_MAX only
    m.c8 = (u32)(512 * q);
    m.c14 = (u32)(16384 * q);
}
// Fermat32 (radix-4 butterflies with a shift for the fourth twiddle product) is selected on top of kind 3 when the
// modulus is 2^16 + 1 and the square root of -1 in the table is the one the policy was written for: roots[1] =
// psi^(n/2) = -2^8.  The reference's root search gives that for every n (psi = 3^(32768/n)).
inline bool fermat_ok(const HostTables &t) {
    if (!(t.q == 65537 && t.n >= 2 && t.n <= (1u << 15) && t.roots[1] == t.q - 256 && t.roots_inv[1] == 256)) return false;
    return t.n < 4 || (t.roots[2] == 4096 && t.roots[3] == 16);  // the shift-only first layer (Fermat32::fwd4_first)
}

// 3: Small32 (q < 2^22 and 2q*n <= 2^32: both transforms free of conditional subtractions), 0: Lazy32 (q < 2^30),
// 1: Lazy64 (q < 2^62), 2: Strict64 (q < 2^63)
inline int modulus_kind(u64 q, int logn) {
    if (q < (1ull << 22) && ((2 * q) << logn) <= (1ull << 32)) return 3;
    return q < (1ull << 30) ? 0 : q < (1ull << 62) ? 1 : 2;
}

// Shoup-expanded tables for policy M.
template <class M> struct ExpandedTables {
    std::vector<typename M::T> fwd, inv;
    std::vector<u32> fwdw, invw;       // radix-4 policies: the twiddles alone, same (device) order
    typename M::T ninv, s_ninv;        // standalone inverse transform
    typename M::T ninv_pw, s_ninv_pw;  // inverse transform after M::pw_mul (absorbs its 2^-wordbits factor)
    M mod;
};
// Radix-4 policies: the slot of the odd child roots[2i+1] of every paired stage holds roots[i] * roots[2i] instead
// (tab in REFERENCE order; forward pairs are the local stages (0,1), (2,3), ... of each pass, inverse pairs run from the
// top of each pass with the transform's last stage left alone -- the same rules as fwd_pass / InvSched in ntt_core.cuh).
inline void radix4_patch(std::vector<u64> &tab, u64 q, int logn, int loge, bool inverse) {
    const int P = ntt_num_passes(logn, loge);
    const std::vector<u64> ref = tab;
    for (int p = 0; p < P; p++) {
        const int s0 = ntt_pass_s0(logn, loge, p), g = (p + 1 < P ? ntt_pass_s0(logn, loge, p + 1) : logn) - s0;
        std::vector<int> parents;  // local stage of the parent of each pair
        if (!inverse) {
            for (int ls = 0; ls + 1 < g; ls += 2) parents.push_back(ls);
        } else {
            const int lo = p == 0 ? 1 : 0;
            for (int ls = g - 1; ls - 1 >= lo; ls -= 2) parents.push_back(ls - 1);
        }
        for (int lp : parents) {
            const int s = s0 + lp;
            for (u64 j = 0; j < (1ull << s); j++)
                tab[(2ull << s) + 2 * j + 1] = hp_mulmod(ref[(1ull << s) + j], ref[(2ull << s) + 2 * j], q);
        }
    }
}

// loge <= 0 selects the library's policy LogE<M>; tables are stored in device order (ntt_core.cuh: tw_slot).
template <class M> void expand_tables(const HostTables &t, ExpandedTables<M> &x, int loge = 0) {
    typedef typename M::W W;
    typedef typename M::T T;
    init_mod(x.mod, t.q);
    x.fwd.resize(t.n);
    x.inv.resize(t.n);
    const int logn = hp_ilog2(t.n);
    if (loge <= 0) loge = LogE<M>::of(logn);
    if (loge > logn) loge = logn;
    std::vector<u64> rf = t.roots, ri = t.roots_inv;
    if (M::RADIX4) {
        radix4_patch(rf, t.q, logn, loge, false);
        radix4_patch(ri, t.q, logn, loge, true);
    }
    for (u64 i = 0; i < t.n; i++) {
        const u64 slot = tw_slot(logn, loge, i, sizeof(T) == 8);
        x.fwd[slot] = make_tw((W)rf[i], (W)t.q, (T *)nullptr);
        x.inv[slot] = make_tw((W)ri[i], (W)t.q, (T *)nullptr);
    }
    if (M::RADIX4) {
        x.fwdw.resize(t.n);
        x.invw.resize(t.n);
        for (u64 i = 0; i < t.n; i++) {
            x.fwdw[i] = (u32)x.fwd[i].w;
            x.invw[i] = (u32)x.inv[i].w;
        }
    }
    x.ninv = make_tw((W)t.n_inv, (W)t.q, (T *)nullptr);
    x.s_ninv = make_tw((W)hp_mulmod(t.roots_inv[1], t.n_inv, t.q), (W)t.q, (T *)nullptr);
    u64 comp = 1;
    if (M::PW_SCALED) comp = (u64)((((u128_t)1) << (8 * sizeof(W))) % t.q);
    const u64 ninv_pw = hp_mulmod(t.n_inv, comp, t.q);
    x.ninv_pw = make_tw((W)ninv_pw, (W)t.q, (T *)nullptr);
    x.s_ninv_pw = make_tw((W)hp_mulmod(t.roots_inv[1], ninv_pw, t.q), (W)t.q, (T *)nullptr);
}

}  // namespace fhe
