// emu_ntt.cpp -- TEST INFRASTRUCTURE ONLY (never loaded by the product).
// Replays, on the CPU, the exact per-thread code the CUDA NTT kernels run (csrc/ntt_core.cuh +
// csrc/modarith.cuh are host+device), thread by thread and pass by pass, with a plain array standing
// in for shared memory.  The CPU test-suite compares this against the oracle, so index / twiddle /
// lazy-reduction bugs are caught without a GPU; the -m gpu tests then check the real kernels.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../../fhe_study_b200/csrc/ntt_core.cuh"
#include "../../fhe_study_b200/csrc/plan_host.hpp"
#include "../../fhe_study_b200/csrc/xp_octet.cuh"

using namespace fhe;

template <class M, int LOGN, int LOGE> struct Emu {
    typedef NttShape<LOGN, LOGE> S;
    typedef typename M::W W;
    typedef typename M::T T;
    std::vector<W> regs;  // [T][E]
    std::vector<W> mem;   // shared memory stand-in
    Emu() : regs((size_t)S::T * S::E), mem(S::N) {}
    W (&r(int tid))[S::E] { return *reinterpret_cast<W(*)[S::E]>(&regs[(size_t)tid * S::E]); }

    void exchange(int from, int to) {
        for (int t = 0; t < S::T; t++)
            for (int e = 0; e < S::E; e++) mem[S::pos(from, t, e)] = regs[(size_t)t * S::E + e];
        for (int t = 0; t < S::T; t++)
            for (int e = 0; e < S::E; e++) regs[(size_t)t * S::E + e] = mem[S::pos(to, t, e)];
    }
    void load(const u64 *g, int to) {
        for (int t = 0; t < S::T; t++)
            for (int e = 0; e < S::E; e++) regs[(size_t)t * S::E + e] = M::load(g[S::pos(0, t, e)]);
        if (to != 0) exchange(0, to);
    }
    void store(u64 *g, int from) {
        if (from != 0) exchange(from, 0);
        for (int t = 0; t < S::T; t++)
            for (int e = 0; e < S::E; e++) g[S::pos(0, t, e)] = M::store(regs[(size_t)t * S::E + e]);
    }
    template <int PASS> void fwd_from(const M &m, const TwSrc<M> &tw) {
        if (PASS > 0) exchange(PASS - 1, PASS);
        for (int t = 0; t < S::T; t++) fwd_pass<M, LOGN, LOGE, PASS>(r(t), t, m, tw);
        if constexpr (PASS + 1 < S::P) fwd_from<PASS + 1>(m, tw);
    }
    template <int PASS> void inv_from(const M &m, const TwSrc<M> &tw, T ninv, T s_ninv) {
        for (int t = 0; t < S::T; t++) inv_pass<M, LOGN, LOGE, PASS>(r(t), t, m, tw, ninv, s_ninv);
        if constexpr (PASS > 0) {
            exchange(PASS, PASS - 1);
            inv_from<PASS - 1>(m, tw, ninv, s_ninv);
        }
    }
    void fwd_canon(const M &m) { for (auto &v : regs) v = m.fwd_canon(v); }
    void fwd_out(const M &m) { for (auto &v : regs) v = m.fwd_out(v); }
    void canon2(const M &m) { for (auto &v : regs) v = m.canon2(v); }
};

template <class M, int LOGN, int LOGE>
static void run(int mode, const ExpandedTables<M> &x, const u64 *a, const u64 *b, u64 *c, u64 *c_evals, int flags) {
    typedef NttShape<LOGN, LOGE> S;
    constexpr int LAST = S::P - 1;
    const TwSrc<M> twf = {x.fwd.data(), x.fwd.data(), x.fwdw.data()};
    const TwSrc<M> twi = {x.inv.data(), x.inv.data(), x.invw.data()};
    Emu<M, LOGN, LOGE> A;
    if (mode == 0) {
        A.load(a, 0);
        A.template fwd_from<0>(x.mod, twf);
        A.fwd_canon(x.mod);
        A.store(c, LAST);
    } else if (mode == 1) {
        A.load(a, LAST);
        A.template inv_from<LAST>(x.mod, twi, x.ninv, x.s_ninv);
        A.canon2(x.mod);
        A.store(c, 0);
    } else {
        Emu<M, LOGN, LOGE> B;
        if (flags & 1) A.load(a, LAST);
        else { A.load(a, 0); A.template fwd_from<0>(x.mod, twf); A.fwd_out(x.mod); }
        if (flags & 2) B.load(b, LAST);
        else { B.load(b, 0); B.template fwd_from<0>(x.mod, twf); B.fwd_out(x.mod); }
        for (size_t i = 0; i < A.regs.size(); i++) B.regs[i] = x.mod.pw_mul(A.regs[i], B.regs[i]);
        if (c_evals) {
            for (size_t i = 0; i < A.regs.size(); i++) A.regs[i] = x.mod.pw_evals(B.regs[i]);
            A.store(c_evals, LAST);
        }
        B.template inv_from<LAST>(x.mod, twi, x.ninv_pw, x.s_ninv_pw);
        B.canon2(x.mod);
        B.store(c, 0);
    }
}

template <class M, int LOGN> struct Disp {
    static void go(int loge, int mode, const ExpandedTables<M> &x, const u64 *a, const u64 *b, u64 *c, u64 *ce, int fl) {
        switch (loge) {
#define C(LE) case LE: if constexpr (LE <= LOGN) run<M, LOGN, LE>(mode, x, a, b, c, ce, fl); break;
            C(1) C(2) C(3) C(4) C(5) C(6)
#undef C
        }
    }
};

template <class M>
static int emu_any(u64 q, u64 n, int loge, int mode, const u64 *a, const u64 *b, u64 *c, u64 *ce, int fl) {
    HostTables t;
    if (!build_host_tables(q, n, t).empty()) return -1;
    int logn = hp_ilog2(n);
    if (loge <= 0) loge = LogE<M>::of(logn);
    if (loge > logn) loge = logn;
    ExpandedTables<M> x;
    expand_tables(t, x, loge);
    switch (logn) {
#define C(L) case L: Disp<M, L>::go(loge, mode, x, a, b, c, ce, fl); return 0;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12) C(13) C(14) C(15)
#undef C
    }
    return -1;
}

// One digit transform of the external product (extprod_fused.cu: digit_ntt): a polynomial of BITS, pass 0 from the
// octet table (xp_octet.cuh), later passes the ordinary csub-free butterflies, final fold below 2^28.
template <int LOGN> static void run_digit(const ExpandedTables<Small32> &x, const u64 *bits, u64 *out) {
    typedef XpOct<LOGN> O;
    typedef typename O::S S;
    const TwSrc<Small32> twf = {x.fwd.data(), x.fwd.data()};
    std::vector<XpQuad> lo(256), hi(256);
    for (int b = 0; b < 256; b++) octet_table_entry(x.mod, twf, b, lo[b], hi[b]);
    Emu<Small32, LOGN, O::LOGE> A;
    for (int t = 0; t < S::T; t++) {
        u32 w = 0;
        for (int o = 0; o < 4; o++)
            for (int jj = 0; jj < 8; jj++) w |= (u32)(bits[S::pos(0, t, O::slot(o, jj))] & 1) << (8 * o + jj);
        digit_pass0<LOGN>(A.r(t), w, lo.data(), hi.data(), t, x.mod, twf);
    }
    if constexpr (S::P > 1) A.template fwd_from<1>(x.mod, twf);
    for (int t = 0; t < S::T; t++)
        for (int e = 0; e < S::E; e++) {
            const u32 v = A.regs[(size_t)t * S::E + e];
            out[S::pos(S::P - 1, t, e)] = v - (v >> 27) * x.mod.q;   // the kernel's fold: < 2^28, congruent mod q
        }
}

extern "C" {
// digit transform of the external product under a 27-bit CRT prime; out = lazy representatives (< 2^28) in NTT order
int emu_xp_digit(uint64_t q, uint64_t n, const uint64_t *bits, uint64_t *out) {
    HostTables t;
    if (!build_host_tables(q, n, t).empty()) return -1;
    ExpandedTables<Small32> x;
    expand_tables(t, x, 5);
    switch (hp_ilog2(n)) {
#define C(L) case L: run_digit<L>(x, bits, out); return 0;
        C(6) C(7) C(8) C(9) C(10) C(13)
#undef C
    }
    return -2;
}
// kind: -1 auto (as the library picks), 0 Lazy32, 1 Lazy64, 2 Strict64, 3 Small32, 4 Fermat32 ; mode 0 fwd, 1 inv, 2 mul
int emu_ntt(int kind, uint64_t q, uint64_t n, int loge, int mode, const uint64_t *a, const uint64_t *b, uint64_t *c,
            uint64_t *c_evals, int flags) {
    int logn = 0;
    while ((1ull << logn) < n) logn++;
    if (kind < 0) {
        kind = modulus_kind(q, logn);
        HostTables t;
        if (kind == 3 && build_host_tables(q, n, t).empty() && fermat_ok(t)) kind = 4;
    }
    if (kind == 4) {
        HostTables t;
        if (!build_host_tables(q, n, t).empty() || !fermat_ok(t)) return -2;
        return emu_any<Fermat32>(q, n, loge, mode, a, b, c, c_evals, flags);
    }
    if (kind == 0) return q < (1ull << 30) ? emu_any<Lazy32>(q, n, loge, mode, a, b, c, c_evals, flags) : -2;
    if (kind == 1) return q < (1ull << 62) ? emu_any<Lazy64>(q, n, loge, mode, a, b, c, c_evals, flags) : -2;
    if (kind == 3) return modulus_kind(q, logn) == 3 ? emu_any<Small32>(q, n, loge, mode, a, b, c, c_evals, flags) : -2;
    return emu_any<Strict64>(q, n, loge, mode, a, b, c, c_evals, flags);
}
int emu_plan(uint64_t q, uint64_t n, uint64_t *psi, uint64_t *n_inv, uint64_t *roots, uint64_t *roots_inv) {
    HostTables t;
    if (!build_host_tables(q, n, t).empty()) return -1;
    *psi = t.psi;
    *n_inv = t.n_inv;
    memcpy(roots, t.roots.data(), n * 8);
    memcpy(roots_inv, t.roots_inv.data(), n * 8);
    return 0;
}
// scalar checks of the modular policies: op 0 = mul (Barrett / double Montgomery)
uint64_t emu_modmul(int kind, uint64_t q, uint64_t a, uint64_t b) {
    if (kind == 0) { Lazy32 m; init_mod(m, q); return m.mul((u32)a, (u32)b); }
    if (kind == 1) { Lazy64 m; init_mod(m, q); return m.mul(a, b); }
    if (kind == 3) { Small32 m; init_mod(m, q); return m.mul((u32)a, (u32)b); }
    Strict64 m; init_mod(m, q); return m.mul(a, b);
}
}
