// fhe_b200.hpp -- C++ host-side mirror of the reference's Rust API for the hot path, over the C ABI of
// fhe_b200.h.  The reference's toolchain (Rust) is not available in this image, so this header plays the role
// of the `gpu`-feature branches of crates `arith` / `tfhe` / `bfv`: same type and method names, same argument
// meaning, same error behaviour (the reference panics on mismatched parameters -> std::runtime_error here).
//   arith::RingParam            arith/src/ring.rs:6-10
//   arith::Rq, NTT              arith/src/ring_nq.rs:19-27,406-607 ; arith/src/ntt.rs:44-110
//   arith::Tn                   arith/src/ring_torus.rs:24-27,118-132,153-327
//   tfhe::TGLWE / TGGSW         tfhe/src/tglwe.rs:30,89-119 ; tfhe/src/tggsw.rs:14,39-62
//   tfhe::TLWE / KSK / bootstrapping   tfhe/src/tlwe.rs:37-40,101-161
//   bfv::RLWE / RLK             bfv/src/lib.rs:35-47,59-90,251-271
//   gfhe::GLWE<Rq> / GLev<Rq> / KSK<Rq>   gfhe/src/glwe.rs:57-66,126-137,197-204,263-280 ; gfhe/src/glev.rs:14,67-80
// Header-only; link with -lfhe_b200.
#pragma once
#include <array>
#include <cstdint>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "fhe_b200.h"

namespace fhe_b200 {

inline void check(int rc) {
    if (rc != 0) throw std::runtime_error(std::string("fhe_b200: ") + fhe_last_error());
}

struct RingParam {
    uint64_t q;  // u64::MAX stands for the torus modulus 2^64 (arith/src/ring_torus.rs)
    size_t n;
    bool operator==(const RingParam &o) const { return q == o.q && n == o.n; }
    bool operator!=(const RingParam &o) const { return !(*this == o); }
};

namespace detail {
struct PlanDeleter { void operator()(fhe_ntt_plan *p) const { fhe_ntt_plan_destroy(p); } };
// The reference's (q, n) -> tables CACHE is permanent (arith/src/ntt.rs:18-38).  The library's own cache is
// reference-counted, so this mirror keeps one reference per (device, q, n) for the life of the process: otherwise every
// operator*, ntt and intt would drop the last reference and redo the root search, the table build, two cudaMalloc
// and two blocking uploads.
inline std::shared_ptr<fhe_ntt_plan> plan(const RingParam &p) {
    static std::mutex mu;
    static std::map<std::tuple<int, uint64_t, uint64_t>, std::shared_ptr<fhe_ntt_plan>> cache;
    int dev = 0;
    check(fhe_current_device(&dev));
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, p.q, (uint64_t)p.n);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    fhe_ntt_plan *h = nullptr;
    check(fhe_ntt_plan_create(p.q, p.n, &h));
    std::shared_ptr<fhe_ntt_plan> sp(h, PlanDeleter());
    cache.emplace(key, sp);
    return sp;
}
inline void same(const RingParam &a, const RingParam &b) {
    if (a != b) throw std::runtime_error("fhe_b200: ring parameter mismatch (the reference's assert_eq!(param))");
}
}  // namespace detail

// ---- Rq = Z_q[X]/(X^n+1) --------------------------------------------------------------------------------
class Rq {
  public:
    RingParam param;
    std::vector<uint64_t> coeffs;
    std::optional<std::vector<uint64_t>> evals;  // cached NTT of coeffs (ring_nq.rs:19-27)

    Rq() : param{0, 0} {}
    Rq(RingParam p, std::vector<uint64_t> c, std::optional<std::vector<uint64_t>> e = std::nullopt)
        : param(p), coeffs(std::move(c)), evals(std::move(e)) {}
    static Rq zero(const RingParam &p) { return Rq(p, std::vector<uint64_t>(p.n, 0)); }
    // Rq::from_vec_u64 (ring_nq.rs:55-63,132-141,156-159): reduce mod q, fold X^n = -1
    static Rq from_vec_u64(const RingParam &p, const std::vector<uint64_t> &v) {
        if (v.size() < p.n) throw std::runtime_error("fhe_b200: Rq::from_vec needs at least n coefficients");
        std::vector<uint64_t> out(p.n);
        check(fhe_rq_from_vec(p.q, p.n, v.data(), v.size(), out.data(), 1));
        return Rq(p, std::move(out));
    }
    void compute_evals() {  // ring_nq.rs:147-150
        std::vector<uint64_t> e(param.n);
        check(fhe_ntt_fwd(detail::plan(param).get(), coeffs.data(), e.data(), 1));
        evals = std::move(e);
    }
    bool operator==(const Rq &o) const { return param == o.param && coeffs == o.coeffs; }  // ignores evals (:401-405)

    Rq operator+(const Rq &o) const { return zip(o, fhe_rq_add); }
    Rq operator-(const Rq &o) const { return zip(o, fhe_rq_sub); }
    Rq operator-() const {
        Rq r = zero(param);
        check(fhe_rq_neg(param.q, coeffs.data(), r.coeffs.data(), param.n));
        return r;
    }
    Rq operator*(uint64_t s) const {  // mul_by_u64 (ring_nq.rs:274-281)
        Rq r = zero(param);
        check(fhe_rq_mul_u64(param.q, coeffs.data(), s, r.coeffs.data(), param.n));
        return r;
    }
    // ring_nq::mul (ring_nq.rs:586-607): reuses cached evals, returns the product WITH its evals
    Rq operator*(const Rq &o) const {
        detail::same(param, o.param);
        std::vector<uint64_t> c(param.n), ce(param.n);
        const int flags = (evals ? FHE_A_IS_EVALS : 0) | (o.evals ? FHE_B_IS_EVALS : 0);
        check(fhe_rq_mul(detail::plan(param).get(), evals ? evals->data() : coeffs.data(),
                         o.evals ? o.evals->data() : o.coeffs.data(), c.data(), 1, flags, ce.data()));
        return Rq(param, std::move(c), std::move(ce));
    }
    // ring_nq::mul_mut (ring_nq.rs:564-583): additionally caches the operands' evals
    static Rq mul_mut(Rq &a, Rq &b) {
        if (!a.evals) a.compute_evals();
        if (!b.evals) b.compute_evals();
        return a * b;
    }
    Rq remodule(uint64_t p) const { return map1(RingParam{p, param.n}, [&](uint64_t *o) { return fhe_rq_remodule(coeffs.data(), p, o, param.n); }); }
    Rq mod_switch(uint64_t p) const { return map1(RingParam{p, param.n}, [&](uint64_t *o) { return fhe_rq_mod_switch(param.q, coeffs.data(), p, o, param.n); }); }
    Rq mul_div_round(uint64_t num, uint64_t den) const { return map1(param, [&](uint64_t *o) { return fhe_rq_mul_div_round(param.q, coeffs.data(), num, den, o, param.n); }); }
    std::vector<Rq> decompose(uint32_t beta, uint32_t l) const {  // ring_nq.rs:67-77
        std::vector<uint64_t> d((size_t)l * param.n);
        check(fhe_rq_decompose(param.q, param.n, coeffs.data(), beta, l, d.data(), 1));
        std::vector<Rq> out;
        for (uint32_t j = 0; j < l; j++) out.emplace_back(param, std::vector<uint64_t>(d.begin() + j * param.n, d.begin() + (j + 1) * param.n));
        return out;
    }

  private:
    template <class F> Rq zip(const Rq &o, F f) const {
        detail::same(param, o.param);
        Rq r = zero(param);
        check(f(param.q, coeffs.data(), o.coeffs.data(), r.coeffs.data(), param.n));
        return r;
    }
    template <class F> static Rq map1(const RingParam &p, F f) {
        Rq r = zero(p);
        check(f(r.coeffs.data()));
        return r;
    }
};

struct NTT {  // arith/src/ntt.rs:44-110
    static Rq ntt(const Rq &a) {
        Rq r = Rq::zero(a.param);
        check(fhe_ntt_fwd(detail::plan(a.param).get(), a.coeffs.data(), r.coeffs.data(), 1));
        return r;
    }
    static Rq intt(const Rq &a) {
        Rq r = Rq::zero(a.param);
        check(fhe_ntt_inv(detail::plan(a.param).get(), a.coeffs.data(), r.coeffs.data(), 1));
        return r;
    }
};

// ---- Tn = T_q[X]/(X^n+1), q = 2^64 ------------------------------------------------------------------------
class Tn {
  public:
    RingParam param;
    std::vector<uint64_t> coeffs;
    Tn() : param{~0ull, 0} {}
    Tn(RingParam p, std::vector<uint64_t> c) : param(p), coeffs(std::move(c)) {}
    static Tn zero(const RingParam &p) { return Tn(p, std::vector<uint64_t>(p.n, 0)); }
    bool operator==(const Tn &o) const { return param == o.param && coeffs == o.coeffs; }
    Tn operator+(const Tn &o) const { Tn r = zero(param); detail::same(param, o.param); check(fhe_tn_add(coeffs.data(), o.coeffs.data(), r.coeffs.data(), param.n)); return r; }
    Tn operator-(const Tn &o) const { Tn r = zero(param); detail::same(param, o.param); check(fhe_tn_sub(coeffs.data(), o.coeffs.data(), r.coeffs.data(), param.n)); return r; }
    Tn operator-() const { Tn r = zero(param); check(fhe_tn_neg(coeffs.data(), r.coeffs.data(), param.n)); return r; }
    Tn operator*(uint64_t s) const { Tn r = zero(param); check(fhe_tn_mul_u64(coeffs.data(), s, r.coeffs.data(), param.n)); return r; }
    // naive_poly_mul (ring_torus.rs:251-298): exact negacyclic product mod 2^64
    Tn operator*(const Tn &o) const { Tn r = zero(param); detail::same(param, o.param); check(fhe_tn_mul(param.n, coeffs.data(), o.coeffs.data(), r.coeffs.data(), 1)); return r; }
    Tn left_rotate(uint64_t h) const {  // ring_torus.rs:118-132
        Tn r = zero(param);
        check(fhe_tn_left_rotate(param.n, coeffs.data(), &h, 1, r.coeffs.data(), 1));
        return r;
    }
    std::vector<Tn> decompose(uint32_t beta, uint32_t l) const {  // ring_torus.rs:67-77 (beta = 2 only, as torus.rs:43-52)
        if (beta != 2) throw std::runtime_error("fhe_b200: T64::decompose supports beta = 2 only");
        std::vector<uint64_t> d((size_t)l * param.n);
        check(fhe_tn_decompose(param.n, coeffs.data(), l, d.data(), 1));
        std::vector<Tn> out;
        for (uint32_t j = 0; j < l; j++) out.emplace_back(param, std::vector<uint64_t>(d.begin() + j * param.n, d.begin() + (j + 1) * param.n));
        return out;
    }
};

// ---- TFHE -----------------------------------------------------------------------------------------------
struct TGLWE {  // GLWE<Tn>: k mask polynomials then the body, flat (tfhe/src/tglwe.rs:30)
    size_t n, k;
    std::vector<uint64_t> data;  // (k+1)*n words
    TGLWE(size_t n_, size_t k_) : n(n_), k(k_), data((k_ + 1) * n_, 0) {}
    TGLWE(size_t n_, size_t k_, std::vector<uint64_t> d) : n(n_), k(k_), data(std::move(d)) {}
    TGLWE operator+(const TGLWE &o) const { TGLWE r(n, k); check(fhe_tn_add(data.data(), o.data.data(), r.data.data(), data.size())); return r; }
    TGLWE operator-(const TGLWE &o) const { TGLWE r(n, k); check(fhe_tn_sub(data.data(), o.data.data(), r.data.data(), data.size())); return r; }
    TGLWE left_rotate(uint64_t h) const {  // tglwe.rs:116-119
        TGLWE r(n, k);
        check(fhe_tn_left_rotate(n, data.data(), &h, k + 1, r.data.data(), k + 1));
        return r;
    }
    Tn decrypt(const std::vector<uint64_t> &sk) const {  // tglwe.rs:86-88: b - sum_i a_i * sk_i (sk = k polynomials, flat)
        if (sk.size() != k * n) throw std::runtime_error("fhe_b200: TGLWE::decrypt needs k*n key words");
        Tn p = Tn::zero(RingParam{~0ull, n});
        check(fhe_tglwe_decrypt(n, k, sk.data(), data.data(), p.coeffs.data(), 1));
        return p;
    }
    std::vector<uint64_t> sample_extraction(uint64_t h) const {  // tglwe.rs:89-115 -> TLWE of dimension k*n
        std::vector<uint64_t> out(k * n + 1);
        check(fhe_sample_extract(n, k, data.data(), h, out.data(), 1));
        return out;
    }
};

class TGGSW {  // tfhe/src/tggsw.rs:14
  public:
    size_t n, k;
    TGGSW(size_t n_, size_t k_, const std::vector<uint64_t> &rows) : n(n_), k(k_) {
        if (rows.size() != (k + 1) * 64 * (k + 1) * n) throw std::runtime_error("fhe_b200: TGGSW needs (k+1)*64*(k+1)*n words");
        fhe_tggsw *h = nullptr;
        check(fhe_tggsw_load(n, k, rows.data(), &h));
        h_.reset(h, [](fhe_tggsw *p) { fhe_tggsw_destroy(p); });
    }
    TGLWE operator*(const TGLWE &ct) const {  // impl Mul<TGLWE> for TGGSW (tggsw.rs:45-62)
        TGLWE r(n, k);
        check(fhe_extprod(h_.get(), ct.data.data(), r.data.data(), 1));
        return r;
    }
    static TGLWE cmux(const TGGSW &bit, const TGLWE &ct1, const TGLWE &ct2) {  // tggsw.rs:39-41
        TGLWE r(bit.n, bit.k);
        check(fhe_cmux(bit.h_.get(), ct1.data.data(), ct2.data.data(), r.data.data(), 1));
        return r;
    }
    const fhe_tggsw *handle() const { return h_.get(); }

  private:
    std::shared_ptr<fhe_tggsw> h_;
};

class KSK {  // tfhe/src/tlwe.rs:84-100
  public:
    size_t kn_in, kn_out, l;
    KSK(size_t kn_in_, size_t kn_out_, size_t l_, const std::vector<uint64_t> &rows) : kn_in(kn_in_), kn_out(kn_out_), l(l_) {
        if (rows.size() != kn_in * l * (kn_out + 1)) throw std::runtime_error("fhe_b200: KSK needs kn_in*l*(kn_out+1) words");
        fhe_ksk *h = nullptr;
        check(fhe_ksk_load(kn_in, kn_out, l, rows.data(), &h));
        h_.reset(h, [](fhe_ksk *p) { fhe_ksk_destroy(p); });
    }
    const fhe_ksk *handle() const { return h_.get(); }

  private:
    std::shared_ptr<fhe_ksk> h_;
};

struct TLWE {  // GLWE<T64>: kn mask words then b (tfhe/src/tlwe.rs:37-40)
    std::vector<uint64_t> data;
    explicit TLWE(std::vector<uint64_t> d) : data(std::move(d)) {}
    size_t kn() const { return data.size() - 1; }
    TLWE key_switch(const KSK &ksk) const {  // tlwe.rs:101-112
        std::vector<uint64_t> out(ksk.kn_out + 1);
        check(fhe_key_switch(ksk.handle(), data.data(), out.data(), 1));
        return TLWE(std::move(out));
    }
    uint64_t decrypt(const std::vector<uint64_t> &sk) const {  // tlwe.rs:80-82: the phase b - <a, sk>
        if (sk.size() != kn()) throw std::runtime_error("fhe_b200: TLWE::decrypt needs a key of the ciphertext's dimension");
        uint64_t p = 0;
        check(fhe_tlwe_decrypt(kn(), sk.data(), data.data(), &p, 1));
        return p;
    }
    static uint64_t decode(uint64_t t, uint64_t p) {  // tlwe.rs:60-63
        uint64_t r = 0, m = 0;
        check(fhe_tn_mul_div_round(&p, t, ~0ull, &r, 1));
        check(fhe_rq_remodule(&r, t, &m, 1));
        return m;
    }
    TLWE mod_switch(uint64_t q2) const {  // tlwe.rs:114-118
        std::vector<uint64_t> out(data.size());
        check(fhe_tlwe_mod_switch(data.data(), q2, out.data(), data.size()));
        return TLWE(std::move(out));
    }
};

// blind_rotation / bootstrapping as the reference executes them (tlwe.rs:121-161)
inline TGLWE blind_rotation(size_t n, size_t k, const TLWE &c, const TGLWE &table) {
    TGLWE r(n, k);
    check(fhe_blind_rotate(n, k, nullptr, 0, table.data.data(), c.data.data(), c.kn(), r.data.data(), 1));
    return r;
}
inline TLWE bootstrapping(size_t n, size_t k, const KSK &ksk, const TGLWE &table, const TLWE &c) {
    std::vector<uint64_t> out(ksk.kn_out + 1);
    check(fhe_bootstrap(n, k, ksk.handle(), table.data.data(), c.data.data(), c.kn(), out.data(), 1));
    return TLWE(std::move(out));
}

// blind rotation with one TGGSW per mask element: the loop tlwe.rs:138-147 spells out (extension, see fhe_b200.h)
inline TGLWE cmux_chain(const std::vector<TGGSW> &bsk, const TGLWE &acc, const std::vector<uint64_t> &h, bool negacyclic) {
    if (h.size() != bsk.size()) throw std::runtime_error("fhe_b200: cmux_chain needs one rotation per TGGSW");
    std::vector<const fhe_tggsw *> hs;
    for (const TGGSW &g : bsk) hs.push_back(g.handle());
    TGLWE r(acc.n, acc.k);
    check(fhe_cmux_chain(acc.n, acc.k, hs.data(), hs.size(), negacyclic ? 1 : 0, acc.data.data(), h.data(), r.data.data(), 1));
    return r;
}
inline TLWE bootstrapping_chain(size_t n, size_t k, const std::vector<TGGSW> &bsk, int mode, const KSK *ksk, const TGLWE &table,
                                const TLWE &c) {
    std::vector<const fhe_tggsw *> hs;
    for (const TGGSW &g : bsk) hs.push_back(g.handle());
    std::vector<uint64_t> out((ksk ? ksk->kn_out : k * n) + 1);
    check(fhe_bootstrap_chain(n, k, hs.data(), hs.size(), mode, ksk ? ksk->handle() : nullptr, table.data.data(), c.data.data(),
                              c.kn(), out.data(), 1));
    return TLWE(std::move(out));
}

// ---- gfhe over Rq -------------------------------------------------------------------------------------------
struct GLWE {  // GLWE<Rq>(TR<Rq>, Rq) (gfhe/src/glwe.rs:57): k mask polynomials, then the body
    RingParam param;
    size_t k;
    std::vector<uint64_t> data;  // (k+1)*n words
    GLWE(RingParam p, size_t k_) : param(p), k(k_), data((k_ + 1) * p.n, 0) {}
    GLWE(RingParam p, size_t k_, std::vector<uint64_t> d) : param(p), k(k_), data(std::move(d)) {
        if (data.size() != (k + 1) * p.n) throw std::runtime_error("fhe_b200: GLWE needs (k+1)*n words");
    }
    bool operator==(const GLWE &o) const { return param == o.param && k == o.k && data == o.data; }
    GLWE operator+(const GLWE &o) const { detail::same(param, o.param); GLWE r(param, k); check(fhe_rq_add(param.q, data.data(), o.data.data(), r.data.data(), data.size())); return r; }
    GLWE operator-(const GLWE &o) const { detail::same(param, o.param); GLWE r(param, k); check(fhe_rq_sub(param.q, data.data(), o.data.data(), r.data.data(), data.size())); return r; }
    GLWE mod_switch(uint64_t p) const {  // glwe.rs:197-204
        GLWE r(RingParam{p, param.n}, k);
        check(fhe_rq_mod_switch(param.q, data.data(), p, r.data.data(), data.size()));
        return r;
    }
    GLWE operator*(const Rq &plaintext) const {  // impl Mul<R> for GLWE<R> (glwe.rs:263-280): every component times the plaintext
        detail::same(param, plaintext.param);
        GLWE r(param, k);
        std::vector<uint64_t> rhs;
        for (size_t c = 0; c <= k; c++) rhs.insert(rhs.end(), plaintext.coeffs.begin(), plaintext.coeffs.end());
        check(fhe_rq_mul(detail::plan(param).get(), data.data(), rhs.data(), r.data.data(), k + 1, 0, nullptr));
        return r;
    }
};
class GLev {  // GLev<Rq>(Vec<GLWE<Rq>>) (gfhe/src/glev.rs:14), or the k*l rows of a KSK<Rq> (glwe.rs:99-125)
  public:
    RingParam param;
    size_t k, rows;
    GLev(RingParam p, size_t k_, const std::vector<GLWE> &glwes) : param(p), k(k_), rows(glwes.size()) {
        std::vector<uint64_t> flat;
        for (const GLWE &g : glwes) { detail::same(p, g.param); flat.insert(flat.end(), g.data.begin(), g.data.end()); }
        plan_ = detail::plan(p);
        fhe_rq_glev *h = nullptr;
        check(fhe_rq_glev_load(plan_.get(), k, rows, flat.data(), &h));
        h_.reset(h, [](fhe_rq_glev *x) { fhe_rq_glev_destroy(x); });
    }
    GLWE operator*(const std::vector<Rq> &v) const {  // impl Mul<Vec<R>> for GLev<R> (glev.rs:67-80)
        if (v.size() != rows) throw std::runtime_error("fhe_b200: GLev * Vec<R> needs one polynomial per level");
        std::vector<uint64_t> flat;
        for (const Rq &x : v) { detail::same(param, x.param); flat.insert(flat.end(), x.coeffs.begin(), x.coeffs.end()); }
        GLWE r(param, k);
        check(fhe_rq_glev_mul(h_.get(), flat.data(), r.data.data(), 1));
        return r;
    }
    GLWE key_switch(const GLWE &ct, uint32_t beta, uint32_t l) const {  // GLWE::key_switch (glwe.rs:126-137), *this = the KSK
        detail::same(param, ct.param);
        GLWE r(param, k);
        check(fhe_glwe_rq_key_switch(h_.get(), beta, l, ct.data.data(), r.data.data(), 1));
        return r;
    }

  private:
    std::shared_ptr<fhe_ntt_plan> plan_;
    std::shared_ptr<fhe_rq_glev> h_;
};

// ---- BFV --------------------------------------------------------------------------------------------------
struct RLWE {  // bfv/src/lib.rs:35-47
    Rq c0, c1;
    // BFV::decrypt (lib.rs:164-178): ((c0 + c1*s).mul_div_round(t, q)).remodule(t)
    static Rq decrypt(uint64_t t, const Rq &sk, const RLWE &c) {
        detail::same(sk.param, c.c0.param);
        const RingParam p = c.c0.param;
        std::vector<uint64_t> flat(c.c0.coeffs);
        flat.insert(flat.end(), c.c1.coeffs.begin(), c.c1.coeffs.end());
        Rq m = Rq::zero(RingParam{t, p.n});
        check(fhe_bfv_decrypt(detail::plan(p).get(), p.q, p.n, t, sk.coeffs.data(), flat.data(), m.coeffs.data(), 1));
        return m;
    }
    // RLWE::mul (lib.rs:87-90) = tensor + relinearize_204; rlk = (rlk0, rlk1) with coefficients mod pq
    static RLWE mul(uint64_t t, uint64_t pq, const std::pair<std::vector<uint64_t>, std::vector<uint64_t>> &rlk, const RLWE &a,
                    const RLWE &b) {
        detail::same(a.c0.param, b.c0.param);
        const RingParam p = a.c0.param;
        auto flat = [&](const RLWE &x) { std::vector<uint64_t> v(x.c0.coeffs); v.insert(v.end(), x.c1.coeffs.begin(), x.c1.coeffs.end()); return v; };
        std::vector<uint64_t> fa = flat(a), fb = flat(b), k(rlk.first), out(2 * p.n);
        k.insert(k.end(), rlk.second.begin(), rlk.second.end());
        check(fhe_bfv_mul_relin(p.q, p.n, t, pq, k.data(), fa.data(), fb.data(), out.data(), 1));
        return RLWE{Rq(p, std::vector<uint64_t>(out.begin(), out.begin() + p.n)), Rq(p, std::vector<uint64_t>(out.begin() + p.n, out.end()))};
    }
    // RLWE::tensor (lib.rs:59-85): (c0, c1, c2)
    static std::array<Rq, 3> tensor(uint64_t t, const RLWE &a, const RLWE &b) {
        detail::same(a.c0.param, b.c0.param);
        const RingParam p = a.c0.param;
        std::vector<uint64_t> fa = a.flat(), fb = b.flat(), out(3 * p.n);
        check(fhe_bfv_tensor(p.q, p.n, t, fa.data(), fb.data(), out.data(), 1));
        auto part = [&](size_t i) { return Rq(p, std::vector<uint64_t>(out.begin() + i * p.n, out.begin() + (i + 1) * p.n)); };
        return {part(0), part(1), part(2)};
    }
    // BFV::relinearize_204 (lib.rs:251-271)
    static RLWE relinearize_204(uint64_t pq, const std::vector<uint64_t> &rlk, const Rq &c0, const Rq &c1, const Rq &c2) {
        const RingParam p = c0.param;
        std::vector<uint64_t> c(c0.coeffs), out(2 * p.n);
        c.insert(c.end(), c1.coeffs.begin(), c1.coeffs.end());
        c.insert(c.end(), c2.coeffs.begin(), c2.coeffs.end());
        check(fhe_bfv_relinearize(p.q, p.n, pq, rlk.data(), c.data(), out.data(), 1));
        return from_flat(p, out);
    }
    // BFV::mul_const (lib.rs:189-200): m is a plaintext polynomial mod t
    static RLWE mul_const(uint64_t pq, const std::vector<uint64_t> &rlk, const RLWE &c, const Rq &m) {
        const RingParam p = c.c0.param;
        std::vector<uint64_t> fc = c.flat(), out(2 * p.n);
        check(fhe_bfv_mul_const(p.q, p.n, m.param.q, pq, rlk.data(), fc.data(), m.coeffs.data(), out.data(), 1));
        return from_flat(p, out);
    }
    // BFV::new_key / rlk_key / encrypt (lib.rs:120-160,202-225) with the device sampler (seeded twins of the rng-taking functions)
    static std::pair<Rq, std::vector<uint64_t>> new_key(const RingParam &p, double sigma, uint64_t seed) {
        Rq sk = Rq::zero(p);
        std::vector<uint64_t> pk(2 * p.n);
        check(fhe_bfv_keygen(detail::plan(p).get(), p.q, p.n, sigma, seed, sk.coeffs.data(), pk.data()));
        return {sk, pk};
    }
    static std::vector<uint64_t> rlk_key(const RingParam &p, uint64_t pmul, const Rq &sk, double sigma, uint64_t seed) {
        std::vector<uint64_t> rlk(2 * p.n);
        check(fhe_bfv_rlk_generate(p.q, p.n, pmul, sigma, seed, sk.coeffs.data(), rlk.data()));
        return rlk;
    }
    static RLWE encrypt(uint64_t t, const std::vector<uint64_t> &pk, const Rq &m, const RingParam &p, double sigma, uint64_t seed) {
        std::vector<uint64_t> out(2 * p.n);
        check(fhe_bfv_encrypt(detail::plan(p).get(), p.q, p.n, t, pk.data(), m.coeffs.data(), sigma, seed, out.data(), 1));
        return from_flat(p, out);
    }
    std::vector<uint64_t> flat() const {
        std::vector<uint64_t> v(c0.coeffs);
        v.insert(v.end(), c1.coeffs.begin(), c1.coeffs.end());
        return v;
    }
    static RLWE from_flat(const RingParam &p, const std::vector<uint64_t> &w) {
        return RLWE{Rq(p, std::vector<uint64_t>(w.begin(), w.begin() + p.n)), Rq(p, std::vector<uint64_t>(w.begin() + p.n, w.end()))};
    }
};

// compute_lookup_table (tfhe/src/tlwe.rs:196-214)
inline TGLWE compute_lookup_table(size_t n, size_t k, uint64_t t) {
    TGLWE table(n, k);
    check(fhe_compute_lookup_table(n, k, t, table.data.data()));
    return table;
}

// ---- CKKS over Rq (ckks/src/lib.rs:46-119); plaintexts are elements of R (int64 coefficients) ------------------------
struct CKKS {
    RingParam ring;
    std::pair<Rq, std::vector<uint64_t>> new_key(double sigma, uint64_t seed) const {  // lib.rs:46-63
        Rq sk = Rq::zero(ring);
        std::vector<uint64_t> pk(2 * ring.n);
        check(fhe_ckks_keygen(detail::plan(ring).get(), ring.q, ring.n, sigma, seed, sk.coeffs.data(), pk.data()));
        return {sk, pk};
    }
    RLWE encrypt(const std::vector<uint64_t> &pk, const std::vector<int64_t> &m, double sigma, uint64_t seed) const {  // lib.rs:66-84
        std::vector<uint64_t> out(2 * ring.n);
        check(fhe_ckks_encrypt(detail::plan(ring).get(), ring.q, ring.n, pk.data(), m.data(), sigma, seed, out.data(), 1));
        return RLWE::from_flat(ring, out);
    }
    std::vector<int64_t> decrypt(const Rq &sk, const RLWE &c) const {  // lib.rs:86-94
        std::vector<uint64_t> fc = c.flat();
        std::vector<int64_t> m(ring.n);
        check(fhe_ckks_decrypt(detail::plan(ring).get(), ring.q, ring.n, sk.coeffs.data(), fc.data(), m.data(), 1));
        return m;
    }
    RLWE add(const RLWE &a, const RLWE &b) const { return addsub(a, b, false); }  // lib.rs:113-115
    RLWE sub(const RLWE &a, const RLWE &b) const { return addsub(a, b, true); }   // lib.rs:116-118 (adds the second components, as written)
  private:
    RLWE addsub(const RLWE &a, const RLWE &b, bool sub) const {
        std::vector<uint64_t> fa = a.flat(), fb = b.flat(), out(2 * ring.n);
        check((sub ? fhe_ckks_sub : fhe_ckks_add)(ring.q, ring.n, fa.data(), fb.data(), out.data(), 1));
        return RLWE::from_flat(ring, out);
    }
};

}  // namespace fhe_b200
