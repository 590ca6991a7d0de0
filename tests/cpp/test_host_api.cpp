// test_host_api.cpp -- the reference's own unit tests for the hot path, re-stated against the C++ host mirror
// (include/fhe_b200.hpp) so that they read like the originals:
//   arith/src/ring_nq.rs:668-704 (test_mul: sage-generated vectors, via mul_mut)
//   arith/src/ring_nq.rs:627-665 (fold / add / sub), :707-729 (decompose)
//   arith/src/ntt.rs:194-234     (NTT round trips, n=4 and n=512)
//   arith/src/ring_torus.rs:334-366 (left_rotate)
// plus error behaviour (the reference panics where these throw).  Needs a GPU; run by tests/test_gpu_cpp_api.py.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "../../include/fhe_b200.hpp"

using namespace fhe_b200;
static int failures = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static std::vector<uint64_t> neg64(std::vector<int64_t> v) {
    std::vector<uint64_t> r;
    for (auto x : v) r.push_back((uint64_t)x);
    return r;
}

int main() {
    const uint64_t Q = 65537;
    {   // test_mul (ring_nq.rs:668-704)
        RingParam p{Q, 4};
        Rq a(p, {1, 2, 3, 4}), b(p, {1, 2, 3, 4});
        Rq c = Rq::mul_mut(a, b);
        EXPECT(c == Rq(p, {65513, 65517, 65531, 20}));
        EXPECT(a.evals.has_value() && c.evals.has_value());
        EXPECT((a * b) == c);  // operands now carry evals: no forward transform is run
        Rq d(p, {0, 0, 0, 2});
        EXPECT((d * d) == Rq(p, {0, 0, 65533, 0}));
        EXPECT(NTT::ntt(Rq(p, {1, 2, 3, 4})).coeffs == (std::vector<uint64_t>{7489, 56514, 17185, 49890}));
    }
    {   // poly_ring_zq_* (ring_nq.rs:627-665)
        EXPECT(Rq::from_vec_u64(RingParam{7, 3}, {0, 1, 2, 3, 4, 5}).coeffs == (std::vector<uint64_t>{4, 4, 4}));
        RingParam p{7, 4};
        Rq a(p, {1, 2, 3, 4}), b(p, {1, 2, 3, 6});
        EXPECT((a + b).coeffs == (std::vector<uint64_t>{2, 4, 6, 3}));
        EXPECT((a - b).coeffs == (std::vector<uint64_t>{0, 0, 0, 5}));
        EXPECT((-a).coeffs == (std::vector<uint64_t>{6, 5, 4, 3}));
    }
    {   // test_rq_decompose (ring_nq.rs:707-729): q=16, n=4, beta=4, l=2
        auto d = Rq(RingParam{16, 4}, {7, 14, 3, 6}).decompose(4, 2);
        EXPECT(d.size() == 2 && d[0].coeffs == (std::vector<uint64_t>{1, 3, 0, 1}) && d[1].coeffs == (std::vector<uint64_t>{3, 2, 3, 2}));
    }
    {   // test_ntt / test_ntt_loop (ntt.rs:194-234)
        std::mt19937_64 rng(1);
        for (size_t n : {size_t(4), size_t(512)}) {
            RingParam p{Q, n};
            for (int it = 0; it < (n == 4 ? 10 : 100); it++) {
                std::vector<uint64_t> v(n);
                for (auto &x : v) x = rng() % Q;
                Rq a(p, v);
                EXPECT(NTT::intt(NTT::ntt(a)) == a);
            }
        }
    }
    {   // test_left_rotate (ring_torus.rs:334-366)
        RingParam p{~0ull, 4};
        Tn f(p, neg64({2, 3, -4, -1}));
        EXPECT(f.left_rotate(3).coeffs == neg64({-1, -2, -3, 4}));
        EXPECT(f.left_rotate(1).coeffs == neg64({3, -4, -1, -2}));
        Tn one(p, {1, 0, 0, 0});
        EXPECT((f * one) == f);
        Tn x(p, {0, 1, 0, 0});  // multiplying by X is the inverse rotation: (f*X).left_rotate(1) == f
        EXPECT((f * x).left_rotate(1) == f);
    }
    {   // error behaviour: the reference panics (ntt.rs:116-130, ring_nq.rs:410,587)
        bool threw = false;
        try { NTT::ntt(Rq(RingParam{Q, 3}, {1, 2, 3})); } catch (const std::runtime_error &) { threw = true; }
        EXPECT(threw);
        threw = false;
        try { (void)(Rq(RingParam{Q, 4}, {1, 2, 3, 4}) + Rq(RingParam{7, 4}, {1, 2, 3, 4})); } catch (const std::runtime_error &) { threw = true; }
        EXPECT(threw);
    }
    {   // TGGSW (x) TGLWE and cmux agree with each other: cmux(g, ct, ct) == ct + g (x) 0 == ct
        const size_t n = 64, k = 1;
        std::mt19937_64 rng(2);
        std::vector<uint64_t> rows((k + 1) * 64 * (k + 1) * n), c((k + 1) * n);
        for (auto &x : rows) x = rng();
        for (auto &x : c) x = rng();
        TGGSW g(n, k, rows);
        TGLWE ct(n, k, c);
        EXPECT(TGGSW::cmux(g, ct, ct).data == ct.data);
        TGLWE zero(n, k);
        EXPECT((g * zero).data == zero.data);
    }
    {   // CMux chain (the loop of tlwe.rs:138-147): equals applying TGGSW::cmux + left_rotate step by step
        const size_t n = 64, k = 1, steps = 3;
        std::mt19937_64 rng(3);
        std::vector<TGGSW> bsk;
        for (size_t j = 0; j < steps; j++) {
            std::vector<uint64_t> rows((k + 1) * 64 * (k + 1) * n);
            for (auto &x : rows) x = rng();
            bsk.emplace_back(n, k, rows);
        }
        std::vector<uint64_t> c((k + 1) * n), h = {5, 0, 63};
        for (auto &x : c) x = rng();
        TGLWE acc(n, k, c), step = acc;
        for (size_t j = 0; j < steps; j++) step = TGGSW::cmux(bsk[j], step, step.left_rotate(h[j]));
        EXPECT(cmux_chain(bsk, acc, h, false).data == step.data);
    }
    {   // gfhe over Rq: GLev * Vec<Rq> equals the sum of GLWE * Rq (glev.rs:67-80), and key_switch with an all-zero
        // key is (0, b) (glwe.rs:126-137)
        RingParam p{Q, 16};
        const size_t k = 2, l = 3;
        std::mt19937_64 rng(4);
        auto rnd = [&](size_t len) { std::vector<uint64_t> v(len); for (auto &x : v) x = rng() % Q; return v; };
        std::vector<GLWE> rows;
        std::vector<Rq> v;
        for (size_t j = 0; j < l; j++) { rows.emplace_back(p, k, rnd((k + 1) * p.n)); v.emplace_back(p, rnd(p.n)); }
        GLWE want = rows[0] * v[0];
        for (size_t j = 1; j < l; j++) want = want + rows[j] * v[j];
        EXPECT((GLev(p, k, rows) * v) == want);
        std::vector<GLWE> zero_rows(k * 4, GLWE(p, k));
        GLWE ct(p, k, rnd((k + 1) * p.n));
        GLWE ks = GLev(p, k, zero_rows).key_switch(ct, 2, 4);
        GLWE expect(p, k);
        std::copy(ct.data.begin() + k * p.n, ct.data.end(), expect.data.begin() + k * p.n);
        EXPECT(ks == expect);
    }
    {   // decrypt entry points: a trivial TLWE (zero mask) decrypts to its body; a trivial TGLWE to its body polynomial;
        // a BFV ciphertext (m * floor(q/t), 0) to m
        std::vector<uint64_t> c(9, 0), sk(8, 1);
        c[8] = 5 * (~0ull / 16);
        TLWE tl(c);
        EXPECT(tl.decrypt(sk) == c[8] && TLWE::decode(16, tl.decrypt(sk)) == 5);
        TGLWE tg(64, 1);
        for (size_t x = 0; x < 64; x++) tg.data[64 + x] = x * (~0ull / 128);
        std::vector<uint64_t> z(64, 1);
        EXPECT(tg.decrypt(z).coeffs == std::vector<uint64_t>(tg.data.begin() + 64, tg.data.end()));
        RingParam p{Q, 16};
        std::vector<uint64_t> m(16), c0(16);
        for (size_t x = 0; x < 16; x++) { m[x] = x % 8; c0[x] = m[x] * (Q / 8); }
        RLWE ct{Rq(p, c0), Rq::zero(p)};
        EXPECT(RLWE::decrypt(8, Rq(p, std::vector<uint64_t>(16, 1)), ct).coeffs == m);
    }
    {   // test_mul_relin / test_tensor (bfv/src/lib.rs:504-601) with keys and ciphertexts generated on the device:
        // decrypt(mul(enc(m1), enc(m2))) == m1 * m2 in Z_t[X]/(X^n+1); mul == relinearize_204(tensor(..))
        RingParam p{Q, 16};
        const uint64_t t = 2, pmul = Q * Q, pq = pmul * Q;
        auto [sk, pk] = RLWE::new_key(p, 3.2, 11);
        std::vector<uint64_t> rlk = RLWE::rlk_key(p, pmul, sk, 3.2, 12);
        std::mt19937_64 rng(5);
        for (int it = 0; it < 20; it++) {
            std::vector<uint64_t> m1(16), m2(16), want(16, 0);
            for (auto &x : m1) x = rng() % t;
            for (auto &x : m2) x = rng() % t;
            for (size_t i = 0; i < 16; i++)
                for (size_t j = 0; j < 16; j++) {
                    const size_t d = (i + j) % 16;
                    const uint64_t v = m1[i] * m2[j] % t;
                    want[d] = (i + j >= 16) ? (want[d] + t - v) % t : (want[d] + v) % t;
                }
            RLWE c1 = RLWE::encrypt(t, pk, Rq(RingParam{t, 16}, m1), p, 3.2, 100 + it);
            RLWE c2 = RLWE::encrypt(t, pk, Rq(RingParam{t, 16}, m2), p, 3.2, 200 + it);
            std::vector<uint64_t> k0(rlk.begin(), rlk.begin() + 16), k1(rlk.begin() + 16, rlk.end());
            RLWE c3 = RLWE::mul(t, pq, {k0, k1}, c1, c2);
            auto tns = RLWE::tensor(t, c1, c2);
            RLWE c3b = RLWE::relinearize_204(pq, rlk, tns[0], tns[1], tns[2]);
            EXPECT(c3.flat() == c3b.flat());
            EXPECT(RLWE::decrypt(t, sk, c3).coeffs == want);
        }
    }
    {   // compute_lookup_table (tlwe.rs:196-214) at the bootstrapping test's parameters: a staircase of t plateaus
        TGLWE table = compute_lookup_table(1024, 1, 128);
        const uint64_t delta = ~0ull / 128;
        bool ok = true;
        for (size_t c = 0; c < 1024; c++) ok = ok && table.data[c] == 0 && table.data[1024 + c] == (c / 8) * delta;
        EXPECT(ok);
    }
    {   // CKKS Rq paths (ckks/src/lib.rs:46-119): decrypt(encrypt(m)) = m up to the noise, add is homomorphic
        CKKS ckks{RingParam{Q, 32}};
        auto [sk, pk] = ckks.new_key(3.2, 3);
        std::vector<int64_t> m0(32), m1(32);
        for (size_t i = 0; i < 32; i++) { m0[i] = (int64_t)(100 * i) - 1500; m1[i] = 7 - (int64_t)i * 13; }
        RLWE c0 = ckks.encrypt(pk, m0, 3.2, 5), c1 = ckks.encrypt(pk, m1, 3.2, 6);
        std::vector<int64_t> d0 = ckks.decrypt(sk, c0), ds = ckks.decrypt(sk, ckks.add(c0, c1));
        bool ok = true;
        for (size_t i = 0; i < 32; i++) ok = ok && std::llabs(d0[i] - m0[i]) < 200 && std::llabs(ds[i] - (m0[i] + m1[i])) < 400;
        EXPECT(ok);
    }
    std::printf(failures ? "%d FAILURES\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
