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
    std::printf(failures ? "%d FAILURES\n" : "ALL OK\n", failures);
    return failures ? 1 : 0;
}
