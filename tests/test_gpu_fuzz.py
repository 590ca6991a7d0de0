"""Seeded random shapes through the C ABI against the oracle: the shapes nobody listed by hand (odd batch sizes,
unusual k / l / beta, degrees outside the fused kernels' lists, non-power-of-two BFV degrees)."""
import os

import numpy as np
import pytest

from primes import Q17, Q22, Q30, Q62, Q63

pytestmark = pytest.mark.gpu

# FHE_FUZZ_SEEDS=<count> widens every seed range below for a soak run (tools/run_fuzz_soak.sh; the log of the last one is
# profiles/r2_fuzz_soak.log); the default ranges keep the suite under a minute
SOAK = int(os.environ.get("FHE_FUZZ_SEEDS", "0"))


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


@pytest.mark.parametrize("seed", range(max(12, SOAK)))
def test_random_ntt_and_polymul(fhe, orc, seed):
    rng = np.random.default_rng(seed)
    q = [Q17, Q22, Q30, Q62, Q63, 12289, 7681][rng.integers(0, 7)]
    max_logn = {12289: 11, 7681: 8}.get(q, 14 if q > 2**32 else 15)
    logn = int(rng.integers(1, min(max_logn, 13) + 1))
    n, batch = 1 << logn, int(rng.integers(1, 20))
    a, b = orc.uniform(seed, (batch, n), q), orc.uniform(seed + 100, (batch, n), q)
    plan = fhe.NttPlan(q, n)
    fa = plan.ntt(a)
    assert np.array_equal(fa, orc.ntt(q, n, a))
    assert np.array_equal(plan.intt(fa), a)
    ev = np.empty_like(a)
    c = plan.mul(a, b, evals_out=ev)
    assert np.array_equal(c, orc.rq_mul_batch(q, n, a, b))
    assert np.array_equal(ev, orc.ntt(q, n, c))
    assert np.array_equal(plan.mul(fa, b, flags=fhe.A_IS_EVALS), c)


@pytest.mark.parametrize("seed", range(max(10, SOAK)))
def test_random_torus_shapes(fhe, orc, seed):
    rng = np.random.default_rng(100 + seed)
    n = 1 << int(rng.integers(1, 12))
    k = int(rng.integers(1, 5 if n <= 256 else 3))
    batch = int(rng.integers(1, 7))
    glwe = (k + 1) * n
    a, b = orc.uniform(seed, (batch, n)), orc.uniform(seed + 1, (batch, n))
    assert np.array_equal(fhe.tn_mul(n, a, b), orc.tn_mul(n, a, b, threads=4))
    if n * (k + 1) <= 4096:  # keep the O(N^2)-free fast oracle quick
        tggsw = orc.uniform(seed + 2, (k + 1) * 64 * glwe)
        ct1, ct2 = orc.uniform(seed + 3, (batch, glwe)), orc.uniform(seed + 4, (batch, glwe))
        g = fhe.Tggsw(n, k, tggsw)
        assert np.array_equal(g.extprod(ct1), orc.extprod(n, k, tggsw, ct1))
        assert np.array_equal(g.cmux(ct1, ct2), orc.cmux(n, k, tggsw, ct1, ct2))
        h = orc.uniform(seed + 5, (batch, 2)) % np.uint64(4 * n)
        neg = bool(seed & 1)
        assert np.array_equal(fhe.cmux_chain(n, k, [g, g], ct1, h, negacyclic=neg),
                              orc.cmux_chain(n, k, np.concatenate([tggsw, tggsw]), ct1, h, neg))
        hh = int(rng.integers(0, n))
        got = fhe.sample_extract(n, k, ct1, hh)
        for i in range(batch):
            want = np.empty(k * n + 1, dtype=np.uint64)
            orc.lib().orc_tglwe_sample_extraction(n, k, orc.ptr(np.ascontiguousarray(ct1[i])), hh, orc.ptr(want))
            assert np.array_equal(got[i], want)


@pytest.mark.parametrize("seed", range(max(10, SOAK)))
def test_random_key_switch_and_bootstrap_shapes(fhe, orc, seed):
    rng = np.random.default_rng(200 + seed)
    kn_in, kn_out = int(rng.integers(1, 200)), int(rng.integers(1, 200))
    l = int([64, 64, 64, 1, 7, 33][rng.integers(0, 6)])
    if seed % 3 == 0:
        kn_out = 32 * int(rng.integers(1, 5))  # the split body column of the tensor-core paths
    batch = int(rng.integers(1, 300))
    ksk = orc.uniform(seed, kn_in * l * (kn_out + 1))
    ct = orc.uniform(seed + 1, (batch, kn_in + 1))
    K = fhe.Ksk(kn_in, kn_out, l, ksk)
    assert np.array_equal(K.key_switch(ct).reshape(-1), orc.key_switch(kn_in, kn_out, l, ksk, ct.reshape(-1), threads=4))
    # bootstrapping as executed on a small ring with this batch
    n, k = 1 << int(rng.integers(1, 8)), int(rng.integers(1, 3))
    kn = n * k
    ksk2 = orc.uniform(seed + 2, kn * 64 * (kn + 1))
    table = orc.uniform(seed + 3, (k + 1) * n)
    c_kn = int(rng.integers(1, 50))
    cts = orc.uniform(seed + 4, (batch, c_kn + 1))
    K2 = fhe.Ksk(kn, kn, 64, ksk2)
    assert np.array_equal(fhe.bootstrap(n, k, K2, table, cts, c_kn).reshape(-1),
                          orc.bootstrapping(n, k, ksk2, table, cts.reshape(-1), c_kn, threads=4))


@pytest.mark.parametrize("seed", range(max(8, SOAK)))
def test_random_bfv_and_gfhe_shapes(fhe, orc, seed):
    rng = np.random.default_rng(300 + seed)
    q = [Q17, 12289, Q30, 257][rng.integers(0, 4)]
    n = int(rng.integers(1, 70))  # BFV works on any degree (linear convolutions, no transform)
    t = int(rng.integers(2, 40))
    p = int([q * q, 2**10, 1, q][rng.integers(0, 4)])
    pq = p * q
    if pq >= 2**63:
        pq, p = q * 4, 4
    batch = int(rng.integers(1, 40))
    a, b = orc.uniform(seed, (batch, 2 * n), q), orc.uniform(seed + 1, (batch, 2 * n), q)
    rlk = orc.uniform(seed + 2, 2 * n, pq)
    assert np.array_equal(fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b).reshape(-1),
                          orc.bfv_mul(q, n, t, pq, rlk, a.reshape(-1), b.reshape(-1)))
    # GLWE<Rq> key switch on a random small shape
    q2 = [Q17, 12289, Q62][rng.integers(0, 3)]
    n2, k2 = 1 << int(rng.integers(1, 9)), int(rng.integers(1, 6))
    beta = int(rng.integers(2, 6))
    l2 = int(rng.integers(1, 12))
    while beta ** l2 >= 2**32 or q2 // beta ** l2 == 0:
        l2 -= 1
    glwe = (k2 + 1) * n2
    ksk = orc.uniform(seed + 3, k2 * l2 * glwe, q2)
    cts = orc.uniform(seed + 4, (batch, glwe), q2)
    K = fhe.RqGlev(fhe.NttPlan(q2, n2), k2, k2 * l2, ksk)
    assert np.array_equal(K.key_switch(beta, l2, cts), orc.glwe_rq_key_switch(q2, n2, k2, beta, l2, ksk, cts))
