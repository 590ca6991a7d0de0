"""GPU parity of the gfhe layer over Rq through the C ABI: GLev<Rq> * Vec<Rq>
(gfhe/src/glev.rs:67-80) and GLWE<Rq>::key_switch (gfhe/src/glwe.rs:126-137) against the oracle, at the parameter
sets of the reference's tests (gfhe/src/glwe.rs:582-594, gfhe/src/glev.rs:92-104), plus the functional property of
test_key_switch (glwe.rs:580-624): the switched ciphertext decrypts under the second key."""
import numpy as np
import pytest

from primes import Q62
from test_oracle_kats import glwe_rq_decode, glwe_rq_fixture

pytestmark = pytest.mark.gpu
Q = 65537


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


def test_key_switch_reference_params_functional(fhe, orc):
    q, n, k, t, beta, l = Q, 128, 16, 2, 2, 16
    sk, sk2, ksk, msgs, cts = glwe_rq_fixture(orc, q, n, k, t, beta, l, batch=3)
    K = fhe.RqGlev(fhe.NttPlan(q, n), k, k * l, ksk)
    got = K.key_switch(beta, l, cts)
    assert np.array_equal(got, orc.glwe_rq_key_switch(q, n, k, beta, l, ksk, cts))
    for i in range(len(msgs)):
        assert np.array_equal(glwe_rq_decode(orc, q, n, k, t, sk2, got[i]), msgs[i])


@pytest.mark.parametrize("q,n,k,beta,l,batch", [(Q, 16, 2, 2, 16, 5), (Q, 8, 1, 4, 8, 3), (Q, 1024, 1, 2, 17, 2),
                                                  (Q, 64, 3, 3, 10, 4), (Q62, 32, 2, 2, 62, 2), (12289, 512, 2, 2, 10, 3),
                                                  (Q, 128, 16, 2, 16, 5), (Q, 1024, 1, 2, 16, 3), (Q, 64, 4, 2, 40, 9),
                                                  (Q, 256, 1, 2, 5, 6), (12289, 512, 1, 2, 14, 7), (Q, 64, 1, 2, 33, 5)])
def test_key_switch_random_inputs(fhe, orc, monkeypatch, q, n, k, beta, l, batch):
    glwe = (k + 1) * n
    ksk = orc.uniform(q % 1000 + n, k * l * glwe, q)
    cts = orc.uniform(n + k, (batch, glwe), q)
    cts[0, :n] = q - 1  # beta = 2 with 2^l <= q-1, or beta^l <= q-1: the saturating branch of Zq::decompose
    K = fhe.RqGlev(fhe.NttPlan(q, n), k, k * l, ksk)
    want = orc.glwe_rq_key_switch(q, n, k, beta, l, ksk, cts)
    assert np.array_equal(K.key_switch(beta, l, cts), want)          # fused kernel where (n, k, q, beta) allow it
    monkeypatch.setenv("FHE_GLWE_KS_PATH", "unfused")
    assert np.array_equal(K.key_switch(beta, l, cts), want)          # building-block path
    monkeypatch.delenv("FHE_GLWE_KS_PATH")


@pytest.mark.parametrize("q,n,k,l,batch", [(Q, 128, 16, 16, 2), (Q, 4, 1, 1, 7), (Q62, 64, 2, 5, 3)])
def test_glev_mul(fhe, orc, q, n, k, l, batch):
    glev = orc.uniform(3 + n, l * (k + 1) * n, q)
    v = orc.uniform(4 + n, (batch, l * n), q)
    G = fhe.RqGlev(fhe.NttPlan(q, n), k, l, glev)
    assert np.array_equal(G.mul(v), orc.glev_rq_mul(q, n, k, l, glev, v))


def test_gfhe_argument_errors(fhe, orc):
    plan = fhe.NttPlan(Q, 16)
    K = fhe.RqGlev(plan, 2, 2 * 4, orc.uniform(1, 2 * 4 * 3 * 16, Q))
    with pytest.raises(RuntimeError):
        K.key_switch(2, 5, orc.uniform(2, (1, 3 * 16), Q))  # handle holds k*4 rows, not k*5
    with pytest.raises(RuntimeError):
        K.key_switch(70000, 4, orc.uniform(2, (1, 3 * 16), Q))  # q / beta^i == 0: the reference divides by zero
