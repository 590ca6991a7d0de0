"""GPU parity of the scheme-level entry points added in round 2: BFV key and relinearisation-key
generation, BFV::mul_const, compute_lookup_table, and the CKKS Rq paths -- bit-exact against the oracle's restatement
(counter-based sampler), plus the reference's own functional properties (decrypt(op(encrypt(m))) == op(m)) on keys and
ciphertexts that were generated on the device."""
import numpy as np
import pytest

from primes import Q17

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fhe():
    import torch

    assert torch.cuda.is_available()
    import fhe_study_b200 as f

    f.set_device(0)
    return f


def _negacyclic(a, b, n, t):
    e = np.zeros(n, dtype=np.int64)
    for i in range(n):
        for j in range(n):
            v = int(a[i]) * int(b[j])
            if i + j >= n:
                e[i + j - n] -= v
            else:
                e[i + j] += v
    return (e % t).astype(np.uint64)


@pytest.mark.parametrize("n", [16, 128, 512])
def test_bfv_keygen_matches_oracle(fhe, orc, n):
    # bfv/src/lib.rs:120-140 (parameter sets of the reference's tests: n = 16, 128, 512)
    plan = fhe.NttPlan(Q17, n)
    for seed in (1, 99):
        sk, pk = fhe.bfv_keygen(plan, 3.2, seed)
        wsk, wpk = orc.bfv_keygen_ctr(seed, Q17, n, 3.2)
        assert (sk == wsk).all() and (pk == wpk).all()
        assert set(np.unique(sk).tolist()) <= {0, 1}


@pytest.mark.parametrize("n,p", [(16, Q17 * Q17), (64, Q17 * Q17), (16, Q17), (1024, Q17)])
def test_bfv_rlk_generate_matches_oracle(fhe, orc, n, p):
    # bfv/src/lib.rs:202-225 with p = q^2 (the reference's tests) and other multipliers
    sk, _ = orc.bfv_keygen_ctr(5, Q17, n, 3.2)
    for seed in (2, 77):
        got = fhe.bfv_rlk_generate(Q17, n, p, sk, 3.2, seed)
        assert (got == orc.bfv_rlk_key_ctr(seed, Q17, n, p, 3.2, sk)).all()


def test_bfv_pipeline_with_device_generated_keys(fhe, orc):
    # the reference's test_mul_relin (bfv/src/lib.rs:557-601) and test_constant_add_mul (:342-377) with EVERY key and
    # ciphertext produced on the device: keygen -> rlk -> encrypt -> mul+relin / mul_const -> decrypt
    q, n, t = Q17, 16, 2
    p = q * q
    pq = p * q
    plan = fhe.NttPlan(q, n)
    sk, pk = fhe.bfv_keygen(plan, 3.2, 11)
    rlk = fhe.bfv_rlk_generate(q, n, p, sk, 3.2, 12)
    rng = np.random.default_rng(4)
    batch = 200
    m1 = rng.integers(0, t, size=(batch, n), dtype=np.uint64)
    m2 = rng.integers(0, t, size=(batch, n), dtype=np.uint64)
    c1 = fhe.bfv_encrypt(plan, t, pk, m1, 3.2, 21)
    c2 = fhe.bfv_encrypt(plan, t, pk, m2, 3.2, 22)
    c3 = fhe.bfv_mul_relin(q, n, t, pq, rlk, c1, c2)
    assert (c3.reshape(-1) == orc.bfv_mul(q, n, t, pq, rlk, c1.reshape(-1), c2.reshape(-1))).all()
    d3 = fhe.bfv_decrypt(plan, t, sk, c3)
    want = np.stack([_negacyclic(m1[i], m2[i], n, t) for i in range(batch)])
    assert (d3 == want).all()
    # mul_const at t = 8 (lib.rs:342-377)
    t8 = 8
    m1 = rng.integers(0, t8, size=(batch, n), dtype=np.uint64)
    mc = rng.integers(0, t8, size=(batch, n), dtype=np.uint64)
    c1 = fhe.bfv_encrypt(plan, t8, pk, m1, 3.2, 23)
    cm = fhe.bfv_mul_const(q, n, t8, pq, rlk, c1, mc)
    assert (cm == orc.bfv_mul_const(q, n, t8, pq, rlk, c1, mc)).all()
    dm = fhe.bfv_decrypt(plan, t8, sk, cm)
    good = sum(int((dm[i] == _negacyclic(m1[i], mc[i], n, t8)).all()) for i in range(batch))
    assert good >= batch * 0.95  # the reference's own test tolerates noise overflow only through its parameter choice


@pytest.mark.parametrize("n,k,t", [(1024, 1, 128), (64, 4, 16), (64, 16, 64), (8, 1, 8), (16, 2, 5)])
def test_compute_lookup_table(fhe, orc, n, k, t):
    # tfhe/src/tlwe.rs:196-214 at the reference's parameter sets (and a t that does not divide n)
    assert (fhe.compute_lookup_table(n, k, t) == orc.lookup_table(n, k, t)).all()


@pytest.mark.parametrize("n", [16, 32])
def test_ckks_rq_paths(fhe, orc, n):
    # ckks/src/lib.rs:46-119 at the reference's rings (q = 65537, n = 32 and 16)
    q = Q17
    plan = fhe.NttPlan(q, n)
    sk, pk = fhe.ckks_keygen(plan, 3.2, 3)
    wsk, wpk = orc.ckks_keygen_ctr(3, q, n, 3.2)
    assert (sk == wsk).all() and (pk == wpk).all()
    rng = np.random.default_rng(n)
    batch = 64
    m0 = rng.integers(-3000, 3000, size=(batch, n), dtype=np.int64)
    m1 = rng.integers(-3000, 3000, size=(batch, n), dtype=np.int64)
    c0, c1 = fhe.ckks_encrypt(plan, pk, m0, 3.2, 5), fhe.ckks_encrypt(plan, pk, m1, 3.2, 6)
    assert (c0 == orc.ckks_encrypt_ctr(5, q, n, 3.2, pk, m0)).all()
    d0 = fhe.ckks_decrypt(plan, sk, c0)
    assert (d0 == orc.ckks_decrypt(q, n, sk, c0)).all()
    assert np.abs(d0 - m0).max() < 200  # noise only
    add = fhe.ckks_add(q, n, c0, c1)
    assert (add == orc.ckks_addsub(q, n, c0, c1, False)).all()
    assert np.abs(fhe.ckks_decrypt(plan, sk, add) - (m0 + m1)).max() < 400
    sub = fhe.ckks_add(q, n, c0, c1, sub=True)
    assert (sub == orc.ckks_addsub(q, n, c0, c1, True)).all()  # as written: second components are ADDED (lib.rs:116-118)
    # extreme plaintext coefficients go through Zq::from_f64's signed reduction
    mx = np.array([[2**62, -(2**62), q, -q, q // 2, -(q // 2)] + [0] * (n - 6)], dtype=np.int64)
    assert (fhe.ckks_encrypt(plan, pk, mx, 3.2, 9) == orc.ckks_encrypt_ctr(9, q, n, 3.2, pk, mx)).all()


def test_bootstrap_requires_power_of_two_kn(fhe):
    # T64::mod_switch asserts q2.is_power_of_two() (torus.rs:58-66): k*n = 3*64 panics in the reference
    n, k = 64, 3
    table = np.zeros((k + 1) * n, dtype=np.uint64)
    ct = np.zeros((1, k * n + 1), dtype=np.uint64)
    with pytest.raises(fhe.FheError):
        fhe.blind_rotate(n, k, table, ct, k * n)


def test_container_files_feed_the_device_paths(fhe, orc, tmp_path):
    # a key-switching key, a TGGSW and packed Rq operands written to the flat container, read back and
    # handed to the C ABI without conversion give the same results as the originals
    kn, l = 64, 64
    ksk = orc.uniform(31, kn * l * (kn + 1))
    fhe.save(str(tmp_path / "ksk.fheb"), "ksk", ksk, kn + 1, q=0, n=kn, k=kn, l=l)
    info, back = fhe.load(str(tmp_path / "ksk.fheb"))
    cts = orc.uniform(32, (9, kn + 1))
    K1, K2 = fhe.Ksk(kn, kn, l, ksk), fhe.Ksk(int(info["k"]), int(info["n"]), int(info["l"]), back.reshape(-1))
    assert (K1.key_switch(cts) == K2.key_switch(cts)).all()
    n, k = 64, 2
    glwe = (k + 1) * n
    tggsw = orc.uniform(33, (k + 1) * 64 * glwe)
    fhe.save(str(tmp_path / "g.fheb"), "tggsw", tggsw, glwe, q=0, n=n, k=k, l=64)
    _, gb = fhe.load(str(tmp_path / "g.fheb"))
    ct = orc.uniform(34, (5, glwe))
    assert (fhe.Tggsw(n, k, gb.reshape(-1)).extprod(ct) == fhe.Tggsw(n, k, tggsw).extprod(ct)).all()
    q, nn = Q17, 1024
    a, b = orc.uniform(35, (6, nn), q), orc.uniform(36, (6, nn), q)
    fhe.save(str(tmp_path / "a.fheb"), "rq", a, nn, q=q, n=nn, encoding=fhe.ENC_PACKED, bits=17)
    ia, pa = fhe.load(str(tmp_path / "a.fheb"))
    assert (pa == a).all()
    plan = fhe.NttPlan(q, nn)
    got = plan.mul_packed(17, fhe.pack_bits(17, pa), fhe.pack_bits(17, b))
    assert (fhe.unpack_bits(17, got) == orc.rq_mul_batch(q, nn, a, b)).all()
