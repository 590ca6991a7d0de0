"""GPU parity (bit-exact, including the f64 scale/round epilogue) of BFV tensor / relinearize_204 / mul
(bfv/src/lib.rs:59-90,251-271) against the oracle at the reference's parameters (q=65537, n=16, t=2, p=q^2;
lib.rs:559-564), plus the reference's functional property decrypt(mul_relin(c1,c2)) == m1*m2, and the
coefficient-wise Rq operations with their known-answer vectors."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Q = 2**16 + 1


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


def _bfv_material(orc, q, n, t, p, trials):
    L = orc.lib()
    out = []
    for trial in range(trials):
        sk = np.empty(n, dtype=np.uint64)
        pk = np.empty(2 * n, dtype=np.uint64)
        L.orc_bfv_keygen(1000 + trial, q, n, orc.ptr(sk), orc.ptr(pk))
        rlk = np.empty(2 * n, dtype=np.uint64)
        L.orc_bfv_rlk_key(2000 + trial, q, n, p, orc.ptr(sk), orc.ptr(rlk))
        m1 = orc.uniform(3000 + trial, n, t)
        m2 = orc.uniform(4000 + trial, n, t)
        c1 = np.empty(2 * n, dtype=np.uint64)
        c2 = np.empty(2 * n, dtype=np.uint64)
        L.orc_bfv_encrypt(5000 + trial, q, n, t, orc.ptr(pk), orc.ptr(m1), orc.ptr(c1))
        L.orc_bfv_encrypt(6000 + trial, q, n, t, orc.ptr(pk), orc.ptr(m2), orc.ptr(c2))
        out.append((sk, rlk, m1, m2, c1, c2))
    return out


def test_mul_relin_reference_params_functional(fhe, orc):
    L = orc.lib()
    q, n, t = Q, 16, 2
    p = q * q
    pq = p * q
    for sk, rlk, m1, m2, c1, c2 in _bfv_material(orc, q, n, t, p, 25):
        c3 = fhe.bfv_mul_relin(q, n, t, pq, rlk, c1, c2)
        assert (c3 == orc.bfv_mul(q, n, t, pq, rlk, c1, c2)).all()
        m3 = np.empty(n, dtype=np.uint64)
        L.orc_bfv_decrypt(q, n, t, orc.ptr(sk), orc.ptr(c3), orc.ptr(m3))
        expect = np.empty(n, dtype=np.uint64)
        L.orc_r_mul_to_rq(n, orc.ptr(orc.i64(m1)), orc.ptr(orc.i64(m2)), t, orc.ptr(expect))
        assert (m3 == expect).all()


@pytest.mark.parametrize("n,t,batch", [(16, 2, 4096), (16, 8, 100), (64, 2, 33), (512, 32, 5), (2, 2, 7)])
def test_mul_relin_batched_uniform(fhe, orc, n, t, batch):
    # uniform ciphertexts and an RLK uniform below q^3: the relinearisation products wrap i64
    L = orc.lib()
    q = Q
    pq = q * q * q
    a = orc.uniform(n + 1, (batch, 2 * n), q)
    b = orc.uniform(n + 2, (batch, 2 * n), q)
    a[0, :] = q - 1
    b[0, :] = q - 1
    rlk = orc.uniform(n + 3, 2 * n, pq)
    want = orc.bfv_mul(q, n, t, pq, rlk, a.reshape(-1), b.reshape(-1), threads=8).reshape(batch, 2 * n)
    assert (fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b) == want).all()
    # tensor and relinearize_204 separately
    c012 = fhe.bfv_tensor(q, n, t, a, b)
    for i in range(min(batch, 8)):
        c0, c1, c2 = (np.empty(n, dtype=np.uint64) for _ in range(3))
        L.orc_bfv_tensor(q, n, t, orc.ptr(np.ascontiguousarray(a[i])), orc.ptr(np.ascontiguousarray(b[i])),
                         orc.ptr(c0), orc.ptr(c1), orc.ptr(c2))
        assert (c012[i] == np.concatenate([c0, c1, c2])).all()
    assert (fhe.bfv_relinearize(q, n, pq, rlk, c012) == want).all()


def test_mul_relin_other_modulus_and_device_buffers(fhe, orc):
    import torch

    q, n, t = 0x3FFC0001, 32, 4  # a 30-bit q: tensor products reach 2^65 and wrap before the f64 scaling
    p = 2**10
    pq = p * q
    a = orc.uniform(1, (50, 2 * n), q)
    b = orc.uniform(2, (50, 2 * n), q)
    rlk = orc.uniform(3, 2 * n, pq)
    want = orc.bfv_mul(q, n, t, pq, rlk, a.reshape(-1), b.reshape(-1)).reshape(50, 2 * n)
    fhe.use_torch_stream()
    da, db, dr = (torch.from_numpy(v.view(np.int64)).cuda() for v in (a, b, rlk))
    got = fhe.bfv_mul_relin(q, n, t, pq, dr, da, db)
    torch.cuda.synchronize()
    assert (got.cpu().numpy().view(np.uint64) == want).all()


def test_rq_coefficientwise_known_answers(fhe, orc):
    L = orc.lib()
    # fold / add / sub strings, arith/src/ring_nq.rs:627-665 (mod 7, n=3 and n=4)
    f = lambda q, n, v: list(fhe.rq_from_vec(q, n, np.array(v, dtype=np.uint64), len(v)).reshape(-1))
    assert f(7, 3, [0, 1, 2, 3, 4, 5]) == [4, 4, 4]
    assert f(7, 3, [0, 1, 9, 3, 4, 5]) == [4, 4, 4]
    assert f(7, 4, [1, 2, 3, 4, 5]) == list(orc.rq_from_vec_u64(7, 4, [1, 2, 3, 4, 5]))
    a = np.array([1, 2, 3, 4], dtype=np.uint64)
    b = np.array([1, 2, 3, 6], dtype=np.uint64)
    assert list(fhe.rq_add(7, a, b)) == [2, 4, 6, 3]
    assert list(fhe.rq_sub(7, a, b)) == [0, 0, 0, 5]
    assert list(fhe.rq_neg(7, a)) == [6, 5, 4, 3]
    # decompose KAT, arith/src/ring_nq.rs:707-729: q=16, n=4, beta=4, l=2
    d = fhe.rq_decompose(16, 4, np.array([7, 14, 3, 6], dtype=np.uint64), 4, 2)
    assert d.reshape(2, 4).tolist() == [[1, 3, 0, 1], [3, 2, 3, 2]]
    # random parity of every map against the oracle, incl. the saturating decompose branch
    q, n = Q, 64
    x = orc.uniform(4, (6, n), q)
    y = orc.uniform(5, (6, n), q)
    flat = x.reshape(-1)
    assert (fhe.rq_add(q, x, y).reshape(-1) == orc.rq_addsub(q, n, flat, y.reshape(-1), 0)).all()
    assert (fhe.rq_sub(q, x, y).reshape(-1) == orc.rq_addsub(q, n, flat, y.reshape(-1), 1)).all()
    w = np.empty_like(flat)
    L.orc_rq_mul_u64(q, flat.size, orc.ptr(flat), 2**63 + 12345, orc.ptr(w))
    assert (fhe.rq_mul_u64(q, x, 2**63 + 12345).reshape(-1) == w).all()
    L.orc_rq_remodule(flat.size, orc.ptr(flat), 257, orc.ptr(w))
    assert (fhe.rq_remodule(x, 257).reshape(-1) == w).all()
    L.orc_rq_mod_switch(q, flat.size, orc.ptr(flat), 257, orc.ptr(w))
    assert (fhe.rq_mod_switch(q, x, 257).reshape(-1) == w).all()
    L.orc_rq_mul_div_round(q, flat.size, orc.ptr(flat), 32, q, orc.ptr(w))
    assert (fhe.rq_mul_div_round(q, x, 32, q).reshape(-1) == w).all()
    for beta, l in [(2, 16), (2, 17), (4, 8), (5, 3), (2, 4)]:
        got = fhe.rq_decompose(q, n, x, beta, l)
        for p in range(6):
            want = np.empty(l * n, dtype=np.uint64)
            L.orc_rq_decompose(q, n, orc.ptr(np.ascontiguousarray(x[p])), beta, l, orc.ptr(want))
            assert (got[p].reshape(-1) == want).all(), (beta, l)


@pytest.mark.parametrize("q,p,n,t", [(2**16, 2**8, 16, 2), (3, 9, 8, 2), (2**61 - 1, 4, 16, 5), (2**32, 2**16, 4, 3),
                                      (2**63 - 25, 1, 8, 2)])
def test_mul_relin_modulus_edge_cases(fhe, orc, q, p, n, t):
    # power-of-two and near-2^63 moduli: Zq::from_f64's reduction of negative / huge rounded values (zq.rs:32-40)
    pq = p * q
    a = orc.uniform(11, (9, 2 * n), q)
    b = orc.uniform(12, (9, 2 * n), q)
    a[0, :] = q - 1
    b[0, :] = q - 1
    rlk = orc.uniform(13, 2 * n, pq)
    want = orc.bfv_mul(q, n, t, pq, rlk, a.reshape(-1), b.reshape(-1)).reshape(9, 2 * n)
    assert (fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b) == want).all()


@pytest.mark.parametrize("q,n,t,batch", [(Q, 512, 32, 9), (Q, 128, 32, 40), (Q, 16, 2, 4096), (0x3FFC0001, 64, 5, 7)])
def test_bfv_decrypt_batched(fhe, orc, q, n, t, batch):
    # BFV::decrypt (bfv/src/lib.rs:164-178) at the reference's parameter sets (lib.rs:283-290: n=512, t=32; :311-318:
    # n=128; :559-564: n=16, t=2): real encryptions decrypt to their messages, random ciphertexts match the oracle
    L = orc.lib()
    sk, pk = np.empty(n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64)
    L.orc_bfv_keygen(31, q, n, orc.ptr(sk), orc.ptr(pk))
    msgs = orc.uniform(32, (batch, n), t)
    cts = np.empty((batch, 2 * n), dtype=np.uint64)
    for i in range(min(batch, 50)):
        L.orc_bfv_encrypt(100 + i, q, n, t, orc.ptr(pk), orc.ptr(np.ascontiguousarray(msgs[i])), orc.ptr(cts[i]))
    cts[50:] = orc.uniform(33, (max(batch - 50, 0), 2 * n), q)
    plan = fhe.NttPlan(q, n)
    got = fhe.bfv_decrypt(plan, t, sk, cts)
    want = np.empty((batch, n), dtype=np.uint64)
    for i in range(batch):
        L.orc_bfv_decrypt(q, n, t, orc.ptr(sk), orc.ptr(np.ascontiguousarray(cts[i])), orc.ptr(want[i]))
    assert np.array_equal(got, want)
    k = min(batch, 50)
    assert np.array_equal(got[:k], msgs[:k])  # decrypt(encrypt(m)) == m (lib.rs:281-307)


def test_rq_mul_broadcast_operand(fhe, orc):
    # FHE_B_BROADCAST: one right operand for the whole batch (GLWE * R, gfhe/src/glwe.rs:263-280)
    q, n, batch = Q, 256, 11
    plan = fhe.NttPlan(q, n)
    a, b = orc.uniform(1, (batch, n), q), orc.uniform(2, n, q)
    want = orc.rq_mul_batch(q, n, a, np.tile(b, batch))
    assert np.array_equal(plan.mul(a, b, flags=fhe.B_BROADCAST), want)
    assert np.array_equal(plan.mul(a, plan.ntt(b), flags=fhe.B_BROADCAST | fhe.B_IS_EVALS), want)


@pytest.mark.parametrize("q,n,t,batch", [(Q, 512, 32, 6), (Q, 128, 32, 25), (Q, 16, 2, 300), (0x3FFC0001, 64, 5, 7)])
def test_bfv_encrypt_on_device_roundtrip(fhe, orc, q, n, t, batch):
    # BFV::encrypt (bfv/src/lib.rs:142-160) sampled on the device: bit-exact against the oracle's counter-based restatement,
    # and decrypt(encrypt(m)) == m as in test_encrypt_decrypt (lib.rs:281-307)
    L = orc.lib()
    sk, pk = np.empty(n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64)
    L.orc_bfv_keygen(41, q, n, orc.ptr(sk), orc.ptr(pk))
    msgs = orc.uniform(42, (batch, n), t)
    plan = fhe.NttPlan(q, n)
    ct = fhe.bfv_encrypt(plan, t, pk, msgs, sigma=3.2, seed=9)
    want = np.empty((batch, 2 * n), dtype=np.uint64)
    L.orc_bfv_encrypt_ctr(9, q, n, t, 3.2, orc.ptr(pk), orc.ptr(msgs), batch, orc.ptr(want))
    assert np.array_equal(ct, want)
    assert np.array_equal(fhe.bfv_decrypt(plan, t, sk, ct), msgs)


def test_bfv_pipeline_on_the_gpu(fhe, orc):
    # encrypt -> ciphertext multiply + relinearise -> decrypt, all batched on the device, at the reference's test_mul_relin
    # parameters (bfv/src/lib.rs:557-601): the decrypted product equals m1 * m2 in Z_t[X]/(X^n+1)
    L = orc.lib()
    q, n, t, batch = Q, 16, 2, 500
    p = q * q
    pq = p * q
    sk, pk, rlk = np.empty(n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64), np.empty(2 * n, dtype=np.uint64)
    L.orc_bfv_keygen(51, q, n, orc.ptr(sk), orc.ptr(pk))
    L.orc_bfv_rlk_key(52, q, n, p, orc.ptr(sk), orc.ptr(rlk))
    m1, m2 = orc.uniform(53, (batch, n), t), orc.uniform(54, (batch, n), t)
    plan = fhe.NttPlan(q, n)
    c1 = fhe.bfv_encrypt(plan, t, pk, m1, seed=1)
    c2 = fhe.bfv_encrypt(plan, t, pk, m2, seed=2)
    m3 = fhe.bfv_decrypt(plan, t, sk, fhe.bfv_mul_relin(q, n, t, pq, rlk, c1, c2))
    want = np.empty((batch, n), dtype=np.uint64)
    for i in range(batch):
        L.orc_r_mul_to_rq(n, orc.ptr(orc.i64(m1[i])), orc.ptr(orc.i64(m2[i])), t, orc.ptr(want[i]))
    # the reference's own test tolerates nothing: every coefficient of every product must match
    assert np.array_equal(m3, want)
