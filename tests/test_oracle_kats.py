"""Pins the CPU oracle (oracle/fhe_oracle.c) against every known-answer vector the reference's own
tests hold for the hot path, and re-runs the reference's functional property tests inside
the oracle.  CPU only."""
import numpy as np
import pytest

Q = 2**16 + 1


def neg64(v):
    return [(x + 2**64) % 2**64 for x in v]


# ---- arith/src/zq.rs:356-435 ---------------------------------------------------------------------
def test_zq_exp_kat(orc):
    L = orc.lib()
    assert L.orc_zq_exp(1021, 3, 3) == 27
    assert L.orc_zq_exp(1021, 1000, 3) == 949


def test_zq_neg_kat(orc):
    L = orc.lib()
    q = 1021
    a = L.orc_zq_from_f64(q, 101.0)
    b = L.orc_zq_from_f64(q, -1.0)
    assert b == q - 1
    assert L.orc_zq_neg(q, a) == L.orc_zq_mul(q, a, b)


def _zq_decompose(orc, q, v, beta, l):
    out = np.zeros(l, dtype=np.uint64)
    orc.lib().orc_zq_decompose(q, v, beta, l, orc.ptr(out))
    return [int(x) for x in out]


def _recompose(q, beta, l, d):
    x = 0
    for i in range(l):
        x += d[i] * q // beta ** (i + 1)
    return x % q


def test_zq_decompose_kats(orc):
    d = _zq_decompose(orc, 16, 9, 2, 4)
    assert d == [1, 0, 0, 1] and _recompose(16, 2, 4, d) == 9
    rng = np.random.default_rng(0)
    for x in rng.integers(0, 125, 200):
        d = _zq_decompose(orc, 125, int(x), 5, 3)
        assert len(d) == 3 and _recompose(125, 5, 3, d) == int(x)
    # saturating ("approx") branch, zq.rs:411-435
    assert _recompose(17, 2, 4, _zq_decompose(orc, 17, 16, 2, 4)) == 15
    assert _recompose(126, 5, 3, _zq_decompose(orc, 126, 125, 5, 3)) == 124
    assert _recompose(Q, 2, 16, _zq_decompose(orc, Q, Q - 1, 2, 16)) == 2**16 - 1


# ---- arith/src/torus.rs:163-190 ------------------------------------------------------------------
def test_t64_decompose_recompose(orc):
    xs = np.array([12345, 0, 2**64 - 2] + list(orc.uniform(7, 100)), dtype=np.uint64)
    out = np.zeros(64 * xs.size, dtype=np.uint64)
    orc.lib().orc_tn_decompose(xs.size, orc.ptr(xs), 64, orc.ptr(out))
    out = out.reshape(64, xs.size)
    for c, x in enumerate(xs):
        acc = 0
        for j in range(64):
            acc = (acc << 1) | int(out[j, c])
        assert acc == int(x)


# ---- arith/src/ring_torus.rs:334-366 ---------------------------------------------------------------
def test_tn_left_rotate_kat(orc):
    f = np.array(neg64([2, 3, -4, -1]), dtype=np.uint64)
    assert list(orc.tn_left_rotate(4, f, 3)) == neg64([-1, -2, -3, 4])
    assert list(orc.tn_left_rotate(4, f, 1)) == neg64([3, -4, -1, -2])


# ---- arith/src/ring_nq.rs:627-729 ------------------------------------------------------------------
def test_rq_fold_and_addsub_kats(orc):
    f = orc.rq_from_vec_u64
    assert list(f(7, 3, [0, 1, 2, 3, 4, 5])) == [4, 4, 4]
    assert list(f(7, 3, [0, 1, 9, 3, 4, 5])) == [4, 4, 4]
    assert list(f(7, 4, [0, 1, 2, 3, 4, 5])) == [3, 3, 2, 3]
    assert list(f(7, 3, [0, 0, 0, 0, 4, 5])) == [0, 3, 2]
    assert list(f(7, 3, [5, 4, 5, 2, 1, 0])) == [3, 3, 5]
    a = f(7, 3, [0, 1, 2, 3, 4, 5])
    b = f(7, 3, [5, 4, 3, 2, 1, 0])
    assert list(b) == [3, 3, 3]
    assert list(orc.rq_addsub(7, 3, a, b, 0)) == [0, 0, 0]
    assert list(orc.rq_addsub(7, 3, a, b, 1)) == [1, 1, 1]


def test_rq_mul_kats(orc):
    assert list(orc.rq_mul(Q, 4, [1, 2, 3, 4], [1, 2, 3, 4])) == [65513, 65517, 65531, 20]
    assert list(orc.rq_mul(Q, 4, [0, 0, 0, 2], [0, 0, 0, 2])) == [0, 0, 65533, 0]


def test_rq_decompose_kat(orc):
    a = np.array([7, 14, 3, 6], dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    orc.lib().orc_rq_decompose(16, 4, orc.ptr(a), 4, 2, orc.ptr(out))
    assert list(out[:4]) == [1, 3, 0, 1]
    assert list(out[4:]) == [3, 2, 3, 2]


# ---- NTT plan values derived in SURVEY 8c (cross-checked by the mul KATs above) -----------------------
def test_ntt_plan_values(orc):
    roots, roots_inv, n_inv = orc.ntt_tables(Q, 4)
    assert list(roots) == [1, 65281, 4096, 16]
    assert list(roots_inv) == [1, 256, 65521, 61441]
    assert n_inv == 49153
    assert list(orc.ntt(Q, 4, [1, 2, 3, 4])) == [7489, 56514, 17185, 49890]
    assert list(orc.ntt(Q, 4, [0, 0, 0, 2])) == [32, 65505, 8192, 57345]
    L = orc.lib()
    for n, psi in [(4, 4096), (512, 19139), (1024, 61869), (4096, 6561), (16384, 9)]:
        assert L.orc_primitive_root_of_unity(Q, 2 * n) == psi


# ---- arith/src/ntt.rs:194-234 (round trip) --------------------------------------------------------------
@pytest.mark.parametrize("n", [4, 512])
def test_ntt_roundtrip(orc, n):
    a = orc.uniform(11, (50, n), Q)
    assert np.array_equal(orc.ntt(Q, n, orc.ntt(Q, n, a), inverse=True), a)


def test_rq_mul_equals_schoolbook(orc):
    """NTT product == negacyclic schoolbook mod q (what the Sage KATs assert, on random inputs)."""
    n = 64
    a = orc.uniform(1, n, Q)
    b = orc.uniform(2, n, Q)
    c = orc.rq_mul(Q, n, a, b)
    ref = [0] * n
    for i in range(n):
        for j in range(n):
            s = int(a[i]) * int(b[j])
            if i + j >= n:
                ref[i + j - n] -= s
            else:
                ref[i + j] += s
    assert [x % Q for x in ref] == [int(x) for x in c]


# ---- arith/src/ring_n.rs:454-483 --------------------------------------------------------------------------
def test_r_linear_mul_kats(orc):
    L = orc.lib()
    for a, expect in [([Q - 1, Q - 1], [0, 8589934592]), ([1, Q - 1], [-4294967295, 131072])]:
        a = orc.i64(a)
        out = np.zeros(3, dtype=np.int64)
        L.orc_r_naive_mul(2, orc.ptr(a), orc.ptr(a), orc.ptr(out))
        ln = L.orc_r_fold(2, orc.ptr(out), 3)
        assert list(out[:ln]) == expect


def test_tn_mul_fast_equals_reference_form(orc):
    n = 128
    a, b = orc.uniform(3, n), orc.uniform(4, n)
    c1 = orc.tn_mul(n, a, b)
    c2 = np.empty(n, dtype=np.uint64)
    orc.lib().orc_tn_mul_fast(n, orc.ptr(a), orc.ptr(b), orc.ptr(c2))
    assert np.array_equal(c1, c2)
    # and against python big ints
    ref = [0] * n
    for i in range(n):
        for j in range(n):
            s = int(a[i]) * int(b[j])
            if i + j >= n:
                ref[i + j - n] -= s
            else:
                ref[i + j] += s
    assert [x % 2**64 for x in ref] == [int(x) for x in c1]


# ---- tfhe/src/tglwe.rs:337-368 (sample extraction property) -------------------------------------------------
def test_sample_extraction_property(orc):
    L = orc.lib()
    n, k, t = 64, 4, 128
    sk = np.empty(k * n, dtype=np.uint64)
    L.orc_tglwe_keygen(5, n, k, orc.ptr(sk))
    m = orc.uniform(6, n, t)
    p = np.empty(n, dtype=np.uint64)
    L.orc_tglwe_encode(n, t, orc.ptr(m), orc.ptr(p))
    ct = np.empty((k + 1) * n, dtype=np.uint64)
    L.orc_tglwe_encrypt_s(7, n, k, 3.2, orc.ptr(sk), orc.ptr(p), 1, orc.ptr(ct))
    delta = (2**64 - 1) // t
    for h in range(n):
        ext = np.empty(k * n + 1, dtype=np.uint64)
        L.orc_tglwe_sample_extraction(n, k, orc.ptr(ct), h, orc.ptr(ext))
        ph = L.orc_tlwe_decrypt(k * n, orc.ptr(sk), orc.ptr(ext))
        mh = L.orc_t64_mul_div_round(ph, t, 2**64 - 1) % t
        assert mh == int(m[h]), h


# ---- tfhe/src/tggsw.rs:157-196 (external product functional test, n=64 k=4 t=16) -------------------------------
def test_external_product_functional(orc):
    L = orc.lib()
    n, k, t = 64, 4, 16
    for trial in range(3):
        sk = np.empty(k * n, dtype=np.uint64)
        L.orc_tglwe_keygen(100 + trial, n, k, orc.ptr(sk))
        m1 = orc.uniform(200 + trial, n, t)
        m2 = orc.uniform(300 + trial, n, t)
        p2 = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_encode(n, t, orc.ptr(m2), orc.ptr(p2))
        tggsw = np.empty((k + 1) * 64 * (k + 1) * n, dtype=np.uint64)
        L.orc_tggsw_encrypt_s(400 + trial, n, k, 3.2, orc.ptr(sk), orc.ptr(m1), 0, orc.ptr(tggsw))
        ct = np.empty((k + 1) * n, dtype=np.uint64)
        L.orc_tglwe_encrypt_s(500 + trial, n, k, 3.2, orc.ptr(sk), orc.ptr(p2), 0, orc.ptr(ct))
        res = orc.extprod(n, k, tggsw, ct, fast=False)
        assert np.array_equal(res, orc.extprod(n, k, tggsw, ct, fast=True))
        p = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decrypt(n, k, orc.ptr(sk), orc.ptr(res), orc.ptr(p))
        rec = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decode(n, t, orc.ptr(p), orc.ptr(rec))
        expect = np.empty(n, dtype=np.uint64)
        L.orc_r_mul_to_rq(n, orc.ptr(orc.i64(m1)), orc.ptr(orc.i64(m2)), t, orc.ptr(expect))
        assert np.array_equal(rec, expect)


def test_cmux_selects(orc):
    """No TGGSW cmux test exists in the reference; check the defining property with TGGSW(0)/TGGSW(1)."""
    L = orc.lib()
    n, k, t = 64, 2, 16
    sk = np.empty(k * n, dtype=np.uint64)
    L.orc_tglwe_keygen(1, n, k, orc.ptr(sk))
    cts, msgs = [], []
    for s in (10, 20):
        m = orc.uniform(s, n, t)
        p = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_encode(n, t, orc.ptr(m), orc.ptr(p))
        ct = np.empty((k + 1) * n, dtype=np.uint64)
        L.orc_tglwe_encrypt_s(s + 1, n, k, 3.2, orc.ptr(sk), orc.ptr(p), 0, orc.ptr(ct))
        cts.append(ct)
        msgs.append(m)
    for bit in (0, 1):
        mb = np.zeros(n, dtype=np.uint64)
        mb[0] = bit
        tggsw = np.empty((k + 1) * 64 * (k + 1) * n, dtype=np.uint64)
        L.orc_tggsw_encrypt_s(30 + bit, n, k, 3.2, orc.ptr(sk), orc.ptr(mb), 0, orc.ptr(tggsw))
        res = orc.cmux(n, k, tggsw, cts[0], cts[1])
        p = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decrypt(n, k, orc.ptr(sk), orc.ptr(res), orc.ptr(p))
        rec = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decode(n, t, orc.ptr(p), orc.ptr(rec))
        assert np.array_equal(rec, msgs[bit])


# ---- tfhe/src/tlwe.rs:423-463 (key switch functional, k=16 n=1 t=128) ----------------------------------------
def test_key_switch_functional(orc):
    L = orc.lib()
    kn, t, l = 16, 128, 64
    sk = np.empty(kn, dtype=np.uint64)
    sk2 = np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(1, kn, orc.ptr(sk))
    L.orc_tlwe_keygen(2, kn, orc.ptr(sk2))
    ksk = np.empty(kn * l * (kn + 1), dtype=np.uint64)
    L.orc_tlwe_new_ksk(3, kn, kn, l, 3.2, orc.ptr(sk), orc.ptr(sk2), 0, orc.ptr(ksk))
    delta = (2**64 - 1) // t
    for m in (0, 1, 77, 127):
        ct = np.empty(kn + 1, dtype=np.uint64)
        L.orc_tlwe_encrypt_s(10 + m, kn, 3.2, orc.ptr(sk), (m * delta) % 2**64, 1, orc.ptr(ct))
        out = orc.key_switch(kn, kn, l, ksk, ct)
        p = L.orc_tlwe_decrypt(kn, orc.ptr(sk2), orc.ptr(out))
        assert L.orc_t64_mul_div_round(p, t, 2**64 - 1) % t == m


# ---- tfhe/src/tlwe.rs:465-504 (bootstrapping as executed, n=1024 k=1 t=128) -------------------------------------
def test_bootstrapping_as_executed_functional(orc):
    L = orc.lib()
    n, k, t = 1024, 1, 128
    kn = n * k
    table = orc.lookup_table(n, k, t)
    assert not table[:n].any()
    delta = (2**64 - 1) // t
    assert int(table[n + 8]) == delta and int(table[n + 1023]) == 127 * delta
    sk = np.empty(kn, dtype=np.uint64)
    sk2 = np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(1, kn, orc.ptr(sk))
    L.orc_tlwe_keygen(2, kn, orc.ptr(sk2))
    ksk = np.empty(kn * 64 * (kn + 1), dtype=np.uint64)
    L.orc_tlwe_new_ksk(3, kn, kn, 64, 3.2, orc.ptr(sk), orc.ptr(sk2), 0, orc.ptr(ksk))
    for m in (0, 5, 100, 127):
        ct = np.empty(kn + 1, dtype=np.uint64)
        # the reference's own sampling: mask from Xi_key (glwe.rs:146-149), so a.s is tiny
        L.orc_tlwe_encrypt_s(10 + m, kn, 3.2, orc.ptr(sk), (m * delta) % 2**64, 0, orc.ptr(ct))
        out = orc.bootstrapping(n, k, ksk, table, ct, kn)
        p = L.orc_tlwe_decrypt(kn, orc.ptr(sk), orc.ptr(out))
        assert L.orc_t64_mul_div_round(p, t, 2**64 - 1) % t == m


# ---- bfv/src/lib.rs:557-601 (mul + relin functional, q=65537 n=16 t=2 p=q^2) ---------------------------------------
def test_bfv_mul_relin_functional(orc):
    L = orc.lib()
    q, n, t = Q, 16, 2
    p = q * q
    pq = p * q
    ok = 0
    trials = 50
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
        c3 = orc.bfv_mul(q, n, t, pq, rlk, c1, c2)
        m3 = np.empty(n, dtype=np.uint64)
        L.orc_bfv_decrypt(q, n, t, orc.ptr(sk), orc.ptr(c3), orc.ptr(m3))
        expect = np.empty(n, dtype=np.uint64)
        L.orc_r_mul_to_rq(n, orc.ptr(orc.i64(m1)), orc.ptr(orc.i64(m2)), t, orc.ptr(expect))
        ok += int(np.array_equal(m3, expect))
    assert ok == trials


# ---- CMux chain / blind rotation with one TGGSW per mask element (extension of tlwe.rs:138-147) ----------
def pbs_fixture(orc, n=64, k=1, m_lwe=6, t=4, seed=77, sigma=3.2):
    """Binary LWE key s, GLWE key z, bsk[j] = TGGSW_z(s_j), table = identity LUT, inputs encrypting 0..t-1."""
    L = orc.lib()
    s = (orc.uniform(seed, m_lwe) & np.uint64(1)).astype(np.uint64)
    z = (orc.uniform(seed + 1, k * n) & np.uint64(1)).astype(np.uint64)
    glwe = (k + 1) * n
    bsk = np.zeros((m_lwe, (k + 1) * 64 * glwe), dtype=np.uint64)
    for j in range(m_lwe):
        msg = np.zeros(n, dtype=np.uint64)
        msg[0] = s[j]
        L.orc_tggsw_encrypt_s(seed + 10 + j, n, k, sigma, orc.ptr(z), orc.ptr(msg), 1, orc.ptr(bsk[j]))
    table = orc.lookup_table(n, k, t)
    delta_n = n // t
    cts = np.zeros((t, m_lwe + 1), dtype=np.uint64)
    for m in range(t):
        phase = (m * delta_n + delta_n // 2) * (2**64 // (2 * n))  # centre of the m-th window of the LUT
        L.orc_tlwe_encrypt_s(seed + 100 + m, m_lwe, sigma, orc.ptr(s), phase, 1, orc.ptr(cts[m]))
    return s, z, bsk, table, cts


def test_bootstrap_chain_recovers_message(orc):
    n, k, m_lwe, t = 64, 1, 6, 4
    s, z, bsk, table, cts = pbs_fixture(orc, n, k, m_lwe, t)
    out = orc.bootstrap_chain(n, k, m_lwe, bsk, None, table, cts, m_lwe, mode=1)  # TLWE of dimension k*n under z
    delta = (2**64 - 1) // t
    for m in range(t):
        phase = int(orc.lib().orc_tlwe_decrypt(k * n, orc.ptr(z), orc.ptr(out[m])))
        assert round(phase / delta) % t == m


def test_cmux_chain_is_composition_of_cmux_and_left_rotate(orc):
    n, k, steps = 16, 2, 3
    glwe = (k + 1) * n
    bsk = orc.uniform(5, steps * (k + 1) * 64 * glwe)
    acc = orc.uniform(6, glwe)
    for neg in (False, True):
        h = np.array([3, n + 5, 2 * n + 1], dtype=np.uint64)
        want = acc.copy()
        for j in range(steps):
            hj = int(h[j])
            rot = np.concatenate([orc.tn_left_rotate(n, want[c * n:(c + 1) * n], hj % n) for c in range(k + 1)])
            if neg and (hj % (2 * n)) >= n:
                rot = np.uint64(0) - rot
            want = orc.cmux(n, k, bsk[j * (k + 1) * 64 * glwe:(j + 1) * (k + 1) * 64 * glwe], want, rot, fast=False)
        assert np.array_equal(orc.cmux_chain(n, k, bsk, acc, h, negacyclic=neg).reshape(-1), want)


# ---- gfhe/src/glwe.rs:580-624 (test_key_switch) re-run inside the oracle -----------------------------------------
def glwe_rq_fixture(orc, q=Q, n=128, k=16, t=2, beta=2, l=16, seed=5, batch=2):
    L = orc.lib()
    glwe = (k + 1) * n
    sk, sk2 = np.empty(k * n, dtype=np.uint64), np.empty(k * n, dtype=np.uint64)
    L.orc_glwe_rq_keygen(seed, q, n, k, orc.ptr(sk))
    L.orc_glwe_rq_keygen(seed + 1, q, n, k, orc.ptr(sk2))
    ksk = np.empty(k * l * glwe, dtype=np.uint64)
    L.orc_glwe_rq_new_ksk(seed + 2, q, n, k, beta, l, 3.2, orc.ptr(sk), orc.ptr(sk2), orc.ptr(ksk))
    msgs = orc.uniform(seed + 3, (batch, n), t)
    cts = np.empty((batch, glwe), dtype=np.uint64)
    for i in range(batch):
        p = (msgs[i] * np.uint64(q // t)) % np.uint64(q)  # GLWE::encode, glwe.rs:183-189
        L.orc_glwe_rq_encrypt_s(seed + 10 + i, q, n, k, 3.2, orc.ptr(sk), orc.ptr(p), orc.ptr(cts[i]))
    return sk, sk2, ksk, msgs, cts


def glwe_rq_decode(orc, q, n, k, t, sk, ct):
    L = orc.lib()
    p = np.empty(n, dtype=np.uint64)
    L.orc_glwe_rq_decrypt(q, n, k, orc.ptr(sk), orc.ptr(np.ascontiguousarray(ct)), orc.ptr(p))
    r = np.empty(n, dtype=np.uint64)
    L.orc_rq_mul_div_round(q, n, orc.ptr(p), t, q, orc.ptr(r))  # GLWE::decode, glwe.rs:191-195
    return r % np.uint64(t)


def test_glwe_rq_key_switch_functional(orc):
    q, n, k, t, beta, l = Q, 128, 16, 2, 2, 16
    sk, sk2, ksk, msgs, cts = glwe_rq_fixture(orc, q, n, k, t, beta, l)
    out = orc.glwe_rq_key_switch(q, n, k, beta, l, ksk, cts)
    for i in range(len(msgs)):
        assert np.array_equal(glwe_rq_decode(orc, q, n, k, t, sk, cts[i]), msgs[i])       # sanity: decrypts under sk
        assert np.array_equal(glwe_rq_decode(orc, q, n, k, t, sk2, out[i]), msgs[i])      # and under sk2 after the switch


def test_counter_based_ksk_decrypts_to_gadget_multiples(orc):
    # orc_tlwe_new_ksk_ctr (the sampler the device generator reproduces): row i*l + lv-1 must decrypt, under new_sk, to
    # sk_i * (u64::MAX / 2^lv) up to the small non-negative error (tlev.rs:53-77)
    L = orc.lib()
    kn_in, kn_out, l = 6, 10, 64
    sk = orc.uniform(1, kn_in) & np.uint64(1)
    sk2 = orc.uniform(2, kn_out) & np.uint64(1)
    for uniform in (True, False):
        ksk = orc.tlwe_new_ksk_ctr(7, kn_in, kn_out, l, 3.2, sk, sk2, uniform).reshape(kn_in, l, kn_out + 1)
        errs = []
        for i in range(kn_in):
            for lv in range(1, l + 1):
                g = (2**64 - 1) // 2**lv if lv < 64 else 1
                p = int(L.orc_tlwe_decrypt(kn_out, orc.ptr(sk2), orc.ptr(np.ascontiguousarray(ksk[i, lv - 1]))))
                errs.append((p - int(sk[i]) * g) % 2**64)
        assert max(errs) < 40 and len(set(errs)) > 3  # errors are small, non-negative (T64::rand's cast) and not constant
