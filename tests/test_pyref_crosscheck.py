"""Cross-check of the C oracle against tests/pyref.py, an independent pure-Python big-integer transliteration of the same
reference functions (the oracle rows the reference holds no vector for: Tn product, external product / CMux, key switch,
bootstrapping as executed, BFV tensor + relinearisation; and the NTT against the schoolbook product).  CPU only."""
import numpy as np
import pytest

import pyref

Q17 = 65537


def _ints(a):
    return [int(x) for x in np.asarray(a).reshape(-1)]


@pytest.mark.parametrize("n", [4, 16, 64])
def test_ntt_and_rq_mul(orc, n):
    a, b = orc.uniform(1, n, Q17), orc.uniform(2, n, Q17)
    assert _ints(orc.ntt(Q17, n, a)) == pyref.ntt(Q17, n, _ints(a))
    got = _ints(orc.rq_mul_batch(Q17, n, a.reshape(1, n), b.reshape(1, n)))
    assert got == pyref.rq_mul(Q17, n, _ints(a), _ints(b)) == pyref.rq_mul_schoolbook(Q17, n, _ints(a), _ints(b))
    q62 = 0x3FFFFFFFFFFF0001
    a, b = orc.uniform(3, n, q62), orc.uniform(4, n, q62)
    assert _ints(orc.rq_mul_batch(q62, n, a.reshape(1, n), b.reshape(1, n))) == pyref.rq_mul_schoolbook(q62, n, _ints(a), _ints(b))


@pytest.mark.parametrize("n", [4, 32])
def test_tn_mul(orc, n):
    a, b = orc.uniform(5, n), orc.uniform(6, n)
    a[0], b[0] = np.uint64(2**64 - 1), np.uint64(2**64 - 1)
    assert _ints(orc.tn_mul(n, a.reshape(1, n), b.reshape(1, n))) == pyref.tn_mul(n, _ints(a), _ints(b))


@pytest.mark.parametrize("n,k", [(8, 1), (16, 2)])
def test_external_product_and_cmux(orc, n, k):
    glwe = (k + 1) * n
    flat = orc.uniform(7, (k + 1) * 64 * glwe)
    ct1, ct2 = orc.uniform(8, glwe), orc.uniform(9, glwe)
    rows = _ints(flat)
    tggsw = [[[rows[((i * 64 + j) * (k + 1) + c) * n:((i * 64 + j) * (k + 1) + c + 1) * n] for c in range(k + 1)]
              for j in range(64)] for i in range(k + 1)]
    polys = lambda w: [_ints(w)[c * n:(c + 1) * n] for c in range(k + 1)]
    want = [x for p in pyref.cmux(n, k, tggsw, polys(ct1), polys(ct2)) for x in p]
    assert _ints(orc.cmux(n, k, flat, ct1, ct2)) == want
    L = orc.lib()
    out = np.empty(glwe, dtype=np.uint64)
    L.orc_extprod_batch(n, k, orc.ptr(flat), orc.ptr(ct1), orc.ptr(out), 1, 1)
    assert _ints(out) == [x for p in pyref.external_product(n, k, tggsw, polys(ct1)) for x in p]


def test_key_switch_and_bootstrapping_as_executed(orc):
    n, k, l = 8, 1, 64
    kn = n * k
    flat = orc.uniform(10, kn * l * (kn + 1))
    rows = _ints(flat)
    ksk = [[rows[(i * l + j) * (kn + 1):(i * l + j + 1) * (kn + 1)] for j in range(l)] for i in range(kn)]
    ct = orc.uniform(11, kn + 1)
    assert _ints(orc.key_switch(kn, kn, l, flat, ct)) == pyref.key_switch(kn, kn, l, ksk, _ints(ct))
    table = orc.uniform(12, (k + 1) * n)
    tp = [_ints(table)[c * n:(c + 1) * n] for c in range(k + 1)]
    for seed in (13, 14, 15):
        c = orc.uniform(seed, kn + 1)
        assert _ints(orc.bootstrapping(n, k, flat, table, c, kn)) == pyref.bootstrapping_as_executed(n, k, ksk, tp, _ints(c), kn)


@pytest.mark.parametrize("n,t", [(4, 2), (16, 2), (16, 8)])
def test_bfv_mul_relin(orc, n, t):
    q = Q17
    pq = q * q * q
    for seed in range(5):
        a, b = orc.uniform(20 + seed, 2 * n, q), orc.uniform(30 + seed, 2 * n, q)
        rlk = orc.uniform(40 + seed, 2 * n, pq)
        if seed == 0:  # extremes: every product term at its maximum, relinearisation sums wrap i64
            a[:], b[:], rlk[:] = q - 1, q - 1, pq - 1
        A, B, K = _ints(a), _ints(b), _ints(rlk)
        want0, want1 = pyref.bfv_mul(q, n, t, pq, (K[:n], K[n:]), (A[:n], A[n:]), (B[:n], B[n:]))
        assert _ints(orc.bfv_mul(q, n, t, pq, rlk, a, b)) == want0 + want1
