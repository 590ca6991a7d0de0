"""GPU parity (bit-exact) of the torus path against the oracle through the C ABI: Tn arithmetic
(arith/src/ring_torus.rs), TGGSW external product and CMux (tfhe/src/tggsw.rs:39-62) at the reference's
own parameter sets (n=64,k=4,t=16 -- tggsw.rs:159-167; n=1024,k=1 -- tlwe.rs:467-475), plus the
reference's functional property decrypt(tggsw(m1) (x) tglwe(m2)) == m1*m2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

M64 = 2**64


def neg64(v):
    return [(x + M64) % M64 for x in v]


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


def test_left_rotate_known_answers(fhe):
    # arith/src/ring_torus.rs:334-366
    f = np.array(neg64([2, 3, -4, -1]), dtype=np.uint64)
    assert list(fhe.tn_left_rotate(4, f, np.array([3], dtype=np.uint64))) == neg64([-1, -2, -3, 4])
    assert list(fhe.tn_left_rotate(4, f, np.array([1], dtype=np.uint64))) == neg64([3, -4, -1, -2])
    assert list(fhe.tn_left_rotate(4, f, np.array([4 + 1], dtype=np.uint64))) == neg64([3, -4, -1, -2])  # h % n


@pytest.mark.parametrize("n", [2, 4, 64, 128, 256, 512, 1024, 4096])
def test_tn_mul_matches_schoolbook(fhe, orc, n, monkeypatch):
    batch = 3 if n >= 1024 else 37  # ragged against the products-per-CTA grouping of the fused kernel
    a = orc.uniform(n + 1, (batch, n))
    b = orc.uniform(n + 2, (batch, n))
    a[0, :] = M64 - 1  # worst case for the exactness bound of the CRT lift
    b[0, :] = M64 - 1
    a[1, :] = 0
    b[2, :] = 1
    want = orc.tn_mul(n, a, b, threads=8)
    assert (fhe.tn_mul(n, a, b) == want).all()
    monkeypatch.setenv("FHE_TN_PATH", "unfused")  # building-block path (the only one outside 64 <= n <= 1024)
    assert (fhe.tn_mul(n, a, b) == want).all()
    monkeypatch.delenv("FHE_TN_PATH")


def test_tn_elementwise(fhe, orc):
    a, b = orc.uniform(1, 1000), orc.uniform(2, 1000)
    assert (fhe.tn_add(a, b) == a + b).all()
    assert (fhe.tn_sub(a, b) == a - b).all()
    assert (fhe.tn_neg(a) == (np.uint64(0) - a)).all()
    assert (fhe.tn_mul_u64(a, 12345678901234567) == a * np.uint64(12345678901234567)).all()
    # Tn::decompose: recomposition identity of arith/src/torus.rs:163-190
    d = fhe.tn_decompose(8, a[:64], 64)
    want = np.zeros((8, 64, 8), dtype=np.uint64)
    orcl = orc.lib()
    for p in range(8):
        orcl.orc_tn_decompose(8, orc.ptr(a[8 * p:]), 64, orc.ptr(want[p]))
    assert (d == want).all()
    for p, q in [(2**10, None), (2**32, None)]:
        got = fhe.tn_mod_switch(a, p)
        w = np.zeros_like(a)
        orcl.orc_tn_mod_switch(a.size, orc.ptr(a), p, orc.ptr(w))
        assert (got == w).all()
    got = fhe.tn_mul_div_round(a, 128, 2**64 - 1)
    w = np.zeros_like(a)
    orcl.orc_tn_mul_div_round(a.size, orc.ptr(a), 128, 2**64 - 1, orc.ptr(w))
    assert (got == w).all()


def _keys_and_cts(orc, n, k, t, seed, batch, uniform_mask):
    L = orc.lib()
    sk = np.empty(k * n, dtype=np.uint64)
    L.orc_tglwe_keygen(seed, n, k, orc.ptr(sk))
    m1 = orc.uniform(seed + 1, n, t)
    tggsw = np.empty((k + 1) * 64 * (k + 1) * n, dtype=np.uint64)
    L.orc_tggsw_encrypt_s(seed + 2, n, k, 3.2, orc.ptr(sk), orc.ptr(m1), uniform_mask, orc.ptr(tggsw))
    ms, cts = [], []
    for i in range(batch):
        m2 = orc.uniform(seed + 10 + i, n, t)
        p2 = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_encode(n, t, orc.ptr(m2), orc.ptr(p2))
        ct = np.empty((k + 1) * n, dtype=np.uint64)
        L.orc_tglwe_encrypt_s(seed + 100 + i, n, k, 3.2, orc.ptr(sk), orc.ptr(p2), uniform_mask, orc.ptr(ct))
        ms.append(m2)
        cts.append(ct)
    return sk, m1, tggsw, ms, np.stack(cts)


def test_external_product_reference_params(fhe, orc):
    # tfhe/src/tggsw.rs:157-196: n=64, k=4, t=16, the reference's sampling (mask from Xi_key)
    L = orc.lib()
    n, k, t = 64, 4, 16
    sk, m1, tggsw, ms, cts = _keys_and_cts(orc, n, k, t, 1000, 6, 0)
    g = fhe.Tggsw(n, k, tggsw)
    got = g.extprod(cts)
    want = orc.extprod(n, k, tggsw, cts.reshape(-1), fast=False).reshape(cts.shape)
    assert (got == want).all()
    for i in range(cts.shape[0]):  # decrypt(extprod) == m1 * m2 in Z_t[X]/(X^n+1)
        p = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decrypt(n, k, orc.ptr(sk), orc.ptr(np.ascontiguousarray(got[i])), orc.ptr(p))
        rec = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decode(n, t, orc.ptr(p), orc.ptr(rec))
        expect = np.empty(n, dtype=np.uint64)
        L.orc_r_mul_to_rq(n, orc.ptr(orc.i64(m1)), orc.ptr(orc.i64(ms[i])), t, orc.ptr(expect))
        assert (rec == expect).all()


FUSED_SHAPES = {(64, 4), (1024, 1), (512, 1), (256, 1), (64, 1), (128, 1), (256, 2), (512, 2)}


@pytest.mark.parametrize("n,k,batch", [(64, 4, 33), (1024, 1, 5), (256, 2, 7), (2048, 1, 2), (512, 1, 3), (128, 1, 9),
                                       (64, 1, 40), (512, 2, 2), (256, 1, 4)])
def test_external_product_and_cmux_dense_inputs(fhe, orc, n, k, batch, monkeypatch):
    # uniformly random TGGSW rows and accumulators: every digit and limb is exercised; both the fused kernel
    # (extprod_fused.cu) and the unfused building blocks must equal the oracle
    glwe = (k + 1) * n
    tggsw = orc.uniform(7 * n + k, (k + 1) * 64 * glwe)
    ct1 = orc.uniform(11 * n + k, (batch, glwe))
    ct2 = orc.uniform(13 * n + k, (batch, glwe))
    ct1[0, :] = M64 - 1
    g = fhe.Tggsw(n, k, tggsw)
    want_e = orc.extprod(n, k, tggsw, ct1.reshape(-1))
    want_c = orc.cmux(n, k, tggsw, ct1.reshape(-1), ct2.reshape(-1))
    for path in (["fused"] if (n, k) in FUSED_SHAPES else []) + ["unfused"]:
        monkeypatch.setenv("FHE_EXTPROD_PATH", path)
        assert (g.extprod(ct1).reshape(-1) == want_e).all(), path
        assert (g.cmux(ct1, ct2).reshape(-1) == want_c).all(), path
    monkeypatch.delenv("FHE_EXTPROD_PATH")


@pytest.mark.parametrize("ct,batch", [("512", 5), ("512", 2), ("512", 1), ("256", 3)])
def test_external_product_n1024_both_cta_shapes(fhe, orc, ct, batch, monkeypatch):
    # n = 1024, k = 1 has two fused kernels reading ONE key layout: 256 threads / one accumulator, and 512 threads / a
    # PAIR of accumulators sharing every key load (picked when batch > SM count).  FHE_XP_CT forces either; an odd
    # batch leaves the last pair half empty.
    n, k = 1024, 1
    glwe = (k + 1) * n
    tggsw = orc.uniform(91, (k + 1) * 64 * glwe)
    ct1 = orc.uniform(92, (batch, glwe))
    ct2 = orc.uniform(93, (batch, glwe))
    ct1[batch - 1, :] = M64 - 1
    g = fhe.Tggsw(n, k, tggsw)
    monkeypatch.setenv("FHE_XP_CT", ct)
    assert (g.extprod(ct1).reshape(-1) == orc.extprod(n, k, tggsw, ct1.reshape(-1))).all()
    assert (g.cmux(ct1, ct2).reshape(-1) == orc.cmux(n, k, tggsw, ct1.reshape(-1), ct2.reshape(-1))).all()
    steps = 2
    size = (k + 1) * 64 * glwe
    bsk = orc.uniform(94, steps * size)
    h = orc.uniform(95, (batch, steps)) % np.uint64(2 * n)
    handles = [fhe.Tggsw(n, k, bsk[j * size:(j + 1) * size]) for j in range(steps)]
    assert np.array_equal(fhe.cmux_chain(n, k, handles, ct1, h, negacyclic=True), orc.cmux_chain(n, k, bsk, ct1, h, negacyclic=True))
    monkeypatch.delenv("FHE_XP_CT")


def test_external_product_n1024_kernels_agree_at_batch(fhe, orc, monkeypatch):
    # a batch on both sides of the switch (SM count): the two kernels and the automatic choice give identical rows
    import torch

    n, k, batch = 1024, 1, 301
    glwe = (k + 1) * n
    tggsw = orc.uniform(96, (k + 1) * 64 * glwe)
    g = fhe.Tggsw(n, k, tggsw)
    gen = torch.Generator(device="cuda").manual_seed(97)
    ct1 = torch.randint(-(2**63), 2**63 - 1, (batch, glwe), dtype=torch.int64, device="cuda", generator=gen)
    ct2 = torch.randint(-(2**63), 2**63 - 1, (batch, glwe), dtype=torch.int64, device="cuda", generator=gen)
    fhe.use_torch_stream()
    auto = g.cmux(ct1, ct2).clone()
    for ct in ("256", "512"):
        monkeypatch.setenv("FHE_XP_CT", ct)
        assert torch.equal(g.cmux(ct1, ct2), auto), ct
    monkeypatch.delenv("FHE_XP_CT")
    rows = [0, 1, 150, 299, 300]
    sel = torch.tensor(rows, device="cuda")
    a1 = np.ascontiguousarray(ct1[sel].cpu().numpy().view(np.uint64))
    a2 = np.ascontiguousarray(ct2[sel].cpu().numpy().view(np.uint64))
    want = orc.cmux(n, k, tggsw, a1.reshape(-1), a2.reshape(-1))
    assert np.array_equal(auto[sel].cpu().numpy().view(np.uint64).reshape(-1), want)


def test_extprod_worst_case_bound(fhe, orc, monkeypatch):
    # all-ones rows and accumulators maximise the integer magnitude the two-prime lift has to carry
    for n, k in ((1024, 1), (64, 4), (2048, 1)):
        glwe = (k + 1) * n
        tggsw = np.full((k + 1) * 64 * glwe, M64 - 1, dtype=np.uint64)
        ct = np.full((1, glwe), M64 - 1, dtype=np.uint64)
        g = fhe.Tggsw(n, k, tggsw)
        want = orc.extprod(n, k, tggsw, ct.reshape(-1))
        for path in (["fused"] if (n, k) in FUSED_SHAPES else []) + ["unfused"]:
            monkeypatch.setenv("FHE_EXTPROD_PATH", path)
            assert (g.extprod(ct).reshape(-1) == want).all(), (n, k, path)
        monkeypatch.delenv("FHE_EXTPROD_PATH")
    # the exactness bound of the lift is enforced at load: (k+1)*64*n*2^32 < P/2
    with pytest.raises(fhe.FheError):
        fhe.Tggsw(8192, 3, np.zeros(4 * 64 * 4 * 8192, dtype=np.uint64))


def test_cmux_selects(fhe, orc):
    # no TGGSW cmux test exists in the reference: check the defining property with TGGSW(0)/TGGSW(1)
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
        res = fhe.Tggsw(n, k, tggsw).cmux(cts[0], cts[1])
        assert (res == orc.cmux(n, k, tggsw, cts[0], cts[1])).all()
        p = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decrypt(n, k, orc.ptr(sk), orc.ptr(res), orc.ptr(p))
        rec = np.empty(n, dtype=np.uint64)
        L.orc_tglwe_decode(n, t, orc.ptr(p), orc.ptr(rec))
        assert (rec == msgs[bit]).all()


def test_extprod_device_buffers_and_linearity(fhe, orc):
    # size-independent property at a larger batch: the external product is additive in the accumulator
    # when the two accumulators have disjoint bit supports (decompose(a+b) = decompose(a)+decompose(b))
    import torch

    n, k, batch = 1024, 1, 256
    glwe = (k + 1) * n
    tggsw = orc.uniform(99, (k + 1) * 64 * glwe)
    g = fhe.Tggsw(n, k, tggsw)
    x = orc.uniform(5, (batch, glwe))
    lo = x & np.uint64(0x00000000FFFFFFFF)
    hi = x & np.uint64(0xFFFFFFFF00000000)
    fhe.use_torch_stream()
    dx, dlo, dhi = (torch.from_numpy(v.view(np.int64)).cuda() for v in (x, lo, hi))
    full = g.extprod(dx)
    parts = g.extprod(dlo) + g.extprod(dhi)  # int64 add wraps mod 2^64
    torch.cuda.synchronize()
    assert torch.equal(full, parts)
    got = full[:3].cpu().numpy().view(np.uint64)
    assert (got.reshape(-1) == orc.extprod(n, k, tggsw, x[:3].reshape(-1))).all()


@pytest.mark.parametrize("n,k,batch", [(1024, 1, 1024), (64, 4, 8191)])
def test_cmux_at_baseline_batch_spot_checked(fhe, orc, n, k, batch):
    # BASELINE configs[3] sizes ("batched accumulators" sharing one TGGSW), ragged for the multi-accumulator CTAs:
    # rows spread over the batch against the oracle + row-permutation invariance of the whole batch
    glwe = (k + 1) * n
    tggsw = orc.uniform(601, (k + 1) * 64 * glwe)
    ct1, ct2 = orc.uniform(602, (batch, glwe)), orc.uniform(603, (batch, glwe))
    g = fhe.Tggsw(n, k, tggsw)
    got = g.cmux(ct1, ct2)
    rows = np.unique(np.concatenate([np.arange(5), batch - 1 - np.arange(5), orc.uniform(604, 6, batch).astype(np.int64)]))
    want = orc.cmux(n, k, tggsw, np.ascontiguousarray(ct1[rows]), np.ascontiguousarray(ct2[rows]))
    assert np.array_equal(got[rows], want)
    perm = np.argsort(orc.uniform(605, batch))
    assert np.array_equal(g.cmux(np.ascontiguousarray(ct1[perm]), np.ascontiguousarray(ct2[perm])), got[perm])
    # cmux(g, x, x) = x for any TGGSW (the external product of zero is zero)
    assert np.array_equal(g.cmux(ct1, ct1), ct1)
