"""GPU parity (bit-exact) of the TLWE path against the oracle: key switch (tfhe/src/tlwe.rs:101-112), sample
extraction (tglwe.rs:89-115), mod switch, blind rotation as executed and as written, and bootstrapping
(tlwe.rs:150-161) at the reference's own parameters (n=1024, k=1, t=128, l=64; tlwe.rs:467-475), plus
the reference's functional property that bootstrapping preserves the message."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


@pytest.fixture(scope="module")
def p5(orc):
    """keys of the reference's bootstrapping test: n=1024, k=1, KSK beta=2, l=64 (537 MB)."""
    L = orc.lib()
    n, k, t = 1024, 1, 128
    kn = n * k
    sk = np.empty(kn, dtype=np.uint64)
    sk2 = np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(1, kn, orc.ptr(sk))
    L.orc_tlwe_keygen(2, kn, orc.ptr(sk2))
    ksk = np.empty(kn * 64 * (kn + 1), dtype=np.uint64)
    L.orc_tlwe_new_ksk(3, kn, kn, 64, 3.2, orc.ptr(sk), orc.ptr(sk2), 1, orc.ptr(ksk))
    return dict(n=n, k=k, t=t, kn=kn, sk=sk, sk2=sk2, ksk=ksk)


@pytest.mark.parametrize("kn_in,kn_out,l,batch", [(16, 16, 64, 5), (16, 8, 16, 19), (100, 130, 64, 17), (1, 1, 1, 3)])
def test_key_switch_small(fhe, orc, kn_in, kn_out, l, batch):
    ksk = orc.uniform(kn_in * 1000 + l, kn_in * l * (kn_out + 1))
    ct = orc.uniform(kn_in + 7, (batch, kn_in + 1))
    ct[0, :] = 2**64 - 1
    k = fhe.Ksk(kn_in, kn_out, l, ksk)
    got = k.key_switch(ct)
    want = orc.key_switch(kn_in, kn_out, l, ksk, ct.reshape(-1)).reshape(batch, kn_out + 1)
    assert (got == want).all()


def test_key_switch_functional(fhe, orc):
    # tfhe/src/tlwe.rs:423-463: k=16, n=1, t=128
    L = orc.lib()
    kn, t, l = 16, 128, 64
    sk = np.empty(kn, dtype=np.uint64)
    sk2 = np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(1, kn, orc.ptr(sk))
    L.orc_tlwe_keygen(2, kn, orc.ptr(sk2))
    ksk = np.empty(kn * l * (kn + 1), dtype=np.uint64)
    L.orc_tlwe_new_ksk(3, kn, kn, l, 3.2, orc.ptr(sk), orc.ptr(sk2), 0, orc.ptr(ksk))
    delta = (2**64 - 1) // t
    K = fhe.Ksk(kn, kn, l, ksk)
    for m in (0, 1, 77, 127):
        ct = np.empty(kn + 1, dtype=np.uint64)
        L.orc_tlwe_encrypt_s(10 + m, kn, 3.2, orc.ptr(sk), (m * delta) % 2**64, 1, orc.ptr(ct))
        out = K.key_switch(ct).reshape(-1)
        assert (out == orc.key_switch(kn, kn, l, ksk, ct)).all()
        p = L.orc_tlwe_decrypt(kn, orc.ptr(sk2), orc.ptr(out))
        assert L.orc_t64_mul_div_round(p, t, 2**64 - 1) % t == m


def test_sample_extract_and_mod_switch(fhe, orc):
    L = orc.lib()
    n, k = 64, 4
    ct = orc.uniform(3, (5, (k + 1) * n))
    for h in (0, 1, 17, 63):
        got = fhe.sample_extract(n, k, ct, h)
        for b in range(5):
            want = np.empty(k * n + 1, dtype=np.uint64)
            L.orc_tglwe_sample_extraction(n, k, orc.ptr(np.ascontiguousarray(ct[b])), h, orc.ptr(want))
            assert (got[b] == want).all()
    x = orc.uniform(9, 1025)
    for q2 in (2, 1024, 2**20):
        want = np.empty_like(x)
        L.orc_tlwe_mod_switch(1024, orc.ptr(x), q2, orc.ptr(want))
        assert (fhe.tlwe_mod_switch(x, q2) == want).all()


def test_bootstrapping_reference_params(fhe, orc, p5):
    # tfhe/src/tlwe.rs:465-504 with the trivial lookup table (compute_lookup_table) ...
    L = orc.lib()
    n, k, t, kn = p5["n"], p5["k"], p5["t"], p5["kn"]
    table = orc.lookup_table(n, k, t)
    delta = (2**64 - 1) // t
    K = fhe.Ksk(kn, kn, 64, p5["ksk"])
    msgs = [0, 5, 100, 127, 64, 1]
    cts = np.empty((len(msgs), kn + 1), dtype=np.uint64)
    for i, m in enumerate(msgs):
        L.orc_tlwe_encrypt_s(10 + m, kn, 3.2, orc.ptr(p5["sk"]), (m * delta) % 2**64, 0, orc.ptr(cts[i]))
    got = fhe.bootstrap(n, k, K, table, cts, kn)
    want = orc.bootstrapping(n, k, p5["ksk"], table, cts.reshape(-1), kn, threads=8).reshape(len(msgs), kn + 1)
    assert (got == want).all()
    for i, m in enumerate(msgs):  # the reference's functional assertion: the message survives
        p = L.orc_tlwe_decrypt(kn, orc.ptr(p5["sk2"]), orc.ptr(np.ascontiguousarray(got[i])))
        assert L.orc_t64_mul_div_round(p, t, 2**64 - 1) % t == m
    # ... and with a dense (non-trivial) table so that the extracted mask has dense digits
    dense = orc.uniform(77, (k + 1) * n)
    rnd = orc.uniform(78, (9, kn + 1))
    got = fhe.bootstrap(n, k, K, dense, rnd, kn)
    want = orc.bootstrapping(n, k, p5["ksk"], dense, rnd.reshape(-1), kn, threads=8).reshape(9, kn + 1)
    assert (got == want).all()
    # key switch alone at full size, ragged batch
    ks = K.key_switch(rnd)
    assert (ks.reshape(-1) == orc.key_switch(kn, kn, 64, p5["ksk"], rnd.reshape(-1), threads=8)).all()


def test_blind_rotation_as_executed_and_as_written(fhe, orc):
    L = orc.lib()
    # as executed (the CMux loop never runs): one public rotation of the table
    n, k, batch = 1024, 1, 4
    table = orc.uniform(5, (k + 1) * n)
    cts = orc.uniform(6, (batch, n * k + 1))
    got = fhe.blind_rotate(n, k, table, cts, n * k)
    for b in range(batch):
        want = np.empty((k + 1) * n, dtype=np.uint64)
        L.orc_blind_rotation_as_executed(n, k, orc.ptr(np.ascontiguousarray(cts[b])), n * k, orc.ptr(table), orc.ptr(want))
        assert (got[b] == want).all()
    # as written (extension): k=4 so the loop body runs for j=1,2,3 (k*n must be a power of two: T64::mod_switch
    # asserts it, torus.rs:58-66, so the reference panics for k=3)
    n, k, batch = 64, 4, 3
    glwe = (k + 1) * n
    bsk = orc.uniform(8, (k, (k + 1) * 64 * glwe))
    table = orc.uniform(9, glwe)
    cts = orc.uniform(10, (batch, n * k + 1))
    handles = [fhe.Tggsw(n, k, bsk[j]) for j in range(k)]
    got = fhe.blind_rotate(n, k, table, cts, n * k, bsk=handles, as_written=True)
    for b in range(batch):
        want = np.empty(glwe, dtype=np.uint64)
        L.orc_blind_rotation_as_written(n, k, orc.ptr(np.ascontiguousarray(cts[b])), n * k, orc.ptr(bsk.reshape(-1)),
                                        orc.ptr(table), orc.ptr(want))
        assert (got[b] == want).all()


@pytest.mark.parametrize("kn_in,kn_out,batch", [(100, 130, 150), (2, 1, 1), (64, 31, 129), (16, 16, 16)])
def test_key_switch_tensor_core_and_cuda_core_paths(fhe, orc, kn_in, kn_out, batch, monkeypatch):
    # l = 64: the byte-plane IMMA GEMM (ks_mma.cu) and the CUDA-core kernel must both equal the oracle
    l = 64
    ksk = orc.uniform(kn_in * 31 + kn_out, kn_in * l * (kn_out + 1))
    ksk[: kn_out + 1] = 2**64 - 1
    ct = orc.uniform(kn_in + 3, (batch, kn_in + 1))
    ct[0, :] = 2**64 - 1  # every digit set: the largest plane sums
    want = orc.key_switch(kn_in, kn_out, l, ksk, ct.reshape(-1), threads=8).reshape(batch, kn_out + 1)
    K = fhe.Ksk(kn_in, kn_out, l, ksk)
    for path in ("tc", "mma", "cuda"):
        monkeypatch.setenv("FHE_KS_PATH", path)
        assert (K.key_switch(ct) == want).all(), path
    monkeypatch.delenv("FHE_KS_PATH")


def test_key_switch_full_size_both_paths(fhe, orc, p5, monkeypatch):
    kn = p5["kn"]
    K = fhe.Ksk(kn, kn, 64, p5["ksk"])
    ct = orc.uniform(4242, (300, kn + 1))
    ct[7, :] = 2**64 - 1
    want = orc.key_switch(kn, kn, 64, p5["ksk"], ct.reshape(-1), threads=8).reshape(300, kn + 1)
    for path in ("tc", "mma", "cuda"):
        monkeypatch.setenv("FHE_KS_PATH", path)
        assert (K.key_switch(ct) == want).all(), path
    monkeypatch.delenv("FHE_KS_PATH")


def test_bootstrap_at_baseline_batch_spot_checked(fhe, orc, p5):
    # BASELINE configs[4] size (8192 TLWE inputs per GPU): every tile of the tensor-core key switch is exercised;
    # 48 rows spread over the batch (first and last tile included) are compared with the oracle, and the whole
    # batch must be invariant under a permutation of its rows (independent units: no cross-row leakage)
    n, k, kn = p5["n"], p5["k"], p5["kn"]
    batch = 8192
    K = fhe.Ksk(kn, kn, 64, p5["ksk"])
    table = orc.uniform(501, (k + 1) * n)
    cts = orc.uniform(502, (batch, kn + 1))
    got = fhe.bootstrap(n, k, K, table, cts, kn)
    rows = np.unique(np.concatenate([np.arange(8), batch - 1 - np.arange(8), orc.uniform(503, 32, batch).astype(np.int64)]))
    want = orc.bootstrapping(n, k, p5["ksk"], table, np.ascontiguousarray(cts[rows]).reshape(-1), kn, threads=8)
    assert np.array_equal(got[rows].reshape(-1), want)
    perm = np.argsort(orc.uniform(504, batch))
    got_p = fhe.bootstrap(n, k, K, table, np.ascontiguousarray(cts[perm]), kn)
    assert np.array_equal(got_p, got[perm])


def test_bootstrap_at_64k_batch_both_cluster_modes(fhe, orc, p5, monkeypatch):
    # upper end of BASELINE configs[4] (8k-64k TLWE inputs): 65536 ciphertexts on one GPU, device-resident; rows spread
    # over the batch against the oracle, and the whole result equal to eight 8192-row calls (independent units).
    # The cluster size of the tensor-core kernel is read once per process, so the single-CTA variant is checked in a
    # child process on a smaller batch.
    import torch

    n, k, kn = p5["n"], p5["k"], p5["kn"]
    batch = 65536
    K = fhe.Ksk(kn, kn, 64, p5["ksk"])
    table = orc.uniform(601, (k + 1) * n)
    g = torch.Generator(device="cuda").manual_seed(602)
    cts = torch.randint(-(2**63), 2**63 - 1, (batch, kn + 1), dtype=torch.int64, device="cuda", generator=g)
    dt = torch.from_numpy(table.view(np.int64)).cuda()
    fhe.use_torch_stream()
    got = fhe.bootstrap(n, k, K, dt, cts, kn)
    rows = np.unique(np.concatenate([np.arange(4), batch - 1 - np.arange(4), orc.uniform(603, 24, batch).astype(np.int64)]))
    sel = cts[torch.from_numpy(rows).cuda()].cpu().numpy().view(np.uint64)
    want = orc.bootstrapping(n, k, p5["ksk"], table, np.ascontiguousarray(sel).reshape(-1), kn, threads=8)
    assert np.array_equal(got[torch.from_numpy(rows).cuda()].cpu().numpy().view(np.uint64).reshape(-1), want)
    for i in range(0, batch, 8192):
        assert torch.equal(fhe.bootstrap(n, k, K, dt, cts[i:i + 8192].contiguous(), kn), got[i:i + 8192])
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); import numpy as np, oracle, fhe_study_b200 as fhe\n"
            "kn, l = 1024, 64\nksk = oracle.uniform(4, kn * l * (kn + 1)); K = fhe.Ksk(kn, kn, l, ksk)\n"
            "ct = oracle.uniform(6, (700, kn + 1)); got = K.key_switch(ct); rows = [0, 255, 256, 511, 699]\n"
            "want = oracle.key_switch(kn, kn, l, ksk, ct[rows].copy(), threads=8).reshape(len(rows), kn + 1)\n"
            "assert (got[rows] == want).all(); print('single-CTA ok')")
    import os

    # FHE_KS_CLUSTER=1: single CTAs; =3: CTA pairs (tcgen05.mma.cta_group::2, each CTA holding half of every key block)
    for mode in ("1", "3"):
        env = dict(os.environ, FHE_KS_CLUSTER=mode)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300,
                           cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        assert r.returncode == 0 and "single-CTA ok" in r.stdout, mode + r.stdout + r.stderr


@pytest.mark.parametrize("kn_in,kn_out,l,uniform", [(16, 16, 64, True), (5, 33, 7, False), (64, 64, 64, False), (3, 1, 1, True)])
def test_ksk_generated_on_device_equals_cpu_restatement(fhe, orc, kn_in, kn_out, l, uniform):
    # TLWE::new_ksk (tlwe.rs:84-100) on the device, counter-based sampler: bit-exact against the
    # oracle's orc_tlwe_new_ksk_ctr (masks, bodies, the f64 error path included)
    sk = orc.uniform(1, kn_in) & np.uint64(1)
    sk2 = orc.uniform(2, kn_out) & np.uint64(1)
    K = fhe.Ksk.generate(kn_in, kn_out, l, sk, sk2, sigma=3.2, seed=1234 + l, uniform_mask=uniform)
    want = orc.tlwe_new_ksk_ctr(1234 + l, kn_in, kn_out, l, 3.2, sk, sk2, uniform)
    assert np.array_equal(K.export(), want)
    # and the generated handle behaves like a loaded one
    ct = orc.uniform(3, (37, kn_in + 1))
    assert np.array_equal(K.key_switch(ct).reshape(-1), orc.key_switch(kn_in, kn_out, l, want, ct.reshape(-1)))


def test_device_generated_ksk_switches_keys_functionally(fhe, orc):
    # the reference's test_key_switch property (tlwe.rs:423-463) with a key that never left the GPU, at n=1024
    L = orc.lib()
    kn, t = 1024, 128
    sk, sk2 = np.empty(kn, dtype=np.uint64), np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(11, kn, orc.ptr(sk))
    L.orc_tlwe_keygen(12, kn, orc.ptr(sk2))
    K = fhe.Ksk.generate(kn, kn, 64, sk, sk2, sigma=3.2, seed=99, uniform_mask=True)
    delta = (2**64 - 1) // t
    msgs = [0, 1, 77, 127]
    cts = np.empty((len(msgs), kn + 1), dtype=np.uint64)
    for i, m in enumerate(msgs):
        L.orc_tlwe_encrypt_s(20 + m, kn, 3.2, orc.ptr(sk), (m * delta) % 2**64, 1, orc.ptr(cts[i]))
    out = K.key_switch(cts)
    for i, m in enumerate(msgs):
        p = L.orc_tlwe_decrypt(kn, orc.ptr(sk2), orc.ptr(np.ascontiguousarray(out[i])))
        assert L.orc_t64_mul_div_round(p, t, 2**64 - 1) % t == m


def test_tlwe_and_tglwe_decrypt_decode(fhe, orc):
    # TLWE / TGLWE decrypt + decode (tlwe.rs:60-63,80-82; tglwe.rs:59-63,86-88) batched on the device: bit-exact phases
    # against the oracle and the reference's functional property decode(decrypt(encrypt(m))) == m
    L = orc.lib()
    kn, t, batch = 1024, 128, 70
    sk = np.empty(kn, dtype=np.uint64)
    L.orc_tlwe_keygen(1, kn, orc.ptr(sk))
    delta = (2**64 - 1) // t
    msgs = orc.uniform(2, batch, t)
    cts = np.empty((batch, kn + 1), dtype=np.uint64)
    for i in range(batch):
        L.orc_tlwe_encrypt_s(10 + i, kn, 3.2, orc.ptr(sk), (int(msgs[i]) * delta) % 2**64, 0, orc.ptr(cts[i]))
    cts[-1] = orc.uniform(3, kn + 1)  # a random one: parity without any structure
    ph = fhe.tlwe_decrypt(kn, sk, cts)
    want = np.array([L.orc_tlwe_decrypt(kn, orc.ptr(sk), orc.ptr(np.ascontiguousarray(cts[i]))) for i in range(batch)], dtype=np.uint64)
    assert np.array_equal(ph, want)
    assert np.array_equal(fhe.torus_decode(ph, t)[:-1], msgs[:-1])
    for n, k in ((64, 4), (1024, 1), (16, 2)):
        glwe = (k + 1) * n
        skp = np.empty(k * n, dtype=np.uint64)
        L.orc_tglwe_keygen(4, n, k, orc.ptr(skp))
        m = orc.uniform(5, (6, n), t)
        ct = np.empty((6, glwe), dtype=np.uint64)
        for i in range(6):
            pt = np.empty(n, dtype=np.uint64)
            L.orc_tglwe_encode(n, t, orc.ptr(np.ascontiguousarray(m[i])), orc.ptr(pt))
            L.orc_tglwe_encrypt_s(20 + i, n, k, 3.2, orc.ptr(skp), orc.ptr(pt), 0, orc.ptr(ct[i]))
        ct[-1] = orc.uniform(6, glwe)
        p = fhe.tglwe_decrypt(n, k, skp, ct)
        wantp = np.empty((6, n), dtype=np.uint64)
        for i in range(6):
            L.orc_tglwe_decrypt(n, k, orc.ptr(skp), orc.ptr(np.ascontiguousarray(ct[i])), orc.ptr(wantp[i]))
        assert np.array_equal(p, wantp)
        assert np.array_equal(fhe.torus_decode(p, t)[:-1], m[:-1])


def test_tlwe_encrypt_on_device_roundtrip(fhe, orc):
    # TLWE::encrypt_s on the device (counter-based sampler): bit-exact against the oracle, and decrypt+decode gives m back
    kn, t, batch = 630, 16, 333
    sk = orc.uniform(1, kn) & np.uint64(1)
    msgs = orc.uniform(2, batch, t)
    enc = msgs * np.uint64((2**64 - 1) // t)
    for uniform in (True, False):
        ct = fhe.tlwe_encrypt(kn, sk, enc, sigma=3.2, seed=42, uniform_mask=uniform)
        want = np.empty((batch, kn + 1), dtype=np.uint64)
        orc.lib().orc_tlwe_encrypt_ctr(42, kn, 3.2, orc.ptr(sk), orc.ptr(enc), batch, int(uniform), orc.ptr(want))
        assert np.array_equal(ct, want)
        assert np.array_equal(fhe.torus_decode(fhe.tlwe_decrypt(kn, sk, ct), t), msgs)


@pytest.mark.parametrize("n,k", [(64, 4), (1024, 1), (16, 2)])
def test_tglwe_encrypt_on_device_roundtrip(fhe, orc, n, k):
    t, batch = 16, 9
    sk = orc.uniform(1, k * n) & np.uint64(1)
    msgs = orc.uniform(2, (batch, n), t)
    enc = msgs * np.uint64((2**64 - 1) // t)  # TGLWE::encode (tglwe.rs:49-58)
    for uniform in (True, False):
        ct = fhe.tglwe_encrypt(n, k, sk, enc, sigma=3.2, seed=7, uniform_mask=uniform)
        want = np.empty((batch, (k + 1) * n), dtype=np.uint64)
        orc.lib().orc_tglwe_encrypt_ctr(7, n, k, 3.2, orc.ptr(sk), orc.ptr(enc), batch, int(uniform), orc.ptr(want))
        assert np.array_equal(ct, want)
        assert np.array_equal(fhe.torus_decode(fhe.tglwe_decrypt(n, k, sk, ct), t), msgs)
