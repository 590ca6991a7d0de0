"""GPU parity of the CMux chain / blind rotation with one TGGSW per mask element (extension of the loop at
tfhe/src/tlwe.rs:138-147) against the oracle's composition of the reference's own cmux and
left_rotate, through the C ABI: fused persistent kernel and the per-step fallback, both rotation modes, plus the
functional property "a working PBS recovers the message"."""
import numpy as np
import pytest

from test_oracle_kats import pbs_fixture

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


@pytest.mark.parametrize("n,k,steps,batch", [(64, 4, 3, 3), (64, 1, 5, 2), (1024, 1, 2, 2), (256, 2, 2, 1), (32, 1, 3, 2)])
@pytest.mark.parametrize("neg", [False, True])
def test_cmux_chain_matches_oracle(fhe, orc, monkeypatch, n, k, steps, batch, neg):
    glwe = (k + 1) * n
    size = (k + 1) * 64 * glwe
    bsk = orc.uniform(1000 + n + k, steps * size)
    acc = orc.uniform(2000 + n, (batch, glwe))
    h = orc.uniform(3000 + n, (batch, steps)) % np.uint64(4 * n)  # beyond 2n on purpose: reduced mod n or 2n
    h[0, 0] = 0
    want = orc.cmux_chain(n, k, bsk, acc, h, negacyclic=neg)
    handles = [fhe.Tggsw(n, k, bsk[j * size:(j + 1) * size]) for j in range(steps)]
    got = fhe.cmux_chain(n, k, handles, acc, h, negacyclic=neg)
    assert np.array_equal(got, want)
    monkeypatch.setenv("FHE_EXTPROD_PATH", "unfused")  # one rotate + CMux launch pair per step
    assert np.array_equal(fhe.cmux_chain(n, k, handles, acc, h, negacyclic=neg), want)
    monkeypatch.delenv("FHE_EXTPROD_PATH")
    assert np.array_equal(fhe.cmux_chain(n, k, [], acc, None), acc)  # empty chain


def test_bootstrap_chain_both_modes_and_functional(fhe, orc):
    n, k, m_lwe, t = 64, 1, 6, 4
    s, z, bsk, table, cts = pbs_fixture(orc, n, k, m_lwe, t)
    handles = [fhe.Tggsw(n, k, bsk[j]) for j in range(m_lwe)]
    for mode in (0, 1):
        want = orc.bootstrap_chain(n, k, m_lwe, bsk, None, table, cts, m_lwe, mode)
        got = fhe.bootstrap_chain(n, k, handles, table, cts, m_lwe, mode=mode)
        assert np.array_equal(got, want), mode
    # functional: the working mode recovers the message under the GLWE key (as in tfhe/src/tlwe.rs:465-504)
    delta = (2**64 - 1) // t
    for m in range(t):
        phase = int(orc.lib().orc_tlwe_decrypt(k * n, orc.ptr(z), orc.ptr(np.ascontiguousarray(got[m]))))
        assert round(phase / delta) % t == m
    # with a key switch behind it, fewer TGGSWs than mask elements, random (dense) inputs
    kn = k * n
    ksk = orc.uniform(91, kn * 64 * (kn + 1))
    K = fhe.Ksk(kn, kn, 64, ksk)
    rnd = orc.uniform(92, (3, m_lwe + 3))
    dense = orc.uniform(93, (k + 1) * n)
    for mode in (0, 1):
        want = orc.bootstrap_chain(n, k, m_lwe, bsk, ksk, dense, rnd, m_lwe + 2, mode)
        got = fhe.bootstrap_chain(n, k, handles, dense, rnd, m_lwe + 2, mode=mode, ksk=K)
        assert np.array_equal(got, want), mode


def test_chain_argument_errors(fhe, orc):
    n, k = 64, 1
    g = fhe.Tggsw(n, k, orc.uniform(1, (k + 1) * 64 * (k + 1) * n))
    g2 = fhe.Tggsw(32, k, orc.uniform(1, (k + 1) * 64 * (k + 1) * 32))
    acc = orc.uniform(2, (1, (k + 1) * n))
    with pytest.raises(RuntimeError):
        fhe.cmux_chain(n, k, [g, g2], acc, np.zeros(2, dtype=np.uint64))  # mismatched handle
    with pytest.raises(RuntimeError):
        fhe.bootstrap_chain(n, k, [g, g], orc.uniform(3, (k + 1) * n), orc.uniform(4, (1, 2)), 1)  # steps > c_kn


@pytest.mark.parametrize("n,k,uniform", [(64, 1, True), (64, 4, False), (1024, 1, True), (16, 2, True)])
def test_tggsw_generated_on_device_equals_cpu_restatement(fhe, orc, n, k, uniform):
    # TGGSW::encrypt_s (tggsw.rs:17-33) sampled on the device; rows bit-exact against the oracle's
    # counter-based restatement, and the resident handle behaves like a loaded one
    sk = orc.uniform(1, k * n) & np.uint64(1)
    m = orc.uniform(2, n) & np.uint64(1)
    rows = np.empty((k + 1) * 64 * (k + 1) * n, dtype=np.uint64)
    g = fhe.Tggsw.generate(n, k, sk, m, sigma=3.2, seed=77 + n, uniform_mask=uniform, rows_out=rows)
    want = orc.tggsw_encrypt_s_ctr(77 + n, n, k, 3.2, sk, m, uniform)
    assert np.array_equal(rows, want)
    ct = orc.uniform(3, (2, (k + 1) * n))
    assert np.array_equal(g.extprod(ct), fhe.Tggsw(n, k, want).extprod(ct))


def test_bootstrap_chain_with_device_generated_keys(fhe, orc):
    # a whole programmable bootstrap whose bootstrapping key never left the GPU recovers the message
    n, k, m_lwe, t = 64, 1, 6, 4
    s, z, _, table, cts = pbs_fixture(orc, n, k, m_lwe, t)
    bsk = []
    for j in range(m_lwe):
        msg = np.zeros(n, dtype=np.uint64)
        msg[0] = s[j]
        bsk.append(fhe.Tggsw.generate(n, k, z, msg, sigma=3.2, seed=500 + j))
    out = fhe.bootstrap_chain(n, k, bsk, table, cts, m_lwe, mode=1)
    delta = (2**64 - 1) // t
    for m in range(t):
        phase = int(orc.lib().orc_tlwe_decrypt(k * n, orc.ptr(z), orc.ptr(np.ascontiguousarray(out[m]))))
        assert round(phase / delta) % t == m


def test_whole_tfhe_pipeline_on_the_gpu(fhe, orc):
    # keys sampled on the device (fhe_tggsw_generate, fhe_ksk_generate), inputs encrypted on the device (fhe_tlwe_encrypt),
    # programmable bootstrap = CMux chain + sample extraction + key switch, outputs decrypted and decoded on the device:
    # the messages come back (the functional property of tfhe/src/tlwe.rs:465-504 for a blind rotation that really runs)
    n, k, m_lwe, t, batch = 64, 1, 8, 4, 40
    kn = n * k
    s = orc.uniform(1, m_lwe) & np.uint64(1)       # input LWE key
    z = orc.uniform(2, kn) & np.uint64(1)          # GLWE key (k polynomials) = extracted TLWE key
    s2 = orc.uniform(3, kn) & np.uint64(1)         # key after the key switch
    bsk = []
    for j in range(m_lwe):
        msg = np.zeros(n, dtype=np.uint64)
        msg[0] = s[j]
        bsk.append(fhe.Tggsw.generate(n, k, z, msg, sigma=3.2, seed=100 + j))
    K = fhe.Ksk.generate(kn, kn, 64, z, s2, sigma=3.2, seed=200)
    table = orc.lookup_table(n, k, t)
    msgs = orc.uniform(4, batch, t)
    delta_n = n // t
    enc = (msgs * np.uint64(delta_n) + np.uint64(delta_n // 2)) * np.uint64(2**64 // (2 * n))  # centre of the LUT window
    cts = fhe.tlwe_encrypt(m_lwe, s, enc, sigma=3.2, seed=300)
    out = fhe.bootstrap_chain(n, k, bsk, table, cts, m_lwe, mode=1, ksk=K)
    got = fhe.torus_decode(fhe.tlwe_decrypt(kn, s2, out), t)
    assert np.array_equal(got, msgs)
