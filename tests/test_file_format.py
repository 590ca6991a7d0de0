"""Round trips of the flat key / ciphertext container (include/fhe_b200_file.h): every encoding, header
validation, checksum, truncation.  Host-side code only: runs without a GPU."""
import os

import numpy as np
import pytest

Q17 = 65537


def test_round_trip_every_encoding(tmp_path):
    import fhe_study_b200 as fhe

    rng = np.random.default_rng(1)
    n, count = 64, 5
    rq = rng.integers(0, Q17, size=(count, n), dtype=np.uint64)
    for enc, bits in ((fhe.ENC_U64, 0), (fhe.ENC_U32, 0), (fhe.ENC_PACKED, 17), (fhe.ENC_PACKED, 24)):
        path = str(tmp_path / f"rq_{enc}_{bits}.fheb")
        info = fhe.save(path, "rq", rq, n, q=Q17, n=n, encoding=enc, bits=bits)
        want_bytes = {fhe.ENC_U64: 8 * n * count, fhe.ENC_U32: 4 * n * count}.get(enc, n // 32 * bits * 4 * count)
        assert info["payload_bytes"] == want_bytes and os.path.getsize(path) == 96 + want_bytes
        got_info, got = fhe.load(path)
        assert (got == rq).all() and got_info["q"] == Q17 and got_info["n"] == n and got_info["kind"] == fhe.FILE_KINDS["rq"]
    # torus objects: a KSK and a TGGSW in the layouts fhe_ksk_load / fhe_tggsw_load take
    kn, l = 8, 64
    ksk = rng.integers(0, 2**64, size=kn * l * (kn + 1), dtype=np.uint64)
    path = str(tmp_path / "ksk.fheb")
    fhe.save(path, "ksk", ksk, kn + 1, q=0, n=kn, k=kn, l=l)
    info, got = fhe.load(path)
    assert (got.reshape(-1) == ksk).all() and info["count"] == kn * l and info["l"] == l


def test_rejects_corruption_and_bad_headers(tmp_path):
    import fhe_study_b200 as fhe

    a = np.arange(64, dtype=np.uint64)
    path = str(tmp_path / "a.fheb")
    fhe.save(path, "tn", a, 64, q=0, n=64)
    raw = bytearray(open(path, "rb").read())
    raw[100] ^= 1  # flip a payload bit
    bad = str(tmp_path / "bad.fheb")
    open(bad, "wb").write(raw)
    with pytest.raises(fhe.FheError):
        fhe.load(bad)
    open(bad, "wb").write(bytes(raw[:-8]))  # truncated
    with pytest.raises(fhe.FheError):
        fhe.load(bad)
    raw2 = bytearray(open(path, "rb").read())
    raw2[0] = ord("X")  # magic
    open(bad, "wb").write(raw2)
    with pytest.raises(fhe.FheError):
        fhe.load(bad)
    with pytest.raises(fhe.FheError):  # packed needs q <= 2^bits
        fhe.save(str(tmp_path / "c.fheb"), "rq", np.zeros(32, dtype=np.uint64), 32, q=Q17, n=32, encoding=fhe.ENC_PACKED, bits=16)
    with pytest.raises(fhe.FheError):  # u32 words cannot hold torus elements
        fhe.save(str(tmp_path / "d.fheb"), "tn", np.zeros(32, dtype=np.uint64), 32, q=0, n=32, encoding=fhe.ENC_U32)
