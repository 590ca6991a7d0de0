"""CPU test of the host-side (de)serialisers of the bit-packed wire format (fhe_pack_bits / fhe_unpack_bits):
coefficient i occupies bits [i*bits, (i+1)*bits) of the little-endian word stream; checked against Python integers."""
import numpy as np
import pytest


@pytest.mark.parametrize("bits", [1, 5, 16, 17, 20, 24, 31, 32])
def test_pack_roundtrip_and_layout(bits):
    import fhe_study_b200 as fhe

    rng = np.random.default_rng(bits)
    a = rng.integers(0, 1 << bits, size=(4, 96), dtype=np.uint64)
    a[0, :] = (1 << bits) - 1
    w = fhe.pack_bits(bits, a)
    assert w.dtype == np.uint32 and w.shape == (4, 3 * bits)
    assert (fhe.unpack_bits(bits, w) == a).all()
    for r in range(4):
        big = 0
        for i, v in enumerate(a[r]):
            big |= int(v) << (bits * i)
        assert [int(x) for x in w[r]] == [(big >> (32 * j)) & 0xFFFFFFFF for j in range(3 * bits)]


def test_pack_rejects_wide_coefficients_and_ragged_lengths():
    import fhe_study_b200 as fhe

    with pytest.raises(fhe.FheError):
        fhe.pack_bits(17, np.full(32, 1 << 17, dtype=np.uint64))
    with pytest.raises(ValueError):
        fhe.pack_bits(17, np.zeros(33, dtype=np.uint64))
