"""CPU: the oracle must reproduce the frozen golden vectors (tests/golden/) and the reference's own
known-answer vectors (tests/golden/reference_kats.json, each citing the reference file:line)."""
import json
import os

import numpy as np
import pytest

from golden import make_golden as G

KATS = json.load(open(os.path.join(os.path.dirname(G.__file__), "reference_kats.json")))
M64 = 2**64


def test_splitmix_numpy_equals_oracle_generator(orc):
    for seed, count, mod in [(1, 1000, 0), (7, 513, 65537), (2**40 + 3, 64, 65537**3)]:
        assert np.array_equal(G.splitmix64(seed, count, mod), orc.uniform(seed, count, mod))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_reproduces_golden(orc, name):
    gold = G.load()
    for key, arr in G.oracle_outputs(orc, name).items():
        assert np.array_equal(np.asarray(arr, dtype=np.uint64).reshape(-1), gold["%s/%s" % (name, key)]), (name, key)


def test_golden_ntt_roundtrip_is_identity():
    gold = G.load()
    for name, (kind, p, _) in G.CASES.items():
        if kind == "ntt":
            assert np.array_equal(gold[name + "/inv_of_fwd"], G.inputs_for(name)[0])  # arith/src/ntt.rs:194-234


def test_reference_kats_on_oracle(orc):
    L = orc.lib()
    for c in KATS["rq_mul"]:
        assert list(orc.rq_mul(c["q"], c["n"], c["a"], c["b"])) == c["c"], c["cite"]
    for c in KATS["rq_decompose"]:
        out = np.zeros((c["l"], c["n"]), dtype=np.uint64)
        L.orc_rq_decompose(c["q"], c["n"], orc.ptr(orc.u64(c["a"])), c["beta"], c["l"], orc.ptr(out))
        assert out.tolist() == c["d"], c["cite"]
    for c in KATS["r_linear_mul_folded"]:
        a, b = orc.i64(c["a"]), orc.i64(c["b"])
        out = np.zeros(2 * c["n"] - 1, dtype=np.int64)
        L.orc_r_naive_mul(c["n"], orc.ptr(a), orc.ptr(b), orc.ptr(out))
        ln = L.orc_r_fold(c["n"], orc.ptr(out), out.size)
        assert list(out[:ln]) == c["c"], c["cite"]
    for c in KATS["tn_left_rotate"]:
        a = orc.u64([x % M64 for x in c["a"]])
        assert list(orc.tn_left_rotate(c["n"], a, c["h"])) == [x % M64 for x in c["c"]], c["cite"]
    for c in KATS["ntt_plan_derived"]:
        roots, roots_inv, n_inv = orc.ntt_tables(c["q"], c["n"])
        assert (list(roots), list(roots_inv), n_inv) == (c["roots"], c["roots_inv"], c["n_inv"])
        assert list(orc.ntt(c["q"], c["n"], [1, 2, 3, 4])) == c["ntt_of_1234"]
        assert list(orc.ntt(c["q"], c["n"], [0, 0, 0, 2])) == c["ntt_of_0002"]
