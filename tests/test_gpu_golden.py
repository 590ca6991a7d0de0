"""GPU parity against the committed golden fixtures (tests/golden/): the CUDA path through the C ABI must
reproduce the frozen vectors and the reference's own known-answer vectors.  The oracle is NOT executed
here -- inputs come from the numpy SplitMix64 in tests/golden/make_golden.py, outputs from the .npz."""
import json
import os

import numpy as np
import pytest

from golden import make_golden as G

pytestmark = pytest.mark.gpu
KATS = json.load(open(os.path.join(os.path.dirname(G.__file__), "reference_kats.json")))
M64 = 2**64


@pytest.fixture(scope="module")
def fhe():
    import fhe_study_b200 as f

    f.set_device(0)
    return f


def _run(fhe, name):
    kind, p, _ = G.CASES[name]
    x = G.inputs_for(name)
    if kind == "ntt":
        plan = fhe.NttPlan(p["q"], p["n"])
        f = plan.ntt(x[0])
        return {"fwd": f, "inv_of_fwd": plan.intt(f)}
    if kind == "rqmul":
        return {"c": fhe.NttPlan(p["q"], p["n"]).mul(x[0], x[1])}
    if kind == "tnmul":
        return {"c": fhe.tn_mul(p["n"], x[0], x[1])}
    if kind == "extprod":
        return {"out": fhe.Tggsw(p["n"], p["k"], x[0]).extprod(x[1])}
    if kind == "cmux":
        return {"out": fhe.Tggsw(p["n"], p["k"], x[0]).cmux(x[1], x[2])}
    if kind == "keyswitch":
        return {"out": fhe.Ksk(p["kn_in"], p["kn_out"], p["l"], x[0]).key_switch(x[1])}
    if kind == "bootstrap":
        kn = p["n"] * p["k"]
        return {"out": fhe.bootstrap(p["n"], p["k"], fhe.Ksk(kn, kn, 64, x[0]), x[1], x[2], kn)}
    if kind == "bfvmul":
        return {"out": fhe.bfv_mul_relin(p["q"], p["n"], p["t"], p["pq"], x[0], x[1], x[2])}
    raise KeyError(kind)


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_cuda_reproduces_golden(fhe, name):
    gold = G.load()
    l0 = fhe.launch_count()
    for key, arr in _run(fhe, name).items():
        assert np.array_equal(np.asarray(arr, dtype=np.uint64).reshape(-1), gold["%s/%s" % (name, key)]), (name, key)
    assert fhe.launch_count() > l0


def test_reference_kats_on_cuda(fhe):
    u = lambda v: np.array([x % M64 for x in v], dtype=np.uint64)
    for c in KATS["rq_mul"]:  # arith/src/ring_nq.rs:668-704
        assert list(fhe.NttPlan(c["q"], c["n"]).mul(u(c["a"]), u(c["b"]))) == c["c"], c["cite"]
    for c in KATS["rq_decompose"]:  # arith/src/ring_nq.rs:707-729
        assert fhe.rq_decompose(c["q"], c["n"], u(c["a"]), c["beta"], c["l"]).reshape(c["l"], c["n"]).tolist() == c["d"]
    for c in KATS["tn_left_rotate"]:  # arith/src/ring_torus.rs:334-366
        assert list(fhe.tn_left_rotate(c["n"], u(c["a"]), np.array([c["h"]], dtype=np.uint64))) == list(u(c["c"]))
    for c in KATS["ntt_plan_derived"]:
        plan = fhe.NttPlan(c["q"], c["n"])
        psi, n_inv, roots, roots_inv = plan.info()
        assert (psi, n_inv, list(roots), list(roots_inv)) == (c["psi"], c["n_inv"], c["roots"], c["roots_inv"])
        assert list(plan.ntt(u([1, 2, 3, 4]))) == c["ntt_of_1234"]
        assert list(plan.ntt(u([0, 0, 0, 2]))) == c["ntt_of_0002"]
