"""Generates tests/golden/hotpath_vectors.npz -- frozen input/output vectors of the hot path.

Provenance (read before trusting): the reference is Rust and no Rust toolchain exists in the build image,
so these vectors are NOT outputs of the reference binary.  They are outputs of oracle/fhe_oracle.c (the
plain-C restatement, itself pinned to the reference's own known-answer tests by tests/test_oracle_kats.py)
frozen at the commit that added this file.  Their job is (i) to stop the oracle from drifting silently and
(ii) to give the GPU tests a comparison target that does not execute the oracle at all.
Inputs are produced by a pure-numpy SplitMix64 (`splitmix64` below), so a fixture stores only seeds/shapes
for large inputs (TGGSW, KSK) and the full output words.

    python tests/golden/make_golden.py          # rewrites hotpath_vectors.npz

Cases follow the parameter sets of the reference's tests.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "hotpath_vectors.npz")

Q = 65537
Q3 = Q * Q * Q


def splitmix64(seed: int, count: int, modulus: int = 0) -> np.ndarray:
    """Pure-numpy SplitMix64 stream; identical to orc_fill_uniform_u64 (checked by the golden test)."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z % np.uint64(modulus) if modulus else z


# every case: name -> (kind, params, input seeds).  Inputs are regenerated with inputs_for().
CASES = {
    "ntt_q65537_n4": ("ntt", dict(q=Q, n=4, batch=1), [101]),
    "ntt_q65537_n512": ("ntt", dict(q=Q, n=512, batch=2), [102]),          # arith/src/ntt.rs:219
    "ntt_q65537_n1024": ("ntt", dict(q=Q, n=1024, batch=2), [103]),        # BASELINE configs[0]
    "ntt_q62_n2048": ("ntt", dict(q=4611686018427322369, n=2048, batch=1), [104]),
    "rqmul_q65537_n1024": ("rqmul", dict(q=Q, n=1024, batch=3), [111, 112]),
    "rqmul_q65537_n16384": ("rqmul", dict(q=Q, n=16384, batch=1), [113, 114]),
    "rqmul_q62_n1024": ("rqmul", dict(q=4611686018427322369, n=1024, batch=1), [115, 116]),
    "tnmul_n64": ("tnmul", dict(n=64, batch=3), [121, 122]),
    "tnmul_n1024": ("tnmul", dict(n=1024, batch=1), [123, 124]),
    "extprod_n64_k4": ("extprod", dict(n=64, k=4, batch=2), [131, 132]),   # tfhe/src/tggsw.rs:159-167
    "cmux_n64_k4": ("cmux", dict(n=64, k=4, batch=2), [133, 134, 135]),
    "cmux_n1024_k1": ("cmux", dict(n=1024, k=1, batch=1), [136, 137, 138]),  # tfhe/src/tlwe.rs:467-475
    "keyswitch_kn64": ("keyswitch", dict(kn_in=64, kn_out=32, l=64, batch=3), [141, 142]),
    "bootstrap_n64_k1": ("bootstrap", dict(n=64, k=1, batch=3), [151, 152, 153]),
    "bootstrap_n256_k1": ("bootstrap", dict(n=256, k=1, batch=2), [154, 155, 156]),
    "bfvmul_n16_t2": ("bfvmul", dict(q=Q, n=16, t=2, pq=Q3, batch=8), [161, 162, 163]),  # bfv/src/lib.rs:559-564
    "bfvmul_n64_t8": ("bfvmul", dict(q=Q, n=64, t=8, pq=Q3, batch=2), [164, 165, 166]),
}


def inputs_for(name: str):
    kind, p, s = CASES[name]
    if kind == "ntt":
        return (splitmix64(s[0], p["batch"] * p["n"], p["q"]),)
    if kind == "rqmul":
        return tuple(splitmix64(x, p["batch"] * p["n"], p["q"]) for x in s)
    if kind == "tnmul":
        return tuple(splitmix64(x, p["batch"] * p["n"]) for x in s)
    if kind in ("extprod", "cmux"):
        glwe = (p["k"] + 1) * p["n"]
        tggsw = splitmix64(s[0], (p["k"] + 1) * 64 * glwe)
        return (tggsw,) + tuple(splitmix64(x, p["batch"] * glwe) for x in s[1:])
    if kind == "keyswitch":
        return (splitmix64(s[0], p["kn_in"] * p["l"] * (p["kn_out"] + 1)), splitmix64(s[1], p["batch"] * (p["kn_in"] + 1)))
    if kind == "bootstrap":
        kn = p["n"] * p["k"]
        return (splitmix64(s[0], kn * 64 * (kn + 1)), splitmix64(s[1], (p["k"] + 1) * p["n"]), splitmix64(s[2], p["batch"] * (kn + 1)))
    if kind == "bfvmul":
        return (splitmix64(s[0], 2 * p["n"], p["pq"]), splitmix64(s[1], p["batch"] * 2 * p["n"], p["q"]),
                splitmix64(s[2], p["batch"] * 2 * p["n"], p["q"]))
    raise KeyError(kind)


def oracle_outputs(orc, name: str):
    """What the CPU oracle computes for a case (dict of named output arrays)."""
    kind, p, _ = CASES[name]
    x = inputs_for(name)
    if kind == "ntt":
        f = orc.ntt(p["q"], p["n"], x[0])
        return {"fwd": f, "inv_of_fwd": orc.ntt(p["q"], p["n"], f, inverse=True)}
    if kind == "rqmul":
        return {"c": orc.rq_mul_batch(p["q"], p["n"], x[0], x[1])}
    if kind == "tnmul":
        return {"c": orc.tn_mul(p["n"], x[0], x[1])}
    if kind == "extprod":
        return {"out": orc.extprod(p["n"], p["k"], x[0], x[1], fast=False)}
    if kind == "cmux":
        return {"out": orc.cmux(p["n"], p["k"], x[0], x[1], x[2], fast=p["n"] > 64)}
    if kind == "keyswitch":
        return {"out": orc.key_switch(p["kn_in"], p["kn_out"], p["l"], x[0], x[1])}
    if kind == "bootstrap":
        return {"out": orc.bootstrapping(p["n"], p["k"], x[0], x[1], x[2], p["n"] * p["k"])}
    if kind == "bfvmul":
        return {"out": orc.bfv_mul(p["q"], p["n"], p["t"], p["pq"], x[0], x[1], x[2])}
    raise KeyError(kind)


def load():
    z = np.load(OUT)
    return {k: z[k] for k in z.files}


def main():
    sys.path.insert(0, ROOT)
    import oracle

    blob = {}
    for name in CASES:
        for key, arr in oracle_outputs(oracle, name).items():
            blob["%s/%s" % (name, key)] = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1)
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes,", len(blob), "arrays")


if __name__ == "__main__":
    main()
