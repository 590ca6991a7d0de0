"""Parity of the polymul variant selected by the FHE_NTT_* knobs in the environment against the oracle
(test infrastructure, like tests/): python tools/staged_check.py 13 14"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import fhe_study_b200 as fhe
import oracle

fhe.set_device(0)
for logn in [int(x) for x in sys.argv[1:]] or [13, 14]:
    for q in (65537, 0x3FFFFFFFFFFF0001):
        n, batch = 1 << logn, 2 * 148 + 5
        p = fhe.NttPlan(q, n)
        a, b = oracle.uniform(70 + logn, (batch, n), q), oracle.uniform(80 + logn, (batch, n), q)
        a[0, :] = q - 1
        b[0, :] = q - 1
        want = oracle.rq_mul_batch(q, n, a, b, threads=8)
        ok = bool((p.mul(a, b) == want).all())
        ok32 = bool((p.mul_u32(a.astype(np.uint32), b.astype(np.uint32)) == want).all()) if q < 2**32 else None
        print("q=%d n=%d config=%s u64 %s u32 %s" % (q, n, p.config(), ok, ok32), flush=True)
        assert ok and ok32 is not False
