"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize.py   (where the tool is available; it is closed on the graft GPU pool)
Sizes are tiny on purpose (the tools slow kernels down by 10-100x); results are still checked against the oracle."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import fhe_study_b200 as fhe
import oracle

fhe.set_device(0)
Q = 65537
for q, n in ((Q, 1024), (Q, 2048), (0x3FFFFFFFFFFF0001, 1024), (Q, 64)):
    plan = fhe.NttPlan(q, n)
    a, b = oracle.uniform(1, (3, n), q), oracle.uniform(2, (3, n), q)
    assert np.array_equal(plan.mul(a, b), oracle.rq_mul_batch(q, n, a, b))
    assert np.array_equal(plan.intt(plan.ntt(a)), a)
plan = fhe.NttPlan(Q, 1024)
a32 = oracle.uniform(3, (5, 1024), Q).astype(np.uint32)
assert np.array_equal(plan.intt_u32(plan.ntt_u32(a32)), a32)
for n, k, batch in ((64, 4, 5), (1024, 1, 2), (256, 2, 3), (2048, 1, 1)):
    glwe = (k + 1) * n
    rows = oracle.uniform(5, (k + 1) * 64 * glwe)
    ct1, ct2 = oracle.uniform(6, (batch, glwe)), oracle.uniform(7, (batch, glwe))
    g = fhe.Tggsw(n, k, rows)
    assert np.array_equal(g.cmux(ct1, ct2), oracle.cmux(n, k, rows, ct1, ct2))
    if n <= 1024:
        h = oracle.uniform(8, (batch, 2)) % np.uint64(2 * n)
        bsk = np.concatenate([rows, rows])
        assert np.array_equal(fhe.cmux_chain(n, k, [g, g], ct1, h, negacyclic=True), oracle.cmux_chain(n, k, bsk, ct1, h, True))
a, b = oracle.uniform(8, (2, 256)), oracle.uniform(9, (2, 256))
assert np.array_equal(fhe.tn_mul(256, a, b), oracle.tn_mul(256, a, b))
n, k, kn = 64, 1, 64
ksk = oracle.uniform(10, kn * 64 * (kn + 1))
table, cts = oracle.uniform(11, (k + 1) * n), oracle.uniform(12, (130, kn + 1))
K = fhe.Ksk(kn, kn, 64, ksk)
want = oracle.bootstrapping(n, k, ksk, table, cts.reshape(-1), kn).reshape(130, kn + 1)
for path in ("tc", "mma", "cuda"):
    os.environ["FHE_KS_PATH"] = path
    assert np.array_equal(fhe.bootstrap(n, k, K, table, cts, kn), want), path
os.environ.pop("FHE_KS_PATH")
q, n, t = Q, 16, 2
pq = q**3
a, b, rlk = oracle.uniform(13, (9, 2 * n), q), oracle.uniform(14, (9, 2 * n), q), oracle.uniform(15, 2 * n, pq)
assert np.array_equal(fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b).reshape(-1), oracle.bfv_mul(q, n, t, pq, rlk, a.reshape(-1), b.reshape(-1)))
k, l = 2, 8
glwe = (k + 1) * 16
kk = oracle.uniform(16, k * l * glwe, q)
cc = oracle.uniform(17, (3, glwe), q)
G = fhe.RqGlev(fhe.NttPlan(q, 16), k, k * l, kk)
assert np.array_equal(G.key_switch(2, l, cc), oracle.glwe_rq_key_switch(q, 16, k, 2, l, kk, cc))
print("sanitize workload ok, launches:", fhe.launch_count())
