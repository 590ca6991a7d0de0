"""Thread safety of the C ABI (SURVEY 8b "Threading": the reference's operators are re-entrant and `cargo test` calls them
from many threads at once): concurrent host-buffer calls from 8 threads -- plan creation racing on the (q, n) cache,
polymul, Tn product, external product on one shared TGGSW handle, key switch on one shared KSK handle, and a failing
call whose error message must stay with its own thread -- each checked bit for bit against the oracle."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
Q = 65537


def test_concurrent_calls_from_many_threads():
    import fhe_study_b200 as fhe
    import oracle as orc

    fhe.set_device(0)
    n, k = 64, 1
    glwe = (k + 1) * n
    tggsw = orc.uniform(1, (k + 1) * 64 * glwe)
    g = fhe.Tggsw(n, k, tggsw)
    kn = 64
    ksk = orc.uniform(2, kn * 64 * (kn + 1))
    K = fhe.Ksk(kn, kn, 64, ksk)
    errors = []

    def worker(tid):
        try:
            fhe.set_device(0)
            for it in range(6):
                seed = 100 * tid + it
                nn = 1 << (6 + (tid + it) % 5)  # 64..1024: threads race on the plan cache
                plan = fhe.NttPlan(Q, nn)
                a, b = orc.uniform(seed, (3, nn), Q), orc.uniform(seed + 1, (3, nn), Q)
                assert np.array_equal(plan.mul(a, b), orc.rq_mul_batch(Q, nn, a, b)), "polymul"
                x, y = orc.uniform(seed + 2, (2, 128)), orc.uniform(seed + 3, (2, 128))
                assert np.array_equal(fhe.tn_mul(128, x, y), orc.tn_mul(128, x, y)), "tn_mul"
                ct1, ct2 = orc.uniform(seed + 4, (3, glwe)), orc.uniform(seed + 5, (3, glwe))
                assert np.array_equal(g.cmux(ct1, ct2), orc.cmux(n, k, tggsw, ct1, ct2)), "cmux"
                c = orc.uniform(seed + 6, (70, kn + 1))
                assert np.array_equal(K.key_switch(c).reshape(-1), orc.key_switch(kn, kn, 64, ksk, c.reshape(-1))), "key_switch"
                with pytest.raises(RuntimeError) as ei:
                    fhe.NttPlan(Q, 3 + 4 * tid)  # never a power of two: the reference panics (ntt.rs:116-117)
                assert "fhe_ntt_plan_create" in str(ei.value)
        except BaseException as ex:  # noqa: BLE001 -- collected (pytest outcomes included) and re-raised in the main thread
            errors.append((tid, repr(ex)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
