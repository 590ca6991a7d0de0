"""A/B tuning sweep of the NTT kernels.  Each configuration runs in its own process (plans read their knobs when
they are created):   python tools/ntt_tune.py "q17:10,11:FHE_NTT_DUAL=1" "q62:10:FHE_NTT_LOGE=3" ...
spec = <q17|q62|q31>:<logn,logn,...>:<ENV=V,ENV=V,...>[:u32]   (u32: time the packed 32-bit device format as well)
Every configuration is first checked bit-for-bit against the oracle on a small batch."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QS = {"q17": 65537, "q62": 0x3FFFFFFFFFFF0001, "q31": 2013265921}


def child(qname, logn, u32):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch

    import bench_extras
    import fhe_study_b200 as fhe
    import oracle

    torch.cuda.set_device(0)
    fhe.use_torch_stream()
    dev = torch.device("cuda", 0)
    q, n = QS[qname], 1 << logn
    plan = fhe.NttPlan(q, n)
    # ---- parity on a ragged small batch -------------------------------------------------------------
    nb = 37
    a, b = oracle.uniform(11, (nb, n), q), oracle.uniform(12, (nb, n), q)
    want = oracle.rq_mul_batch(q, n, a, b)
    ok = bool((plan.mul(a, b) == want).all()) and bool((plan.ntt(a) == oracle.ntt(q, n, a)).all())
    ok = ok and bool((plan.intt(plan.ntt(a)) == a).all())
    ev = np.empty_like(a)
    ok = ok and bool((plan.mul(a, b, evals_out=ev) == want).all()) and bool((plan.mul(plan.ntt(a), b, flags=1) == want).all())
    if q <= 2**32 and not os.environ.get("FHE_NTT_LOGE"):  # the alternative shapes exist for u64 words only
        a32, b32 = a.astype(np.uint32), b.astype(np.uint32)
        ok32 = bool((plan.mul_u32(a32, b32).astype(np.uint64) == want).all())
        ok32 = ok32 and bool((plan.ntt_u32(a32).astype(np.uint64) == oracle.ntt(q, n, a)).all())
        ok32 = ok32 and bool((plan.intt_u32(plan.ntt_u32(a32)) == a32).all())
    else:
        ok32 = None
    # ---- throughput -----------------------------------------------------------------------------------
    peak = bench_extras._hbm_peak()
    batch = (1 << 26) // n
    out = []
    for dt, wb in ((torch.int64, 8),) + (((torch.int32, 4),) if (u32 and q <= 2**32) else ()):
        A = torch.randint(0, min(q, 2**31 - 1 if wb == 4 else 2**62), (batch, n), dtype=dt, device=dev)
        B = torch.randint(0, min(q, 2**31 - 1 if wb == 4 else 2**62), (batch, n), dtype=dt, device=dev)
        C = torch.empty_like(A)
        fns = ((lambda: plan.ntt(A, out=C), 2), (lambda: plan.intt(A, out=C), 2), (lambda: plan.mul(A, B, out=C), 3)) if wb == 8 else \
              ((lambda: plan.ntt_u32(A, out=C), 2), (lambda: plan.intt_u32(A, out=C), 2), (lambda: plan.mul_u32(A, B, out=C), 3))
        row = []
        for fn, nbuf in fns:
            ms = bench_extras._time(fn, 10)
            row.append((batch / ms / 1e3, nbuf * n * wb * batch / (ms * 1e-3) / 1e9 / peak))
        out.append("u%d: ntt %.1f M/s (%.3f hbm) intt %.1f (%.3f) polymul %.2f M/s (%.3f)" % (wb * 8, *row[0], *row[1], *row[2]))
        del A, B, C
    print("%s n=%5d ok=%s ok32=%s | %s" % (qname, n, ok, ok32, " | ".join(out)), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), sys.argv[4] == "1")
        sys.exit(0)
    for spec in sys.argv[1:]:
        parts = spec.split(":")
        qname, logns = parts[0], [int(x) for x in parts[1].split(",")]
        envs = dict(kv.split("=") for kv in parts[2].split(",") if kv) if len(parts) > 2 else {}
        u32 = len(parts) > 3 and parts[3] == "u32"
        for logn in logns:
            env = dict(os.environ, **envs)
            sys.stdout.write("[%s] " % ",".join(f"{k}={v}" for k, v in envs.items()))
            sys.stdout.flush()
            r = subprocess.run([sys.executable, __file__, "--child", qname, str(logn), "1" if u32 else "0"], env=env,
                               capture_output=True, text=True, timeout=600)
            sys.stdout.write(r.stdout if r.returncode == 0 else "FAILED rc=%d %s\n" % (r.returncode, r.stderr[-800:]))
            sys.stdout.flush()
