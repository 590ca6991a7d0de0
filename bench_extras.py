"""Secondary measurements printed under `extras` in bench.py's JSON line (device-resident, CUDA events):
the NTT / INTT / polymul sweep of BASELINE configs[1] and -- once built -- the TFHE and BFV paths of
configs[2..4].  Each entry carries its own roofline fraction against the measured HBM peak."""
from __future__ import annotations

import json
import os

ROOT = os.path.dirname(os.path.abspath(__file__))
Q17 = 65537
Q62 = 0x3FFFFFFFFFFF0001


def _hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def _time(fn, reps, warm=3):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps  # ms


def ntt_sweep(fhe, dev, quick):
    import torch

    peak = _hbm_peak()
    out = []
    reps = 3 if quick else 20
    for q in (Q17, Q62):
        for logn in (10, 11, 12, 13, 14):
            n = 1 << logn
            batch = (1 << 26) // n  # 512 MiB per operand
            plan = fhe.NttPlan(q, n)
            a = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device=dev)
            b = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device=dev)
            c = torch.empty_like(a)
            row = {"q": q, "n": n, "batch": batch}
            for name, fn, nbuf in (
                ("ntt", lambda: plan.ntt(a, out=c), 2),
                ("intt", lambda: plan.intt(a, out=c), 2),
                ("polymul", lambda: plan.mul(a, b, out=c), 3),
            ):
                ms = _time(fn, reps)
                gbs = nbuf * n * 8 * batch / (ms * 1e-3) / 1e9
                row[name] = {"per_s": batch / (ms * 1e-3), "ms": ms, "hbm_gbs": gbs, "hbm_frac": gbs / peak}
            out.append(row)
            del a, b, c
    # batch-1 latency (BASELINE configs[0]: the crate's own test path), device-resident
    plan = fhe.NttPlan(Q17, 1024)
    a = torch.randint(0, Q17, (1, 1024), dtype=torch.int64, device=dev)
    c = torch.empty_like(a)
    lat = _time(lambda: plan.mul(a, a, out=c), 200, warm=20)
    return {"sweep": out, "polymul_n1024_batch1_us": lat * 1e3}


def run(fhe, dev, quick=False):
    res = {"ntt": ntt_sweep(fhe, dev, quick)}
    return res
