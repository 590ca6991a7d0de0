"""Compact NTT / INTT / polymul sweep (fraction of the measured HBM peak) for A/B runs of library variants:
FHE_B200_LIB=fhe_study_b200/variants/lib_x.so python tools/ntt_ab.py [logn ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_extras
import fhe_study_b200 as fhe

torch.cuda.set_device(0)
fhe.use_torch_stream()
dev = torch.device("cuda", 0)
logns = [int(x) for x in sys.argv[1:]] or [10, 11, 12, 13, 14]
peak = bench_extras._hbm_peak()
print("lib", os.environ.get("FHE_B200_LIB", "default"), "loge", os.environ.get("FHE_NTT_LOGE", "default"))
for q in (bench_extras.Q17, bench_extras.Q62):
    for logn in logns:
        n = 1 << logn
        batch = (1 << 26) // n
        plan = fhe.NttPlan(q, n)
        a = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device=dev)
        b = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device=dev)
        c = torch.empty_like(a)
        row = []
        for fn, nbuf in ((lambda: plan.ntt(a, out=c), 2), (lambda: plan.intt(a, out=c), 2), (lambda: plan.mul(a, b, out=c), 3)):
            ms = bench_extras._time(fn, 10)
            row.append(nbuf * n * 8 * batch / (ms * 1e-3) / 1e9 / peak)
        print("q=%d n=%5d  ntt %.3f  intt %.3f  polymul %.3f  (%.2f M polymul/s)" % (q, n, row[0], row[1], row[2], batch / ms / 1e3), flush=True)
        del a, b, c
