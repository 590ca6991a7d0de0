"""Tiny driver for ncu captures: runs one hot-path op a few times on cuda:0 (device-resident buffers).
  python tools/prof.py polymul|ntt|intt [logn] [q] [batch] [reps] | tn [n] [batch] | extprod [n k batch] | bootstrap [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fhe_study_b200 as fhe

op = sys.argv[1] if len(sys.argv) > 1 else "polymul"
logn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
q = int(sys.argv[3], 0) if len(sys.argv) > 3 else 65537
n = 1 << logn
batch = int(sys.argv[4]) if len(sys.argv) > 4 else (1 << 26) // n
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
torch.cuda.set_device(0)
fhe.use_torch_stream()
if op == "tn":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randint(-(2**63), 2**63 - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    b = torch.randint(-(2**63), 2**63 - 1, (batch, n), dtype=torch.int64, device="cuda", generator=g)
    c = torch.empty_like(a)
    for _ in range(3):
        fhe.tn_mul(n, a, b, out=c)
    torch.cuda.synchronize()
    print("done tn", n, batch)
    sys.exit(0)
if op == "bfv":
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
    q, n, t = 65537, 16, 2
    pq = q * q * q
    a = torch.randint(0, q, (batch, 2 * n), dtype=torch.int64, device="cuda")
    b = torch.randint(0, q, (batch, 2 * n), dtype=torch.int64, device="cuda")
    rlk = torch.randint(0, pq, (2 * n,), dtype=torch.int64, device="cuda")
    out = torch.empty_like(a)
    for _ in range(3):
        fhe.bfv_mul_relin(q, n, t, pq, rlk, a, b, out=out)
    torch.cuda.synchronize()
    print("done bfv", batch)
    sys.exit(0)
if op in ("bootstrap", "extprod"):
    g = torch.Generator(device="cuda").manual_seed(1)
    r = lambda *shape: torch.randint(-(2**63), 2**63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
    if op == "bootstrap":
        kn = 1024
        batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
        K = fhe.Ksk(kn, kn, 64, r(kn * 64 * (kn + 1)))
        table, cts = r(2 * 1024), r(batch, kn + 1)
        out = torch.empty_like(cts)
        for _ in range(3):
            fhe.bootstrap(1024, 1, K, table, cts, kn, out=out)
    else:
        n, k = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1024, 1)
        batch = int(sys.argv[4]) if len(sys.argv) > 4 else 512
        glwe = (k + 1) * n
        G = fhe.Tggsw(n, k, r((k + 1) * 64 * glwe))
        ct = r(batch, glwe)
        out = torch.empty_like(ct)
        for _ in range(3):
            G.extprod(ct, out=out)
    torch.cuda.synchronize()
    print("done", op)
    sys.exit(0)
plan = fhe.NttPlan(q, n)
if op == "polymul32":  # the packed 32-bit device format
    a = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda")
    b = torch.randint(0, q, (batch, n), dtype=torch.int32, device="cuda")
    c = torch.empty_like(a)
    for _ in range(reps):
        plan.mul_u32(a, b, out=c)
    torch.cuda.synchronize()
    print("done", op, n, q, batch, reps)
    sys.exit(0)
a = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device="cuda")
b = torch.randint(0, min(q, 2**62), (batch, n), dtype=torch.int64, device="cuda")
c = torch.empty_like(a)
for _ in range(reps):
    if op == "polymul":
        plan.mul(a, b, out=c)
    elif op == "ntt":
        plan.ntt(a, out=c)
    else:
        plan.intt(a, out=c)
torch.cuda.synchronize()
print("done", op, n, q, batch, reps)
