"""A/B timing of the fused external product / CMux chain for library variants (n = 1024, k = 1):
FHE_B200_LIB=fhe_study_b200/variants/lib_x.so python tools/xp_ab.py [batch] [steps]
Prints throughput and a checksum of the outputs (same seeded inputs in every run: equal checksums = identical results)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_extras
import fhe_study_b200 as fhe

torch.cuda.set_device(0)
fhe.use_torch_stream()
dev = torch.device("cuda", 0)
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 32
n, k = int(os.environ.get("XP_N", 1024)), int(os.environ.get("XP_K", 1))
glwe = (k + 1) * n
g = torch.Generator(device="cuda").manual_seed(7)
r = lambda *shape: torch.randint(-(2**63), 2**63 - 1, shape, dtype=torch.int64, device="cuda", generator=g)
G = fhe.Tggsw(n, k, r((k + 1) * 64 * glwe))
ct1, ct2 = r(batch, glwe), r(batch, glwe)
out = torch.empty_like(ct1)
ms_x = bench_extras._time(lambda: G.extprod(ct1, out=out), 10, warm=3)
sum_x = int(out.sum().item())
ms_c = bench_extras._time(lambda: G.cmux(ct1, ct2, out=out), 10, warm=3)
sum_c = int(out.sum().item())
print("lib %s  extprod %.3f M/s  cmux %.3f M/s  checksums %d %d" % (os.environ.get("FHE_B200_LIB", "default"), batch / ms_x / 1e3,
                                                                  batch / ms_c / 1e3, sum_x, sum_c), flush=True)
if steps:
    cb = 592 if n == 1024 else 4096
    gs = [fhe.Tggsw(n, k, r((k + 1) * 64 * glwe)) for _ in range(steps)]
    acc = r(cb, glwe)
    h = torch.randint(0, 2 * n, (cb, steps), dtype=torch.int64, device="cuda", generator=g)
    o = torch.empty_like(acc)
    ms = bench_extras._time(lambda: fhe.cmux_chain(n, k, gs, acc, h, negacyclic=True, out=o), 5, warm=2)
    print("   chain %d steps x %d: %.3f M CMux/s  checksum %d" % (steps, cb, cb * steps / ms / 1e3, int(o.sum().item())), flush=True)
