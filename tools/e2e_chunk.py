"""End-to-end packed-wire polymul (pinned host buffers, fhe_rq_mul_packed) at the stage size in FHE_PIPE_CHUNK_MB:
  for mb in 4 8 16 32; do FHE_PIPE_CHUNK_MB=$mb python tools/e2e_chunk.py; done"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import fhe_study_b200 as fhe

torch.cuda.set_device(0)
fhe.use_torch_stream()
q, n, batch, bits = 65537, 1024, 65536, 17
plan = fhe.NttPlan(q, n)
pw = n // 32 * bits
g = torch.Generator().manual_seed(1)
ha = torch.randint(-(2**31), 2**31 - 1, (batch, pw), dtype=torch.int32, generator=g).pin_memory()
hb = torch.randint(-(2**31), 2**31 - 1, (batch, pw), dtype=torch.int32, generator=g).pin_memory()
hc = torch.empty((batch, pw), dtype=torch.int32).pin_memory()
for wire, call in (("packed17", lambda: plan.mul_packed(bits, ha, hb, out=hc)),):
    for rep in range(2):
        call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            call()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("chunk_mb=%s %s: %.3f ms per 65536 polymuls = %.2f M polymul/s (H2D %.1f GB/s)" % (
            os.environ.get("FHE_PIPE_CHUNK_MB", "default"), wire, ms, batch / ms / 1e3, 2 * batch * pw * 4 / ms / 1e6), flush=True)
