"""Key switch through the tensor-core engine against the oracle on a few rows (n=1024, k=1, l=64); used while tuning
the kernel (FHE_KS_CLUSTER=1|2).  python tools/ks_check.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import fhe_study_b200 as fhe
import oracle

torch.cuda.set_device(0)
kn, l = 1024, 64
ksk = oracle.uniform(4, kn * l * (kn + 1))
K = fhe.Ksk(kn, kn, l, ksk)
for batch in (300, 512, 1000):
    cts = oracle.uniform(6, (batch, kn + 1))
    got = K.key_switch(cts)
    rows = [0, 1, 255, 256, batch - 1]
    want = oracle.key_switch(kn, kn, l, ksk, cts[rows].copy(), threads=8).reshape(len(rows), kn + 1)
    print(batch, "ok" if (got[rows] == want).all() else "MISMATCH", flush=True)
