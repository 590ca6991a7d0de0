"""Quick device-resident measurements of one family: python tools/quick.py tfhe|bfv|tn|ntt"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench_extras
import fhe_study_b200 as fhe

torch.cuda.set_device(0)
fhe.use_torch_stream()
dev = torch.device("cuda", 0)
what = sys.argv[1] if len(sys.argv) > 1 else "tfhe"
fn = {"tfhe": lambda: bench_extras.tfhe_paths(fhe, dev, False, cpu=False), "bfv": lambda: bench_extras.bfv_path(fhe, dev, False, cpu=False),
      "tn": lambda: bench_extras.tn_mul_path(fhe, dev, False), "ntt": lambda: bench_extras.ntt_sweep(fhe, dev, False),
      "gfhe": lambda: bench_extras.gfhe_path(fhe, dev, False, cpu=True)}[what]
print(json.dumps(fn(), indent=1))
