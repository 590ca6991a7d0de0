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
def boot():
    """bootstrapping as executed + key switch alone, n=1024, k=1, l=64, batches 8192 and 65536"""
    import torch

    n, k, kn, l = 1024, 1, 1024, 64
    ksk = bench_extras._u64_rand(torch, (kn * l * (kn + 1),), dev, 4)
    K = fhe.Ksk(kn, kn, l, ksk)
    del ksk
    table = bench_extras._u64_rand(torch, ((k + 1) * n,), dev, 5)
    res = {}
    for batch in (8192, 65536):
        cts = bench_extras._u64_rand(torch, (batch, kn + 1), dev, 6)
        out = torch.empty_like(cts)
        ms_b = bench_extras._time(lambda: fhe.bootstrap(n, k, K, table, cts, kn, out=out), 5, warm=2)
        ms_k = bench_extras._time(lambda: K.key_switch(cts, out=out), 5, warm=2)
        res[batch] = {"bootstraps_per_s": batch / ms_b * 1e3, "key_switch_per_s": batch / ms_k * 1e3,
                      "int8_pops": batch / ms_b * 1e3 * kn * l * (kn + 1) * 16 / 1e15}
        del cts, out
    return res


fn = {"boot": boot, "tfhe": lambda: bench_extras.tfhe_paths(fhe, dev, False, cpu=False), "bfv": lambda: bench_extras.bfv_path(fhe, dev, False, cpu=False),
      "tn": lambda: bench_extras.tn_mul_path(fhe, dev, False), "ntt": lambda: bench_extras.ntt_sweep(fhe, dev, False),
      "gfhe": lambda: bench_extras.gfhe_path(fhe, dev, False, cpu=True)}[what]
print(json.dumps(fn(), indent=1))
