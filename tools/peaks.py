import sys; sys.path.insert(0,'.')
import torch, fhe_study_b200 as fhe
torch.cuda.set_device(0)
for k in (0,1,2,3,4,3):
    print(k, fhe.int_peak(k)/1e12, flush=True)
