"""K2 and K3 generator times at N patients (experiment helper).  Usage: python scripts/dbg/time_k2_k3.py N"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite import counterfactual as cfm
import b200_insite.cancer_simulation as cs
n = int(sys.argv[1]); T, H = 60, 5
dev.require_cuda()
g = torch.Generator(device='cuda'); g.manual_seed(7)
np.random.seed(3)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
noise = 0.01 * torch.randn((n, T + H), generator=g, device='cuda', dtype=torch.float64)
rec, chemo, radio = (torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64) for _ in range(3))
def med(fn, reps=5):
    ts = []
    for _ in range(reps + 1):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)); del r
    return float(np.median(ts[1:]))
k3 = med(lambda: cfm.sim_cf_treatment_seq(block, noise, rec, chemo, radio, T, H))
k2 = med(lambda: cfm.sim_cf_one_step(block, noise[:, :T].contiguous(), rec, chemo, radio, T))
print(f"K3 {k3:.3f} ms  K2 {k2:.3f} ms")
