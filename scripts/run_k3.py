"""Runs the treatment-sequence counterfactual kernel (K3) on a synthetic cohort (profiling target).
Usage: python scripts/run_k3.py N reps"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite import counterfactual as cfm
import b200_insite.cancer_simulation as cs
n, reps = int(sys.argv[1]), int(sys.argv[2])
T, H = 60, 5
dev.require_cuda()
g = torch.Generator(device='cuda'); g.manual_seed(7)
np.random.seed(3)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
noise = 0.01 * torch.randn((n, T + H), generator=g, device='cuda', dtype=torch.float64)
rec, chemo, radio = (torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64) for _ in range(3))
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); coh = cfm.sim_cf_treatment_seq(block, noise, rec, chemo, radio, T, H); e1.record(); torch.cuda.synchronize()
    print("ms", e0.elapsed_time(e1), "rows", coh.total_rows, "levels", coh.levels)
    del coh
