"""Times the INSITE per-row BFGS kernel (K7) on rows of a generated factual cohort.  Usage: python scripts/run_k7.py [rows] [reps]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 60
dev.require_cuda()
np.random.seed(4)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol, codes, sl, pm, _ = dev.sim_factual_rng(block, T, 11)
st = dev.theta_gram_codes(vol, codes, sl, static, pm)
prior, _ = dev.stlsq_population(st)
x = vol.contiguous(); cd = codes[:, :T].contiguous(); fl = sl.to(torch.int32)
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c7, status, fval = dev.insite_bfgs(x, cd, fl, 1, static, prior.contiguous(), 10.0); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"ms {ms:.2f}  fits/s {n / ms * 1e3:.0f}  mean objective ratio {float((fval[:, 1] / fval[:, 0].clamp_min(1e-300)).mean()):.6f}")
