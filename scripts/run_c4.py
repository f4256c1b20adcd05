"""Runs the individualisation kernels of config C4 on a generated factual cohort (profiling target):
K5b stlsq_batched, K6 ode_rollout (per-row coefficients), optionally with per-row interval lengths.
Usage: python scripts/run_c4.py N reps [irregular]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n, reps = int(sys.argv[1]), int(sys.argv[2])
irregular = len(sys.argv) > 3
T = 60
dev.require_cuda()
np.random.seed(4)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol, codes, sl, pm, _ = dev.sim_factual_rng(block, T, seed=9, pitch=T)
x = vol.contiguous(); cd = codes[:, :T].contiguous(); fit_len = sl.to(torch.int32)
st = dev.theta_gram_codes(vol, codes, sl, static, pm)
prior, _ = dev.stlsq_population(st)
g = torch.Generator(device='cuda'); g.manual_seed(1)
dts = (dev.STANDARD_DT * (0.3 + 2.2 * torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64))) if irregular else None
x32 = x.to(torch.float32).contiguous()
def t(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), r
for i in range(reps):
    ms5, pc = t(lambda: dev.stlsq_batched(x, cd, fit_len, static, prior, 10.0, dts=dts))
    ms5f, pcf = t(lambda: dev.stlsq_batched(x32, cd, fit_len, static, prior, 10.0, dts=dts))
    x0 = x[:, 0].contiguous(); cd1 = cd[:, :T - 1].contiguous(); d1 = None if dts is None else dts[:, :T - 1].contiguous()
    ms6, pred = t(lambda: dev.ode_rollout(x0, static, cd1, pc, drop_below=-1.0, dts=d1))
    ms6f, pred32 = t(lambda: dev.ode_rollout(x0, static, cd1, pc, drop_below=-1.0, dts=d1, fp32=True))
    ms6p, _ = t(lambda: dev.ode_rollout(x0, static, cd1, prior, dts=d1))
    print(f"K5b {ms5:.3f} ms  K5b(f32 storage) {ms5f:.3f} ms  K6 {ms6:.3f} ms  K6 f32 {ms6f:.3f} ms  K6 shared coefs {ms6p:.3f} ms", flush=True)
