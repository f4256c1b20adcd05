"""Times sim_factual_rng (K1L, device-generated draws) alone.  Usage: python scripts/run_k1_rng.py [N] [reps]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
T = 60
dev.require_cuda()
np.random.seed(0)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol = dev.alloc_rows(n, T, 64)
codes = torch.empty((n, 64), dtype=torch.uint8, device='cuda')
sl = torch.empty((n,), dtype=torch.float64, device='cuda')
pm = torch.empty((6, n), dtype=torch.float64, device='cuda')
for name, kw in (("plain", dict(moments=False)), ("moments", dict(patient_moments=pm)), ("fused", dict(fused_static=static))):
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = dev.sim_factual_rng(block, T, 1234, volume=vol, codes=codes, sequence_lengths=sl,
                                  variant=int(os.environ.get('VARIANT', '0')), **kw)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "ms per launch:", ["%.3f" % t for t in ts], "mean seq len", sl.mean().item(), flush=True)
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = dev.theta_gram_codes(vol, codes, sl, static, pm)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("theta_gram_codes ms:", ["%.3f" % t for t in ts])
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d = dev.philox_draws(n, T, 1234, pitch=64)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("philox_draws ms:", ["%.3f" % t for t in ts])
