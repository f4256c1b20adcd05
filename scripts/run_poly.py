"""Times the degree-4 library kernels (csrc/poly_library.cu) on a generated factual cohort.
Usage: python scripts/run_poly.py N reps"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n, reps = int(sys.argv[1]), int(sys.argv[2])
T = 60
dev.require_cuda()
np.random.seed(4)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol, codes, sl, pm, _ = dev.sim_factual_rng(block, T, seed=9, pitch=T)
x = vol.contiguous(); cd = codes[:, :T].contiguous()
chemo = (cd & 1).to(torch.float64); radio = ((cd >> 1) & 1).to(torch.float64)
seq = sl.to(torch.float64)
def t(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b), r
for i in range(reps):
    ms_q, rf = t(lambda: dev.poly_tsqr(x, chemo, radio, seq, static))
    ms_s, (coefs, sup) = t(lambda: dev.poly_stlsq(rf))
    ms_r, pred = t(lambda: dev.poly_rollout(x[:, 0].contiguous(), static, cd[:, :T - 1].contiguous(), coefs))
    ms_g, st = t(lambda: dev.theta_gram(x, chemo, radio, seq, static))
    rows = float(rf[4 * 256:].sum())
    print(f"poly_tsqr {ms_q:.3f} ms ({rows / ms_q / 1e6:.2f} G sample rows/s)  poly_stlsq {ms_s:.3f} ms  "
          f"poly_rollout {ms_r:.3f} ms  [4-term theta_gram gen-1 {ms_g:.3f} ms]  support {sup.sum(1).tolist()}", flush=True)
