"""Runs the simulate_factual kernel a few times on a synthetic cohort (profiling target).
Usage: python scripts/run_k1.py N variant fused reps"""
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import torch
from b200_insite import device as dev
sys.path.insert(0, ROOT)
from bench import synth_inputs

n, variant, fused, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
T = 60
dev.require_cuda()
params, block, static, draws = synth_inputs(n, T, 0)
block, static = block.cuda(), static.cuda()
pitch = int(os.environ.get('PITCH', str(T)))
def _rows(src):
    t = dev.alloc_rows(n, T, pitch); t.copy_(src); return t
draws = [_rows(d) for d in draws]
out = {k: dev.alloc_rows(n, T, pitch) for k in dev.FACTUAL_OUT_KEYS}
out['sequence_lengths'] = torch.empty((n,), dtype=torch.float64, device='cuda')
_codes = torch.zeros((n, 64), dtype=torch.uint8, device='cuda'); _pm = torch.empty((6, n), dtype=torch.float64, device='cuda')
ts = []
for i in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if os.environ.get('SIDE'):
        dev.sim_factual_side(block, *draws, T, out=out, codes=_codes, patient_moments=_pm, variant=variant)
    else:
        dev.sim_factual(block, *draws, T, out=out, variant=variant, fused_static=static if fused else None)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print("ms per launch:", ["%.3f" % t for t in ts], "mean seq len", out['sequence_lengths'].mean().item())
