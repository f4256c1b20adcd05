"""Times theta_gram (K4) alone on the outputs of one simulate_factual launch.  Usage: python scripts/time_k4.py [N]"""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from bench import synth_inputs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
T = 60
dev.require_cuda()
params, block, static, draws = synth_inputs(n, T, 0)
block, static = block.cuda(), static.cuda()
pitch = int(os.environ.get('PITCH', str(T)))
def _rows(src):
    t = dev.alloc_rows(n, T, pitch); t.copy_(src); return t
draws = [_rows(d) for d in draws]
out, _ = dev.sim_factual(block, *draws, T)
del draws
ts = []
for i in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'], out['sequence_lengths'],
                        static, out['chemo_dosage'], out['radio_dosage'])
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[2:]))
byt = n * (5 * T * 8 + 16)
print(f"theta_gram: {ms:.3f} ms  {byt / ms / 1e6:.0f} GB/s over 5 arrays ({byt/1e9:.2f} GB)", ["%.3f" % t for t in ts])
print("stats head", st[:6].cpu().numpy())
