"""Experiment: resident generated-draws step with K1L chunks on one stream and the chunks' statistics (K4 lean) on a
second one, against the sequential two-launch step."""
import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n, T = 1_000_000, 60
dev.require_cuda()
np.random.seed(0)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol = dev.alloc_rows(n, T, 64); codes = torch.empty((n, 64), dtype=torch.uint8, device='cuda')
sl = torch.empty((n,), dtype=torch.float64, device='cuda'); pm = torch.empty((6, n), dtype=torch.float64, device='cuda')
def seq_step():
    dev.sim_factual_rng(block, T, 1, volume=vol, codes=codes, sequence_lengths=sl, patient_moments=pm)
    return dev.theta_gram_codes(vol, codes, sl, static, pm)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("sequential", timeit(seq_step))
sB = torch.cuda.Stream()
for C in (2, 4, 8, 16):
    step = -(-n // C); step = -(-step // 32) * 32
    bounds = [(a, min(a + step, n)) for a in range(0, n, step)]
    evs = [torch.cuda.Event() for _ in bounds]
    def ovl_step():
        main = torch.cuda.current_stream()
        sB.wait_stream(main)
        for i, (a, b) in enumerate(bounds):
            dev.sim_factual_rng(block, T, 1, volume=vol, codes=codes, sequence_lengths=sl, patient_moments=pm, rows=(a, b))
            evs[i].record(main)
            with torch.cuda.stream(sB):
                sB.wait_event(evs[i])
                dev.theta_gram_codes(vol[a:b], codes[a:b], sl[a:b], static[a:b], pm[:, a:b], tag=f"c{i}")
        main.wait_stream(sB)
    print("chunks", C, timeit(ovl_step))
