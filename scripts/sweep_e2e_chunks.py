"""End-to-end time of GeneratedFitPipeline.step_host (host parameters -> result) against the number of H2D chunks."""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite.cohort import GeneratedFitPipeline
import b200_insite.cancer_simulation as cs
n, T = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, 60
dev.require_cuda()
np.random.seed(0)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).pin_memory()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).pin_memory()
res = torch.empty(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory()
uniform = dev.uniform_param_rows(params) if os.environ.get('UNIFORM', '1') == '1' else None
for chunks in (1, 2, 4, 8, 16, 32):
    pipe = GeneratedFitPipeline(n, T, seed=1, chunks=chunks)
    for _ in range(3):
        pipe.step_host(block, static, res, uniform=uniform)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        pipe.step_host(block, static, res, uniform=uniform)
    torch.cuda.synchronize()
    ms = 1e3 * (time.perf_counter() - t0) / 10
    print(f"chunks {len(pipe.bounds):3d}: {ms:.3f} ms per step, coef[0,0] {res[0].item():.6f}", flush=True)
    del pipe
