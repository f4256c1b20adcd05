"""Tuning sweep: times every tile shape of the TMA-tiled simulate_factual kernel (and the fused-gram
variant) at N patients with CUDA events.  Usage: python scripts/sweep_k1.py [N] [reps]"""
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

warnings.filterwarnings('ignore')
from b200_insite import device as dev


VARIANTS = [1, 2, 10, 12]
FUSED = [bool(int(v)) for v in os.environ.get('FUSED', '0').split(',')]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    T = int(os.environ.get('T', '60'))
    dev.require_cuda()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    # synthetic cohort with the parameter distribution's typical magnitudes (timing only)
    params = torch.empty((10, n), dtype=torch.float64, device='cuda')
    params[0] = torch.exp(torch.randn(n, generator=g, device='cuda', dtype=torch.float64) * 1.5 + 2.0).clamp(0.02, 1000)
    params[1] = (0.0398 + 0.168 * torch.randn(n, generator=g, device='cuda', dtype=torch.float64)).abs() + 1e-3
    params[2] = (7e-5 + 7.23e-3 * torch.randn(n, generator=g, device='cuda', dtype=torch.float64)).abs() + 1e-5
    params[3] = params[1] / 10
    params[4] = 0.028
    params[5] = 14137.166941154068
    params[6] = 6.499999999999999; params[7] = 6.499999999999999
    params[8] = 2.0 / 12.999999999999998; params[9] = 2.0 / 12.999999999999998
    pitch = int(os.environ.get('PITCH', str(T)))
    def rows(src):
        t = dev.alloc_rows(n, T, pitch); t.copy_(src); return t
    noise = rows(0.01 * torch.randn((n, T), generator=g, device='cuda', dtype=torch.float64))
    rec, chemo, radio = (rows(torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64)) for _ in range(3))
    static = torch.randint(1, 4, (n,), generator=g, device='cuda').double()
    out = {k: dev.alloc_rows(n, T, pitch) for k in dev.FACTUAL_OUT_KEYS}
    out['sequence_lengths'] = torch.empty((n,), dtype=torch.float64, device='cuda')
    bytes_alg = n * (4 * T * 8 + 9 * T * 8 + 10 * 8 + 8)
    res = []
    for fused in FUSED:
        for variant in VARIANTS:
            try:
                for _ in range(2):
                    dev.sim_factual(params, noise, rec, chemo, radio, T, out=out, variant=variant,
                                    fused_static=static if fused else None)
                torch.cuda.synchronize()
                ts = []
                for _ in range(reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    dev.sim_factual(params, noise, rec, chemo, radio, T, out=out, variant=variant,
                                    fused_static=static if fused else None)
                    e1.record(); torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = float(np.median(ts))
                r = dict(variant=variant, fused=fused, ms=ms, gbs=bytes_alg / ms / 1e6,
                         mean_len=float(out['sequence_lengths'].mean().item()))
            except RuntimeError as ex:
                r = dict(variant=variant, fused=fused, error=str(ex)[:200])
            print(json.dumps(r), flush=True)
            res.append(r)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'sweep_k1.json'), 'w') as f:
        json.dump(res, f, indent=1)


if __name__ == '__main__':
    main()
