import sys, os, warnings
sys.path.insert(0, '/root/repo'); warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
import b200_insite.cancer_simulation as cs
n, T = 20000, 60
np.random.seed(4)
params = cs.generate_params(n, 2.0, 2.0, 15, 0)
block = torch.from_numpy(dev.pack_params(params)).cuda()
static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64)).cuda()
vol, codes, sl, pm, _ = dev.sim_factual_rng(block, T, seed=9, pitch=T)
x = vol.contiguous(); cd = codes[:, :T].contiguous(); fit_len = sl.to(torch.int32)
st = dev.theta_gram_codes(vol, codes, sl, static, pm)
prior, _ = dev.stlsq_population(st)
x32 = x.to(torch.float32).contiguous()
for lam in (10.0, 1e4):
    pc = dev.stlsq_batched(x, cd, fit_len, static, prior, lam)
    pc32 = dev.stlsq_batched(x32, cd, fit_len, static, prior, lam)
    scale = pc.abs().amax(dim=(1, 2), keepdim=True)
    print('lam', lam, 'fit dev', float(((pc32 - pc).abs() / scale).max()), 'equal', bool(torch.equal(pc, pc32)), 'x32==x', bool(torch.equal(x32.double(), x)))
x0 = x[:, 0].contiguous(); cd1 = cd[:, :T - 1].contiguous()
p = dev.ode_rollout(x0, static, cd1, pc, drop_below=-1.0)
p32 = dev.ode_rollout(x0, static, cd1, pc, drop_below=-1.0, fp32=True)
print('roll dev', float(((p32 - p).abs() / p.abs().clamp_min(1e-3 * float(p.abs().max()))).max()), torch.isnan(p).any().item(), p.abs().max().item())
