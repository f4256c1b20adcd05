import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite.config import default_config
from b200_insite.dataset import SyntheticCancerDatasetCollection
from b200_insite.sindy import SINDY
col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 1000, 'val': 100, 'test': 100}, seed=1)
col.process_data_multi()
m = SINDY(default_config(insite=True), col); m.fit(col.train_f, col.val_f)
ds = col.test_cf_one_step
prev, static, codes, seq = m._unscaled_inputs(ds)
x = dev.to_device(prev); cd = dev.to_device(codes, dtype=torch.uint8); st = dev.to_device(static)
theta0 = dev.to_device(m.joint_coefs)
c16, status, fval = dev.insite_bfgs(x, cd, dev.to_device(seq, dtype=torch.int32), 1, st, theta0, lam=10.0)
stn = status.cpu().numpy(); low = stn & 255
for policy in ("keep", "fallback3"):
    c = c16.clone()
    if policy == "fallback3":
        c[torch.from_numpy(low == 3).cuda()] = theta0
    un = m._rollout(prev, static, codes, c.contiguous(), -1.0)
    err = un[..., None] - ds.data['unscaled_outputs']
    act = ds.data['active_entries']
    rmse_all = np.sqrt(((err ** 2) * act).sum() / act.sum()) / ds.norm_const * 100
    print(policy, "rmse_all", rmse_all, "log 1.0837636799825472", "status hist", np.bincount(low[stn >= 0], minlength=7).tolist())
