import sys, os, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite import device as dev
from b200_insite.config import default_config
from b200_insite.dataset import SyntheticCancerDatasetCollection
from b200_insite.sindy import SINDY, _JOINT_TO_PER_TREATMENT
col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 1000, 'val': 100, 'test': 100}, seed=10, treatment_mode='multilabel')
col.process_data_multi()
m = SINDY(default_config(insite=True, seed=10, treatment_mode='multilabel', joint_model=True), col); m.fit(col.train_f, col.val_f)
ds = col.test_cf_one_step
prev, static, codes, seq = m._unscaled_inputs(ds)
x = dev.to_device(prev); cd = dev.to_device(codes, dtype=torch.uint8); st = dev.to_device(static)
theta0 = dev.to_device(m.joint_coefs)
for max_iter in (200, 1, 2, 3, 5, 10):
    c11, status, fval = dev.insite_bfgs(x, cd, dev.to_device(seq, dtype=torch.int32), 1, st, theta0, lam=10.0, joint=True, max_iter=max_iter)
    stn = status.cpu().numpy(); low = stn & 255
    for policy in ("keep", "fallback3"):
        c = c11.clone()
        if policy == "fallback3":
            c[torch.from_numpy(low == 3).cuda()] = theta0.reshape(-1)
        c44 = torch.matmul(c, dev.to_device(_JOINT_TO_PER_TREATMENT).T).reshape(-1, 4, 4).contiguous()
        un = m._rollout(prev, static, codes, c44, -1.0)
        sp = ds.scaling_params
        err = un[..., None] - ds.data['unscaled_outputs']
        act = ds.data['active_entries']
        rmse_all = np.sqrt(((err ** 2) * act).sum() / act.sum()) / ds.norm_const * 100
        print(max_iter, policy, "rmse_all", rmse_all, "status hist", np.bincount(low[stn >= 0], minlength=7).tolist(), "iters", float((stn[stn >= 0] >> 8).mean()))
