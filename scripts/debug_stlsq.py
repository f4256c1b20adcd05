import sys, os, warnings; warnings.filterwarnings('ignore')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch, helpers as h
from oracle import sim_oracle as so
from b200_insite import device as dev
inputs = h.collection_inputs(1, 2.0, 1000, 100, 100)
tr = so.sim_factual(inputs['train'][0], 60, inputs['train'][1])
stats = dev.theta_gram(dev.to_device(tr['cancer_volume']), dev.to_device(tr['chemo_application']),
                       dev.to_device(tr['radio_application']), dev.to_device(tr['sequence_lengths']),
                       dev.to_device(np.asarray(tr['patient_types'], dtype=np.float64)))
torch.cuda.synchronize()
np.set_printoptions(precision=12, linewidth=200)
print(stats.cpu().numpy())
for thr, alpha in [(1e-3, 0.5), (0.08, 0.5), (0.6, 0.05), (5.0, 0.5)]:
    c, s = dev.stlsq_population(stats, thr, alpha)
    torch.cuda.synchronize()
    print(thr, alpha); print(c.cpu().numpy()); print(s.cpu().numpy())
