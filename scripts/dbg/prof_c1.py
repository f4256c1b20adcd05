import sys, cProfile, pstats, warnings
sys.path.insert(0, '/root/repo'); warnings.filterwarnings('ignore')
import numpy as np, torch
from b200_insite.config import default_config
from b200_insite.dataset import SyntheticCancerDatasetCollection
from b200_insite.sindy import SINDY
col = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 10000, 'val': 1000, 'test': 1000}, seed=1)
col.process_data_multi()
m = SINDY(default_config(insite=True, n_train=10000, n_val=1000, n_test=1000), col); m.fit(col.train_f, col.val_f)
m.get_normalised_masked_rmse(col.test_cf_one_step, one_step_counterfactual=True)
pr = cProfile.Profile(); pr.enable()
m.get_normalised_masked_rmse(col.test_cf_one_step, one_step_counterfactual=True)
m.get_normalised_n_step_rmses(col.test_cf_treatment_seq)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
pr = cProfile.Profile(); pr.enable()
col2 = SyntheticCancerDatasetCollection(2.0, 2.0, {'train': 10000, 'val': 1000, 'test': 1000}, seed=1)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(14)
