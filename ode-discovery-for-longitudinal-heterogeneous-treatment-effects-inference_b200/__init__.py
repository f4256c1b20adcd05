"""b200_insite -- B200-native (sm_100a) implementation of INSITE's data-parallel hot path:
the CT cancer PK-PD simulator and the SINDy fit-and-predict loop.

Host layer (this package, Python) mirrors the reference interfaces:
    cancer_simulation   generate_params / simulate_factual / simulate_counterfactual_1_step /
                        simulate_counterfactuals_treatment_seq / get_scaling_params
    dataset             SyntheticCancerDataset / SyntheticCancerDatasetCollection
    sindy               SINDY (fit / get_predictions / get_autoregressive_predictions / RMSE metrics)
Device layer:
    device              typed wrappers over the C ABI (include/b200i.h) of csrc/libb200insite.so
    cohort              device-resident 1M+ patient pipeline, sharded over GPUs (distributed)
Import `b200_insite` (alias directory at the repository root).
"""
__version__ = "0.1.0"
