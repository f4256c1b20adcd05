"""TEST INFRASTRUCTURE ONLY.  Golden vectors of the reference's EQ_5 simulators (libs_m/ct/src/data/continuous/
continuous.py), produced by running the UNMODIFIED reference in the build container (oracle/ref_loader.py):
tests/golden/ref_continuous_small.npz.  For EQ_5_A (one patient type, no observation noise) and EQ_5_D (three types,
per-patient beta_c, observation noise): parameters, factual / one-step / treatment-sequence outputs and the state of
the global RNG afterwards (a digest), seq_length 30 to keep the file small."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

T, H = 30, 5
SIZES = dict(factual=40, one=6, seq=5)


def rng_digest():
    return hashlib.sha256(np.random.get_state()[1].tobytes()).hexdigest()


def main():
    ref = ref_loader.load_reference_continuous()
    Eq = sys.modules["src.data.pkpd.pkpd_simulation"].Equation
    out = {}
    for eq in ('EQ_5_A', 'EQ_5_D'):
        np.random.seed(17)
        for kind, n in SIZES.items():
            p = ref.generate_params(n, 2.0, 2.0, 15, 0, Eq[eq])
            if kind == 'factual':
                sim = ref.simulate_factual(p, T, Eq[eq])
            elif kind == 'one':
                sim = ref.simulate_counterfactual_1_step(p, T, Eq[eq])
            else:
                sim = ref.simulate_counterfactuals_treatment_seq(p, T, H, Eq[eq])
            for k, v in p.items():
                out[f'{eq}/{kind}/params/{k}'] = np.asarray(v)
            for k, v in sim.items():
                out[f'{eq}/{kind}/out/{k}'] = np.asarray(v)
            out[f'{eq}/{kind}/rng_after'] = np.array(rng_digest())
    path = os.path.join(ROOT, 'tests', 'golden', 'ref_continuous_small.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), len(out))


if __name__ == '__main__':
    main()
