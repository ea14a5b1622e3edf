"""Generate tests/golden/sinkhorn.npz from the UNMODIFIED reference (dev container only): `python -m oracle.gen_golden_assoc`.
TEST INFRASTRUCTURE ONLY.  For affinity matrices shaped like the association step's (entries = sigmoid outputs of the
Affinity MLP, m previous x n current objects) it stores what the reference's own `log_optimal_transport`
(models/utils/track4d_utils.py:414) and `Track4D.sinkhorn_module` (models/track4d.py:166) return."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

CASES = [(1, 1), (3, 5), (7, 2), (20, 20), (12, 33), (40, 31), (1, 9), (64, 64)]


def make_aff(m, n, seed):
    rng = np.random.default_rng([seed, m, n])
    a = 1.0 / (1.0 + np.exp(-rng.normal(-2.0, 1.5, (1, m, n))))          # mostly low affinities ...
    for k in range(min(m, n)):                                            # ... and a noisy permutation of true matches
        if rng.random() < 0.8:
            a[0, k, (k * 7 + 3) % n] = 1.0 / (1.0 + np.exp(-rng.normal(3.0, 1.0)))
    return a.astype(np.float32)


if __name__ == "__main__":
    from oracle import ref_harness

    net = ref_harness.make_track4d(npoints=512)
    from models.utils.track4d_utils import log_optimal_transport
    save = {"cases": np.array(CASES)}
    for m, n in CASES:
        aff = torch.from_numpy(make_aff(m, n, 1234))
        scores = log_optimal_transport(aff, torch.tensor(0.9), 500)
        idx1 = net.sinkhorn_module(aff, None)
        save[f"aff_{m}_{n}"] = aff.numpy()
        save[f"scores_{m}_{n}"] = scores.numpy()
        save[f"idx1_{m}_{n}"] = idx1.numpy()
        print(m, n, "matched", int((idx1 >= 0).sum()), "of", n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "sinkhorn.npz"), **save)
