import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tpod():
    d = np.load(os.path.join(GOLDEN, "tpod.npz"))
    return d["y"].astype(np.float64), d["gen"].astype(np.int8)


def synth(n, p, k=1, seed=20261018, h2=0.5, causal=0.01):
    """Synthetic genotypes/phenotypes of SURVEY.md 8d: f_j~U(.05,.5), X_ij~Binom(2,f_j), 1% causal."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.05, 0.5, size=p)
    X = (rng.random((n, p)) < f).astype(np.int8) + (rng.random((n, p)) < f).astype(np.int8)
    nc = max(1, int(round(causal * p)))
    Y = np.empty((n, k))
    for t in range(k):
        beta = np.zeros(p)
        beta[rng.choice(p, nc, replace=False)] = rng.normal(size=nc)
        g = X @ beta
        g = (g - g.mean()) / (g.std() + 1e-12) * np.sqrt(h2)
        Y[:, t] = g + rng.normal(size=n) * np.sqrt(1 - h2)
    return np.asfortranarray(X), (Y[:, 0] if k == 1 else np.asfortranarray(Y))
