"""Shared synthetic data sets (seeded) for the parity tests."""
import numpy as np

BERNOULLI = {"y": np.array([0, 1, 0, 0, 0, 0, 0, 0, 0, 1], dtype=np.int32)}  # C/examples/bernoulli/bernoulli.data.json
EIGHT_SCHOOLS = {"y": np.array([28, 8, -3, 7, -1, 1, 18, 12], dtype=np.float64),
                 "sigma": np.array([15, 10, 16, 11, 9, 11, 10, 18], dtype=np.float64)}  # C/src/test/test-models/eight_schools.data.json


def logistic_data(N, K, seed=20240301):
    """SURVEY.md section 8(d) c3 recipe at arbitrary size: X ~ N(0,1), alpha* = 0.5, beta* ~ N(0, 0.5^2)."""
    rng = np.random.default_rng(seed)
    X = rng.normal(size=(N, K))
    beta = rng.normal(0, 0.5, size=K)
    p = 1.0 / (1.0 + np.exp(-(0.5 + X @ beta)))
    y = (rng.uniform(size=N) < p).astype(np.int32)
    return {"X": X, "y": y}


def poisson_data(N, K, seed=20240303):
    """c5 recipe: X ~ N(0, 0.1^2), alpha* = 1, beta* ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    X = rng.normal(0, 0.1, size=(N, K))
    beta = rng.normal(size=K)
    y = rng.poisson(np.exp(1.0 + X @ beta)).astype(np.int32)
    return {"X": X, "y": y}


def irt_data(I, J, seed=20240302):
    """c4 recipe: theta ~ N(0,1), a ~ LogNormal(0, 0.3), b ~ N(0,1)."""
    rng = np.random.default_rng(seed)
    theta = rng.normal(size=I)
    a = rng.lognormal(0, 0.3, size=J)
    b = rng.normal(size=J)
    p = 1.0 / (1.0 + np.exp(-(a[None, :] * (theta[:, None] - b[None, :]))))
    y = (rng.uniform(size=(I, J)) < p).astype(np.int8)
    return {"y": y}


SMALL = {
    "std_normal": {"D": 3},
    "bernoulli": BERNOULLI,
    "eight_schools_nc": EIGHT_SCHOOLS,
}
