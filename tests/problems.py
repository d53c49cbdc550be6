"""Seeded build-defined objectives of BASELINE.json configs[2]."""
import numpy as np

SEED = 20261018


def gaussian_problem(n=128, key=SEED + 11):
    """Λ = A Aᵀ / n + I, μ ~ N(0,1)."""
    rng = np.random.Generator(np.random.Philox(key=key))
    A = rng.standard_normal((n, n))
    lam = A @ A.T / n + np.eye(n)
    lam = 0.5 * (lam + lam.T)
    mu = rng.standard_normal(n)
    return mu, np.asfortranarray(lam)


def logistic_problem(m=1024, n=128, key=SEED + 12, lam=1.0):
    """X ~ N(0,1)/√n (m×n), y ~ Bernoulli(σ(X θ_true)), ridge λ."""
    rng = np.random.Generator(np.random.Philox(key=key))
    X = rng.standard_normal((m, n)) / np.sqrt(n)
    theta_true = rng.standard_normal(n)
    p = 1.0 / (1.0 + np.exp(-(X @ theta_true)))
    y = (rng.random(m) < p).astype(np.float64)
    return np.asfortranarray(X), y, lam
