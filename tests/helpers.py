"""Shared synthetic inputs for the parity tests."""
import math

import numpy as np


def f_sin(x):
    return math.sin(2 * math.pi * x)


def f_cos(x):
    return math.cos(2 * math.pi * x)


def f_gauss(x):
    return math.exp(-2 * math.pi ** 2 * (x - 0.5) ** 2)


def relerr(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def product_state(oracle, D, k, n, f, scheme="sparse"):
    v1 = oracle.coeffs_1d(k, n, f)
    return oracle.tensor_construct(D, k, n, [v1] * D, scheme=scheme)


def random_state(N, seed=0):
    return np.random.default_rng(seed).standard_normal(N)
