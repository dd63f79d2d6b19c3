"""Synthetic workload inputs: a counter-based splitmix64 generator so that the product
and the oracle can be fed identical doubles without sharing code.  Go's math/rand stream
(bench/mlp.go:60,95,105) cannot be reproduced without a Go toolchain; the distributions match:
U[-1,1) weights, biases, inputs and R-vectors."""
from __future__ import annotations

import numpy as np

_M = (1 << 64) - 1


def uniform(seed: int, count: int, lo: float = -1.0, hi: float = 1.0) -> np.ndarray:
    z = np.arange(count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z * np.uint64(0x9E3779B97F4A7C15) + np.uint64((seed * 0xD1B54A32D192ED03 + 0x9E3779B97F4A7C15) & _M)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return lo + (hi - lo) * ((z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0))


def mlp_params(layer_sizes, seed=123, rseed=125, weight_scale=None):
    """Parameters and R-vectors in bench/mlp.go:59-116 order: [W0, b0, W1, b1, ...]."""
    params, rvecs = [], []
    for k, out_size in enumerate(layer_sizes[1:]):
        in_size = layer_sizes[k]
        sc = 1.0 if weight_scale is None else weight_scale(in_size)
        params.append(uniform(seed + 1000 * k, in_size * out_size) * sc)
        params.append(uniform(seed + 1000 * k + 1, out_size))
    for j, p in enumerate(params):
        rvecs.append(uniform(rseed + 1000 * j, p.size))
    return params, rvecs


def mlp_input(layer_sizes, n, seed=124):
    return uniform(seed, n * layer_sizes[0])
