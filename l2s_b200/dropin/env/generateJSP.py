"""Synthetic instance generator with the reference's semantics (env/generateJSP.py:14-18):
durations = randint(low, high) (high exclusive), machine rows = independent permutations of 1..n_m."""
import numpy as np


def uni_instance_gen(n_j, n_m, low, high):
    times = np.random.randint(low=low, high=high, size=(n_j, n_m))
    order = np.random.sample((n_j, n_m)).argsort(axis=1)
    machines = (order + 1).astype(np.int64)
    return times, machines
