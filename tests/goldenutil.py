"""Loaders for tests/golden/*.npz (fixtures generated from the compiled reference by tests/golden/make_golden.py)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def ofdm_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "ofdm_*.npz")))


def viterbi_cases():
    return [str(n) for n in load("viterbi.npz")["names"]]


def viterbi_case(g, oracle_mod, key):
    name = key.split("__")[0]
    spec = [(int(pi), int(n)) for pi, n in g[f"{name}__spec"]]
    segs = [(oracle_mod.puncture_code(pi) if pi else oracle_mod.PI_X, n) for pi, n in spec]
    nbytes = int(g[f"{name}__nbytes"][0])
    return segs, nbytes, g[f"{key}__soft"], g[f"{key}__out"], int(g[f"{key}__err"][0])
