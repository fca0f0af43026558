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


def ensemble_case():
    """-> (subs as tuples, frames [n][nb_frame_bits] int8, fib_bytes, fib_valid, msc_len, msc bytes as [frame][cif][sub] arrays)"""
    g = load("ensemble.npz")
    nb_cifs, nb_fic_bits, nb_cif_bits = 4, 9216, 55296
    used = int(g["used_bits"][0])
    frames = []
    f = 0
    while f"frame{f}_fic" in g:
        msc = np.zeros((nb_cifs, nb_cif_bits), np.int8)
        msc[:, :used] = g[f"frame{f}_msc"]
        frames.append(np.concatenate([g[f"frame{f}_fic"], msc.reshape(-1)]))
        f += 1
    lens = g["msc_len"]
    flat = g["msc_bytes"]
    pos = 0
    msc_bytes = []
    for fi in range(lens.shape[0]):
        per_cif = []
        for c in range(lens.shape[1]):
            per_sub = []
            for k in range(lens.shape[2]):
                n = int(lens[fi, c, k])
                per_sub.append(flat[pos:pos + n])
                pos += n
            per_cif.append(per_sub)
        msc_bytes.append(per_cif)
    subs = [tuple(int(v) for v in row) for row in g["subs"]]
    return subs, frames, g["fib_bytes"], g["fib_valid"], lens, msc_bytes, g["scrambler"]
