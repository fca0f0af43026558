"""Finds the first Process() call at which a stream of the modes bench diverges from the oracle and prints the fine-time-sync
margin there (not part of the product).  usage: python tools/diverge_probe.py mode block stream"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ofdm = importlib.import_module("dab-radio_b200.ofdm")
mode, block, stream = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fl = bench.MODE_FRAME_LEN[mode]
iq, _ = bench.build_streams_on_device(torch, 1024, 11, seed=4321 + mode, mode=mode, frame_len=fl, period=8)
row = iq[stream].cpu().numpy()
batched = len(sys.argv) > 4
d = ofdm.OfdmDemodBatch(mode, n_streams=1024 if batched else 1, device=0, max_block_samples=block)
if batched:
    d.disable_callback()
    d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
o = po.OracleOfdmDemod(mode)
p = po.params(mode)
cp = p["nb_cyclic_prefix"]
for k in range(0, 10 * fl // block):
    x = row[k * block:(k + 1) * block]
    if batched:
        d.advance_uniform(block)
    else:
        d.process(0, x)
    o.process(x)
    gs = stream if batched else 0
    so, sd = o.state(), d.state(gs)
    imp_o, imp_d = o.impulse_response(), d.impulse_response(gs)
    if (so["state"], so["total_frames_read"], so["total_frames_desync"]) != (sd["state"], sd["total_frames_read"], sd["total_frames_desync"]):
        print("diverged at call", k, "sample", k * block)
        print(" oracle", so)
        print(" gpu   ", sd)
        for name, imp in (("oracle", imp_o), ("gpu", imp_d)):
            w = (1.0 - 0.85 * np.abs(cp - np.arange(imp.size)) / p["nb_symbol_period"]) * imp
            print(f" {name}: impulse max {imp.max():.3f} at {imp.argmax()}, weighted max {w.max():.3f} at {w.argmax()}, mean {imp.mean():.3f}, margin {w.max() - imp.mean():.3f} dB (threshold 20)")
        break
else:
    print("no divergence in", 10 * fl // block, "calls")
