"""Back-to-back resident calls without host synchronisation vs the oracle, repeated (not part of the product).
usage: python tools/async_probe.py mode block [sync]"""
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
mode, block = int(sys.argv[1]), int(sys.argv[2])
sync_each = len(sys.argv) > 3
fl = bench.MODE_FRAME_LEN[mode]
n = int(os.environ.get("PROBE_STREAMS", "1024"))
iq, _ = bench.build_streams_on_device(torch, n, 11, seed=4321 + mode, mode=mode, frame_len=fl, period=8)
calls = 10 * fl // block
want = {}
for s in range(0, n, int(os.environ.get("PROBE_STEP", "37"))):
    o = po.OracleOfdmDemod(mode)
    o.process_blocks(iq[s].cpu().numpy()[:calls * block], block)
    st = o.state()
    want[s] = (st["state"], st["total_frames_read"], st["total_frames_desync"])
    o.close()
for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
    d = ofdm.OfdmDemodBatch(mode, n_streams=n, device=0, max_block_samples=block)
    d.disable_callback()
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        d.set_cuda_stream(stream.cuda_stream)
        d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
        for k in range(calls):
            d.advance_uniform(block)
            if sync_each:
                d.sync()
        d.sync()
    bad = []
    for s, w in want.items():
        st = d.state(s)
        g = (st["state"], st["total_frames_read"], st["total_frames_desync"])
        if g != w:
            bad.append((s, g, w))
    print(f"rep {rep} sync_each={sync_each}: {len(bad)} of {len(want)} streams differ from the oracle", bad[:6], flush=True)
    d.close()
