"""Steady-state Mode I steps for ncu captures (not part of the product): 1024 resident streams, one-frame blocks.
usage: [DAB_B200_PIPELINE_WAYS=1] ncu ... python tools/ncu_target.py [steps]   (8 steps by default; all streams lock in the first 5)"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ofdm = importlib.import_module("dab-radio_b200.ofdm")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n_streams, fl = 1024, bench.FRAME_LEN
iq, _ = bench.build_streams_on_device(torch, n_streams, steps + 2, seed=100, mode=1, frame_len=fl)
d = ofdm.OfdmDemodBatch(1, n_streams=n_streams, device=0, max_block_samples=fl)
d.disable_callback()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    d.set_cuda_stream(st.cuda_stream)
    d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
    for _ in range(steps):
        d.advance_uniform(fl)
    d.join()
    st.synchronize()
print("frames", sum(d.state(s)["total_frames_read"] for s in range(n_streams)))
d.close()
