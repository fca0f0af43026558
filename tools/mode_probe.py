"""Per-mode throughput probe (not part of the product): resident streams of one transmission mode, several block sizes,
with the library's per-kernel CUDA-event timing.  usage: python tools/mode_probe.py [n_streams]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ofdm = importlib.import_module("dab-radio_b200.ofdm")
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for mode in (1, 2, 3, 4):
    fl = bench.MODE_FRAME_LEN[mode]
    n_frames = 12
    iq, _ = bench.build_streams_on_device(torch, n_streams, n_frames + 1, seed=99 + mode, mode=mode, frame_len=fl)
    for block in (4096, fl // 4, fl):
        for ways in ("4", "1"):
            os.environ["DAB_B200_PIPELINE_WAYS"] = ways
            d = ofdm.OfdmDemodBatch(mode, n_streams=n_streams, device=0, max_block_samples=block)
            d.disable_callback()
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                d.set_cuda_stream(st.cuda_stream)
                d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
                per_frame = fl // block
                for _ in range(4 * per_frame):
                    d.advance_uniform(block)
                d.join(); st.synchronize()
                f0 = sum(d.state(s)["total_frames_read"] for s in range(0, n_streams, n_streams // 16))
                if ways == "1":
                    d.set_kernel_timing(True)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K = 6 * per_frame
                for _ in range(K):
                    d.advance_uniform(block)
                d.join()
                e1.record(); st.synchronize()
            ms = e0.elapsed_time(e1)
            f1 = sum(d.state(s)["total_frames_read"] for s in range(0, n_streams, n_streams // 16))
            locked = sum(1 for s in range(0, n_streams, n_streams // 16) if d.state(s)["state"] == 4)
            line = f"mode {mode} block {block:6d} ways {ways}: {n_streams * block * K / ms / 1e3:9.1f} MS/s  {ms / 6:.3f} ms/frame-period  frames/stream {(f1 - f0) / 16:.2f} of 6  reading_symbols {locked}/16"
            if ways == "1":
                kt = d.kernel_times()
                line += f"  frame_ms {sum(kt['frame_ms']) / 6:.3f} control_ms {sum(kt['control_ms'][:7]) / 6:.3f} l1_ms {kt['control_ms'][7] / 6:.3f}"
            print(line, flush=True)
            d.close()
    del iq
    torch.cuda.empty_cache()
