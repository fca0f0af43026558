"""Mode I step probe (not part of the product): 1024 resident streams, one-frame blocks, under several library switches.
usage: python tools/step_probe.py [n_streams]"""
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
mode, fl = 1, bench.FRAME_LEN
iq, _ = bench.build_streams_on_device(torch, n_streams, 30, seed=100, mode=mode, frame_len=fl)
CASES = [{}]
for extra in sys.argv[2:]:
    CASES.append(dict(kv.split("=") for kv in extra.split(",")))
for case in CASES:
    for ways in os.environ.get("PROBE_WAYS", "4,1").split(","):
        env = dict(case, DAB_B200_PIPELINE_WAYS=ways)
        for k, v in env.items():
            os.environ[k] = v
        d = ofdm.OfdmDemodBatch(mode, n_streams=n_streams, device=0, max_block_samples=fl)
        for k in env:
            del os.environ[k]
        d.disable_callback()
        if os.environ.get("PROBE_DECIMATE"):
            cfg = d.get_config(0)
            cfg.signal_l1_nb_decimate = int(os.environ["PROBE_DECIMATE"])
            d.set_config(cfg)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            d.set_cuda_stream(st.cuda_stream)
            d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
            for _ in range(5):
                d.advance_uniform(fl)
            d.join(); st.synchronize()
            if ways == "1":
                d.set_kernel_timing(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            K = 24
            for _ in range(K):
                d.advance_uniform(fl)
            d.join()
            e1.record(); st.synchronize()
        ms = e0.elapsed_time(e1) / K
        frames = sum(d.state(s)["total_frames_read"] for s in range(0, n_streams, max(1, n_streams // 16)))
        line = f"{str(case):60s} ways {ways}: {ms:.4f} ms/step  frames(16 streams) {frames}"
        if ways == "1":
            kt = d.kernel_times()
            line += "  frame " + " ".join(f"{v / K:.3f}" for v in kt["frame_ms"][:3]) + "  control " + " ".join(f"{v / K:.3f}" for v in kt["control_ms"][:4])
        print(line, flush=True)
        d.close()
