"""Per-kernel times of one mode / block / signal_l1 configuration (not part of the product).
usage: python tools/cfg_probe.py mode block [K decimate]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ofdm = importlib.import_module("dab-radio_b200.ofdm")
mode, block = int(sys.argv[1]), int(sys.argv[2])
l1 = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else None
fl = bench.MODE_FRAME_LEN[mode]
n = int(os.environ.get("PROBE_STREAMS", "1024"))
iq, _ = bench.build_streams_on_device(torch, n, 11, seed=4321 + mode, mode=mode, frame_len=fl, period=8)
for ways in os.environ.get("PROBE_WAYS", "auto,1").split(","):
    if ways != "auto":
        os.environ["DAB_B200_PIPELINE_WAYS"] = ways.split(":")[0]
        if ":" in ways:
            os.environ["DAB_B200_WAY_MIN_SAMPLES"] = ways.split(":")[1]
    d = ofdm.OfdmDemodBatch(mode, n_streams=n, device=0, max_block_samples=block)
    os.environ.pop("DAB_B200_PIPELINE_WAYS", None)
    os.environ.pop("DAB_B200_WAY_MIN_SAMPLES", None)
    d.disable_callback()
    if l1:
        cfg = d.get_config(0)
        cfg.signal_l1_nb_samples, cfg.signal_l1_nb_decimate = l1
        d.set_config(cfg)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d.set_cuda_stream(st.cuda_stream)
        ring = bench.ResidentRing(d, iq, fl, 8)
        per = fl // block
        for _ in range(4 * per):
            ring.step(block)
        d.join(); st.synchronize()
        if ways == "1":
            d.set_kernel_timing(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        e0.record()
        K = 6
        th0 = time.perf_counter()
        for _ in range(K * per):
            ring.step(block)
        d.join()
        host_ms = (time.perf_counter() - th0) * 1e3 / K
        e1.record(); st.synchronize()
    ms = e0.elapsed_time(e1) / K
    locked = sum(1 for s in range(n) if d.state(s)["state"] != 0)
    line = f"mode {mode} block {block} l1 {l1} ways {ways}: host issue {host_ms:.3f} ms/frame-period, device {ms:.4f} ms/frame-period  {n * fl / ms / 1e3:.0f} MS/s  locked {locked}/{n}"
    if ways == "1":
        kt = d.kernel_times()
        line += "\n   per frame-period: frame " + " ".join(f"{v / K:.3f}" for v in kt["frame_ms"][:4]) + "  control " + " ".join(f"{v / K:.3f}" for v in kt["control_ms"][:5]) + f"  l1 {kt['control_ms'][7] / K:.3f}"
    print(line, flush=True)
    d.close()
