"""Where the host-buffer path spends its step (not part of the product): raw-u8 and complex64 e2e with / without the soft-bit
download, for several pipeline way counts.  usage: python tools/e2e_probe.py"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

ofdm = importlib.import_module("dab-radio_b200.ofdm")
n, FL, K = 1024, bench.FRAME_LEN, 8
iq, _ = bench.build_streams_on_device(torch, n, 2, seed=1)
for u8 in (True, False):
    if u8:
        q = torch.clamp(torch.view_as_real(iq[:, :FL]) * 127.5 + 127.5, 0.0, 255.0).to(torch.uint8)
        host = torch.empty((n, FL, 2), dtype=torch.uint8).pin_memory(); host.copy_(q); del q
    else:
        host = torch.empty((n, FL), dtype=torch.complex64).pin_memory(); host.copy_(iq[:, :FL])
    torch.cuda.synchronize()
    for ways in ("4", "8", "16", "2"):
        for cb in (True, False):
            os.environ["DAB_B200_PIPELINE_WAYS"] = ways
            d = ofdm.OfdmDemodBatch(1, n_streams=n, device=0, max_block_samples=FL, raw_u8=u8)
            if cb:
                d.use_counting_callback()
            else:
                d.disable_callback()
            p, nn = d.pointer_arrays([host[s].data_ptr() for s in range(n)], [FL] * n)
            for _ in range(3):
                d.process_batch_prepared(p, nn, u8)
            d.sync(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(K):
                d.process_batch_prepared(p, nn, u8)
            d.sync(); torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / K
            print(f"u8={u8} ways={ways} callback={cb}: {dt * 1e3:.2f} ms/step  {n * FL / dt / 1e9:.2f} GS/s", flush=True)
            d.close()
    del host
