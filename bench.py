#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native DAB receive hot path.

Metric (BASELINE.json): MSamples/s IQ -> soft bits, and real-time (2.048 MS/s) DAB Mode I streams, at 1/2/4/8 B200.
Workload (BASELINE.json configs[1]): DAB Mode I, 1024 independent synthetic streams batched per GPU, OFDM demodulation to
int8 soft bits.  One step = every stream advances by one transmission frame (196 608 samples) through the full receive path
(state machine, PRS coarse-frequency + fine-time sync, PLL, cyclic-prefix phase error, FFT, DQPSK, de-interleave, quantise).

  python bench.py --gpus N --steps K --warmup W          our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W  the reference's own CPU implementation on the host cores

`value`      device-time throughput with the IQ already resident in HBM (dab_ofdm_attach_device_streams + advance)
`e2e`        the same metric through the reference-facing C-ABI call dab_ofdm_process_batch with pinned HOST buffers:
             H2D of every block and D2H of every frame's soft bits (callback) inside the timed region
`roofline`   dominant kernel (ofdm_frame_kernel): algorithmic bytes / CUDA-event time vs the measured HBM copy peak
`cpu_baseline` the reference (oracle/_ref, compiled from its own sources) on the host cores, bounded sample, rank 0 at N = 1
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MSamples/s IQ->soft bits and real-time DAB Mode I streams per GPU at 1/2/4/8 B200"
MODE = 1
FS = 2.048e6
FRAME_LEN = 196608                      # Mode I transmission frame, samples (96 ms)
FRAME_BITS = 230400
ALGO_BYTES_PER_FRAME = 196608 * 8 + 230400   # complex64 in + int8 out = 9.172 B/sample
N_STREAMS = 1024
LOCK_FRAMES = 5                         # untimed acquisition frames before the warm-up (see run_ours)
RING_PERIOD = 8                         # frames after which the synthetic content of a resident stream repeats (ResidentRing)
_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line, on the real stdout (fd 1 is parked on stderr while the run is in progress)."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)
POOL_FRAMES = 12                        # distinct modulated frames the synthetic streams are drawn from


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML is polled from a thread every few ms (the timed
    region of a 10-step run lasts tens of ms, far below what `nvidia-smi -lms` can resolve); same fields as the
    B200_PROFILING.md nvidia-smi line."""

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.rows = []
        self.stop_flag = False
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                power = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((time.time(), sm, reasons, power))
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            time.sleep(0.002)

    def stop(self, t0, t1):
        self.stop_flag = True
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {self.err}"]}
        self.thread.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["no samples"]}
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        # nvml clocks event reason bits
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake"}
        reasons = sorted(n for b, n in names.items() if bits & b)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": reasons, "samples": len(rows),
                "power_w_max": round(max(r[3] for r in rows), 1)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = [p.strip() for p in vis.split(",") if p.strip()]
        if local_rank < len(parts) and parts[local_rank].isdigit():
            return int(parts[local_rank])
    return local_rank


def aggregate(ms_local, samples_per_rank, world, dist=None, device="cuda"):
    """MAX over ranks of the device time; value = whole-job MSamples/s (every rank processes samples_per_rank)."""
    import torch
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    return ms_max, world * samples_per_rank / (ms_max * 1e-3) / 1e6


def shard_streams(n_global, rank, world):
    """global stream ids owned by `rank`: stream i -> rank i mod world; no data crosses ranks"""
    return range(rank, n_global, world)


def build_streams_on_device(torch, n_streams, n_frames, seed, mode=None, frame_len=None, period=None, stream_ids=None):
    """Synthetic Mode I streams in HBM: [n_streams, n_frames * FRAME_LEN] complex64.  With `period` the frame sequence of every
    stream repeats after `period` frames, so the buffer can be used as a ring (dab_ofdm_rebase_device_streams): the signal at
    sample i + period * FRAME_LEN continues the one at sample i (same frames, continuous CFO phase; the noise differs).

    A pool of POOL_FRAMES frames is modulated on the CPU exactly as simulate_transmitter does (random payload ->
    OFDM modulator, scale 4/1536); each stream is a random sequence of pool frames, rotated by a random start offset in
    [0, FRAME_LEN), shifted by its own carrier frequency offset (multiples of Fs/FRAME_LEN so that the one-frame host block of
    the e2e leg is phase continuous) and given its own AWGN (SNR 25 dB)."""
    mode = MODE if mode is None else mode
    frame_len = FRAME_LEN if frame_len is None else frame_len
    import dabgen
    from oracle import pyoracle as po
    rng = np.random.default_rng(seed)
    pool = np.stack([po.modulate(mode, rng.integers(0, 256, dabgen.payload_bytes(mode), dtype=np.uint8)) for _ in range(POOL_FRAMES)])
    pool = (pool * np.float32(4.0 / 1536)).astype(np.complex64)
    sig_pow = float(np.mean(np.abs(pool) ** 2))
    noise_sigma = float(np.sqrt(sig_pow / 10 ** (25.0 / 10) / 2))
    d_pool = torch.from_numpy(pool.view(np.float32).reshape(POOL_FRAMES, frame_len, 2)).cuda()
    d_pool = torch.view_as_complex(d_pool)
    total = n_frames * frame_len
    out = torch.empty((n_streams, total), dtype=torch.complex64, device="cuda")
    g = torch.Generator(device="cuda")
    g.manual_seed(seed if stream_ids is None else seed + 7919 * int(stream_ids[0]))
    max_bin = 4800 * frame_len // FRAME_LEN                   # x Fs/frame_len: +-50 kHz in every mode (10.4 Hz steps in Mode I)
    if stream_ids is None:
        starts = rng.integers(0, frame_len, n_streams)
        cfo_bins = rng.integers(-max_bin, max_bin + 1, n_streams)
        choice = rng.integers(0, POOL_FRAMES, (n_streams, n_frames + 1))
    else:
        # what a stream looks like depends on its GLOBAL id only (shard_streams: stream i lives on rank i mod N): the same job on
        # 1, 2, 4 or 8 GPUs demodulates the same streams
        per = [np.random.default_rng([seed, int(g)]) for g in stream_ids]
        starts = np.array([r.integers(0, frame_len) for r in per])
        cfo_bins = np.array([r.integers(-max_bin, max_bin + 1) for r in per])
        choice = np.stack([r.integers(0, POOL_FRAMES, n_frames + 1) for r in per])
    if period:
        choice = choice[:, np.arange(n_frames + 1) % period]
    chunk = 16
    ar = torch.arange(total, device="cuda", dtype=torch.int64)
    for s0 in range(0, n_streams, chunk):
        s1 = min(n_streams, s0 + chunk)
        st = torch.from_numpy(starts[s0:s1]).cuda().view(-1, 1)
        idx = ar.view(1, -1) + st                              # position in the un-rotated stream
        fi = idx // frame_len
        within = idx - fi * frame_len
        ch = torch.from_numpy(choice[s0:s1]).cuda()
        frame_id = torch.gather(ch, 1, fi)
        x = d_pool[frame_id, within]
        f = torch.from_numpy(cfo_bins[s0:s1].astype(np.float64) / frame_len).cuda().view(-1, 1)
        phase = torch.remainder(f * ar.view(1, -1).to(torch.float64), 1.0) * (2.0 * np.pi)
        rot = torch.polar(torch.ones_like(phase, dtype=torch.float32), phase.to(torch.float32))
        noise = torch.randn((s1 - s0, total, 2), device="cuda", generator=g) * noise_sigma
        out[s0:s1] = x * rot + torch.view_as_complex(noise)
        del idx, fi, within, frame_id, x, phase, rot, noise
    if period and n_frames > period:
        # the ring is exactly periodic, noise included: a frame that straddles a rebase is read by the frame kernel from the other
        # copy of the same samples
        out[:, period * frame_len:] = out[:, :total - period * frame_len].clone()
    return out, {"starts": starts, "cfo_bins": cfo_bins}


class ResidentRing:
    """n_streams device-resident streams whose content repeats every `period` frames, advanced block by block for as long as
    wanted: once the cursor is period + 2 frames into the buffer the origin moves forward by `period` frames
    (dab_ofdm_rebase_device_streams) and the demodulator continues in the same buffer, on the same signal."""

    def __init__(self, d, iq, frame_len, period, raw_u8=False):
        self.d, self.iq, self.fl, self.period, self.raw_u8 = d, iq, frame_len, period, raw_u8
        self.total = iq.shape[1] // 2 if raw_u8 else iq.shape[1]      # samples per row
        self.pos = 0            # samples advanced since the origin
        self.fed = []           # (first sample, count) of every call so far, in buffer coordinates: the oracle replays them
        d.attach_device_streams(iq.data_ptr(), self.total, self.total)

    def step(self, n=None):
        n = self.fl if n is None else n
        if self.pos + n > self.total:   # the buffer holds period + 2 frames + one block: at least two frames stay behind the cursor
            self.d.rebase_device_streams(self.period * self.fl)
            self.pos -= self.period * self.fl
        self.d.advance_uniform(n)
        self.fed.append((self.pos, n))
        self.pos += n

    def replay(self, stream):
        """the samples stream `stream` has been fed so far, as one host array"""
        row = self.iq[stream].cpu().numpy()
        if self.raw_u8:   # what the reference's raw_u8 reader hands to OFDM_Demod (app_iq_readers.h:36-43, 76-88)
            row = ((row.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).view(np.complex64)
        return np.concatenate([row[a:a + n] for a, n in self.fed])


def check_streams_against_oracle(d, ring, mode, streams, block, l1_config=None):
    """Feeds the oracle (oracle/dab_oracle.c, the CPU restatement of OFDM_Demod) exactly what the timed run fed the chosen streams
    and compares: frames completed, desyncs, state, frequency offsets within 1e-3 bin, signal average, and the soft bits of
    the last frame within +-1 LSB on >= 99.9 % (north_star's acceptance).  Raises on mismatch."""
    import dabgen
    from oracle import pyoracle as po
    nfft = po.params(mode)["nb_fft"]
    out = []
    for s in streams:
        x = ring.replay(s)
        o = po.OracleOfdmDemod(mode)
        if l1_config:
            o.config.signal_l1_nb_samples, o.config.signal_l1_nb_decimate = l1_config
        o.process_blocks(x, block)
        so, sd = o.state(), d.state(s)
        assert sd["state"] == so["state"] and sd["total_frames_read"] == so["total_frames_read"] and \
            sd["total_frames_desync"] == so["total_frames_desync"], f"stream {s}: state {sd} vs oracle {so}"
        assert abs(sd["fine_frequency_offset"] - so["fine_offset"]) * nfft < 1e-3 and abs(sd["coarse_frequency_offset"] - so["coarse_offset"]) * nfft < 1e-3, \
            f"stream {s}: offsets {sd} vs oracle {so}"
        assert abs(sd["signal_average"] - so["signal_average"]) <= 1e-4 * abs(so["signal_average"]), f"stream {s}: signal average"
        lsb1 = eq = None
        if o.frames_done() > 0:
            eq, lsb1, mx = dabgen.compare_bits(d.frame_data_bits(s), o.frame(o.frames_done() - 1)[1])
            assert lsb1 >= 0.999, f"stream {s}: last frame only {lsb1:.5f} within +-1 LSB (max {mx})"
        out.append({"stream": int(s), "frames": int(so["total_frames_read"]), "desyncs": int(so["total_frames_desync"]), "locked": bool(so["state"] != 0),
                    "last_frame_within_1lsb": None if lsb1 is None else round(lsb1, 5), "last_frame_identical": None if eq is None else round(eq, 5)})
        o.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # NCCL prints its version banner on stdout when NCCL_DEBUG is set; stdout must carry the ONE JSON line only, so fd 1 is
    # pointed at stderr until the result is printed (see emit()).
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    pkg = importlib.import_module("dab-radio_b200")
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    if rank == 0 and not os.path.exists(pkg.capi.LIB_PATH):
        pkg.build()
    if world > 1:
        dist.barrier()

    K, W = args.steps, args.warmup
    n_streams = args.streams
    # acquisition is set-up, not the measured workload: LOCK_FRAMES untimed frames let every stream find its NULL symbol and lock
    # (a stream that is still searching runs FindNullPowerDip over whole blocks), then come the W warm-up and the K timed steps.
    # The streams live in a ring of RING_PERIOD + 3 frames per stream whose content repeats every RING_PERIOD frames, so a run
    # can be as long as wanted (the sustained leg) in 17.7 GB of HBM.
    my_streams = shard_streams(world * n_streams, rank, world)      # global stream ids of this rank: i -> rank i mod N
    iq, gen = build_streams_on_device(torch, n_streams, RING_PERIOD + 3, seed=1234, period=RING_PERIOD, stream_ids=my_streams)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ------------------------------------------------------------------ value: IQ resident in HBM
    d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=local_rank, max_block_samples=FRAME_LEN)
    d.disable_callback()
    # a dedicated (non-default) torch stream carries both the library's kernels and the timing events
    work_stream = torch.cuda.Stream()
    torch.cuda.set_stream(work_stream)
    d.set_cuda_stream(work_stream.cuda_stream)
    ring = ResidentRing(d, iq, FRAME_LEN, RING_PERIOD)
    for _ in range(LOCK_FRAMES + W):
        ring.step()
    d.join()
    barrier()
    probe = range(0, n_streams, max(1, n_streams // 16))
    frames_before = sum(d.state(s)["total_frames_read"] for s in probe)
    launches0 = d.kernel_launches()
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(K):
        ring.step()
    d.join()            # the handle's stream (= work_stream) now waits for every pipeline way: ev1 closes the whole job
    ev1.record()
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms = ev0.elapsed_time(ev1)
    launches = d.kernel_launches() - launches0
    frames_after = sum(d.state(s)["total_frames_read"] for s in probe)
    frames_per_stream = (frames_after - frames_before) / len(probe)
    samples_per_rank = n_streams * FRAME_LEN * K
    ms_max, value = aggregate(ms, samples_per_rank, world, dist)
    # the timed path checked against the oracle: one stream per pipeline way (the ways run on different CUDA streams), fed to the
    # CPU restatement of OFDM_Demod sample for sample as the ring fed it here
    t0 = time.time()
    parity = {"checked_streams": check_streams_against_oracle(d, ring, MODE, [w * n_streams // 4 + (7 * w) % max(1, n_streams // 4) for w in range(4)],
                                                              FRAME_LEN),
              "oracle": "oracle/dab_oracle.c (OFDM_Demod restated, pinned against the compiled reference); identical frame / desync counts and "
                        "state, offsets within 1e-3 bin, signal average within 1e-4, last frame's soft bits within +-1 LSB on >= 99.9 %",
              "frames_per_stream_checked": LOCK_FRAMES + W + K}
    parity["seconds"] = round(time.time() - t0, 2)
    states = [d.state(s) for s in range(n_streams)] if n_streams <= 4096 else []
    locked = sum(1 for st in states if st["state"] != 0)      # anything but FINDING_NULL_POWER_DIP: the stream is tracking frames

    # ------------------------------------------------------------------ sustained: the same step back to back for >= 2 s
    sustained = None
    if not args.no_sustained:
        n_sus = max(K, int(args.sustained_seconds / max(1e-6, ms_max / K * 1e-3)))
        sampler2 = ClockSampler(physical_gpu_index(local_rank))
        sampler2.start()
        f0 = sum(d.state(s)["total_frames_read"] for s in probe)
        ev4, ev5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        tw0 = time.time()
        ev4.record()
        for _ in range(n_sus):
            ring.step()
        d.join()
        ev5.record()
        barrier()
        tw1 = time.time()
        sus_ms = aggregate(ev4.elapsed_time(ev5), 0, world, dist)[0]
        f1 = sum(d.state(s)["total_frames_read"] for s in probe)
        sustained = {"value": round(world * n_streams * FRAME_LEN * n_sus / (sus_ms * 1e-3) / 1e6, 1), "unit": "MSamples/s", "steps": n_sus,
                     "seconds": round(sus_ms * 1e-3, 3), "ms_per_step": round(sus_ms / n_sus, 4),
                     "frames_per_stream_per_step": round((f1 - f0) / len(probe) / n_sus, 4), "clocks": sampler2.stop(tw0, tw1),
                     "note": "back-to-back steps in the resident ring (origin rebased every %d frames), device time, max over ranks" % RING_PERIOD}
    d.close()
    del d

    # ------------------------------------------------------------------ roofline: the same K steps once more with the pipeline
    # ways switched off, so that every kernel runs alone on the stream and its CUDA-event time is its own
    os.environ["DAB_B200_PIPELINE_WAYS"] = "1"
    d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=local_rank, max_block_samples=FRAME_LEN)
    del os.environ["DAB_B200_PIPELINE_WAYS"]
    d.disable_callback()
    d.set_cuda_stream(work_stream.cuda_stream)
    ring2 = ResidentRing(d, iq, FRAME_LEN, RING_PERIOD)
    for _ in range(LOCK_FRAMES + W):
        ring2.step()
    barrier()
    d.set_kernel_timing(True)
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    for _ in range(K):
        ring2.step()
    ev3.record()
    barrier()
    serial_ms = ev2.elapsed_time(ev3)
    kt = d.kernel_times()
    peak, peak_src = measured_peaks()
    frame_ms = kt["frame_ms"][0] / max(1, kt["frame_launches"][0])
    algo_bytes_step = ALGO_BYTES_PER_FRAME * n_streams * frames_per_stream / K
    achieved = algo_bytes_step / (frame_ms * 1e-3) / 1e9 if frame_ms > 0 else 0.0
    step_achieved = algo_bytes_step / (ms_max / K * 1e-3) / 1e9
    kernel_ms_total = sum(kt["frame_ms"]) + sum(kt["control_ms"])
    # DRAM traffic of the same kernel from the committed ncu --set full capture, scaled to the frames of one timed launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f)["ofdm_frame_v3_kernel<2048>"]
        traffic = int((t["dram_bytes_read"] + t["dram_bytes_write"]) / t["frames_in_launch"] * n_streams * frames_per_stream / K)
        traffic_src = t["source"]
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "hbm", "kernel": "ofdm_frame_v3_kernel<2048>", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms_per_launch": round(frame_ms, 4), "algorithmic_bytes_per_launch": int(ALGO_BYTES_PER_FRAME * n_streams),
                "timed": "separate pass of the same K steps with DAB_B200_PIPELINE_WAYS=1 (every kernel alone on the stream)",
                "step_achieved": round(step_achieved, 1), "step_frac": round(step_achieved / peak, 4),
                "step_note": "the whole timed step (all kernels, 4 pipeline ways) against the same peak: algorithmic bytes / ms_per_step",
                "serial_ms_per_step": round(serial_ms / K, 4),
                "share_of_step": round(kt["frame_ms"][0] / kernel_ms_total, 4) if kernel_ms_total > 0 else None,
                "control_ms_per_step": round(sum(kt["control_ms"][:7]) / K, 4), "frames_per_stream_per_step": round(frames_per_stream / K, 3),
                "per_pass_ms_per_step": {"frame": [round(v / K, 4) for v in kt["frame_ms"][:3]],
                                         "control": [round(v / K, 4) for v in kt["control_ms"][:3]],
                                         "l1_windows_side_stream": round(kt["control_ms"][7] / K, 4)}}
    d.set_kernel_timing(False)
    d.close()
    del d

    # ------------------------------------------------------------------ the same resident workload as raw 8-bit IQ (what an RTL-SDR style
    # front end delivers and the reference's apps read by default: examples/app_helpers/app_iq_readers.h raw_u8), dequantised in the kernels
    resident_u8 = None
    try:
        iq8 = torch.clamp(torch.view_as_real(iq) * (127.5 * 16.0) + 127.5, 0.0, 255.0).to(torch.uint8).reshape(n_streams, -1).contiguous()
        d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=local_rank, max_block_samples=FRAME_LEN, raw_u8=True)
        d.disable_callback()
        d.set_cuda_stream(work_stream.cuda_stream)
        ring8 = ResidentRing(d, iq8, FRAME_LEN, RING_PERIOD, raw_u8=True)
        for _ in range(LOCK_FRAMES + W):
            ring8.step()
        d.join()
        barrier()
        ev6, ev7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev6.record()
        for _ in range(K):
            ring8.step()
        d.join()
        ev7.record()
        barrier()
        ms8, v8 = aggregate(ev6.elapsed_time(ev7), samples_per_rank, world, dist)
        resident_u8 = {"value": round(v8, 1), "unit": "MSamples/s", "ms_per_step": round(ms8 / K, 4), "realtime_streams": round(v8 * 1e6 / FS, 1),
                       "algorithmic_bytes_per_sample": round((2 * FRAME_LEN + FRAME_BITS) / FRAME_LEN, 3),
                       "api": "dab_ofdm_attach_device_streams_raw (uint8 IQ resident in HBM, gain x16 so that the 8 bits are used) + dab_ofdm_advance_uniform",
                       "checked_streams": check_streams_against_oracle(d, ring8, MODE, [3, n_streams // 2 + 5], FRAME_LEN)}
        d.close()
        del d, iq8, ring8
        torch.cuda.empty_cache()
    except Exception as ex:  # noqa: BLE001
        resident_u8 = {"unavailable": repr(ex)}

    # ------------------------------------------------------------------ e2e: host buffers through dab_ofdm_process_batch
    def e2e_leg(u8):
        """K calls of the reference-facing batch entry point with pinned HOST blocks (one frame period per stream; fed repeatedly
        it is a continuous stream: the demodulator is differential per symbol and re-synchronises on every PRS, and the CFO
        phase is continuous across the block boundary by construction).  Every call uploads the blocks, demodulates and hands
        each completed frame's soft bits to the frame callback in host memory before it returns."""
        if u8:
            # examples/app_helpers/app_iq_readers.h:17-69 in reverse: what an 8-bit SDR front end delivers
            src = torch.view_as_real(iq[:, :FRAME_LEN])
            q = torch.clamp(src * 127.5 + 127.5, 0.0, 255.0).to(torch.uint8)
            host = torch.empty((n_streams, FRAME_LEN, 2), dtype=torch.uint8).pin_memory()
            host.copy_(q)
            del q, src
            bytes_per_sample = 2
        else:
            host = torch.empty((n_streams, FRAME_LEN), dtype=torch.complex64).pin_memory()
            host.copy_(iq[:, :FRAME_LEN])
            bytes_per_sample = 8
        torch.cuda.synchronize()
        d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=local_rank, max_block_samples=FRAME_LEN, raw_u8=u8)
        d.set_cuda_stream(work_stream.cuda_stream)
        counter = d.use_counting_callback()     # the library's own C callback: Python stays out of the delivery loop
        p, n = d.pointer_arrays([host[s].data_ptr() for s in range(n_streams)], [FRAME_LEN] * n_streams)
        for _ in range(LOCK_FRAMES + max(W, 3)):
            d.process_batch_prepared(p, n, u8)
        barrier()
        f0, b0 = int(counter.frames), int(counter.bits)
        t0 = time.perf_counter()
        for _ in range(K):
            d.process_batch_prepared(p, n, u8)   # returns after the soft bits of every completed frame were delivered
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt_max = aggregate(dt * 1e3, 0, world, dist)[0] * 1e-3
        frames, bits = int(counter.frames) - f0, int(counter.bits) - b0
        out = {"value": round(world * n_streams * FRAME_LEN * K / dt_max / 1e6, 1), "unit": "MSamples/s",
               "h2d_bytes_per_step": int(n_streams * FRAME_LEN * bytes_per_sample), "d2h_bytes_per_step": int(bits / K),
               "frames_delivered_per_step": frames / K, "ms_per_step": round(dt_max / K * 1e3, 3),
               "api": ("dab_ofdm_process_batch_u8 (pinned host uint8 IQ)" if u8 else "dab_ofdm_process_batch (pinned host complex64)") +
                      " + frame callback (dab_ofdm_count_frames_cb)",
               "realtime_streams": round(world * n_streams * FRAME_LEN * K / dt_max / FS, 1)}
        d.close()
        del host
        return out

    def full_chain_leg():
        """Raw 8-bit IQ in pinned host memory -> decoded FIBs + sub-channel bytes in pinned host memory, per call: upload,
        demodulate (dab_ofdm_process_batch_u8, soft bits stay in HBM), decode (dab_ensemble_decode_frames_device: FIC + 18 x EEP 3-A
        48 CU sub-channels = a full Mode I ensemble), download only the decoded bytes and CRC flags.  The synthetic IQ carries
        random payload, so the decoded bytes are not checked here (tests/test_ensemble_gpu.py::test_demod_to_bytes_on_device checks
        the chain against the oracle); the work per frame does not depend on the payload."""
        from cuda.bindings import runtime as cudart
        ens = importlib.import_module("dab-radio_b200.ensemble")
        src = torch.view_as_real(iq[:, :FRAME_LEN])
        host = torch.empty((n_streams, FRAME_LEN, 2), dtype=torch.uint8).pin_memory()
        host.copy_(torch.clamp(src * 127.5 + 127.5, 0.0, 255.0).to(torch.uint8))
        torch.cuda.synchronize()
        d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=local_rank, max_block_samples=FRAME_LEN, raw_u8=True)
        d.set_cuda_stream(work_stream.cuda_stream)
        d.disable_callback()
        dec = ens.EnsembleDecoder(1, n_streams=n_streams, device=local_rank, max_subchannels=18)
        dec.set_cuda_stream(work_stream.cuda_stream)
        decode_stream = torch.cuda.Stream()
        dec.set_decode_stream(decode_stream.cuda_stream)   # ingest on the demodulator's stream, decode + read-back beside the next upload
        dec.set_subchannels(-1, [ens.subchannel(48 * k, 48, 0, 0, 2, 0) for k in range(18)])
        res = dec.device_results()
        sizes = {"msc_bytes": n_streams * res.nb_cifs * res.msc_cif_bytes, "fib_bytes": n_streams * res.nb_cifs * res.fib_group_bytes,
                 "fib_valid": n_streams * res.nb_cifs * res.nb_fibs_per_cif, "msc_nbytes": n_streams * res.nb_cifs * res.max_subchannels * 4}
        h_out = {k: torch.empty(v, dtype=torch.uint8).pin_memory() for k, v in sizes.items()}
        p, n = d.pointer_arrays([host[s].data_ptr() for s in range(n_streams)], [FRAME_LEN] * n_streams)
        d_bits, n_bits, slots, d_fic = d.device_bits()

        def step():
            # frame k: upload + demodulate (asynchronous); meanwhile the decode of frame k - 1 and the read-back of its bytes finish on
            # the decode stream; then frame k is ingested (soft bits -> de-interleaver ring, on the demodulator's stream) and its decode
            # and read-back are queued behind that.  Every frame's bytes are in pinned host memory one step later; the timed region
            # ends with everything drained.
            d.process_batch_prepared(p, n, True)
            decode_stream.synchronize()
            d.join()
            for slot in range(slots):
                dec.decode_frames_device(d_bits + slot * n_bits, slots * n_bits, d_fic, slot)
            r = dec.device_results()
            for k in sizes:
                (err,) = cudart.cudaMemcpyAsync(h_out[k].data_ptr(), getattr(r, k), sizes[k], cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost,
                                                decode_stream.cuda_stream)
                assert int(err) == 0, err

        for _ in range(max(W, 5)):
            step()
        decode_stream.synchronize()
        work_stream.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        decode_stream.synchronize()
        work_stream.synchronize()
        dt = time.perf_counter() - t0
        dt_max = aggregate(dt * 1e3, 0, world, dist)[0] * 1e-3
        nbytes = h_out["msc_nbytes"].view(torch.int32)
        out = {"value": round(world * n_streams * FRAME_LEN * K / dt_max / 1e6, 1), "unit": "MSamples/s",
               "ensemble_frames_per_s": round(world * n_streams * K / dt_max, 1), "realtime_ensembles": round(world * n_streams * K / dt_max * 0.096, 1),
               "ms_per_step": round(dt_max / K * 1e3, 3), "h2d_bytes_per_step": int(n_streams * FRAME_LEN * 2),
               "d2h_bytes_per_step": int(sum(sizes.values())), "subchannel_cifs_decoded_last_step": int((nbytes > 0).sum().item()),
               "api": "dab_ofdm_process_batch_u8 (pinned host uint8 IQ) -> dab_ensemble_decode_frames_device with dab_ensemble_set_decode_stream "
                      "(the decode of frame k runs beside the upload of frame k + 1) -> decoded bytes to pinned host memory"}
        dec.close()
        d.close()
        del host
        return out

    def pcie_ceiling(bytes_per_sample):
        """What the host link gives a bare pinned copy of the same size IN THIS RUN (every rank copying at once, as in the e2e leg):
        one step's H2D block and one step's D2H soft bits on two CUDA streams, K times.  The e2e leg cannot be faster than this."""
        h_in = torch.empty(n_streams * FRAME_LEN * bytes_per_sample, dtype=torch.uint8).pin_memory()
        d_in = torch.empty_like(h_in, device="cuda")
        d_out = torch.empty(n_streams * FRAME_BITS, dtype=torch.int8, device="cuda")
        h_out = torch.empty(n_streams * FRAME_BITS, dtype=torch.int8).pin_memory()
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        def once():
            with torch.cuda.stream(s_up):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_down):
                h_out.copy_(d_out, non_blocking=True)
        for _ in range(2):
            once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(K):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt_max = aggregate(dt * 1e3, 0, world, dist)[0] * 1e-3
        return {"ms_per_step": round(dt_max / K * 1e3, 3), "h2d_gbs_per_gpu": round(h_in.numel() * K / dt_max / 1e9, 1),
                "d2h_gbs_per_gpu": round(h_out.numel() * K / dt_max / 1e9, 1),
                "note": "bare pinned cudaMemcpyAsync of one step's input and output, both directions at once, all ranks at once, max over ranks"}

    e2e = None
    if not args.no_e2e:
        e2e = e2e_leg(False)
        try:
            ceil = pcie_ceiling(8)
            ceil["e2e_fraction_of_ceiling"] = round(ceil["ms_per_step"] / e2e["ms_per_step"], 3)
            e2e["pcie_ceiling"] = ceil
        except Exception as ex:  # noqa: BLE001
            e2e["pcie_ceiling"] = {"unavailable": repr(ex)}
        e2e["raw_u8_ingest"] = e2e_leg(True)   # raw 8-bit IQ uploaded and dequantised on the device
        try:
            ceil = pcie_ceiling(2)
            ceil["e2e_fraction_of_ceiling"] = round(ceil["ms_per_step"] / e2e["raw_u8_ingest"]["ms_per_step"], 3)
            e2e["raw_u8_ingest"]["pcie_ceiling"] = ceil
        except Exception as ex:  # noqa: BLE001
            e2e["raw_u8_ingest"]["pcie_ceiling"] = {"unavailable": repr(ex)}
        try:   # SURVEY 8(f) ranks 1-3 chained: IQ in, FIBs + sub-channel bytes out (a secondary line must not take the headline down)
            e2e["raw_u8_to_decoded_bytes"] = full_chain_leg()
        except Exception as ex:  # noqa: BLE001
            e2e["raw_u8_to_decoded_bytes"] = {"unavailable": repr(ex)}

    # ------------------------------------------------------------------ single stream through the per-object call (rank 0, N = 1)
    single = None
    if rank == 0 and world == 1 and not args.no_e2e:
        single = single_stream_leg(ofdm, iq, local_rank)

    # ------------------------------------------------------------------ Viterbi (secondary line; rank 0, N = 1 only)
    viterbi = None
    modes = None
    impaired = None
    if rank == 0 and world == 1 and not args.no_viterbi:
        del iq
        torch.cuda.empty_cache()
        try:
            impaired = impairments_leg(torch, ofdm, n_streams, 10, local_rank)
        except Exception as ex:  # noqa: BLE001
            impaired = {"unavailable": repr(ex)}
        torch.cuda.empty_cache()
        viterbi = viterbi_leg(torch, pkg, n_streams, 5, not args.no_cpu)
        try:
            viterbi["ensemble"] = ensemble_leg(torch, pkg, n_streams, 5, not args.no_cpu)
        except Exception as ex:  # noqa: BLE001
            viterbi["ensemble"] = {"unavailable": repr(ex)}
        torch.cuda.empty_cache()
        try:
            modes = modes_leg(torch, pkg, n_streams, 6)
        except Exception as ex:  # noqa: BLE001
            modes = {"unavailable": repr(ex)}

    # ------------------------------------------------------------------ cpu baseline (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, sample_frames=args.cpu_frames)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "MSamples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": round(ms_max / K, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_text(n_streams),
                       "streams_per_gpu": n_streams, "samples_per_step_per_gpu": n_streams * FRAME_LEN, "block_samples": FRAME_LEN,
                       "l2": "inputs (1.6 GB per step) are larger than L2 and read once; no flush needed",
                       "resident_buffer": f"ring of {RING_PERIOD + 3} frames per stream, content periodic over {RING_PERIOD} frames, rebased every {RING_PERIOD} steps",
                       "snr_db": 25, "cfo": "+-50 kHz per stream", "acquisition": f"{LOCK_FRAMES} untimed frames per stream before the warm-up (streams lock)", "parallelism": f"streams sharded, {world} rank(s), no data-path collective"},
            "realtime_streams": round(value * 1e6 / FS, 1), "realtime_streams_per_gpu": round(value * 1e6 / FS / world, 1),
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "viterbi": viterbi, "modes": modes,
            "frames_per_stream_in_timed_region": frames_per_stream, "locked_streams": locked, "locked_fraction": round(locked / max(1, len(states)), 4),
            "parity": parity, "sustained": sustained, "single_stream": single, "impairments": impaired, "resident_raw_u8": resident_u8,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def single_stream_leg(ofdm, iq, device):
    """The drop-in class's own call pattern: ONE stream, dab_ofdm_process (what OFDM_Demod::Process of the mirror class issues) in
    65 536-sample blocks from pageable host memory, frames delivered through the callback before each call returns
    (examples/app_helpers/app_ofdm_blocks.h:45-58, examples/basic_radio_app.cpp:78).  Launch latency bound: a real-time factor."""
    x = iq[0].cpu().numpy()
    block = 65536
    d = ofdm.OfdmDemodBatch(MODE, n_streams=1, device=device, max_block_samples=block)
    counter = d.use_counting_callback()
    n_blocks = x.size // block
    for b in range(min(n_blocks, 12)):      # lock
        d.process(0, x[b * block:(b + 1) * block])
    f0 = int(counter.frames)
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 1.0:
        for b in range(n_blocks):
            d.process(0, x[b * block:(b + 1) * block])
        reps += 1
    dt = time.perf_counter() - t0
    samples = reps * n_blocks * block
    out = {"value": round(samples / dt / 1e6, 2), "unit": "MSamples/s", "realtime_factor": round(samples / dt / FS, 1), "block_samples": block,
           "ms_per_call": round(dt / (reps * n_blocks) * 1e3, 3), "frames_delivered": int(counter.frames) - f0, "seconds": round(dt, 2),
           "api": "dab_ofdm_process, one stream per handle (the OFDM_Demod mirror class's path), pageable host blocks, callback per frame"}
    d.close()
    return out


def impairments_leg(torch, ofdm, n_streams, steps, device):
    """BASELINE.json configs[2] as a throughput line: the headline workload under a three-tap multipath channel (delays up to half the
    cyclic prefix, -3 / -9 dB), +-50 kHz CFO and 12 dB SNR.  Two streams are replayed through the oracle (parity under impairments
    is what tests/test_ofdm_gpu.py::test_impairments_match_oracle covers case by case, incl. sample-rate drift)."""
    iq, _ = build_streams_on_device(torch, n_streams, RING_PERIOD + 3, seed=77, period=RING_PERIOD)
    # multipath: y[n] = x[n] + g1 x[n - d1] + g2 x[n - d2] (circular over the periodic ring content, so the ring stays periodic)
    per = RING_PERIOD * FRAME_LEN
    base = iq[:, :per]
    y = base.clone()
    for delay, gain_db, phase in ((100, -3.0, 2.0), (230, -9.0, -1.0)):
        g = 10.0 ** (gain_db / 20.0) * np.exp(1j * phase)
        y += torch.roll(base, delay, dims=1) * complex(g)
    gen = torch.Generator(device="cuda")
    gen.manual_seed(5)
    sigma = float(torch.sqrt(torch.mean(torch.abs(y[:8]) ** 2)).item()) * 10 ** (-12.0 / 20) / np.sqrt(2.0)
    y += torch.view_as_complex(torch.randn((n_streams, per, 2), device="cuda", generator=gen) * sigma)
    iq[:, :per] = y
    iq[:, per:] = y[:, :iq.shape[1] - per]
    del y, base
    d = ofdm.OfdmDemodBatch(MODE, n_streams=n_streams, device=device, max_block_samples=FRAME_LEN)
    d.disable_callback()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        d.set_cuda_stream(st.cuda_stream)
        ring = ResidentRing(d, iq, FRAME_LEN, RING_PERIOD)
        for _ in range(LOCK_FRAMES + 3):
            ring.step()
        d.join()
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            ring.step()
        d.join()
        e1.record()
        st.synchronize()
    ms = e0.elapsed_time(e1) / steps
    locked = sum(1 for s in range(n_streams) if d.state(s)["state"] != 0)
    desync = sum(d.state(s)["total_frames_desync"] for s in range(0, n_streams, max(1, n_streams // 64)))
    checked = check_streams_against_oracle(d, ring, MODE, [1, n_streams // 2 + 1], FRAME_LEN)
    out = {"value": round(n_streams * FRAME_LEN / ms / 1e3, 1), "unit": "MSamples/s", "ms_per_step": round(ms, 4), "streams": n_streams,
           "channel": "3 taps (0 / 100 / 230 samples, 0 / -3 / -9 dB), +-50 kHz CFO, 12 dB SNR", "locked_fraction": round(locked / n_streams, 4),
           "desyncs_in_64_sampled_streams": int(desync), "streams_checked_against_oracle": checked}
    d.close()
    return out


def viterbi_leg(torch, pkg, n_streams, reps, with_cpu):
    """Secondary line (BASELINE.json configs[3]): full-ensemble Viterbi decode.  Per stream and transmission frame: the FIC
    (4 groups: PI_16 x 21 blocks, PI_15 x 3, PI_X; fic_decoder.cpp:74-87) plus the whole MSC as 18 DAB+ sub-channels of 48 CU at
    EEP 3-A (PI_8 x 45 blocks, PI_7 x 3, PI_X; subchannel_protection_tables.h:121-126) x 4 CIFs = 76 trellises over exactly the
    230 400 soft bits one demodulated frame holds.  Soft bits, jobs and outputs are device resident; CUDA-event timing."""
    from oracle import pyoracle as po
    vit = importlib.import_module("dab-radio_b200.viterbi")
    rng = np.random.default_rng(99)
    fic = [(po.puncture_code(16), 128 * 21), (po.puncture_code(15), 128 * 3), (po.PI_X, 24)]
    eep = [(po.puncture_code(8), 128 * 45), (po.puncture_code(7), 128 * 3), (po.PI_X, 24)]
    layout = [(fic, 96, 2304)] * 4 + [(eep, 192, 3072)] * 72
    base = []
    for segs, nbytes, nsoft in layout:
        tx = po.puncture(po.conv_encode(rng.integers(0, 256, nbytes, dtype=np.uint8)), segs)
        assert tx.size == nsoft
        base.append(np.clip(np.rint(tx + 60.0 * rng.standard_normal(tx.size)), -128, 127).astype(np.int8))
    base = np.concatenate(base)
    frame_soft, frame_out = base.size, sum(nb for _, nb, _ in layout)
    vb = vit.ViterbiBatch(torch.cuda.current_device())
    vb.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
    sid_fic = vb.add_schedule(vit.make_schedule(fic, 96))
    sid_eep = vb.add_schedule(vit.make_schedule(eep, 192))
    jobs = np.zeros(n_streams * len(layout), pkg.capi.VIT_JOB_DTYPE)
    # job order: trellis slot major, stream minor -- the 32 trellises a warp of the bulk kernel runs in lock step share a
    # schedule; jobs[slot * n_streams + s] is trellis `slot` of stream s (buffers stay [stream][slot])
    so_slot = np.concatenate([[0], np.cumsum([ns for _, _, ns in layout])[:-1]])
    oo_slot = np.concatenate([[0], np.cumsum([nb for _, nb, _ in layout])[:-1]])
    for i, (segs, nbytes, nsoft) in enumerate(layout):
        for s in range(n_streams):
            jobs[i * n_streams + s] = (sid_fic if nbytes == 96 else sid_eep, nsoft, s * frame_soft + int(so_slot[i]), s * frame_out + int(oo_slot[i]))
    d_base = torch.from_numpy(base).cuda()
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    # every stream: the same code words under its own extra noise; stream 0 keeps the CPU-generated soft bits for the check
    d_soft = (d_base.view(1, -1).to(torch.int16) + torch.randint(-24, 25, (n_streams, frame_soft), device="cuda", generator=g, dtype=torch.int16))
    d_soft = d_soft.clamp_(-128, 127).to(torch.int8).contiguous()
    d_soft[0] = d_base
    d_out = torch.zeros(n_streams * frame_out, dtype=torch.uint8, device="cuda")
    d_err = torch.zeros(jobs.size, dtype=torch.int64, device="cuda")
    d_st = torch.zeros(jobs.size, dtype=torch.int32, device="cuda")
    # the job list of an ensemble configuration is prepared once (ordered by schedule, short FIC trellises packed two to a warp)
    plan = vb.prepare_jobs(jobs)
    run = lambda: vb.decode_prepared(plan, d_soft.data_ptr(), d_soft.numel(), d_out.data_ptr(), d_out.numel(), d_err.data_ptr(), d_st.data_ptr())
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    assert int(d_st.abs().max().item()) == 0
    # stream 0 against the oracle: bit-exact bytes and path error for all 76 trellises
    out0 = d_out[:frame_out].cpu().numpy()
    err0 = d_err.view(len(layout), n_streams)[:, 0].cpu().numpy()
    ov = po.OracleViterbi()
    so = oo = 0
    for i, (segs, nbytes, nsoft) in enumerate(layout):
        ov.set_traceback_length(nbytes * 8)
        ref_out, ref_err, _ = ov.decode_job(base[so:so + nsoft], segs, nbytes)
        assert np.array_equal(out0[oo:oo + nbytes], ref_out) and int(err0[i]) == ref_err, f"viterbi job {i} differs from the oracle"
        so += nsoft
        oo += nbytes
    steps = n_streams * (4 * 774 + 72 * 1542)
    bits = n_streams * frame_out * 8
    res = {"metric": "full-ensemble Viterbi decode (FIC + 18 x EEP 3-A 48 CU x 4 CIF per Mode I frame)", "value": round(bits / ms / 1e3, 1),
           "unit": "decoded Mbit/s", "ms_per_launch": round(ms, 4), "trellises": int(jobs.size), "streams": n_streams,
           "trellis_steps_per_s": round(steps / ms * 1e3, 0), "ensemble_frames_per_s": round(n_streams / ms * 1e3, 1),
           "realtime_ensembles": round(n_streams / ms * 1e3 * 0.096, 1), "bit_exact_vs_oracle": f"{len(layout)} trellises of stream 0",
           "kernel": "viterbi_lanes_kernel (one trellis per thread; batches below 4096 trellises take viterbi_kernel, one per warp)",
           "bound": "integer issue (fma-heavy + alu pipes); HBM traffic = soft bits in + bits/8 out + 8 B per trellis step of survivor "
                    "decisions written once and read once by the traceback"}
    if with_cpu:
        try:
            from oracle import pyref
            import ctypes as C
            L = pyref.fast_lib()
            cores = os.cpu_count() or 1
            n_jobs = cores * 400
            soft = np.tile(base[4 * 2304:4 * 2304 + 3072], n_jobs)
            codes, lens, nout = po.pack_segments(eep)
            out = np.zeros(n_jobs * 192, np.uint8)
            P = lambda a: a.ctypes.data_as(C.c_void_p)
            dt, rounds = 0.0, 0
            L.ref_vit_bench(cores, P(soft), 3072, n_jobs, P(codes), P(lens), P(nout), len(eep), 1536, P(out), 192)   # warm-up
            while dt < 2.0:
                dt += float(L.ref_vit_bench(cores, P(soft), 3072, n_jobs, P(codes), P(lens), P(nout), len(eep), 1536, P(out), 192))
                rounds += 1
            res["cpu_baseline"] = {"value": round(rounds * n_jobs * 1536 / dt / 1e6, 1), "unit": "decoded Mbit/s", "cores": cores, "kind": "reference",
                                   "sample": f"{rounds} x {n_jobs} EEP 3-A 48 CU trellises, DAB_Viterbi_Decoder (AVX2 u16), one decoder per core, {dt:.2f} s"}
        except (FileNotFoundError, OSError, AttributeError) as e:  # noqa: PERF203
            res["cpu_baseline"] = {"unavailable": repr(e)}
    vb.close()
    return res


def ensemble_leg(torch, pkg, n_streams, reps, with_cpu):
    """SURVEY 8(f) rows 2-3 measured: one OFDM frame of soft bits per stream, resident in HBM, through dab_ensemble_decode_frames_device
    (CIF time de-interleave ring -> FIC + 18 x EEP 3-A Viterbi -> energy dispersal -> FIB CRC16), i.e. everything between the
    demodulator's output and the decoded bytes.  Stream 0 is checked against the oracle's FIC_Decoder / MSC_Decoder."""
    import ensgen
    from oracle import pyoracle as po
    ens = importlib.import_module("dab-radio_b200.ensemble")
    layout = [(48 * k, 48, 0, 0, 2, 0) for k in range(18)]     # 18 DAB+ sub-channels, EEP 3-A, 48 CU each = the whole CIF
    subs_o = [po.subchannel(*a) for a in layout]
    n_frames = 5                                               # 20 CIFs: the 16-CIF de-interleaver is full from frame 4 on
    tx = ensgen.EnsembleTx(1, subs_o, seed=77, sigma=45.0)
    frames = np.stack([tx.next_frame() for _ in range(n_frames)])
    want = ensgen.oracle_decode_stream(1, subs_o, frames)
    d_base = torch.from_numpy(frames).cuda()
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    # every stream: the same code words under its own extra noise; stream 0 keeps the CPU-generated frames for the check
    d_frames = d_base.view(n_frames, 1, -1).to(torch.int16) + torch.randint(-20, 21, (n_frames, n_streams, frames.shape[1]), device="cuda",
                                                                            generator=g, dtype=torch.int16)
    d_frames = d_frames.clamp_(-127, 127).to(torch.int8).contiguous()
    d_frames[:, 0] = d_base
    dec = ens.EnsembleDecoder(1, n_streams=n_streams, device=torch.cuda.current_device(), max_subchannels=18)
    dec.set_cuda_stream(torch.cuda.current_stream().cuda_stream)
    dec.set_subchannels(-1, [ens.subchannel(*a) for a in layout])
    frame_bits = frames.shape[1]
    run = lambda f: dec.decode_frames_device(d_frames[f].data_ptr(), frame_bits, None, 0)
    for f in range(n_frames):
        run(f)
    dec.sync()
    # stream 0 after the last distinct frame: FIBs, CRC flags, sub-channel bytes and path errors equal the oracle's
    fb, fv, fe, msc = want[-1]
    got_b, got_v, got_e = dec.read_fic(0)
    for c in range(4):
        assert np.array_equal(got_b[c], fb[c]) and np.array_equal(got_v[c], fv[c]) and int(got_e[c]) == fe[c], f"ensemble FIC cif {c} differs from the oracle"
        for k in range(len(layout)):
            b, n, e = dec.read_msc(0, c, k)
            wb, we = msc[c][k]
            assert n == wb.size == 192 and np.array_equal(b, wb) and e == we, f"ensemble MSC cif {c} sub {k} differs from the oracle"
    launches0 = dec.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(reps):
        run(i % n_frames)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    trellises, steps = dec.last_work()
    launches = (dec.kernel_launches() - launches0) // reps
    bits_out = n_streams * (4 * 768 + 4 * 18 * 1536)
    res = {"metric": "ensemble decode: soft-bit frame -> FIBs + sub-channel bytes (CIF de-interleave, FIC + 18 x EEP 3-A 48 CU x 4 CIF Viterbi, "
                     "energy dispersal, FIB CRC16), Mode I, soft bits resident in HBM",
           "value": round(n_streams / ms * 1e3, 1), "unit": "ensemble frames/s", "ms_per_call": round(ms, 4), "streams": n_streams,
           "realtime_ensembles": round(n_streams / ms * 1e3 * 0.096, 1), "decoded_mbit_s": round(bits_out / ms / 1e3, 1),
           "trellises_per_call": trellises, "trellis_steps_per_s": round(steps / ms * 1e3, 0), "kernel_launches_per_call": int(launches),
           "bit_exact_vs_oracle": "stream 0: 4 FIB groups + 72 sub-channel CIFs (bytes, CRC flags, path errors)",
           "api": "dab_ensemble_decode_frames_device"}
    if with_cpu:
        try:
            from concurrent.futures import ThreadPoolExecutor
            from oracle import pyref
            cores = os.cpu_count() or 1
            n_rep = 40

            def worker(_):
                fic = pyref.RefFicDecoder(2304, 3)
                decs = [pyref.RefMscDecoder(*a) for a in layout]
                for _ in range(n_rep):
                    for fr in frames:
                        for c in range(4):
                            fic.decode_group(fr[c * 2304:(c + 1) * 2304], c)
                            cif = fr[9216 + c * 55296: 9216 + (c + 1) * 55296]
                            for d in decs:
                                d.decode_cif(cif)
            t0 = time.perf_counter()
            with ThreadPoolExecutor(cores) as ex:
                list(ex.map(worker, range(cores)))
            dt = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": round(cores * n_rep * n_frames / dt, 1), "unit": "ensemble frames/s", "cores": cores, "kind": "reference",
                                   "sample": f"{cores} threads x {n_rep * n_frames} frames through the reference's FIC_Decoder + 18 MSC_Decoder "
                                             f"(AVX2 Viterbi) via ctypes, {dt:.1f} s"}
        except (FileNotFoundError, OSError, AttributeError) as e:  # noqa: PERF203
            res["cpu_baseline"] = {"unavailable": repr(e)}
    dec.close()
    return res


MODE_FRAME_LEN = {1: 196608, 2: 49152, 3: 49152, 4: 98304}   # dab_ofdm_params_ref.cpp:13-52
# SURVEY 8(d) algorithmic bytes per frame: 8 (null + S Tsym) + 2 (S - 1) Ncarr
MODE_ALGO_BYTES = {1: 1803264, 2: 450816, 3: 451584, 4: 901632}


def modes_leg(torch, pkg, n_streams, frame_periods):
    """BASELINE.json configs[4]: Modes II / III / IV (512 / 256 / 1024-point FFT), n_streams resident streams each, fed in
    4096-sample blocks (the size at which the reference itself locks in Modes II / III) and in whole frames, with
    the reference's default OFDM_Demod_Config and -- Modes II / III -- with signal_l1 = 25-sample windows, every second one (the
    reference's own knob; under the default 100-sample windows its power-dip detector misses the 664 / 345-sample NULL symbol for
    most start offsets, in the reference as here).  Every line carries the fraction of streams that locked, the same for 6 sample
    streams in the oracle (and their parity: equal frame / desync counts and state, soft bits of the last frame), and the
    fraction of the HBM roofline.  Then a mixed-mode batch: four handles (Modes I-IV, n_streams / 4 streams each) advancing
    concurrently on their own CUDA streams, 96 ms of air time per stream per round."""
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    W, period = 4, 8
    peak, _ = measured_peaks()
    res = {}

    def make(mode, n, block, l1_config):
        fl = MODE_FRAME_LEN[mode]
        iq, _ = build_streams_on_device(torch, n, period + 2 + max(1, block // fl), seed=4321 + mode, mode=mode, frame_len=fl, period=period)
        d = ofdm.OfdmDemodBatch(mode, n_streams=n, device=torch.cuda.current_device(), max_block_samples=block)
        d.disable_callback()
        if l1_config:
            cfg = d.get_config(0)
            cfg.signal_l1_nb_samples, cfg.signal_l1_nb_decimate = l1_config
            d.set_config(cfg)
        return d, iq

    def frames_read(d, n):
        return sum(d.state(s)["total_frames_read"] for s in range(0, n, max(1, n // 16)))

    for mode in (2, 3, 4):
        fl = MODE_FRAME_LEN[mode]
        for block in (4096, fl):
            for l1_config in ((None, (25, 2)) if mode in (2, 3) else (None,)):
                d, iq = make(mode, n_streams, block, l1_config)
                stream = torch.cuda.Stream()
                per_period = fl // block
                with torch.cuda.stream(stream):
                    d.set_cuda_stream(stream.cuda_stream)
                    ring = ResidentRing(d, iq, fl, period)
                    for _ in range(W * per_period):
                        ring.step(block)
                    d.join()
                    stream.synchronize()
                    f0 = frames_read(d, n_streams)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    for _ in range(frame_periods * per_period):
                        ring.step(block)
                    d.join()
                    e1.record()
                    stream.synchronize()
                ms = e0.elapsed_time(e1) / frame_periods
                f1 = frames_read(d, n_streams)
                n_probe = len(range(0, n_streams, max(1, n_streams // 16)))
                frames_per_period = (f1 - f0) / n_probe / frame_periods
                locked = sum(1 for s in range(n_streams) if d.state(s)["state"] != 0)
                sample = [k * n_streams // 6 + 3 for k in range(6)] if n_streams >= 64 else list(range(min(6, n_streams)))
                checked = check_streams_against_oracle(d, ring, mode, sample, block, l1_config)
                gbs = MODE_ALGO_BYTES[mode] * n_streams * frames_per_period / (ms * 1e-3) / 1e9
                key = f"mode_{mode}_block_{'frame' if block == fl else block}_{'l1_25x2' if l1_config else 'default_config'}"
                res[key] = {"value": round(n_streams * fl / ms / 1e3, 1), "unit": "MSamples/s", "ms_per_frame_period": round(ms, 4), "streams": n_streams,
                            "block_samples": block, "frames_per_stream_per_frame_period": round(frames_per_period, 3),
                            "locked_fraction": round(locked / n_streams, 4),
                            "oracle_locked_fraction_of_sample": round(sum(c["locked"] for c in checked) / len(checked), 3),
                            "sample_streams_match_oracle": len(checked),
                            "roofline_frac": round(gbs / peak, 4), "achieved_gbs": round(gbs, 1)}
                d.close()
                del d, iq, ring
                torch.cuda.empty_cache()

    # mixed-mode batch
    block = 196608
    n_each = max(1, n_streams // 4)
    handles = []
    for mode in (1, 2, 3, 4):
        d, iq = make(mode, n_each, block, (25, 2) if mode in (2, 3) else None)
        st = torch.cuda.Stream()
        d.set_cuda_stream(st.cuda_stream)
        handles.append((mode, d, ResidentRing(d, iq, MODE_FRAME_LEN[mode], period), st))
    for _ in range(W):
        for _, _, ring, _ in handles:
            ring.step(block)
    for _, d, _, st in handles:
        d.join()
        st.synchronize()
    f0 = {m: frames_read(d, n_each) for m, d, _, _ in handles}
    main = torch.cuda.current_stream()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _, _, _, st in handles:
        st.wait_stream(main)
    for _ in range(frame_periods):
        for _, _, ring, _ in handles:
            ring.step(block)
    for _, d, _, st in handles:
        d.join()
        main.wait_stream(st)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / frame_periods
    n_probe = len(range(0, n_each, max(1, n_each // 16)))
    res["mixed"] = {"value": round(4 * n_each * block / ms / 1e3, 1), "unit": "MSamples/s", "ms_per_round": round(ms, 4),
                    "streams_per_mode": n_each, "samples_per_stream_per_round": block,
                    "config": "Modes II / III with signal_l1 = 25 x 2, Modes I / IV default",
                    "locked_fraction": {f"mode_{m}": round(sum(1 for s in range(n_each) if d.state(s)["state"] != 0) / n_each, 4) for m, d, _, _ in handles},
                    "frames_per_stream_per_round": {f"mode_{m}": round((frames_read(d, n_each) - f0[m]) / n_probe / frame_periods, 3) for m, d, _, _ in handles}}
    for _, d, _, _ in handles:
        d.close()
    return res


REF_STREAM_FRAMES = 4      # frames per reference instance: the content repeats after them (continuous CFO phase), as in the GPU ring


def host_streams(n, n_frames, seed):
    """n synthetic Mode I streams on the host, generated like build_streams_on_device: frames from a modulated pool, per-stream
    start offset in [0, FRAME_LEN), CFO within +-50 kHz in multiples of Fs / FRAME_LEN (so that repeating the buffer is a continuous
    signal), AWGN at 25 dB."""
    import dabgen
    from oracle import pyoracle as po
    rng = np.random.default_rng(seed)
    pool = [po.modulate(MODE, rng.integers(0, 256, dabgen.payload_bytes(MODE), dtype=np.uint8)) * np.float32(4.0 / 1536) for _ in range(4)]
    sig_pow = float(np.mean(np.abs(pool[0]) ** 2))
    sigma = np.sqrt(sig_pow / 10 ** 2.5 / 2)
    total = n_frames * FRAME_LEN
    t = np.arange(total, dtype=np.float64)
    out = []
    for _ in range(n):
        seq = np.concatenate([pool[k] for k in rng.integers(0, len(pool), n_frames)])
        x = np.roll(seq, -int(rng.integers(0, FRAME_LEN)))
        cfo = int(rng.integers(-4800, 4801)) / FRAME_LEN
        x = x * np.exp(2j * np.pi * np.remainder(cfo * t, 1.0))
        x = x + sigma * (rng.standard_normal(total) + 1j * rng.standard_normal(total))
        out.append(np.ascontiguousarray(x, np.complex64))
    return out


def reference_runner(cores, threads_each=1, n_instances=None):
    """(run(reps) -> seconds, kind, note, close): `reps` passes of REF_STREAM_FRAMES frames through every instance of the
    reference's OFDM_Demod (oracle/_ref, compiled from its own sources) -- or the oracle port where that library is missing -- one
    instance per core, each on its own stream, FRAME_LEN-sample Process() calls like the GPU arm."""
    n_instances = cores if n_instances is None else n_instances
    xs = host_streams(n_instances, REF_STREAM_FRAMES, seed=7)
    try:
        from oracle import pyref
        pool = pyref.RefOfdmPool(MODE, n_instances, threads_each, fast=True)
        note = ("reference sources compiled unmodified (oracle/_ref/libdabref_fast.so, -O3 -march=x86-64-v3 -ffast-math, AVX2 PLL); "
                "FFTW3 absent -> single-precision radix-4 Stockham stand-in FFT (AVX2 + FMA intrinsics)")
        return (lambda reps: pool.run_multi(xs, FRAME_LEN, reps)), "reference", note, pool
    except (FileNotFoundError, OSError, AttributeError):
        from oracle import pyoracle as po
        import ctypes as C

        def run(reps):
            frames = C.c_uint64()
            return float(po.lib().orc_ofdm_bench(MODE, n_instances, xs[0].ctypes.data_as(C.c_void_p), xs[0].size, FRAME_LEN, reps, C.byref(frames)))
        return run, "port", "oracle/dab_oracle.c (scalar C restatement, double-precision FFT; every instance on stream 0)", None


def fft_share(pool, msamples_per_s_per_instance):
    """How much of the reference's CPU time per frame is the FFT stand-in: 76 symbol transforms + the NULL slot + 5 synchronisation
    transforms per Mode I frame (ofdm_demodulator.cpp:360-548, 705, 891-894).  FFTW3's codelets would shrink exactly this share."""
    if pool is None:
        return None
    reps = 20000
    t_fft = pool.fft_seconds(2048, reps) / reps
    t_frame = FRAME_LEN / (msamples_per_s_per_instance * 1e6)
    share = min(1.0, 82 * t_fft / t_frame)
    return {"fft_us_per_2048_transform": round(t_fft * 1e6, 2), "transforms_per_frame": 82, "share_of_cpu_time": round(share, 3),
            "note": "time of the FFT stand-in linked in place of FFTW3 (absent from the image); with an FFT k times faster the CPU "
                    "arm would run 1 / (1 - share + share / k) times faster (k = 3: x%.2f, k = inf: x%.2f)" %
                    (1.0 / (1.0 - share + share / 3.0), 1.0 / max(1e-9, 1.0 - share))}


def cpu_baseline(args, sample_frames):
    """The reference's own OFDM_Demod (oracle/_ref, timing build) on the host cores: one instance + feeder thread per core,
    nb_desired_threads = 1 each, bounded sample; plus one instance using every core
    (nb_desired_threads = 0, ofdm_demodulator.cpp:148-183)."""
    cores = os.cpu_count() or 1
    run, kind, note, pool = reference_runner(cores)
    reps = max(1, sample_frames // REF_STREAM_FRAMES)
    run(1)
    dt = run(reps)
    v = cores * reps * REF_STREAM_FRAMES * FRAME_LEN / dt / 1e6
    out = {"value": round(v, 1), "unit": "MSamples/s", "cores": cores, "kind": kind,
           "sample": f"{cores} instances (each its own stream: start offset, +-50 kHz CFO, 25 dB) x {reps * REF_STREAM_FRAMES} Mode I frames "
                     f"({cores * reps * REF_STREAM_FRAMES * FRAME_LEN / 1e6:.0f} MSamples), {FRAME_LEN}-sample Process() calls, {dt:.1f} s",
           "note": note, "realtime_streams": round(v * 1e6 / FS, 1), "fft": fft_share(pool, v / cores)}
    if pool is not None:
        pool.close()
        try:   # one demodulator with the reference's own symbol-level thread pool on every core
            run1, _, _, pool1 = reference_runner(cores, threads_each=0, n_instances=1)
            run1(1)
            reps1 = max(1, reps // 2)
            dt1 = run1(reps1)
            v1 = reps1 * REF_STREAM_FRAMES * FRAME_LEN / dt1 / 1e6
            out["single_instance_all_threads"] = {"value": round(v1, 1), "unit": "MSamples/s", "nb_desired_threads": 0, "threads": cores,
                                                  "realtime_factor": round(v1 * 1e6 / FS, 1),
                                                  "sample": f"1 instance x {reps1 * REF_STREAM_FRAMES} frames, {dt1:.1f} s"}
            pool1.close()
        except Exception as ex:  # noqa: BLE001
            out["single_instance_all_threads"] = {"unavailable": repr(ex)}
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads, same metric/config: one OFDM_Demod per
    core, every instance on its own synthetic stream (start offset, +-50 kHz CFO, 25 dB SNR), one-frame Process() calls."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    run, kind, note, pool = reference_runner(cores)
    reps = max(1, args.ref_frames // REF_STREAM_FRAMES)
    frames_per_step = reps * REF_STREAM_FRAMES
    for _ in range(max(1, args.warmup)):
        run(1)
    t = 0.0
    for _ in range(args.steps):
        t += run(reps)
    samples = cores * frames_per_step * FRAME_LEN * args.steps
    v = samples / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(v, 1), "unit": "MSamples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(t / args.steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_text(args.streams),
                   "streams_per_gpu": args.streams, "samples_per_step_per_gpu": args.streams * FRAME_LEN, "block_samples": FRAME_LEN,
                   "snr_db": 25, "cfo": "+-50 kHz per stream",
                   "reference_sample": f"bounded sample of that workload on the host CPU: {cores} OFDM_Demod instances (one per core, each its own "
                                       f"stream), {frames_per_step} frames per instance per step"},
        "realtime_streams": round(v * 1e6 / FS, 1),
        "cpu_baseline": {"value": round(v, 1), "unit": "MSamples/s", "cores": cores, "kind": kind, "note": note,
                         "sample": f"{cores} instances x {frames_per_step} frames per step x {args.steps} steps, {FRAME_LEN}-sample Process() calls",
                         "fft": fft_share(pool, v / cores)},
        "e2e": {"value": round(v, 1), "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if pool is not None:
        pool.close()
    print(json.dumps(line), flush=True)


def workload_text(n_streams):
    return (f"DAB Mode I, {n_streams} independent synthetic streams per GPU, OFDM demod IQ->int8 soft bits "
            f"(BASELINE.json configs[1]); one step = one 196608-sample frame per stream")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=N_STREAMS)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-viterbi", action="store_true")
    ap.add_argument("--no-sustained", action="store_true")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained leg")
    ap.add_argument("--cpu-frames", type=int, default=2000, help="frames per instance in the cpu_baseline sample")
    ap.add_argument("--ref-frames", type=int, default=24, help="frames per instance per step for --impl reference")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
