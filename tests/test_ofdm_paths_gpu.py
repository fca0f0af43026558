"""GPU parity of the demodulator paths test_ofdm_gpu.py does not reach: the batched device-resident path with the pipeline ways
active and ragged per-stream blocks (what bench.py times), the generic-geometry frame kernel (DAB modes forced through it, and a
non-DAB OFDM_Params geometry), the GUI taps, the raw integer IQ formats, and the host-side contracts of the C ABI (callback
re-entrancy, snapshot getters).  Oracle = oracle/dab_oracle.c, pinned against the reference (tests/test_oracle_vs_reference.py)."""
import ctypes as C
import importlib

import numpy as np
import pytest

import dabgen
from test_ofdm_gpu import FREQ_TOL_BINS, LSB1_MIN, STREAM_CASES, _assert_stream_parity

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ofdm(pkg):
    return importlib.import_module("dab-radio_b200.ofdm")


def _torch():
    import torch
    return torch


# ---------------------------------------------------------------------------------------------------------------------------
# the benchmarked path: dab_ofdm_attach_device_streams + dab_ofdm_advance with the pipeline ways active
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("uniform", [True, False])
def test_resident_streams_with_pipeline_ways(ofdm, oracle, uniform):
    """288 device-resident Mode I streams (>= 256: the handle splits them into 4 pipeline ways, each on its own CUDA stream).  9 distinct base streams x 32 replicas; the call pattern is uniform blocks (advance_uniform,
    the ways stay un-joined between calls) or ragged per-stream blocks (advance: per-stream sample counts, join per call).  Every
    base stream must match the oracle fed the same block sequence, every replica its base stream bit for bit."""
    torch = _torch()
    mode, n_base, n_rep = 1, 9, 32
    n_streams = n_base * n_rep
    cfos = [0.0, 333.0, -2500.0, 50000.0, -333.0, 2500.0, -50000.0, 1000.0, 12345.0]
    base = [dabgen.make_stream(mode, 4, seed=500 + b, cfo_hz=cfos[b], start=b * 20011 + 3, snr_db=25.0) for b in range(n_base)]
    total = base[0].size
    stack = np.stack([base[s % n_base] for s in range(n_streams)])
    t = torch.from_numpy(stack.view(np.float32)).cuda()
    max_block = 131072
    d = ofdm.OfdmDemodBatch(mode, n_streams=n_streams, max_block_samples=max_block)
    d.attach_device_streams(t.data_ptr(), total, total)
    # block sequence per base stream (replicas share their base stream's sequence)
    rng = np.random.default_rng(7)
    if uniform:
        seq = [[65536] * (total // 65536)] * n_base
    else:
        seq = []
        for b in range(n_base):
            sizes, left = [], total
            while left > 0:
                n = int(min(left, rng.choice([0, 1, 99, 4096, 30000, 65536, 100000, max_block])))
                sizes.append(n)
                left -= n
            seq.append(sizes)
    n_calls = max(len(q) for q in seq)
    for k in range(n_calls):
        if uniform:
            d.advance_uniform(seq[0][k])
        else:
            d.advance([seq[s % n_base][k] if k < len(seq[s % n_base]) else 0 for s in range(n_streams)])
    d.sync()
    for b in range(n_base):
        o = oracle.OracleOfdmDemod(mode)
        off = 0
        for n in seq[b]:
            o.process(base[b][off:off + n])
            off += n
        _assert_stream_parity(oracle, mode, o, d, stream=b, min_frames=2)
        o.close()
    for s in range(n_base, n_streams):
        got, want = d.frames[s], d.frames[s % n_base]
        assert len(got) == len(want) >= 2
        for (gi, gb), (wi, wb) in zip(got, want):
            assert gi == wi and np.array_equal(gb, wb), f"stream {s} differs from its base stream {s % n_base}"
        assert d.state(s) == d.state(s % n_base)
    d.close()


@pytest.mark.parametrize("mode,block", [(2, 4096), (2, 49152), (3, 4096), (1, 65536)])
def test_back_to_back_calls_every_stream_matches_oracle(ofdm, oracle, mode, block):
    """1024 resident streams, calls queued back to back without any host synchronisation (what bench.py's modes leg does).
    Regression for a shared-memory race in the control kernel's state loop that only showed under that load: a warp late to the
    loop head could read the state thread 0 had already advanced and take a different case (a handful of the 1024 streams then
    lost lock at random; compute-sanitizer synccheck: divergent barrier).  Every stream's frame / desync counts and state must
    equal the oracle's."""
    torch = _torch()
    import bench
    fl = bench.MODE_FRAME_LEN[mode]
    n = 1024
    iq, _ = bench.build_streams_on_device(torch, n, 9, seed=777 + mode, mode=mode, frame_len=fl)
    torch.cuda.synchronize()
    calls = 8 * fl // block
    d = ofdm.OfdmDemodBatch(mode, n_streams=n, max_block_samples=block)
    d.disable_callback()
    d.attach_device_streams(iq.data_ptr(), iq.shape[1], iq.shape[1])
    for _ in range(calls):
        d.advance_uniform(block)
    d.sync()
    host = iq[::4].cpu().numpy()
    bad = []
    for i, s in enumerate(range(0, n, 4)):
        o = oracle.OracleOfdmDemod(mode)
        o.process_blocks(host[i][:calls * block], block)
        so, sd = o.state(), d.state(s)
        if (so["state"], so["total_frames_read"], so["total_frames_desync"]) != (sd["state"], sd["total_frames_read"], sd["total_frames_desync"]):
            bad.append((s, sd, so))
        o.close()
    assert not bad, bad[:4]
    d.close()


# ---------------------------------------------------------------------------------------------------------------------------
# the generic-geometry frame kernel
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode,block,cfo_hz,start,min_frames", [c for c in STREAM_CASES if c[4] > 0])
def test_generic_frame_kernel_on_dab_modes(ofdm, oracle, monkeypatch, mode, block, cfo_hz, start, min_frames):
    """DAB_B200_GENERIC_FRAME_KERNEL=1: the DAB transmission modes through ofdm_frame_kernel (the kernel every non-DAB OFDM_Params
    geometry runs on) instead of ofdm_frame_v3_kernel -- same acceptance as test_stream_matches_oracle"""
    monkeypatch.setenv("DAB_B200_GENERIC_FRAME_KERNEL", "1")
    x = dabgen.make_stream(mode, 5, seed=mode, cfo_hz=cfo_hz, start=start)
    o = oracle.OracleOfdmDemod(mode)
    o.process_blocks(x, block)
    d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=max(block, 4096))
    for off in range(0, x.size, block):
        d.process(0, x[off:off + block])
    _assert_stream_parity(oracle, mode, o, d, min_frames=min_frames)
    d.close()
    o.close()


def _custom_geometry_stream(oracle, n_frames, seed, cfo_norm, start):
    """A non-DAB OFDM geometry the reference's constructor accepts (OFDM_Params is free-form, ofdm_demodulator.cpp:80-146):
    512-point FFT, cyclic prefix 100 (DAB would be 126), 256 carriers (DAB: 384), 20 symbols per frame, NULL of 700 samples;
    random QPSK phase reference symbol and a random carrier interleaver.  Modulated here as ofdm_modulator.cpp:49-93 does:
    NULL = silence, PRS, then differentially encoded data symbols, each with its cyclic prefix."""
    P = oracle.Params(20, 612, 700, 100, 512, 256)
    rng = np.random.default_rng(seed)
    nfft, ncarr, S, cp = 512, 256, 20, 100
    bins = np.array([(nfft + k) % nfft for k in list(range(-ncarr // 2, 0)) + list(range(1, ncarr // 2 + 1))])
    prs = np.zeros(nfft, np.complex64)
    prs[bins] = np.exp(1j * np.pi / 2 * rng.integers(0, 4, ncarr)).astype(np.complex64)
    mapper = rng.permutation(ncarr).astype(np.int32)
    frames = []
    for _ in range(n_frames):
        spec = prs.copy()
        out = [np.zeros(700, np.complex64)]
        for s in range(S):
            if s > 0:
                dq = np.exp(1j * (np.pi / 4 + np.pi / 2 * rng.integers(0, 4, ncarr))).astype(np.complex64)
                nxt = np.zeros(nfft, np.complex64)
                nxt[bins] = spec[bins] * dq
                spec = nxt
            sym = np.fft.ifft(spec).astype(np.complex64) * np.float32(nfft)
            out.append(np.concatenate([sym[-cp:], sym]))
        frames.append(np.concatenate(out))
    x = np.concatenate(frames)
    x = x * np.exp(2j * np.pi * cfo_norm * np.arange(x.size))
    x = (x * (4.0 / ncarr)).astype(np.complex64)
    sig = float(np.mean(np.abs(x) ** 2))
    x = x + np.sqrt(sig / 10 ** 2.5 / 2) * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))
    return P, prs, mapper, np.ascontiguousarray(np.roll(x, -start), np.complex64)


@pytest.mark.parametrize("block,cfo_bins,start", [(4096, 0.3, 0), (1000, -7.4, 5000), (20000, 40.2, 123)])
def test_non_dab_geometry(ofdm, oracle, pkg, block, cfo_bins, start):
    """the product path for any OFDM_Params that is not one of the four DAB modes: ofdm_frame_kernel + the control kernel with a
    caller-supplied phase reference symbol and carrier mapper, against the oracle built from the same tables"""
    P, prs, mapper, x = _custom_geometry_stream(oracle, 12, seed=3, cfo_norm=cfo_bins / 512.0, start=start)
    o = oracle.OracleOfdmDemod(None, custom=(P, prs, mapper))
    o.process_blocks(x, block)
    assert o.frames_done() >= 5
    cp = pkg.capi.OfdmParams(*[getattr(P, k) for k, _ in P._fields_])
    d = ofdm.OfdmDemodBatch(None, n_streams=1, max_block_samples=max(block, 4096), params=cp, prs=prs, mapper=mapper)
    for off in range(0, x.size, block):
        d.process(0, x[off:off + block])
    got = d.frames[0]
    assert len(got) == o.frames_done()
    for i, (info, bits) in enumerate(got):
        oinfo, obits = o.frame(i)
        assert info["frame_start"] == oinfo["frame_start"] and info["fine_time_offset"] == oinfo["fine_time_offset"]
        for key in ("coarse_offset", "fine_offset_used", "fine_offset_after"):
            assert abs(info[key] - oinfo[key]) * 512 < FREQ_TOL_BINS, (i, key, info[key], oinfo[key])
        eq, lsb1, mx = dabgen.compare_bits(bits, obits)
        assert lsb1 >= LSB1_MIN, (i, eq, lsb1, mx)
    so, sd = o.state(), d.state(0)
    assert sd["state"] == so["state"] and sd["total_frames_read"] == so["total_frames_read"] and sd["total_frames_desync"] == so["total_frames_desync"]
    d.close()
    o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# GUI taps
# ---------------------------------------------------------------------------------------------------------------------------
def _db_close(a, b, floor_db=50.0, tol_db=0.05):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    keep = np.isfinite(a) & np.isfinite(b)
    keep &= np.maximum(a, b) > max(np.max(a[keep]), np.max(b[keep])) - floor_db
    return keep.sum() > 0 and float(np.max(np.abs(a[keep] - b[keep]))) <= tol_db


@pytest.mark.parametrize("mode,block,cfo_hz,start,generic", [(1, 65536, 333.0, 5000, False), (2, 4096, -2500.0, 0, False),
                                                             (3, 4096, 2500.0, 0, False), (4, 65536, 20000.0, 44444, False),
                                                             (1, 65536, -333.0, 777, True)])
def test_gui_taps_match_oracle(ofdm, oracle, monkeypatch, mode, block, cfo_hz, start, generic):
    """keep_debug_taps: GetFrameFFT / GetFrameDataVec / GetFrameDataBits / GetImpulseResponse / GetCoarseFrequencyResponse /
    GetCorrelationTimeBuffer (ofdm_demodulator.h:133-139, read by examples/gui/ofdm/render_ofdm_demod.cpp:124-338) against the
    oracle's buffers after the same stream (the oracle's are pinned to the reference's getters in test_oracle_vs_reference.py)"""
    if generic:
        monkeypatch.setenv("DAB_B200_GENERIC_FRAME_KERNEL", "1")
    x = dabgen.make_stream(mode, 4, seed=mode + 70, cfo_hz=cfo_hz, start=start, snr_db=25.0)
    o = oracle.OracleOfdmDemod(mode)
    d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=max(block, 4096), keep_debug_taps=True)
    for off in range(0, x.size, block):
        o.process(x[off:off + block])
        d.process(0, x[off:off + block])
    assert o.frames_done() == len(d.frames[0]) >= 2
    p = oracle.params(mode)
    S, nfft, ncarr = p["nb_frame_symbols"], p["nb_fft"], p["nb_data_carriers"]
    fo, fg = o.frame_fft()[:S * nfft], d.frame_fft(0)
    assert fg.size == S * nfft
    # the kernel's PLL differs from the reference's by a few 1e-4 turns of phase at most (ofdm_frame_v3.cuh): compare per bin
    # against the spectrum's scale
    assert np.max(np.abs(fg - fo)) <= 2e-3 * np.max(np.abs(fo)), float(np.max(np.abs(fg - fo)) / np.max(np.abs(fo)))
    vo, vg = o.frame_data_vec(), d.frame_data_vec(0)
    assert vg.size == (S - 1) * ncarr
    assert np.max(np.abs(vg - vo)) <= 4e-3 * np.max(np.abs(vo))
    assert _db_close(o.impulse_response(), d.impulse_response(0))
    assert _db_close(o.coarse_freq_response(), d.coarse_frequency_response(0))
    eq, lsb1, mx = dabgen.compare_bits(d.frame_data_bits(0), o.frame(o.frames_done() - 1)[1])
    assert lsb1 >= LSB1_MIN
    co, length = o.correlation_time_buffer()
    cg = d.correlation_time_buffer(0)
    assert cg.size == co.size and np.array_equal(cg[:length], co[:length])
    d.close()
    o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# raw integer IQ
# ---------------------------------------------------------------------------------------------------------------------------
def _quantise(x, fmt):
    """what a front end of that sample type delivers (QuantisedIQ<T>::from_iq, app_iq_readers.h:44-62) and what the reference's
    reader turns it back into (to_c32 + scale, :36-43, 76-88)"""
    spec = {"u8": (np.uint8, 127.5, 127.5, "<"), "s8": (np.int8, 0.0, 127.0, "<"), "s16le": (np.int16, 0.0, 32767.0, "<"),
            "s16be": (np.int16, 0.0, 32767.0, ">"), "u16le": (np.uint16, 32767.5, 32767.5, "<"), "u16be": (np.uint16, 32767.5, 32767.5, ">")}
    dt, bias, amp, order = spec[fmt]
    info = np.iinfo(dt)
    pairs = np.stack([x.real, x.imag], -1).reshape(-1).astype(np.float32)
    q = np.clip(pairs * np.float32(amp) + np.float32(bias), info.min, info.max).astype(dt)   # C cast: truncation
    host = ((q.astype(np.float32) - np.float32(bias)) * np.float32(1.0 / amp)).view(np.complex64)
    raw = q.astype(np.dtype(dt).newbyteorder(order)).view(np.uint8)
    return raw, np.ascontiguousarray(host)


@pytest.mark.parametrize("fmt", ["u8", "s8", "s16le", "s16be", "u16le", "u16be"])
@pytest.mark.parametrize("mode,block,start", [(1, 65536, 4321), (2, 4096, 30000)])
def test_raw_formats_match_host_dequantised_path(ofdm, oracle, fmt, mode, block, start):
    """raw_u8 / raw_s8 / raw_s16l / raw_s16b / raw_u16l / raw_u16b (app_iq_readers.h:109-135) dequantised on the device = the
    reference's reader dequantising on the host and the oracle demodulating the floats"""
    x = dabgen.make_stream(mode, 5, seed=31, cfo_hz=1500.0, start=start, snr_db=30.0, u8=False)
    x = x * np.float32(0.25 / np.sqrt(np.mean(np.abs(x) ** 2)))   # rms at a quarter of full scale (the transmitter's 4 / carriers is far below)
    raw, host = _quantise(x, fmt)
    o = oracle.OracleOfdmDemod(mode)
    o.process_blocks(host, block)
    assert o.frames_done() >= 2
    d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=max(block, 4096), sample_format=fmt)
    sb = d.sample_bytes
    assert sb == {"u8": 2, "s8": 2}.get(fmt, 4)
    for off in range(0, host.size, block):
        d.process_batch_raw([raw[off * sb:(off + block) * sb]])
    _assert_stream_parity(oracle, mode, o, d, min_frames=2)
    d.close()
    o.close()


@pytest.mark.parametrize("fmt", ["u8", "s16be"])
def test_raw_resident_streams(ofdm, oracle, fmt):
    """dab_ofdm_attach_device_streams_raw: rows of raw integer samples already in HBM (what a capture card DMAs there), read in
    place by the kernels; odd row stride so that the rows are not 16-byte aligned"""
    torch = _torch()
    mode, block, n_streams = 1, 65536, 5
    xs = [dabgen.make_stream(mode, 4, seed=80 + s, cfo_hz=700.0 * s, start=11111 * s + 1, snr_db=25.0, u8=False) for s in range(n_streams)]
    hosts, raws = [], []
    for x in xs:
        x = x * np.float32(0.25 / np.sqrt(np.mean(np.abs(x) ** 2)))
        raw, host = _quantise(x, fmt)
        hosts.append(host)
        raws.append(raw)
    d = ofdm.OfdmDemodBatch(mode, n_streams=n_streams, max_block_samples=block, sample_format=fmt)
    sb = d.sample_bytes
    total = hosts[0].size
    stride = total + 3                                  # samples
    buf = np.zeros((n_streams, stride * sb), np.uint8)
    for s in range(n_streams):
        buf[s, :total * sb] = raws[s]
    t = torch.from_numpy(buf).cuda()
    d.attach_device_streams(t.data_ptr(), stride, total)
    for off in range(0, total - block + 1, block):
        d.advance_uniform(block)
    d.sync()
    n_used = (total // block) * block
    for s in range(n_streams):
        o = oracle.OracleOfdmDemod(mode)
        o.process_blocks(hosts[s][:n_used], block)
        _assert_stream_parity(oracle, mode, o, d, stream=s, min_frames=2)
        o.close()
    d.close()


# ---------------------------------------------------------------------------------------------------------------------------
# host-side contracts
# ---------------------------------------------------------------------------------------------------------------------------
def test_callback_may_call_back_into_the_handle(ofdm, oracle, pkg):
    """an On_OFDM_Frame observer calling GetState() / Reset() on the same demodulator (the reference allows it) must not
    deadlock: the handle's lock is recursive and the getters are served from the host snapshot"""
    capi = pkg.capi
    mode, block = 1, 65536
    x = dabgen.make_stream(mode, 4, seed=12, cfo_hz=333.0, start=5000)
    d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=block)
    seen = []

    def on_frame(user, stream, bits, n_bits, info):
        st = capi.OfdmState()
        assert d.L.dab_ofdm_get_state(d.h, stream, C.byref(st)) == 0      # re-entrant getter
        cfg = capi.OfdmConfig()
        assert d.L.dab_ofdm_get_config(d.h, stream, C.byref(cfg)) == 0
        if len(seen) == 1:
            assert d.L.dab_ofdm_reset(d.h, stream) == 0                   # Reset() from inside the observer
        seen.append(st.total_frames_read)

    cb = capi.FRAME_CB(on_frame)
    capi.check(d.L.dab_ofdm_set_frame_callback(d.h, cb, None))
    for off in range(0, x.size, block):
        d.process(0, x[off:off + block])
    assert len(seen) >= 2
    assert d.state(0)["total_frames_desync"] >= 1                          # the Reset() took effect
    d.close()


def test_getters_between_calls_come_from_the_snapshot(ofdm, oracle):
    """after a synchronous process call the scalar getters, the sync responses and the last frame's bits are host copies: they
    must equal what the device holds (forced by an asynchronous advance-style read through dab_ofdm_sync + a second handle path)"""
    mode, block = 2, 4096
    x = dabgen.make_stream(mode, 4, seed=9, cfo_hz=-2500.0, start=0, snr_db=25.0)
    o = oracle.OracleOfdmDemod(mode)
    d = ofdm.OfdmDemodBatch(mode, n_streams=2, max_block_samples=block)
    for off in range(0, x.size, block):
        o.process(x[off:off + block])
        d.process_batch([x[off:off + block], None if off % (3 * block) == 0 else x[off:off + block][:0]])
        so, sd = o.state(), d.state(0)
        assert sd["state"] == so["state"] and sd["total_frames_read"] == so["total_frames_read"]
        assert abs(sd["signal_average"] - so["signal_average"]) <= 1e-4 * max(1e-9, abs(so["signal_average"]))
    assert len(d.frames[0]) == o.frames_done() >= 2
    assert np.array_equal(d.frame_data_bits(0), d.frames[0][-1][1])
    assert _db_close(o.impulse_response(), d.impulse_response(0))
    assert d.state(1)["total_frames_read"] == 0
    d.close()
    o.close()


# ---------------------------------------------------------------------------------------------------------------------------
# UpdateSignalAverage windows: lengths, strides and alignments of the paired-load fold (ofdm_control.cuh l1_windows)
# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("K,decimate", [(33, 1), (64, 3), (99, 5), (100, 1), (101, 2), (250, 4), (257, 1)])
@pytest.mark.parametrize("resident_offset", [None, 0, 1])
def test_signal_average_window_geometries(ofdm, oracle, K, decimate, resident_offset):
    """CalculateL1Average / UpdateSignalAverage (ofdm_demodulator.cpp:922-950) for window lengths either side of the default 100 (odd and
    even: with / without a lone sample in front of or behind the 16-byte pairs), strides from contiguous to sparse, windows longer
    than one batch of loads, through the stream ring of the host path (windows that wrap around the ring take the masked loads) and
    over resident streams whose first sample is / is not 16-byte aligned.  The running average must stay within 1e-4 of the oracle
    and the null search that uses it must find the same frames."""
    torch = _torch()
    mode, block = 4, 30000   # Mode IV: 98304-sample frames; the block is no multiple of any window stride
    x = dabgen.make_stream(mode, 6, seed=11, cfo_hz=1234.0, start=4097, snr_db=25.0)
    o = oracle.OracleOfdmDemod(mode)
    o.config.signal_l1_nb_samples = K
    o.config.signal_l1_nb_decimate = decimate
    d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=32768)
    cfg = d.get_config(0)
    cfg.signal_l1_nb_samples = K
    cfg.signal_l1_nb_decimate = decimate
    d.set_config(cfg)
    if resident_offset is None:
        for off in range(0, x.size, block):
            o.process(x[off:off + block])
            d.process(0, x[off:off + block])
    else:
        pad = np.zeros(2 + resident_offset, np.complex64)   # cudaMalloc is 256-byte aligned: + 2 samples keeps 16 bytes, + 3 does not
        t = torch.from_numpy(np.concatenate([pad, x]).view(np.float32)).cuda()
        d.attach_device_streams(t.data_ptr() + 8 * pad.size, x.size, x.size)
        n_blocks = x.size // block
        for k in range(n_blocks):
            o.process(x[k * block:(k + 1) * block])
            d.advance_uniform(block)
        d.sync()
    # long windows lose lock now and then (in the oracle as here): frames, desyncs and the average must agree whatever happens
    _assert_stream_parity(oracle, mode, o, d, min_frames=1)
    d.close()
    o.close()


def test_stage_layout_of_a_custom_carrier_map(ofdm, oracle, pkg):
    """The v3 frame kernel stages soft bits where a host-side search puts them (csrc/stage_layout.h) for whatever carrier map the
    caller passes with a DAB geometry: a reversed and a block-interleaved map must come out in exactly the order the map asks for."""
    mode = 2
    p = oracle.params(mode)
    ncarr = p["nb_data_carriers"]
    x = dabgen.make_stream(mode, 5, seed=3, cfo_hz=-777.0, start=1500, snr_db=30.0)
    ref = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=65536)
    for off in range(0, x.size, 65536):
        ref.process(0, x[off:off + 65536])
    base_map = np.asarray(ofdm.mapper_reference(mode), np.int32)
    for name, perm in (("reversed", np.arange(ncarr - 1, -1, -1, dtype=np.int32)),
                       ("blocks", (np.arange(ncarr, dtype=np.int32).reshape(8, -1).T.reshape(-1)).astype(np.int32))):
        # mapper[i] = carrier whose soft bit lands at position i
        mapper = base_map[perm]
        d = ofdm.OfdmDemodBatch(mode, n_streams=1, max_block_samples=65536, mapper=mapper)
        for off in range(0, x.size, 65536):
            d.process(0, x[off:off + 65536])
        assert len(d.frames[0]) == len(ref.frames[0]) >= 3, name
        for (gi, gb), (ri, rb) in zip(d.frames[0], ref.frames[0]):
            assert gi["frame_start"] == ri["frame_start"]
            S1 = gb.size // (2 * ncarr)
            g3 = gb.reshape(S1, 2, ncarr)
            r3 = rb.reshape(S1, 2, ncarr)
            assert np.array_equal(g3, r3[:, :, perm]), f"{name}: soft bits are not the reference order permuted by the map"
        d.close()
    ref.close()
