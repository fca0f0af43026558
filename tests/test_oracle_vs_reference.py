"""CPU, build container only: the oracle against the reference's own sources compiled live (oracle/_ref/libdabref.so).
Skipped where that library does not exist.  Mirrors the vendored library's own tests (run_tests.cpp, run_punctured_decoder.cpp,
run_simple.cpp: noiseless round trips) and adds noisy / adversarial inputs."""
import numpy as np
import pytest

import dabgen


def test_tables_match(oracle, ref):
    for mode in (1, 2, 3, 4):
        assert ref.params(mode) == oracle.params(mode)
        assert np.array_equal(ref.prs(mode).view(np.uint32), oracle.prs(mode).view(np.uint32))
        assert np.array_equal(ref.mapper(mode), oracle.mapper(mode))


def test_modulator_matches(oracle, ref):
    rng = np.random.default_rng(4)
    for mode in (1, 2, 3, 4):
        data = rng.integers(0, 256, dabgen.payload_bytes(mode), dtype=np.uint8)
        a, b = ref.modulate(mode, data), oracle.modulate(mode, data)
        assert np.abs(a - b).max() <= 4e-7 * np.abs(a).max()


@pytest.mark.parametrize("f,dt,n", [(1.6e-4, 0.0, 2552), (0.0244, 1234.5, 2552), (-0.12, 48000.0, 2552), (0.001, 3.2, 319), (1e-5, 0.0, 638)])
def test_apply_pll_bit_exact(oracle, ref, f, dt, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    assert np.array_equal(ref.apply_pll(x, f, dt).view(np.uint32), oracle.apply_pll(x, f, dt).view(np.uint32))


def _vit_pair(oracle, ref, soft, segs, nbytes):
    r, o = ref.RefViterbi(), oracle.OracleViterbi()
    r.set_traceback_length(nbytes * 8)
    o.set_traceback_length(nbytes * 8)
    r.reset()
    o.reset()
    u = 0
    for code, n in segs:
        a, b = r.update(soft[u:], code, n), o.update(soft[u:], code, n)
        assert a == b
        u += a
    assert r.get_current_decoded_bit() == o.get_current_decoded_bit()
    (br, er), (bo, eo) = r.chainback(nbytes), o.chainback(nbytes)
    return br, er, bo, eo


def test_viterbi_roundtrip_and_noise(oracle, ref):
    """run_punctured_decoder.cpp:139-191 (FIC schedule, zero errors without noise) + noise sweep, bit-exact vs AVX2 decoder"""
    rng = np.random.default_rng(8)
    segs = [(oracle.puncture_code(16), 128 * 21), (oracle.puncture_code(15), 128 * 3), (oracle.PI_X, 24)]
    for sigma in (0, 40, 80, 120, 200):
        for _ in range(10):
            data = rng.integers(0, 256, 96, dtype=np.uint8)
            tx = oracle.puncture(oracle.conv_encode(data), segs)
            rx = np.clip(np.rint(tx + sigma * rng.standard_normal(tx.size)), -128, 127).astype(np.int8)
            br, er, bo, eo = _vit_pair(oracle, ref, rx, segs, 96)
            assert np.array_equal(br, bo) and er == eo
            if sigma == 0:
                assert np.array_equal(br, data)


def test_viterbi_unpunctured_64_bytes(oracle, ref):
    """run_tests.cpp:194-241: 64 random bytes, code {109,79,83,109}, no noise, zero errors"""
    rng = np.random.default_rng(1)
    data = rng.integers(0, 256, 64, dtype=np.uint8)
    mother = oracle.conv_encode(data)
    segs = [(oracle.puncture_code(24), mother.size)]
    br, er, bo, eo = _vit_pair(oracle, ref, mother, segs, 64)
    assert np.array_equal(br, data) and np.array_equal(bo, data) and er == eo == 0


def test_viterbi_adversarial(oracle, ref):
    rng = np.random.default_rng(3)
    gens = [lambda n: np.zeros(n, np.int8), lambda n: np.full(n, -128, np.int8), lambda n: np.full(n, 127, np.int8),
            lambda n: rng.integers(-128, 128, n).astype(np.int8), lambda n: np.where(np.arange(n) % 2, 127, -128).astype(np.int8)]
    for g in gens:
        for pi in (1, 8, 16, 24):
            L = 100
            segs = [(oracle.puncture_code(pi), 128 * L), (oracle.PI_X, 24)]
            n_in = int(oracle.puncture_code(pi).sum()) * 4 * L + 12
            br, er, bo, eo = _vit_pair(oracle, ref, g(n_in), segs, 128 * L // 32)
            assert np.array_equal(br, bo) and er == eo


@pytest.mark.parametrize("mode,block,cfo,start,snr", [(1, 65536, 333.0, 5000, 20.0), (1, 1000, -50000.0, 100000, 12.0), (2, 4096, 2500.0, 0, None),
                                                      (4, 65536, -333.0, 44444, 25.0)])
def test_ofdm_stream_parity(oracle, ref, mode, block, cfo, start, snr):
    """the reference in real-time order (oracle/ref_harness.cpp) vs the oracle: north_star tolerances"""
    x = dabgen.make_stream(mode, 4, seed=mode + 40, cfo_hz=cfo, start=start, snr_db=snr)
    r, o = ref.RefOfdmDemod(mode, 1), oracle.OracleOfdmDemod(mode)
    r.process_blocks(x, block)
    o.process_blocks(x, block)
    nfft = oracle.params(mode)["nb_fft"]
    assert r.frames_done() == o.frames_done() >= 2
    for i in range(r.frames_done()):
        ir, br = r.frame(i)
        io, bo = o.frame(i)
        assert ir["frame_start"] == io["frame_start"] and ir["fine_time_offset"] == io["fine_time_offset"]
        for k in ("coarse_offset", "fine_offset_used", "fine_offset_after"):
            assert abs(ir[k] - io[k]) * nfft < 1e-3
        eq, lsb1, mx = dabgen.compare_bits(br, bo)
        assert lsb1 >= 0.999 and eq >= 0.9
    assert r.state()["total_frames_desync"] == o.state()["total_frames_desync"]
    r.close()
    o.close()


def _db_close(a, b, floor_db=50.0, tol_db=0.05):
    """dB curves agree where they are within floor_db of the peak (deep nulls are rounding noise in both)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    keep = np.isfinite(a) & np.isfinite(b) & (np.maximum(a, b) > max(np.max(a[np.isfinite(a)]), np.max(b[np.isfinite(b)])) - floor_db)
    return keep.sum() > 0 and float(np.max(np.abs(a[keep] - b[keep]))) <= tol_db


@pytest.mark.parametrize("mode,block,cfo,start", [(1, 65536, 333.0, 5000), (2, 4096, -2500.0, 0), (4, 65536, 20000.0, 44444)])
def test_ofdm_gui_taps_parity(oracle, ref, mode, block, cfo, start):
    """the GUI-visible buffers (ofdm_demodulator.h:133-139: GetFrameFFT, GetFrameDataVec, GetImpulseResponse,
    GetCoarseFrequencyResponse, GetCorrelationTimeBuffer) of the reference vs the oracle after the same stream"""
    x = dabgen.make_stream(mode, 4, seed=mode + 70, cfo_hz=cfo, start=start, snr_db=25.0)
    r, o = ref.RefOfdmDemod(mode, 1), oracle.OracleOfdmDemod(mode)
    r.process_blocks(x, block)
    o.process_blocks(x, block)
    assert r.frames_done() == o.frames_done() >= 2
    p = oracle.params(mode)
    S, nfft = p["nb_frame_symbols"], p["nb_fft"]
    fr, fo = r.frame_fft()[:S * nfft], o.frame_fft()[:S * nfft]
    assert np.max(np.abs(fr - fo)) <= 2e-4 * np.max(np.abs(fr))
    vr, vo = r.frame_data_vec(), o.frame_data_vec()
    assert np.max(np.abs(vr - vo)) <= 4e-4 * np.max(np.abs(vr))
    assert _db_close(r.impulse_response(), o.impulse_response())
    assert _db_close(r.coarse_freq_response(), o.coarse_freq_response())
    cr = r.correlation_time_buffer()
    co, length = o.correlation_time_buffer()
    assert np.array_equal(cr[:length], co[:length])
    r.close()
    o.close()
