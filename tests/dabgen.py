"""Synthetic DAB IQ generator for the tests and the bench (test infrastructure).

Follows the reference's examples/simulate_transmitter.cpp:138-178: random payload -> OFDM_Modulator::ProcessBlock (through the
oracle's restatement, oracle/dab_oracle.c: orc_modulate) -> optional frequency shift (apply_pll) -> scale 4/nb_data_carriers ->
8-bit quantise -> dequantise as the apps do (examples/app_helpers/app_iq_readers.h:17-69).  Impairments the reference has no
generator for (AWGN, multipath, sample-rate drift) are ours, seeded.
"""
import numpy as np

from oracle import pyoracle as po


def frame_len(mode):
    p = po.params(mode)
    return p["nb_null_period"] + p["nb_symbol_period"] * p["nb_frame_symbols"]


def payload_bytes(mode):
    p = po.params(mode)
    return (p["nb_frame_symbols"] - 1) * p["nb_data_carriers"] * 2 // 8


def quantise_u8(x):
    q = lambda v: np.clip(v * 127.5 + 127.5, 0, 255).astype(np.uint8)
    return np.stack([q(x.real), q(x.imag)], -1).reshape(-1)


def dequantise_u8(iq8):
    return ((iq8.astype(np.float32) - np.float32(127.5)) * np.float32(1.0 / 127.5)).view(np.complex64).reshape(-1)


def make_stream(mode, n_frames, seed=1, cfo_hz=0.0, start=0, snr_db=None, u8=True, multipath=None, drift_ppm=0.0, same_frame=False,
                return_u8=False):
    """Returns complex64 samples: n_frames modulated frames, rotated so that the stream starts `start` samples into a frame."""
    p = po.params(mode)
    rng = np.random.default_rng(seed)
    nb = payload_bytes(mode)
    if same_frame:
        one = po.modulate(mode, rng.integers(0, 256, nb, dtype=np.uint8))
        x = np.tile(one, n_frames)
    else:
        x = np.concatenate([po.modulate(mode, rng.integers(0, 256, nb, dtype=np.uint8)) for _ in range(n_frames)])
    if multipath:
        y = np.zeros_like(x)
        for delay, gain_db, phase in multipath:
            g = np.float32(10.0 ** (gain_db / 20.0)) * np.exp(1j * phase).astype(np.complex64)
            y[delay:] += g * x[:x.size - delay]
        x = y
    if drift_ppm:
        # resample by (1 + ppm*1e-6) with a windowed-sinc interpolator (8 taps each side)
        ratio = 1.0 + drift_ppm * 1e-6
        t = np.arange(x.size, dtype=np.float64) * ratio
        base = np.floor(t).astype(np.int64)
        frac = t - base
        y = np.zeros(x.size, np.complex128)
        for k in range(-7, 9):
            idx = np.clip(base + k, 0, x.size - 1)
            arg = frac - k
            w = np.sinc(arg) * (0.5 + 0.5 * np.cos(np.pi * np.clip(arg / 8.0, -1, 1)))
            y += w * x[idx]
        x = y.astype(np.complex64)
    if cfo_hz:
        x = po.apply_pll(x, cfo_hz / 2.048e6)
    x = (x * np.float32(4.0 / p["nb_data_carriers"])).astype(np.complex64)
    if snr_db is not None:
        sig = float(np.mean(np.abs(x) ** 2))
        nz = np.sqrt(sig / 10 ** (snr_db / 10) / 2)
        x = (x + nz * (rng.standard_normal(x.size) + 1j * rng.standard_normal(x.size))).astype(np.complex64)
    if start:
        x = np.roll(x, -start)
    if u8 or return_u8:
        iq8 = quantise_u8(x)
        if return_u8:
            return iq8
        x = dequantise_u8(iq8)
    return np.ascontiguousarray(x, np.complex64)


def aligned_frame(mode, seed=1, cfo_hz=0.0, snr_db=None):
    """One frame without its NULL symbol: nb_frame_symbols * nb_symbol_period samples starting at the PRS cyclic prefix."""
    p = po.params(mode)
    x = make_stream(mode, 1, seed=seed, cfo_hz=cfo_hz, snr_db=snr_db, u8=False)
    return np.ascontiguousarray(x[p["nb_null_period"]:])


def compare_bits(a, b):
    """fraction equal, fraction within +-1 LSB, max abs difference"""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16))
    return float((d == 0).mean()), float((d <= 1).mean()), int(d.max())
