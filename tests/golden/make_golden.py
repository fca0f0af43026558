"""Generates tests/golden/*.npz from the reference's own sources compiled here (oracle/_ref/libdabref.so, built by
`make -C oracle ref` from /root/reference).  Run in the build container only; the fixtures are committed so that the oracle and
the CUDA path can be checked against the real reference where /root/reference does not exist (the GPU box).

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dabgen  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle import pyref  # noqa: E402


def golden_tables():
    out = {}
    for mode in (1, 2, 3, 4):
        p = pyref.params(mode)
        out[f"params_{mode}"] = np.array([p[k] for k in ("nb_frame_symbols", "nb_symbol_period", "nb_null_period", "nb_cyclic_prefix",
                                                         "nb_fft", "nb_data_carriers")], np.int64)
        out[f"prs_{mode}"] = pyref.prs(mode)
        out[f"mapper_{mode}"] = pyref.mapper(mode)
    np.savez_compressed(os.path.join(HERE, "tables.npz"), **out)


def golden_dsp():
    rng = np.random.default_rng(2024)
    out = {}
    cases = [(1.6e-4, 0.0, 2552), (0.0244, 1234.5, 2552), (-0.12, 48000.0, 2552), (0.3, 7.0, 2048), (0.001, 3.2, 319), (1e-5, 0.0, 638)]
    for i, (f, dt, n) in enumerate(cases):
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        out[f"pll_in_{i}"] = x
        out[f"pll_args_{i}"] = np.array([f, dt], np.float32)
        out[f"pll_out_{i}"] = pyref.apply_pll(x, np.float32(f), np.float32(dt))
    for i, n in enumerate((504, 252, 126, 63)):
        a = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        b = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        out[f"cms_a_{i}"], out[f"cms_b_{i}"] = a, b
        out[f"cms_out_{i}"] = np.array([pyref.conj_mul_sum(a, b)], np.complex64)
    np.savez_compressed(os.path.join(HERE, "dsp.npz"), **out)


def golden_viterbi():
    """DAB_Viterbi_Decoder (AVX2 u16) outputs: FIC / EEP / UEP schedules under noise, plus adversarial inputs."""
    rng = np.random.default_rng(77)
    PI = po.puncture_code
    schedules = {
        "fic": ([(16, 128 * 21), (15, 128 * 3), (0, 24)], 96),
        "eep3a_48cu": ([(8, 128 * 45), (7, 128 * 3), (0, 24)], 192),
        "uep_row0": ([(5, 128 * 3), (3, 128 * 4), (2, 128 * 17), (0, 24)], 96),
        "pi24_short": ([(24, 128 * 4), (0, 24)], 16),
        "pi1_long": ([(1, 128 * 40), (0, 24)], 160),
    }
    out = {}
    names = []
    for name, (spec, nbytes) in schedules.items():
        segs = [(PI(pi) if pi else po.PI_X, n) for pi, n in spec]
        n_soft = sum(int(np.resize(c, n // 4).astype(np.int64).sum()) for c, n in segs)
        gens = []
        for sigma in (0, 60, 110, 200):
            data = rng.integers(0, 256, nbytes, dtype=np.uint8)
            tx = po.puncture(po.conv_encode(data), segs)
            assert tx.size == n_soft
            gens.append((f"s{sigma}", np.clip(np.rint(tx + sigma * rng.standard_normal(tx.size)), -128, 127).astype(np.int8)))
        gens += [("zeros", np.zeros(n_soft, np.int8)), ("m128", np.full(n_soft, -128, np.int8)), ("p127", np.full(n_soft, 127, np.int8)),
                 ("alt", np.where(np.arange(n_soft) % 2, 127, -128).astype(np.int8)), ("rand", rng.integers(-128, 128, n_soft).astype(np.int8))]
        for tag, soft in gens:
            v = pyref.RefViterbi()
            v.set_traceback_length(nbytes * 8)
            v.reset()
            used = 0
            for code, n in segs:
                used += v.update(soft[used:], code, n)
            assert used == n_soft
            dec, err = v.chainback(nbytes)
            key = f"{name}__{tag}"
            names.append(key)
            out[f"{key}__soft"] = soft
            out[f"{key}__out"] = dec
            out[f"{key}__err"] = np.array([err], np.uint64)
            v.close()
        out[f"{name}__spec"] = np.array(spec, np.int64)
        out[f"{name}__nbytes"] = np.array([nbytes], np.int64)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "viterbi.npz"), **out)


OFDM_CASES = [
    # name, mode, frames, cfo_hz, start, snr_db, block
    ("mode1_cfo333", 1, 3, 333.0, 77777, 20.0, 65536),
    ("mode2_cfo2500", 2, 4, -2500.0, 30000, 25.0, 4096),
    ("mode3_start0", 3, 4, -2500.0, 0, None, 4096),
    ("mode3_neverlocks", 3, 3, 333.0, 20000, None, 4096),
    ("mode4_cfo50k", 4, 4, 50000.0, 12345, 20.0, 65536),
]


def golden_ofdm():
    """OFDM_Demod in real-time order (oracle/ref_harness.cpp): raw 8-bit IQ in, per-frame sync results and soft bits out."""
    for name, mode, frames, cfo, start, snr, block in OFDM_CASES:
        iq8 = dabgen.make_stream(mode, frames, seed=100 + mode, cfo_hz=cfo, start=start, snr_db=snr, return_u8=True)
        x = dabgen.dequantise_u8(iq8)
        r = pyref.RefOfdmDemod(mode, 1)
        r.process_blocks(x, block)
        n = r.frames_done()
        infos, bits = [], []
        for i in range(n):
            info, b = r.frame(i)
            infos.append([info["frame_start"], info["fine_time_offset"], info["total_desync"]])
            bits.append(b)
        floats = [[r.frame(i)[0][k] for k in ("coarse_offset", "fine_offset_used", "fine_offset_after")] for i in range(n)]
        st = r.state()
        np.savez_compressed(os.path.join(HERE, f"ofdm_{name}.npz"), iq_u8=iq8, mode=np.array([mode]), block=np.array([block]),
                            frame_ints=np.array(infos, np.int64).reshape(n, 3), frame_floats=np.array(floats, np.float32).reshape(n, 3),
                            bits=np.array(bits, np.int8).reshape(n, -1) if n else np.zeros((0, r.frame_bits), np.int8),
                            final_state=np.array([st["state"], st["total_frames_read"], st["total_frames_desync"]], np.int64),
                            final_signal_average=np.array([st["signal_average"]], np.float32))
        print(name, "frames", n, "desync", st["total_frames_desync"])
        r.close()


ENSEMBLE_SUBS = [
    # start_address, length, is_uep, uep_prot_index, eep_prot_level, eep_type_b
    (0, 12, 0, 0, 2, 0),     # EEP 3-A, n = 2
    (12, 27, 0, 0, 0, 1),    # EEP 1-B, n = 1
    (40, 16, 1, 0, 0, 0),    # UEP row 0 (32 kbit/s, level 5)
    (56, 8, 0, 0, 1, 0),     # EEP 2-A with 8 CU: the special row
    (64, 64, 1, 34, 0, 0),   # UEP row 34: the reference's table gives it 64 CU, its last segments underrun and are dropped
    (128, 24, 0, 0, 3, 0),   # EEP 4-A, n = 6
    (860, 8, 0, 0, 2, 0),    # overflows the CIF: never decoded (msc_decoder.cpp:49-54)
]
ENSEMBLE_USED_CU = 160


def golden_ensemble():
    """FIC_Decoder + MSC_Decoder (reference sources) over 6 consecutive Mode I frames of a noisy synthetic ensemble."""
    import ensgen
    subs = [po.subchannel(*a) for a in ENSEMBLE_SUBS]
    tx = ensgen.EnsembleTx(1, subs[:-1], seed=321, sigma=75.0)
    frames = [tx.next_frame(corrupt_fibs=[(1, 2), (3, 0)] if f == 2 else ()) for f in range(6)]
    nb_cifs, nb_fic_bits, nb_fib_cif_bits, nb_fibs, nb_cif_bits = ensgen.MODE_GEOM[1]
    used = ENSEMBLE_USED_CU * 64
    fic = pyref.RefFicDecoder(nb_fib_cif_bits, nb_fibs)
    mscs = [pyref.RefMscDecoder(*a) for a in ENSEMBLE_SUBS]
    out = {"subs": np.array(ENSEMBLE_SUBS, np.int32), "used_bits": np.array([used])}
    fib_bytes, fib_valid, msc_len, msc_bytes = [], [], [], []
    for f, frame in enumerate(frames):
        msc = frame[nb_fic_bits:].reshape(nb_cifs, nb_cif_bits)   # a view: the unused capacity units are blanked to keep the fixture small
        msc[:, used:] = 0
        out[f"frame{f}_fic"] = frame[:nb_fic_bits]
        out[f"frame{f}_msc"] = np.ascontiguousarray(msc[:, :used])
        for c in range(nb_cifs):
            b, v = fic.decode_group(frame[c * nb_fib_cif_bits:(c + 1) * nb_fib_cif_bits], c)
            fib_bytes.append(b)
            fib_valid.append(v)
            for m in mscs:
                d = m.decode_cif(msc[c])
                msc_len.append(d.size)
                msc_bytes.append(d)
    out["fib_bytes"] = np.array(fib_bytes, np.uint8).reshape(len(frames), nb_cifs, 96)
    out["fib_valid"] = np.array(fib_valid, np.uint8).reshape(len(frames), nb_cifs, nb_fibs)
    out["msc_len"] = np.array(msc_len, np.int32).reshape(len(frames), nb_cifs, len(mscs))
    out["msc_bytes"] = np.concatenate(msc_bytes)
    out["scrambler"] = pyref.scrambler_bytes(256)
    np.savez_compressed(os.path.join(HERE, "ensemble.npz"), **out)
    print("ensemble: valid FIBs", int(out["fib_valid"].sum()), "of", out["fib_valid"].size, "msc bytes", out["msc_bytes"].size,
          "lens", out["msc_len"][-1, -1].tolist())


if __name__ == "__main__":
    only = sys.argv[1:]
    if only == ["ensemble"]:
        golden_ensemble()
        sys.exit(0)
    golden_ensemble()
    golden_tables()
    golden_dsp()
    golden_viterbi()
    golden_ofdm()
    print(pyref.lib().ref_build_info().decode())
