"""GPU: the C++ mirror classes (dab-radio_b200/cpp) driven through the reference's own call patterns by tests/cpp/test_dropin.cpp
(OFDM_Block::run and FIC_Decoder::DecodeFIBGroup), checked against the golden outputs of the reference."""
import os
import subprocess

import numpy as np
import pytest

import dabgen
import ensgen
import goldenutil

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=["mirror_classes", "reference_callers_on_mirrors"])
def dropin(pkg, tmp_path_factory, request):
    """the driver binary in two builds: (1) every class from dab-radio_b200/cpp, compiled here; (2) oracle/_ref/dropin_ref_callers
    (make -C oracle refcallers, built where /root/reference exists): the REFERENCE's unchanged FIC_Decoder / MSC_Decoder /
    CIF_Deinterleaver / table code / Create_OFDM_Demodulator on top of the mirror OFDM_Demod and DAB_Viterbi_Decoder, overlaid as
    INTEGRATION.md section 2 describes -- the same caller code, the same expected outputs"""
    if request.param == "reference_callers_on_mirrors":
        exe = os.path.join(ROOT, "oracle", "_ref", "dropin_ref_callers")
        if not os.path.exists(exe):
            pytest.skip("oracle/_ref/dropin_ref_callers not built (needs /root/reference at build time)")
        return exe
    out = str(tmp_path_factory.mktemp("dropin") / "test_dropin")
    cpp = os.path.join(ROOT, "dab-radio_b200", "cpp")
    cmd = ["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"), "-I", cpp, "-I", os.path.join(cpp, "standalone"),
           os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), os.path.join(cpp, "ofdm", "ofdm_demodulator.cpp"),
           os.path.join(cpp, "dab", "algorithms", "dab_viterbi_decoder.cpp"), os.path.join(cpp, "dab", "fic", "fic_decoder.cpp"),
           os.path.join(cpp, "dab", "msc", "msc_decoder.cpp"), "-o", out, "-L", os.path.dirname(pkg.capi.LIB_PATH),
           "-ldab_b200", "-Wl,-rpath," + os.path.dirname(pkg.capi.LIB_PATH)]
    subprocess.check_call(cmd)
    return out


def test_ofdm_demod_class(dropin, tmp_path):
    g = goldenutil.load("ofdm_mode1_cfo333.npz")
    x = dabgen.dequantise_u8(g["iq_u8"])
    x.tofile(tmp_path / "iq.c64")
    out = tmp_path / "bits.bin"
    res = subprocess.run([dropin, "ofdm", "1", str(int(g["block"][0])), str(tmp_path / "iq.c64"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = np.fromfile(out, np.uint8)
    frames, off = [], 0
    while off < raw.size:
        nb = int(raw[off:off + 8].view(np.int64)[0])
        frames.append(raw[off + 8:off + 8 + nb].view(np.int8))
        off += 8 + nb
    assert len(frames) == g["bits"].shape[0] >= 2
    for i, bits in enumerate(frames):
        eq, lsb1, mx = dabgen.compare_bits(bits, g["bits"][i])
        assert lsb1 >= 0.999, (i, eq, lsb1, mx)
    assert f"frames={len(frames)} read={len(frames)} desync=0" in res.stdout, res.stdout


def test_viterbi_decoder_class(dropin, oracle, tmp_path):
    g = goldenutil.load("viterbi.npz")
    keys = [k for k in goldenutil.viterbi_cases() if k.startswith("fic__")]
    soft = np.concatenate([g[f"{k}__soft"] for k in keys])
    soft.tofile(tmp_path / "soft.i8")
    out = tmp_path / "dec.bin"
    res = subprocess.run([dropin, "fic", str(tmp_path / "soft.i8"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = np.fromfile(out, np.uint8).reshape(len(keys), 96 + 8)
    for i, k in enumerate(keys):
        assert np.array_equal(raw[i, :96], g[f"{k}__out"]), k
        assert int(raw[i, 96:].view(np.uint64)[0]) == int(g[f"{k}__err"][0]), k


def test_fic_decoder_class(dropin, oracle, tmp_path):
    """FIC_Decoder mirror (dab-radio_b200/cpp/dab/fic) driven as BasicRadio does: the CRC-valid FIBs it notifies, in order, equal
    those of the oracle's FIC_Decoder restatement -- noise, a corrupted FIB, and a Mode III group size that decodes nothing"""
    rng = np.random.default_rng(31)
    groups, want = [], []
    for g in range(12):
        fibs = ensgen.make_fib_group(rng, corrupt=[1] if g == 3 else ())
        tx = ensgen.encode_fic_group(fibs).astype(np.float64)
        sigma = (0.0, 50.0, 90.0, 140.0)[g % 4]
        rx = np.clip(np.rint(tx + sigma * rng.standard_normal(tx.size)), -127, 127).astype(np.int8)
        groups.append(rx)
        b, v, e = oracle.fic_decode_group(rx, 3)
        for i in range(3):
            if v[i]:
                want.append((g, b[i * 32:i * 32 + 30]))
    np.concatenate(groups).tofile(tmp_path / "fic.i8")
    out = tmp_path / "fibs.bin"
    res = subprocess.run([dropin, "ficdec", "2304", "3", str(tmp_path / "fic.i8"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = np.fromfile(out, np.uint8)
    got, off = [], 0
    while off < raw.size:
        g, n = (int(v) for v in raw[off:off + 8].view(np.int32))
        got.append((g, raw[off + 8:off + 8 + n]))
        off += 8 + n
    assert len(got) == len(want) and len(want) >= 20
    for (g0, b0), (g1, b1) in zip(got, want):
        assert g0 == g1 and np.array_equal(b0, b1)
    # Mode III: 3072-bit groups are rejected by the reference (fic_decoder.cpp:68-75) -> no FIB is ever notified
    np.zeros(3072 * 2, np.int8).tofile(tmp_path / "fic3.i8")
    res = subprocess.run([dropin, "ficdec", "3072", "4", str(tmp_path / "fic3.i8"), str(out)], capture_output=True, text=True)
    assert res.returncode == 0 and os.path.getsize(out) == 0


@pytest.mark.parametrize("layout", [(0, 48, 0, 0, 2, 0), (150, 35, 1, 4, 0, 0), (96, 54, 0, 0, 0, 1), (850, 48, 0, 0, 2, 0)])
def test_msc_decoder_class(dropin, oracle, tmp_path, layout):
    """MSC_Decoder mirror (dab-radio_b200/cpp/dab/msc): DecodeCIF over 24 CIFs -- empty while the 16-CIF de-interleaver fills,
    then the oracle's bytes bit for bit (EEP-A, UEP, EEP-B); a sub-channel that overflows the CIF always returns nothing"""
    sc = oracle.subchannel(*layout)
    nb_cif_bits = 55296
    overflow = (layout[0] + layout[1]) * 64 > nb_cif_bits
    rng = np.random.default_rng(layout[0] + 1)
    ref = oracle.OracleMscDecoder(sc)
    cifs, want = [], []
    for c in range(24):
        cif = rng.integers(-40, 41, nb_cif_bits).astype(np.int8)
        if not overflow:
            payload = rng.integers(0, 256, ensgen.sub_decoded_bytes(sc), dtype=np.uint8)
            tx = ensgen.encode_subchannel(sc, payload).astype(np.float64)
            cif[layout[0] * 64:(layout[0] + layout[1]) * 64] = np.clip(np.rint(tx + 45.0 * rng.standard_normal(tx.size)), -127, 127)
        cifs.append(cif)
        want.append(ref.decode_cif(cif)[0])
    np.concatenate(cifs).tofile(tmp_path / "cifs.i8")
    out = tmp_path / "msc.bin"
    res = subprocess.run([dropin, "mscdec"] + [str(v) for v in layout] + [str(nb_cif_bits), str(tmp_path / "cifs.i8"), str(out)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    raw = np.fromfile(out, np.uint8)
    off = 0
    for c in range(24):
        n = int(raw[off:off + 4].view(np.int32)[0])
        assert n == want[c].size, (c, n, want[c].size)
        assert np.array_equal(raw[off + 4:off + 4 + n], want[c]), c
        off += 4 + n
    assert off == raw.size
    assert overflow or want[-1].size > 0
