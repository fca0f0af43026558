"""GPU: the C++ mirror classes (dab-radio_b200/cpp) driven through the reference's own call patterns by tests/cpp/test_dropin.cpp
(OFDM_Block::run and FIC_Decoder::DecodeFIBGroup), checked against the golden outputs of the reference."""
import os
import subprocess

import numpy as np
import pytest

import dabgen
import goldenutil

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dropin(pkg, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dropin") / "test_dropin")
    cpp = os.path.join(ROOT, "dab-radio_b200", "cpp")
    cmd = ["g++", "-std=c++20", "-O2", "-I", os.path.join(ROOT, "include"), "-I", cpp, "-I", os.path.join(cpp, "standalone"),
           os.path.join(ROOT, "tests", "cpp", "test_dropin.cpp"), os.path.join(cpp, "ofdm", "ofdm_demodulator.cpp"),
           os.path.join(cpp, "dab", "algorithms", "dab_viterbi_decoder.cpp"), "-o", out, "-L", os.path.dirname(pkg.capi.LIB_PATH),
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
