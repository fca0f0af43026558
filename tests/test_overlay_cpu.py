"""CPU: the mirror classes overlaid on a COPY of the reference tree -- the reference's own callers
(FIC_Decoder / MSC_Decoder on the mirror DAB_Viterbi_Decoder; BasicRadio and Basic_Audio_Channel on the mirror FIC_Decoder /
MSC_Decoder; OFDM_Block on the mirror OFDM_Demod) must compile unchanged.  Syntax check only: no GPU, nothing is linked or run.
Skipped where /root/reference does not exist (the GPU box)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CPP = os.path.join(ROOT, "dab-radio_b200", "cpp")

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "ofdm", "ofdm_demodulator.cpp")), reason="needs the reference tree")

VARIANTS = {
    # mirror files copied over the reference's (same relative paths under src/) -> reference sources that must still compile
    "viterbi_and_ofdm": (["dab/algorithms/dab_viterbi_decoder.h", "dab/algorithms/dab_viterbi_decoder.cpp", "ofdm/ofdm_demodulator.h",
                          "ofdm/ofdm_demodulator.cpp"],
                         ["dab/fic/fic_decoder.cpp", "dab/msc/msc_decoder.cpp", "dab/algorithms/dab_viterbi_decoder.cpp",
                          "ofdm/ofdm_demodulator.cpp", "basic_radio/basic_radio.cpp"]),
    "decoders_too": (["dab/algorithms/dab_viterbi_decoder.h", "dab/algorithms/dab_viterbi_decoder.cpp", "ofdm/ofdm_demodulator.h",
                      "ofdm/ofdm_demodulator.cpp", "dab/fic/fic_decoder.h", "dab/fic/fic_decoder.cpp", "dab/msc/msc_decoder.h",
                      "dab/msc/msc_decoder.cpp"],
                     ["dab/fic/fic_decoder.cpp", "dab/msc/msc_decoder.cpp", "basic_radio/basic_radio.cpp",
                      "basic_radio/basic_audio_channel.cpp"]),
}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_reference_callers_compile_on_the_overlay(variant, tmp_path):
    overlay, sources = VARIANTS[variant]
    src = tmp_path / "src"
    shutil.copytree(os.path.join(REF, "src"), src)
    for rel in overlay:
        shutil.copy(os.path.join(CPP, rel), src / rel)
    inc = ["-I", str(src), "-I", os.path.join(ROOT, "include"), "-I", os.path.join(REF, "vendor", "fmt", "include"),
           "-I", os.path.join(REF, "vendor", "viterbi_decoder", "include")]
    for rel in sources:
        res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-DFMT_HEADER_ONLY"] + inc + [str(src / rel)], capture_output=True, text=True)
        assert res.returncode == 0, f"{variant}: {rel}\n{res.stderr[-2000:]}"


def test_ofdm_block_compiles_against_the_mirror_demodulator(tmp_path):
    """examples/app_helpers/app_ofdm_blocks.h (OFDM_Block: Process loop + On_OFDM_Frame().Attach) with the mirror header found first"""
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "app_helpers/app_ofdm_blocks.h"\nint main() { return 0; }\n')
    res = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-I", CPP, "-I", os.path.join(REF, "src"),
                          "-I", os.path.join(REF, "examples"), str(tu)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
