"""Host logic of the frame kernel's staging layout (dab-radio_b200/csrc/stage_layout.h): compiled and run on the CPU."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stage_layout_is_a_consistent_bijection_and_reduces_conflicts(tmp_path):
    exe = tmp_path / "test_stage_layout"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "dab-radio_b200", "csrc"),
                    os.path.join(ROOT, "tests", "cpp", "test_stage_layout.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    lines = out.strip().splitlines()
    assert len(lines) == 4, out
    for line in lines:
        n_inst, before, after = map(int, re.search(r"(\d+) store instructions, (\d+) wavefronts in position order, (\d+) as laid out", line).groups())
        assert after <= before and after >= n_inst, line
    # Mode I (what the bench runs): 3.4 -> about 2 wavefronts per store instruction
    n_inst, before, after = map(int, re.search(r"(\d+) store instructions, (\d+) wavefronts in position order, (\d+) as laid out", lines[0]).groups())
    assert before >= 3 * n_inst and after <= 2.2 * n_inst, lines[0]
