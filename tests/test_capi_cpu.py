"""CPU: libdab_b200.so builds, loads and exports every symbol include/dab_b200.h declares; host-side logic that needs no GPU
(tables, schedule digestion, argument validation); and the no-fallback rule: without a device every create call fails loudly."""
import ctypes as C
import importlib
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "dab_b200.h")).read()
    declared = set(re.findall(r"DAB_API\s+[^;(]*?\b(dab_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 40
    L = pkg.capi.load()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"not exported: {missing}"
    assert declared == set(pkg.capi.EXPORTED_SYMBOLS), declared ^ set(pkg.capi.EXPORTED_SYMBOLS)


def test_tables_match_oracle(pkg, oracle):
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    vit = importlib.import_module("dab-radio_b200.viterbi")
    for mode in (1, 2, 3, 4):
        assert ofdm.ofdm_params(mode).asdict() == oracle.params(mode)
        assert np.array_equal(ofdm.prs_reference(mode).view(np.uint32), oracle.prs(mode).view(np.uint32))
        assert np.array_equal(ofdm.mapper_reference(mode), oracle.mapper(mode))
    with pytest.raises(pkg.capi.DabError):
        ofdm.ofdm_params(0)
    for pi in range(1, 25):
        assert np.array_equal(vit.puncture_code(pi), oracle.puncture_code(pi))
    assert np.array_equal(vit.puncture_code(0), oracle.PI_X)


def test_schedule_digest(pkg):
    vit = importlib.import_module("dab-radio_b200.viterbi")
    assert vit.schedule_soft_symbols(vit.fic_schedule()) == 2304          # 3 FIBs x 256 bits x 3 (fic_decoder.cpp:40-41)
    eep = vit.make_schedule([(vit.puncture_code(8), 128 * 45), (vit.puncture_code(7), 128 * 3), (vit.puncture_code(0), 24)], 192)
    assert vit.schedule_soft_symbols(eep) == 48 * 64                      # 48 CU x 64 bits (msc_decoder.cpp:22-29)
    bad = vit.make_schedule([(vit.puncture_code(8), 130)], 1)             # not a multiple of the code rate
    assert vit.schedule_soft_symbols(bad) == pkg.capi.DAB_ERR_INVALID


def test_no_cpu_fallback(pkg):
    """On a machine without a B200 the product must fail loudly, never compute on the CPU."""
    L = pkg.capi.load()
    if L.dab_device_count() > 0:
        pytest.skip("a B200 is present")
    vit = importlib.import_module("dab-radio_b200.viterbi")
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    with pytest.raises(pkg.capi.DabError) as e:
        vit.ViterbiBatch(0)
    assert e.value.status == pkg.capi.DAB_ERR_NO_DEVICE
    with pytest.raises(pkg.capi.DabError) as e:
        ofdm.OfdmDemodBatch(1)
    assert e.value.status == pkg.capi.DAB_ERR_NO_DEVICE
    assert L.dab_ofdm_process(None, 0, None, 0) == pkg.capi.DAB_ERR_INVALID


def test_product_never_touches_the_oracle():
    """the product path must not import, link or call anything under oracle/"""
    for base, _, files in os.walk(os.path.join(ROOT, "dab-radio_b200")):
        if os.path.basename(base) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "pyoracle" not in text and "pyref" not in text and "dab_oracle" not in text and "liboracle" not in text, f
