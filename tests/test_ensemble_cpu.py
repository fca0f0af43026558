"""CPU: the oracle's FIC_Decoder / MSC_Decoder / CIF_Deinterleaver restatement against the committed golden fixture (outputs of
the reference's own sources, tests/golden/make_golden.py::golden_ensemble) and, where oracle/_ref exists, against the reference
live; plus the device-free part of the dab_ensemble_* C ABI (tables and schedules)."""
import importlib

import numpy as np
import pytest

import ensgen
import goldenutil


def test_golden_ensemble(oracle):
    subs, frames, fib_bytes, fib_valid, lens, msc_bytes, scr = goldenutil.ensemble_case()
    assert np.array_equal(oracle.scrambler_bytes(scr.size), scr)
    res = ensgen.oracle_decode_stream(1, [oracle.subchannel(*a) for a in subs], frames)
    assert len(res) == len(frames) == 6
    for f, (fb, fv, fe, msc) in enumerate(res):
        for c in range(4):
            assert np.array_equal(fb[c], fib_bytes[f, c]), (f, c)
            assert np.array_equal(fv[c], fib_valid[f, c]), (f, c)
            for k, (b, e) in enumerate(msc[c]):
                assert b.size == lens[f, c, k], (f, c, k, b.size)
                assert np.array_equal(b, msc_bytes[f][c][k]), (f, c, k)
    assert fib_valid.sum() == fib_valid.size - 2          # the two FIBs corrupted before encoding fail their CRC
    assert lens[3, 3, :6].tolist() == [48, 96, 96, 24, 140, 144] and not lens[:3].any() and not lens[:, :, 6].any()


def test_fic_rejects_other_group_sizes(oracle):
    """fic_decoder.cpp:68-75: only the Mode I group size (2304 soft bits -> 768 bits) is decoded; Mode III's 3072 is not"""
    out, valid, err = oracle.fic_decode_group(np.zeros(3072, np.int8), 4)
    assert err is None and not valid.any() and not out.any()


def test_oracle_vs_reference_fic_msc_deint(oracle, ref):
    rng = np.random.default_rng(11)
    assert np.array_equal(oracle.scrambler_bytes(1000), ref.scrambler_bytes(1000))
    a, b = oracle.OracleDeinterleaver(64 * 5), ref.RefDeinterleaver(64 * 5)
    for k in range(36):
        x = rng.integers(-128, 128, 64 * 5, dtype=np.int8)
        oa, ob = a.push(x), b.push(x)
        assert (oa is None) == (ob is None) == (k < 15)
        if oa is not None:
            assert np.array_equal(oa, ob)
    rf = ref.RefFicDecoder()
    for sigma in (0, 70, 130):
        for t in range(4):
            g = ensgen.make_fib_group(rng, corrupt=[t] if t < 3 else [])
            soft = ensgen.encode_fic_group(g)
            soft = np.clip(np.rint(soft + sigma * rng.standard_normal(soft.size)), -128, 127).astype(np.int8)
            o1, v1, _ = oracle.fic_decode_group(soft)
            o2, v2 = rf.decode_group(soft)
            assert np.array_equal(o1, o2) and np.array_equal(v1, v2)


@pytest.mark.parametrize("args", [(10, 12 * 3, 0, 0, 0, 0), (3, 8, 0, 0, 1, 0), (3, 16, 0, 0, 1, 0), (0, 6 * 8, 0, 0, 2, 0), (7, 4 * 5, 0, 0, 3, 0),
                                  (1, 27, 0, 0, 0, 1), (1, 21 * 2, 0, 0, 1, 1), (1, 18, 0, 0, 2, 1), (1, 15 * 3, 0, 0, 3, 1),
                                  (5, 16, 1, 0, 0, 0), (5, 35, 1, 4, 0, 0), (5, 84, 1, 33, 0, 0), (5, 64, 1, 34, 0, 0), (5, 416, 1, 63, 0, 0),
                                  (860, 8, 0, 0, 2, 0)])
def test_oracle_vs_reference_msc(oracle, ref, args):
    rng = np.random.default_rng(args[1])
    om, rm = oracle.OracleMscDecoder(oracle.subchannel(*args)), ref.RefMscDecoder(*args)
    total = 0
    for k in range(19):
        cif = rng.integers(-127, 128, 864 * 64, dtype=np.int8)
        a, _ = om.decode_cif(cif)
        b = rm.decode_cif(cif)
        assert a.size == b.size and np.array_equal(a, b), (k, a.size, b.size)
        total += a.size
    assert (total > 0) == (args[0] + args[1] <= 864)


def test_capi_dab_parameters_and_schedules(pkg, oracle):
    """no device needed: get_dab_parameters and MSC_Decoder's puncturing schedule for every protection profile"""
    ens = importlib.import_module("dab-radio_b200.ensemble")
    for mode, (cifs, fic, fibcif, fibs, cif) in ensgen.MODE_GEOM.items():
        p = ens.dab_parameters(mode)
        assert (p.nb_cifs, p.nb_fic_bits, p.nb_fib_cif_bits, p.nb_fibs_per_cif, p.nb_cif_bits) == (cifs, fic, fibcif, fibs, cif)
        assert p.nb_frame_bits == p.nb_fic_bits + p.nb_msc_bits
    with pytest.raises(pkg.capi.DabError):
        ens.dab_parameters(5)                 # get_dab_parameters throws on an invalid mode (dab_parameters.h:74-76)

    def check(sc_args):
        sch, n_soft = ens.subchannel_schedule(ens.subchannel(*sc_args))
        segs = oracle.msc_segments(oracle.subchannel(*sc_args))
        assert n_soft == sc_args[1] * 64
        # the oracle lists every update(); the C ABI drops the ones that underrun (the reference decodes nothing for them)
        left, kept = n_soft, []
        for code, n_out in segs:
            need = int(np.resize(code, n_out // 4).astype(np.int64).sum())
            if n_out and need <= left:
                kept.append((code, n_out))
                left -= need
        assert sch.n_seg == len(kept)
        for i, (code, n_out) in enumerate(kept):
            assert sch.seg[i].n_out == n_out and sch.seg[i].code_len == code.size
            assert list(sch.seg[i].counts[:code.size]) == code.tolist()
        assert sch.n_out_bytes == (sum(n for _, n in kept) // 4 - 6) // 8

    for lvl in range(4):
        for n in (1, 2, 7):
            check((0, [12, 8, 6, 4][lvl] * n, False, 0, lvl, False))
            check((0, [27, 21, 18, 15][lvl] * n, False, 0, lvl, True))
    for idx in range(64):
        check((0, oracle.uep_subchannel_size(idx), True, idx, 0, False))
    with pytest.raises(pkg.capi.DabError):
        ens.subchannel_schedule(ens.subchannel(0, 16, True, 64))
