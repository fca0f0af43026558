"""CPU: the oracle (oracle/dab_oracle.c, our restatement) against the committed golden fixtures, i.e. against outputs of the
reference's own sources (tests/golden/make_golden.py).  This is what pins the oracle where /root/reference is absent."""
import numpy as np
import pytest

import dabgen
import goldenutil


def test_tables(oracle):
    g = goldenutil.load("tables.npz")
    keys = ("nb_frame_symbols", "nb_symbol_period", "nb_null_period", "nb_cyclic_prefix", "nb_fft", "nb_data_carriers")
    for mode in (1, 2, 3, 4):
        p = oracle.params(mode)
        assert [p[k] for k in keys] == g[f"params_{mode}"].tolist()
        assert np.array_equal(oracle.prs(mode).view(np.uint32), g[f"prs_{mode}"].view(np.uint32))  # bit-exact floats
        assert np.array_equal(oracle.mapper(mode), g[f"mapper_{mode}"])
    with pytest.raises(ValueError):
        oracle.params(5)  # get_DAB_OFDM_params throws on an invalid mode (dab_ofdm_params_ref.cpp:54)


def test_pll_and_conj_mul_sum_bit_exact(oracle):
    g = goldenutil.load("dsp.npz")
    i = 0
    while f"pll_in_{i}" in g:
        f, dt = g[f"pll_args_{i}"]
        got = oracle.apply_pll(g[f"pll_in_{i}"], float(f), float(dt))
        assert np.array_equal(got.view(np.uint32), g[f"pll_out_{i}"].view(np.uint32)), f"apply_pll case {i}"
        i += 1
    assert i >= 6
    i = 0
    while f"cms_a_{i}" in g:
        got = oracle.conj_mul_sum(g[f"cms_a_{i}"], g[f"cms_b_{i}"])
        assert np.array([got], np.complex64).view(np.uint32).tolist() == g[f"cms_out_{i}"].view(np.uint32).tolist()
        i += 1
    assert i >= 4


@pytest.mark.parametrize("key", goldenutil.viterbi_cases())
def test_viterbi_bit_exact(oracle, key):
    g = goldenutil.load("viterbi.npz")
    segs, nbytes, soft, want, want_err = goldenutil.viterbi_case(g, oracle, key)
    v = oracle.OracleViterbi()
    v.set_traceback_length(nbytes * 8)
    out, err, used = v.decode_job(soft, segs, nbytes)
    assert used == soft.size
    assert np.array_equal(out, want)
    assert err == want_err


@pytest.mark.parametrize("case", goldenutil.ofdm_cases())
def test_ofdm_stream(oracle, case):
    g = goldenutil.load(f"ofdm_{case}.npz")
    mode, block = int(g["mode"][0]), int(g["block"][0])
    nfft = oracle.params(mode)["nb_fft"]
    x = dabgen.dequantise_u8(g["iq_u8"])
    o = oracle.OracleOfdmDemod(mode)
    o.process_blocks(x, block)
    assert o.frames_done() == g["bits"].shape[0]
    for i in range(o.frames_done()):
        info, bits = o.frame(i)
        assert [info["frame_start"], info["fine_time_offset"], info["total_desync"]] == g["frame_ints"][i].tolist()
        got = np.array([info["coarse_offset"], info["fine_offset_used"], info["fine_offset_after"]], np.float32)
        assert np.all(np.abs(got - g["frame_floats"][i]) * nfft < 1e-3)       # north_star: 1e-3 of the sub-carrier spacing
        eq, lsb1, mx = dabgen.compare_bits(bits, g["bits"][i])
        assert lsb1 >= 0.999 and eq >= 0.9, (i, eq, lsb1, mx)                 # north_star: >= 99.9 % within +-1 LSB
    st = o.state()
    assert [st["state"], st["total_frames_read"], st["total_frames_desync"]] == g["final_state"].tolist()
    assert abs(st["signal_average"] - float(g["final_signal_average"][0])) <= 1e-5 * float(g["final_signal_average"][0])
    o.close()
