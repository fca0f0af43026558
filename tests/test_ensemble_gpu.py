"""GPU parity: the CUDA ensemble decoder (dab_ensemble_* through the C ABI: CIF de-interleave -> Viterbi -> energy dispersal ->
FIB CRC) against the oracle's restatement of FIC_Decoder / MSC_Decoder and the golden fixture generated from the reference.
Bit-exact bytes, CRC flags, byte counts and u64 path errors."""
import importlib

import numpy as np
import pytest

import ensgen
import goldenutil

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["warp", "lanes"])
def ens(pkg, request, monkeypatch):
    """every case runs with the Viterbi stage in both forms: one trellis per warp and one per thread (the bulk form the
    default dispatch takes from 4096 trellises per call) -- DAB_B200_VITERBI_LANES is read at every launch"""
    monkeypatch.setenv("DAB_B200_VITERBI_LANES", "1" if request.param == "lanes" else "0")
    return importlib.import_module("dab-radio_b200.ensemble")


def _check_stream(dec, stream, want, subs_n, tag):
    fb, fv, fe, msc = want
    got_b, got_v, got_e = dec.read_fic(stream)
    for c in range(len(fb)):
        if fe[c] is None:      # group size the reference does not decode
            assert not got_v[c].any()
            continue
        assert np.array_equal(got_b[c], fb[c]), (tag, stream, c)
        assert np.array_equal(got_v[c], fv[c]), (tag, stream, c)
        assert int(got_e[c]) == fe[c], (tag, stream, c)
    for c in range(len(msc)):
        for k in range(subs_n):
            b, n, e = dec.read_msc(stream, c, k)
            wb, we = msc[c][k]
            assert max(n, 0) == wb.size, (tag, stream, c, k, n, wb.size)
            assert np.array_equal(b, wb), (tag, stream, c, k)
            if wb.size:
                assert e == we, (tag, stream, c, k)


def test_golden_fixture(ens, oracle):
    """outputs of the reference's own FIC_Decoder / MSC_Decoder (tests/golden/ensemble.npz)"""
    subs, frames, fib_bytes, fib_valid, lens, msc_bytes, _ = goldenutil.ensemble_case()
    dec = ens.EnsembleDecoder(1, n_streams=1, max_subchannels=8)
    dec.set_subchannels(0, [ens.subchannel(*a) for a in subs])
    for f, frame in enumerate(frames):
        dec.decode_frames(frame)
        b, v, _ = dec.read_fic(0)
        assert np.array_equal(b, fib_bytes[f]) and np.array_equal(v, fib_valid[f]), f
        for c in range(4):
            for k in range(len(subs)):
                got, n, _ = dec.read_msc(0, c, k)
                if k == len(subs) - 1:
                    assert n == -1                    # overflows the CIF
                    continue
                assert n == lens[f, c, k], (f, c, k, n)
                assert np.array_equal(got, msc_bytes[f][c][k]), (f, c, k)
    assert dec.kernel_launches() == 4 * len(frames)
    dec.close()


LAYOUTS = [
    [(0, 48, 0, 0, 2, 0), (48, 48, 0, 0, 2, 0), (96, 54, 0, 0, 0, 1), (150, 35, 1, 4, 0, 0), (185, 8, 0, 0, 1, 0), (200, 84, 1, 33, 0, 0)],
    [(10, 12, 0, 0, 0, 0), (30, 16, 0, 0, 1, 0), (60, 40, 0, 0, 3, 0), (100, 42, 0, 0, 1, 1), (150, 64, 1, 34, 0, 0)],
    [(0, 140, 1, 37, 0, 0), (140, 18, 0, 0, 2, 1), (200, 15, 0, 0, 3, 1), (856, 8, 0, 0, 1, 0)],
    [],
]


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
def test_multi_stream_vs_oracle(ens, oracle, mode):
    """several streams with different sub-channel sets, noise, a stream that skips a frame, a sub-channel added mid-way"""
    n_streams = 4
    nb_cifs = ensgen.MODE_GEOM[mode][0]
    n_frames = {1: 6, 2: 20, 3: 19, 4: 10}[mode]
    layouts = [[a for a in LAYOUTS[s]] for s in range(n_streams)]
    subs_o = [[oracle.subchannel(*a) for a in lay] for lay in layouts]
    txs = [ensgen.EnsembleTx(mode, subs_o[s], seed=10 * mode + s, sigma=[0.0, 60.0, 90.0, 40.0][s]) for s in range(n_streams)]
    dec = ens.EnsembleDecoder(mode, n_streams=n_streams, max_subchannels=8)
    for s in range(n_streams):
        dec.set_subchannels(s, [ens.subchannel(*a) for a in layouts[s]])
    mscs = [[oracle.OracleMscDecoder(sc) for sc in subs_o[s]] for s in range(n_streams)]
    geom = ensgen.MODE_GEOM[mode]
    skip = (n_frames // 2, 1)             # (frame, stream): no frame for that stream in that call
    add_at = n_frames // 3                # stream 2 gains a sub-channel here; the old ones keep their history
    for f in range(n_frames):
        if f == add_at:
            layouts[2] = layouts[2] + [(400, 24, 0, 0, 2, 0)]
            mscs[2].append(oracle.OracleMscDecoder(oracle.subchannel(*layouts[2][-1])))
            dec.set_subchannels(2, [ens.subchannel(*a) for a in layouts[2]])
        frames = np.stack([tx.next_frame(corrupt_fibs=[(0, 1)] if f == 1 else ()) for tx in txs])
        present = np.ones(n_streams, np.uint8)
        if (f, 1) == skip:
            present[1] = 0
        dec.decode_frames(frames, present)
        for s in range(n_streams):
            if not present[s]:
                for k in range(len(layouts[s])):
                    assert dec.read_msc(s, 0, k)[1] == 0
                continue
            fb, fv, fe = [], [], []
            for c in range(nb_cifs):
                b, v, e = oracle.fic_decode_group(frames[s, c * geom[2]:(c + 1) * geom[2]], geom[3])
                fb.append(b)
                fv.append(v)
                fe.append(e)
            msc = []
            for c in range(nb_cifs):
                cif = frames[s, geom[1] + c * geom[4]: geom[1] + (c + 1) * geom[4]]
                msc.append([d.decode_cif(cif) for d in mscs[s]])
            _check_stream(dec, s, (fb, fv, fe, msc), len(layouts[s]), (mode, f))
    # noiseless stream 0 decodes what was sent 15 CIFs earlier
    b, n, _ = dec.read_msc(0, nb_cifs - 1, 0)
    assert n == 192 and np.array_equal(b, txs[0].sent_payloads[n_frames * nb_cifs - 16][0])
    dec.close()


def test_long_subchannel_spills(ens, oracle):
    """a 416 CU UEP sub-channel (9222 trellis steps) does not fit the shared-memory decision window: global spill path"""
    args = (100, 416, 1, 63, 0, 0)
    small = (0, 24, 0, 0, 2, 0)
    dec = ens.EnsembleDecoder(2, n_streams=2, max_subchannels=2)
    dec.set_subchannels(-1, [ens.subchannel(*small), ens.subchannel(*args)])
    subs_o = [oracle.subchannel(*small), oracle.subchannel(*args)]
    txs = [ensgen.EnsembleTx(2, subs_o, seed=5 + s, sigma=50.0 * s) for s in range(2)]
    mscs = [[oracle.OracleMscDecoder(sc) for sc in subs_o] for _ in range(2)]
    for f in range(18):
        frames = np.stack([tx.next_frame() for tx in txs])
        dec.decode_frames(frames)
        for s in range(2):
            cif = frames[s, 2304:]
            for k in range(2):
                wb, we = mscs[s][k].decode_cif(cif)
                b, n, e = dec.read_msc(s, 0, k)
                assert n == wb.size and np.array_equal(b, wb), (f, s, k, n, wb.size)
                if wb.size:
                    assert e == we
    assert dec.read_msc(0, 0, 1)[1] == 1152      # 384 kbit/s x 24 ms; 9222 trellis steps > the ~7100-step window
    dec.close()


def test_demod_to_bytes_on_device(ens, oracle, pkg):
    """OFDM soft bits stay on the device: dab_ofdm_device_bits -> dab_ensemble_decode_frames_device.  The oracle decodes the
    same soft bits (copied back through the frame callback) on the CPU."""
    import dabgen
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    n_streams, n_frames = 3, 7
    d = ofdm.OfdmDemodBatch(1, n_streams=n_streams, max_block_samples=196608)
    layout = [(0, 48, 0, 0, 2, 0), (48, 35, 1, 4, 0, 0), (100, 27, 0, 0, 0, 1)]
    dec = ens.EnsembleDecoder(1, n_streams=n_streams, max_subchannels=4)
    dec.set_subchannels(-1, [ens.subchannel(*a) for a in layout])
    xs = [dabgen.make_stream(1, n_frames + 1, seed=40 + s, cfo_hz=200.0 * s, start=1000 * s, snr_db=12.0) for s in range(n_streams)]
    mscs = [[oracle.OracleMscDecoder(oracle.subchannel(*a)) for a in layout] for _ in range(n_streams)]
    seen = [0] * n_streams
    decoded_any = 0
    for off in range(0, xs[0].size, 196608):
        d.process_batch([x[off:off + 196608] for x in xs])
        d.sync()
        d_bits, n_bits, slots, d_fic = d.device_bits()
        assert n_bits == 230400
        for slot in range(slots):
            dec.decode_frames_device(d_bits + slot * n_bits, slots * n_bits, d_fic, slot)
            dec.sync()
            for s in range(n_streams):
                frames = d.frames[s]
                if len(frames) <= seen[s] + slot:
                    continue
                bits = frames[seen[s] + slot][1]
                got_b, got_v, got_e = dec.read_fic(s)
                for c in range(4):
                    b, v, e = oracle.fic_decode_group(bits[c * 2304:(c + 1) * 2304])
                    assert np.array_equal(got_b[c], b) and np.array_equal(got_v[c], v) and int(got_e[c]) == e
                    cif = bits[9216 + c * 55296: 9216 + (c + 1) * 55296]
                    for k in range(len(layout)):
                        wb, we = mscs[s][k].decode_cif(cif)
                        gb, n, ge = dec.read_msc(s, c, k)
                        assert n == wb.size and np.array_equal(gb, wb)
                        decoded_any += int(n > 0)
        for s in range(n_streams):
            seen[s] = len(d.frames[s])
    assert min(seen) >= n_frames - 2 and decoded_any > 0
    d.close()
    dec.close()


def test_schedules_are_garbage_collected_across_reconfigurations(ens, oracle):
    """repeated sub-channel reconfiguration must not grow the schedule tables without bound.  Every EEP-A profile /
    length combination passes through one stream (539 distinct schedules); the table never holds more than the FIC's schedule and
    the ones in use, and afterwards the decoder still matches the oracle on a fresh layout."""
    dec = ens.EnsembleDecoder(1, n_streams=1, max_subchannels=8)
    n = 0
    for length in range(6, 866, 2):                      # EEP A needs a multiple of {12, 8, 6, 4}; the schedule differs per length
        for level, type_b in ((0, 0), (1, 0), (2, 0), (3, 0)):
            unit = (12, 8, 6, 4)[level]
            if length % unit:
                continue
            dec.set_subchannels(0, [ens.subchannel(0, length, 0, 0, level, type_b)])
            n += 1
            assert dec.L.dab_ensemble_schedule_count(dec.h) == 2        # FIC + the one sub-channel
    assert n > 500
    layout = LAYOUTS[1]
    subs_o = [oracle.subchannel(*a) for a in layout]
    dec.set_subchannels(0, [ens.subchannel(*a) for a in layout])
    tx = ensgen.EnsembleTx(1, subs_o, seed=91, sigma=50.0)
    frames = np.stack([tx.next_frame() for _ in range(5)])
    want = ensgen.oracle_decode_stream(1, subs_o, frames)
    for f in range(5):
        dec.decode_frames(frames[f])
        _check_stream(dec, 0, want[f], len(layout), f"frame {f}")
    dec.close()


def test_decode_stream_overlaps_the_next_frame(ens, oracle, pkg):
    """dab_ensemble_set_decode_stream: only the ingest of a frame stays on the stream that produced the soft bits; de-interleave, Viterbi,
    descramble, CRC and commit run on a second stream while the demodulator already overwrites its soft-bit buffer with the next
    frame.  A decoder in that mode must report, frame after frame, exactly what a decoder that is synchronised after every call reports."""
    import dabgen
    import torch
    ofdm = importlib.import_module("dab-radio_b200.ofdm")
    n_streams, n_frames, block = 4, 8, 196104   # one frame per call at most: one soft-bit slot per stream
    d = ofdm.OfdmDemodBatch(1, n_streams=n_streams, max_block_samples=block)
    layout = [(0, 48, 0, 0, 2, 0), (48, 35, 1, 4, 0, 0), (100, 27, 0, 0, 0, 1), (200, 96, 0, 0, 3, 0)]
    serial = ens.EnsembleDecoder(1, n_streams=n_streams, max_subchannels=4)
    split = ens.EnsembleDecoder(1, n_streams=n_streams, max_subchannels=4)
    for dec in (serial, split):
        dec.set_subchannels(-1, [ens.subchannel(*a) for a in layout])
    work, side = torch.cuda.Stream(), torch.cuda.Stream()
    d.set_cuda_stream(work.cuda_stream)
    split.set_cuda_stream(work.cuda_stream)
    split.set_decode_stream(side.cuda_stream)
    serial.set_cuda_stream(work.cuda_stream)
    xs = [dabgen.make_stream(1, n_frames + 1, seed=70 + s, cfo_hz=-300.0 * s, start=777 * s + 5, snr_db=14.0) for s in range(n_streams)]

    def snapshot(dec):
        out = []
        for s in range(n_streams):
            b, v, e = dec.read_fic(s)
            out.append((b.copy(), v.copy(), np.array(e).copy()))
            for c in range(4):
                for k in range(len(layout)):
                    gb, n, ge = dec.read_msc(s, c, k)
                    out.append((np.array(gb).copy(), int(n), int(ge)))
        return out

    def same(a, b):
        return len(a) == len(b) and all(all(np.array_equal(np.asarray(x), np.asarray(y)) for x, y in zip(p, q)) for p, q in zip(a, b))

    expected = None
    decoded_calls = 0
    for off in range(0, xs[0].size - block + 1, block):
        d.process_batch([x[off:off + block] for x in xs])   # overwrites the soft bits the split decoder ingested in the last round
        d.sync()
        if expected is not None:
            assert same(snapshot(split), expected), f"call at sample {off}: the overlapped decoder differs from the synchronised one"
        d_bits, n_bits, slots, d_fic = d.device_bits()
        assert slots == 1
        serial.decode_frames_device(d_bits, n_bits, d_fic, 0)
        serial.sync()
        expected = snapshot(serial)
        split.decode_frames_device(d_bits, n_bits, d_fic, 0)   # not synchronised: the next process_batch races with it by design
        decoded_calls += 1
    assert same(snapshot(split), expected)
    assert decoded_calls >= n_frames and sum(len(f) for f in d.frames) >= n_streams * (n_frames - 2)
    assert any(int(item[1]) > 0 for item in expected if isinstance(item[1], int))   # sub-channel bytes did come out
    for h in (d, serial, split):
        h.close()
