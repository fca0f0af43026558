"""GPU parity: CUDA Viterbi (through the C ABI) vs the oracle restatement of DAB_Viterbi_Decoder -- bit-exact bytes and
equal u64 path error on identical soft bits (north_star)."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _segments(po, spec):
    return [(po.puncture_code(pi) if pi else po.PI_X, n) for pi, n in spec]


def _make_case(po, rng, spec, n_bytes, sigma, garbage=None):
    segs = _segments(po, spec)
    if garbage is None:
        data = rng.integers(0, 256, n_bytes, dtype=np.uint8)
        tx = po.puncture(po.conv_encode(data), segs)
        rx = np.clip(np.rint(tx + sigma * rng.standard_normal(tx.size)), -128, 127).astype(np.int8)
    else:
        n_soft = sum(int(np.resize(c, n // 4).astype(np.int64).sum()) for c, n in segs)
        rx = garbage(n_soft)
    return segs, rx


FIC = [(16, 128 * 21), (15, 128 * 3), (0, 24)]
EEP_3A_48CU = [(8, 128 * 45), (7, 128 * 3), (0, 24)]      # 6n-3, 3 blocks with n = 8 (subchannel_protection_tables.h:121-126)
UEP_ROW0 = [(5, 128 * 3), (3, 128 * 4), (2, 128 * 17), (0, 24)]


@pytest.fixture(params=["warp", "lanes"])
def vit(pkg, request, monkeypatch):
    """every parity case runs through both kernel forms: one trellis per warp (viterbi_kernel) and one per thread
    (viterbi_lanes_kernel, the bulk form) -- DAB_B200_VITERBI_LANES is read at every launch"""
    monkeypatch.setenv("DAB_B200_VITERBI_LANES", "1" if request.param == "lanes" else "0")
    v = importlib.import_module("dab-radio_b200.viterbi")
    return v


def _run_batch(vit, po, cases, n_out_bytes_list):
    vb = vit.ViterbiBatch(0)
    jobs = np.zeros(len(cases), vit.capi.VIT_JOB_DTYPE)
    soft_all, out_off, soft_off = [], 0, 0
    sched_cache = {}
    for i, ((segs, rx), nb) in enumerate(zip(cases, n_out_bytes_list)):
        key = (tuple((tuple(c.tolist()), n) for c, n in segs), nb)
        if key not in sched_cache:
            sched_cache[key] = vb.add_schedule(vit.make_schedule(segs, nb))
        jobs[i] = (sched_cache[key], rx.size, soft_off, out_off)
        soft_all.append(rx)
        soft_off += rx.size
        out_off += nb
    out, err, st = vb.decode_batch(np.concatenate(soft_all), jobs, out_off)
    assert np.all(st == 0)
    o = po.OracleViterbi()
    off = 0
    for i, ((segs, rx), nb) in enumerate(zip(cases, n_out_bytes_list)):
        o.set_traceback_length(nb * 8)
        ref_out, ref_err, used = o.decode_job(rx, segs, nb)
        assert used == rx.size
        got = out[off:off + nb]
        assert np.array_equal(got, ref_out), f"job {i}: {np.count_nonzero(got != ref_out)} bytes differ"
        assert int(err[i]) == ref_err, f"job {i}: path error {int(err[i])} != {ref_err}"
        off += nb
    assert vb.kernel_launches() >= 1
    vb.close()


def test_fic_noise_sweep(vit, oracle):
    rng = np.random.default_rng(11)
    cases = [_make_case(oracle, rng, FIC, 96, sigma) for sigma in (0, 40, 80, 120, 200) for _ in range(40)]
    _run_batch(vit, oracle, cases, [96] * len(cases))


def test_fic_noiseless_roundtrip(vit, oracle):
    """reference vendor test run_punctured_decoder.cpp:139-191: FIC schedule, no noise, zero bit errors"""
    rng = np.random.default_rng(3)
    vb = vit.ViterbiBatch(0)
    sid = vb.add_schedule(vit.fic_schedule())
    n = 64
    data = rng.integers(0, 256, (n, 96), dtype=np.uint8)
    segs = _segments(oracle, FIC)
    soft = np.concatenate([oracle.puncture(oracle.conv_encode(d), segs) for d in data])
    jobs = np.zeros(n, vit.capi.VIT_JOB_DTYPE)
    for i in range(n):
        jobs[i] = (sid, 2304, i * 2304, i * 96)
    out, err, st = vb.decode_batch(soft, jobs, n * 96)
    assert np.array_equal(out.reshape(n, 96), data)
    assert np.all(err == 792 * 127)  # every punctured position costs |+-127 - 0| on the surviving path
    vb.close()


def test_mixed_schedules_one_launch(vit, oracle):
    rng = np.random.default_rng(5)
    cases, nbytes = [], []
    for spec, nb in ((FIC, 96), (EEP_3A_48CU, 192), (UEP_ROW0, 96)):
        for sigma in (0, 60, 130):
            for _ in range(6):
                cases.append(_make_case(oracle, rng, spec, nb, sigma))
                nbytes.append(nb)
    _run_batch(vit, oracle, cases, nbytes)


def test_adversarial_ties_and_saturation(vit, oracle):
    """all-zero (every comparison ties), all -128 / +127 and alternating extremes (fast metric growth, renormalisation,
    saturating adds), random garbage -- SURVEY.md 8(c)"""
    rng = np.random.default_rng(9)
    gens = [lambda n: np.zeros(n, np.int8), lambda n: np.full(n, -128, np.int8), lambda n: np.full(n, 127, np.int8),
            lambda n: rng.integers(-128, 128, n).astype(np.int8), lambda n: np.where(np.arange(n) % 2, 127, -128).astype(np.int8),
            lambda n: np.where(np.arange(n) % 5 < 2, -128, 127).astype(np.int8)]
    cases, nbytes = [], []
    for g in gens:
        for pi in (1, 8, 16, 24):
            L = 60
            cases.append(_make_case(oracle, rng, [(pi, 128 * L), (0, 24)], 0, 0, garbage=g))
            nbytes.append(128 * L // 4 // 8)
    _run_batch(vit, oracle, cases, nbytes)


def test_long_trellis_spills_to_global(vit, oracle):
    """a sub-channel longer than the shared-memory window (EEP 4-B filling most of a CIF)"""
    rng = np.random.default_rng(21)
    L = 24 * 12 - 3  # n = 12 -> 180 CU at 4-B
    spec = [(2, 128 * L), (1, 128 * 3), (0, 24)]
    nb = (128 * (L + 3)) // 4 // 8
    cases = [_make_case(oracle, rng, spec, nb, sigma) for sigma in (0, 100)]
    _run_batch(vit, oracle, cases, [nb] * 2)


def test_underrun_is_reported(vit, oracle):
    vb = vit.ViterbiBatch(0)
    sid = vb.add_schedule(vit.fic_schedule())
    jobs = np.zeros(1, vit.capi.VIT_JOB_DTYPE)
    jobs[0] = (sid, 2000, 0, 0)
    out, err, st = vb.decode_batch(np.zeros(2304, np.int8), jobs, 96, raise_on_job_error=False)
    assert st[0] == vit.capi.DAB_ERR_UNDERRUN
    vb.close()


def test_against_reference_golden_vectors(vit, oracle):
    """CUDA decoder vs outputs of the reference's own AVX2 decoder (tests/golden/viterbi.npz): bit-exact bytes and path error"""
    import goldenutil
    g = goldenutil.load("viterbi.npz")
    vb = vit.ViterbiBatch(0)
    keys = goldenutil.viterbi_cases()
    sched = {}
    jobs = np.zeros(len(keys), vit.capi.VIT_JOB_DTYPE)
    softs, want, off_s, off_o = [], [], 0, 0
    for i, key in enumerate(keys):
        segs, nbytes, soft, out, err = goldenutil.viterbi_case(g, oracle, key)
        name = key.split("__")[0]
        if name not in sched:
            sched[name] = vb.add_schedule(vit.make_schedule(segs, nbytes))
        jobs[i] = (sched[name], soft.size, off_s, off_o)
        softs.append(soft)
        want.append((off_o, out, err))
        off_s += soft.size
        off_o += nbytes
    got, errs, st = vb.decode_batch(np.concatenate(softs), jobs, off_o)
    assert np.all(st == 0)
    for i, (o, out, err) in enumerate(want):
        assert np.array_equal(got[o:o + out.size], out), keys[i]
        assert int(errs[i]) == err, keys[i]
    vb.close()


def test_unaligned_offsets_and_ragged_warps(vit, oracle):
    """jobs start at odd byte offsets (the bulk kernel reads aligned words and funnel-shifts), 37 jobs = one full + one
    ragged warp, three schedules of different length interleaved"""
    rng = np.random.default_rng(77)
    vb = vit.ViterbiBatch(0)
    specs = [(FIC, 96), (EEP_3A_48CU, 192), (UEP_ROW0, 96)]
    sids = [vb.add_schedule(vit.make_schedule(_segments(oracle, sp), nb)) for sp, nb in specs]
    n = 37
    jobs = np.zeros(n, vit.capi.VIT_JOB_DTYPE)
    chunks, cases, soft_off, out_off = [], [], 0, 0
    for i in range(n):
        k = i % 3
        segs, rx = _make_case(oracle, rng, specs[k][0], specs[k][1], (0, 70, 140)[(i // 3) % 3])
        pad = int(rng.integers(0, 4)) | 1
        chunks += [np.full(pad, -128, np.int8), rx]
        soft_off += pad
        jobs[i] = (sids[k], rx.size, soft_off, out_off)
        cases.append((segs, rx, specs[k][1], out_off))
        soft_off += rx.size
        out_off += specs[k][1]
    out, err, st = vb.decode_batch(np.concatenate(chunks), jobs, out_off)
    assert np.all(st == 0)
    o = oracle.OracleViterbi()
    for i, (segs, rx, nb, off) in enumerate(cases):
        o.set_traceback_length(nb * 8)
        ref_out, ref_err, _ = o.decode_job(rx, segs, nb)
        assert np.array_equal(out[off:off + nb], ref_out), i
        assert int(err[i]) == ref_err, i
    vb.close()


def test_bulk_batch_default_dispatch(pkg, oracle, monkeypatch):
    """8192 trellises in one call: the default dispatch takes the one-trellis-per-thread form; every job must equal the
    one-trellis-per-warp form bit for bit, and a sample of them the oracle"""
    v = importlib.import_module("dab-radio_b200.viterbi")
    rng = np.random.default_rng(123)
    specs = [(FIC, 96), (EEP_3A_48CU, 192)]
    base = [[_make_case(oracle, rng, sp, nb, sigma) for sigma in (0, 60, 110, 160) for _ in range(4)] for sp, nb in specs]
    n = 8192
    vb = v.ViterbiBatch(0)
    sids = [vb.add_schedule(v.make_schedule(_segments(oracle, sp), nb)) for sp, nb in specs]
    jobs = np.zeros(n, v.capi.VIT_JOB_DTYPE)
    chunks, meta, soft_off, out_off = [], [], 0, 0
    for i in range(n):
        k = int(rng.integers(0, 2))
        segs, rx = base[k][int(rng.integers(0, len(base[k])))]
        # perturb a few symbols so that the jobs are not copies of one another
        rx = rx.copy()
        idx = rng.integers(0, rx.size, 8)
        rx[idx] = rng.integers(-128, 128, 8).astype(np.int8)
        jobs[i] = (sids[k], rx.size, soft_off, out_off)
        chunks.append(rx)
        meta.append((segs, rx, specs[k][1], out_off))
        soft_off += rx.size
        out_off += specs[k][1]
    soft = np.concatenate(chunks)
    results = {}
    for form in ("", "0", "1"):
        if form:
            monkeypatch.setenv("DAB_B200_VITERBI_LANES", form)
        else:
            monkeypatch.delenv("DAB_B200_VITERBI_LANES", raising=False)
        out, err, st = vb.decode_batch(soft, jobs, out_off)
        assert np.all(st == 0)
        results[form] = (out.copy(), err.copy())
    assert np.array_equal(results[""][0], results["0"][0]) and np.array_equal(results[""][1], results["0"][1])
    assert np.array_equal(results["1"][0], results["0"][0]) and np.array_equal(results["1"][1], results["0"][1])
    o = oracle.OracleViterbi()
    for i in range(0, n, 97):
        segs, rx, nb, off = meta[i]
        o.set_traceback_length(nb * 8)
        ref_out, ref_err, _ = o.decode_job(rx, segs, nb)
        assert np.array_equal(results[""][0][off:off + nb], ref_out), i
        assert int(results[""][1][i]) == ref_err, i
    vb.close()


def test_prepared_jobs_pack_short_trellises(pkg, oracle, monkeypatch):
    """dab_viterbi_prepare_jobs / decode_prepared: a job list uploaded once (long EEP trellises and half-length FIC groups
    interleaved, so that FIC groups share warps), device buffers in and out; every job equals dab_viterbi_decode_batch, in both
    kernel forms, across repeated calls"""
    import torch
    v = importlib.import_module("dab-radio_b200.viterbi")
    rng = np.random.default_rng(321)
    specs = [(FIC, 96), (EEP_3A_48CU, 192)]
    base = [[_make_case(oracle, rng, sp, nb, sigma) for sigma in (0, 70, 130) for _ in range(3)] for sp, nb in specs]
    n = 4608
    vb = v.ViterbiBatch(0)
    sids = [vb.add_schedule(v.make_schedule(_segments(oracle, sp), nb)) for sp, nb in specs]
    jobs = np.zeros(n, v.capi.VIT_JOB_DTYPE)
    chunks, soft_off, out_off = [], 0, 0
    for i in range(n):
        k = 0 if i % 3 == 0 else 1
        rx = base[k][int(rng.integers(0, len(base[k])))][1].copy()
        rx[rng.integers(0, rx.size, 6)] = rng.integers(-128, 128, 6).astype(np.int8)
        jobs[i] = (sids[k], rx.size, soft_off, out_off)
        chunks.append(rx)
        soft_off += rx.size
        out_off += specs[k][1]
    soft = np.concatenate(chunks)
    monkeypatch.setenv("DAB_B200_VITERBI_LANES", "0")
    want_out, want_err, st = vb.decode_batch(soft, jobs, out_off)
    assert np.all(st == 0)
    d_soft = torch.from_numpy(soft).cuda()
    plan = vb.prepare_jobs(jobs)
    assert plan >= 1
    for form in ("1", "0", "1"):
        monkeypatch.setenv("DAB_B200_VITERBI_LANES", form)
        d_out = torch.zeros(out_off, dtype=torch.uint8, device="cuda")
        d_err = torch.zeros(n, dtype=torch.int64, device="cuda")
        d_st = torch.full((n,), -99, dtype=torch.int32, device="cuda")
        vb.decode_prepared(plan, d_soft.data_ptr(), d_soft.numel(), d_out.data_ptr(), d_out.numel(), d_err.data_ptr(), d_st.data_ptr())
        vb.sync()
        assert int(d_st.abs().max().item()) == 0
        assert np.array_equal(d_out.cpu().numpy(), want_out), form
        assert np.array_equal(d_err.cpu().numpy().view(np.uint64), want_err), form
    vb.release_jobs(plan)
    with pytest.raises(v.capi.DabError):
        vb.decode_prepared(plan, d_soft.data_ptr(), d_soft.numel(), d_out.data_ptr(), d_out.numel())
    vb.close()


def test_start_and_end_state(vit, oracle):
    """reset(start_state) / chainback(end_state) other than 0 (DAB_Viterbi_Decoder::reset, ::chainback arguments), states on both
    sides of 32 (the two halves of a packed metric register), noisy input"""
    rng = np.random.default_rng(55)
    segs = _segments(oracle, [(8, 128 * 6), (0, 24)])
    nb = ((128 * 6 + 24) // 4 - 6) // 8   # 192 decoded bits = all 198 trellis steps minus the tail
    vb = vit.ViterbiBatch(0)
    cases = []
    jobs = np.zeros(12, vit.capi.VIT_JOB_DTYPE)
    chunks, soft_off = [], 0
    for i, (start, end) in enumerate([(0, 0), (5, 0), (37, 0), (63, 0), (0, 9), (0, 41), (0, 63), (17, 50), (33, 2), (62, 31), (1, 32), (32, 1)]):
        sid = vb.add_schedule(vit.make_schedule(segs, nb, start_state=start, end_state=end))
        n_soft = sum(int(np.resize(c, n // 4).astype(np.int64).sum()) for c, n in segs)
        rx = rng.integers(-128, 128, n_soft).astype(np.int8)
        jobs[i] = (sid, rx.size, soft_off, i * nb)
        chunks.append(rx)
        soft_off += rx.size
        cases.append((start, end, rx))
    out, err, st = vb.decode_batch(np.concatenate(chunks), jobs, len(cases) * nb)
    assert np.all(st == 0)
    o = oracle.OracleViterbi()
    o.set_traceback_length(nb * 8)
    for i, (start, end, rx) in enumerate(cases):
        o.reset(start)
        used = 0
        for code, n_out in segs:
            used += o.update(rx[used:], code, n_out)
        ref_out, ref_err = o.chainback(nb, end)
        assert np.array_equal(out[i * nb:(i + 1) * nb], ref_out), (i, start, end)
        assert int(err[i]) == ref_err, (i, start, end)
    vb.close()


def test_every_protection_profile_both_forms(pkg, oracle, monkeypatch):
    """all 64 UEP rows, EEP 1-A..4-A and 1-B..4-B at several sizes (incl. the 2-A n = 1 special case) and the FIC schedule, 70
    trellises of each in one bulk call, random soft bits: the one-trellis-per-thread form equals the one-trellis-per-warp form
    job by job (bytes and path error), and one trellis of every profile equals the oracle"""
    v = importlib.import_module("dab-radio_b200.viterbi")
    ens = importlib.import_module("dab-radio_b200.ensemble")
    rng = np.random.default_rng(2024)
    subs = [(ens.subchannel(0, oracle.uep_subchannel_size(i), True, i), oracle.subchannel(0, oracle.uep_subchannel_size(i), True, i)) for i in range(64)]
    for type_b, mult in ((False, (12, 8, 6, 4)), (True, (27, 21, 18, 15))):
        for level in range(4):
            for n in (1, 2, 5):
                L = mult[level] * n
                subs.append((ens.subchannel(0, L, False, 0, level, type_b), oracle.subchannel(0, L, False, 0, level, type_b)))
    vb = v.ViterbiBatch(0)
    profiles = []   # (schedule id, n_soft, n_out_bytes, oracle segments)
    for sg, so in subs:
        sch, n_soft = ens.subchannel_schedule(sg)
        if sch.n_out_bytes == 0:
            continue
        profiles.append((vb.add_schedule(sch), n_soft, int(sch.n_out_bytes), oracle.msc_segments(so)))
    profiles.append((vb.add_schedule(v.fic_schedule()), 2304, 96, _segments(oracle, FIC)))
    reps = 70
    n = len(profiles) * reps
    assert n >= 4096
    jobs = np.zeros(n, v.capi.VIT_JOB_DTYPE)
    soft_off = out_off = 0
    meta = []
    order = rng.permutation(n)            # profiles interleaved at random: the library has to group them itself
    for i in order:
        sid, n_soft, nb, segs = profiles[i % len(profiles)]
        jobs[i] = (sid, n_soft, soft_off, out_off)
        meta.append((i, soft_off, n_soft, out_off, nb, segs))
        soft_off += n_soft
        out_off += nb
    soft = rng.integers(-128, 128, soft_off).astype(np.int8)
    soft[rng.random(soft_off) < 0.5] //= 4   # mix of confident and weak symbols
    res = {}
    for form in ("0", "1"):
        monkeypatch.setenv("DAB_B200_VITERBI_LANES", form)
        out, err, st = vb.decode_batch(soft, jobs, out_off)
        assert np.all(st == 0)
        res[form] = (out, err)
    assert np.array_equal(res["0"][0], res["1"][0])
    assert np.array_equal(res["0"][1], res["1"][1])
    o = oracle.OracleViterbi()
    seen = set()
    for i, s_off, n_soft, o_off, nb, segs in meta:
        k = i % len(profiles)
        if k in seen:
            continue
        seen.add(k)
        o.set_traceback_length(sum(nn for _, nn in segs) // 4)
        ref_out, ref_err, used = o.decode_job(soft[s_off:s_off + n_soft], segs, nb)
        assert np.array_equal(res["1"][0][o_off:o_off + nb], ref_out), k
        assert int(res["1"][1][i]) == ref_err, k
    assert len(seen) == len(profiles)
    vb.close()
