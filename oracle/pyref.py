"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/_ref/libdabref.so (the reference's own sources, compiled
unmodified by oracle/Makefile through oracle/ref_harness.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product path (dab-radio_b200/) never does.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libdabref.so")            # accuracy build (double-precision FFT stand-in): parity tests
REF_FAST_SO = os.path.join(HERE, "_ref", "libdabref_fast.so")  # timing build (vectorised float FFT stand-in): CPU baselines


class FrameInfo(C.Structure):
    _fields_ = [
        ("frame_start", C.c_int64),
        ("fine_time_offset", C.c_int32),
        ("total_desync", C.c_int32),
        ("coarse_offset", C.c_float),
        ("fine_offset_used", C.c_float),
        ("fine_offset_after", C.c_float),
        ("signal_average", C.c_float),
    ]


class OfdmState(C.Structure):
    _fields_ = [
        ("state", C.c_int32),
        ("fine_time_offset", C.c_int32),
        ("total_frames_read", C.c_int32),
        ("total_frames_desync", C.c_int32),
        ("signal_average", C.c_float),
        ("fine_offset", C.c_float),
        ("coarse_offset", C.c_float),
        ("pad", C.c_int32),
    ]


_lib = None


def available():
    return os.path.exists(REF_SO)


_fast = None


def fast_lib():
    """timing build of the reference (same sources, -O3, vectorised single-precision FFT stand-in)"""
    global _fast
    if _fast is None:
        if not os.path.exists(REF_FAST_SO):
            raise FileNotFoundError(f"{REF_FAST_SO} missing: run `make -C oracle ref` where /root/reference exists")
        _fast = _bind(C.CDLL(REF_FAST_SO))
    return _fast


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise FileNotFoundError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
    _lib = _bind(C.CDLL(REF_SO))
    return _lib


def _bind(L):
    vp, u64, i32, f32 = C.c_void_p, C.c_uint64, C.c_int, C.c_float
    fp = C.POINTER(C.c_float)
    L.ref_get_params.argtypes = [i32, C.POINTER(u64)]
    L.ref_get_prs.argtypes = [i32, vp]
    L.ref_get_mapper.argtypes = [i32, vp]
    L.ref_apply_pll.argtypes = [vp, vp, u64, f32, f32]
    L.ref_conj_mul_sum.argtypes = [vp, vp, u64, fp]
    L.ref_modulate.argtypes = [i32, vp, u64, vp, u64]
    L.ref_ofdm_create.argtypes = [i32, i32, i32]
    L.ref_ofdm_create.restype = vp
    L.ref_ofdm_destroy.argtypes = [vp]
    L.ref_ofdm_process.argtypes = [vp, vp, u64, i32]
    L.ref_ofdm_frames_done.argtypes = [vp]
    L.ref_ofdm_frames_done.restype = u64
    L.ref_ofdm_frame_bits.argtypes = [vp]
    L.ref_ofdm_frame_bits.restype = u64
    L.ref_ofdm_get_frame.argtypes = [vp, u64, C.POINTER(FrameInfo), vp]
    L.ref_ofdm_get_state.argtypes = [vp, C.POINTER(OfdmState)]
    for name in ("ref_ofdm_get_frame_fft", "ref_ofdm_get_frame_data_vec", "ref_ofdm_get_impulse_response",
                 "ref_ofdm_get_coarse_freq_response", "ref_ofdm_get_correlation_time_buffer"):
        getattr(L, name).argtypes = [vp, vp]
    L.ref_ofdm_bench.argtypes = [i32, i32, i32, vp, u64, u64, i32, C.POINTER(u64)]
    L.ref_ofdm_bench.restype = C.c_double
    L.ref_ofdm_pool_create.argtypes = [i32, i32, i32]
    L.ref_ofdm_pool_create.restype = vp
    L.ref_ofdm_pool_destroy.argtypes = [vp]
    L.ref_ofdm_pool_frames.argtypes = [vp]
    L.ref_ofdm_pool_frames.restype = u64
    L.ref_ofdm_pool_run.argtypes = [vp, vp, u64, u64, i32]
    L.ref_ofdm_pool_run.restype = C.c_double
    L.ref_ofdm_pool_run_multi.argtypes = [vp, C.POINTER(vp), u64, u64, i32]
    L.ref_ofdm_pool_run_multi.restype = C.c_double
    L.ref_fft_bench.argtypes = [i32, i32]
    L.ref_fft_bench.restype = C.c_double
    L.ref_vit_create.restype = vp
    L.ref_vit_destroy.argtypes = [vp]
    L.ref_vit_set_traceback_length.argtypes = [vp, u64]
    L.ref_vit_get_traceback_length.argtypes = [vp]
    L.ref_vit_get_traceback_length.restype = u64
    L.ref_vit_get_current_decoded_bit.argtypes = [vp]
    L.ref_vit_get_current_decoded_bit.restype = u64
    L.ref_vit_reset.argtypes = [vp, u64]
    L.ref_vit_update.argtypes = [vp, vp, u64, vp, u64, u64]
    L.ref_vit_update.restype = u64
    L.ref_vit_chainback.argtypes = [vp, vp, u64, u64]
    L.ref_vit_chainback.restype = u64
    L.ref_vit_decode_job.argtypes = [vp, vp, u64, vp, vp, vp, C.c_uint32, vp, u64, C.POINTER(u64)]
    L.ref_vit_decode_job.restype = u64
    L.ref_vit_bench.argtypes = [i32, vp, u64, u64, vp, vp, vp, C.c_uint32, u64, vp, u64]
    L.ref_vit_bench.restype = C.c_double
    L.ref_build_info.restype = C.c_char_p
    L.ref_fic_create.argtypes = [u64, u64]
    L.ref_fic_create.restype = vp
    L.ref_fic_destroy.argtypes = [vp]
    L.ref_fic_decode_group.argtypes = [vp, vp, u64, u64, vp, vp]
    L.ref_msc_create.argtypes = [i32] * 6
    L.ref_msc_create.restype = vp
    L.ref_msc_destroy.argtypes = [vp]
    L.ref_msc_decode_cif.argtypes = [vp, vp, u64, vp, u64]
    L.ref_msc_decode_cif.restype = C.c_int64
    L.ref_deint_create.argtypes = [i32]
    L.ref_deint_create.restype = vp
    L.ref_deint_destroy.argtypes = [vp]
    L.ref_deint_push.argtypes = [vp, vp, vp, u64]
    L.ref_scrambler_bytes.argtypes = [C.c_uint16, vp, u64]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def params(mode):
    out = (C.c_uint64 * 6)()
    if lib().ref_get_params(mode, out) != 0:
        raise RuntimeError("invalid transmission mode")
    keys = ("nb_frame_symbols", "nb_symbol_period", "nb_null_period", "nb_cyclic_prefix", "nb_fft", "nb_data_carriers")
    return dict(zip(keys, (int(v) for v in out)))


def prs(mode):
    p = params(mode)
    out = np.zeros(p["nb_fft"], np.complex64)
    assert lib().ref_get_prs(mode, _p(out)) == 0
    return out


def mapper(mode):
    p = params(mode)
    out = np.zeros(p["nb_data_carriers"], np.int32)
    assert lib().ref_get_mapper(mode, _p(out)) == 0
    return out


def apply_pll(x, freq_norm, dt_norm=0.0):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.empty_like(x)
    lib().ref_apply_pll(_p(x), _p(y), x.size, freq_norm, dt_norm)
    return y


def conj_mul_sum(x0, x1):
    x0 = np.ascontiguousarray(x0, np.complex64)
    x1 = np.ascontiguousarray(x1, np.complex64)
    out = (C.c_float * 2)()
    lib().ref_conj_mul_sum(_p(x0), _p(x1), x0.size, out)
    return np.complex64(complex(out[0], out[1]))


def modulate(mode, data_bytes):
    p = params(mode)
    n = p["nb_null_period"] + p["nb_symbol_period"] * p["nb_frame_symbols"]
    data_bytes = np.ascontiguousarray(data_bytes, np.uint8)
    out = np.zeros(n, np.complex64)
    rc = lib().ref_modulate(mode, _p(data_bytes), data_bytes.size, _p(out), n)
    if rc != 0:
        raise RuntimeError(f"ref_modulate failed rc={rc}")
    return out


class RefOfdmDemod:
    """The reference OFDM_Demod behind the determinism harness (real-time order)."""

    def __init__(self, mode, nb_threads=1, collect=True):
        self.L = lib()
        self.mode = mode
        self.h = self.L.ref_ofdm_create(mode, nb_threads, 1 if collect else 0)
        if not self.h:
            raise RuntimeError("ref_ofdm_create failed")
        self.frame_bits = int(self.L.ref_ofdm_frame_bits(self.h))

    def process(self, iq, realtime=True):
        iq = np.ascontiguousarray(iq, np.complex64)
        return self.L.ref_ofdm_process(self.h, _p(iq), iq.size, 1 if realtime else 0)

    def process_blocks(self, iq, block, realtime=True):
        iq = np.ascontiguousarray(iq, np.complex64)
        for off in range(0, iq.size, block):
            self.process(iq[off:off + block], realtime)

    def frames_done(self):
        return int(self.L.ref_ofdm_frames_done(self.h))

    def frame(self, i):
        info = FrameInfo()
        bits = np.zeros(self.frame_bits, np.int8)
        if self.L.ref_ofdm_get_frame(self.h, i, C.byref(info), _p(bits)) != 0:
            raise IndexError(i)
        d = {k: getattr(info, k) for k, _ in FrameInfo._fields_}
        return d, bits

    def state(self):
        s = OfdmState()
        self.L.ref_ofdm_get_state(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in OfdmState._fields_ if k != "pad"}

    # GUI getters (ofdm_demodulator.h:133-139) of the latest frame / synchronisation
    def _tap(self, fn, n, dtype):
        out = np.zeros(n, dtype)
        getattr(self.L, fn)(self.h, _p(out))
        return out

    def frame_fft(self):
        p = params(self.mode)
        return self._tap("ref_ofdm_get_frame_fft", (p["nb_frame_symbols"] + 1) * p["nb_fft"], np.complex64)

    def frame_data_vec(self):
        p = params(self.mode)
        return self._tap("ref_ofdm_get_frame_data_vec", (p["nb_frame_symbols"] - 1) * p["nb_data_carriers"], np.complex64)

    def impulse_response(self):
        return self._tap("ref_ofdm_get_impulse_response", params(self.mode)["nb_fft"], np.float32)

    def coarse_freq_response(self):
        return self._tap("ref_ofdm_get_coarse_freq_response", params(self.mode)["nb_fft"], np.float32)

    def correlation_time_buffer(self):
        p = params(self.mode)
        return self._tap("ref_ofdm_get_correlation_time_buffer", p["nb_null_period"] + p["nb_symbol_period"], np.complex64)

    def close(self):
        if self.h:
            self.L.ref_ofdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefViterbi:
    """The reference DAB_Viterbi_Decoder (AVX2 u16 build)."""

    def __init__(self):
        self.L = lib()
        self.h = self.L.ref_vit_create()

    def set_traceback_length(self, n):
        self.L.ref_vit_set_traceback_length(self.h, n)

    def get_traceback_length(self):
        return int(self.L.ref_vit_get_traceback_length(self.h))

    def get_current_decoded_bit(self):
        return int(self.L.ref_vit_get_current_decoded_bit(self.h))

    def reset(self, start_state=0):
        self.L.ref_vit_reset(self.h, start_state)

    def update(self, soft, code, n_out):
        soft = np.ascontiguousarray(soft, np.int8)
        code = np.ascontiguousarray(code, np.uint8)
        return int(self.L.ref_vit_update(self.h, _p(soft), soft.size, _p(code), code.size, n_out))

    def chainback(self, nbytes, end_state=0):
        out = np.zeros(nbytes, np.uint8)
        err = int(self.L.ref_vit_chainback(self.h, _p(out), nbytes, end_state))
        return out, err

    def close(self):
        if self.h:
            self.L.ref_vit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RefOfdmPool:
    """n independent reference OFDM_Demod instances, one feeder thread each (bench.py --impl reference / cpu_baseline)."""

    def __init__(self, mode, n_instances, threads_each=1, fast=True):
        self.L = fast_lib() if fast else lib()
        self.h = self.L.ref_ofdm_pool_create(mode, n_instances, threads_each)
        if not self.h:
            raise RuntimeError("ref_ofdm_pool_create failed")
        self.n_instances = n_instances

    def run(self, iq, block, repeats):
        iq = np.ascontiguousarray(iq, np.complex64)
        return float(self.L.ref_ofdm_pool_run(self.h, _p(iq), iq.size, block, repeats))

    def run_multi(self, iqs, block, repeats):
        """instance i consumes iqs[i] (equal lengths) `repeats` times"""
        arrs = [np.ascontiguousarray(x, np.complex64) for x in iqs]
        assert len(arrs) == self.n_instances and all(a.size == arrs[0].size for a in arrs)
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        return float(self.L.ref_ofdm_pool_run_multi(self.h, ptrs, arrs[0].size, block, repeats))

    def fft_seconds(self, nfft, reps):
        """seconds for reps forward FFTs of nfft points through the FFT this build of the reference is linked against"""
        return float(self.L.ref_fft_bench(nfft, reps))

    def frames(self):
        return int(self.L.ref_ofdm_pool_frames(self.h))

    def close(self):
        if self.h:
            self.L.ref_ofdm_pool_destroy(self.h)
            self.h = None


# ---------------------------------------------------------------------------------------------- FIC / MSC decoders
def scrambler_bytes(n, syncword=0xFFFF):
    out = np.zeros(n, np.uint8)
    lib().ref_scrambler_bytes(syncword, _p(out), n)
    return out


class RefFicDecoder:
    """The reference FIC_Decoder (src/dab/fic/fic_decoder.cpp:53-116)."""

    def __init__(self, nb_encoded_bits=2304, nb_fibs=3):
        self.L = lib()
        self.n, self.fibs = nb_encoded_bits, nb_fibs
        self.h = self.L.ref_fic_create(nb_encoded_bits, nb_fibs)

    def decode_group(self, bits, cif_index=0):
        bits = np.ascontiguousarray(bits, np.int8)
        out = np.zeros(self.n // 24, np.uint8)
        valid = np.zeros(self.fibs, np.uint8)
        self.L.ref_fic_decode_group(self.h, _p(bits), bits.size, cif_index, _p(out), _p(valid))
        return out, valid

    def __del__(self):
        try:
            if self.h:
                self.L.ref_fic_destroy(self.h)
                self.h = None
        except Exception:
            pass


class RefDeinterleaver:
    """The reference CIF_Deinterleaver (src/dab/msc/cif_deinterleaver.cpp:21-70)."""

    def __init__(self, nb_bits):
        self.L = lib()
        self.n = nb_bits
        self.h = self.L.ref_deint_create(nb_bits // 8)

    def push(self, bits):
        bits = np.ascontiguousarray(bits, np.int8)
        out = np.zeros(self.n, np.int8)
        ok = self.L.ref_deint_push(self.h, _p(bits), _p(out), self.n)
        return out if ok else None

    def __del__(self):
        try:
            if self.h:
                self.L.ref_deint_destroy(self.h)
                self.h = None
        except Exception:
            pass


class RefMscDecoder:
    """The reference MSC_Decoder (src/dab/msc/msc_decoder.cpp:27-170) for one sub-channel."""

    def __init__(self, start_address, length, is_uep=False, uep_prot_index=0, eep_prot_level=0, eep_type_b=False):
        self.L = lib()
        self.length = length
        self.h = self.L.ref_msc_create(int(start_address), int(length), int(bool(is_uep)), int(uep_prot_index), int(eep_prot_level),
                                       int(bool(eep_type_b)))

    def decode_cif(self, cif_bits):
        cif_bits = np.ascontiguousarray(cif_bits, np.int8)
        out = np.zeros(self.length * 8 + 8, np.uint8)
        n = int(self.L.ref_msc_decode_cif(self.h, _p(cif_bits), cif_bits.size, _p(out), out.size))
        return out[:max(n, 0)].copy()

    def __del__(self):
        try:
            if self.h:
                self.L.ref_msc_destroy(self.h)
                self.h = None
        except Exception:
            pass
