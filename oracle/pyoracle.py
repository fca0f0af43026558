"""TEST INFRASTRUCTURE ONLY: ctypes loader for oracle/liboracle.so (oracle/dab_oracle.c, our plain-C restatement of
the reference hot path).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.
The product path (dab-radio_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")


class Params(C.Structure):
    _fields_ = [(k, C.c_int) for k in
                ("nb_frame_symbols", "nb_symbol_period", "nb_null_period", "nb_cyclic_prefix", "nb_fft", "nb_data_carriers")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Config(C.Structure):
    _fields_ = [
        ("signal_l1_update_beta", C.c_float), ("signal_l1_nb_samples", C.c_int), ("signal_l1_nb_decimate", C.c_int),
        ("thresh_null_start", C.c_float), ("thresh_null_end", C.c_float), ("fine_freq_update_beta", C.c_float),
        ("is_coarse_freq_correction", C.c_int), ("max_coarse_freq_correction_norm", C.c_float),
        ("coarse_freq_slow_beta", C.c_float), ("impulse_peak_threshold_db", C.c_float),
        ("impulse_peak_distance_probability", C.c_float),
    ]


class FrameInfo(C.Structure):
    _fields_ = [
        ("frame_start", C.c_int64), ("fine_time_offset", C.c_int32), ("total_desync", C.c_int32),
        ("coarse_offset", C.c_float), ("fine_offset_used", C.c_float), ("fine_offset_after", C.c_float),
        ("signal_average", C.c_float),
    ]


class OfdmState(C.Structure):
    _fields_ = [
        ("state", C.c_int32), ("fine_time_offset", C.c_int32), ("total_frames_read", C.c_int32),
        ("total_frames_desync", C.c_int32), ("signal_average", C.c_float), ("fine_offset", C.c_float),
        ("coarse_offset", C.c_float), ("pad", C.c_int32),
    ]


class Subchannel(C.Structure):
    """Subchannel (src/dab/database/dab_database_entities.h:179-190), the fields the MSC decoder reads."""
    _fields_ = [(k, C.c_int32) for k in ("start_address", "length", "is_uep", "uep_prot_index", "eep_prot_level", "eep_type_b")]


class C32(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


_lib = None


def build(force=False):
    if force or not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "dab_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "oracle"], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(ORACLE_SO)
    vp, sz, i32, f32, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_uint32
    L.orc_get_params.argtypes = [i32, C.POINTER(Params)]
    L.orc_get_prs.argtypes = [i32, vp]
    L.orc_get_mapper.argtypes = [i32, vp]
    L.orc_default_config.argtypes = [C.POINTER(Config)]
    L.orc_apply_pll.argtypes = [vp, vp, sz, f32, f32]
    L.orc_conj_mul_sum.argtypes = [vp, vp, sz]
    L.orc_conj_mul_sum.restype = C32
    L.orc_fft.argtypes = [vp, vp, i32, i32]
    L.orc_modulate.argtypes = [i32, vp, sz, vp, sz]
    L.orc_ofdm_create.argtypes = [i32]
    L.orc_ofdm_create.restype = vp
    L.orc_ofdm_create_custom.argtypes = [C.POINTER(Params), vp, vp]
    L.orc_ofdm_create_custom.restype = vp
    L.orc_ofdm_destroy.argtypes = [vp]
    L.orc_ofdm_config.argtypes = [vp]
    L.orc_ofdm_config.restype = C.POINTER(Config)
    L.orc_ofdm_process.argtypes = [vp, vp, sz]
    L.orc_ofdm_reset.argtypes = [vp]
    L.orc_ofdm_get_state.argtypes = [vp, C.POINTER(OfdmState)]
    L.orc_ofdm_frames_done.argtypes = [vp]
    L.orc_ofdm_frames_done.restype = sz
    L.orc_ofdm_get_frame.argtypes = [vp, sz, C.POINTER(FrameInfo), vp]
    L.orc_ofdm_frame_bits.argtypes = [vp]
    L.orc_ofdm_frame_bits.restype = sz
    for name, rt in (("orc_ofdm_frame_fft", vp), ("orc_ofdm_frame_data_vec", vp), ("orc_ofdm_impulse_response", vp),
                     ("orc_ofdm_coarse_freq_response", vp)):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = rt
    L.orc_ofdm_demod_frame.argtypes = [C.POINTER(Params), vp, vp, f32, vp, C.POINTER(C.c_float)]
    L.orc_vit_create.restype = vp
    L.orc_vit_destroy.argtypes = [vp]
    L.orc_vit_set_traceback_length.argtypes = [vp, sz]
    L.orc_vit_get_traceback_length.argtypes = [vp]
    L.orc_vit_get_traceback_length.restype = sz
    L.orc_vit_get_current_decoded_bit.argtypes = [vp]
    L.orc_vit_get_current_decoded_bit.restype = sz
    L.orc_vit_reset.argtypes = [vp, sz]
    L.orc_vit_update.argtypes = [vp, vp, sz, vp, sz, sz]
    L.orc_vit_update.restype = sz
    L.orc_vit_chainback.argtypes = [vp, vp, sz, sz]
    L.orc_vit_chainback.restype = C.c_uint64
    L.orc_vit_decode_job.argtypes = [vp, vp, sz, vp, vp, vp, u32, vp, sz, C.POINTER(sz)]
    L.orc_vit_decode_job.restype = C.c_uint64
    L.orc_puncture_code.argtypes = [i32]
    L.orc_puncture_code.restype = C.POINTER(C.c_uint8)
    L.orc_puncture_code_tail.restype = C.POINTER(C.c_uint8)
    L.orc_conv_encode.argtypes = [vp, sz, vp]
    L.orc_conv_encode.restype = sz
    L.orc_puncture.argtypes = [vp, sz, vp, vp, vp, u32, vp]
    L.orc_puncture.restype = sz
    L.orc_scrambler_bytes.argtypes = [C.c_uint16, vp, sz]
    L.orc_crc16_fib.argtypes = [vp, sz]
    L.orc_crc16_fib.restype = C.c_uint16
    L.orc_fic_decode_group.argtypes = [vp, vp, sz, sz, vp, vp]
    L.orc_fic_decode_group.restype = C.c_uint64
    L.orc_deint_create.argtypes = [sz]
    L.orc_deint_create.restype = vp
    L.orc_deint_destroy.argtypes = [vp]
    L.orc_deint_push.argtypes = [vp, vp, vp]
    L.orc_uep_subchannel_size.argtypes = [i32]
    L.orc_msc_segments.argtypes = [C.POINTER(Subchannel), vp, vp]
    L.orc_msc_create.argtypes = [C.POINTER(Subchannel)]
    L.orc_msc_create.restype = vp
    L.orc_msc_destroy.argtypes = [vp]
    L.orc_msc_decode_cif.argtypes = [vp, vp, sz, vp, C.POINTER(C.c_uint64)]
    L.orc_msc_decode_cif.restype = C.c_int64
    L.orc_ofdm_bench.argtypes = [i32, i32, vp, sz, sz, i32, C.POINTER(C.c_uint64)]
    L.orc_ofdm_bench.restype = C.c_double
    L.orc_vit_bench.argtypes = [i32, vp, sz, sz, vp, vp, vp, u32, sz, vp, sz]
    L.orc_vit_bench.restype = C.c_double
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def params_struct(mode):
    p = Params()
    if lib().orc_get_params(mode, C.byref(p)) != 0:
        raise ValueError(f"invalid transmission mode {mode}")
    return p


def params(mode):
    return params_struct(mode).asdict()


def prs(mode):
    out = np.zeros(params(mode)["nb_fft"], np.complex64)
    lib().orc_get_prs(mode, _p(out))
    return out


def mapper(mode):
    out = np.zeros(params(mode)["nb_data_carriers"], np.int32)
    lib().orc_get_mapper(mode, _p(out))
    return out


def apply_pll(x, freq_norm, dt_norm=0.0):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.empty_like(x)
    lib().orc_apply_pll(_p(x), _p(y), x.size, freq_norm, dt_norm)
    return y


def conj_mul_sum(x0, x1):
    x0 = np.ascontiguousarray(x0, np.complex64)
    x1 = np.ascontiguousarray(x1, np.complex64)
    r = lib().orc_conj_mul_sum(_p(x0), _p(x1), x0.size)
    return np.complex64(complex(r.re, r.im))


def fft(x, sign=-1):
    x = np.ascontiguousarray(x, np.complex64)
    y = np.empty_like(x)
    lib().orc_fft(_p(x), _p(y), x.size, sign)
    return y


def modulate(mode, data_bytes):
    p = params(mode)
    n = p["nb_null_period"] + p["nb_symbol_period"] * p["nb_frame_symbols"]
    data_bytes = np.ascontiguousarray(data_bytes, np.uint8)
    out = np.zeros(n, np.complex64)
    rc = lib().orc_modulate(mode, _p(data_bytes), data_bytes.size, _p(out), n)
    if rc != 0:
        raise RuntimeError(f"orc_modulate failed rc={rc}")
    return out


def demod_frame(mode, frame, freq_offset):
    """Stage-level: aligned frame (S*nb_symbol_period samples) -> (bits, phase_error_sum)."""
    p = params_struct(mode)
    m = mapper(mode)
    frame = np.ascontiguousarray(frame, np.complex64)
    bits = np.zeros((p.nb_frame_symbols - 1) * p.nb_data_carriers * 2, np.int8)
    pe = C.c_float()
    lib().orc_ofdm_demod_frame(C.byref(p), _p(m), _p(frame), freq_offset, _p(bits), C.byref(pe))
    return bits, float(pe.value)


class OracleOfdmDemod:
    def __init__(self, mode, custom=None):
        """mode: DAB transmission mode 1-4, or None with custom = (Params, prs_fft_ref complex64[nb_fft], mapper int32[nb_data_carriers])
        for any geometry OFDM_Demod's constructor accepts (ofdm_demodulator.cpp:80-146)"""
        self.L = lib()
        if custom is not None:
            self._custom = (custom[0], np.ascontiguousarray(custom[1], np.complex64), np.ascontiguousarray(custom[2], np.int32))
            self.h = self.L.orc_ofdm_create_custom(C.byref(self._custom[0]), _p(self._custom[1]), _p(self._custom[2]))
        else:
            self.h = self.L.orc_ofdm_create(mode)
        if not self.h:
            raise ValueError(f"invalid transmission mode {mode}")
        self.frame_bits = int(self.L.orc_ofdm_frame_bits(self.h))
        self.mode = mode

    @property
    def config(self):
        return self.L.orc_ofdm_config(self.h).contents

    def process(self, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        self.L.orc_ofdm_process(self.h, _p(iq), iq.size)

    def process_blocks(self, iq, block):
        iq = np.ascontiguousarray(iq, np.complex64)
        for off in range(0, iq.size, block):
            self.process(iq[off:off + block])

    def reset(self):
        self.L.orc_ofdm_reset(self.h)

    def frames_done(self):
        return int(self.L.orc_ofdm_frames_done(self.h))

    def frame(self, i):
        info = FrameInfo()
        bits = np.zeros(self.frame_bits, np.int8)
        if self.L.orc_ofdm_get_frame(self.h, i, C.byref(info), _p(bits)) != 0:
            raise IndexError(i)
        return {k: getattr(info, k) for k, _ in FrameInfo._fields_}, bits

    def state(self):
        s = OfdmState()
        self.L.orc_ofdm_get_state(self.h, C.byref(s))
        return {k: getattr(s, k) for k, _ in OfdmState._fields_ if k != "pad"}

    def _tap(self, fn, n, dtype):
        ptr = getattr(self.L, fn)(self.h)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n * (2 if dtype == np.complex64 else 1),)).copy().view(dtype)

    def frame_fft(self):
        p = params(self.mode)
        return self._tap("orc_ofdm_frame_fft", (p["nb_frame_symbols"] + 1) * p["nb_fft"], np.complex64)

    def frame_data_vec(self):
        p = params(self.mode)
        return self._tap("orc_ofdm_frame_data_vec", (p["nb_frame_symbols"] - 1) * p["nb_data_carriers"], np.complex64)

    def impulse_response(self):
        return self._tap("orc_ofdm_impulse_response", params(self.mode)["nb_fft"], np.float32)

    def coarse_freq_response(self):
        return self._tap("orc_ofdm_coarse_freq_response", params(self.mode)["nb_fft"], np.float32)

    def correlation_time_buffer(self):
        """(buffer of nb_null_period + nb_symbol_period samples, filled length)"""
        p = params(self.mode)
        n = p["nb_null_period"] + p["nb_symbol_period"]
        length = C.c_size_t(0)
        self.L.orc_ofdm_correlation_time_buffer.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        self.L.orc_ofdm_correlation_time_buffer.restype = C.c_void_p
        ptr = self.L.orc_ofdm_correlation_time_buffer(self.h, C.byref(length))
        buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(2 * n,)).copy().view(np.complex64)
        return buf, int(length.value)

    def close(self):
        if self.h:
            self.L.orc_ofdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def puncture_code(pi):
    ptr = lib().orc_puncture_code(pi)
    if not ptr:
        raise ValueError(pi)
    return np.array([ptr[i] for i in range(8)], np.uint8)


PI_X = np.array([2, 2, 2, 2, 2, 2], np.uint8)


def pack_segments(segments):
    """segments: list of (code ndarray, n_out) -> (codes[n_seg][8] u8, code_len u32, n_out u32)."""
    n = len(segments)
    codes = np.zeros((n, 8), np.uint8)
    lens = np.zeros(n, np.uint32)
    nout = np.zeros(n, np.uint32)
    for i, (code, n_out) in enumerate(segments):
        code = np.asarray(code, np.uint8)
        codes[i, :code.size] = code
        lens[i] = code.size
        nout[i] = n_out
    return codes, lens, nout


def conv_encode(data_bytes):
    data_bytes = np.ascontiguousarray(data_bytes, np.uint8)
    out = np.zeros((data_bytes.size * 8 + 6) * 4, np.int8)
    n = lib().orc_conv_encode(_p(data_bytes), data_bytes.size, _p(out))
    assert n == out.size
    return out


def puncture(mother, segments):
    mother = np.ascontiguousarray(mother, np.int8)
    codes, lens, nout = pack_segments(segments)
    out = np.zeros(mother.size, np.int8)
    n = lib().orc_puncture(_p(mother), mother.size, _p(codes), _p(lens), _p(nout), len(segments), _p(out))
    return out[:n].copy()


class OracleViterbi:
    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_vit_create()

    def set_traceback_length(self, n):
        self.L.orc_vit_set_traceback_length(self.h, n)

    def get_traceback_length(self):
        return int(self.L.orc_vit_get_traceback_length(self.h))

    def get_current_decoded_bit(self):
        return int(self.L.orc_vit_get_current_decoded_bit(self.h))

    def reset(self, start_state=0):
        self.L.orc_vit_reset(self.h, start_state)

    def update(self, soft, code, n_out):
        soft = np.ascontiguousarray(soft, np.int8)
        code = np.ascontiguousarray(code, np.uint8)
        return int(self.L.orc_vit_update(self.h, _p(soft), soft.size, _p(code), code.size, n_out))

    def chainback(self, nbytes, end_state=0):
        out = np.zeros(nbytes, np.uint8)
        err = int(self.L.orc_vit_chainback(self.h, _p(out), nbytes, end_state))
        return out, err

    def decode_job(self, soft, segments, n_out_bytes):
        soft = np.ascontiguousarray(soft, np.int8)
        codes, lens, nout = pack_segments(segments)
        out = np.zeros(n_out_bytes, np.uint8)
        used = C.c_size_t()
        err = int(self.L.orc_vit_decode_job(self.h, _p(soft), soft.size, _p(codes), _p(lens), _p(nout), len(segments), _p(out),
                                            n_out_bytes, C.byref(used)))
        return out, err, int(used.value)

    def close(self):
        if self.h:
            self.L.orc_vit_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------- FIC / MSC decode
def scrambler_bytes(n, syncword=0xFFFF):
    out = np.zeros(n, np.uint8)
    lib().orc_scrambler_bytes(syncword, _p(out), n)
    return out


def crc16_fib(data):
    data = np.ascontiguousarray(data, np.uint8)
    return int(lib().orc_crc16_fib(_p(data), data.size))


def subchannel(start_address, length, is_uep=False, uep_prot_index=0, eep_prot_level=0, eep_type_b=False):
    return Subchannel(int(start_address), int(length), int(bool(is_uep)), int(uep_prot_index), int(eep_prot_level), int(bool(eep_type_b)))


def uep_subchannel_size(index):
    return int(lib().orc_uep_subchannel_size(index))


def msc_segments(sc):
    """The update() schedule of MSC_Decoder for a sub-channel as [(puncture code, n_out)], tail included."""
    pi = np.zeros(5, np.int32)
    n_out = np.zeros(5, np.uint32)
    n = lib().orc_msc_segments(C.byref(sc), _p(pi), _p(n_out))
    return [(puncture_code(int(pi[i])) if pi[i] else PI_X, int(n_out[i])) for i in range(n)]


def fic_decode_group(bits, nb_fibs=3):
    """FIC_Decoder::DecodeFIBGroup -> (descrambled bytes, crc-valid flags, path error or None when the size is rejected)."""
    bits = np.ascontiguousarray(bits, np.int8)
    v = OracleViterbi()
    out = np.zeros(bits.size // 24, np.uint8)
    valid = np.zeros(nb_fibs, np.uint8)
    err = int(lib().orc_fic_decode_group(v.h, _p(bits), bits.size, nb_fibs, _p(out), _p(valid)))
    v.close()
    return out, valid, (None if err == 0xFFFFFFFFFFFFFFFF else err)


class OracleDeinterleaver:
    def __init__(self, nb_bits):
        self.L = lib()
        self.n = nb_bits
        self.h = self.L.orc_deint_create(nb_bits)

    def push(self, bits):
        bits = np.ascontiguousarray(bits, np.int8)
        assert bits.size == self.n
        out = np.zeros(self.n, np.int8)
        ok = self.L.orc_deint_push(self.h, _p(bits), _p(out))
        return out if ok else None

    def __del__(self):
        try:
            if self.h:
                self.L.orc_deint_destroy(self.h)
                self.h = None
        except Exception:
            pass


class OracleMscDecoder:
    """MSC_Decoder for one sub-channel (msc_decoder.cpp:27-170)."""

    def __init__(self, sc):
        self.L = lib()
        self.sc = sc
        self.h = self.L.orc_msc_create(C.byref(sc))

    def decode_cif(self, cif_bits):
        """-> (bytes ndarray (empty while the de-interleaver fills), path error or None)."""
        cif_bits = np.ascontiguousarray(cif_bits, np.int8)
        out = np.zeros(self.sc.length * 8 + 8, np.uint8)
        err = C.c_uint64(0)
        n = int(self.L.orc_msc_decode_cif(self.h, _p(cif_bits), cif_bits.size, _p(out), C.byref(err)))
        if n <= 0:
            return np.zeros(0, np.uint8), None
        return out[:n].copy(), int(err.value)

    def __del__(self):
        try:
            if self.h:
                self.L.orc_msc_destroy(self.h)
                self.h = None
        except Exception:
            pass
