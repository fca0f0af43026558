"""ctypes bindings of include/dab_b200.h (libdab_b200.so).  No torch types, no CPU fallback: if the library or a
sm_100 device is missing, the constructors raise."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdab_b200.so")

DAB_OK = 0
DAB_ERR_INVALID = -1
DAB_ERR_NO_DEVICE = -2
DAB_ERR_CUDA = -3
DAB_ERR_UNDERRUN = -4
DAB_ERR_CAPACITY = -5
DAB_ERR_TRACEBACK = -6

DAB_VIT_MAX_SEGMENTS = 8


class DabError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"dab_b200 status {status}: {message}")
        self.status = status


class C32(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class OfdmParams(C.Structure):
    _fields_ = [(k, C.c_size_t) for k in
                ("nb_frame_symbols", "nb_symbol_period", "nb_null_period", "nb_cyclic_prefix", "nb_fft", "nb_data_carriers")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class OfdmConfig(C.Structure):
    _fields_ = [
        ("signal_l1_update_beta", C.c_float), ("signal_l1_nb_samples", C.c_int), ("signal_l1_nb_decimate", C.c_int),
        ("null_l1_thresh_null_start", C.c_float), ("null_l1_thresh_null_end", C.c_float),
        ("sync_fine_freq_update_beta", C.c_float), ("sync_is_coarse_freq_correction", C.c_int),
        ("sync_max_coarse_freq_correction_norm", C.c_float), ("sync_coarse_freq_slow_beta", C.c_float),
        ("sync_impulse_peak_threshold_db", C.c_float), ("sync_impulse_peak_distance_probability", C.c_float),
    ]


class FrameCounter(C.Structure):
    _fields_ = [("frames", C.c_uint64), ("bits", C.c_uint64), ("checksum", C.c_uint64)]


class OfdmState(C.Structure):
    _fields_ = [
        ("state", C.c_int32), ("fine_time_offset", C.c_int32), ("total_frames_read", C.c_int32),
        ("total_frames_desync", C.c_int32), ("signal_average", C.c_float), ("fine_frequency_offset", C.c_float),
        ("coarse_frequency_offset", C.c_float), ("reserved", C.c_int32),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class OfdmFrameInfo(C.Structure):
    _fields_ = [
        ("frame_start", C.c_int64), ("fine_time_offset", C.c_int32), ("total_desync", C.c_int32),
        ("coarse_offset", C.c_float), ("fine_offset_used", C.c_float), ("fine_offset_after", C.c_float),
        ("signal_average", C.c_float),
    ]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# dab_iq_format
IQ_F32, IQ_U8, IQ_S8, IQ_S16LE, IQ_S16BE, IQ_U16LE, IQ_U16BE = range(7)
IQ_FORMATS = {"f32": IQ_F32, "u8": IQ_U8, "s8": IQ_S8, "s16le": IQ_S16LE, "s16be": IQ_S16BE, "u16le": IQ_U16LE, "u16be": IQ_U16BE}


class OfdmOptions(C.Structure):
    _fields_ = [("n_streams", C.c_int), ("device", C.c_int), ("max_block_samples", C.c_size_t), ("keep_debug_taps", C.c_int),
                ("sample_format", C.c_int)]


class OfdmKernelTimes(C.Structure):
    _fields_ = [("frame_ms", C.c_double * 8), ("frame_launches", C.c_uint64 * 8), ("control_ms", C.c_double * 8),
                ("control_launches", C.c_uint64 * 8)]


FRAME_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_int8), C.c_size_t, C.POINTER(OfdmFrameInfo))


class VitSegment(C.Structure):
    _fields_ = [("counts", C.c_uint8 * 8), ("code_len", C.c_uint32), ("n_out", C.c_uint32)]


class VitSchedule(C.Structure):
    _fields_ = [("seg", VitSegment * DAB_VIT_MAX_SEGMENTS), ("n_seg", C.c_uint32), ("n_out_bytes", C.c_uint32),
                ("start_state", C.c_uint32), ("end_state", C.c_uint32)]


class DabParameters(C.Structure):
    """DAB_Parameters (reference src/dab/constants/dab_parameters.h:5-21)."""
    _fields_ = [(k, C.c_int) for k in
                ("nb_frame_bits", "nb_symbols", "nb_fic_symbols", "nb_msc_symbols", "nb_fibs", "nb_cifs", "nb_fibs_per_cif",
                 "nb_sym_bits", "nb_fic_bits", "nb_msc_bits", "nb_fib_bits", "nb_fib_cif_bits", "nb_cif_bits")]

    def asdict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Subchannel(C.Structure):
    """The fields of Subchannel (reference src/dab/database/dab_database_entities.h:179-190) the decoder reads."""
    _fields_ = [(k, C.c_int32) for k in
                ("id", "start_address", "length", "is_uep", "uep_prot_index", "eep_prot_level", "eep_type_b", "reserved")]


class EnsembleOptions(C.Structure):
    _fields_ = [("n_streams", C.c_int), ("device", C.c_int), ("max_subchannels", C.c_int)]


class EnsembleResults(C.Structure):
    _fields_ = [("fib_bytes", C.c_void_p), ("fib_valid", C.c_void_p), ("fic_error", C.c_void_p), ("msc_bytes", C.c_void_p),
                ("msc_nbytes", C.c_void_p), ("msc_error", C.c_void_p), ("decoded", C.c_void_p), ("fib_group_bytes", C.c_size_t),
                ("msc_cif_bytes", C.c_size_t), ("nb_cifs", C.c_int), ("nb_fibs_per_cif", C.c_int), ("max_subchannels", C.c_int)]


VIT_JOB_DTYPE = np.dtype([("schedule", np.uint32), ("n_soft", np.uint32), ("soft_offset", np.uint64), ("out_offset", np.uint64)])

# every symbol include/dab_b200.h declares (tests/test_capi_symbols.py checks the library exports each of them)
EXPORTED_SYMBOLS = (
    "dab_last_error", "dab_version", "dab_device_count", "dab_get_ofdm_params", "dab_get_prs_reference",
    "dab_get_mapper_reference", "dab_get_puncture_code",
    "dab_ofdm_create", "dab_ofdm_destroy", "dab_ofdm_set_cuda_stream", "dab_ofdm_set_frame_callback", "dab_ofdm_set_config",
    "dab_ofdm_get_config", "dab_ofdm_default_config", "dab_ofdm_process", "dab_ofdm_process_batch", "dab_ofdm_process_batch_u8",
    "dab_ofdm_process_batch_raw", "dab_iq_format_bytes", "dab_ofdm_rebase_device_streams", "dab_ofdm_attach_device_streams_raw",
    "dab_ofdm_attach_device_streams", "dab_ofdm_advance", "dab_ofdm_advance_uniform", "dab_ofdm_device_bits", "dab_ofdm_reset",
    "dab_ofdm_get_state", "dab_ofdm_sync", "dab_ofdm_join", "dab_ofdm_count_frames_cb", "dab_ofdm_frame_bits", "dab_ofdm_get_params", "dab_ofdm_get_impulse_response",
    "dab_ofdm_get_coarse_frequency_response", "dab_ofdm_get_correlation_time_buffer", "dab_ofdm_get_frame_data_bits", "dab_ofdm_get_frame_fft", "dab_ofdm_get_frame_data_vec",
    "dab_ofdm_kernel_launches", "dab_ofdm_set_kernel_timing", "dab_ofdm_get_kernel_times", "dab_ofdm_demod_frames_device",
    "dab_viterbi_create", "dab_viterbi_destroy", "dab_viterbi_set_cuda_stream", "dab_viterbi_add_schedule",
    "dab_viterbi_schedule_soft_symbols", "dab_viterbi_decode_batch", "dab_viterbi_decode_batch_device",
    "dab_viterbi_decode_jobs_device", "dab_viterbi_prepare_jobs", "dab_viterbi_decode_prepared", "dab_viterbi_release_jobs", "dab_viterbi_decode_one", "dab_viterbi_sync", "dab_viterbi_kernel_launches",
    "dab_get_dab_parameters", "dab_ensemble_create", "dab_ensemble_destroy", "dab_ensemble_set_cuda_stream", "dab_ensemble_set_decode_stream",
    "dab_ensemble_set_subchannels", "dab_ensemble_subchannel_schedule", "dab_ensemble_decode_frames_device",
    "dab_ensemble_decode_frames", "dab_ensemble_device_results", "dab_ensemble_read_fic", "dab_ensemble_read_msc",
    "dab_ensemble_sync", "dab_ensemble_kernel_launches", "dab_ensemble_last_work", "dab_ensemble_schedule_count",
)

_lib = None


def load():
    """Load libdab_b200.so (built in-tree by build.py).  Raises if it is missing: there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python dab-radio_b200/build.py` (nvcc, sm_100a)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64
    ip = C.POINTER(C.c_int)
    L.dab_last_error.restype = C.c_char_p
    L.dab_version.restype = C.c_char_p
    L.dab_device_count.restype = i32
    L.dab_get_ofdm_params.argtypes = [i32, C.POINTER(OfdmParams)]
    L.dab_get_prs_reference.argtypes = [i32, vp, sz]
    L.dab_get_mapper_reference.argtypes = [vp, sz, sz]
    L.dab_get_puncture_code.argtypes = [i32, C.POINTER(C.c_uint8 * 8)]
    # OFDM
    L.dab_ofdm_create.argtypes = [C.POINTER(OfdmParams), vp, vp, C.POINTER(OfdmOptions), ip]
    L.dab_ofdm_create.restype = vp
    L.dab_ofdm_destroy.argtypes = [vp]
    L.dab_ofdm_destroy.restype = None
    L.dab_ofdm_set_cuda_stream.argtypes = [vp, vp]
    L.dab_ofdm_set_frame_callback.argtypes = [vp, FRAME_CB, vp]
    L.dab_ofdm_set_config.argtypes = [vp, i32, C.POINTER(OfdmConfig)]
    L.dab_ofdm_get_config.argtypes = [vp, i32, C.POINTER(OfdmConfig)]
    L.dab_ofdm_default_config.argtypes = [C.POINTER(OfdmConfig)]
    L.dab_ofdm_default_config.restype = None
    L.dab_ofdm_process.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_process_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.dab_ofdm_process_batch_u8.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.dab_ofdm_process_batch_raw.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
    L.dab_iq_format_bytes.argtypes = [i32]
    L.dab_iq_format_bytes.restype = sz
    L.dab_ofdm_attach_device_streams.argtypes = [vp, vp, sz, sz]
    L.dab_ofdm_attach_device_streams_raw.argtypes = [vp, vp, sz, sz]
    L.dab_ofdm_advance.argtypes = [vp, C.POINTER(sz)]
    L.dab_ofdm_advance_uniform.argtypes = [vp, sz]
    L.dab_ofdm_rebase_device_streams.argtypes = [vp, sz]
    L.dab_ofdm_device_bits.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), ip, C.POINTER(vp)]
    L.dab_ofdm_reset.argtypes = [vp, i32]
    L.dab_ofdm_join.argtypes = [vp]
    L.dab_ofdm_get_state.argtypes = [vp, i32, C.POINTER(OfdmState)]
    L.dab_ofdm_sync.argtypes = [vp]
    L.dab_ofdm_frame_bits.argtypes = [vp]
    L.dab_ofdm_frame_bits.restype = sz
    L.dab_ofdm_get_params.argtypes = [vp, C.POINTER(OfdmParams)]
    L.dab_ofdm_get_impulse_response.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_get_coarse_frequency_response.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_get_correlation_time_buffer.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_get_frame_data_bits.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_get_frame_fft.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_get_frame_data_vec.argtypes = [vp, i32, vp, sz]
    L.dab_ofdm_kernel_launches.argtypes = [vp]
    L.dab_ofdm_kernel_launches.restype = u64
    L.dab_ofdm_set_kernel_timing.argtypes = [vp, i32]
    L.dab_ofdm_get_kernel_times.argtypes = [vp, C.POINTER(OfdmKernelTimes)]
    L.dab_ofdm_demod_frames_device.argtypes = [vp, vp, sz, i32, vp, vp, vp]
    _bind_viterbi(L)
    _bind_ensemble(L)
    _lib = L
    return L


def _bind_viterbi(L):
    vp, sz, i32, u32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint32, C.c_uint64
    ip = C.POINTER(C.c_int)
    L.dab_viterbi_create.argtypes = [i32, ip]
    L.dab_viterbi_create.restype = vp
    L.dab_viterbi_destroy.argtypes = [vp]
    L.dab_viterbi_destroy.restype = None
    L.dab_viterbi_set_cuda_stream.argtypes = [vp, vp]
    L.dab_viterbi_add_schedule.argtypes = [vp, C.POINTER(VitSchedule)]
    L.dab_viterbi_schedule_soft_symbols.argtypes = [C.POINTER(VitSchedule)]
    L.dab_viterbi_schedule_soft_symbols.restype = C.c_int64
    L.dab_viterbi_decode_batch.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, vp]
    L.dab_viterbi_decode_batch_device.argtypes = [vp, vp, sz, vp, i32, vp, sz, vp, vp]
    L.dab_viterbi_decode_jobs_device.argtypes = [vp, vp, sz, vp, i32, u32, vp, sz, vp, vp]
    L.dab_viterbi_prepare_jobs.argtypes = [vp, vp, i32]
    L.dab_viterbi_decode_prepared.argtypes = [vp, i32, vp, sz, vp, sz, vp, vp]
    L.dab_viterbi_release_jobs.argtypes = [vp, i32]
    L.dab_viterbi_decode_one.argtypes = [vp, C.POINTER(VitSchedule), vp, sz, vp, C.POINTER(u64)]
    L.dab_viterbi_sync.argtypes = [vp]
    L.dab_viterbi_kernel_launches.argtypes = [vp]
    L.dab_viterbi_kernel_launches.restype = u64
    return L


def _bind_ensemble(L):
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    ip = C.POINTER(C.c_int)
    L.dab_get_dab_parameters.argtypes = [i32, C.POINTER(DabParameters)]
    L.dab_ensemble_create.argtypes = [C.POINTER(DabParameters), C.POINTER(EnsembleOptions), ip]
    L.dab_ensemble_create.restype = vp
    L.dab_ensemble_destroy.argtypes = [vp]
    L.dab_ensemble_destroy.restype = None
    L.dab_ensemble_set_cuda_stream.argtypes = [vp, vp]
    L.dab_ensemble_set_decode_stream.argtypes = [vp, vp]
    L.dab_ensemble_set_subchannels.argtypes = [vp, i32, vp, i32]
    L.dab_ensemble_subchannel_schedule.argtypes = [C.POINTER(Subchannel), C.POINTER(VitSchedule), C.POINTER(C.c_uint32)]
    L.dab_ensemble_decode_frames_device.argtypes = [vp, vp, sz, vp, i32]
    L.dab_ensemble_decode_frames.argtypes = [vp, vp, vp]
    L.dab_ensemble_device_results.argtypes = [vp, C.POINTER(EnsembleResults)]
    L.dab_ensemble_read_fic.argtypes = [vp, i32, vp, vp, vp]
    L.dab_ensemble_read_msc.argtypes = [vp, i32, i32, i32, vp, sz, C.POINTER(C.c_int32), C.POINTER(u64)]
    L.dab_ensemble_sync.argtypes = [vp]
    L.dab_ensemble_schedule_count.argtypes = [vp]
    L.dab_ensemble_kernel_launches.argtypes = [vp]
    L.dab_ensemble_kernel_launches.restype = u64
    L.dab_ensemble_last_work.argtypes = [vp, C.POINTER(u64), C.POINTER(u64)]
    return L


def check(status):
    if status < 0:
        raise DabError(status, load().dab_last_error().decode("utf-8", "replace"))
    return status


def ptr(a):
    """void* of a numpy array (or pass through ints / None)."""
    if a is None:
        return None
    if isinstance(a, (int, C.c_void_p)):
        return a
    return a.ctypes.data_as(C.c_void_p)
