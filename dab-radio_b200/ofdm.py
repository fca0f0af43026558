"""Thin Python plumbing over the dab_ofdm_* C ABI (tests and bench use it; the product interface is the C ABI and the C++
mirror class in cpp/ofdm_demodulator.h)."""
import ctypes as C

import numpy as np

from . import capi


def ofdm_params(mode):
    p = capi.OfdmParams()
    capi.check(capi.load().dab_get_ofdm_params(mode, C.byref(p)))
    return p


def prs_reference(mode):
    p = ofdm_params(mode)
    out = np.zeros(p.nb_fft, np.complex64)
    capi.check(capi.load().dab_get_prs_reference(mode, capi.ptr(out), p.nb_fft))
    return out


def mapper_reference(mode):
    p = ofdm_params(mode)
    out = np.zeros(p.nb_data_carriers, np.int32)
    capi.check(capi.load().dab_get_mapper_reference(capi.ptr(out), p.nb_data_carriers, p.nb_fft))
    return out


class OfdmDemodBatch:
    """n_streams independent OFDM demodulators of one DAB transmission mode on one GPU."""

    def __init__(self, mode, n_streams=1, device=0, max_block_samples=0, keep_debug_taps=False, raw_u8=False, params=None, prs=None,
                 mapper=None, sample_format=None):
        self.L = capi.load()
        self.params = params if params is not None else ofdm_params(mode)
        prs = prs if prs is not None else prs_reference(mode)
        mapper = mapper if mapper is not None else mapper_reference(mode)
        fmt = capi.IQ_U8 if raw_u8 else capi.IQ_F32
        if sample_format is not None:
            fmt = capi.IQ_FORMATS[sample_format] if isinstance(sample_format, str) else int(sample_format)
        self.sample_format = fmt
        self.sample_bytes = int(self.L.dab_iq_format_bytes(fmt))
        opts = capi.OfdmOptions(n_streams, device, max_block_samples, 1 if keep_debug_taps else 0, fmt)
        status = C.c_int(0)
        prs = np.ascontiguousarray(prs, np.complex64)
        mapper = np.ascontiguousarray(mapper, np.int32)
        self.h = self.L.dab_ofdm_create(C.byref(self.params), capi.ptr(prs), capi.ptr(mapper), C.byref(opts), C.byref(status))
        if not self.h:
            capi.check(status.value)
            raise capi.DabError(status.value, "dab_ofdm_create failed")
        self.n_streams = n_streams
        self.frame_bits = int(self.L.dab_ofdm_frame_bits(self.h))
        self.frames = [[] for _ in range(n_streams)]  # (info dict, bits copy) per stream, filled by the callback
        self._cb = capi.FRAME_CB(self._on_frame)
        self.collect = True
        capi.check(self.L.dab_ofdm_set_frame_callback(self.h, self._cb, None))

    def _on_frame(self, user, stream, bits, n_bits, info):
        if self.collect:
            # the span is only valid during the callback (reference ofdm_demodulator.cpp:110,635): copy
            arr = np.ctypeslib.as_array(bits, shape=(n_bits,)).copy()
            self.frames[stream].append((info.contents.asdict(), arr))

    def disable_callback(self):
        capi.check(self.L.dab_ofdm_set_frame_callback(self.h, capi.FRAME_CB(), None))

    def set_cuda_stream(self, stream_ptr):
        capi.check(self.L.dab_ofdm_set_cuda_stream(self.h, stream_ptr))

    def process(self, stream, iq):
        iq = np.ascontiguousarray(iq, np.complex64)
        capi.check(self.L.dab_ofdm_process(self.h, stream, capi.ptr(iq), iq.size))

    def process_batch(self, blocks):
        """blocks: list of n_streams complex64 arrays (or None)."""
        arrs = [None if b is None else np.ascontiguousarray(b, np.complex64) for b in blocks]
        ptrs = (C.c_void_p * self.n_streams)(*[None if a is None else a.ctypes.data for a in arrs])
        ns = (C.c_size_t * self.n_streams)(*[0 if a is None else a.size for a in arrs])
        capi.check(self.L.dab_ofdm_process_batch(self.h, ptrs, ns))

    def process_batch_u8(self, blocks):
        arrs = [None if b is None else np.ascontiguousarray(b, np.uint8) for b in blocks]
        ptrs = (C.c_void_p * self.n_streams)(*[None if a is None else a.ctypes.data for a in arrs])
        ns = (C.c_size_t * self.n_streams)(*[0 if a is None else a.size // 2 for a in arrs])
        capi.check(self.L.dab_ofdm_process_batch_u8(self.h, ptrs, ns))

    def process_batch_raw(self, blocks):
        """blocks: per stream a byte array (np.uint8) of raw samples in the handle's sample_format, or None"""
        arrs = [None if b is None else np.ascontiguousarray(b).view(np.uint8) for b in blocks]
        ptrs = (C.c_void_p * self.n_streams)(*[None if a is None else a.ctypes.data for a in arrs])
        ns = (C.c_size_t * self.n_streams)(*[0 if a is None else a.size // self.sample_bytes for a in arrs])
        capi.check(self.L.dab_ofdm_process_batch_raw(self.h, ptrs, ns))

    def process_batch_ptrs(self, ptrs, ns):
        """raw host pointers (ints) and sample counts: used by bench.py with pinned torch tensors"""
        p = (C.c_void_p * self.n_streams)(*ptrs)
        n = (C.c_size_t * self.n_streams)(*ns)
        capi.check(self.L.dab_ofdm_process_batch(self.h, p, n))

    def pointer_arrays(self, ptrs, ns):
        """ctypes argument arrays for process_batch_prepared, built once and reused every step"""
        return (C.c_void_p * self.n_streams)(*ptrs), (C.c_size_t * self.n_streams)(*ns)

    def process_batch_prepared(self, p, n, u8=False):
        capi.check((self.L.dab_ofdm_process_batch_u8 if u8 else self.L.dab_ofdm_process_batch)(self.h, p, n))

    def use_counting_callback(self):
        """deliver frames to the library's own dab_ofdm_count_frames_cb (no Python in the delivery loop); returns the counter"""
        self.counter = capi.FrameCounter()
        cb = C.cast(self.L.dab_ofdm_count_frames_cb, capi.FRAME_CB)
        capi.check(self.L.dab_ofdm_set_frame_callback(self.h, cb, C.cast(C.pointer(self.counter), C.c_void_p)))
        return self.counter

    def attach_device_streams(self, d_ptr, stride_samples, total_samples):
        fn = self.L.dab_ofdm_attach_device_streams if self.sample_format == capi.IQ_F32 else self.L.dab_ofdm_attach_device_streams_raw
        capi.check(fn(self.h, d_ptr, stride_samples, total_samples))

    def advance_uniform(self, n):
        capi.check(self.L.dab_ofdm_advance_uniform(self.h, n))

    def rebase_device_streams(self, delta_samples):
        capi.check(self.L.dab_ofdm_rebase_device_streams(self.h, delta_samples))

    def advance(self, ns):
        n = (C.c_size_t * self.n_streams)(*ns)
        capi.check(self.L.dab_ofdm_advance(self.h, n))

    def device_bits(self):
        d_bits, n_bits, slots, d_frames = C.c_void_p(), C.c_size_t(), C.c_int(), C.c_void_p()
        capi.check(self.L.dab_ofdm_device_bits(self.h, C.byref(d_bits), C.byref(n_bits), C.byref(slots), C.byref(d_frames)))
        return d_bits.value, int(n_bits.value), int(slots.value), d_frames.value

    def demod_frames_device(self, d_frames, frame_stride, n_frames, freq_offsets, d_bits, d_phase_err):
        f = np.ascontiguousarray(freq_offsets, np.float32)
        capi.check(self.L.dab_ofdm_demod_frames_device(self.h, d_frames, frame_stride, n_frames, capi.ptr(f), d_bits, d_phase_err))

    def reset(self, stream):
        capi.check(self.L.dab_ofdm_reset(self.h, stream))

    def state(self, stream):
        s = capi.OfdmState()
        capi.check(self.L.dab_ofdm_get_state(self.h, stream, C.byref(s)))
        return s.asdict()

    def get_config(self, stream=0):
        c = capi.OfdmConfig()
        capi.check(self.L.dab_ofdm_get_config(self.h, stream, C.byref(c)))
        return c

    def set_config(self, cfg, stream=-1):
        capi.check(self.L.dab_ofdm_set_config(self.h, stream, C.byref(cfg)))

    def sync(self):
        capi.check(self.L.dab_ofdm_sync(self.h))

    def join(self):
        """order the handle's CUDA stream after all queued work (device side, the host does not block)"""
        capi.check(self.L.dab_ofdm_join(self.h))

    def impulse_response(self, stream):
        out = np.zeros(self.params.nb_fft, np.float32)
        capi.check(self.L.dab_ofdm_get_impulse_response(self.h, stream, capi.ptr(out), out.size))
        return out

    def coarse_frequency_response(self, stream):
        out = np.zeros(self.params.nb_fft, np.float32)
        capi.check(self.L.dab_ofdm_get_coarse_frequency_response(self.h, stream, capi.ptr(out), out.size))
        return out

    def frame_data_bits(self, stream):
        out = np.zeros(self.frame_bits, np.int8)
        capi.check(self.L.dab_ofdm_get_frame_data_bits(self.h, stream, capi.ptr(out), out.size))
        return out

    def correlation_time_buffer(self, stream):
        out = np.zeros(self.params.nb_null_period + self.params.nb_symbol_period, np.complex64)
        capi.check(self.L.dab_ofdm_get_correlation_time_buffer(self.h, stream, capi.ptr(out), out.size))
        return out

    def frame_fft(self, stream):
        out = np.zeros(self.params.nb_frame_symbols * self.params.nb_fft, np.complex64)
        capi.check(self.L.dab_ofdm_get_frame_fft(self.h, stream, capi.ptr(out), out.size))
        return out

    def frame_data_vec(self, stream):
        out = np.zeros((self.params.nb_frame_symbols - 1) * self.params.nb_data_carriers, np.complex64)
        capi.check(self.L.dab_ofdm_get_frame_data_vec(self.h, stream, capi.ptr(out), out.size))
        return out

    def set_kernel_timing(self, enable=True):
        capi.check(self.L.dab_ofdm_set_kernel_timing(self.h, 1 if enable else 0))

    def kernel_times(self):
        t = capi.OfdmKernelTimes()
        capi.check(self.L.dab_ofdm_get_kernel_times(self.h, C.byref(t)))
        return {k: list(getattr(t, k)) for k in ("frame_ms", "frame_launches", "control_ms", "control_launches")}

    def kernel_launches(self):
        return int(self.L.dab_ofdm_kernel_launches(self.h))

    def close(self):
        if self.h:
            self.L.dab_ofdm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
