"""Thin Python plumbing over the dab_ensemble_* C ABI (tests and bench use it; the product interface is the C ABI and the
C++ mirror classes FIC_Decoder / MSC_Decoder in cpp/dab/)."""
import ctypes as C

import numpy as np

from . import capi


def dab_parameters(mode):
    """get_dab_parameters (reference src/dab/constants/dab_parameters.h:26-93)."""
    p = capi.DabParameters()
    capi.check(capi.load().dab_get_dab_parameters(mode, C.byref(p)))
    return p


def subchannel(start_address, length, is_uep=False, uep_prot_index=0, eep_prot_level=0, eep_type_b=False, id=0):
    return capi.Subchannel(int(id), int(start_address), int(length), int(bool(is_uep)), int(uep_prot_index), int(eep_prot_level),
                           int(bool(eep_type_b)), 0)


def subchannel_schedule(sub):
    """-> (capi.VitSchedule, n_soft): MSC_Decoder's update() sequence for the sub-channel."""
    sch = capi.VitSchedule()
    n_soft = C.c_uint32(0)
    capi.check(capi.load().dab_ensemble_subchannel_schedule(C.byref(sub), C.byref(sch), C.byref(n_soft)))
    return sch, int(n_soft.value)


class EnsembleDecoder:
    """One dab_ensemble handle: n_streams ensembles of one transmission mode; decode one OFDM frame per stream per call."""

    def __init__(self, params, n_streams=1, device=0, max_subchannels=0):
        self.L = capi.load()
        self.params = params if isinstance(params, capi.DabParameters) else dab_parameters(params)
        self.n_streams = n_streams
        opt = capi.EnsembleOptions(n_streams, device, max_subchannels)
        status = C.c_int(0)
        self.h = self.L.dab_ensemble_create(C.byref(self.params), C.byref(opt), C.byref(status))
        if not self.h:
            capi.check(status.value)
            raise capi.DabError(status.value, "dab_ensemble_create failed")
        self.subs = {}

    def set_cuda_stream(self, stream_ptr):
        capi.check(self.L.dab_ensemble_set_cuda_stream(self.h, stream_ptr))

    def set_decode_stream(self, stream_ptr):
        """second CUDA stream for everything after the ingest of a frame (dab_ensemble_set_decode_stream); None / 0 switches it off"""
        capi.check(self.L.dab_ensemble_set_decode_stream(self.h, stream_ptr or None))

    def set_subchannels(self, stream, subs):
        subs = list(subs)
        arr = (capi.Subchannel * max(len(subs), 1))(*subs)
        capi.check(self.L.dab_ensemble_set_subchannels(self.h, stream, C.cast(arr, C.c_void_p), len(subs)))
        for s in (range(self.n_streams) if stream < 0 else [stream]):
            self.subs[s] = subs

    def decode_frames(self, bits, present=None):
        """bits: int8 [n_streams][nb_frame_bits] host array."""
        bits = np.ascontiguousarray(bits, np.int8)
        frame = self.params.nb_fic_bits + self.params.nb_cifs * self.params.nb_cif_bits
        assert bits.size == self.n_streams * frame, (bits.size, self.n_streams, frame)
        if present is not None:
            present = np.ascontiguousarray(present, np.uint8)
        capi.check(self.L.dab_ensemble_decode_frames(self.h, capi.ptr(bits), capi.ptr(present)))

    def decode_frames_device(self, d_bits, stream_stride, d_frames_in_call=None, slot=0):
        capi.check(self.L.dab_ensemble_decode_frames_device(self.h, d_bits, stream_stride, d_frames_in_call, slot))

    def device_results(self):
        r = capi.EnsembleResults()
        capi.check(self.L.dab_ensemble_device_results(self.h, C.byref(r)))
        return r

    def read_fic(self, stream):
        """-> (bytes [nb_cifs][group bytes], valid [nb_cifs][fibs], path error [nb_cifs])."""
        p = self.params
        gb = p.nb_fib_cif_bits // 24
        out = np.zeros((p.nb_cifs, max(gb, 1)), np.uint8)
        valid = np.zeros((p.nb_cifs, max(p.nb_fibs_per_cif, 1)), np.uint8)
        err = np.zeros(p.nb_cifs, np.uint64)
        capi.check(self.L.dab_ensemble_read_fic(self.h, stream, capi.ptr(out), capi.ptr(valid), capi.ptr(err)))
        return out[:, :gb], valid, err

    def read_msc(self, stream, cif, sub_index):
        """-> (bytes ndarray, n_bytes (0 = de-interleaver filling, -1 = overflow), path error)."""
        out = np.zeros(self.params.nb_cif_bits // 8 + 8, np.uint8)
        n = C.c_int32(0)
        err = C.c_uint64(0)
        capi.check(self.L.dab_ensemble_read_msc(self.h, stream, cif, sub_index, capi.ptr(out), out.size, C.byref(n), C.byref(err)))
        return out[:max(n.value, 0)].copy(), int(n.value), int(err.value)

    def sync(self):
        capi.check(self.L.dab_ensemble_sync(self.h))

    def kernel_launches(self):
        return int(self.L.dab_ensemble_kernel_launches(self.h))

    def last_work(self):
        t, s = C.c_uint64(0), C.c_uint64(0)
        capi.check(self.L.dab_ensemble_last_work(self.h, C.byref(t), C.byref(s)))
        return int(t.value), int(s.value)

    def close(self):
        if self.h:
            self.L.dab_ensemble_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
