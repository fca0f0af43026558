"""Thin Python plumbing over the dab_viterbi_* C ABI (tests and bench use it; the product interface is the C ABI and the
C++ mirror class in cpp/dab_viterbi_decoder.h)."""
import ctypes as C

import numpy as np

from . import capi


def puncture_code(pi):
    """PI_1..PI_24 count table; pi = 0 gives the tail code PI_X (length 6)."""
    buf = (C.c_uint8 * 8)()
    n = capi.check(capi.load().dab_get_puncture_code(pi, C.byref(buf)))
    return np.array(buf[:n], np.uint8)


def make_schedule(segments, n_out_bytes, start_state=0, end_state=0):
    """segments: iterable of (code counts, requested_output_symbols) as passed to DAB_Viterbi_Decoder::update."""
    s = capi.VitSchedule()
    segments = list(segments)
    if len(segments) > capi.DAB_VIT_MAX_SEGMENTS:
        raise ValueError("too many segments")
    for i, (code, n_out) in enumerate(segments):
        code = np.asarray(code, np.uint8)
        for j, c in enumerate(code):
            s.seg[i].counts[j] = int(c)
        s.seg[i].code_len = code.size
        s.seg[i].n_out = int(n_out)
    s.n_seg = len(segments)
    s.n_out_bytes = int(n_out_bytes)
    s.start_state = start_state
    s.end_state = end_state
    return s


def fic_schedule():
    """FIC_Decoder::DecodeFIBGroup (reference src/dab/fic/fic_decoder.cpp:74-87): PI_16 x 21, PI_15 x 3, PI_X; 96 bytes."""
    return make_schedule([(puncture_code(16), 128 * 21), (puncture_code(15), 128 * 3), (puncture_code(0), 24)], 96)


def schedule_soft_symbols(schedule):
    return int(capi.load().dab_viterbi_schedule_soft_symbols(C.byref(schedule)))


class ViterbiBatch:
    """One dab_viterbi handle: register schedules, decode batches of independent trellises."""

    def __init__(self, device=0):
        self.L = capi.load()
        status = C.c_int(0)
        self.h = self.L.dab_viterbi_create(device, C.byref(status))
        if not self.h:
            capi.check(status.value)
            raise capi.DabError(status.value, "dab_viterbi_create failed")

    def set_cuda_stream(self, stream_ptr):
        capi.check(self.L.dab_viterbi_set_cuda_stream(self.h, stream_ptr))

    def add_schedule(self, schedule):
        return capi.check(self.L.dab_viterbi_add_schedule(self.h, C.byref(schedule)))

    def decode_batch(self, soft, jobs, out_bytes, raise_on_job_error=True):
        """soft: int8 host array; jobs: array of capi.VIT_JOB_DTYPE. Returns (out bytes, path_error u64, job_status i32)."""
        soft = np.ascontiguousarray(soft, np.int8)
        jobs = np.ascontiguousarray(jobs, capi.VIT_JOB_DTYPE)
        out = np.zeros(out_bytes, np.uint8)
        err = np.zeros(jobs.size, np.uint64)
        st = np.zeros(jobs.size, np.int32)
        rc = self.L.dab_viterbi_decode_batch(self.h, capi.ptr(soft), soft.size, capi.ptr(jobs), jobs.size, capi.ptr(out), out.size,
                                             capi.ptr(err), capi.ptr(st))
        if rc < 0 and (raise_on_job_error or not np.any(st != 0)):
            capi.check(rc)
        return out, err, st

    def decode_jobs_device(self, d_soft, soft_bytes, d_jobs, n_jobs, max_steps, d_out, out_bytes, d_err=None, d_status=None):
        capi.check(self.L.dab_viterbi_decode_jobs_device(self.h, d_soft, soft_bytes, d_jobs, n_jobs, max_steps, d_out, out_bytes, d_err,
                                                         d_status))

    def prepare_jobs(self, jobs):
        """Upload + order + pack a job list that will be decoded repeatedly; returns the plan id."""
        jobs = np.ascontiguousarray(jobs)
        return capi.check(self.L.dab_viterbi_prepare_jobs(self.h, capi.ptr(jobs), jobs.size))

    def decode_prepared(self, plan, d_soft, soft_bytes, d_out, out_bytes, d_err=None, d_status=None):
        capi.check(self.L.dab_viterbi_decode_prepared(self.h, plan, d_soft, soft_bytes, d_out, out_bytes, d_err, d_status))

    def release_jobs(self, plan):
        capi.check(self.L.dab_viterbi_release_jobs(self.h, plan))

    def sync(self):
        capi.check(self.L.dab_viterbi_sync(self.h))

    def kernel_launches(self):
        return int(self.L.dab_viterbi_kernel_launches(self.h))

    def close(self):
        if self.h:
            self.L.dab_viterbi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
