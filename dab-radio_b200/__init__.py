"""dab-radio_b200: B200-native (sm_100a) implementation of DAB-Radio's receive hot path -- OFDM_Demod (IQ -> int8 soft bits)
and DAB_Viterbi_Decoder (punctured soft bits -> bytes) -- behind the C ABI in include/dab_b200.h.

Layout:  csrc/  hand-written CUDA kernels + the C ABI (libdab_b200.so, built by build.py)
         cpp/   C++ mirror classes with the reference's own signatures (OFDM_Demod, DAB_Viterbi_Decoder)
         *.py   ctypes plumbing used by tests/ and bench.py

The directory name carries a hyphen; import it with importlib.import_module("dab-radio_b200").
"""
from . import capi  # noqa: F401
from .build import build  # noqa: F401
