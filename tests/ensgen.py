"""Synthetic DAB ensemble generator for the ensemble-decoder tests and the bench (test infrastructure).

Transmit side of what FIC_Decoder / MSC_Decoder undo (ETSI EN 300 401 clauses 5.2, 10, 11, 12): FIBs with CRC16 -> energy
dispersal -> mother code -> puncturing (PI_16 x21 + PI_15 x3 + PI_X for the FIC, the sub-channel's EEP / UEP profile for the MSC)
-> time interleaving over 16 CIFs -> soft bits (+-127) with additive Gaussian noise.  Frame layout as BasicRadio::Process splits
it (reference src/basic_radio/basic_radio.cpp:49-50): [FIC bits | MSC bits], MSC = nb_cifs CIFs of 864 CU x 64 bits.
"""
import numpy as np

from oracle import pyoracle as po

CIF_OFFSETS = np.array([0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15])
FIC_SEGS = None

# (nb_cifs, nb_fic_bits, nb_fib_cif_bits, nb_fibs_per_cif, nb_cif_bits) per transmission mode (dab_parameters.h:26-93)
MODE_GEOM = {1: (4, 9216, 2304, 3, 55296), 2: (1, 2304, 2304, 3, 55296), 3: (1, 3072, 3072, 4, 55296), 4: (2, 4608, 2304, 3, 55296)}


def fic_segments():
    global FIC_SEGS
    if FIC_SEGS is None:
        FIC_SEGS = [(po.puncture_code(16), 128 * 21), (po.puncture_code(15), 128 * 3), (po.PI_X, 24)]
    return FIC_SEGS


def make_fib_group(rng, corrupt=()):
    """96 bytes = 3 FIBs of 30 data bytes + CRC16; FIB indices in `corrupt` get a flipped data bit (CRC then fails)."""
    data = rng.integers(0, 256, 96, dtype=np.uint8)
    for i in range(3):
        crc = po.crc16_fib(data[32 * i:32 * i + 30])
        data[32 * i + 30], data[32 * i + 31] = crc >> 8, crc & 0xFF
    for i in corrupt:
        data[32 * i + 7] ^= 0x10
    return data


def encode_fic_group(fib_bytes):
    tx = fib_bytes ^ po.scrambler_bytes(fib_bytes.size)
    return po.puncture(po.conv_encode(tx), fic_segments())


def sub_decoded_bytes(sc):
    """bytes MSC_Decoder produces per CIF for the sub-channel (msc_decoder.cpp:103-108), ignoring the reference's underrun quirk"""
    segs = po.msc_segments(sc)
    return (sum(n for _, n in segs) // 4 - 6) // 8


def encode_subchannel(sc, payload):
    """payload bytes -> length*64 soft bits (zero padded: UEP padding / unused tail)"""
    segs = po.msc_segments(sc)
    tx = payload ^ po.scrambler_bytes(payload.size)
    soft = po.puncture(po.conv_encode(tx), segs)
    out = np.zeros(sc.length * 64, np.int8)
    n = min(out.size, soft.size)
    out[:n] = soft[:n]
    return out


class EnsembleTx:
    """Generates consecutive frames of soft bits for one stream, time-interleaved across CIFs."""

    def __init__(self, mode, subs, seed=0, sigma=0.0):
        self.mode = mode
        self.nb_cifs, self.nb_fic_bits, self.nb_fib_cif_bits, self.nb_fibs, self.nb_cif_bits = MODE_GEOM[mode]
        self.subs = list(subs)
        self.rng = np.random.default_rng(seed)
        self.sigma = sigma
        self.pending = np.zeros((16, self.nb_cif_bits), np.int8)   # pending[d] = what is known of the CIF d steps ahead
        self.sent_fibs = []      # per frame: [nb_cifs][96]
        self.sent_payloads = []  # per logical CIF: list of payload arrays per sub-channel

    def _next_cif(self):
        logical = np.zeros(self.nb_cif_bits, np.int8)
        payloads = []
        for sc in self.subs:
            end = (sc.start_address + sc.length) * 64
            n = sub_decoded_bytes(sc)
            payload = self.rng.integers(0, 256, max(n, 0), dtype=np.uint8)
            payloads.append(payload)
            if end <= self.nb_cif_bits and n > 0:
                logical[sc.start_address * 64:end] = encode_subchannel(sc, payload)
        self.sent_payloads.append(payloads)
        # time interleaver: bit i of logical frame m goes out in CIF m + offset[i % 16] (inverse of cif_deinterleaver.cpp:49-66)
        idx = np.arange(self.nb_cif_bits)
        self.pending[CIF_OFFSETS[idx % 16], idx] = logical
        out = self.pending[0].copy()
        self.pending = np.roll(self.pending, -1, axis=0)
        self.pending[15] = 0
        return out

    def next_frame(self, corrupt_fibs=()):
        fic = np.zeros(self.nb_fic_bits, np.int8)
        fibs = []
        if self.nb_fib_cif_bits == 2304:
            for c in range(self.nb_cifs):
                g = make_fib_group(self.rng, corrupt=[i for (cc, i) in corrupt_fibs if cc == c])
                fibs.append(g)
                fic[c * 2304:(c + 1) * 2304] = encode_fic_group(g)
        else:
            fic[:] = np.where(self.rng.integers(0, 2, fic.size), 127, -127)
        self.sent_fibs.append(fibs)
        msc = np.concatenate([self._next_cif() for _ in range(self.nb_cifs)])
        frame = np.concatenate([fic, msc]).astype(np.float64)
        if self.sigma:
            frame = frame + self.sigma * self.rng.standard_normal(frame.size)
        return np.clip(np.rint(frame), -127, 127).astype(np.int8)


def oracle_decode_stream(mode, subs, frames):
    """Run the oracle's FIC_Decoder / MSC_Decoder restatement over consecutive frames of one stream.
    -> per frame: (fib_bytes [nb_cifs][96], fib_valid [nb_cifs][fibs], fic_err [nb_cifs], msc: [nb_cifs][n_subs] of (bytes, err))"""
    nb_cifs, nb_fic_bits, nb_fib_cif_bits, nb_fibs, nb_cif_bits = MODE_GEOM[mode]
    decs = [po.OracleMscDecoder(sc) for sc in subs]
    out = []
    for frame in frames:
        fb, fv, fe = [], [], []
        for c in range(nb_cifs):
            b, v, e = po.fic_decode_group(frame[c * nb_fib_cif_bits:(c + 1) * nb_fib_cif_bits], nb_fibs)
            fb.append(b); fv.append(v); fe.append(e)
        msc = []
        for c in range(nb_cifs):
            cif = frame[nb_fic_bits + c * nb_cif_bits: nb_fic_bits + (c + 1) * nb_cif_bits]
            msc.append([d.decode_cif(cif) for d in decs])
        out.append((fb, fv, fe, msc))
    return out
