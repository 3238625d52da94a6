"""VDIF payloads: offset-binary codes, LSB first, in 32-bit words.

Class surface of baseband/vdif/payload.py:117-198; the codec tables
(:137-145) hold CUDA-backed callables (baseband_b200/codecs.py).  EDV 0xab
frames carry Mark 5B payloads (:151-154, :185-194).
"""
from collections import namedtuple

import numpy as np

from .. import codecs
from ..base.payload import PayloadBase

__all__ = ['init_luts', 'decode_1bit', 'decode_2bit', 'decode_4bit',
           'encode_1bit', 'encode_2bit', 'encode_4bit', 'VDIFPayload']


def init_luts():
    """Byte -> samples look-up tables of the reference (vdif/payload.py:
    25-66), LSB-first offset binary: constants, kept for API parity (the GPU
    kernels build their own shared-memory tables from the same levels)."""
    from ..levels import decoder_levels
    b = np.arange(256)[:, np.newaxis]
    lut1bit = decoder_levels[1][(b >> np.arange(8)) & 1]
    lut2bit = decoder_levels[2][(b >> np.arange(0, 8, 2)) & 3]
    lut4bit = decoder_levels[4][(b >> np.arange(0, 8, 4)) & 0xf]
    return lut1bit, lut2bit, lut4bit


lut1bit, lut2bit, lut4bit = init_luts()
# The operator callables of the codec tables under their reference names
# (vdif/payload.py:69-114): words -> float32, values -> packed bytes, on GPU.
decode_1bit, decode_2bit, decode_4bit = (codecs.VDIF_DECODERS[bps]
                                         for bps in (1, 2, 4))
encode_1bit, encode_2bit, encode_4bit = (codecs.VDIF_ENCODERS[bps]
                                         for bps in (1, 2, 4))


class VDIFPayload(PayloadBase):
    _decoders = codecs.VDIF_DECODERS
    _encoders = codecs.VDIF_ENCODERS
    _sample_shape_maker = namedtuple('SampleShape', 'nchan')

    def __init__(self, words, header=None, sample_shape=(1,), bps=2,
                 complex_data=False):
        if header is not None and header.edv == 0xab:
            self._decoders = codecs.MARK5B_DECODERS
            self._encoders = codecs.MARK5B_ENCODERS
        super().__init__(words, header=header, sample_shape=sample_shape,
                         bps=bps, complex_data=complex_data)
        if self.bps & (self.bps - 1):
            # Samples never straddle 32-bit words (vdif/payload.py:158-169).
            if tuple(self.sample_shape) != (1,):
                raise ValueError('multi-channel VDIF data requires bits per '
                                 'sample that is a power of two.')
            per_word = 32 // self._bpfs
            if per_word & (per_word - 1):
                raise ValueError(
                    'cannot yet sensibly handle {} data with bps={}'.format(
                        'complex' if self.complex_data else 'real', bps))
            self._bpfs = 32 // per_word

    @classmethod
    def fromdata(cls, data, header=None, bps=2, edv=None):
        if (edv if header is None else header.edv) == 0xab:
            from ..mark5b.payload import Mark5BPayload
            data = np.asanyarray(data)
            bps = bps if header is None else header.bps
            inner = Mark5BPayload.fromdata(data, bps=bps)
            return cls(inner.words, header, sample_shape=data.shape[1:],
                       bps=bps, complex_data=False)
        return super().fromdata(data, header=header, bps=bps)
