"""VDIF payloads: offset-binary codes, LSB first, in 32-bit words.

Class surface of baseband/vdif/payload.py:117-198; the codec tables
(:137-145) hold CUDA-backed callables (baseband_b200/codecs.py).  EDV 0xab
frames carry Mark 5B payloads (:151-154, :185-194).
"""
from collections import namedtuple

import numpy as np

from .. import codecs
from ..base.payload import PayloadBase

__all__ = ['VDIFPayload']


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
