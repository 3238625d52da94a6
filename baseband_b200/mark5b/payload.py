"""Mark 5B payloads: 10000 bytes of sign/magnitude bit streams, real only
(baseband/mark5b/payload.py:112-145); codecs are CUDA backed."""
from collections import namedtuple

import numpy as np

from .. import codecs
from ..base.payload import PayloadBase

__all__ = ['decode_1bit', 'decode_2bit', 'encode_1bit', 'encode_2bit',
           'Mark5BPayload']

# codec callables under their reference names (mark5b/payload.py:78-106)
decode_1bit, decode_2bit = (codecs.MARK5B_DECODERS[bps] for bps in (1, 2))
encode_1bit, encode_2bit = (codecs.MARK5B_ENCODERS[bps] for bps in (1, 2))


class Mark5BPayload(PayloadBase):
    _nbytes = 10000
    _decoders = codecs.MARK5B_DECODERS
    _encoders = codecs.MARK5B_ENCODERS
    _sample_shape_maker = namedtuple('SampleShape', 'nchan')

    def __init__(self, words, header=None, *, sample_shape=(1,), bps=2,
                 complex_data=False):
        if complex_data:
            raise ValueError('Mark5B format does not support complex data.')
        super().__init__(words, sample_shape=sample_shape, bps=bps,
                         complex_data=False)

    @classmethod
    def fromfile(cls, fh, *args, **kwargs):
        kwargs.pop('header', None)
        raw = fh.read(cls._nbytes)
        if len(raw) < cls._nbytes:
            raise EOFError('could not read full payload.')
        return cls(np.frombuffer(raw, dtype=cls._dtype_word), **kwargs)

    @classmethod
    def fromdata(cls, data, header=None, bps=2):
        data = np.asanyarray(data)
        if data.dtype.kind == 'c':
            raise ValueError('Mark5B format does not support complex data.')
        return super().fromdata(data, header=None, bps=bps)
